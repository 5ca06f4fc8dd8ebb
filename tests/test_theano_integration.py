"""INTEGRATION.md section 3 executed: the Op classes of g3py_b200/theano_ops.py spliced into the graph of the UNMODIFIED
reference (its own hyper-parameter variables, mean, warping, shared data), compiled with `theano.function` and
differentiated with `tt.grad` — all through the Theano/PyMC3 stand-in of tests/golden/refshim.  The fused GPLogpOp must
reproduce the reference's own logp and dlogp, with the gradient reaching the mean / warping hypers through `delta` and
`det_m` (plain Theano expressions) and the kernel hypers through GPLogpOp.grad -> GPLogpGradOp.

Needs the reference tree (/root/reference, present in the build container only) and no GPU: the device is the NumPy
double of tests/fake_ctx.py.  Skipped where the reference is absent."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/g3py"), reason="reference tree not available")


@pytest.fixture(scope="module")
def ref():
    import torch
    prev = torch.get_default_dtype()
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import refshim
    import make_reference_goldens as mr
    g3ref = mr._import_reference()
    yield g3ref
    torch.set_default_dtype(prev)
    refshim.uninstall()
    for name in [n for n in sys.modules if n == "g3py" or n.startswith("g3py.")]:
        del sys.modules[name]
    sys.path[:] = [p for p in sys.path if p != "/root/reference"]


@pytest.fixture
def fake(monkeypatch):
    import g3py_b200 as g3
    from fake_ctx import FakeContext
    ctx = FakeContext()
    monkeypatch.setattr(g3.processes, "get_context", lambda device=0: ctx)
    return ctx


CASES = {
    "gp_se_mat52": ("GP", "Bias", lambda g, X: g.SE(X) + g.MAT52(X), None),
    "wgp_boxcox_rq": ("WGP", "Linear", lambda g, X: g.RQ(X) * g.SE(X, name="SE2"), "BoxCoxShifted"),
    "tp_ou": ("TP", "Bias", lambda g, X: g.OU(X), None),
}


@pytest.mark.parametrize("case", list(CASES))
def test_fused_op_inside_the_reference_graph(ref, fake, case):
    import theano as th
    import theano.tensor as tt
    import pymc3 as pm
    import g3py_b200 as g3
    from g3py_b200 import _cabi as cabi, theano_ops
    from helpers import scaled_err

    cls, mean, kern, mapping = CASES[case]
    rng = np.random.default_rng(3)
    N, D = 40, 2
    X = rng.uniform(0.2, 4.0, size=(N, D))
    y = np.sin(X[:, 0]) + 0.3 * X[:, 1] + 0.1 * rng.standard_normal(N)
    if mapping:
        y = np.exp(0.4 * y) + 0.3
    # ---- the reference process (its own graph objects) and the product-side process (descriptor + slot order)
    rargs = [X, getattr(ref, mean)(X), kern(ref, X)] + ([getattr(ref, mapping)()] if mapping else [])
    rp = getattr(ref, cls)(*rargs)
    rp.observed(X, y)
    pargs = [X, getattr(g3, mean)(X), kern(g3, X)] + ([getattr(g3, mapping)()] if mapping else [])
    pp = getattr(g3, cls)(*pargs)
    pp.observed(X, y)
    names = [m.var for m in rp.active.bijection.ordering.vmap]
    assert [h.tname for h in pp.registry.vars] == names
    th0 = pp.dict_to_array(pp.params_default) + 0.05 * rng.standard_normal(pp.ndim)
    if cls == "TP":
        th0[-1] = np.log(4.0)

    # ---- delta and det_m stay Theano expressions, the kernel part is one fused Op
    ops = theano_ops.build_ops()
    model = rp.model
    theta_k = tt.concatenate([tt.flatten(model[h.name]) if h is not None else tt.as_tensor_variable(np.asarray(c, dtype=np.float64))
                              for h, off, size, c in pp._slots])          # natural space, descriptor slot order
    delta = rp.f_mapping.inv(rp.th_outputs) - rp.prior_location_inputs
    det_m = rp.f_mapping.logdet_dinv(rp.th_outputs)
    n = rp.th_outputs.shape[0].astype("float64")
    if cls == "TP":
        nu = rp.th_freedom(prior=True)
        core, beta, logdet = ops.GPLogpOp(pp.desc, cabi.KIND_STUDENT)(rp.th_inputs, delta, theta_k, nu)
        np5, np2, npi = np.float32(0.5), np.float32(2.0), np.float32(np.pi)
        r2 = tt.gammaln((nu + n) * np5) - tt.gammaln(nu * np5) - np5 * n * tt.log((nu - np2) * npi)   # studentT.py:128
        loglike = core + r2 + det_m
    else:
        core, beta, logdet = ops.GPLogpOp(pp.desc, cabi.KIND_GAUSS)(rp.th_inputs, delta, theta_k)
        loglike = np.float32(-0.5) * n * tt.log(np.float32(2.0 * np.pi)) + core + det_m             # gaussian.py:218
    prior = tt.add(*[tt.sum(v.logpt) for v in model.free_RVs])
    logp = loglike + prior
    free = list(model.vars)
    fn = th.function(free, [logp] + [tt.grad(logp, v) for v in free])
    vals = fn(**rp.active.array_to_dict(th0))
    got_lp = float(vals[0])
    got_g = np.concatenate([np.atleast_1d(np.asarray(v, dtype=np.float64)).ravel() for v in vals[1:]])

    # ---- against the reference's own compiled logp / dlogp (per variable name: its order is traversal order)
    want_lp = float(rp.logp(th0, array=True))
    wrt = pm.inputvars(pm.cont_inputs(rp.th_logp()))
    flat = np.asarray(rp.dlogp(th0, array=True), dtype=np.float64)
    off, by_name = 0, {}
    for v in wrt:
        size = int(np.prod(np.shape(v.tag.test_value)))
        by_name[v.name] = flat[off:off + size]
        off += size
    want_g = np.concatenate([by_name.get(v.name, np.zeros(int(np.prod(np.shape(v.tag.test_value))))) for v in free])
    assert abs(got_lp - want_lp) <= 1e-10 * abs(want_lp)
    # Matern rates: the reference's autodiff loses them to NaN -> 0, the fused Op returns the analytic value
    quirk = np.concatenate([np.full(int(np.prod(np.shape(v.tag.test_value))), ("MAT32_rate" in v.name) or ("MAT52_rate" in v.name))
                            for v in free])
    assert scaled_err(got_g[~quirk], want_g[~quirk]) < 1e-9
    if quirk.any():
        assert np.all(want_g[quirk] == 0.0) and np.all(got_g[quirk] != 0.0)
    # and the same numbers as the product's own front end
    assert abs(pp.logp(th0, array=True) - got_lp) <= 1e-10 * abs(got_lp)
    assert scaled_err(pp.dlogp(th0, array=True), got_g) < 1e-9


def test_gram_and_posterior_ops_inside_the_reference_graph(ref, fake):
    """GramOp (+ its gradient through GramVJPOp) against the reference's own `Kernel.cov` expression and its Theano
    gradient, and GPPosteriorOp against the reference's posterior selectors, inside the same graph."""
    import theano as th
    import theano.tensor as tt
    import g3py_b200 as g3
    from g3py_b200 import theano_ops
    from helpers import scaled_err

    rng = np.random.default_rng(5)
    N, M, D = 30, 9, 2
    X = rng.uniform(0.2, 4.0, size=(N, D))
    Xs = rng.uniform(0.2, 4.0, size=(M, D))
    y = np.sin(X[:, 0]) + 0.3 * X[:, 1] + 0.1 * rng.standard_normal(N)
    rp = ref.GP(X, ref.Bias(X), ref.SE(X) + ref.RQ(X))
    rp.observed(X, y)
    pp = g3.GP(X, g3.Bias(X), g3.SE(X) + g3.RQ(X))
    pp.observed(X, y)
    th0 = pp.dict_to_array(pp.params_default) + 0.05 * rng.standard_normal(pp.ndim)
    params = rp.active.array_to_dict(th0)
    ops = theano_ops.build_ops()
    model = rp.model
    free = list(model.vars)

    def slots(sl):
        return tt.concatenate([tt.flatten(model[h.name]) if h is not None else tt.as_tensor_variable(np.asarray(c, dtype=np.float64))
                               for h, off, size, c in sl])
    # ---- Gram: f_kernel.cov(space, inputs) and d/dtheta of a weighted sum of its entries
    W = rng.standard_normal((M, N))
    xs_var, x_var = th.shared(Xs, name="xs"), th.shared(X, name="x")
    K_op = ops.GramOp(pp.desc_f)(xs_var, x_var, slots(pp._slots_f))
    K_ref = rp.f_kernel.cov(xs_var, x_var)
    outs = []
    for Kx in (K_op, K_ref):
        s = tt.sum(Kx * W)
        outs.append(th.function(free, [Kx] + [tt.grad(s, v) for v in free])(**params))
    assert scaled_err(outs[0][0], outs[1][0]) < 1e-13
    for a, b in zip(outs[0][1:], outs[1][1:]):
        assert scaled_err(np.atleast_1d(a), np.atleast_1d(b)) < 1e-11 or (np.max(np.abs(b)) == 0 and np.max(np.abs(a)) == 0)
    # ---- posterior: location - m(X*) and variance from one Op against the reference's selectors
    delta = rp.f_mapping.inv(rp.th_outputs) - rp.prior_location_inputs
    for noise in (False, True):
        mean, var = ops.GPPosteriorOp(pp.desc, noise)(rp.th_inputs, xs_var, delta, slots(pp._slots))
        got_m, got_v = th.function(free, [mean, var])(**params)
        kw = dict(params=params, space=Xs, inputs=X, outputs=y, prior=False, noise=noise)
        want_loc = np.asarray(rp.location(**kw)) - np.asarray(rp.location(params=params, space=Xs, inputs=X, outputs=y, prior=True))
        assert scaled_err(got_m, want_loc) < 1e-9
        assert scaled_err(got_v, rp.kernel_diag(**kw)) < 1e-8


def test_cholesky_robust_gpu_replaces_the_bare_factor(ref, fake, monkeypatch):
    """CholeskyRobustGPU in place of `cholesky_robust` for the non-differentiated
    selectors (`th_cholesky`: the factors the samplers use).  Patched into the reference's elliptical module before the
    process is built, it must return the factor the reference's own Op returns - also through the jitter ladder."""
    import g3py_b200  # noqa: F401
    from g3py_b200 import theano_ops
    from helpers import scaled_err
    import g3py.processes.elliptical as ell

    rng = np.random.default_rng(7)
    X = rng.uniform(0.2, 4.0, size=(25, 1))
    X[1::2] = X[0::2][:12]                               # duplicated inputs: the noise-free prior Gram is singular
    y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(25)
    Xs = np.vstack([X[:6], X[:6]])                       # singular on `space` too -> ladder inside the selector
    want = {}
    for patched in (False, True):
        if patched:
            monkeypatch.setattr(ell, "cholesky_robust", theano_ops.build_ops().CholeskyRobustGPU())
        rp = ref.GP(X, ref.Bias(X), ref.SE(X))
        rp.observed(X, y)
        for noise in (True, False):
            L = np.asarray(rp.cholesky(space=Xs, prior=True, noise=noise))
            if not patched:
                want[noise] = L
            else:
                assert scaled_err(L, want[noise]) < 1e-12
                assert np.allclose(L @ L.T, np.asarray(rp.kernel(space=Xs, prior=True, noise=noise)), atol=1e-4)


def test_cholesky_robust_gpu_gradient_through_the_bare_factor(ref, fake, monkeypatch):
    """`tt.grad` THROUGH CholeskyRobustGPU (its symbolic Murray reverse mode) when it replaces `cholesky_robust` in the
    reference's own logp graph: the reference's compiled dlogp must not change."""
    import g3py_b200  # noqa: F401
    from g3py_b200 import theano_ops
    from helpers import scaled_err
    import g3py.processes.elliptical as ell

    rng = np.random.default_rng(9)
    X = rng.uniform(0.2, 4.0, size=(30, 2))
    y = np.exp(0.4 * (np.sin(X[:, 0]) + 0.3 * X[:, 1] + 0.1 * rng.standard_normal(30))) + 0.3
    got = {}
    for patched in (False, True):
        if patched:
            monkeypatch.setattr(ell, "cholesky_robust", theano_ops.build_ops().CholeskyRobustGPU())
        rp = ref.WTP(X, ref.Bias(X), ref.SE(X) + ref.RQ(X), ref.BoxCoxShifted())
        rp.observed(X, y)
        th0 = rp.active.dict_to_array(rp.params_default) + 0.05 * np.random.default_rng(1).standard_normal(rp.ndim)
        got[patched] = (float(rp.logp(th0, array=True)), np.asarray(rp.dlogp(th0, array=True), dtype=np.float64))
    assert abs(got[True][0] - got[False][0]) <= 1e-12 * abs(got[False][0])
    assert np.max(np.abs(got[False][1])) > 0
    assert scaled_err(got[True][1], got[False][1]) < 1e-10
