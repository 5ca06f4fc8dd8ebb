"""Parity against the REFERENCE ITSELF: tests/golden/reference_g3py.json holds outputs of the unmodified
/root/reference/g3py source (logp, dlogp, posterior moments, Gram matrices ...) executed in the build container
through the Theano/PyMC3 API stand-in of tests/golden/refshim (generator: tests/golden/make_reference_goldens.py).

CPU tests pin the oracle to those outputs; GPU tests hold the CUDA path (through the public API, i.e. ctypes ->
libg3b.so) to them at the north-star tolerance of 1e-9 relative.  Nothing here reads /root/reference.
"""
import json
import os

import numpy as np
import pytest

from oracle import g3_oracle as orc
from helpers import build_process, scaled_err

HERE = os.path.dirname(os.path.abspath(__file__))
REF = json.load(open(os.path.join(HERE, "golden", "reference_g3py.json")))
CASES = list(REF)
POST = [c for c in CASES if "post_noise0" in REF[c]]
TRANSPORT = [c for c in CASES if REF[c]["spec"].get("kind") == "transport"]
# Cholesky-route posterior (the CUDA path) is comparable with the reference's LU route only where K is positive
# definite; on an indefinite K (SIN with a large rate) the reference itself is inconsistent: logp sees
# the jittered factor, predict the raw LU solve.
POST_PD = [c for c in POST if REF[c]["min_eig_K"] > 1e-6]

TOL = 1e-9          # north star: 1e-9 relative on fp64 logp, gradients and posterior moments


def _load(name):
    rec = REF[name]
    return rec, np.array(rec["X"]), np.array(rec["y"]), np.array(rec["Xs"]), np.array(rec["theta"])


def _ref_dlogp(rec):
    """Reference gradient re-ordered into bijection (= theta) order; variables the reference graph does not reach
    are absent from its dlogp and count as 0."""
    parts = []
    for nm, lay in zip(rec["ref_names"], rec["layout"]):
        parts.append(np.asarray(rec["dlogp"].get(nm, [0.0] * lay[1]), dtype=np.float64))
    return np.concatenate(parts) if parts else np.zeros(0)


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


# ------------------------------------------------------------------------------------ oracle (CPU)
@pytest.mark.parametrize("name", CASES)
def test_oracle_layout_and_logp_match_reference(name):
    rec, X, y, Xs, th = _load(name)
    op = orc.build_process(rec["spec"], X.shape[1])
    assert [list(l) for l in op.layout()] == rec["layout"]
    assert _rel(op.logp(th, X, y), rec["logp"]) < 1e-12
    assert _rel(op.loglike(th, X, y), rec["loglike"]) < 1e-12
    assert op.logprior(th) == rec["logp_prior"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_dlogp_matches_reference(name):
    rec, X, y, Xs, th = _load(name)
    if rec.get("skip_dlogp"):
        pytest.skip("gradient through the tt_to_cov shift: Theano's tie-multiplied max gradient is not reproduced")
    op = orc.build_process(rec["spec"], X.shape[1])
    want = _ref_dlogp(rec)
    # the reference's autodiff yields NaN -> 0 for the rate of sqrt-kernels: nan_quirk mode
    for method in ("analytic", "murray"):
        kw = {} if name in TRANSPORT else {"method": method}
        got = op.dlogp(th, X, y, nan_quirk=True, **kw)
        assert scaled_err(got, want) < (1e-10 if name != "jitter_ladder" else 1e-6), method


SELECTORS = [(sel, prior, noise) for sel in ("transport", "transport_inv", "transport_diag") for prior in (False, True)
             for noise in (False, True)]


def _transport_input(rec, sel, prior, noise):
    """The vector each golden selector call was given (the prior inverse gets the transported draw)."""
    if sel == "transport_inv" and prior:
        return np.array(rec["transport_prior1_noise%d" % noise])
    return np.array(rec["vector"])


@pytest.mark.parametrize("name", TRANSPORT)
def test_oracle_transport_selectors_match_reference(name):
    rec, X, y, Xs, th = _load(name)
    op = orc.build_process(rec["spec"], X.shape[1])
    for sel, prior, noise in SELECTORS:
        got = getattr(op, sel)(th, Xs, _transport_input(rec, sel, prior, noise), X, y, prior=prior, noise=noise)
        assert scaled_err(got, rec["%s_prior%d_noise%d" % (sel, prior, noise)]) < 1e-9, (sel, prior, noise)
    # the prior inverse undoes the prior transport
    back = op.transport_inv(th, Xs, np.array(rec["transport_prior1_noise1"]), prior=True, noise=True)
    assert scaled_err(back, rec["vector"]) < 1e-8


@pytest.mark.parametrize("name", POST)
@pytest.mark.parametrize("noise", [False, True])
def test_oracle_posterior_matches_reference(name, noise):
    rec, X, y, Xs, th = _load(name)
    op = orc.OracleProcess(rec["spec"], X.shape[1])
    r = rec["post_noise%d" % noise]
    po = op.posterior(th, Xs, X, y, noise=noise, cov=True, solver="lu")
    assert scaled_err(po["location"], r["location"]) < 1e-10
    assert scaled_err(po["kernel"], r["kernel"]) < 1e-10
    assert scaled_err(po["kernel_diag"], r["kernel_diag"]) < 1e-9
    pr = op.predict(th, Xs, X, y, noise=noise)
    for key in ("mean", "median", "variance", "std", "quantile_up", "quantile_down"):
        assert scaled_err(pr[key], r[key]) < 1e-9, key
    # the Cholesky route the CUDA path takes agrees with the reference's LU route on these inputs
    if name not in POST_PD:
        return
    pc = op.posterior(th, Xs, X, y, noise=noise, solver="chol")
    assert scaled_err(pc["location"], r["location"]) < 1e-9
    assert scaled_err(pc["kernel_diag"], r["kernel_diag"]) < 1e-8


@pytest.mark.parametrize("name", POST)
def test_oracle_gram_matches_reference(name):
    """prior kernel on `space` = f_kernel.cov(Xs) (noise=False) and tt_to_cov(f_kernel_noise.cov(Xs)) (noise=True)."""
    rec, X, y, Xs, th = _load(name)
    op = orc.OracleProcess(rec["spec"], X.shape[1])
    nat = op.natural(th)
    t_ker = op.split(nat)[1]
    Kf = op.f_kernel.cov(t_ker[:op.f_kernel.n_theta()], Xs, Xs, True)
    assert scaled_err(Kf, rec["post_noise0"]["prior_kernel"]) < 1e-13
    Kn = orc.tt_to_cov(op.k_noise.cov(t_ker, Xs, Xs, True), op.consts)
    assert scaled_err(Kn, rec["post_noise1"]["prior_kernel"]) < 1e-13


def test_standin_reproduces_real_theano_notebook_output():
    """tests/golden/reference_shim_check.json: the reference executed through the Theano/PyMC3 stand-in against the
    values a real Theano run printed in the reference's notebook (float32 prints), and the oracle against both."""
    chk = json.load(open(os.path.join(HERE, "golden", "reference_shim_check.json")))
    spec = {"kind": "student", "warped": True, "location": {"type": "Bias"}, "kernel": {"type": "SE"},
            "mapping": {"type": "ArcsinhLinear"}}
    op = orc.OracleProcess(spec, 1)
    X2, y2 = np.array([[0.0], [1.0]]), np.array([0.0, 1.0])
    for row in chk["rows"]:
        assert abs(row["shim_logp"] - row["notebook_sum"]) < 5e-6
        assert abs(op.logp(np.array(row["theta"]), X2, y2) - row["shim_logp"]) < 1e-12


def test_reference_confirms_matern_rate_gradient_quirk():
    """Executed evidence for SURVEY a3-iv: the reference's dlogp has exact zeros for Matern rates."""
    rec = REF["C2_gp_se_mat52"]
    assert all(v == 0.0 for v in rec["dlogp"]["GP_MAT52_rate_log__"])
    assert all(v != 0.0 for v in rec["dlogp"]["GP_SE_rate_log__"])
    assert all(v == 0.0 for v in REF["leaf_mat32"]["dlogp"]["GP_MAT32_rate_log__"])


def test_reference_default_params_are_float32_rounded():
    """params_default goes through get_hypers_floatX = np.float32 (hypers/__init__.py:12-16) even in an fp64 run."""
    rec = REF["C1_gp_se"]
    for k, v in rec["default_params"].items():
        assert np.all(np.asarray(v) == np.asarray(v, dtype=np.float32).astype(np.float64)), k


# ------------------------------------------------------------------- product: host logic (CPU) and CUDA path (GPU)
# The same checks run twice: on CPU with the NumPy test double of the device context (tests/fake_ctx.py: pins the
# host-side assembly -- bijection, means, warpings, Student-t terms, selectors -- to the reference), and on the GPU
# through ctypes -> libg3b.so (pins the CUDA kernels).
def _check_logp_dlogp(name):
    rec, X, y, Xs, th = _load(name)
    gp = build_process(rec["spec"], X)
    gp.observed(X, y)
    assert [h.tname for h in gp.registry.vars] == rec["ref_names"]          # same names, same order as the reference
    assert _rel(gp.logp(th, array=True), rec["logp"]) < TOL
    assert _rel(gp.loglike(th, array=True), rec["loglike"]) < TOL
    assert gp.logp(th, array=True, prior=True) == rec["logp_prior"]
    lp, g, info = gp.logp_dlogp_batch(np.stack([th, th]))
    assert _rel(lp[0], rec["logp"]) < TOL and lp[0] == lp[1]
    if rec.get("skip_dlogp"):
        return
    want = _ref_dlogp(rec)
    tol = TOL if name != "jitter_ladder" else 1e-5         # ladder case: K is singular up to the 1e-6 jitter
    gq = gp.dlogp(th, array=True, reference_nan_quirk=True)
    assert scaled_err(gq, want) < tol
    # default mode: identical except where the reference loses the component to NaN -> 0 (analytic value instead)
    quirk = np.zeros(len(th), bool)
    for h in gp.f_kernel_noise.nan_quirk_hypers():
        quirk[h.offset:h.offset + h.size] = True
    assert scaled_err(g[0][~quirk], want[~quirk]) < tol
    if quirk.any():
        assert np.all(g[0][quirk] != 0.0) and np.all(gq[quirk] == 0.0) and np.all(want[quirk] == 0.0)
    # dict-in API with the reference's own variable names
    params = {n: np.array(v) for n, v in zip(rec["ref_names"], np.split(th, np.cumsum([l[1] for l in rec["layout"]])[:-1]))}
    assert _rel(gp.logp(params), rec["logp"]) < TOL


def _check_transport(name):
    rec, X, y, Xs, th = _load(name)
    gp = build_process(rec["spec"], X)
    gp.observed(X, y)
    for sel, prior, noise in SELECTORS:
        got = getattr(gp, sel)(th, space=Xs, vector=_transport_input(rec, sel, prior, noise), prior=prior, noise=noise,
                               array=True)
        # joint (N+M) Cholesky: the trailing block is the factor of a Schur complement, conditioning ~1e3..1e6
        assert scaled_err(got, rec["%s_prior%d_noise%d" % (sel, prior, noise)]) < 1e-7, (sel, prior, noise)
    s = gp.sampler(th, space=Xs, samples=4, array=True, rng=np.random.default_rng(0))
    assert s.shape == (len(Xs), 4) and np.all(np.isfinite(s))
    out = gp.predict(th, space=Xs, array=True, quantiles=True, simulations=16, rng=np.random.default_rng(1))
    assert np.all(out["quantile_down"] <= out["quantile_up"]) and out["mean"].shape == (len(Xs),)


def _check_posterior(name, noise):
    rec, X, y, Xs, th = _load(name)
    gp = build_process(rec["spec"], X)
    gp.observed(X, y)
    r = rec["post_noise%d" % noise]
    kw = dict(space=Xs, array=True, noise=noise)
    assert scaled_err(gp.location(th, **kw), r["location"]) < TOL
    assert scaled_err(gp.kernel_diag(th, **kw), r["kernel_diag"]) < 1e-8     # k** - |V|^2 cancellation, LU vs Cholesky
    assert scaled_err(gp.kernel_sd(th, **kw), r["kernel_sd"]) < 1e-8
    assert scaled_err(gp.kernel(th, **kw), r["kernel"]) < 1e-8
    out = gp.predict(th, mean=True, var=True, std=True, median=True, quantiles=True,
                     cov=("covariance" in r), **kw)
    for key in ("mean", "median", "quantile_up", "quantile_down"):
        assert scaled_err(out[key], r[key]) < 1e-8, key
    for key in ("variance", "std"):
        assert scaled_err(out[key], r[key]) < 1e-7, key
    if "covariance" in r:
        assert scaled_err(out["covariance"], r["covariance"]) < 1e-8
    assert scaled_err(gp.quantiler(th, q=0.025, **kw), r["quantile_down"]) < 1e-8
    assert scaled_err(gp.mean(th, **kw), r["mean"]) < 1e-8
    if "freedom_post" in rec:
        assert gp.freedom(th, array=True) == pytest.approx(rec["freedom_post"], rel=1e-14)


def _check_gram(name):
    """Gram matrices straight from the gram kernel (g3_gram) against the reference's Kernel.cov."""
    rec, X, y, Xs, th = _load(name)
    gp = build_process(rec["spec"], X)
    gp.observed(X, y)
    Kf = gp.kernel(th, space=Xs, array=True, prior=True, noise=False)
    assert scaled_err(Kf, rec["post_noise0"]["prior_kernel"]) < 1e-12
    Kn = gp.kernel(th, space=Xs, array=True, prior=True, noise=True)
    assert scaled_err(Kn, rec["post_noise1"]["prior_kernel"]) < 1e-12


def _check_logpredictive(name):
    rec, X, y, Xs, th = _load(name)
    gp = build_process(rec["spec"], X)
    gp.observed(X, y)
    out = gp.predict(th, space=Xs, array=True, distribution=True, noise=True)
    got = out["logpredictive"](np.array(rec["logpredictive_at"]))
    assert _rel(got, rec["logpredictive"]) < 1e-8


LOGPRED = [c for c in POST_PD if "logpredictive" in REF[c]]
C1MAP = json.load(open(os.path.join(HERE, "golden", "reference_c1_find_map.json")))


def _check_c1_find_map():
    """BASELINE config 1 at full size (N=200 tutorial example): the reference's default hypers, logp / dlogp there, and
    the optimum its own find_MAP (scipy BFGS through its logp / dlogp) reached."""
    import g3py_b200 as g3
    from g3py_b200 import workloads
    x, y = workloads.c1_inputs()
    gp = g3.GP(x, g3.Bias(), g3.SE(x))
    gp.observed(x, y)
    assert [h.tname for h in gp.registry.vars] == C1MAP["names"]
    th0, thm = np.array(C1MAP["theta_default"]), np.array(C1MAP["theta_map"])
    # params_default of the reference is float32-rounded (hypers/__init__.py:12-16); ours is the unrounded value
    assert np.allclose(gp.dict_to_array(gp.params_default), th0, rtol=3e-7, atol=1e-7)
    for th, lp, g in ((th0, C1MAP["logp_default"], C1MAP["dlogp_default"]), (thm, C1MAP["logp_map"], C1MAP["dlogp_map"])):
        assert _rel(gp.logp(th, array=True), lp) < TOL
        assert np.max(np.abs(gp.dlogp(th, array=True) - np.array(g))) < TOL * max(np.max(np.abs(C1MAP["dlogp_default"])), 1.0)
    best = gp.dict_to_array(gp.find_MAP(start=gp.array_to_dict(th0)))
    assert gp.logp(best, array=True) >= C1MAP["logp_map"] - 1e-6 * abs(C1MAP["logp_map"])   # same optimum (or better)
    assert np.max(np.abs(best - thm)) < 1e-3


FULL = json.load(open(os.path.join(HERE, "golden", "reference_fullsize.json")))
C2_SPEC = {"kind": "gauss", "location": {"type": "Bias"}, "kernel": {"type": "sum", "k1": {"type": "SE"}, "k2": {"type": "MAT52"}}}
C3_SPEC = {"kind": "gauss", "warped": True, "location": {"type": "Bias"},
           "kernel": {"type": "prod", "k1": {"type": "SIN"}, "k2": {"type": "SE"}}, "mapping": {"type": "BoxCoxShifted"}}


def _by_name(d, names, sizes):
    return np.concatenate([np.asarray(d.get(n, [0.0] * s), dtype=np.float64) for n, s in zip(names, sizes)])


def _check_fullsize_c2():
    """BASELINE config 2 at FULL size (N=4096, D=3): two rows of the benchmark's own theta batch against the
    reference executed at that size (logp 1e-9; gradient 1e-9 except the Matern rates the reference zeroes)."""
    import g3py_b200 as g3
    from g3py_b200 import workloads
    X, y, Theta = workloads.c2_inputs(4096, 64)
    gp = g3.GP(X, g3.Bias(), g3.SE(X) + g3.MAT52(X))
    gp.observed(X, y)
    rec = FULL["C2"]
    assert [h.tname for h in gp.registry.vars] == rec["names"]
    sizes = [h.size for h in gp.registry.vars]
    rows = rec["rows"]
    Th = np.array([r["theta"] for r in rows])
    assert np.array_equal(Th, Theta[[r["b"] for r in rows]])
    lp, g, info = gp.logp_dlogp_batch(Th, reference_nan_quirk=True)
    for i, r in enumerate(rows):
        assert _rel(lp[i], r["logp"]) < TOL
        assert scaled_err(g[i], _by_name(r["dlogp"], rec["names"], sizes)) < TOL


def _check_fullsize_c3():
    """BASELINE config 3 at FULL size (N=2048 warped GP, periodic x SE): logp, gradient and posterior moments at 64 of
    the 10 000 test points against the reference executed at that size."""
    import g3py_b200 as g3
    from g3py_b200 import workloads
    x, y, xs = workloads.c3_inputs(2048, 10000)
    gp = g3.WGP(x, g3.Bias(), g3.SIN(x) * g3.SE(x), g3.BoxCoxShifted())
    gp.observed(x, y)
    rec = FULL["C3"]
    assert [h.tname for h in gp.registry.vars] == rec["names"]
    sizes = [h.size for h in gp.registry.vars]
    th = np.array(rec["theta"])
    assert _rel(gp.logp(th, array=True), rec["logp"]) < TOL
    assert scaled_err(gp.dlogp(th, array=True, reference_nan_quirk=True), _by_name(rec["dlogp"], rec["names"], sizes)) < TOL
    Xs = xs[np.array(rec["space_index"])]
    kw = dict(space=Xs, array=True, noise=False)
    assert scaled_err(gp.location(th, **kw), rec["location"]) < 1e-8
    assert scaled_err(gp.kernel_diag(th, **kw), rec["kernel_diag"]) < 1e-6      # LU (reference) vs Cholesky, kappa ~ 1e5
    out = gp.predict(th, mean=True, var=True, **kw)
    assert scaled_err(out["mean"], rec["mean"]) < 1e-8
    assert scaled_err(out["variance"], rec["variance"]) < 1e-6


def _check_c4_4096():
    """BASELINE config 4 (Student-t process, SE ARD on D=5) on the first 4096 rows of its generator: logp, gradient,
    predictive location / variance (with the t scaling) / quantile against the reference executed at that size."""
    import g3py_b200 as g3
    from g3py_b200 import workloads
    X, y, Xs = workloads.c4_inputs(16384, 4096)
    X, y, Xs = X[:4096], y[:4096], Xs[:32]
    gp = g3.TP(X, g3.Bias(), g3.SE(X))
    gp.observed(X, y)
    rec = FULL["C4_4096"]
    assert [h.tname for h in gp.registry.vars] == rec["names"]
    th = np.array(rec["theta"])
    assert _rel(gp.logp(th, array=True), rec["logp"]) < TOL
    assert scaled_err(gp.dlogp(th, array=True), _by_name(rec["dlogp"], rec["names"], [h.size for h in gp.registry.vars])) < TOL
    kw = dict(space=Xs, array=True, noise=True)
    assert scaled_err(gp.location(th, **kw), rec["location"]) < 1e-8
    assert scaled_err(gp.variance(th, **kw), rec["variance"]) < 1e-7
    assert scaled_err(gp.quantiler(th, q=0.975, **kw), rec["quantile_up"]) < 1e-8


def test_host_c4_4096_oracle_matches_reference():
    from g3py_b200 import workloads
    X, y, Xs = workloads.c4_inputs(16384, 4096)
    X, y, Xs = X[:4096], y[:4096], Xs[:32]
    op = orc.OracleProcess({"kind": "student", "location": {"type": "Bias"}, "kernel": {"type": "SE"}}, 5)
    rec = FULL["C4_4096"]
    th = np.array(rec["theta"])
    assert _rel(op.logp(th, X, y), rec["logp"]) < 1e-11
    assert scaled_err(op.dlogp(th, X, y), _by_name(rec["dlogp"], rec["names"], [l[1] for l in op.layout()])) < 1e-9
    pr = op.predict(th, Xs, X, y, noise=True)
    assert scaled_err(pr["variance"], rec["variance"]) < 1e-8
    assert scaled_err(pr["quantile_up"], rec["quantile_up"]) < 1e-9


@pytest.mark.gpu
def test_cuda_matches_reference_c4_at_4096():
    _check_c4_4096()


def test_oracle_matches_reference_at_full_size():
    from g3py_b200 import workloads
    X, y, Theta = workloads.c2_inputs(4096, 64)
    op = orc.OracleProcess(C2_SPEC, 3)
    rec = FULL["C2"]
    sizes = [l[1] for l in op.layout()]
    r = rec["rows"][0]
    th = np.array(r["theta"])
    assert _rel(op.logp(th, X, y), r["logp"]) < 1e-11
    assert scaled_err(op.dlogp(th, X, y, nan_quirk=True), _by_name(r["dlogp"], rec["names"], sizes)) < 1e-9
    x, y3, xs = workloads.c3_inputs(2048, 10000)
    op3 = orc.OracleProcess(C3_SPEC, 1)
    rec3 = FULL["C3"]
    th3 = np.array(rec3["theta"])
    assert _rel(op3.logp(th3, x, y3), rec3["logp"]) < 1e-11
    assert scaled_err(op3.dlogp(th3, x, y3, nan_quirk=True),
                      _by_name(rec3["dlogp"], rec3["names"], [l[1] for l in op3.layout()])) < 1e-9
    po = op3.posterior(th3, xs[np.array(rec3["space_index"])], x, y3, noise=False, solver="lu")
    assert scaled_err(po["location"], rec3["location"]) < 1e-9
    assert scaled_err(po["kernel_diag"], rec3["kernel_diag"]) < 1e-7


@pytest.mark.gpu
def test_cuda_matches_reference_at_full_size_c2():
    _check_fullsize_c2()


@pytest.mark.gpu
def test_cuda_matches_reference_at_full_size_c3():
    _check_fullsize_c3()


def test_oracle_matches_reference_at_c1_default_and_map():
    from g3py_b200 import workloads
    x, y = workloads.c1_inputs()
    op = orc.OracleProcess({"kind": "gauss", "location": {"type": "Bias"}, "kernel": {"type": "SE"}}, 1)
    for key in ("default", "map"):
        th = np.array(C1MAP["theta_" + key])
        assert _rel(op.logp(th, x, y), C1MAP["logp_" + key]) < 1e-12
        assert np.max(np.abs(op.dlogp(th, x, y) - np.array(C1MAP["dlogp_" + key]))) < 1e-9


@pytest.fixture
def fake(monkeypatch):
    import g3py_b200 as g3
    from fake_ctx import FakeContext
    ctx = FakeContext()
    monkeypatch.setattr(g3.processes, "get_context", lambda device=0: ctx)
    return ctx


# the test double has no jitter ladder (that lives in the library): PD cases only on CPU, all cases on the GPU
@pytest.mark.parametrize("name", [c for c in CASES if REF[c]["min_eig_K"] > 1e-6])
def test_host_logp_dlogp_match_reference(fake, name):
    _check_logp_dlogp(name)


@pytest.mark.parametrize("name", POST_PD)
@pytest.mark.parametrize("noise", [False, True])
def test_host_posterior_matches_reference(fake, name, noise):
    _check_posterior(name, noise)


@pytest.mark.parametrize("name", TRANSPORT)
def test_host_transport_matches_reference(fake, name):
    _check_transport(name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", TRANSPORT)
def test_cuda_transport_matches_reference(name):
    _check_transport(name)


@pytest.mark.parametrize("name", POST)
def test_host_gram_matches_reference(fake, name):
    _check_gram(name)


@pytest.mark.parametrize("name", LOGPRED)
def test_host_logpredictive_matches_reference(fake, name):
    _check_logpredictive(name)


def test_host_c1_find_map_matches_reference(fake):
    _check_c1_find_map()


@pytest.mark.gpu
def test_cuda_c1_find_map_matches_reference():
    _check_c1_find_map()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_logp_dlogp_match_reference(name):
    _check_logp_dlogp(name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", POST_PD)
@pytest.mark.parametrize("noise", [False, True])
def test_cuda_posterior_matches_reference(name, noise):
    _check_posterior(name, noise)


@pytest.mark.gpu
@pytest.mark.parametrize("name", POST)
def test_cuda_gram_matches_reference(name):
    _check_gram(name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", LOGPRED)
def test_cuda_logpredictive_matches_reference(name):
    _check_logpredictive(name)
