"""The SURVEY §8f features (transports, dot-product / Brownian / constant leaves, KernelMax, potentials, Logistic and the
Newton-inverse warpings) at sizes that span several 128-tiles and in batches, against the oracle — which
tests/test_reference_goldens.py pins to the executed reference at small N.  CPU: host logic with the NumPy device
double; GPU: the CUDA path."""
import numpy as np
import pytest

from oracle import g3_oracle as orc
from helpers import build_process, scaled_err

K = lambda t, **kw: dict(type=t, **kw)
SPECS = {
    "tgp": dict(kind="transport", chain=[dict(t="TMapping", mapping=K("BoxCoxShifted")), dict(t="TLocation", location=K("Bias")),
                                         dict(t="TKernel", kernel=K("SE"), noisy=True)]),
    "tgp_scale": dict(kind="transport", chain=[dict(t="TMapping", mapping=K("BoxCoxShifted")), dict(t="TScale", scale=K("Bias", name="Scale")),
                                               dict(t="TLocation", location=K("Bias")), dict(t="TKernel", kernel=K("SE"), noisy=True)]),
    "lin_se": dict(kind="gauss", location=K("Zero"), kernel=K("sum", k1=K("LIN"), k2=K("SE"))),
    "pol3": dict(kind="gauss", location=K("Bias"), kernel=K("sum", k1=K("POL", p=3), k2=K("SE"))),
    "dot_bw_var": dict(kind="gauss", location=K("Zero"), kernel=K("sum", k1=K("sum", k1=K("KernelDot"), k2=K("BW")), k2=K("VAR"))),
    "max": dict(kind="student", location=K("Bias"), kernel=K("max", k1=K("SE"), k2=K("scale", c=0.5, k=K("MAT32")))),
    "warptanh": dict(kind="gauss", warped=True, location=K("Bias"), kernel=K("SE"), mapping=K("WarpingTanh", n=2)),
    "warpboxcox": dict(kind="gauss", warped=True, location=K("Bias"), kernel=K("SE"), mapping=K("WarpingBoxCox", n=2)),
    "logistic": dict(kind="gauss", warped=True, location=K("Bias"), kernel=K("SE"), mapping=K("Logistic")),
    "potentials": dict(kind="gauss", warped=True, location=dict(type="Bias", potential=["Bias", "L2", 0.3]),
                       kernel=dict(type="sum", potential=["var", "L1", 0.7], k1=K("SE"), k2=dict(type="RQ", potential=["alpha", "L2", 0.2])),
                       mapping=dict(type="BoxCoxShifted", potential=["power", "L1", 0.4])),
    # NN is defined for the training Gram only (kernels.py:351): logp / gradient, no posterior
    "nn": dict(kind="gauss", location=K("Zero"), kernel=K("sum", k1=K("NN"), k2=K("SE"))),
    "nil_equals": dict(kind="gauss", location=K("Bias"),
                       kernel=K("sum", k1=K("sum", k1=K("SE"), k2=K("NIL")),
                                k2=K("sum", k1=K("scale", c=0.3, k=K("KernelEquals", eq=1.0, dims=[0, 1])),
                                     k2=K("scale", c=0.002, k=K("KernelEquals2", eq1=0.0, eq2=1.0, dims=[0, 1]))))),
}
POSITIVE = {"tgp", "tgp_scale", "warpboxcox", "logistic", "potentials"}
LOGP_ONLY = {"nn"}


def _problem(name, N, B, seed=0):
    rng = np.random.default_rng(seed)
    D = 2
    X = rng.uniform(0.2, 5.0, size=(N, D))
    y = np.sin(X[:, 0]) + 0.4 * np.cos(0.7 * X[:, 1]) + 0.1 * rng.standard_normal(N)
    if name in POSITIVE:
        y = np.exp(0.5 * y) + 0.2
    if name == "nil_equals":
        X[:, 0] = np.floor(X[:, 0] / 2.0)                   # {0, 1, 2}: the equality metrics are not identically zero
    op = orc.build_process(SPECS[name], D)
    th = []
    for nm, size, pos in op.layout():
        v = 0.1 * rng.standard_normal((B, size))
        if nm.endswith("Scale_Bias"):          # the TScale factor must stay positive
            v += 1.3
        elif "Noise" in nm:                    # max(k1, k2) is not PSD in general: more noise keeps K definite
            v += np.log((2.0 if name in ("max", "nn", "nil_equals") else 0.05) * np.var(y))
        elif nm.endswith("_var"):
            v += np.log(np.var(y))
        elif nm.endswith("_bias"):
            v += np.log(0.5)
        elif nm.endswith("LIN_rate") or nm.endswith("POL_rate") or nm.endswith("KernelDot_rate"):
            v += np.log(0.3)
        elif nm.endswith("_power"):
            v = np.full((B, size), np.log(0.7)) + 0.02 * rng.standard_normal((B, size))
        elif nm.endswith("Freedom_degree"):
            v = np.full((B, size), np.log(5.0))
        elif nm.endswith("Logistic_lower"):
            v = np.full((B, size), np.min(y) - 0.4)
        elif nm.endswith("Logistic_high"):
            v = np.full((B, size), np.log(np.max(y) - np.min(y) + 0.9))
        elif nm.endswith("WarpingTanh_a"):
            v += np.log(0.3)
        elif nm.endswith("WarpingTanh_c"):
            v += -np.mean(y)
        elif nm.endswith("WarpingBoxCox_w"):
            v += np.log(0.5)
        elif nm.endswith("_Bias") and name not in POSITIVE and "warp" not in name:
            v += np.mean(y)
        th.append(v)
    return X, y, op, np.concatenate(th, axis=1)


def _check(name, N, B, M):
    X, y, op, Th = _problem(name, N, B)
    gp = build_process(SPECS[name], X)
    gp.observed(X, y)
    assert [(n, s) for n, s, _ in op.layout()] == [(v.name[len(gp.name) + 1:], v.size) for v in gp.registry.vars]
    lp, g, info = gp.logp_dlogp_batch(Th, reference_nan_quirk=True)
    for b in range(B):
        want = op.logp(Th[b], X, y)
        assert abs(lp[b] - want) <= 1e-9 * abs(want), (name, b)
        assert scaled_err(g[b], op.dlogp(Th[b], X, y, nan_quirk=True)) < 1e-9, (name, b)
    if name in LOGP_ONLY:
        return
    Xs = X[:M] + 0.05
    if name == "nil_equals":
        Xs[:, 0] = X[:M, 0]
    if SPECS[name].get("kind") == "transport":
        v = np.random.default_rng(5).standard_normal(M)
        for prior in (True, False):
            got = gp.transport(Th[0], space=Xs, vector=v, prior=prior, noise=True, array=True)
            assert scaled_err(got, op.transport(Th[0], Xs, v, X, y, prior=prior, noise=True)) < 1e-7
        return
    out = gp.predict(Th[0], space=Xs, array=True, var=True, median=True, quantiles=True, noise=True)
    pr = op.predict(Th[0], Xs, X, y, noise=True)
    post, _, _ = gp._posterior(Th[0], Xs, noise=True)
    pc = op.posterior(Th[0], Xs, X, y, noise=True, solver="chol")
    assert scaled_err(post["location"], pc["location"]) < 1e-9
    assert scaled_err(post["kernel_diag"], pc["kernel_diag"]) < 1e-8
    for key in ("mean", "median", "quantile_up"):
        assert scaled_err(out[key], pr[key]) < 1e-7, (name, key)
    assert scaled_err(out["variance"], pr["variance"]) < 1e-6


@pytest.fixture
def fake(monkeypatch):
    import g3py_b200 as g3
    from fake_ctx import FakeContext
    ctx = FakeContext()
    monkeypatch.setattr(g3.processes, "get_context", lambda device=0: ctx)
    return ctx


@pytest.mark.parametrize("name", list(SPECS))
def test_host_widened_features(fake, name):
    _check(name, 60, 2, 7)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SPECS))
def test_cuda_widened_features_multi_tile(name):
    _check(name, 300 if "warp" in name else 520, 3, 40)


COMPOSED = [
    ("LogShifted", "LinearMapping"), ("BoxCoxShifted", "SinhArcsinh"), ("LinearMapping", "ArcsinhLinear"),
    ("BoxCoxLinear", "Logistic"), ("ArcsinhLinear", "WarpingTanh"), ("LogShifted", "WarpingBoxCox"),
    ("SinhArcsinh", "SinhArcsinh2"), ("WarpingTanh", "SinhArcsinh"), ("WarpingBoxCox", "BoxCoxLinear"),
]


@pytest.mark.parametrize("pair", COMPOSED, ids=lambda p: "@".join(p))
def test_composed_mapping_gradients_are_the_analytic_chain(pair):
    """m1 @ m2 (reference mappings.py:57-70): d inv / d theta and d logdet_dinv / d theta of the composition - including
    the term through which m1's hypers move the argument of m2's log-Jacobian - against 4th-order central differences
    of the composition's own inv / logdet_dinv."""
    import g3py_b200 as g3
    from g3py_b200.hypers import HyperVar
    rng = np.random.default_rng(11)
    y = rng.uniform(0.8, 2.2, size=9)
    vals = {
        "LogShifted": dict(shift=0.1), "LinearMapping": dict(shift=0.3, scale=1.3), "BoxCoxShifted": dict(shift=0.2, power=0.7),
        "BoxCoxLinear": dict(shift=0.2, scale=1.2, power=0.7), "SinhArcsinh": dict(shift=0.1, scale=1.2),
        "SinhArcsinh2": dict(shift=-0.2, scale=0.8), "ArcsinhLinear": dict(shift=0.2, scale=1.4),
        "Logistic": dict(lower=-3.0, high=9.0, location=0.3, scale=0.8),
        "WarpingTanh": dict(a=np.array([0.3, 0.5]), b=np.array([0.8, 1.2]), c=np.array([-0.2, 0.4])),
        "WarpingBoxCox": dict(shift=np.array([4.0, 5.0]), power=np.array([0.7, 1.3]), w=np.array([0.6, 0.4])),
    }
    maps, theta, where = [], [], []
    for nm in pair:
        cls = getattr(g3, nm.rstrip("2"))
        m = cls(n=2) if nm.startswith("Warping") else cls(name=nm)
        for k, v in vals[nm].items():
            hv = HyperVar(nm + "_" + k, np.size(v), scalar=np.size(v) == 1)
            where.append((hv, k, len(theta), np.size(v)))
            theta.extend(np.atleast_1d(v).tolist())
            setattr(m, k, hv)
        maps.append(m)
    comp = maps[0] @ maps[1]
    theta = np.array(theta)

    def lookup(t):
        def p(h):
            _, _, off, size = next(w for w in where if w[0] is h)
            return t[off] if size == 1 else t[off:off + size]
        return p
    dinv, dld = comp.grads(y, lookup(theta))
    for key, k, off, size in where:
        for j in range(size):
            h = 1e-4 * max(1.0, abs(theta[off + j]))
            f = lambda s: (lambda t: (comp.inv(y, lookup(t)), comp.logdet_dinv(y, lookup(t))))(theta + s * h * np.eye(len(theta))[off + j])
            (a2, b2), (a1, b1), (c1, d1), (c2, d2) = f(2), f(1), f(-1), f(-2)
            fd_inv = (-a2 + 8 * a1 - 8 * c1 + c2) / (12 * h)
            fd_ld = (-b2 + 8 * b1 - 8 * d1 + d2) / (12 * h)
            got_inv = np.asarray(dinv[key]).reshape(size, -1)[j] if size > 1 else np.asarray(dinv[key])
            got_ld = np.atleast_1d(dld[key])[j]
            assert scaled_err(got_inv, fd_inv) < 1e-8, (pair, k, j)
            assert abs(got_ld - fd_ld) <= 1e-8 * max(1.0, abs(fd_ld)), (pair, k, j)
    # and the slope itself
    fd = g3.hypers.mappings.Mapping.dlog_dinv_dy(comp, y, lookup(theta))
    assert scaled_err(comp.dlog_dinv_dy(y, lookup(theta)), fd) < 1e-8


def test_power_blackbox_means_and_new_mappings_on_fake(monkeypatch):
    """Power / BlackBox means (means.py:32-41,162-182), BoxCoxLinear2 and MappingInvSum (mappings.py:73-85,218-251) through
    the public API on the NumPy device double: Power(n) equals Linear on x**n, BlackBox(m) equals a GP on y - m, and the
    gradients of the new warpings agree with central differences of logp."""
    import g3py_b200 as g3
    from fake_ctx import FakeContext
    ctx = FakeContext()
    monkeypatch.setattr(g3.processes, "get_context", lambda device=0: ctx)
    rng = np.random.default_rng(5)
    X = rng.uniform(0.5, 3.0, size=(40, 2))
    y = np.sin(X[:, 0]) + 0.3 * X[:, 1] ** 2 + 0.05 * rng.standard_normal(40)
    a = g3.GP(X, g3.Power(X, n=2), g3.SE(X))
    b = g3.GP(X ** 2, g3.Linear(X ** 2), g3.SE(X ** 2))
    a.observed(X, y)
    b.observed(X ** 2, y)
    assert [n.split("_", 2)[2] for n, _, _ in a.layout][:2] == ["Constant", "Coeff"]
    pa, pb = a.dict_to_array(a.params_default), b.dict_to_array(b.params_default)
    assert np.allclose(pa[:3], pb[:3])                       # same defaults for constant / coeff (means.py:178-180)
    th = pa.copy()
    th[3:] = [0.1, -0.2, 0.3, np.log(0.05)]
    tb = th.copy()
    tb[4:6] = th[4:6] - np.log(2.0) - np.log(np.mean(X, axis=0)) * 0       # different metric: compare means only
    m = rng.standard_normal(40)
    c = g3.GP(X, g3.BlackBox(m), g3.SE(X))
    d = g3.GP(X, g3.Zero(), g3.SE(X))
    c.observed(X, y)
    d.observed(X, y - m)
    t0 = np.array([0.1, -0.2, 0.3, np.log(0.05)])
    assert c.logp(t0, array=True) == pytest.approx(d.logp(t0, array=True), rel=1e-12)
    assert np.allclose(c.dlogp(t0, array=True), d.dlogp(t0, array=True), rtol=1e-10)
    # Power mean value / jacobian
    nat = a.natural(th)
    p = a._accessor(nat)
    assert np.allclose(a.f_location(X, p), nat[0] + (X ** 2) @ nat[1:3])
    g = a.dlogp(th, array=True)
    for i in range(3):
        e = np.zeros_like(th)
        e[i] = 1e-5
        fd = (a.logp(th + e, array=True) - a.logp(th - e, array=True)) / 2e-5
        assert g[i] == pytest.approx(fd, rel=1e-5, abs=1e-6)
    # new warpings: gradient of logp wrt every hyper against central differences
    yp = np.exp(0.4 * y) + 0.3
    for mp in (g3.BoxCoxLinear2(), g3.MappingInvSum(g3.BoxCoxLinear2(name="A"), g3.LogShifted(name="B"))):
        w = g3.WGP(X, g3.Bias(), g3.SE(X), mp)
        w.observed(X, yp)
        tw = w.dict_to_array(w.params_default)
        lay = [n for n, s, _ in w.layout for _ in range(s)]
        tw[lay.index("WGP_Noise_var")] = np.log(0.05)
        for k, nm in enumerate(lay):
            if nm.endswith("_shift"):
                tw[k] = 0.4
        gw = w.dlogp(tw, array=True)
        for i in range(len(tw)):
            e = np.zeros_like(tw)
            e[i] = 1e-5
            fd = (w.logp(tw + e, array=True) - w.logp(tw - e, array=True)) / 2e-5
            assert gw[i] == pytest.approx(fd, rel=2e-5, abs=1e-5), (type(mp).__name__, lay[i])
        pr = w.predict(tw, space=X[:5], array=True, var=True, median=True)
        assert np.all(np.isfinite(pr["mean"])) and np.all(pr["variance"] >= 0)


def test_nn_nil_equals_kernels_on_fake(monkeypatch):
    """NN / NIL / KernelEquals / KernelEquals2 (kernels.py:262-288,309-351) through the public API on the NumPy device
    double: NIL adds nothing, the equality kernels count coordinates equal to their constants, the NN gradient agrees with
    central differences, and the posterior of an NN model raises as the reference's two-argument cov does (kernels.py:351)."""
    import g3py_b200 as g3
    from fake_ctx import FakeContext
    ctx = FakeContext()
    monkeypatch.setattr(g3.processes, "get_context", lambda device=0: ctx)
    rng = np.random.default_rng(11)
    X = rng.uniform(0.0, 2.0, size=(30, 2))
    X[:, 0] = np.floor(X[:, 0] * 1.5)                                  # {0, 1, 2}
    y = np.sin(X[:, 1]) + 0.3 * (X[:, 0] == 1.0) + 0.05 * rng.standard_normal(30)
    base = g3.GP(X, g3.Zero(), g3.SE(X))
    with_nil = g3.GP(X, g3.Zero(), g3.SE(X) + g3.NIL(X))
    base.observed(X, y)
    with_nil.observed(X, y)
    t0 = np.array([0.1, -0.2, 0.3, np.log(0.05)])
    assert with_nil.ndim == base.ndim                                  # NIL has no hypers
    assert with_nil.logp(t0, array=True) == pytest.approx(base.logp(t0, array=True), rel=1e-13)
    K1 = g3.KernelEquals(X[:, :1], eq=1.0).cov(X[:, :1])
    assert np.array_equal(K1, np.outer(X[:, 0] == 1.0, X[:, 0] == 1.0).astype(float))
    K2 = g3.KernelEquals2(X[:, :1], eq1=0.0, eq2=2.0).cov(X[:, :1], X[:5, :1])
    a, b = X[:, 0][:, None], X[:5, 0][None, :]
    assert np.array_equal(K2, ((a == 0.0) & (b == 2.0)).astype(float) + ((a == 2.0) & (b == 0.0)))
    nn = g3.GP(X, g3.Zero(), g3.NN(X))
    nn.observed(X, y)
    assert [n.split("_", 1)[1] for n, _, _ in nn.layout] == ["NN_var", "NN_rate", "NN_bias", "Noise_var"]
    th = np.array([np.log(0.2), np.log(0.9), np.log(1.1), np.log(0.6), np.log(0.2)])
    g = nn.dlogp(th, array=True)
    for i in range(len(th)):
        e = np.zeros_like(th)
        e[i] = 1e-5
        fd = (nn.logp(th + e, array=True) - nn.logp(th - e, array=True)) / 2e-5
        assert g[i] == pytest.approx(fd, rel=2e-5, abs=1e-6)
    with pytest.raises(NotImplementedError):
        nn.predict(th, space=X[:4], array=True)
