"""Parity of the CUDA path (through the C ABI / ctypes) against the NumPy oracle.  Needs a B200.

Tolerances: the north star asks for 1e-9 relative on fp64 logp, gradients and posterior moments;
element-wise Gram values are held to 1e-12.
"""
import numpy as np
import pytest
import scipy.linalg as sla

import g3py_b200 as g3
from g3py_b200 import _cabi as cabi
from oracle import g3_oracle as orc
from helpers import build_process, build_kernel, rel_err, scaled_err

pytestmark = pytest.mark.gpu

TOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    return g3.processes.get_context(0)


def _desc_and_theta(kspec, X, rng):
    """compile a kernel spec through the product algebra; random natural hypers in oracle order."""
    D = X.shape[1]
    ok = orc.build_kernel(kspec, D)
    k = build_kernel(kspec, X)
    reg = g3.Registry()
    k.check_dims(X)
    k.check_hypers("", reg)
    b = g3.DescBuilder(D)
    k.compile(b)
    desc = b.finish()
    off = 0
    for v in reg.vars:
        v.offset = off
        off += v.size
    names_o = [(h.name, h.size) for h in ok.layout()]
    names_p = [(v.name, v.size) for v in reg.vars]
    assert names_o == names_p, (names_o, names_p)
    th_o = np.exp(rng.normal(0.0, 0.4, size=off))
    th_p = np.empty(desc.n_theta)
    for h, o, size, const in b.slots:
        th_p[o:o + size] = const if h is None else th_o[h.offset:h.offset + h.size]
    return ok, desc, th_o, th_p, b.slots


KSPECS = {
    "SE": {"type": "SE"},
    "OU": {"type": "OU"},
    "MAT32": {"type": "MAT32"},
    "MAT52": {"type": "MAT52"},
    "RQ": {"type": "RQ"},
    "SIN": {"type": "SIN"},
    "WN": {"type": "WN"},
    "COS": {"type": "COS"},
    "SINC": {"type": "SINC"},
    "SM": {"type": "SM"},
    "SM+COS+SINC": {"type": "sum", "k1": {"type": "sum", "k1": {"type": "SM"}, "k2": {"type": "COS"}}, "k2": {"type": "SINC"}},
    "SMxSE": {"type": "prod", "k1": {"type": "SM"}, "k2": {"type": "SE"}},
    "SE+MAT52+Noise": {"type": "sum", "k1": {"type": "sum", "k1": {"type": "SE"}, "k2": {"type": "MAT52"}},
                       "k2": {"type": "Noise", "name": "Noise"}},
    "SINxSE": {"type": "prod", "k1": {"type": "SIN"}, "k2": {"type": "SE"}},
    "2*(RQ)+0.5": {"type": "shift", "c": 0.5, "k": {"type": "scale", "c": 2.0, "k": {"type": "RQ"}}},
    "SE[0:1]*OU[1:3]+MAT32": {"type": "sum", "k1": {"type": "prod", "k1": {"type": "SE", "dims": [0, 1]},
                                                     "k2": {"type": "OU", "dims": [1, 3], "var": 1.0}},
                              "k2": {"type": "MAT32"}},
}


@pytest.mark.parametrize("name", list(KSPECS))
@pytest.mark.parametrize("n1,n2,D", [(1, 1, 1), (5, 3, 2), (127, 129, 3), (300, 257, 3)])
def test_gram_parity(ctx, name, n1, n2, D):
    kspec = KSPECS[name]
    if "dims" in str(kspec) and D < 3:
        pytest.skip("spec needs 3 columns")
    rng = np.random.default_rng(hash((name, n1, n2, D)) % 2 ** 32)
    X1 = rng.uniform(0, 3, size=(n1, D))
    X2 = rng.uniform(0, 3, size=(n2, D))
    X2[: min(n1, n2) // 2] = X1[: min(n1, n2) // 2]           # exact coincidences (WN, d = 0)
    ok, desc, th_o, th_p, _ = _desc_and_theta(kspec, X1, rng)
    Kc, st = ctx.gram(desc, X1, X2, th_p[None, :])
    Ko = orc.tt_to_num(ok.cov(th_o, X1, X2, False))
    assert scaled_err(Kc[0], Ko) < 1e-12
    Ks, st = ctx.gram(desc, X1, None, th_p[None, :])
    Kso = orc.tt_to_num(ok.cov(th_o, X1, X1, True))
    assert scaled_err(Ks[0], Kso) < 1e-12
    assert np.array_equal(Ks[0], Ks[0].T)                      # symmetric to the bit


def test_gram_batched_and_nonfinite(ctx):
    rng = np.random.default_rng(7)
    X = rng.uniform(0, 3, size=(200, 2))
    ok, desc, th_o, th_p, slots = _desc_and_theta(KSPECS["SE+MAT52+Noise"], X, rng)
    B = 5
    Th = np.tile(th_p, (B, 1)) * np.exp(rng.normal(0, 0.2, size=(B, len(th_p))))
    K, st = ctx.gram(desc, X, None, Th)
    for b in range(B):
        th_b = np.empty_like(th_o)
        for h, o, size, const in slots:
            th_b[h.offset:h.offset + size] = Th[b, o:o + size]
        assert scaled_err(K[b], ok.cov(th_b, X, X, True)) < 1e-12
    assert np.all(st == 0)
    Th[2, 0] = np.inf                                           # var = inf -> inf*exp(-d): scrubbed like tt_to_num
    K, st = ctx.gram(desc, X, None, Th)
    assert st[2] & cabi.ST_NONFINITE_INPUT and st[0] == 0
    assert np.all(np.isfinite(K))


@pytest.mark.parametrize("name", list(KSPECS))
def test_gram_vjp_parity(ctx, name):
    kspec = KSPECS[name]
    rng = np.random.default_rng(11)
    n1, n2, D = 150, 131, 3
    X1 = rng.uniform(0, 3, size=(n1, D))
    X2 = rng.uniform(0, 3, size=(n2, D))
    ok, desc, th_o, th_p, slots = _desc_and_theta(kspec, X1, rng)
    for same in (False, True):
        xb = X1 if same else X2
        W = rng.standard_normal((n1, xb.shape[0]))
        g = ctx.gram_vjp(desc, X1, None if same else X2, th_p[None, :], W[None])[0]
        dK = ok.dcov(th_o, X1, xb, same)
        go = np.array([np.sum(W * d) for d in dK])
        gp = np.empty_like(go)
        for h, o, size, const in slots:
            if h is not None:
                gp[h.offset:h.offset + size] = g[o:o + size]
        assert scaled_err(gp, go) < 1e-11, (name, same)


@pytest.mark.parametrize("n", [1, 2, 5, 128, 129, 300, 1000])
def test_potrf_parity(ctx, n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n + 3))
    K = A @ A.T + n * 1e-3 * np.eye(n)
    L, info, jit = ctx.potrf_robust(K)
    Lo = sla.cholesky(K, lower=True)
    assert info == 0 and jit == 0.0
    assert np.all(np.triu(L, 1) == 0.0)
    assert scaled_err(L, Lo) < 1e-11
    assert scaled_err(L @ L.T, K) < 1e-13


def test_potrf_jitter_ladder(ctx):
    """Rank-deficient and indefinite inputs: same ladder outcome as CholeskyRobust (tensors.py:197-222)."""
    rng = np.random.default_rng(3)
    n = 200
    A = rng.standard_normal((n, 20))
    K_psd = A @ A.T                                              # rank 20: dpotrf fails, jitter repairs it
    K_neg = K_psd - 5.0 * np.eye(n)                              # negative diagonal entries
    K_bad = -np.eye(n) * 1e30                                    # ladder exhausted? (pre-shift makes it PD)
    for K in (K_psd, K_neg, K_bad):
        Lo, info_o = orc.cholesky_robust(K, return_info=True)
        L, info, jit = ctx.potrf_robust(K)
        assert info == info_o, (info, info_o)
        if info_o > 0:
            assert scaled_err(L @ L.T, Lo @ Lo.T) < 1e-9
    Kb = np.stack([K_psd + np.eye(n), K_psd, K_neg])
    Lb, infob, jitb = ctx.potrf_robust(Kb)
    assert infob[0] == 0 and infob[1] > 0 and infob[2] > 0
    assert scaled_err(Lb[0], sla.cholesky(Kb[0], lower=True)) < 1e-11
    Knan = K_psd.copy()
    Knan[3, 3] = np.nan
    L, info, jit = ctx.potrf_robust(Knan)
    assert info == -1                                            # caller applies the 1e-10*I fallback


# ------------------------------------------------------------------------------------------- logp / dlogp
SPECS = {
    "C1": {"kind": "gauss", "location": {"type": "Bias"}, "kernel": {"type": "SE"}},
    "C2": {"kind": "gauss", "location": {"type": "Bias"},
           "kernel": {"type": "sum", "k1": {"type": "SE"}, "k2": {"type": "MAT52"}}},
    "C3": {"kind": "gauss", "warped": True, "location": {"type": "Bias"},
           "kernel": {"type": "prod", "k1": {"type": "SIN"}, "k2": {"type": "SE"}},
           "mapping": {"type": "BoxCoxShifted"}},
    "C4": {"kind": "student", "location": {"type": "Bias"}, "kernel": {"type": "SE"}},
    "WTP": {"kind": "student", "warped": True, "location": {"type": "Linear"},
            "kernel": {"type": "sum", "k1": {"type": "RQ"}, "k2": {"type": "OU"}},
            "mapping": {"type": "ArcsinhLinear"}},
    "noiseless": {"kind": "gauss", "location": {"type": "Zero"}, "noisy": False,
                  "kernel": {"type": "sum", "k1": {"type": "MAT32"}, "k2": {"type": "WN"}}},
}


def _data(name, N):
    if name == "C1":
        x, y = orc.c1_inputs()
        return x[:N], y[:N]
    if name in ("C2", "noiseless"):
        X, y, _ = orc.c2_inputs(N, 1)
        return X, y
    if name == "C3":
        x, y, _ = orc.c3_inputs(N, 8)
        return x, y
    X, y, _ = orc.c4_inputs(N, 8)
    return X, y


def _theta0(name, gp, op, X, y, rng, B):
    """default hypers of the product, perturbed; kernels kept well conditioned (noise floor)."""
    th = gp.dict_to_array(gp.params_default)
    Th = np.tile(th, (B, 1)) + 0.1 * rng.standard_normal((B, len(th)))
    lay = gp.layout
    off = 0
    for nm, size, pos in lay:
        if nm.endswith("SIN_rate"):
            Th[:, off:off + size] = np.log(0.1)                  # a3-iii: keeps the periodic kernel PD
        if nm.endswith("SIN_freq"):
            Th[:, off:off + size] = np.log(0.2)
        if nm.endswith("Freedom_degree"):
            Th[:, off:off + size] = np.log(5.0) + 0.1 * rng.standard_normal((B, size))
        if nm.endswith("BoxShift_power"):
            Th[:, off:off + size] = np.log(0.7) + 0.05 * rng.standard_normal((B, size))
        off += size
    return Th


@pytest.mark.parametrize("name,N", [("C1", 200), ("C2", 333), ("C3", 256), ("C4", 300), ("WTP", 130), ("noiseless", 257)])
def test_logp_dlogp_parity(name, N):
    spec = SPECS[name]
    X, y = _data(name, N)
    rng = np.random.default_rng(5)
    gp = build_process(spec, X)
    gp.observed(X, y)
    op = orc.OracleProcess(spec, X.shape[1])
    assert [(a.split("_", 1)[1], b, c) for a, b, c in gp.layout] == op.layout()
    B = 3
    Th = _theta0(name, gp, op, X, y, rng, B)
    lp, g, info = gp.logp_dlogp_batch(Th)
    for b in range(B):
        t = op.logp_terms(Th[b], X, y)
        assert t["info"] == 0
        assert abs(info["beta"][b] - t["beta"]) <= TOL * abs(t["beta"])
        assert abs(info["logdet"][b] - t["logdet"]) <= TOL * max(abs(t["logdet"]), 1.0)
        lo = op.logp(Th[b], X, y)
        assert abs(lp[b] - lo) <= TOL * abs(lo), (name, b, lp[b], lo)
        go = op.dlogp(Th[b], X, y, method="analytic")
        assert scaled_err(g[b], go) < TOL, (name, b, g[b], go)
        gm = op.dlogp(Th[b], X, y, method="murray")             # the reference's own reverse-mode route
        assert scaled_err(g[b], gm) < 1e-7
    # single-theta entries agree with the batch, dict and array forms agree
    assert abs(gp.logp(Th[0], array=True) - lp[0]) <= 1e-12 * abs(lp[0])
    assert abs(gp.logp(gp.array_to_dict(Th[0])) - lp[0]) <= 1e-12 * abs(lp[0])
    assert scaled_err(gp.dlogp(Th[1], array=True), g[1]) < 1e-12


def test_student_kat_notebook():
    """The only known-answer vectors in the reference tree (notebooks/07-Student-t-Process.ipynb:206-218):
    WarpedStudentTProcess(x, Bias(), SE(x), ArcsinhLinear()) on X=[[0],[1]], y=[0,1]."""
    X = np.array([[0.0], [1.0]])
    y = np.array([0.0, 1.0])
    tp = g3.WTP(X, g3.Bias(), g3.SE(X), g3.ArcsinhLinear(y))
    tp.observed(X, y)
    assert [n for n, _, _ in tp.layout] == ["WTP_Bias_Bias", "WTP_SE_var", "WTP_SE_rate", "WTP_Noise_var",
                                            "WTP_ArcsinhLinear_shift", "WTP_ArcsinhLinear_scale", "WTP_Freedom_degree"]
    kat = [(np.zeros(7), -2.620996), (np.array([0.5, np.log(0.25), np.log(0.5), np.log(0.25), 0.5, np.log(0.5), 0.0]), -2.654731)]
    for th, want in kat:
        got = tp.logp(th, array=True)
        assert abs(got - want) < 5e-6, (got, want)              # float32 prints in the notebook


def test_jitter_ladder_in_logp():
    """Duplicated inputs without noise -> singular K: the ladder result equals the oracle's."""
    rng = np.random.default_rng(9)
    X = rng.uniform(0, 5, size=(150, 2))
    X[75:] = X[:75]
    y = np.sin(X[:, 0])
    spec = {"kind": "gauss", "location": {"type": "Zero"}, "kernel": {"type": "SE"}, "noisy": False}
    gp = build_process(spec, X)
    gp.observed(X, y)
    op = orc.OracleProcess(spec, 2)
    th = np.array([0.0, np.log(0.3), np.log(0.3)])
    t = op.logp_terms(th, X, y)
    ll, g, info = gp.logp_dlogp_batch(th[None])
    assert t["info"] > 0
    assert (info["status"][0] & cabi.ST_JITTER) and ((info["status"][0] >> 8) & 0xFF) == t["info"]
    assert abs(info["logdet"][0] - t["logdet"]) <= 1e-6 * abs(t["logdet"])
    assert abs(ll[0] - t["loglike"]) <= 1e-6 * abs(t["loglike"])


def test_guards():
    """Non-finite delta -> float32(-1e30) like gaussian.py:234-241."""
    x, y = orc.c1_inputs()
    y = y.copy()
    y[5] = np.nan
    gp = build_process(SPECS["C1"], x)
    gp.observed(x, y)
    assert gp.logp() == float(np.float32(-1e30))
    assert np.all(gp.dlogp() == 0.0)


# ------------------------------------------------------------------------------------------- posterior
@pytest.mark.parametrize("name,N,M", [("C1", 200, 333), ("C3", 256, 500), ("C4", 300, 129), ("noiseless", 200, 64)])
@pytest.mark.parametrize("noise", [False, True])
def test_posterior_parity(name, N, M, noise):
    spec = SPECS[name]
    X, y = _data(name, N)
    rng = np.random.default_rng(13)
    Xs = X[rng.integers(0, N, size=M)] + 0.05 * rng.standard_normal((M, X.shape[1]))
    gp = build_process(spec, X)
    gp.observed(X, y)
    op = orc.OracleProcess(spec, X.shape[1])
    th = _theta0(name, gp, op, X, y, rng, 1)[0]
    out = gp.predict(th, space=Xs, array=True, mean=True, std=True, var=True, cov=True, median=True, quantiles=True,
                     noise=noise)
    for solver, tol in (("chol", TOL), ("lu", 1e-8)):            # the reference's LU route: bounded by kappa(K) * eps
        po = op.posterior(th, Xs, X, y, noise=noise, cov=True, solver=solver)
        post, _, _ = gp._posterior(th, Xs, noise=noise, cov=True)
        assert scaled_err(post["location"], po["location"]) < tol
        assert scaled_err(post["kernel_diag"], po["kernel_diag"]) < tol * 10
        assert scaled_err(post["kernel"], po["kernel"]) < tol * 10
    pr = op.predict(th, Xs, X, y, noise=noise)
    for k in ("mean", "variance", "std", "median", "quantile_up", "quantile_down"):
        assert scaled_err(out[k], pr[k]) < 1e-8, k


def test_find_map_c1():
    """BASELINE config 1: SE + noise on the 1-D sine, BFGS through logp/dlogp."""
    x, y = orc.c1_inputs()
    gp = g3.GP(x, g3.Bias(), g3.SE(x))
    gp.observed(x, y)
    op = orc.OracleProcess(SPECS["C1"], 1)
    th0 = gp.dict_to_array(gp.params_default)
    assert abs(gp.logp(th0, array=True) - op.logp(th0, x, y)) <= TOL * abs(op.logp(th0, x, y))
    best, pts = gp.find_MAP(return_points=True)
    thm = gp.dict_to_array(best)
    assert gp.logp(thm, array=True) > gp.logp(th0, array=True)
    assert np.linalg.norm(gp.dlogp(thm, array=True)) < 1e-2
    assert abs(gp.logp(thm, array=True) - op.logp(thm, x, y)) <= TOL * abs(op.logp(thm, x, y))


# ------------------------------------------------------------------------------------------- full size
def test_c2_full_size_properties():
    """N=4096, B=64 (BASELINE config 2): two items against the oracle, the rest through
    size-independent properties (batch == single evaluation; finite, ordered status)."""
    X, y, Theta = orc.c2_inputs(4096, 64)
    gp = build_process(SPECS["C2"], X)
    gp.observed(X, y)
    op = orc.OracleProcess(SPECS["C2"], 3)
    lp, g, info = gp.logp_dlogp_batch(Theta)
    assert np.all(info["status"] == 0) and np.all(np.isfinite(lp)) and np.all(np.isfinite(g))
    for b in (0, 63):
        lo = op.logp(Theta[b], X, y)
        go = op.dlogp(Theta[b], X, y)
        assert abs(lp[b] - lo) <= TOL * abs(lo)
        assert scaled_err(g[b], go) < TOL
    l1, g1, _ = gp.logp_dlogp_batch(Theta[17:18])
    assert abs(l1[0] - lp[17]) <= 1e-12 * abs(lp[17])
    assert scaled_err(g1[0], g[17]) < 1e-11
    # directional derivative by central differences of the batched logp itself
    rng = np.random.default_rng(0)
    v = rng.standard_normal(Theta.shape[1])
    v /= np.linalg.norm(v)
    h = 1e-4
    lpp = gp.logp_batch(np.stack([Theta[5] + h * v, Theta[5] - h * v]))
    fd = (lpp[0] - lpp[1]) / (2 * h)
    assert abs(fd - g[5].dot(v)) <= 1e-6 * max(1.0, abs(fd))


def test_repeat_runs_bitwise_identical(ctx):
    """Determinism stress of the batched factorisation: the same 64-item batch factored repeatedly must give
    bit-identical tiles (guards the TMA-pipeline stage-release ordering in dgemm_nt_kernel)."""
    X, y, Theta = orc.c2_inputs(2048, 64)
    gp = build_process(SPECS["C2"], X)
    gp.observed(X, y)
    thk = gp._kernel_theta(gp.natural(Theta))
    out, _ = gp.ctx.debug_potrf_stress(gp.desc, thk, 12)
    assert not out[:, 0].any(), out
    res = []
    for _ in range(3):
        lp, g, info = gp.logp_dlogp_batch(Theta)
        res.append(np.concatenate([lp, g.ravel()]))
    assert np.array_equal(res[0], res[1]) and np.array_equal(res[0], res[2])


def _gpu_count():
    import ctypes
    try:
        n = ctypes.c_int(0)
        rt = ctypes.CDLL("libcuda.so.1")
        rt.cuInit(0)
        rt.cuDeviceGetCount(ctypes.byref(n))
        return n.value
    except OSError:
        return 0


@pytest.mark.parametrize("N,nb,lookahead", [(2048, 256, True), (2048, 256, False), (1536, 512, True), (1024, 1024, True)])
def test_block_cyclic_cholesky_single_gpu(N, nb, lookahead):
    """g3_dist_factor / g3_dist_solve / g3_dist_residual on a 1 x 1 grid (no communicator): every piece, u, beta and
    log-det against NumPy's Cholesky; the residual probe agrees with a host computation of the same quantity."""
    from g3py_b200.dist import run_dist_cholesky
    from g3py_b200 import workloads
    ctx = g3.Context(0)
    try:
        r = run_dist_cholesky(ctx, N, nb=nb, lookahead=lookahead, verify=4)
        X, y = workloads.c5_inputs(N)
        d = ((X[:, None, :] - X[None, :, :]) ** 2 * 0.5).sum(-1)
        L = np.linalg.cholesky(np.exp(-d) + 0.01 * np.eye(N))
        assert r["info"] == 0 and r["grid"] == [1, 1]
        assert abs(r["logdet"] - np.log(np.diag(L)).sum()) <= 1e-10 * abs(np.log(np.diag(L)).sum())
        for J in range(N // nb):
            P = ctx.dist_read_piece(J, nb)
            want = L[J * nb:, J * nb:(J + 1) * nb]
            assert P.shape == want.shape
            assert np.abs(np.tril(P[:nb]) - want[:nb]).max() < 1e-11
            if P.shape[0] > nb:
                assert np.abs(P[nb:] - want[nb:]).max() < 1e-11
        s = ctx.dist_solve(y, want_u=True)
        uref = np.linalg.solve(L, y)
        assert np.abs(s["u"] - uref).max() < 1e-10
        assert abs(r["beta"] - uref @ uref) <= 1e-10 * (uref @ uref) and s["beta"] == r["beta"]
        want_logp = -0.5 * N * np.log(2 * np.pi) - 0.5 * (uref @ uref) - np.log(np.diag(L)).sum()
        assert abs(r["logp"] - want_logp) <= 1e-10 * abs(want_logp)
        assert len(r["residual"]) == 4 and max(r["residual"]) < 1e-12
    finally:
        ctx.close()


@pytest.mark.parametrize("N,nb", [(1536, 256), (2048, 512), (1024, 1024)])
def test_block_cyclic_gradient_and_posterior_single_gpu(N, nb):
    """g3_dist_grad / g3_dist_posterior on a 1 x 1 grid against the single-GPU fused path (g3_gp_logp_grad /
    g3_gp_posterior) on the same kernel, data and hypers: dtheta, ddelta, posterior mean and variance."""
    from g3py_b200.dist import se_noise_desc
    from g3py_b200 import workloads
    X, y = workloads.c5_inputs(N)
    desc = se_noise_desc(X)
    th = np.array([1.3, 0.9, 1.1, 0.8, 0.02])
    ref = g3.Context(0)
    ctx = g3.Context(0)
    try:
        ref.set_data(X)
        want = ref.gp_logp_grad(desc, cabi.KIND_GAUSS, y, th[None], want_grad=True)
        Xs = np.random.default_rng(3).uniform(0, N ** (1.0 / 3.0), size=(333, 3))
        wp = ref.gp_posterior(desc, Xs, y, th, noise=False)
        wpn = ref.gp_posterior(desc, Xs, y, th, noise=True)
        ctx.set_data(X)
        f = ctx.dist_factor(desc, th, nb, 1, 1)
        s = ctx.dist_solve(y)
        assert f["info"] == 0
        assert abs(f["logdet"] - want["logdet"][0]) <= 1e-11 * abs(want["logdet"][0])
        assert abs(s["beta"] - want["beta"][0]) <= 1e-11 * abs(want["beta"][0])
        m, v = ctx.dist_posterior(Xs, noise=False)
        assert scaled_err(m, wp["mean"]) < 1e-10 and scaled_err(v, wp["var"]) < 1e-10
        m2, v2 = ctx.dist_posterior(Xs, noise=True)
        assert scaled_err(m2, wpn["mean"]) < 1e-10 and scaled_err(v2, wpn["var"]) < 1e-10
        g = ctx.dist_grad(desc.n_theta, cfac=1.0)
        assert scaled_err(g["dtheta"], want["dtheta"][0]) < 1e-10
        assert scaled_err(g["ddelta"], want["ddelta"][0]) < 1e-10
        with pytest.raises(cabi.G3Error):
            ctx.dist_grad(desc.n_theta)                          # the factor was consumed
    finally:
        ref.close()
        ctx.close()


def test_block_cyclic_cholesky_detects_indefinite_matrix():
    """*info = 1-based index of the first non-positive pivot (no ladder on the distributed path): K = -exp(-d) shifted by
    tt_to_cov to a 1e-6 diagonal is indefinite at its second pivot; a pivot failure deep inside (block 3 of 4) is located too."""
    from g3py_b200.dist import se_noise_desc
    X = np.random.default_rng(0).uniform(0, 3, size=(1024, 2))
    ctx = g3.Context(0)
    try:
        ctx.set_data(X)
        f = ctx.dist_factor(se_noise_desc(X), np.array([-1.0, 1.0, 1.0, 0.0]), 256, 1, 1)
        assert f["info"] == 2
        Xw = np.random.default_rng(1).uniform(0, 300, size=(1024, 2))     # nearly diagonal K: well conditioned ...
        Xw[700:] = Xw[:324]                                                # ... but rows 700.. repeat rows 0..: singular from 701 on
        ctx.set_data(Xw)
        f = ctx.dist_factor(se_noise_desc(Xw), np.array([1.0, 1.0, 1.0, -1e-9]), 256, 1, 1)
        assert 701 <= f["info"] <= 704
    finally:
        ctx.close()


@pytest.mark.parametrize("world,grid", [(2, "1x2"), (2, "2x1"), (4, "2x2"), (4, "1x4"), (4, "4x1"), (8, "2x4")])
def test_block_cyclic_cholesky_multi_gpu(world, grid):
    """The same checks on `world` GPUs (one process per GPU, NCCL inside libg3b.so, file rendezvous): every rank
    compares its pieces, u, beta and log-det with NumPy's Cholesky and runs the residual probe.  Needs `world` GPUs."""
    import os
    import subprocess
    import sys
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(root, "tools", "dist_chol.py"), "4096", "256", "--grid", grid,
           "--check", "--verify", "--reps", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("check ok") == world
    import json
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert max(res["residual"]) < 1e-12 and res["max_abs_err_vs_numpy"] < 1e-10


def test_c3_full_size_posterior():
    """BASELINE config 3 at full size: warped GP, periodic x SE kernel, N=2048, posterior mean/variance on 10k
    test points, against the oracle (Cholesky route 1e-9, the reference's LU route 1e-7)."""
    x, y, xs = orc.c3_inputs(2048, 10000)
    gp = build_process(SPECS["C3"], x)
    gp.observed(x, y)
    op = orc.OracleProcess(SPECS["C3"], 1)
    rng = np.random.default_rng(21)
    th = _theta0("C3", gp, op, x, y, rng, 1)[0]
    lo = op.logp(th, x, y)
    assert abs(gp.logp(th, array=True) - lo) <= TOL * abs(lo)
    assert scaled_err(gp.dlogp(th, array=True), op.dlogp(th, x, y)) < TOL
    for noise in (False, True):
        post, _, _ = gp._posterior(th, xs, noise=noise)
        pc = op.posterior(th, xs, x, y, noise=noise, solver="chol")
        assert scaled_err(post["location"], pc["location"]) < TOL
        assert scaled_err(post["kernel_diag"], pc["kernel_diag"]) < 1e-8
    pl = op.posterior(th, xs, x, y, noise=False, solver="lu")
    post, _, _ = gp._posterior(th, xs, noise=False)
    assert scaled_err(post["location"], pl["location"]) < 1e-7
    out = gp.predict(th, space=xs, array=True, var=True, median=True, quantiles=True)
    pr = op.predict(th, xs, x, y)
    for k in ("mean", "variance", "median", "quantile_up", "quantile_down"):
        assert scaled_err(out[k], pr[k]) < 1e-7, k
    # 8 chains in lockstep (one batched launch per leapfrog step), finite and moving
    chain, lp, acc = gp.sample_hmc(start=th, samples=3, chains=8, step=0.01, n_leapfrog=3, seed=0)
    assert np.all(np.isfinite(lp)) and chain.shape == (3, 8, gp.ndim)


def test_c4_full_size_properties(ctx):
    """BASELINE config 4: StudentTProcess, N=16384, D=5 (generic Gram interpreter path, D > 4).  The oracle
    needs minutes at this size, so the checks are size-independent: K alpha = delta from the returned
    d logp / d delta, a directional finite difference of the batched logp, and predictive moments at
    training points."""
    X, y, Xs = orc.c4_inputs(16384, 512)
    gp = build_process(SPECS["C4"], X)
    gp.observed(X, y)
    th = gp.dict_to_array(gp.params_default)
    lay = [n for n, s, _ in gp.layout for _ in range(s)]
    th[lay.index("TP_Freedom_degree")] = np.log(5.0)
    th[lay.index("TP_Noise_var")] = np.log(0.05)
    lp, g, info = gp.logp_dlogp_batch(th[None])
    assert info["status"][0] == 0 and np.isfinite(lp[0]) and np.all(np.isfinite(g))
    nat = gp.natural(th)
    nu, beta, n = info["nu"][0], info["beta"][0], len(y)
    c = (nu + n) / (nu - 2.0 + beta)
    delta = y - nat[0]
    res = gp.ctx.gp_logp_grad(gp.desc, cabi.KIND_STUDENT, delta, gp._kernel_theta(nat[None]), nu=np.array([nu]))
    alpha = -res["ddelta"][0] / c
    Ka = np.zeros(n)
    thk = gp._kernel_theta(nat[None])
    for r0 in range(0, n, 2048):                                # K alpha through the Gram entry, row blocks
        Kb, _ = gp.ctx.gram(gp.desc, X[r0:r0 + 2048], X, thk)
        Ka[r0:r0 + 2048] = Kb[0] @ alpha
    Ka += nat[lay.index("TP_Noise_var")] * alpha                # cross-form Gram carries no Noise term
    assert np.max(np.abs(Ka - delta)) <= 1e-9 * np.max(np.abs(delta))
    assert abs(alpha @ delta - beta) <= 1e-9 * beta
    rng = np.random.default_rng(4)
    v = rng.standard_normal(len(th))
    v /= np.linalg.norm(v)
    h = 1e-4
    lpp = gp.logp_batch(np.stack([th + h * v, th - h * v]))
    fd = (lpp[0] - lpp[1]) / (2 * h)
    assert abs(fd - g[0].dot(v)) <= 1e-6 * max(1.0, abs(fd))
    out = gp.predict(th, space=X[:512], array=True, var=True, noise=False)
    resid = out["mean"] - y[:512]
    assert np.sqrt(np.mean(resid ** 2)) < 0.3 and np.all(out["variance"] >= 0) and np.all(out["variance"] < nat[1] * 5)


def test_jitter_ladder_inside_grouped_batch():
    """A 40-item batch (4 stream groups) in which every third item has a numerically singular K (noise 1e-30):
    those run the ladder, the others must be untouched; every item equals its single-item evaluation."""
    X, y, _ = orc.c2_inputs(384, 1)
    X[192:] = X[:192]                                      # duplicated inputs: K is singular without noise
    spec = {"kind": "gauss", "location": {"type": "Zero"}, "kernel": {"type": "SE"}}
    gp = build_process(spec, X)
    gp.observed(X, y)
    op = orc.OracleProcess(spec, 3)
    rng = np.random.default_rng(8)
    B = 40
    Th = np.tile(np.array([0.0, np.log(0.4), np.log(0.4), np.log(0.4), np.log(0.05)]), (B, 1)) + 0.05 * rng.standard_normal((B, 5))
    Th[::3, 4] = np.log(1e-30)
    lp, g, info = gp._eval_batch(Th)                         # log-likelihood part (the 1e-30 noise is behind the prior barrier)
    assert np.all(gp.logprior_batch(Th)[::3] == -np.inf)
    jit = (info["status"] & cabi.ST_JITTER) != 0
    assert np.array_equal(jit, np.arange(B) % 3 == 0)
    for b in (0, 1, 2, 3, 38, 39):
        l1, g1, i1 = gp._eval_batch(Th[b:b + 1])
        # a single item takes the blocked schedule, the batch the left-looking one: same result up to rounding, which the
        # jittered items (K singular up to the 1e-6 shift, condition number ~1e7) amplify
        tol_l, tol_g = (1e-7, 1e-5) if jit[b] else (1e-12, 1e-10)
        assert abs(l1[0] - lp[b]) <= tol_l * abs(lp[b]) and scaled_err(g1[0], g[b]) < tol_g, (b, l1[0], lp[b])
        assert i1["status"][0] == info["status"][b]
        t = op.logp_terms(Th[b], X, y)
        assert (t["info"] > 0) == bool(jit[b])
        if not jit[b]:
            assert abs(lp[b] - t["loglike"]) <= TOL * abs(t["loglike"])
        else:
            assert ((info["status"][b] >> 8) & 0xFF) == t["info"]
            assert abs(info["logdet"][b] - t["logdet"]) <= 1e-6 * abs(t["logdet"])


def test_posterior_under_jitter_ladder():
    """A theta whose K needs the jitter ladder (periodic x SE with little noise): the posterior path must use the
    same repaired factor as the oracle's cholesky_robust."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as mg
    spec, N = mg.SPECS["C3"]
    x, y, xs = mg.data("C3", N)
    gp = build_process(spec, x)
    gp.observed(x, y)
    op = orc.OracleProcess(spec, 1)
    th = mg.theta("C3", op, x, y)
    names = [n for n, s_, _ in op.layout() for _ in range(s_)]
    th[names.index("SIN_rate")] = np.log(0.1)
    th[names.index("Noise_var")] = np.log(0.05 * np.var(y))
    t = op.logp_terms(th, x, y)
    assert t["info"] > 0
    ll, g, info = gp._eval_batch(th[None])
    assert ((info["status"][0] >> 8) & 0xFF) == t["info"]
    assert abs(info["logdet"][0] - t["logdet"]) <= 1e-7 * abs(t["logdet"])
    assert abs(info["beta"][0] - t["beta"]) <= 1e-6 * abs(t["beta"])
    for noise in (False, True):
        post, _, _ = gp._posterior(th, xs, noise=noise)
        po = op.posterior(th, xs, x, y, noise=noise, solver="chol")
        assert scaled_err(post["location"], po["location"]) < 1e-6, noise
        assert scaled_err(post["kernel_diag"], po["kernel_diag"]) < 1e-5, noise


def test_cabi_error_convention(ctx):
    """Bad arguments -> negative return code + message (never a crash); numerical trouble -> status bits only."""
    import ctypes as C
    lib = cabi.load()
    X = np.random.default_rng(0).uniform(0, 1, size=(20, 2))
    gp = g3.GP(X, g3.Zero(), g3.SE(X))
    desc = gp.desc
    th = np.ones((1, desc.n_theta))
    # malformed descriptor: binary node without two operands
    bad = cabi.KernelDesc()
    bad.n_nodes, bad.n_theta = 1, 0
    bad.nodes[0].op = cabi.K_SUM
    with pytest.raises(g3.G3Error, match="malformed|child"):
        ctx.gram(bad, X, None, np.zeros((1, 0)))
    # leaf whose dims exceed D
    bad2 = cabi.KernelDesc()
    bad2.n_nodes, bad2.n_theta = 1, 4
    n0 = bad2.nodes[0]
    n0.op, n0.dim0, n0.dim1, n0.var_idx, n0.p0_idx, n0.p1_idx = cabi.K_SE, 0, 3, 0, 1, -1
    with pytest.raises(g3.G3Error, match="dims"):
        ctx.gram(bad2, X, None, np.ones((1, 4)))
    # theta index outside theta
    bad3 = cabi.KernelDesc()
    bad3.n_nodes, bad3.n_theta = 1, 2
    n0 = bad3.nodes[0]
    n0.op, n0.dim0, n0.dim1, n0.var_idx, n0.p0_idx, n0.p1_idx = cabi.K_SE, 0, 2, 0, 1, -1
    with pytest.raises(g3.G3Error, match="outside theta"):
        ctx.gram(bad3, X, None, np.ones((1, 2)))
    # logp before set_data on a fresh context, wrong delta length
    c2 = g3.Context(0)
    with pytest.raises(g3.G3Error, match="g3_set_data"):
        c2.gp_upload(desc, 0, np.zeros(20), th)
    c2.set_data(X)
    with pytest.raises(ValueError):
        c2.gp_logp_grad(desc, 0, np.zeros(19), th)
    rc = lib.g3_gp_run(c2._h)                                   # nothing uploaded
    assert rc < 0 and b"nothing uploaded" in lib.g3_last_error(c2._h)
    # numerical trouble is not an error: NaN theta -> status bits, finite return code
    r = c2.gp_logp_grad(desc, 0, np.zeros(20), np.full((1, desc.n_theta), np.nan))
    assert r["status"][0] != 0
    c2.close()
    with pytest.raises(g3.G3Error):
        g3.Context(99)                                          # no such device


@pytest.mark.gpu
def test_blocked_schedules_agree(ctx):
    """Few large matrices factor (and invert) by outer blocks of tile columns, batches column by column
    (g3_set_potrf_block; 0 = chosen from B and N).  Both schedules must give the same logp / gradient, and the
    automatic choice must match them, for a single matrix and for a small batch."""
    X, y, Th = orc.c2_inputs(1500, 3)                       # T = 12 tile columns
    gp = build_process(SPECS["C2"], X)
    gp.observed(X, y)
    op = orc.OracleProcess(SPECS["C2"], 3)
    want_lp = np.array([op.logp(t, X, y) for t in Th])
    want_g = np.array([op.dlogp(t, X, y) for t in Th])
    try:
        for w, look, sk, pipe in ((1 << 20, 1, 1, 1), (1 << 20, 1, 0, 1), (4, 1, 1, 1), (4, 1, 1, 0), (4, 0, 1, 1),
                                  (4, 1, 0, 1), (5, 1, 1, 1), (1, 1, 1, 1), (2, 1, 1, 1), (2, 1, 1, 0), (0, 1, 1, 1),
                                  (0, 1, 0, 0)):
            gp.ctx.set_potrf_block(w)
            gp.ctx.set_lookahead(look)
            gp.ctx.set_splitk(sk)
            gp.ctx.set_trtri_pipeline(pipe)
            for sel in (slice(0, 1), slice(0, 3)):
                lp, g, info = gp.logp_dlogp_batch(Th[sel])
                assert np.all(info["status"] == 0)
                assert np.max(np.abs(lp - want_lp[sel]) / np.abs(want_lp[sel])) < TOL, (w, look, sk, pipe)
                assert scaled_err(g, want_g[sel]) < TOL, (w, look, sk, pipe)
                lp2, g2, _ = gp.logp_dlogp_batch(Th[sel])          # split-K adds partial tiles in a fixed order
                assert np.array_equal(lp, lp2) and np.array_equal(g, g2), (w, look, sk, pipe)
    finally:
        gp.ctx.set_potrf_block(0)
        gp.ctx.set_lookahead(1)
        gp.ctx.set_splitk(1)
        gp.ctx.set_trtri_pipeline(1)


@pytest.mark.gpu
def test_pure_c_client_matches_ctypes_path(tmp_path, ctx):
    """The boundary used from plain C (tests/cabi/client.c, no Python in the process) gives the numbers the ctypes
    binding gives for the same descriptor, data and hyper samples."""
    import subprocess
    from test_host import _build_c_client
    exe = _build_c_client(tmp_path)
    N, D = 300, 2
    r = subprocess.run([str(exe), str(N)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [l.split() for l in r.stdout.strip().splitlines()]
    i = np.arange(N)
    X = np.stack([np.fmod(0.37 * i, 5.0), np.fmod(0.91 * i + 0.5, 3.0)], axis=1)
    y = np.sin(X[:, 0]) + 0.3 * np.cos(2.0 * X[:, 1])
    gp = g3.GP(X, g3.Zero(), g3.SE(X))
    gp.observed(X, y)
    theta = np.array([[1.0, 0.8, 1.3, 0.05], [0.6, 1.1, 0.7, 0.1]])
    res = gp.ctx.gp_logp_grad(gp.desc, cabi.KIND_GAUSS, y, theta)
    op = orc.OracleProcess({"kind": "gauss", "location": {"type": "Zero"}, "kernel": {"type": "SE"}}, D)
    for b, row in enumerate(rows):
        assert int(row[0]) == b and int(row[1]) == 0
        vals = np.array([float(v) for v in row[2:]])
        want = np.concatenate([[res["beta"][b], res["logdet"][b]], res["dtheta"][b], [res["ddelta"][b][N // 2]]])
        assert scaled_err(vals, want) < 1e-13
        t = op.logp_terms(np.log(theta[b]), X, y)
        assert abs(vals[0] - t["beta"]) <= 1e-9 * t["beta"] and abs(vals[1] - t["logdet"]) <= 1e-9 * abs(t["logdet"])


def test_c4_full_size_against_oracle_golden(ctx):
    """BASELINE config 4 at its real size (StudentTProcess, SE ARD + noise, N = 16384, D = 5): logp, the gradient and
    predictive moments at 64 of the test points against the committed oracle golden
    (tests/golden/oracle_c4_16384.json, generator tests/golden/make_c4_fullsize_golden.py).  1e-9 on logp, gradient,
    location and the Student-t scaling; the raw `k** - |V|^2` is held to 1e-7 (Cholesky here, LU in the reference /
    oracle, elliptical.py:78-92)."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_c4_16384.json")) as f:
        gold = json.load(f)
    X, y, Xs = orc.c4_inputs(gold["N"], gold["M_all"])
    Xs = Xs[np.array(gold["space_index"])]
    gp = build_process(SPECS["C4"], X)
    gp.observed(X, y)
    assert [(a.split("_", 1)[1], b, c) for a, b, c in gp.layout] == [tuple(v) for v in gold["layout"]]
    th = np.array(gold["theta"])
    lp, g, info = gp.logp_dlogp_batch(th[None])
    assert info["status"][0] == 0
    assert abs(lp[0] - gold["logp"]) <= TOL * abs(gold["logp"])
    assert abs(info["beta"][0] - gold["beta"]) <= TOL * abs(gold["beta"])
    assert abs(info["logdet"][0] - gold["logdet"]) <= TOL * abs(gold["logdet"])
    assert scaled_err(g[0], np.array(gold["dlogp"])) <= TOL
    out = gp.predict(th, space=Xs, array=True, var=True, quantiles=True, noise=False)
    assert scaled_err(out["mean"], np.array(gold["mean"])) <= TOL
    assert scaled_err(out["variance"], np.array(gold["variance"])) <= 1e-7
    assert scaled_err(out["quantile_up"], np.array(gold["quantile_up"])) <= 1e-7
    assert scaled_err(gp.location(th, space=Xs, array=True), np.array(gold["location"])) <= TOL


@pytest.mark.parametrize("N,B", [(333, 3), (1100, 2), (700, 12)])
def test_gradient_resumed_from_resident_factor(ctx, N, B):
    """g3_gp_grad_resume: the gradient finished from the factor a logp-only call left on the device equals the fused
    logp+gradient call (single matrices, blocked and batched left-looking schedules), and is refused once anything
    else ran on the context."""
    X, y, Theta = orc.c2_inputs(N, B)
    gp = build_process(SPECS["C2"], X)
    gp.observed(X, y)
    nat = gp.natural(Theta)
    delta, _, _, _ = gp._host_terms(nat, X, y, False)
    delta = np.array(delta)
    thk = gp._kernel_theta(nat)
    c = gp.ctx
    full = c.gp_logp_grad(gp.desc, cabi.KIND_GAUSS, delta, thk, want_grad=True)
    lo = c.gp_logp_grad(gp.desc, cabi.KIND_GAUSS, delta, thk, want_grad=False)
    assert c.resident_matches(gp.desc, cabi.KIND_GAUSS, delta, thk, None)
    assert not c.resident_matches(gp.desc, cabi.KIND_GAUSS, delta, thk * 1.0000001, None)
    dth, ddl = c.gp_grad_resume()
    assert np.array_equal(lo["beta"], full["beta"]) and np.array_equal(lo["logdet"], full["logdet"])
    assert scaled_err(dth, full["dtheta"]) < 1e-12 and scaled_err(ddl, full["ddelta"]) < 1e-12
    with pytest.raises(cabi.G3Error):
        c.gp_grad_resume()                                   # consumed: K^-1 overwrote L
    c.gp_logp_grad(gp.desc, cabi.KIND_GAUSS, delta, thk, want_grad=False)
    c.gram(gp.desc, X[:10], None, thk[:1])                  # Gram entry does not touch the gp workspaces ...
    assert c.resident_matches(gp.desc, cabi.KIND_GAUSS, delta, thk, None)
    c.gp_posterior(gp.desc, X[:5], np.atleast_2d(delta)[0], thk[0])      # ... the posterior refactors for ONE theta
    assert not c.resident_matches(gp.desc, cabi.KIND_GAUSS, delta, thk, None)


@pytest.mark.parametrize("N,B,ladder", [(600, 1, False), (1100, 2, False), (300, 1, True), (520, 3, True)])
def test_gradient_resumed_with_speculative_inverse(ctx, N, B, ladder):
    """g3_set_speculate_grad: a logp-only evaluation that also pipelines U = L^-T behind its factorisation (what
    GPLogpOp.perform asks for) followed by g3_gp_grad_resume gives the fused gradient - several times in a row (CUDA-graph
    replay of the speculative sequence) and also when the jitter ladder refactors after the speculative pass (duplicated
    inputs without noise: the U of the first, failed pass must not be used)."""
    X, y, Theta = orc.c2_inputs(N, B)
    spec = SPECS["C2"]
    if ladder:
        X = X.copy()
        X[N // 2:] = X[:N - N // 2]
        spec = {"kind": "gauss", "location": {"type": "Zero"}, "kernel": {"type": "SE"}, "noisy": False}
    gp = build_process(spec, X)
    gp.observed(X, y)
    Theta = np.tile(gp.dict_to_array(gp.params_default), (B, 1)) + 0.01 * np.arange(B)[:, None] if ladder else Theta
    nat = gp.natural(Theta)
    delta, _, _, _ = gp._host_terms(nat, X, y, False)
    delta = np.array(delta)
    thk = gp._kernel_theta(nat)
    c = gp.ctx
    full = c.gp_logp_grad(gp.desc, cabi.KIND_GAUSS, delta, thk, want_grad=True)
    if ladder:
        assert np.all(full["status"] & cabi.ST_JITTER)
    c.set_speculate_grad(1)
    try:
        for rep in range(4):                                 # plain run, capture, replays
            lo = c.gp_logp_grad(gp.desc, cabi.KIND_GAUSS, delta, thk, want_grad=False)
            dth, ddl = c.gp_grad_resume()
            assert np.array_equal(lo["status"], full["status"])
            assert scaled_err(lo["logdet"], full["logdet"]) < 1e-13 and scaled_err(lo["beta"], full["beta"]) < 1e-12
            tol = 1e-6 if ladder else 1e-12                  # singular up to the jitter: the two schedules differ by conditioning
            assert scaled_err(dth, full["dtheta"]) < tol and scaled_err(ddl, full["ddelta"]) < tol, rep
    finally:
        c.set_speculate_grad(0)


@pytest.mark.parametrize("N,B", [(700, 1), (1100, 3)])
def test_scheduling_knobs_do_not_change_results(ctx, N, B):
    """The latency switches of round 2 (clustered few-tile GEMM launches, one-launch triangular solves, the two diagonal-tile
    kernels, CUDA-graph replay) only reschedule the same arithmetic: logp and the gradient agree to rounding whatever they
    are set to, and the defaults are restored."""
    X, y, Theta = orc.c2_inputs(N, B)
    gp = build_process(SPECS["C2"], X)
    gp.observed(X, y)
    c = gp.ctx
    ref_lp, ref_g, _ = gp.logp_dlogp_batch(Theta)
    for setter, vals in ((c.set_tile_split, (0, 1)), (c.set_trsv_fused, (0, 1)), (c.set_diag_variant, (1, 2)),
                         (c.set_graphs, (0, 1))):
        try:
            for v in vals:
                setter(v)
                for _ in range(3):                               # plain run, graph capture, replay
                    lp, g, info = gp.logp_dlogp_batch(Theta)
                    assert np.all(info["status"] == 0)
                    assert scaled_err(lp, ref_lp) < 1e-11 and scaled_err(g, ref_g) < 1e-9, (setter.__name__, v)
        finally:
            setter(vals[-1])                                     # the default of every switch is the last value tried


def test_threads_get_their_own_context():
    """Contexts are per thread (a g3_ctx is not re-entrant and ctypes releases the GIL): four threads evaluating
    different hyper samples on the SAME process object concurrently give exactly the sequential results
    (the reference analogue is emcee's `threads > 1`, g3py/bayesian/average.py:29,36)."""
    import threading
    X, y, Theta = orc.c2_inputs(500, 24)
    gp = build_process(SPECS["C2"], X)
    gp.observed(X, y)
    want_lp, want_g, _ = gp.logp_dlogp_batch(Theta)
    out, ctxs, errs = {}, {}, []

    def work(t):
        try:
            rows = Theta[t::4]
            acc = []
            for _ in range(3):                               # repeated: keeps the four contexts busy at the same time
                acc.append(gp.logp_dlogp_batch(rows)[:2])
            out[t] = acc
            ctxs[t] = id(g3.processes.get_context(0))
        except Exception as e:                               # pragma: no cover
            errs.append(e)
    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    assert len(set(ctxs.values())) == 4                      # one context per thread
    for t in range(4):
        for lp, g in out[t]:
            assert scaled_err(lp, want_lp[t::4]) < 1e-11      # 6-item batches take the blocked schedule, 24 the batched one
            assert scaled_err(g, want_g[t::4]) < 1e-9
        assert np.array_equal(out[t][0][0], out[t][2][0]) and np.array_equal(out[t][0][1], out[t][2][1])   # repeatable


@pytest.mark.parametrize("N,B,min_k", [(1100, 10, 256), (2048, 12, 256), (1920, 9, 512)])
def test_ozaki_int8_panel_updates_match_dmma(N, B, min_k):
    """g3_set_gemm_mode(G3_GEMM_OZAKI): the deep block-column updates of the batched Cholesky run as exact int8 slice
    products on the tcgen05 tensor cores (csrc/ozaki.cu).  Factor element-wise, beta, log-det against the DMMA mode at
    1e-13, logp / gradient against the oracle at 1e-9; odd tile counts exercise the 128-wide last block column."""
    X, y, Theta = orc.c2_inputs(N, B)
    gp = build_process(SPECS["C2"], X)
    gp.observed(X, y)
    ctx = g3.Context(0)
    try:
        ctx.set_jitter(gp.consts.jitter, 20)
        ctx.set_data(X)
        nat = gp.natural(Theta)
        delta, _, _, _ = gp._host_terms(nat, X, y, False)
        delta = np.array(delta)
        thk = gp._kernel_theta(nat)
        Np = (N + 127) // 128 * 128
        out = {}
        for mode in ("dmma", "ozaki"):
            ctx.set_gemm_mode(mode, min_k)
            n0 = ctx.ozaki_launch_count()
            r = ctx.gp_logp_grad(gp.desc, cabi.KIND_GAUSS, delta, thk, want_grad=False)
            L = np.tril(ctx.debug_read("gp_A", (B, Np, Np)))
            g = ctx.gp_logp_grad(gp.desc, cabi.KIND_GAUSS, delta, thk, want_grad=True)
            out[mode] = (r, L, g, ctx.ozaki_launch_count() - n0)
        (r0, L0, g0, k0), (r1, L1, g1, k1) = out["dmma"], out["ozaki"]
        assert k0 == 0 and k1 > 0                                # the int8 kernel ran in ozaki mode only
        assert np.all(r1["status"] == 0)
        assert np.max(np.abs(L1 - L0)) <= 1e-13 * np.max(np.abs(L0))
        assert scaled_err(r1["beta"], r0["beta"]) < 1e-12 and scaled_err(r1["logdet"], r0["logdet"]) < 1e-13
        assert scaled_err(g1["dtheta"], g0["dtheta"]) < 1e-11 and scaled_err(g1["ddelta"], g0["ddelta"]) < 1e-11
        op = orc.OracleProcess(SPECS["C2"], X.shape[1])
        t = op.logp_terms(Theta[0], X, y)
        assert abs(r1["beta"][0] - t["beta"]) <= TOL * abs(t["beta"]) and abs(r1["logdet"][0] - t["logdet"]) <= TOL * abs(t["logdet"])
    finally:
        ctx.close()
