"""Op-protocol tests of g3py_b200.theano_ops through the fake-Theano harness (make_node / perform / grad /
__props__ / pickling).  CPU: device replaced by tests/fake_ctx.py; GPU: the same through libg3b.so."""
import pickle

import numpy as np
import pytest

import g3py_b200 as g3
from g3py_b200 import _cabi as cabi, theano_ops
from oracle import g3_oracle as orc
from fake_ctx import FakeContext
from fake_theano import make_module, Var
from helpers import scaled_err

SPEC = {"kind": "gauss", "location": {"type": "Zero"}, "kernel": {"type": "sum", "k1": {"type": "SE"}, "k2": {"type": "MAT52"}}}


def _setup(kind="gauss"):
    rng = np.random.default_rng(0)
    X = rng.uniform(0, 4, size=(60, 2))
    y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(60)
    cls = g3.GP if kind == "gauss" else g3.TP
    gp = cls(X, g3.Zero(), g3.SE(X) + g3.MAT52(X))
    gp.observed(X, y)
    spec = dict(SPEC, kind=kind)
    op = orc.OracleProcess(spec, 2)
    th = 0.2 * rng.standard_normal(op.P)
    th[6] = np.log(0.3)
    if kind == "student":
        th[-1] = np.log(4.0)
    return gp, op, X, y, th


def _run(kind):
    ops = theano_ops.build_ops(make_module())
    gp, op, X, y, th = _setup(kind)
    nat = gp.natural(th)
    thk = gp._kernel_theta(nat[None])[0]
    nu = 2.0 + nat[-1] if kind == "student" else 3.0
    k = cabi.KIND_STUDENT if kind == "student" else cabi.KIND_GAUSS
    lop = ops.GPLogpOp(gp.desc, k)
    assert lop == ops.GPLogpOp(gp.desc, k) and hash(lop) == hash(ops.GPLogpOp(gp.desc, k))     # __props__ merge
    assert pickle.loads(pickle.dumps(lop)) == lop
    core, beta, logdet = lop(Var(X), Var(y), Var(thk), Var(nu))
    t = op.logp_terms(th, X, y)
    assert beta.eval() == pytest.approx(t["beta"], rel=1e-9) and logdet.eval() == pytest.approx(t["logdet"], rel=1e-9)
    n = len(y)
    if kind == "gauss":
        total = -0.5 * n * orc.Consts().log_2pi + core.eval()
    else:
        total = t["r2"] + core.eval()
    assert total == pytest.approx(t["loglike"], rel=1e-9)
    gX, gdelta, gtheta, gnu = lop.grad(core.owner.inputs, [Var(1.0)])
    assert gX == "disconnected"
    go = op.dlogp(th, X, y)                                  # log-space gradient, bijection order
    g_nat = gtheta.eval()
    g_log = np.empty(gp.ndim)
    for h, off, size, const in gp._slots:
        if h is not None:
            g_log[h.offset:h.offset + size] = g_nat[off:off + size] * nat[h.offset:h.offset + size]
    if kind == "student":
        from scipy.special import digamma
        d_r2 = 0.5 * digamma((nu + n) * 0.5) - 0.5 * digamma(nu * 0.5) - 0.5 * n / (nu - 2.0)
        g_log[-1] = (float(gnu.eval()) + d_r2) * nat[-1]
    assert scaled_err(g_log, go) < 1e-9
    # d core / d delta = -c * alpha: check against a directional difference of the oracle in y
    v = np.random.default_rng(1).standard_normal(n)
    h = 1e-5
    fd = (op.loglike(th, X, y + h * v) - op.loglike(th, X, y - h * v)) / (2 * h)
    assert float(gdelta.eval() @ v) == pytest.approx(fd, rel=1e-6)
    # Gram Op and its VJP Op
    gop = ops.GramOp(gp.desc)
    xv = Var(X)
    K = gop(xv, xv, Var(thk))
    Ko = op.k_noise.cov(nat, X, X, True)
    assert scaled_err(K.eval(), Ko) < 1e-12
    W = np.random.default_rng(2).standard_normal(Ko.shape)
    _, _, gth = gop.grad(K.owner.inputs, [Var(W)])
    dK = op.k_noise.dcov(nat, X, X, True)
    want = np.array([np.sum(W * d) for d in dK])
    got = gth.eval()
    back = np.empty_like(want)
    for h_, off, size, const in gp._slots:
        if h_ is not None:
            back[h_.offset:h_.offset + size] = got[off:off + size]
    assert scaled_err(back, want) < 1e-10
    # posterior Op
    Xs = X[:9] + 0.03
    m, vr = ops.GPPosteriorOp(gp.desc, noise=False)(Var(X), Var(Xs), Var(y), Var(thk))
    po = op.posterior(th, Xs, X, y, noise=False, solver="chol")
    assert scaled_err(m.eval(), po["location"]) < 1e-9 and scaled_err(vr.eval(), po["kernel_diag"]) < 1e-8


@pytest.mark.parametrize("kind", ["gauss", "student"])
def test_ops_cpu_harness(monkeypatch, kind):
    ctx = FakeContext()
    ctx.gram_vjp = lambda desc, x1, x2, theta, w: np.array([[np.sum(np.asarray(w) * g) for i, g in sorted(
        __import__("fake_ctx")._eval_desc(desc, np.asarray(theta), np.asarray(x1), np.asarray(x1 if x2 is None else x2), x2 is None, grad=True)[1].items())]])
    ctx.potrf_robust = lambda A: (np.linalg.cholesky(A), 0, 0.0)
    monkeypatch.setattr(g3.processes, "get_context", lambda device=0: ctx)
    _run(kind)
    # value-then-gradient on the same inputs finished from the resident factor (no second factorisation), and X was
    # uploaded once although every perform() receives it
    assert ctx.resumed >= 1 and ctx.uploads == 1


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["gauss", "student"])
def test_ops_gpu(kind):
    ctx = g3.processes.get_context(0)
    r0 = getattr(ctx, "op_resumed", 0)
    _run(kind)
    assert getattr(ctx, "op_resumed", 0) > r0       # GPLogpGradOp finished from GPLogpOp's resident factor (g3_gp_grad_resume)
    ops = theano_ops.build_ops(make_module())
    rng = np.random.default_rng(3)
    A = rng.standard_normal((150, 160))
    K = A @ A.T
    L = ops.CholeskyRobustGPU()(Var(K)).eval()
    assert scaled_err(L @ L.T, K) < 1e-12
