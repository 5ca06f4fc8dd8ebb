"""Random kernel expression trees (sums, products, scalar scale / shift over all supported leaves with random column
slices): the post-order descriptor compiled by the product algebra must evaluate to the oracle's Kernel.cov and
dK/dtheta - on the CPU through the NumPy test double, on the GPU through gram_fwd / gram_vjp (generic interpreter and
additive fast path)."""
import numpy as np
import pytest

import g3py_b200 as g3
from oracle import g3_oracle as orc
from helpers import build_kernel, scaled_err
from fake_ctx import _eval_desc

LEAVES = ["SE", "OU", "MAT32", "MAT52", "RQ", "SIN", "WN", "COS", "SINC", "SM"]


def random_spec(rng, D, depth, names):
    if depth == 0 or rng.random() < 0.35:
        t = LEAVES[rng.integers(len(LEAVES))]
        lo = int(rng.integers(0, D))
        hi = int(rng.integers(lo + 1, D + 1))
        n = "%s%d" % (t, len(names))
        names.append(n)
        return {"type": t, "name": n, "dims": [lo, hi]}
    r = rng.random()
    if r < 0.45:
        return {"type": "sum", "k1": random_spec(rng, D, depth - 1, names), "k2": random_spec(rng, D, depth - 1, names)}
    if r < 0.8:
        return {"type": "prod", "k1": random_spec(rng, D, depth - 1, names), "k2": random_spec(rng, D, depth - 1, names)}
    if r < 0.9:
        return {"type": "scale", "c": float(np.round(rng.uniform(0.3, 2.0), 3)), "k": random_spec(rng, D, depth - 1, names)}
    return {"type": "shift", "c": float(np.round(rng.uniform(0.1, 1.0), 3)), "k": random_spec(rng, D, depth - 1, names)}


def n_nodes(spec):
    t = spec["type"]
    if t in ("sum", "prod"):
        return 1 + n_nodes(spec["k1"]) + n_nodes(spec["k2"])
    if t in ("scale", "shift"):
        return 1 + n_nodes(spec["k"])
    return 1


def cases(n=24):
    rng = np.random.default_rng(2024)
    out = []
    while len(out) < n:
        D = int(rng.integers(1, 5))
        spec = random_spec(rng, D, 3, [])
        if n_nodes(spec) <= 16:
            out.append((D, spec))
    return out


def compile_both(D, spec, rng):
    X = rng.uniform(0, 2, size=(37, D))
    ok = orc.build_kernel(spec, D)
    k = build_kernel(spec, X)
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
    assert [(h.name, h.size) for h in ok.layout()] == [(v.name, v.size) for v in reg.vars]
    th_o = np.exp(rng.normal(0.0, 0.3, size=off))
    th_p = np.ones(max(desc.n_theta, 1))
    for h, o, size, const in b.slots:
        th_p[o:o + size] = const if h is None else th_o[h.offset:h.offset + size]
    return X, ok, desc, th_o, th_p[:desc.n_theta], b.slots


# second family: the non-stationary leaves (dot-product, Brownian, constant) and the max node mixed in
LEAVES2 = LEAVES + ["KernelDot", "LIN", "POL", "BW", "VAR"] * 2


def random_spec2(rng, D, depth, names):
    if depth == 0 or rng.random() < 0.35:
        t = LEAVES2[rng.integers(len(LEAVES2))]
        lo = int(rng.integers(0, D))
        hi = int(rng.integers(lo + 1, D + 1))
        n = "%s%d" % (t, len(names))
        names.append(n)
        sp = {"type": t, "name": n, "dims": [lo, hi]}
        if t == "POL":
            sp["p"] = int(rng.integers(1, 5))
        return sp
    r = rng.random()
    kids = lambda: (random_spec2(rng, D, depth - 1, names), random_spec2(rng, D, depth - 1, names))
    if r < 0.35:
        a, b = kids()
        return {"type": "sum", "k1": a, "k2": b}
    if r < 0.6:
        a, b = kids()
        return {"type": "prod", "k1": a, "k2": b}
    if r < 0.8:
        a, b = kids()
        return {"type": "max", "k1": a, "k2": b}
    if r < 0.9:
        return {"type": "scale", "c": float(np.round(rng.uniform(0.3, 2.0), 3)), "k": random_spec2(rng, D, depth - 1, names)}
    return {"type": "shift", "c": float(np.round(rng.uniform(0.1, 1.0), 3)), "k": random_spec2(rng, D, depth - 1, names)}


def n_nodes2(spec):
    t = spec["type"]
    if t in ("sum", "prod", "max"):
        return 1 + n_nodes2(spec["k1"]) + n_nodes2(spec["k2"])
    if t in ("scale", "shift"):
        return 1 + n_nodes2(spec["k"])
    return 1


def cases2(n=16):
    rng = np.random.default_rng(4048)
    out = []
    while len(out) < n:
        D = int(rng.integers(1, 6))
        spec = random_spec2(rng, D, 3, [])
        if n_nodes2(spec) <= 16:
            out.append((D, spec))
    return out


# third family: trees beyond the first ABI limits (17..32 nodes, up to 64 hyper slots): generic interpreter only
def cases3(n=6):
    rng = np.random.default_rng(8096)
    out = []
    while len(out) < n:
        D = int(rng.integers(2, 6))
        names = []
        spec = random_spec2(rng, D, 2, names)
        for _ in range(int(rng.integers(5, 9))):        # a long left-nested sum / product chain keeps the stack shallow
            spec = {"type": "sum" if rng.random() < 0.7 else "prod", "k1": spec, "k2": random_spec2(rng, D, 1, names)}
        if 17 <= n_nodes2(spec) <= 32:
            out.append((D, spec))
    return out


def _case(idx):
    if idx < 24:
        return cases()[idx]
    return cases2()[idx - 24] if idx < 40 else cases3()[idx - 40]


@pytest.mark.parametrize("idx", range(46))
def test_descriptor_matches_oracle_cpu(idx):
    D, spec = _case(idx)
    rng = np.random.default_rng(idx)
    X, ok, desc, th_o, th_p, slots = compile_both(D, spec, rng)
    X2 = rng.uniform(0, 2, size=(11, D))
    for x2, same in ((X, True), (X2, False)):
        K, dK = _eval_desc(desc, th_p, X, x2, same, grad=True)
        assert scaled_err(K, orc.tt_to_num(ok.cov(th_o, X, x2, same))) < 1e-13
        want = ok.dcov(th_o, X, x2, same)
        got = [None] * len(want)
        for h, o, size, const in slots:
            if h is not None:
                for q in range(size):
                    got[h.offset + q] = dK.get(o + q, np.zeros_like(K))
        for a, w in zip(got, want):
            assert scaled_err(a, w) < 1e-12 or np.max(np.abs(w)) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(46))
def test_descriptor_matches_oracle_gpu(idx):
    D, spec = _case(idx)
    rng = np.random.default_rng(idx)
    X, ok, desc, th_o, th_p, slots = compile_both(D, spec, rng)
    rng2 = np.random.default_rng(1000 + idx)
    X = rng2.uniform(0, 2, size=(150, D))
    X2 = rng2.uniform(0, 2, size=(131, D))
    X2[:40] = X[:40]
    ctx = g3.processes.get_context(0)
    for x2, same in ((None, True), (X2, False)):
        xb = X if same else X2
        K, st = ctx.gram(desc, X, x2, th_p[None])
        assert scaled_err(K[0], orc.tt_to_num(ok.cov(th_o, X, xb, same))) < 1e-12
        W = rng2.standard_normal((150, xb.shape[0]))
        g = ctx.gram_vjp(desc, X, x2, th_p[None], W[None])[0]
        want = np.array([np.sum(W * d) for d in ok.dcov(th_o, X, xb, same)])
        got = np.zeros_like(want)
        for h, o, size, const in slots:
            if h is not None:
                got[h.offset:h.offset + size] = g[o:o + size]
        assert scaled_err(got, want) < 1e-10
