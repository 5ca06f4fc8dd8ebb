"""CPU tests: the C-ABI library loads and exports every declared symbol, ctypes mirrors the C structs,
the host-side operator algebra / bijection / gradient assembly agree with the oracle (device calls replaced
by the NumPy test double in tests/fake_ctx.py), and the theta-batch sharding works over gloo (world 2)."""
import ctypes
import os
import pickle
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import g3py_b200 as g3
from g3py_b200 import _cabi as cabi, workloads, sharding
from oracle import g3_oracle as orc
from helpers import build_process, scaled_err
from fake_ctx import FakeContext

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "g3b.h")).read()
    declared = set(re.findall(r"\b(g3_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(cabi.lib_path())
    for name in sorted(declared):
        assert hasattr(lib, name), "libg3b.so does not export %s" % name
    assert declared >= set(cabi.SIGNATURES)              # every bound symbol is declared in the header
    cabi.load()


def test_struct_layout_matches_c():
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "g3b.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(g3_knode), sizeof(g3_kernel_desc), offsetof(g3_knode, value), offsetof(g3_kernel_desc, nodes));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out] == [ctypes.sizeof(cabi.KNode), ctypes.sizeof(cabi.KernelDesc), cabi.KNode.value.offset,
                                     cabi.KernelDesc.nodes.offset]


def test_header_is_standalone_c99(tmp_path):
    """include/g3b.h is the C boundary: it must compile on its own as plain C (no C++, no implicit includes)."""
    src = tmp_path / "h.c"
    src.write_text('#include "g3b.h"\nint main(void) { g3_kernel_desc d; d.n_nodes = 0; (void)d; return G3_K_MAX == 20 ? 0 : 1; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_no_cpu_fallback():
    """Without a GPU the product path fails loudly (this container has none)."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is visible")
    except ImportError:
        pass
    with pytest.raises(g3.G3Error, match="no CPU fallback"):
        g3.Context(0)
    x, y = workloads.c1_inputs()
    gp = g3.GP(x, g3.Bias(), g3.SE(x))
    gp.observed(x, y)
    with pytest.raises(g3.G3Error):
        gp.logp()
    for mod in ("oracle", "oracle.g3_oracle"):
        pass
    src = "".join(open(os.path.join(ROOT, "g3py_b200", f)).read() for f in ("processes.py", "_cabi.py", "sharding.py", "workloads.py"))
    assert "oracle" not in src.replace("(bench.py must not import the oracle)", "").replace("tests check they equal the oracle's", "")


def test_oracle_is_imported_only_where_allowed():
    """oracle/ is test infrastructure: nothing under g3py_b200/ or tools/ may import it (AST scan of every module), and
    bench.py only inside its CPU-baseline / reference-arm function."""
    import ast

    def oracle_imports(path):
        tree = ast.parse(open(path).read())
        hits = []
        for fn in ast.walk(tree):
            for node in ast.iter_child_nodes(fn):
                mods = []
                if isinstance(node, ast.Import):
                    mods = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    mods = [node.module or ""]
                if any(m == "oracle" or m.startswith("oracle.") for m in mods):
                    hits.append(getattr(fn, "name", "<module>"))
        return hits

    for base in ("g3py_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith(".py"):
                    assert oracle_imports(os.path.join(dirpath, f)) == [], (dirpath, f)
    where = oracle_imports(os.path.join(ROOT, "bench.py"))
    assert where and all("cpu" in w or "reference" in w for w in where), where
    assert set(oracle_imports(os.path.join(ROOT, "__graft_entry__.py"))) <= {"smoke", "build"}


def test_workloads_equal_oracle_inputs():
    for a, b in ((workloads.c1_inputs(), orc.c1_inputs()), (workloads.c2_inputs(64, 4), orc.c2_inputs(64, 4)),
                 (workloads.c3_inputs(32, 8), orc.c3_inputs(32, 8)), (workloads.c4_inputs(32, 8), orc.c4_inputs(32, 8)),
                 (workloads.c5_inputs(64), orc.c5_inputs(64))):
        for u, v in zip(a, b):
            assert np.array_equal(u, v)


def test_layout_names_and_desc():
    X = np.random.default_rng(0).uniform(0, 1, size=(10, 3))
    gp = g3.GP(X, g3.Bias(), g3.SE(X) + g3.MAT52(X))
    assert [n for n, _, _ in gp.layout] == ["GP_Bias_Bias", "GP_SE_var", "GP_SE_rate", "GP_MAT52_var", "GP_MAT52_rate",
                                            "GP_Noise_var"]
    d = gp.desc
    ops = [d.nodes[i].op for i in range(d.n_nodes)]
    assert ops == [cabi.K_SE, cabi.K_MAT52, cabi.K_SUM, cabi.K_NOISE, cabi.K_SUM] and d.n_theta == 9
    assert d.nodes[3].flags == cabi.KF_PROCESS_NOISE and (d.nodes[4].dim0, d.nodes[4].dim1) == (2, 3)
    # KernelProd fixes the second var to 1.0 (kernels.py:215-219); SIN creates freq before rate
    tp = g3.WTP(X[:, :1], g3.Bias(), g3.SIN(X[:, :1]) * g3.SE(X[:, :1]), g3.ArcsinhLinear())
    assert [n for n, _, _ in tp.layout] == ["WTP_Bias_Bias", "WTP_SIN_var", "WTP_SIN_freq", "WTP_SIN_rate", "WTP_SE_rate",
                                            "WTP_Noise_var", "WTP_ArcsinhLinear_shift", "WTP_ArcsinhLinear_scale",
                                            "WTP_Freedom_degree"]
    se = [tp.desc.nodes[i] for i in range(tp.desc.n_nodes) if tp.desc.nodes[i].op == cabi.K_SE][0]
    assert se.var_idx == -1 and se.value == 1.0
    # scalar algebra and fixed hypers
    k = 2.0 * g3.SE(X, var=3.0) + 0.5
    reg = g3.Registry()
    k.check_hypers("", reg)
    b = g3.DescBuilder(3)
    k.compile(b)
    dd = b.finish()
    assert [dd.nodes[i].op for i in range(dd.n_nodes)] == [cabi.K_SE, cabi.K_SCALE, cabi.K_SHIFT]
    assert [v.name for v in reg.vars] == ["SE_rate"]
    with pytest.raises(ValueError):
        g3.GP(X, g3.Bias(), g3.SE(X) + g3.SE(X))           # duplicate hyper names, as PyMC3 would refuse


def test_bijection_params_and_pickle():
    x, y = workloads.c1_inputs()
    gp = g3.GP(x, g3.Bias(), g3.SE(x))
    assert gp.params_test == {"GP_Bias_Bias": 0.0, "GP_SE_var_log__": 0.0, "GP_SE_rate_log__": pytest.approx([0.0]),
                              "GP_Noise_var_log__": 0.0}
    gp.observed(x, y)
    pd = gp.params_default                                   # Appendix B of SURVEY.md
    assert pd["GP_Bias_Bias"] == pytest.approx(y.mean())
    assert pd["GP_SE_var_log__"] == pytest.approx(np.log(y.var()))
    assert pd["GP_Noise_var_log__"] == pytest.approx(np.log(y.var()))
    assert pd["GP_SE_rate_log__"][0] == pytest.approx(np.log(0.5 / np.abs(np.diff(x[:, 0])).mean()))
    th = gp.dict_to_array(pd)
    assert gp.array_to_dict(th) == pytest.approx(pd)
    assert gp.dict_to_array({"GP_SE_var": 2.0})[1] == pytest.approx(np.log(2.0))     # bare name = natural value
    assert gp.dict_to_array({"GP_SE_var_log_": 0.3})[1] == pytest.approx(0.3)        # pymc3 3.0 spelling
    g2 = pickle.loads(pickle.dumps(gp))
    assert g2.layout == gp.layout and g2.desc.n_nodes == gp.desc.n_nodes and np.array_equal(g2.inputs, gp.inputs)
    assert gp.logp(prior=True) == 0.0
    th_bad = th.copy()
    th_bad[1] = np.log(1e-7)
    assert gp.logp(th_bad, array=True, prior=True) == -np.inf                         # NonTransformLog barrier


SPECS = {
    "gp": {"kind": "gauss", "location": {"type": "Bias"}, "kernel": {"type": "sum", "k1": {"type": "SE"}, "k2": {"type": "MAT52"}}},
    "wgp": {"kind": "gauss", "warped": True, "location": {"type": "Linear"},
            "kernel": {"type": "prod", "k1": {"type": "SIN"}, "k2": {"type": "SE"}}, "mapping": {"type": "BoxCoxShifted"}},
    "wtp": {"kind": "student", "warped": True, "location": {"type": "Bias"},
            "kernel": {"type": "sum", "k1": {"type": "RQ"}, "k2": {"type": "OU"}}, "mapping": {"type": "ArcsinhLinear"}},
    "tp": {"kind": "student", "location": {"type": "Zero"}, "kernel": {"type": "sum", "k1": {"type": "MAT32"}, "k2": {"type": "WN"}},
           "mapping": {"type": "SinhArcsinh"}, "noisy": False},
    "boxlin": {"kind": "gauss", "location": {"type": "Bias"},
               "kernel": {"type": "shift", "c": 0.3, "k": {"type": "scale", "c": 2.0, "k": {"type": "SE"}}},
               "mapping": {"type": "BoxCoxLinear"}},
    "logshift": {"kind": "gauss", "location": {"type": "Bias"}, "kernel": {"type": "SE"}, "mapping": {"type": "LogShifted"}},
}


def _problem(spec, n=40, D=2, seed=0, B=3):
    rng = np.random.default_rng(seed)
    X = rng.uniform(0, 4, size=(n, D))
    y = 1.5 + np.exp(0.4 * np.sin(X[:, 0]) + 0.1 * rng.standard_normal(n))
    op = orc.OracleProcess(spec, D)
    Th = 0.2 * rng.standard_normal((B, op.P))
    off = 0
    for nm, size, pos in op.layout():
        if nm.endswith("SIN_rate"):
            Th[:, off:off + size] = np.log(0.05)
        if nm.endswith("SIN_freq"):
            Th[:, off:off + size] = np.log(0.2)
        if nm.endswith("LogShifted_shift"):
            Th[:, off:off + size] = 0.3
        if nm.endswith("Freedom_degree"):
            Th[:, off:off + size] += np.log(4.0)
        if nm.endswith("Noise_var"):
            Th[:, off:off + size] += np.log(0.3)
        off += size
    return op, X, y, Th


@pytest.fixture
def fake(monkeypatch):
    ctx = FakeContext()
    monkeypatch.setattr(g3.processes, "get_context", lambda device=0: ctx)
    return ctx


@pytest.mark.parametrize("name", list(SPECS))
def test_host_assembly_against_oracle(fake, name):
    """logp / dlogp / predict assembled on the host around the (faked) device call == oracle."""
    spec = SPECS[name]
    op, X, y, Th = _problem(spec)
    gp = build_process(spec, X)
    gp.observed(X, y)
    assert [(a.split("_", 1)[1], b, c) for a, b, c in gp.layout] == op.layout()
    lp, g, info = gp.logp_dlogp_batch(Th)
    for b in range(len(Th)):
        lo = op.logp(Th[b], X, y)
        assert abs(lp[b] - lo) < 1e-10 * abs(lo)
        assert scaled_err(g[b], op.dlogp(Th[b], X, y)) < 1e-9
    assert gp.logp(Th[0], array=True) == pytest.approx(lp[0], rel=1e-13)
    assert gp.loglike(gp.array_to_dict(Th[1])) == pytest.approx(lp[1], rel=1e-13)
    assert np.allclose(gp.logp_chain(Th), lp, rtol=1e-13)
    Xs = X[:15] + 0.05
    for noise in (False, True):
        out = gp.predict(Th[0], space=Xs, array=True, var=True, median=True, quantiles=True, noise=noise)
        pr = op.predict(Th[0], Xs, X, y, noise=noise)
        for k in ("mean", "variance", "std", "median", "quantile_up", "quantile_down"):
            assert scaled_err(out[k], pr[k]) < 1e-9, (k, noise)


def test_strict_vs_exact_constants(fake):
    spec = SPECS["gp"]
    op, X, y, Th = _problem(spec)
    a = build_process(spec, X, strict=True)
    b = build_process(spec, X, strict=False)
    a.observed(X, y)
    b.observed(X, y)
    la, lb = a.logp(Th[0], array=True), b.logp(Th[0], array=True)
    n = len(y)
    assert la - lb == pytest.approx(-0.5 * n * (1.8378770351409912 - np.log(2 * np.pi)), rel=1e-6)
    assert lb == pytest.approx(orc.OracleProcess(spec, 2, strict=False).logp(Th[0], X, y), rel=1e-12)


def test_find_map_on_fake(fake):
    x, y = workloads.c1_inputs()
    x, y = x[::4], y[::4]
    gp = g3.GP(x, g3.Bias(), g3.SE(x))
    gp.observed(x, y)
    p0 = gp.dict_to_array(gp.params_default)
    best = gp.find_MAP()
    assert gp.logp(best) > gp.logp(p0, array=True)
    assert np.linalg.norm(gp.dlogp(best)) < 1e-3


def test_shard_bounds():
    for B in (1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["G3_ROOT"]); sys.path.insert(0, os.path.join(os.environ["G3_ROOT"], "tests"))
import numpy as np, torch, torch.distributed as dist
import g3py_b200 as g3
from g3py_b200 import sharding, workloads
from fake_ctx import FakeContext
ctx = FakeContext()
g3.processes.get_context = lambda device=0: ctx
dist.init_process_group("gloo")
X, y, Theta = workloads.c2_inputs(48, 7)
gp = g3.GP(X, g3.Bias(), g3.SE(X) + g3.MAT52(X)); gp.observed(X, y)
lp, g = sharding.logp_dlogp_batch_sharded(gp, Theta)
calls_sharded = ctx.calls
lp0, g0, _ = gp.logp_dlogp_batch(Theta)
assert np.array_equal(lp, lp0) and np.array_equal(g, g0), (lp, lp0)
lo, hi = sharding.shard_bounds(7, dist.get_rank(), dist.get_world_size())
assert (hi - lo) in (3, 4) and calls_sharded == 1
Xs = X[:9] + 0.05
m, v = sharding.predict_sharded(gp, Theta[0], Xs)
full = gp.predict(Theta[0], space=Xs, array=True, var=True)
assert np.allclose(m, full["mean"], rtol=1e-13) and np.allclose(v, full["variance"], rtol=1e-13, atol=1e-15)
if dist.get_rank() == 0:
    print("SHARD_OK", flush=True)
dist.destroy_process_group()
'''


def test_theta_sharding_gloo_world2():
    env = dict(os.environ, G3_ROOT=ROOT, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "w.py")
        open(f, "w").write(_WORKER)
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                            "--master-addr", "127.0.0.1", "--master-port", "29517", f],
                           capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARD_OK" in r.stdout


_DIST_WORKER = r"""
import os, sys
sys.path.insert(0, os.environ["G3_ROOT"]); sys.path.insert(0, os.path.join(os.environ["G3_ROOT"], "tests"))
import numpy as np
from dist_model import DistModel, LocalComm, GlooComm
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")
    comm = GlooComm()
else:
    comm = LocalComm()
N, nb = 1536, 128
rng = np.random.default_rng(0)
A = rng.standard_normal((N, N + 8)); K = A @ A.T / N + np.eye(N)
Lref = np.linalg.cholesky(K)
dvec = np.sin(np.arange(N) * 0.01)
uref = np.linalg.solve(Lref, dvec)
for Pr in range(1, world + 1):
    if world % Pr:
        continue
    m = DistModel(K, nb, Pr, world // Pr, comm).factor()
    m.check_against(Lref)
    ld = float(comm.allreduce(np.array([m.logdet]))[0])
    assert abs(ld - np.log(np.diag(Lref)).sum()) < 1e-9
    u, beta = m.solve(dvec)
    assert np.abs(u - uref).max() < 1e-10 and abs(beta - uref @ uref) < 1e-9 * (uref @ uref)
    held = sum(v.size for v in m.store.values())
    tot = float(comm.allreduce(np.array([float(held)]))[0])
    assert tot == nb * nb * (N // nb) * (N // nb + 1) / 2            # every block of the lower triangle exactly once
if comm.rank == 0: print("DIST_OK", flush=True)
if world > 1: dist.destroy_process_group()
"""


@pytest.mark.parametrize("world", [1, 2, 4])
def test_block_cyclic_cholesky_schedule(world):
    """The 2-D block-cyclic schedule of csrc/dist.cu (pieces, diagonal-block exchange inside a process column, grouped
    piece broadcasts, per-panel all-reduce of the substitution) restated with NumPy blocks (tests/dist_model.py), one
    process per rank over gloo, on every Pr x Pc grid of the world size."""
    env = dict(os.environ, G3_ROOT=ROOT, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "w.py")
        open(f, "w").write(_DIST_WORKER)
        if world == 1:
            cmd = [sys.executable, f]
        else:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
                   "--master-addr", "127.0.0.1", "--master-port", str(29520 + world), f]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "DIST_OK" in r.stdout


def test_dist_layout_of_the_library_matches_the_model():
    """g3_dist_layout (pure index arithmetic of libg3b.so, no device) against the Python restatement: owner, first
    block, block count and in-piece index for every block of several grids; every block has exactly one owner."""
    from dist_model import first_blk, cnt_blk
    from g3py_b200 import _cabi
    for (nP, Pr, Pc) in [(6, 1, 1), (7, 1, 3), (8, 2, 1), (9, 2, 2), (12, 2, 4), (10, 4, 2), (5, 3, 1)]:
        nb, N = 128, 128 * nP
        seen = {}
        for J in range(nP):
            for I in range(J, nP):
                for p in range(Pr):
                    lay = _cabi.dist_layout(N, nb, Pr, Pc, I, J, p)
                    assert lay["first"] == first_blk(J, p, Pr) and lay["count"] == cnt_blk(J, p, Pr, nP)
                    assert lay["owner"] == (J % Pc) * Pr + I % Pr
                    assert lay["index"] == (I - first_blk(J, I % Pr, Pr)) // Pr
                seen[(I, J)] = lay["owner"]
            assert sum(cnt_blk(J, p, Pr, nP) for p in range(Pr)) == nP - J
        assert len(seen) == nP * (nP + 1) // 2
    with pytest.raises(ValueError):
        _cabi.dist_layout(1000, 128, 1, 1, 0, 0, 0)                  # N not a multiple of nb


def test_comm_rendezvous_file(tmp_path, monkeypatch):
    """g3py_b200.comm: rank 0 publishes the 128-byte NCCL id atomically, the other ranks read the same bytes."""
    from g3py_b200 import comm
    monkeypatch.setenv("G3_RDV_FILE", str(tmp_path / "rdv"))
    monkeypatch.setattr(comm, "_SEQ", [0])
    uid0, path = comm.exchange_id(0, 2)
    monkeypatch.setattr(comm, "_SEQ", [0])
    uid1, _ = comm.exchange_id(1, 2, timeout=5)
    assert uid0 == uid1 and len(uid0) == 128 and os.path.exists(path)
    assert comm.grid_for(8) == (2, 4) and comm.grid_for(4) == (2, 2) and comm.grid_for(2) == (1, 2) and comm.grid_for(1) == (1, 1)


def test_batched_drivers_on_fake(fake):
    """sample_hypers (stretch move on logp_batch) and multi-start MAP: one device call per half-ensemble."""
    x, y = workloads.c1_inputs()
    x, y = x[::5], y[::5]
    gp = g3.GP(x, g3.Bias(), g3.SE(x))
    gp.observed(x, y)
    calls0 = fake.calls
    chain, lp = gp.sample_hypers(samples=5, chains=8, seed=0)
    assert chain.shape == (5, 8, gp.ndim) and lp.shape == (5, 8)
    assert fake.calls - calls0 == 1 + 5 * 2                         # init + two half-ensembles per sweep
    assert np.all(np.isfinite(lp))
    for c in range(8):                                              # stored logp is the logp of the stored position
        assert gp.logp(chain[-1, c], array=True) == pytest.approx(lp[-1, c], rel=1e-12)
    starts = [gp.params_default, gp.params_test]
    best = gp.find_MAP_multistart(starts)
    assert gp.logp(best) >= max(gp.logp(s) for s in starts)


def test_batched_hmc_on_fake(fake):
    x, y = workloads.c1_inputs()
    x, y = x[::5], y[::5]
    gp = g3.GP(x, g3.Bias(), g3.SE(x))
    gp.observed(x, y)
    best = gp.find_MAP()
    c0 = fake.calls
    chain, lp, acc = gp.sample_hmc(start=best, samples=6, chains=4, step=0.05, n_leapfrog=5, seed=1)
    assert fake.calls - c0 == 1 + 6 * 5                              # one batched launch per leapfrog step
    assert chain.shape == (6, 4, gp.ndim) and np.all(np.isfinite(lp)) and np.all(acc > 0.3)
    assert gp.logp(chain[-1, 2], array=True) == pytest.approx(lp[-1, 2], rel=1e-12)


def test_logpredictive_and_sampler_on_fake(fake):
    x, y = workloads.c1_inputs()
    x, y = x[::4], y[::4]
    gp = g3.GP(x, g3.Bias(), g3.SE(x))
    gp.observed(x, y)
    fake.potrf_robust = lambda A: (np.linalg.cholesky(A + 1e-9 * np.eye(len(A))), 0, 0.0)
    th = gp.dict_to_array(gp.params_default)
    xs = x[:7] + 0.1
    v = gp.predict(th, space=xs, array=True, var=True, samples=5, distribution=True, noise=True)
    assert v["samples"].shape == (7, 5)
    # logpredictive = sum of independent normal log-densities with the noisy predictive variance
    from scipy import stats
    want = stats.norm.logpdf(y[:7], loc=v["mean"], scale=np.sqrt(v["variance"])).sum()
    got = v["logpredictive"](y[:7])
    strict_shift = -0.5 * 7 * (1.8378770351409912 - np.log(2 * np.pi))
    assert got == pytest.approx(want + strict_shift, rel=1e-9)


def _build_c_client(tmp_path):
    exe = tmp_path / "cabi_client"
    libdir = os.path.dirname(g3.lib_path())
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "cabi", "client.c"), "-o", str(exe), "-L", libdir, "-lg3b", "-lm",
                        "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_pure_c_client_links_against_the_boundary(tmp_path):
    """tests/cabi/client.c uses only include/g3b.h: it must compile as C99 and link against libg3b.so; without a GPU
    it fails loudly at g3_ctx_create (exit code 2), never silently."""
    exe = _build_c_client(tmp_path)
    r = subprocess.run([str(exe), "40"], capture_output=True, text=True)
    assert r.returncode in (0, 2), (r.returncode, r.stderr)
    if r.returncode == 2:
        assert "g3_ctx_create failed" in r.stderr


def test_per_call_inputs_do_not_replace_observations(fake):
    """`logp(inputs=, outputs=)` / `predict(inputs=, ...)` substitute the data for that call only, like the reference's
    lambda_method (g3py/processes/stochastic.py:385-430); the observed data are untouched afterwards."""
    spec = SPECS["gp"]
    op, X, y, Th = _problem(spec)
    gp = build_process(spec, X)
    gp.observed(X, y)
    base = gp.loglike(Th[0], array=True)
    X7, y7 = X[:7] + 0.1, y[:7] * 0.5
    held = gp.loglike(Th[0], inputs=X7, outputs=y7, array=True)
    assert held == pytest.approx(op.logp(Th[0], X7, y7), rel=1e-10)
    assert len(gp.inputs) == len(X) and np.array_equal(gp.inputs, X) and np.array_equal(gp.outputs, y)
    assert gp.loglike(Th[0], array=True) == base                      # device copy followed the restore
    out = gp.predict(Th[0], space=X[:5], inputs=X7, outputs=y7, array=True, var=True)
    pr = op.predict(Th[0], X[:5], X7, y7)
    assert scaled_err(out["mean"], pr["mean"]) < 1e-9
    assert len(gp.inputs) == len(X)
    loc = gp.location(Th[0], space=X[:5], inputs=X7, outputs=y7, array=True)
    assert scaled_err(loc, pr["location"] if "location" in pr else op.posterior(Th[0], X[:5], X7, y7)["location"]) < 1e-9
    assert np.array_equal(gp.outputs, y) and gp.loglike(Th[0], array=True) == base


def test_library_carries_blackwell_native_code():
    """SASS evidence (B200_PROFILING.md): the fp64 GEMM is DMMA fed by TMA, the int8 panel update is tcgen05 (UTCIMMA) with
    tensor-memory loads (LDTM) and TMA - checked on the built libg3b.so with cuobjdump (no GPU needed)."""
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    r = subprocess.run(["cuobjdump", "-sass", g3.lib_path()], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0
    sass = r.stdout
    for mnemonic in ("DMMA.8x8x4", "UTMALDG", "UTCIMMA", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert "HGMMA" not in sass and "sm_100a" in sass
