#!/usr/bin/env python
"""bench.py — BASELINE metric 1: fp64 GP logp+grad evaluations/s at N=4096, B=64 theta (config 2:
SE + Matern-5/2 ARD on D=3 synthetic data, Bias mean, white noise).

One step = one pass of the hot path over one batch of 64 hyper samples (Gram build, Cholesky,
solves, K^-1, gradient contraction) per GPU.  `value` is timed with the inputs resident in HBM
(g3_gp_upload before the timed region, g3_gp_run inside it, CUDA events on the library's stream);
`e2e` goes through the public API (NumPy in / NumPy out: `process.logp_dlogp_batch` on one GPU,
`sharding.logp_dlogp_batch_sharded` - all-gather of the results over NCCL included - on several) and
includes host work and both copies.

Multi-GPU (one process per GPU, launched by torchrun; PyTorch-free inside: the NCCL communicator lives
in libg3b.so, the id is handed over through g3py_b200/comm.py): the theta batch shards naturally - every
rank evaluates its own 64 samples (weak scaling).  Beside the headline the N > 1 line carries
  strong   : ONE 64-sample batch split over the N GPUs (all-gather included)
  config3  : BASELINE config 3 - posterior moments on 10 000 test points split over the GPUs, and MCMC chains one per GPU
  metric3  : BASELINE config 5 - exact GP N=131072, 2-D block-cyclic Cholesky + solve, with the 1-GPU run of the same
             code measured in the same job (parallel efficiency) and an on-hardware residual check of the factor.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE JSON line (rank 0).  Libraries write banners to the C-level stdout (NCCL prints
# "NCCL version ..." at communicator creation), so file descriptor 1 is pointed at stderr for the whole run and the
# result line goes to the saved, real stdout.
_REAL_STDOUT = None


def _claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    if _REAL_STDOUT is None:
        print(json.dumps(line), flush=True)
    else:
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


N_OBS, D_IN, B_THETA = 4096, 3, 64
METRIC = "fp64 GP logp+grad evals/s at N=4096 x64 theta batch"
UNIT = "evals/s"


def config_of(N, B):
    """The SAME dict on both arms (ours / reference): the workload, not how an arm executes it."""
    return {"workload": "BASELINE config 2: GP Bias + SE+MAT52 ARD + noise, N=%d D=%d, B=%d theta per GPU, logp+grad" % (N, D_IN, B),
            "N": N, "D": D_IN, "B": B,
            "parallelism": "theta batch sharded over the ranks (each rank its own B rows), results all-gathered",
            "l2": "inputs larger than L2 (working set %.1f GiB per GPU)" % (3 * B * N * N * 8 / 2 ** 30)}


def peaks(ctx):
    """Roofline denominators.  fp64: measured IN THIS RUN on this GPU (g3_debug_fp64_peak: ~0.5 s of back-to-back
    DMMA.8x8x4 / DFMA loops, CUDA events) - MEASURED_PEAKS.json has no fp64 entry.  HBM: the driver-written
    MEASURED_PEAKS.json copy bandwidth (fallback of B200_PROFILING.md if absent); the in-run copy figure is printed
    beside it."""
    m = ctx.fp64_peak(0.5, copy=True)
    out = {"fp64_tflops": m["dmma_tflops"], "dfma_tflops": m["dfma_tflops"], "copy_gbs_in_run": m["copy_gbs"],
           "fp64_source": "in-run: g3_debug_fp64_peak, sustained DMMA.8x8x4 loop on this GPU (DFMA pipe: %.2f TFLOP/s)" % m["dfma_tflops"]}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            mp = json.load(f)
        out["hbm_gbs"] = mp.get("hbm_gbs")
        out["hbm_source"] = "MEASURED_PEAKS.json (driver-written copy bandwidth); in-run copy: %.0f GB/s" % m["copy_gbs"]
    except Exception:
        out["hbm_gbs"] = 6650.0
        out["hbm_source"] = "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent); in-run copy: %.0f GB/s" % m["copy_gbs"]
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_process(X, y, device):
    import g3py_b200 as g3
    gp = g3.GP(X, g3.Bias(), g3.SE(X) + g3.MAT52(X), device=device)
    gp.observed(X, y)
    return gp


def cpu_reference_sample(n_evals, theta_seed=2):
    """The reference schedule on the host cores (oracle port of g3py's Theano/LAPACK path): N x N x D
    broadcast Gram, dpotrf, triangular solve, Murray reverse-mode Cholesky gradient.  n_evals theta rows
    of the same workload; returns (evals/s, seconds)."""
    from oracle import g3_oracle as orc
    try:        # torchrun exports OMP_NUM_THREADS=1: give the CPU baseline every host core back
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    X, y, Theta = orc.c2_inputs(N_OBS, B_THETA)
    spec = {"kind": "gauss", "location": {"type": "Bias"},
            "kernel": {"type": "sum", "k1": {"type": "SE"}, "k2": {"type": "MAT52"}}}
    op = orc.OracleProcess(spec, D_IN)
    t0 = time.perf_counter()
    for b in range(n_evals):
        op.logp(Theta[b], X, y)                       # the reference compiles logp and dlogp as two functions,
        op.dlogp(Theta[b], X, y, method="murray")     # dlogp recomputes the forward (stochastic.py:300-313)
    dt = time.perf_counter() - t0
    return n_evals / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count()
    rows = 1                                            # bounded sample: theta rows of the 64-row batch per step
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_sample(1)
    secs = 0.0
    for _ in range(args.steps):
        v, dt = cpu_reference_sample(rows)
        secs += dt
    value = rows * args.steps / secs
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(N_OBS, B_THETA),
            "step_is": "a bounded sample of the workload: %d of the 64 theta rows per step (the CPU needs ~9 s per row), so "
                       "ms_per_step is the time of %d evaluation(s); `value` is evaluations per second either way" % (rows, rows),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d theta row(s) of the 64 per step; oracle port of the reference schedule "
                                       "(NxNxD broadcast gram, dpotrf, Murray reverse-mode gradient), NumPy/SciPy "
                                       "OpenBLAS on all host cores; Theano itself is not installable here" % rows},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def single_eval_latency():
    """One theta per call through the public API (NumPy in / out), the pattern of one NUTS / BFGS chain: wall time of
    `logp` and of the fused `logp_dlogp` at N = 200 (BASELINE config 1), 1024, 2048, 4096 (config 2 inputs), device default
    settings (CUDA-graph replay, tile split, one-launch solves)."""
    import g3py_b200 as g3
    from g3py_b200 import workloads
    out = {}
    for n_obs in (200, 1024, 2048, 4096):
        x, y = workloads.c1_inputs() if n_obs == 200 else workloads.c2_inputs(n_obs, 1)[:2]
        gp = g3.GP(x, g3.Bias(), g3.SE(x))
        gp.observed(x, y)
        th = gp.dict_to_array(gp.params_default)
        reps = 40 if n_obs <= 2048 else 12
        res = []
        for fn in (lambda: gp.logp(th, array=True), lambda: gp.logp_dlogp(th)):
            for _ in range(4):
                fn()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            res.append(1e6 * (time.perf_counter() - t0) / reps)
        out[str(n_obs)] = {"logp_us": round(res[0], 1), "logp_dlogp_us": round(res[1], 1),
                           "tflops": round(n_obs ** 3 / res[1] / 1e6, 3)}
    out["note"] = "wall clock around the public call, host work and copies included; flops = N^3 per logp+grad"
    return out


def op_path_rate(gp, Theta, steps):
    """Value-and-gradient through the Theano Op boundary (GPLogpOp.perform followed by GPLogpGradOp.perform on the
    same inputs, one theta per call - what PyMC3's NUTS / find_MAP drive): evaluations per second, and how many of the
    gradients were finished from the resident factor (g3_gp_grad_resume) instead of refactoring."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_theano import make_module, Var       # minimal theano.gof.Op / Apply stand-in (Theano is not installable)
    from g3py_b200 import theano_ops, _cabi as cabi
    ops = theano_ops.build_ops(make_module())
    lop = ops.GPLogpOp(gp.desc, cabi.KIND_GAUSS, gp.device)
    nat = gp.natural(Theta)
    thk = gp._kernel_theta(nat)
    X, y = gp.inputs, gp.outputs
    ctx = gp.ctx
    r0 = getattr(ctx, "op_resumed", 0)

    def one(b):
        delta = y - nat[b, 0]
        core, beta, logdet = lop(Var(X), Var(delta), Var(thk[b]), Var(3.0))
        v = core.eval()
        gX, gdelta, gtheta, gnu = lop.grad(core.owner.inputs, [Var(1.0)])
        return v, gtheta.eval()
    one(0)
    ctx.sync()
    t0 = time.perf_counter()
    n = 0
    for s in range(steps):
        for b in range(4):
            one((4 * s + b) % len(Theta))
            n += 1
    ctx.sync()
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "ms_per_eval": 1e3 * dt / n, "evals": n,
            "resumed_from_resident_factor": int(getattr(ctx, "op_resumed", 0) - r0 - 1),
            "note": "GPLogpOp.perform + GPLogpGradOp.perform per theta through a theano.gof.Op stand-in; NumPy in/out"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-metric2", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip strong / config3 / op-path figures")
    ap.add_argument("--dist-n", type=int, default=131072, help="N of the multi-GPU exact-GP Cholesky (runs when --gpus > 1)")
    ap.add_argument("--groups", type=int, default=4, help="batch groups run concurrently on separate streams")
    ap.add_argument("--n", type=int, default=N_OBS, help=argparse.SUPPRESS)
    ap.add_argument("--b", type=int, default=B_THETA, help=argparse.SUPPRESS)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if max(args.warmup, 0) < 3:
        args.warmup = 3

    from g3py_b200 import workloads, comm, sharding
    from g3py_b200.processes import get_context
    ctx = get_context(local)
    comm.init(ctx, rank, world)                           # NCCL communicator inside libg3b.so (no-op for one rank)

    def allmax(v):
        return float(ctx.comm_allreduce([v], "max")[0])

    N, B = args.n, args.b
    X, y, Theta = workloads.c2_inputs(N, B, theta_seed=2 + rank)      # each rank: its own 64 hyper samples
    gp = build_process(X, y, local)
    assert gp.ctx is ctx
    ctx.set_groups(args.groups)
    nat = gp.natural(Theta)
    delta, det_m, _, _ = gp._host_terms(nat, X, y, False)
    thk = gp._kernel_theta(nat)
    P_k = thk.shape[1]

    # ---- device-resident timing --------------------------------------------------------------
    ctx.gp_upload(gp.desc, 0, delta, thk, want_grad=True)
    for _ in range(args.warmup):
        ctx.gp_run()
    ctx.sync()
    chk = ctx.gp_download()
    assert np.all(chk["status"] == 0) and np.all(np.isfinite(chk["dtheta"])), "bench inputs must factor cleanly"
    sampler = ClockSampler(local)
    ctx.comm_barrier()
    l0 = ctx.launch_count()
    sampler.start()
    ctx.timer_begin()
    for _ in range(args.steps):
        ctx.gp_run()
    ms = ctx.timer_end()
    ctx.comm_barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - l0
    # per-kernel-class device times: same step on ONE stream (the concurrent batch groups of the timed region
    # overlap kernels of different classes, so per-launch event pairs are only meaningful serialised)
    ctx.set_groups(1)
    ctx.gp_run()
    ctx.sync()
    ctx.prof_enable(True)
    prof_steps = 2
    ctx.timer_begin()
    for _ in range(prof_steps):
        ctx.gp_run()
    ms_serial = ctx.timer_end() / prof_steps
    prof = ctx.prof_read()
    ctx.prof_enable(False)
    ctx.set_groups(args.groups)
    res = ctx.gp_download()
    assert np.all(res["status"] == 0)
    ms = allmax(ms)
    ms_step = ms / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- the same step with the deep panel updates on the int8 tensor cores (g3_set_gemm_mode(G3_GEMM_OZAKI)) ----------
    gemm_modes = None
    if world == 1 and not args.no_extras:
        try:
            ctx.set_gemm_mode("ozaki", 1024)
            k0 = ctx.ozaki_launch_count()
            ctx.gp_run()
            ctx.sync()
            ctx.timer_begin()
            for _ in range(args.steps):
                ctx.gp_run()
            ms_oz = ctx.timer_end() / args.steps
            ro = ctx.gp_download()
            gemm_modes = {"dmma": {"ms_per_step": ms_step, "value": value},
                          "ozaki_int8": {"ms_per_step": ms_oz, "value": B / (ms_oz * 1e-3), "min_k": 1024,
                                         "int8_update_launches_per_step": (ctx.ozaki_launch_count() - k0) // (args.steps + 1),
                                         "max_rel_diff_vs_dmma": {"logdet": float(np.max(np.abs(ro["logdet"] - res["logdet"]) / np.abs(res["logdet"]))),
                                                                  "dtheta": float(np.max(np.abs(ro["dtheta"] - res["dtheta"])) / np.max(np.abs(res["dtheta"])))},
                                         "status_ok": bool(np.all(ro["status"] == 0))},
                          "default": "dmma",
                          "note": "tcgen05.mma kind::i8 slice products (9 slices of 7 bits, exact int32 accumulation in tensor memory) "
                                  "for the block-column updates of the batched Cholesky with contraction >= 1024; fp64-equivalent "
                                  "results, but at N=4096 x 64 the left-looking block-column update has no A-operand reuse across "
                                  "CTAs and re-reads the slice planes from HBM (45 slice-pair passes), so it only matches the DMMA "
                                  "rate here - DMMA stays the default (DESIGN.md section 6)"}
        except Exception as e:                                 # reported, never hidden
            gemm_modes = {"error": repr(e)}
        finally:
            ctx.set_gemm_mode("dmma")

    # ---- end to end through the public API ----------------------------------------------------
    if world == 1:
        call = lambda: gp.logp_dlogp_batch(Theta)
        e2e_note = ("process.logp_dlogp_batch(Theta): NumPy in/out through ctypes, host O(N) terms included; NumPy arrays "
                    "staged through page-locked buffers inside the library (async DMA both ways)")
    else:
        Theta_all = np.concatenate([workloads.c2_inputs(N, B, theta_seed=2 + r)[2] for r in range(world)])
        call = lambda: sharding.logp_dlogp_batch_sharded(gp, Theta_all)
        e2e_note = ("sharding.logp_dlogp_batch_sharded(process, Theta[%d x P]): every rank evaluates its %d rows, (logp, dlogp) "
                    "all-gathered over NCCL (g3_comm_allgather) - NumPy in/out, host terms and copies included" % (world * B, B))
    for _ in range(2):
        call()
    ctx.comm_barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        call()
    ctx.sync()
    e2e_s = allmax(time.perf_counter() - t0)
    e2e_value = world * B * args.steps / e2e_s
    h2d = 8 * (B * P_k + delta.size)
    d2h = 8 * (2 * B + B * P_k + B * N) + 4 * B

    extras = {}
    if world > 1 and not args.no_extras:
        # strong scaling: ONE 64-sample batch split over the GPUs (B / world rows per GPU: blocked schedule), all-gather included
        for _ in range(2):
            sharding.logp_dlogp_batch_sharded(gp, Theta)
        ctx.comm_barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            sharding.logp_dlogp_batch_sharded(gp, Theta)
        ctx.sync()
        ss = allmax(time.perf_counter() - t0) / args.steps
        extras["strong"] = {"metric": "one %d-theta batch split over %d GPUs, all-gather included" % (B, world), "value": B / ss,
                            "unit": UNIT, "ms_per_batch": 1e3 * ss, "rows_per_gpu": B // world}
    if not args.no_extras and N == N_OBS:
        try:
            extras["config3"] = config3(ctx, local, world, allmax)
        except Exception as e:                                 # reported, never hidden
            extras["config3"] = {"error": repr(e)}
        try:
            if rank == 0:
                extras["op_path"] = op_path_rate(gp, Theta, max(args.steps, 2))
        except Exception as e:
            extras["op_path"] = {"error": repr(e)}
        try:
            if rank == 0 and world == 1:
                extras["single_eval_latency"] = single_eval_latency()
        except Exception as e:
            extras["single_eval_latency"] = {"error": repr(e)}
        ctx.comm_barrier()

    pk = peaks(ctx)
    # exact GP too large for one GPU: block-cyclic Cholesky across the ranks (all ranks take part)
    dist_metric = None
    if world > 1 and not args.no_metric2:
        try:
            dist_metric = metric3(ctx, local, world, args.dist_n, pk)
        except Exception as e:
            dist_metric = {"error": repr(e)}
    if rank != 0:
        ctx.comm_barrier()
        return

    flops_step = float(B) * float(N) ** 3                     # one logp+grad evaluation = N^3 flop
    gemm = prof["dgemm_nt"]
    gemm_ms_step = gemm["ms"] / prof_steps
    achieved = flops_step / (gemm_ms_step * 1e-3) / 1e12 if gemm_ms_step > 0 else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_of(N, B),
        "stream_groups": args.groups,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "note": e2e_note},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["fp64_tflops"], "unit": "TFLOP/s",
                     "frac": (achieved / pk["fp64_tflops"]) if achieved else None,
                     "traffic": None,
                     "traffic_note": "not measured in this run (needs ncu); last capture: 1.10e9 B of dram__bytes_read+write per "
                                     "dgemm_nt launch, mean over the 125 launches of one step = 137.2 GB/step "
                                     "(profiles/r01b_dgemm_dram_per_launch.csv, round-1 kernel, unchanged since)",
                     "kernel": "dgemm_nt_kernel (all level-3 steps of potrf/trtri/lauum; %d launches/step, %.2f ms/step = %.0f%% of the step)"
                               % (gemm["launches"] // prof_steps, gemm_ms_step, 100 * gemm_ms_step / ms_serial),
                     "algorithmic": "B*N^3 flop per step / summed dgemm_nt time per step (CUDA event pairs on the launch stream, single-stream pass of the same step: %.2f ms/step)" % ms_serial,
                     "peak_source": pk["fp64_source"]},
        "step_roofline": {"tflops": flops_step / (ms_step * 1e-3) / 1e12, "frac": flops_step / (ms_step * 1e-3) / 1e12 / pk["fp64_tflops"]},
        "stage_ms_per_step_serial": {k: v["ms"] / prof_steps for k, v in prof.items()},
    }
    gram_ms = prof["gram_fwd"]["ms"] / prof_steps
    if gram_ms > 0:
        gb = B * 4.0 * N * (N + 1)                             # lower-triangle-only variant: 4*N*(N+1) bytes per Gram
        elems = B * 0.5 * N * (N + 1)
        line["gram_roofline"] = {"bound": "hbm", "achieved": gb / (gram_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                 "frac": gb / (gram_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "peak_source": pk["hbm_source"],
                                 "fp64_alu_note": "SE + Matern-5/2 on D = 3 needs ~80 fp64-pipe instructions per element: "
                                                  "%.2f G elements/step -> %.2f ms at the measured DFMA rate (%.1f TFLOP/s); "
                                                  "measured gram_fwd %.2f ms, gram_vjp %.2f ms"
                                                  % (elems / 1e9, elems * 80 * 2 / (pk["dfma_tflops"] * 1e12) * 1e3, pk["dfma_tflops"],
                                                     gram_ms, prof["gram_vjp"]["ms"] / prof_steps)}
    line.update(extras)
    if gemm_modes is not None:
        line["gemm_modes"] = gemm_modes
    if not args.no_metric2:
        try:
            ctx.trim()
            line["metric2"] = metric2(ctx, pk)
        except Exception as e:                                  # reported, never hidden
            line["metric2"] = {"error": str(e)}
    if dist_metric is not None:
        line["metric3"] = dist_metric
    if not args.no_cpu_baseline:
        v, dt = cpu_reference_sample(1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                "sample": "1 of the 64 theta rows (%.1f s); oracle port of the reference schedule (NxNxD "
                                          "broadcast gram, dpotrf, Murray reverse-mode gradient) on all host cores" % dt}
    emit(line)
    ctx.comm_barrier()


def metric2(ctx, pk, N=65536):
    """BASELINE metric 2: fp64 Cholesky TFLOP/s at N=65536 (SE kernel, D=3, lower triangle built on
    the device, factored in place)."""
    from g3py_b200 import workloads
    from g3py_b200.dist import se_noise_desc
    X, y = workloads.c5_inputs(N)
    desc = se_noise_desc(X)
    th = np.array([1.0, 1.0, 1.0, 1.0, 0.01])
    ctx.set_data(X)
    ctx._data_tag = None
    r = ctx.gram_potrf_device(desc, th)      # warm-up (allocations, first launches)
    r = ctx.gram_potrf_device(desc, th)
    fl = float(N) ** 3 / 3.0
    return {"metric": "fp64 Cholesky TFLOP/s at N=%d" % N, "value": fl / (r["ms_potrf"] * 1e-3) / 1e12, "unit": "TFLOP/s",
            "ms_potrf": r["ms_potrf"], "ms_gram": r["ms_gram"], "info": r["info"], "logdet": r["logdet"],
            "frac_of_fp64_peak": fl / (r["ms_potrf"] * 1e-3) / 1e12 / pk["fp64_tflops"],
            "gram_gbs": 4.0 * N * (N + 1) / (r["ms_gram"] * 1e-3) / 1e9}


def metric3(ctx, local, world, N, pk):
    """BASELINE config 5: exact GP N=131072 (SE, D=3), 2-D block-cyclic Cholesky + solve over the ranks.  The 1-GPU run
    of the SAME code (1 x 1 grid, no communicator) is measured in this job on every GPU at once for the efficiency
    baseline; the distributed factor is verified on the hardware (residual probe) and against the 1-GPU beta / log-det."""
    from g3py_b200 import comm
    from g3py_b200._cabi import Context
    from g3py_b200.dist import run_dist_cholesky
    ctx.trim()
    solo = Context(local)                                 # no communicator: world = 1 on every GPU
    try:
        run_dist_cholesky(solo, 8192, nb=1024)            # warm-up
        base = run_dist_cholesky(solo, N, nb=1024)
    finally:
        solo.close()
    ctx.comm_barrier()
    grid = comm.grid_for(world)
    run_dist_cholesky(ctx, 16384, nb=1024, grid=grid)     # warm-up: NCCL channels, allocations
    r = run_dist_cholesky(ctx, N, nb=1024, grid=grid, verify=4)
    alt = None
    if grid[0] > 1:                                       # the 1 x G layout of the same code, for comparison
        ctx.dist_free()
        alt = run_dist_cholesky(ctx, N, nb=1024, grid=(1, world))
    ctx.dist_free()
    t1 = float(ctx.comm_allreduce([base["ms_potrf"]], "max")[0])
    rel_beta = abs(r["beta"] - base["beta"]) / abs(base["beta"])
    rel_ld = abs(r["logdet"] - base["logdet"]) / abs(base["logdet"])
    verified = bool(max(r["residual"]) <= 1e-12 and rel_beta <= 1e-10 and rel_ld <= 1e-10 and r["info"] == 0)
    out = {"metric": "exact-GP Cholesky N=%d, 2-D block-cyclic (nb=1024, %dx%d grid, look-ahead, NCCL inside libg3b.so)" % (N, grid[0], grid[1]),
           "value": r["tflops"], "unit": "TFLOP/s", "ms_potrf": r["ms_potrf"], "ms_gram": r["ms_gram"], "ms_solve": r["ms_solve"],
           "n_gpus": world, "grid": list(grid), "per_gpu_frac_of_fp64_peak": r["tflops"] / world / pk["fp64_tflops"],
           "one_gpu_same_code": {"ms_potrf": t1, "tflops": float(N) ** 3 / 3 / (t1 * 1e-3) / 1e12, "ms_solve": base["ms_solve"],
                                 "measured": "in this run, 1x1 grid on every GPU at once, max over ranks"},
           "parallel_efficiency": t1 / (world * r["ms_potrf"]),
           "solve_speedup_vs_1gpu": base["ms_solve"] / r["ms_solve"],
           "logdet": r["logdet"], "beta": r["beta"], "logp": r["logp"], "info": r["info"], "local_gib": r["local_gib"],
           "residual_max_rel": max(r["residual"]), "rel_diff_beta_vs_1gpu": rel_beta, "rel_diff_logdet_vs_1gpu": rel_ld,
           "verified": verified}
    if alt is not None:
        out["grid_1xG"] = {"ms_potrf": alt["ms_potrf"], "tflops": alt["tflops"], "parallel_efficiency": t1 / (world * alt["ms_potrf"])}
    return out


def config3(ctx, local, world, allmax):
    """BASELINE config 3: warped GP, periodic x SE kernel, N=2048: posterior mean / variance on 10 000 test points with
    the test points split over the GPUs, and HMC chains one per GPU (lock-step leapfrog launches within a rank)."""
    import g3py_b200 as g3
    from g3py_b200 import workloads, sharding
    x, y, xs = workloads.c3_inputs(2048, 10000)
    gp = g3.WGP(x, g3.Bias(), g3.SIN(x) * g3.SE(x), g3.BoxCoxShifted(), device=local)
    gp.observed(x, y)
    th = gp.dict_to_array(gp.params_default)
    lay = [n for n, s, _ in gp.layout for _ in range(s)]
    th[lay.index("WGP_SIN_rate")] = np.log(0.1)
    th[lay.index("WGP_SIN_freq")] = np.log(0.2)
    th[lay.index("WGP_Noise_var")] = np.log(0.05)
    for _ in range(2):
        m, v = sharding.predict_sharded(gp, th, xs)
    ctx.comm_barrier()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        m, v = sharding.predict_sharded(gp, th, xs)
    tp = allmax(time.perf_counter() - t0) / reps
    chains_per_rank = 1 if world > 1 else 8
    sharding.chains_sharded(gp, th, samples=1, chains_per_rank=chains_per_rank, step=0.01, n_leapfrog=2)
    ctx.comm_barrier()
    t0 = time.perf_counter()
    samples, nl = 4, 5
    ch, lp = sharding.chains_sharded(gp, th, samples=samples, chains_per_rank=chains_per_rank, step=0.01, n_leapfrog=nl)
    tc = allmax(time.perf_counter() - t0)
    nchains = ch.shape[1]
    return {"metric": "BASELINE config 3 (warped GP, SINxSE, N=2048)",
            "predict_10k": {"ms": 1e3 * tp, "points_per_s": len(xs) / tp, "finite": bool(np.all(np.isfinite(m)) and np.all(v >= 0)),
                            "note": "sharding.predict_sharded: test points split over %d GPU(s), every rank factors K, results all-gathered" % world},
            "chains": {"chains": int(nchains), "chains_per_gpu": chains_per_rank, "ms_per_leapfrog_step": 1e3 * tc / (samples * nl + 1),
                       "chain_gradient_evals_per_s": nchains * (samples * nl + 1) / tc, "finite": bool(np.all(np.isfinite(lp)))}}


if __name__ == "__main__":
    _claim_stdout()
    main()
