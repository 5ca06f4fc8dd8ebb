#!/usr/bin/env python
"""bench.py — BASELINE metric 1: fp64 GP logp+grad evaluations/s at N=4096, B=64 theta (config 2:
SE + Matern-5/2 ARD on D=3 synthetic data, Bias mean, white noise).

One step = one pass of the hot path over one batch of 64 hyper samples (Gram build, Cholesky,
solves, K^-1, gradient contraction) per GPU.  `value` is timed with the inputs resident in HBM
(g3_gp_upload before the timed region, g3_gp_run inside it, CUDA events on the library's stream);
`e2e` goes through the public API (`process.logp_dlogp_batch`, NumPy in / NumPy out) and includes
host work and both copies.  Multi-GPU: the theta batch shards naturally - every rank evaluates its
own 64 samples (weak scaling), no data-path collective.

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
        out["hbm_source"] = "MEASURED_PEAKS.json (driver-written copy bandwidth)"
    except Exception:
        out["hbm_gbs"] = 6650.0
        out["hbm_source"] = "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"
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
    per_step = 1                                       # one logp+dlogp evaluation of the 64-batch per step
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_sample(1)
    vals, secs = [], 0.0
    for _ in range(args.steps):
        v, dt = cpu_reference_sample(per_step)
        vals.append(v)
        secs += dt
    value = per_step * args.steps / secs
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE config 2: GP Bias + SE+MAT52 ARD + noise, N=4096 D=3, B=64 theta, logp+grad",
                       "N": N_OBS, "D": D_IN, "B": B_THETA},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d theta row(s) of the 64 per step; oracle port of the reference schedule "
                                       "(NxNxD broadcast gram, dpotrf, Murray reverse-mode gradient), NumPy/SciPy "
                                       "OpenBLAS on all host cores; Theano itself is not installable here" % per_step},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-metric2", action="store_true")
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

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from g3py_b200 import workloads
    from g3py_b200.processes import get_context
    N, B = args.n, args.b
    X, y, Theta = workloads.c2_inputs(N, B, theta_seed=2 + rank)      # each rank: its own 64 hyper samples
    gp = build_process(X, y, local)
    ctx = gp.ctx
    ctx.set_groups(args.groups)
    nat = gp.natural(Theta)
    delta, det_m, _, _ = gp._host_terms(nat, X, y, False)
    thk = gp._kernel_theta(nat)
    P_k = thk.shape[1]

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing --------------------------------------------------------------
    ctx.gp_upload(gp.desc, 0, delta, thk, want_grad=True)
    for _ in range(args.warmup):
        ctx.gp_run()
    ctx.sync()
    chk = ctx.gp_download()
    assert np.all(chk["status"] == 0) and np.all(np.isfinite(chk["dtheta"])), "bench inputs must factor cleanly"
    sampler = ClockSampler(local)
    barrier()
    l0 = ctx.launch_count()
    sampler.start()
    ctx.timer_begin()
    for _ in range(args.steps):
        ctx.gp_run()
    ms = ctx.timer_end()
    barrier()
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
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- end to end through the public API ----------------------------------------------------
    for _ in range(2):
        gp.logp_dlogp_batch(Theta)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lp, g, info = gp.logp_dlogp_batch(Theta)
    ctx.sync()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * B * args.steps / e2e_s
    h2d = 8 * (B * P_k + delta.size)
    d2h = 8 * (2 * B + B * P_k + B * N) + 4 * B

    # exact GP too large for one GPU: block-cyclic Cholesky across the ranks (all ranks take part)
    dist_metric = None
    if world > 1 and not args.no_metric2:
        try:
            from g3py_b200.dist_potrf import run_dist_cholesky
            run_dist_cholesky(16384, nb=1024)                      # warm-up: NCCL channels, allocator
            r = run_dist_cholesky(args.dist_n, nb=1024)
            one_gpu_tflops = 35.59                                  # measured, same code, world=1 (profiles/r01_dist_cholesky_131072.jsonl)
            dist_metric = {"metric": "exact-GP Cholesky N=%d, block-cyclic (nb=1024, 1x%d grid, panel broadcast over NCCL, look-ahead)" % (args.dist_n, world),
                           "value": r["tflops"], "unit": "TFLOP/s", "ms_potrf": r["ms_potrf"], "ms_gram": r["ms_gram"],
                           "n_gpus": world,
                           "parallel_efficiency_vs_1gpu_same_code": r["tflops"] / world / one_gpu_tflops,
                           "logdet": r["logdet"], "info": r["info"], "local_gib": r["local_gib"]}
        except Exception as e:
            dist_metric = {"error": repr(e)}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    pk = peaks(ctx)
    flops_step = float(B) * float(N) ** 3                     # one logp+grad evaluation = N^3 flop
    gemm = prof["dgemm_nt"]
    gemm_ms_step = gemm["ms"] / prof_steps
    achieved = flops_step / (gemm_ms_step * 1e-3) / 1e12 if gemm_ms_step > 0 else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "BASELINE config 2: GP Bias + SE+MAT52 ARD + noise, N=%d D=%d, B=%d theta per GPU, logp+grad" % (N, D_IN, B),
                   "N": N, "D": D_IN, "B": B, "parallelism": "theta-batch sharded, %d rank(s), no collective" % world, "stream_groups": args.groups,
                   "l2": "inputs larger than L2 (working set %.1f GiB per GPU)" % (3 * B * N * N * 8 / 2 ** 30)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": "process.logp_dlogp_batch(Theta): NumPy in/out through ctypes, host O(N) terms included; NumPy arrays staged through page-locked buffers inside the library (async DMA both ways)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["fp64_tflops"], "unit": "TFLOP/s",
                     "frac": (achieved / pk["fp64_tflops"]) if achieved else None,
                     "traffic": 1.0977e9 if (N == N_OBS and B == B_THETA) else None,
                     "traffic_note": "dram__bytes_read+write per dgemm_nt launch, mean over the 125 launches of one step (137.2 GB/step; ncu, profiles/r01b_dgemm_dram_per_launch.csv); the one lauum launch moves 8.83 GB for 8.6 GB of algorithmic operand+result bytes",
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
        line["gram_roofline"] = {"bound": "hbm", "achieved": gb / (gram_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                 "frac": gb / (gram_ms * 1e-3) / 1e9 / pk["hbm_gbs"]}
    if not args.no_metric2:
        try:
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
    if dist is not None:
        dist.destroy_process_group()


def metric2(ctx, pk, N=65536):
    """BASELINE metric 2: fp64 Cholesky TFLOP/s at N=65536 (SE kernel, D=3, lower triangle built on
    the device, factored in place)."""
    import g3py_b200 as g3
    from g3py_b200 import workloads
    X, y = workloads.c5_inputs(N)
    k = g3.SE(X) + g3.KernelNoise(name="Noise")
    reg = g3.Registry()
    k.check_dims(X)
    k.check_hypers("", reg)
    b = g3.DescBuilder(3)
    k.compile(b)
    desc = b.finish()
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


if __name__ == "__main__":
    _claim_stdout()
    main()
