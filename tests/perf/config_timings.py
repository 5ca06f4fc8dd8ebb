"""Per-config timings of the BASELINE.json configs on one GPU (SURVEY §8d "per-config throughput"), through the public
API (NumPy in / NumPy out, host<->device copies included), with the oracle port timed beside it on a bounded sample.

    python tests/perf/config_timings.py > profiles/r01_config_timings.jsonl

One JSON line per config.  Flop conventions are SURVEY §8d: logp = N^3/3, logp+grad = N^3, posterior variance = N^2 M.
The oracle is used only for the CPU reference numbers (like bench.py's cpu_baseline leg); the script lives under
tests/ because only tests/, smoke() and bench.py may import oracle/.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import g3py_b200 as g3  # noqa: E402
from g3py_b200 import workloads as wl  # noqa: E402


def best(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), float(np.median(ts))


def cpu_time(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    from oracle import g3_oracle as orc
    cores = os.cpu_count()

    # ---- C1: tutorial GP, N=200 -----------------------------------------------------------------
    x, y = wl.c1_inputs()
    gp = g3.GP(x, g3.Bias(), g3.SE(x))
    gp.observed(x, y)
    th = gp.dict_to_array(gp.params_default)
    t_lp = best(lambda: gp.logp(th, array=True), 20, 5)
    t_g = best(lambda: gp.logp_dlogp(th), 20, 5)
    t0 = time.perf_counter()
    gp.executed["logp"] = 0
    gp.find_MAP(start=gp.params_default)
    t_map = time.perf_counter() - t0
    op = orc.OracleProcess({"kind": "gauss", "location": {"type": "Bias"}, "kernel": {"type": "SE"}}, 1)
    c_lp = min(cpu_time(lambda: op.logp(th, x, y)) for _ in range(5))
    c_g = min(cpu_time(lambda: op.dlogp(th, x, y, method="murray")) for _ in range(5))
    emit(config="C1 GP SE+noise N=200 (tutorial)", logp_us=t_lp[0] * 1e6, logp_dlogp_us=t_g[0] * 1e6,
         find_MAP_s=t_map, find_MAP_evals=int(gp.executed["logp"]), cpu_logp_us=c_lp * 1e6,
         cpu_dlogp_murray_us=c_g * 1e6, cpu_cores=cores,
         note="latency-bound: one 128x128 tile; API call = ctypes + 2 small copies + 11 launches")

    # ---- C3: warped GP, periodic x SE, N=2048, 8 chains, M=10k posterior ----------------------------
    x, y, xs = wl.c3_inputs(2048, 10000)
    gp = g3.WGP(x, g3.Bias(), g3.SIN(x) * g3.SE(x), g3.BoxCoxShifted())
    gp.observed(x, y)
    th = gp.dict_to_array(gp.params_default)
    lay = [n for n, s, _ in gp.layout for _ in range(s)]
    th[lay.index("WGP_SIN_rate")] = np.log(0.01)
    th[lay.index("WGP_SIN_freq")] = np.log(0.2)
    th[lay.index("WGP_BoxShift_power")] = np.log(0.7)
    Th = np.tile(th, (8, 1)) + 0.02 * np.random.default_rng(0).standard_normal((8, len(th)))
    N, M = 2048, 10000
    t_b = best(lambda: gp.logp_dlogp_batch(Th), 5, 2)
    t_p = best(lambda: gp.predict(th, space=xs, array=True, var=True), 5, 2)
    spec = {"kind": "gauss", "warped": True, "location": {"type": "Bias"},
            "kernel": {"type": "prod", "k1": {"type": "SIN"}, "k2": {"type": "SE"}}, "mapping": {"type": "BoxCoxShifted"}}
    op = orc.OracleProcess(spec, 1)
    c_g = cpu_time(lambda: op.dlogp(th, x, y, method="murray"))
    c_p = cpu_time(lambda: op.predict(th, xs[:2000], x, y))
    emit(config="C3 WGP SINxSE N=2048, 8 chains, posterior on M=10000",
         chains8_logp_dlogp_ms=t_b[0] * 1e3, chain_evals_per_s=8 / t_b[0], chains_tflops=8 * N ** 3 / t_b[0] / 1e12,
         predict_mean_var_ms=t_p[0] * 1e3, predict_points_per_s=M / t_p[0], predict_tflops=N * N * M / t_p[0] / 1e12,
         cpu_dlogp_murray_s_per_chain=c_g, cpu_predict_s_per_2000_points=c_p, cpu_cores=cores,
         note="reference posterior = LU solve with M right-hand sides + full MxM product (elliptical.py:78-107); "
              "CPU predict timed on 2000 of the 10000 points (the MxM product is 800 MB x2 at full size)")

    # ---- C4: Student-t process N=16384, D=5 ---------------------------------------------------------
    X, y, Xs = wl.c4_inputs(16384, 4096)
    gp = g3.TP(X, g3.Bias(), g3.SE(X))
    gp.observed(X, y)
    th = gp.dict_to_array(gp.params_default)
    lay = [n for n, s, _ in gp.layout for _ in range(s)]
    th[lay.index("TP_Freedom_degree")] = np.log(5.0)
    th[lay.index("TP_Noise_var")] = np.log(0.05)
    N, M = 16384, 4096
    t_l = best(lambda: gp.logp(th, array=True), 3, 1)
    t_g = best(lambda: gp.logp_dlogp(th), 3, 1)
    t_p = best(lambda: gp.variance(th, space=Xs, array=True), 3, 1)
    import scipy.linalg as sla
    K = np.eye(4096) * 2 + 0.1
    c_chol = min(cpu_time(lambda: sla.cholesky(K, lower=True)) for _ in range(2))
    emit(config="C4 TP SE ARD N=16384 D=5, predictive moments on M=4096",
         logp_ms=t_l[0] * 1e3, logp_tflops=N ** 3 / 3 / t_l[0] / 1e12,
         logp_dlogp_ms=t_g[0] * 1e3, logp_dlogp_tflops=N ** 3 / t_g[0] / 1e12,
         predict_var_ms=t_p[0] * 1e3, predict_tflops=(N ** 3 / 3 + N * N * M) / t_p[0] / 1e12,
         cpu_dpotrf_4096_s=c_chol, cpu_dpotrf_16384_extrapolated_s=c_chol * 64, cpu_cores=cores,
         note="single matrix: right-looking outer blocks; predict re-factorises K (stateless API call); "
              "CPU: dpotrf timed at N=4096 and scaled by 4^3 (extrapolation, stated as such)")


if __name__ == "__main__":
    main()
