"""Call latency of the public API at small and mid sizes with CUDA-graph replay on / off (round 2).
Rows: BASELINE config 1 (N=200), one chain at N=1024 / 2048 / 4096, config 3 (8 chains in lockstep, warped GP SINxSE N=2048)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import g3py_b200 as g3
from g3py_b200 import workloads


def timeit(f, n):
    for _ in range(4):
        f()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    return (time.perf_counter() - t0) / n


for graphs in (False, True):
    for N in (200, 1024, 2048, 4096):
        if N == 200:
            x, y = workloads.c1_inputs()
        else:
            x, y, _ = workloads.c2_inputs(N, 1)
        gp = g3.GP(x, g3.Bias(), g3.SE(x)); gp.observed(x, y)
        gp.ctx.set_graphs(graphs)
        th = gp.dict_to_array(gp.params_default)
        n = 60 if N <= 2048 else 15
        r0 = gp.ctx.graph_replays()
        tl = timeit(lambda: gp.logp(th, array=True), n)
        tg = timeit(lambda: gp.logp_dlogp(th), n)
        print("graphs %d  N=%-5d logp %7.0f us   logp+grad (one fused call) %7.0f us = %5.2f TFLOP/s   replays %d"
              % (graphs, N, 1e6 * tl, 1e6 * tg, N ** 3 / tg / 1e12, gp.ctx.graph_replays() - r0), flush=True)
    x, y, xs = workloads.c3_inputs(2048, 10000)
    gp = g3.WGP(x, g3.Bias(), g3.SIN(x) * g3.SE(x), g3.BoxCoxShifted()); gp.observed(x, y)
    gp.ctx.set_graphs(graphs)
    th = gp.dict_to_array(gp.params_default)
    lay = [n_ for n_, s, _ in gp.layout for _ in range(s)]
    th[lay.index("WGP_SIN_rate")] = np.log(0.1); th[lay.index("WGP_SIN_freq")] = np.log(0.2); th[lay.index("WGP_Noise_var")] = np.log(0.05)
    Th = np.tile(th, (8, 1)) + 0.01 * np.random.default_rng(0).standard_normal((8, len(th)))
    t8 = timeit(lambda: gp.logp_dlogp_batch(Th), 20)
    tp = timeit(lambda: gp.predict(th, space=xs, array=True, var=True), 10)
    print("graphs %d  config 3: 8 chains logp+grad %7.0f us = %5.2f TFLOP/s;  predict 10k points %7.0f us" % (graphs, 1e6 * t8, 8 * 2048 ** 3 / t8 / 1e12, 1e6 * tp), flush=True)
    gp.ctx.set_graphs(True)
