"""Generates tests/golden/oracle_c4_16384.json: BASELINE config 4 AT ITS REAL SIZE (StudentTProcess, SE ARD + noise,
N = 16384, D = 5) evaluated by the CPU oracle (oracle/g3_oracle.py, itself pinned to the executed reference up to
N = 4096, tests/golden/reference_fullsize.json: C4_4096).  The executed reference cannot run this size (its N x N x D
metric tensor alone is 10.7 GB under the torch-based stand-in), so the full-size pin is the oracle: logp, the analytic
gradient, and the predictive location / t-scaled variance / 97.5 % quantile at 64 of the config's test points (LU
route, like elliptical.py:78-92).  Takes ~10 minutes and ~40 GB of host memory.

    python tests/golden/make_c4_fullsize_golden.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import g3_oracle as orc  # noqa: E402

SPEC = {"kind": "student", "location": {"type": "Bias"}, "kernel": {"type": "SE"}}
N, M_ALL, M_KEEP = 16384, 4096, 64


def theta_for(op, X, y):
    """Deterministic hypers: g3py's defaults (kernels.py:33-40, metrics.py:104-108, means.py:133-134) with the noise
    floor of SURVEY §8d and nu = 2 + 5."""
    th = np.zeros(op.P)
    off = 0
    for nm, size, pos in op.layout():
        if nm.endswith("Bias_Bias"):
            th[off] = np.mean(y)
        elif nm.endswith("SE_var"):
            th[off] = np.log(np.var(y))
        elif nm.endswith("SE_rate"):
            th[off:off + size] = np.log(0.5 / np.mean(np.abs(X[1:] - X[:-1]), axis=0))
        elif nm.endswith("Noise_var"):
            th[off] = np.log(0.05)
        elif nm.endswith("Freedom_degree"):
            th[off] = np.log(5.0)
        off += size
    return th


def main():
    t0 = time.time()
    X, y, Xs = orc.c4_inputs(N, M_ALL)
    idx = np.linspace(0, M_ALL - 1, M_KEEP).astype(int)
    op = orc.OracleProcess(SPEC, X.shape[1])
    th = theta_for(op, X, y)
    terms = op.logp_terms(th, X, y)
    assert terms["info"] == 0
    print("logp terms %.1f s" % (time.time() - t0), terms["loglike"], flush=True)
    g = op.dlogp(th, X, y)
    print("dlogp %.1f s" % (time.time() - t0), g, flush=True)
    pr = op.predict(th, Xs[idx], X, y, noise=False)
    po = op.posterior(th, Xs[idx], X, y, noise=False)
    print("predict %.1f s" % (time.time() - t0), flush=True)
    rec = {"spec": SPEC, "N": N, "M_all": M_ALL, "space_index": idx.tolist(), "layout": op.layout(), "theta": th.tolist(),
           "logp": op.logprior(th) + terms["loglike"], "beta": terms["beta"], "logdet": terms["logdet"], "nu": terms["nu"],
           "dlogp": g.tolist(), "location": po["location"].tolist(), "kernel_diag": po["kernel_diag"].tolist(),
           "scaling": float(po["scaling"]), "mean": pr["mean"].tolist(), "variance": pr["variance"].tolist(),
           "quantile_up": pr["quantile_up"].tolist(),
           "generator": "tests/golden/make_c4_fullsize_golden.py (CPU oracle, NumPy/SciPy fp64)"}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_c4_16384.json"), "w") as f:
        json.dump(rec, f, indent=0)
    print("done %.1f s" % (time.time() - t0), rec["logp"])


if __name__ == "__main__":
    main()
