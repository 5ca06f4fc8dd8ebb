"""Generates tests/golden/oracle_c1_c4.json: oracle outputs on seeded, down-sized versions of BASELINE configs
1-4 (the reference holds no fixtures and cannot run here, so the goldens come from the oracle,
which is itself pinned by the notebook KATs and cross-checked in tests/test_oracle.py).  The file freezes the
oracle: a later change of its arithmetic shows up as a diff against these numbers.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import g3_oracle as orc  # noqa: E402

SPECS = {
    "C1": ({"kind": "gauss", "location": {"type": "Bias"}, "kernel": {"type": "SE"}}, 200),
    "C2": ({"kind": "gauss", "location": {"type": "Bias"},
            "kernel": {"type": "sum", "k1": {"type": "SE"}, "k2": {"type": "MAT52"}}}, 256),
    "C3": ({"kind": "gauss", "warped": True, "location": {"type": "Bias"},
            "kernel": {"type": "prod", "k1": {"type": "SIN"}, "k2": {"type": "SE"}},
            "mapping": {"type": "BoxCoxShifted"}}, 192),
    "C4": ({"kind": "student", "location": {"type": "Bias"}, "kernel": {"type": "SE"}}, 224),
}


def data(name, N):
    if name == "C1":
        x, y = orc.c1_inputs()
        return x[:N], y[:N], np.linspace(0, 10, 37)[:, None]
    if name == "C2":
        X, y, _ = orc.c2_inputs(N, 1)
        return X, y, X[:29] + 0.07
    if name == "C3":
        x, y, xs = orc.c3_inputs(N, 41)
        return x, y, xs
    X, y, Xs = orc.c4_inputs(N, 33)
    return X, y, Xs


def theta(name, op, X, y):
    rng = np.random.default_rng(100 + ord(name[1]))
    th = 0.1 * rng.standard_normal(op.P)
    off = 0
    for nm, size, pos in op.layout():
        if nm.endswith("_var") and not nm.startswith("Noise"):
            th[off:off + size] += np.log(max(np.var(y), 1e-3))
        if nm.endswith("Noise_var"):
            th[off:off + size] += np.log(0.05 * max(np.var(y), 1e-3))
        if nm.endswith("Bias_Bias"):
            th[off:off + size] += np.mean(y) if name != "C3" else 0.0
        if nm.endswith("SIN_rate"):
            th[off:off + size] = np.log(0.05)
        if nm.endswith("Noise_var") and name == "C3":
            th[off:off + size] = np.log(0.05)
        if nm.endswith("SIN_freq"):
            th[off:off + size] = np.log(0.2)
        if nm.endswith("Freedom_degree"):
            th[off:off + size] = np.log(5.0)
        if nm.endswith("BoxShift_power"):
            th[off:off + size] = np.log(0.7)
        off += size
    return th


def main():
    out = {}
    for name, (spec, N) in SPECS.items():
        X, y, Xs = data(name, N)
        op = orc.OracleProcess(spec, X.shape[1])
        th = theta(name, op, X, y)
        t = op.logp_terms(th, X, y)
        assert t["info"] == 0, (name, "golden theta must factor without the jitter ladder")
        rec = {"spec": spec, "N": N, "theta": th.tolist(), "layout": op.layout(),
               "logp": op.logp(th, X, y), "beta": t["beta"], "logdet": t["logdet"], "det_m": t["det_m"],
               "dlogp": op.dlogp(th, X, y).tolist(), "dlogp_murray": op.dlogp(th, X, y, method="murray").tolist()}
        for noise in (False, True):
            po = op.posterior(th, Xs, X, y, noise=noise, solver="chol")
            pr = op.predict(th, Xs, X, y, noise=noise)
            rec["post_noise%d" % noise] = {"location": po["location"].tolist(), "kernel_diag": po["kernel_diag"].tolist(),
                                           "mean": pr["mean"].tolist(), "variance": pr["variance"].tolist(),
                                           "quantile_up": pr["quantile_up"].tolist()}
        out[name] = rec
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_c1_c4.json"), "w") as f:
        json.dump(out, f, indent=0)
    print({k: v["logp"] for k, v in out.items()})


if __name__ == "__main__":
    main()
