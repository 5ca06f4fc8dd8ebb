"""Committed golden vectors (tests/golden/oracle_c1_c4.json, made by tests/golden/make_golden.py from the oracle on
seeded down-sized BASELINE configs 1-4): the oracle must keep reproducing them (CPU), and the CUDA path must match
them through the public API (GPU)."""
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as mg  # noqa: E402
from oracle import g3_oracle as orc  # noqa: E402
from helpers import build_process, scaled_err  # noqa: E402

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_c1_c4.json")))


@pytest.mark.parametrize("name", list(GOLD))
def test_oracle_reproduces_golden(name):
    rec = GOLD[name]
    X, y, Xs = mg.data(name, rec["N"])
    op = orc.OracleProcess(rec["spec"], X.shape[1])
    th = np.array(rec["theta"])
    assert [list(l) for l in op.layout()] == rec["layout"]
    assert op.logp(th, X, y) == pytest.approx(rec["logp"], rel=1e-11)
    assert scaled_err(op.dlogp(th, X, y), rec["dlogp"]) < 1e-9
    assert scaled_err(rec["dlogp_murray"], rec["dlogp"]) < 1e-7
    po = op.posterior(th, Xs, X, y, noise=True, solver="chol")
    assert scaled_err(po["location"], rec["post_noise1"]["location"]) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(GOLD))
def test_cuda_matches_golden(name):
    rec = GOLD[name]
    X, y, Xs = mg.data(name, rec["N"])
    gp = build_process(rec["spec"], X)
    gp.observed(X, y)
    th = np.array(rec["theta"])
    lp, g, info = gp.logp_dlogp_batch(th[None])
    assert abs(lp[0] - rec["logp"]) <= 1e-9 * abs(rec["logp"])
    assert abs(info["beta"][0] - rec["beta"]) <= 1e-9 * abs(rec["beta"])
    assert abs(info["logdet"][0] - rec["logdet"]) <= 1e-9 * max(abs(rec["logdet"]), 1.0)
    assert scaled_err(g[0], rec["dlogp"]) < 1e-9
    for noise in (False, True):
        out = gp.predict(th, space=Xs, array=True, var=True, quantiles=True, noise=noise)
        post, _, _ = gp._posterior(th, Xs, noise=noise)
        r = rec["post_noise%d" % noise]
        assert scaled_err(post["location"], r["location"]) < 1e-9
        assert scaled_err(post["kernel_diag"], r["kernel_diag"]) < 1e-8
        assert scaled_err(out["mean"], r["mean"]) < 1e-8
        assert scaled_err(out["variance"], r["variance"]) < 1e-8
        assert scaled_err(out["quantile_up"], r["quantile_up"]) < 1e-8
