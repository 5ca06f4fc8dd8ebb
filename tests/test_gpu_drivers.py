"""Batched drivers (sample_hypers, find_MAP_multistart, sample_hmc) on the CUDA context against the same
drivers run on the oracle-backed NumPy test double (tests/fake_ctx.py), same seeds.  The reference analogues are the
emcee loop of g3py/bayesian/average.py:20-54 and the start-point scoring of g3py/processes/stochastic.py:606-613; there
the proposals are evaluated one logp call at a time, here one batched launch per half-ensemble / leapfrog step.

The ensemble sampler's positions depend on the log-densities only through accept / reject decisions, so the two
trajectories must be IDENTICAL (bit for bit) and the stored log-densities agree to 1e-9; HMC positions integrate the
gradients, so they agree to the gradient tolerance."""
import numpy as np
import pytest

import g3py_b200 as g3
from g3py_b200 import workloads
from fake_ctx import FakeContext

pytestmark = pytest.mark.gpu


def _both(monkeypatch, build, run):
    """run(process) with the NumPy double, then with the CUDA context."""
    fake = FakeContext()
    with monkeypatch.context() as m:
        m.setattr(g3.processes, "get_context", lambda device=0: fake)
        want = run(build())
    real = g3.processes.get_context(0)
    l0 = real.launch_count()
    got = run(build())
    assert real.launch_count() > l0                      # the CUDA path did the work
    return want, got


def _c1(stride=2):
    x, y = workloads.c1_inputs()
    x, y = x[::stride], y[::stride]

    def build():
        gp = g3.GP(x, g3.Bias(), g3.SE(x))
        gp.observed(x, y)
        return gp
    return build


def _c2(n=200):
    X, y, _ = workloads.c2_inputs(n, 1)

    def build():
        gp = g3.GP(X, g3.Bias(), g3.SE(X) + g3.MAT52(X))
        gp.observed(X, y)
        return gp
    return build


@pytest.mark.parametrize("build", [_c1(), _c2()], ids=["C1_N100", "C2_N200"])
def test_sample_hypers_trajectory_matches_oracle_backed_run(monkeypatch, build):
    run = lambda gp: gp.sample_hypers(samples=12, chains=2 * gp.ndim + 2, seed=7)
    (c0, lp0), (c1, lp1) = _both(monkeypatch, build, run)
    assert np.array_equal(c0, c1)                         # same accept / reject decisions -> identical walkers
    assert np.all(np.isfinite(lp1))
    assert np.max(np.abs(lp1 - lp0) / np.maximum(np.abs(lp0), 1.0)) < 1e-9
    assert len(np.unique(c1[:, 0, 0])) > 3                # the walkers moved


def test_sample_hmc_trajectory_matches_oracle_backed_run(monkeypatch):
    build = _c1()
    start = build().find_MAP()

    def run(gp):
        return gp.sample_hmc(start=start, samples=8, chains=6, step=0.04, n_leapfrog=6, seed=3)
    (c0, lp0, a0), (c1, lp1, a1) = _both(monkeypatch, build, run)
    assert np.array_equal(a0, a1)                         # same acceptances
    assert np.max(np.abs(c1 - c0)) < 1e-7 * max(1.0, np.max(np.abs(c0)))
    assert np.max(np.abs(lp1 - lp0) / np.maximum(np.abs(lp0), 1.0)) < 1e-8
    assert np.all(a1 > 0.3)


def test_find_map_multistart_matches_oracle_backed_run(monkeypatch):
    build = _c1()
    gp0 = build()
    rng = np.random.default_rng(11)
    base = gp0.dict_to_array(gp0.params_default)
    starts = [base + 0.3 * rng.standard_normal(len(base)) for _ in range(6)] + [gp0.dict_to_array(gp0.params_test)]

    def run(gp):
        best = gp.find_MAP_multistart(starts)
        th = gp.dict_to_array(best)
        return th, gp.logp(th, array=True), gp.logp_batch(np.array(starts))
    (t0, v0, s0), (t1, v1, s1) = _both(monkeypatch, build, run)
    assert np.max(np.abs(s1 - s0) / np.maximum(np.abs(s0), 1.0)) < 1e-9      # the batched scoring of the starts
    assert abs(v1 - v0) < 1e-6 * max(1.0, abs(v0))                            # same optimum ...
    assert np.max(np.abs(t1 - t0)) < 1e-3                                     # ... reached at the same place (BFGS tolerance)
    assert v1 >= np.max(s1) - 1e-9
