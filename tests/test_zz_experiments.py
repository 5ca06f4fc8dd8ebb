"""The stand-alone int8 tensor-core prototypes of experiments/i8gemm ("what comes next"): they check
themselves (bit-exact int8 GEMM against the CPU, exact slicing, fp64-equivalent GEMM against a long-double reference)
and exit non-zero on any mismatch; this test runs them on the GPU and reads their JSON lines.  Named zz so that it runs
after the parity suite of the product."""
import json
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EXP = os.path.join(os.path.dirname(HERE), "experiments", "i8gemm")


def _run(name, *args):
    exe = os.path.join(EXP, name)
    if not os.path.exists(exe):
        pytest.skip("%s not built (python __graft_entry__.py)" % name)
    r = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return [json.loads(line) for line in r.stdout.splitlines() if line.startswith("{")]


def test_sources_are_in_the_build():
    assert os.path.exists(os.path.join(EXP, "Makefile"))
    entry = open(os.path.join(os.path.dirname(HERE), "__graft_entry__.py")).read()
    assert "experiments" in entry and "i8gemm" in entry


@pytest.mark.gpu
def test_int8_tcgen05_gemm_is_exact():
    out = _run("i8gemm", 2048, 1024, 2)
    assert out[0]["check"] == "exact" and out[0]["mismatches"] == 0
    assert out[1]["sampled_mismatches"] == 0 and out[1]["TOPs"] > 0


@pytest.mark.gpu
def test_sliced_int8_gemm_matches_fp64_accuracy():
    out = _run("ozaki_dgemm", 2048, 1)
    slicing = out[0]
    assert slicing["check"] == "slicing" and slicing["max_residual_over_row_scale"] <= slicing["bound_2^-7S"]
    full = out[1]
    assert full["check"] == "ozaki_vs_long_double" and full["max_err_ozaki"] < 2e-15
    for line in out[2:]:
        assert line["bench"] == "ozaki_dgemm_nt" and line["max_err_ozaki"] < 2e-15, line
