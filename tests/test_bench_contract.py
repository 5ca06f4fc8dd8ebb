"""The bench.py JSON contract, checked on the committed lines of the last GPU runs (profiles/) and on a live run of the
reference arm (CPU).  Guards the keys the driver and the judge read."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "e2e"}


def _load(name):
    txt = open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()
    return json.loads(txt[-1])


@pytest.mark.parametrize("name", ["r01h_bench_final.json", "r01g_bench_2gpu_theta_only.json"])
def test_committed_bench_line_follows_the_contract(name):
    d = _load(name)
    assert BASE <= set(d), BASE - set(d)
    assert d["metric"].startswith("fp64 GP logp+grad evals/s") and d["unit"] == "evals/s" and d["dtype"] == "f64"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"] and "l2" in d["config"]
    assert d["value"] == pytest.approx(d["n_gpus"] * d["config"]["B"] / (d["ms_per_step"] * 1e-3), rel=1e-9)
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["h2d_bytes_per_step"] > 0
    assert 0.9 * d["value"] < e["value"] < d["value"]                 # measured separately, through the public API
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-12) and 0 < r["frac"] <= 1
    assert r["traffic"] is None or r["traffic"] > 0
    if d["n_gpus"] == 1 and "cpu_baseline" in d:
        b = d["cpu_baseline"]
        assert {"value", "unit", "cores", "kind", "sample"} <= set(b) and b["kind"] in ("reference", "port") and b["cores"] >= 1


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and BASE <= set(d)
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
