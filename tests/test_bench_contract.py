"""bench.py prints ONE JSON line with the keys the measurement contract names (both arms)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
             "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e"}


def run_bench(*args):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True,
                         text=True, cwd=ROOT, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, res.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    """--impl reference: the oracle port on the host cores, same metric / unit / config keys."""
    d = run_bench("--impl", "reference", "--particles", "20000", "--steps", "2", "--warmup", "1")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "particle_updates_per_sec" and d["unit"] == "particle-updates/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


@pytest.mark.gpu
def test_b200_arm_line():
    d = run_bench("--particles", "200000", "--steps", "5", "--warmup", "3")
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 5 and d["value"] > 0 and d["dtype"] == "f32"
    assert d["value"] == pytest.approx(200000 / (d["ms_per_step"] * 1e-3), rel=1e-6)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9) and "traffic" in r
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 200000 * 32 == e["d2h_bytes_per_step"]
    assert 0 < e["value"] < d["value"]                      # host copies are inside the timed region
    assert d["gpu_launches"] >= 5 * 5                       # >= one kernel per stage and step
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and isinstance(c["reasons"], list)
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and set(cb["stage_ms"]) == {"bin_sort", "density", "update"}
    assert cb["single_thread"]["value"] > 0
    assert set(d["stage_ms"]) == {"hash_count", "scan", "reorder", "density", "update"}
