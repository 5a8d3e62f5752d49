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
    assert cb["sample_particles"] == 20000 and "20000-particle" in cb["sample"]


def test_reference_arm_under_torchrun_uses_all_host_threads_and_the_arms_config():
    """torchrun exports OMP_NUM_THREADS=1 to its workers: the reference arm still times the
    oracle on every thread it may use, says how many, prints the SAME config dict the B200 arm
    prints for that command line (the bounded sample is described in cpu_baseline), and only
    rank 0 prints."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")
    args = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
            "--particles", "60000", "--steps", "1", "--warmup", "1"]
    res = subprocess.run(args, capture_output=True, text=True, cwd=ROOT, timeout=900, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    d = json.loads([ln for ln in res.stdout.strip().splitlines() if ln.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    cfg = d["config"]
    assert cfg["particles"] == 60000 and cfg["parallelism"] == "z-slabs x2"
    assert len(cfg["slab_cuts"]) == 3 and sum(cfg["particles_per_rank"]) == 60000
    sys.path.insert(0, ROOT)
    import bench

    ns = bench.parse_args.__globals__["argparse"].Namespace  # same helper the B200 arm calls
    a = ns(workload="dam_break", particles=60000, particles_per_gpu=8_000_000, neighbours=50.0,
           simple_kernels=False, exchange="peer")
    assert bench.arm_config(a, 2) == cfg
    env["RANK"] = "1"
    res = subprocess.run(args, capture_output=True, text=True, cwd=ROOT, timeout=900, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


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
    assert "roofline_16m" not in d            # only the default workload carries the 16M case


@pytest.mark.gpu
def test_b200_arm_default_line_carries_the_16m_roofline_case():
    d = run_bench("--steps", "5", "--warmup", "3", "--no-cpu-baseline")
    assert d["config"]["particles"] == 1_000_000
    m = d["roofline_16m"]
    assert m["particles"] == 16_000_000 and m["bound"] == "hbm" and 0 < m["frac"] < 1
    assert m["frac"] == pytest.approx(168 * 16e6 / (m["ms_per_step"] * 1e-3) / 1e9 / m["peak"], rel=1e-6)
    assert m["dominant_kernel"]["kernel"] in ("k_density_tile", "k_update_tile")
