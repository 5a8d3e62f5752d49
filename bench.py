#!/usr/bin/env python
"""bench.py -- particle-updates/s of the full SPH step (sort + density + force + integrate).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload dam_break|uniform_box|default] [--particles P]

N = 1 (default): BASELINE.json configs[1], "Dam break, 1M particles, 1xB200, full SPH step".
N > 1 (under torchrun): z-slab decomposed dam break, weak scaling at --particles-per-gpu
(default 8M, so N = 8 is BASELINE's 64M configuration).

One JSON line on stdout (rank 0).  `value` is device-timed with inputs resident in HBM;
`e2e` is the same step through the C-ABI with pinned HOST buffers (H2D + step + D2H per
step).  `roofline` is for the dominant kernel against MEASURED_PEAKS.json; `cpu_baseline`
is the CPU oracle (a port of the reference's shaders: the reference itself needs
Cinder + OpenGL 4.6 + Win32 and cannot be built) timed on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from watercube_b200 import scenes  # noqa: E402

METRIC = "particle_updates_per_sec"
UNIT = "particle-updates/s"
FRAME_DT = 1.0 / 60.0
# compulsory DRAM bytes per particle-update, by stage (sum = 168 B).
ALGO_BYTES = {"hash_count": 16, "scan": 0, "reorder": 64, "density": 24, "update": 64}
ALGO_BYTES_STEP = 168
KERNEL_NAMES = {"hash_count": "k_hash_count", "scan": "k_scan", "reorder": "k_reorder",
                "density": "k_density_tile", "update": "k_update_tile"}
L2_BYTES = 126 * 1024 * 1024
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dam_break",
                    choices=["dam_break", "uniform_box", "default"])
    ap.add_argument("--particles", type=int, default=0, help="total particles (N=1 default 1M)")
    ap.add_argument("--particles-per-gpu", type=int, default=8_000_000)
    ap.add_argument("--neighbours", type=float, default=50.0, help="uniform_box: target count")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--simple-kernels", action="store_true")
    ap.add_argument("--e2e-separate-calls", action="store_true",
                    help="e2e through upload + step + download instead of wc_step_host")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: slab-neighbour transport (peer memory over NVLink, or NCCL P2P)")
    return ap.parse_args()


def make_scene(args, n):
    if args.workload == "default":
        return scenes.dam_break(80000, seed=0)
    if args.workload == "uniform_box":
        h = scenes.smoothing_length_for_neighbours(args.neighbours)
        size = float((n / (1.0 / 0.0175 ** 3)) ** (1.0 / 3.0))
        return scenes.uniform_box(n, size=size, h=h, seed=0)
    return scenes.dam_break(n, seed=0)


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def recorded(kernel_key, n, field="dram_bytes"):
    """Per-launch figure from the committed `ncu --set full` capture (profiles/traffic.json,
    written by tools/make_traffic.py), or None when this kernel / size was not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        e = t.get(f"{kernel_key}@{n}")
        return float(e[field]) if e else None
    except Exception:
        return None


def recorded_traffic(kernel_key, n):
    return recorded(kernel_key, n, "dram_bytes")


def issue_roofline(kernel_key, n, ms_per_launch, sm_mhz):
    """How close the kernel runs to the SM instruction-issue ceiling: warp instructions per
    launch (from the committed ncu capture) / duration, against 148 SMs x 4 schedulers x the
    SM clock sampled during the run.  Explains why the HBM fraction is small."""
    inst = recorded(kernel_key, n, "warp_inst")
    if inst is None or not sm_mhz:
        return None
    peak = 148 * 4 * sm_mhz * 1e6
    achieved = inst / (ms_per_launch * 1e-3)
    return {"bound": "issue", "kernel": kernel_key, "achieved": achieved / 1e9, "peak": peak / 1e9,
            "unit": "G warp-inst/s", "frac": achieved / peak,
            "warp_inst_per_particle": inst / n}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (recipe line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU-side implementation of the path: the oracle port (the GLSL
    pipeline itself needs an OpenGL 4.6 context and Cinder), all host threads,
    on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as ob

    n_total = args.particles or (1_000_000 if args.gpus == 1 else args.particles_per_gpu * args.gpus)
    n = min(n_total, 1_000_000)
    sc = make_scene(args, n)
    p = ob.default_params(num_particles=sc.n, size=sc.size, grid_res=sc.grid_res,
                          particle_radius=sc.particle_radius)
    threads = ob.max_threads()
    st = ob.Stepper(sc.particles, p, nthreads=threads)
    for _ in range(args.warmup):
        st.step(FRAME_DT)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st.step(FRAME_DT)
    dt = time.perf_counter() - t0
    value = sc.n * args.steps / dt
    sample = (f"{args.steps} full steps (after {args.warmup} warm-up) of the {sc.n}-particle "
              f"{args.workload} scene" + ("" if n == n_total else f" (bounded sample of {n_total})"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, sc, n_total, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(args, sc, n_total, n_gpus, extra=None):
    cfg = {"workload": f"{args.workload}, {n_total} particles, full SPH step "
                       f"(sort + density + force + integrate)",
           "particles": int(n_total), "box_size": sc.size, "grid_res": sc.grid_res,
           "particle_radius": sc.particle_radius, "frame_dt": FRAME_DT,
           "parallelism": "single GPU" if n_gpus == 1 else f"z-slabs x{n_gpus}"}
    if extra:
        cfg.update(extra)
    return cfg


def cpu_baseline(args, sc, budget_s=20.0):
    from oracle import binding as ob

    n = min(sc.n, 1_000_000)
    sub = sc if n == sc.n else make_scene(args, n)
    p = ob.default_params(num_particles=sub.n, size=sub.size, grid_res=sub.grid_res,
                          particle_radius=sub.particle_radius)
    threads = ob.max_threads()
    st = ob.Stepper(sub.particles, p, nthreads=threads)
    st.step(FRAME_DT)  # warm-up (page faults, thread pool)
    steps, t0 = 0, time.perf_counter()
    while steps < 3 or (time.perf_counter() - t0 < budget_s and steps < 20):
        st.step(FRAME_DT)
        steps += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    out = {"value": sub.n * steps / dt, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"{steps} full steps of the {sub.n}-particle {args.workload} scene, "
                     f"oracle/libwc_oracle.so, OpenMP {threads} threads, {dt:.1f} s"}
    # per-stage split (all threads) and the single-threaded rate, one step each.
    d = ob.derive(p)
    frame_dt = np.float32(FRAME_DT) * np.float32(p.time_scale)

    def staged(nthreads):
        t = [time.perf_counter()]
        srt = ob.sort(sub.particles, d.bin_size, p.grid_res)      # serial in the oracle
        t.append(time.perf_counter())
        P, _ = ob.density(srt["sorted"], srt["counts"], srt["offsets"], p, nthreads=nthreads)
        t.append(time.perf_counter())
        ob.update(P, srt["counts"], srt["offsets"], p, frame_dt, nthreads=nthreads)
        t.append(time.perf_counter())
        return [1e3 * (b - a) for a, b in zip(t[:-1], t[1:])]

    ms = staged(threads)
    out["stage_ms"] = {"bin_sort": ms[0], "density": ms[1], "update": ms[2]}
    ms1 = staged(1)
    out["single_thread"] = {"value": sub.n / (1e-3 * sum(ms1)), "unit": UNIT, "sample": "1 step",
                            "stage_ms": {"bin_sort": ms1[0], "density": ms1[1], "update": ms1[2]}}
    return out


# ---------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch

    from watercube_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        from watercube_b200 import slab_bench

        return slab_bench.run(args, rank, world, local)

    n_total = args.particles or 1_000_000
    sc = make_scene(args, n_total)
    n = sc.n
    # a non-default torch stream, so torch.cuda.Event sees the library's launches
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    flags = capi.FLAG_STAGE_TIMING | (capi.FLAG_SIMPLE_KERNELS if args.simple_kernels else 0)
    fl = capi.Fluid(num_particles=n, grid_res=sc.grid_res, size=sc.size,
                    particle_radius=sc.particle_radius, device=local, flags=flags,
                    stream=stream.cuda_stream)
    fl.upload(sc.particles)
    working_set = n * 32 * 2 + n * 16 + int(fl.derived.num_bins) * 8
    flush = None
    l2_note = "working set %.0f MiB > L2, no flush" % (working_set / 2 ** 20)
    if working_set < 2 * L2_BYTES:
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
        l2_note = "L2 flushed between timed steps (256 MiB write)"

    def one_step(timed):
        if flush is not None:
            flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fl.step(FRAME_DT)
        e1.record(stream)
        return e0, e1

    for _ in range(max(args.warmup, 3)):
        one_step(False)
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    launches0 = fl.launch_count()
    stage_ms = {k: 0.0 for k in capi.STAGES}
    torch.cuda.synchronize()
    t_wall0 = time.time()
    events = []
    for _ in range(args.steps):
        events.append(one_step(True))
        for k, v in fl.stage_times().items():  # syncs on the step's last event
            stage_ms[k] += v
    torch.cuda.synchronize()
    t_wall1 = time.time()
    launches = fl.launch_count() - launches0
    total_ms = sum(e0.elapsed_time(e1) for e0, e1 in events)
    ms_per_step = total_ms / args.steps
    value = n / (ms_per_step * 1e-3)

    # ---- e2e: pinned host buffers in and out of the C-ABI every step
    e2e = None
    if not args.no_e2e:
        h_in = torch.empty((n, 8), dtype=torch.float32, pin_memory=True)
        h_out = torch.empty((n, 8), dtype=torch.float32, pin_memory=True)
        h_in.copy_(torch.from_numpy(fl.download(1)))
        ev = []
        for it in range(3 + args.steps):
            if flush is not None:
                flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if args.e2e_separate_calls:
                fl.upload((h_in.data_ptr(), n))
                fl.step(FRAME_DT)
                fl.download(1, out=(h_out.data_ptr(), n))   # syncs: the result is on the host
            else:
                # wc_step_host: H2D of the pinned input, the step, and the result stored into
                # the pinned output by the update kernel itself; returns with it complete
                fl.step_host((h_in.data_ptr(), n), h_out.data_ptr(), FRAME_DT)
            e1.record(stream)
            e1.synchronize()
            if it >= 3:
                ev.append(e0.elapsed_time(e1))
            h_in, h_out = h_out, h_in
        e2e_ms = float(np.mean(ev))
        e2e = {"value": n / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": n * 32,
               "api": "wc_upload_particles + wc_step + wc_download_particles"
                      if args.e2e_separate_calls else
                      "wc_step_host (pinned host AoS in and out; D2H fused into the update kernel)"}
        launches_e2e = 3  # aos->soa, soa->aos + the step's kernels (reported for context)
    clocks = sampler.stop(t_wall0, t_wall1)

    # ---- roofline of the dominant kernel (stage times are CUDA events on the same stream)
    peak, peak_src = peak_hbm()
    per_stage = {k: v / args.steps for k, v in stage_ms.items()}
    dom = max(per_stage, key=per_stage.get)
    achieved = ALGO_BYTES[dom] * n / (per_stage[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": KERNEL_NAMES[dom], "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak,
                "traffic": recorded_traffic(KERNEL_NAMES[dom], n),
                "algorithmic_bytes_per_particle": ALGO_BYTES[dom], "peak_source": peak_src,
                "ms_per_launch": per_stage[dom]}
    step_gbs = ALGO_BYTES_STEP * n / (ms_per_step * 1e-3) / 1e9
    roofline_step = {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s",
                     "frac": step_gbs / peak, "algorithmic_bytes_per_particle": ALGO_BYTES_STEP,
                     "note": "whole step; the gathers are FP32-issue bound, see DESIGN.md"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, sc, n, 1, {"l2": l2_note,
                                                   "kernels": "simple" if args.simple_kernels
                                                   else "tiled"}),
        "stage_ms": per_stage, "roofline": roofline, "roofline_step": roofline_step,
        "clocks": clocks, "gpu_launches": int(launches),
    }
    ir = issue_roofline(KERNEL_NAMES[dom], n, per_stage[dom], clocks.get("sm_mhz"))
    if ir:
        line["issue_roofline"] = ir
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, sc)
    print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
