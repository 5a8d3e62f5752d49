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
    ap.add_argument("--steps", type=int, default=120)   # >= 120 timed steps
    ap.add_argument("--warmup", type=int, default=20)   # ... after 20 warm-up steps
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dam_break",
                    choices=["dam_break", "uniform_box", "default"])
    ap.add_argument("--particles", type=int, default=0, help="total particles (N=1 default 1M)")
    ap.add_argument("--particles-per-gpu", type=int, default=8_000_000)
    ap.add_argument("--neighbours", type=float, default=50.0, help="uniform_box: target count")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-16m", action="store_true",
                    help="N = 1 default workload: skip the 16M roofline case (BASELINE configs[2])")
    ap.add_argument("--no-weak-baseline", action="store_true",
                    help="N > 1: skip rank 0's single-GPU run at the per-GPU load")
    ap.add_argument("--no-rebalance", action="store_true",
                    help="N > 1: skip the mid-run refresh of the slab cuts")
    ap.add_argument("--no-slab-parity", action="store_true",
                    help="N > 1: skip the bit-for-bit check against rank 0's whole-grid run")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--simple-kernels", action="store_true")
    ap.add_argument("--e2e-separate-calls", action="store_true",
                    help="e2e through upload + step + download instead of wc_step_host")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: slab-neighbour transport (peer memory over NVLink, or NCCL P2P)")
    return ap.parse_args()


def host_threads():
    """Host threads this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers,
    which says nothing about the box: the CPU legs size their thread teams from the affinity
    mask and pass the count to the oracle explicitly (its loops carry num_threads(nt))."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


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


# ---------------------------------------------------------------------------- workload config
def scene_params(args, n):
    """(n, size, grid_res, particle_radius) of make_scene(args, n) without building the particles."""
    if args.workload == "default":
        return 80000, 1.0, 21, scenes.DEFAULT_RADIUS
    if args.workload == "uniform_box":
        h = scenes.smoothing_length_for_neighbours(args.neighbours)
        size = float((n / (1.0 / 0.0175 ** 3)) ** (1.0 / 3.0))
        return n, size, max(int(np.floor(size / h)), 1), float(h) / 4.0
    size, grid_res = scenes.scaled_box(n)
    return n, size, grid_res, scenes.DEFAULT_RADIUS


def total_particles(args, world):
    if args.workload == "default":
        return 80000
    return args.particles or (1_000_000 if world == 1 else args.particles_per_gpu * world)


def arm_config(args, world, slab_layout=None):
    """The workload a line is quoted on.  BOTH arms print this same dict for the same command
    line (the reference arm times a bounded sample of it and says so in cpu_baseline.sample)."""
    n, size, grid_res, radius = scene_params(args, total_particles(args, world))
    cfg = {"workload": f"{args.workload}, {n} particles, full SPH step "
                       f"(sort + density + force + integrate)",
           "particles": int(n), "box_size": size, "grid_res": grid_res,
           "particle_radius": radius, "frame_dt": FRAME_DT,
           "parallelism": "single GPU" if world == 1 else f"z-slabs x{world}",
           "kernels": "simple" if args.simple_kernels else "tiled"}
    if world == 1:
        working_set = n * 32 * 2 + n * 16 + grid_res ** 3 * 8
        cfg["l2"] = ("L2 flushed between timed steps (256 MiB write)" if working_set < 2 * L2_BYTES
                     else "working set %.0f MiB > L2, no flush" % (working_set / 2 ** 20))
        return cfg
    from watercube_b200 import slab_bench

    hist, cuts = slab_layout or slab_bench.slab_layout(n, world)
    cfg.update({
        "l2": "per-rank working set > L2, no flush",
        "slab_cuts": [int(c) for c in cuts],
        "particles_per_rank": [int(hist[cuts[r]:cuts[r + 1]].sum()) for r in range(world)],
        "exchange": ("peer memory (CUDA IPC over NVLink): the producing kernels store into the "
                     "neighbour's buffers + device-side signals" if args.exchange == "peer"
                     else "NCCL P2P") + " with slab neighbours: layer counts, halo positions, "
                    "halo rho/P/velocity, migrants"})
    return cfg


# ---------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU-side implementation of the path: the oracle port (the GLSL
    pipeline itself needs an OpenGL 4.6 context and Cinder), on all host threads
    this process may use, each step one full step of a bounded sample of the workload.  Under
    torchrun rank 0 alone runs it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as ob

    world = max(1, args.gpus)
    n_total = total_particles(args, world)
    n = min(n_total, 1_000_000)
    sc = make_scene(args, n)
    p = ob.default_params(num_particles=sc.n, size=sc.size, grid_res=sc.grid_res,
                          particle_radius=sc.particle_radius)
    threads = host_threads()
    st = ob.Stepper(sc.particles, p, nthreads=threads)
    for _ in range(args.warmup):
        st.step(FRAME_DT)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st.step(FRAME_DT)
    dt = time.perf_counter() - t0
    value = sc.n * args.steps / dt
    sample = (f"{args.steps} full steps (after {args.warmup} warm-up) of a {sc.n}-particle "
              f"{args.workload} scene (box {sc.size:.4f}, gridRes {sc.grid_res}), "
              f"oracle/libwc_oracle.so, OpenMP {threads} threads")
    if sc.n != n_total:
        sample += (f"; bounded sample of the {n_total}-particle workload: same generator, number "
                   f"density, kernel radius and bin/h ratio, so the work per particle-update is "
                   f"the workload's")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": arm_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample, "sample_particles": int(sc.n)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cpu_baseline(args, sc, budget_s=20.0):
    from oracle import binding as ob

    n = min(sc.n, 1_000_000)
    sub = sc if n == sc.n else make_scene(args, n)
    p = ob.default_params(num_particles=sub.n, size=sub.size, grid_res=sub.grid_res,
                          particle_radius=sub.particle_radius)
    threads = host_threads()
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
def measure_single_gpu(args, sc, steps, warmup, local, with_e2e, sample_clocks=True):
    """Whole-grid run of scene `sc` on cuda:local.  Device-timed steps (CUDA events on the
    library's stream), per-stage times, the roofline figures and, optionally, the end-to-end
    rate through wc_step_host with pinned host buffers."""
    import torch

    from watercube_b200 import capi

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
    if working_set < 2 * L2_BYTES:   # same rule as arm_config()'s "l2" note
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def one_step():
        if flush is not None:
            flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fl.step(FRAME_DT)
        e1.record(stream)
        return e0, e1

    warmup = max(warmup, 3)
    for _ in range(warmup):
        one_step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local) if sample_clocks else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    launches0 = fl.launch_count()
    stage_ms = {k: 0.0 for k in capi.STAGES}
    torch.cuda.synchronize()
    t_wall0 = time.time()
    events = []
    for _ in range(steps):
        events.append(one_step())
        for k, v in fl.stage_times().items():  # syncs on the step's last event
            stage_ms[k] += v
    torch.cuda.synchronize()
    t_wall1 = time.time()
    launches = fl.launch_count() - launches0
    ms_per_step = sum(e0.elapsed_time(e1) for e0, e1 in events) / steps

    # ---- e2e: pinned host buffers in and out of the C-ABI every step
    e2e = None
    if with_e2e:
        h_in = torch.empty((n, 8), dtype=torch.float32, pin_memory=True)
        h_out = torch.empty((n, 8), dtype=torch.float32, pin_memory=True)
        h_in.copy_(torch.from_numpy(fl.download(1)))
        ev = []
        for it in range(3 + steps):
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
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    fl.close()
    del flush
    torch.cuda.empty_cache()

    # ---- roofline of the dominant kernel (stage times are CUDA events on the same stream)
    peak, peak_src = peak_hbm()
    per_stage = {k: v / steps for k, v in stage_ms.items()}
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
                     "note": "whole step; the gathers are FP32-issue / shared-memory bound, "
                             "see DESIGN.md"}
    out = {"particles": n, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
           "value": n / (ms_per_step * 1e-3), "stage_ms": per_stage, "roofline": roofline,
           "roofline_step": roofline_step, "gpu_launches": int(launches), "dominant": dom}
    if clocks is not None:
        out["clocks"] = clocks
    if e2e:
        out["e2e"] = e2e
    return out


def run_b200(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        from watercube_b200 import slab_bench

        return slab_bench.run(args, rank, world, local)

    sc = make_scene(args, total_particles(args, 1))
    m = measure_single_gpu(args, sc, args.steps, args.warmup, local, with_e2e=not args.no_e2e)
    line = {
        "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": m["warmup"], "ms_per_step": m["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": arm_config(args, 1),
        "stage_ms": m["stage_ms"], "roofline": m["roofline"], "roofline_step": m["roofline_step"],
        "clocks": m["clocks"], "gpu_launches": m["gpu_launches"],
    }
    ir = issue_roofline(KERNEL_NAMES[m["dominant"]], sc.n, m["stage_ms"][m["dominant"]],
                        m["clocks"].get("sm_mhz"))
    if ir:
        line["issue_roofline"] = ir
    if "e2e" in m:
        line["e2e"] = m["e2e"]
    # BASELINE.json configs[2], the single-GPU roofline case (north star: fraction of the HBM
    # roofline for the full step at 16M): measured in the same run so it is driver-visible.
    default_workload = args.workload == "dam_break" and not args.particles
    if default_workload and not args.no_16m:
        sc16 = scenes.dam_break(16_000_000, seed=0)
        m16 = measure_single_gpu(args, sc16, min(args.steps, 20), 3, local, with_e2e=False)
        line["roofline_16m"] = {
            "config": "Dam break, 16M particles, 1xB200 (BASELINE.json configs[2]); "
                      "working set > L2, no flush",
            "particles": sc16.n, "box_size": sc16.size, "grid_res": sc16.grid_res,
            "steps": m16["steps"], "warmup": m16["warmup"], "ms_per_step": m16["ms_per_step"],
            "value": m16["value"], "unit": UNIT, "stage_ms": m16["stage_ms"],
            "bound": "hbm", "achieved": m16["roofline_step"]["achieved"],
            "peak": m16["roofline_step"]["peak"], "frac": m16["roofline_step"]["frac"],
            "algorithmic_bytes_per_particle": ALGO_BYTES_STEP,
            "dominant_kernel": m16["roofline"], "clocks": m16["clocks"],
            "north_star_target_frac": 0.60}
        del sc16
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
