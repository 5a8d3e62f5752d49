"""bench.py's multi-GPU leg: z-slab decomposed dam break, one rank per GPU (torchrun).

Weak scaling: every rank owns about --particles-per-gpu particles (8M x 8 GPUs = BASELINE's
64M configuration).  The scene is the jittered dam-break lattice of scenes.dam_break; a rank
generates only the lattice planes that can fall into its slab (the lattice index is z-major),
so set-up cost does not grow with the world size.  Slab cuts give equal particle counts
(slab.slab_cuts over the global z-layer histogram).

Exchange: torch.distributed P2P over NCCL between slab neighbours only (layer counts, halo
positions, halo density/pressure/velocity, migrants) -- see watercube_b200/slab.py.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np

from . import scenes, slab

FRAME_DT = 1.0 / 60.0


def _layer_hist_and_ranges(n, d, spacing, jitter, half, bin_size, grid_res, seed, chunk=1 << 22):
    """Global z-layer histogram of the n-particle dam break from the z coordinates alone."""
    hist = np.zeros(grid_res, np.int64)
    for s0 in range(0, n, chunk):
        idx = np.arange(s0, min(n, s0 + chunk), dtype=np.int64)
        z = (idx // (d * d)).astype(np.float32) * spacing
        z += scenes.hash_u01(3 * idx + 2, seed) * jitter - half
        hist += np.bincount(slab.layer_of(z, bin_size, grid_res), minlength=grid_res)
    return hist


def _lattice(n_total):
    radius = scenes.DEFAULT_RADIUS
    size, grid_res = scenes.scaled_box(n_total, radius)
    d = scenes.lattice_side(n_total)
    spacing = np.float32(radius) * np.float32(scenes.DEFAULT_SPACING_FACTOR)
    jitter = spacing * np.float32(0.5)
    half = jitter / np.float32(2.0)
    bin_size = np.float32(size) / np.float32(grid_res)
    return radius, size, grid_res, d, spacing, jitter, half, bin_size


def slab_layout(n_total, world, seed=0):
    """-> (global z-layer histogram, equal-count cuts) of the n_total-particle dam break."""
    radius, size, grid_res, d, spacing, jitter, half, bin_size = _lattice(n_total)
    hist = _layer_hist_and_ranges(n_total, d, spacing, jitter, half, bin_size, grid_res, seed)
    return hist, slab.slab_cuts(hist, world)


def make_rank_scene(n_total, rank, world, seed=0):
    """-> (scene params, cuts, this rank's particles [m, 8])."""
    radius, size, grid_res, d, spacing, jitter, half, bin_size = _lattice(n_total)
    hist, cuts = slab_layout(n_total, world, seed)
    # lattice planes that can reach this rank's layers (jitter < one plane spacing)
    z_lo, z_hi = cuts[rank] * float(bin_size), cuts[rank + 1] * float(bin_size)
    p_lo = max(int(np.floor(z_lo / float(spacing))) - 1, 0)
    p_hi = min(int(np.ceil(z_hi / float(spacing))) + 2, d)
    i0, i1 = min(p_lo * d * d, n_total), min(p_hi * d * d, n_total)
    out = []
    chunk = 1 << 22
    for s0 in range(i0, i1, chunk):
        idx = np.arange(s0, min(i1, s0 + chunk), dtype=np.int64)
        xyz = np.stack([idx % d, (idx // d) % d, idx // (d * d)], axis=1).astype(np.float32)
        pos = xyz * spacing
        for k in range(3):
            pos[:, k] += scenes.hash_u01(3 * idx + k, seed) * jitter - half
        lay = slab.layer_of(pos[:, 2], bin_size, grid_res)
        keep = (lay >= cuts[rank]) & (lay < cuts[rank + 1])
        part = np.zeros((int(keep.sum()), scenes.PARTICLE_FLOATS), np.float32)
        part[:, 0:3] = pos[keep]
        out.append(part)
    mine = np.concatenate(out) if out else np.zeros((0, scenes.PARTICLE_FLOATS), np.float32)
    params = dict(grid_res=int(grid_res), size=float(size), particle_radius=float(radius))
    return params, cuts, hist, mine


def check_slab_parity(b, one_step, n_total, rank, world, local, stream, steps=2):
    """Decomposed run == whole-grid run, bit for bit, on real ranks: every rank advances its
    slab `steps` steps; rank 0 also runs the SAME scene undecomposed on its own GPU; the ranks'
    buffer 1 (download(1)), concatenated in rank order, must equal the whole-grid buffer 1
    (slab.py: the concatenation IS the whole-grid state in the whole-grid order).  Collective:
    every rank calls it.  -> dict for the bench line (raises on a mismatch)."""
    import torch
    import torch.distributed as dist

    from . import capi

    for _ in range(steps):
        one_step()
    torch.cuda.synchronize()
    mine = torch.from_numpy(b.download(1)).view(torch.int32).cuda()
    counts = [None] * world
    dist.all_gather_object(counts, int(mine.shape[0]))
    ok = True
    if rank == 0:
        sc = scenes.dam_break(n_total, seed=0)
        with capi.Fluid(num_particles=sc.n, grid_res=sc.grid_res, size=sc.size,
                        particle_radius=sc.particle_radius, device=local,
                        stream=stream.cuda_stream) as fl:
            fl.upload(sc.particles)
            del sc
            for _ in range(steps):
                fl.step(FRAME_DT)
            whole = torch.from_numpy(fl.download(1)).view(torch.int32)
        off = 0
        for r in range(world):
            ref = whole[off:off + counts[r]].cuda()
            if r == 0:
                got = mine
            else:
                got = torch.empty((counts[r], 8), dtype=torch.int32, device="cuda")
                dist.recv(got, src=r)
            ok = ok and ref.shape == got.shape and bool(torch.equal(ref, got))
            off += counts[r]
        ok = ok and off == whole.shape[0] == n_total
        del whole
    else:
        dist.send(mine, dst=0)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    if not int(flag.item()):
        raise RuntimeError("slab parity FAILED: the decomposed run differs from the whole-grid run")
    return {"status": "bit-exact", "steps": steps, "particles": int(n_total),
            "against": "whole-grid run of the same scene on rank 0's GPU; ranks' buffer 1 "
                       "concatenated in rank order, compared as raw 32-bit words"}


def run(args, rank, world, local):
    import torch
    import torch.distributed as dist

    from . import capi
    import bench  # the launching script: shared helpers (clock sampler, peaks, config)

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    n_total = args.particles or args.particles_per_gpu * world
    params, cuts, hist, mine = make_rank_scene(n_total, rank, world)
    n_mine = mine.shape[0]
    layer_max = int(hist.max())
    cap = int(n_mine * 1.25) + 4 * layer_max + 1024
    ghost_cap = int(layer_max * 1.5) + 1024
    mig_cap = max(layer_max // 2, 65536)
    flags = capi.FLAG_STAGE_TIMING

    def make_backend(z0, z1):
        return slab.CudaSlabBackend(params, z0, z1, capacity=cap, ghost_capacity=ghost_cap,
                                    migrant_capacity=mig_cap, device=local, flags=flags,
                                    stream=stream.cuda_stream)

    b = make_backend(cuts[rank], cuts[rank + 1])
    b.upload(mine)
    del mine
    drv = slab.SlabDriver(b, rank, world)
    peer = args.exchange == "peer"
    if peer:
        slab.attach_peers_ipc(b, rank, world)

    def one_step():
        if peer:
            slab.run_step_peer(b, FRAME_DT, wait=False)   # queued only: no host wait in a step
        else:
            slab.run_step(drv, FRAME_DT)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    parity = None
    if not args.no_slab_parity:
        parity = check_slab_parity(b, one_step, n_total, rank, world, local, stream)

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        one_step()
    barrier()

    sampler = bench.ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    launches0 = b.fluid.launch_count()
    stage_ms = {k: 0.0 for k in capi.STAGES}
    barrier()
    t_wall0 = time.time()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):   # no per-step sync: the host runs ahead like a real frame loop
        one_step()
    e1.record(stream)
    barrier()
    t_wall1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)       # device time, max over ranks
    ms_per_step = float(ms.item()) / args.steps
    launches = b.fluid.launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    # stage split from a few extra steps (reading the stage events syncs, so not in the timed loop)
    # (every rank starts such a step together: the waits for a neighbour's message sit inside the
    # kernels now, so a stage time includes them -- this keeps them to the real dependencies)
    n_stage = max(1, min(3, args.steps))
    for _ in range(n_stage):
        barrier()
        one_step()
        for k, v in b.fluid.stage_times().items():
            stage_ms[k] += v
    barrier()
    n_now = torch.tensor([b.num_particles], device="cuda", dtype=torch.int64)
    dist.all_reduce(n_now)
    assert int(n_now.item()) == n_total, (int(n_now.item()), n_total)  # nothing lost in migration
    value = n_total / (ms_per_step * 1e-3)

    # ---- e2e: every rank's slab goes host -> device and back every step.  Two routes through
    # the C-ABI are timed and the faster one is the line's e2e (the other is reported beside it):
    # wc_slab_step_peer_host (the update kernel stores the result into the pinned output itself:
    # no extra pass, but the stores cross PCIe at kernel speed) and upload + step + download
    # (pack kernel + copy engine).
    e2e = None
    if not args.no_e2e:
        cur = b.download(1)
        h_in = torch.empty((cap, 8), dtype=torch.float32, pin_memory=True)
        h_out = torch.empty((cap, 8), dtype=torch.float32, pin_memory=True)
        n_cur = cur.shape[0]
        h_in[:n_cur].copy_(torch.from_numpy(cur))
        del cur
        steps_e = max(3, min(args.steps, 6))

        def timed_e2e(fused):
            nonlocal h_in, h_out, n_cur
            h2d = d2h = 0
            barrier()
            ev0 = torch.cuda.Event(enable_timing=True)
            ev1 = torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            for _ in range(steps_e):
                h2d += n_cur * 32
                if fused:
                    info = b.fluid.slab_step_peer_host((h_in.data_ptr(), n_cur), h_out.data_ptr(),
                                                       cap, FRAME_DT)
                    n_cur = info["n_owned"]
                else:
                    b.fluid.upload((h_in.data_ptr(), n_cur))
                    one_step()
                    n_cur = b.num_particles
                    b.fluid.download(1, out=(h_out.data_ptr(), n_cur))
                d2h += n_cur * 32
                h_in, h_out = h_out, h_in
            ev1.record(stream)
            barrier()
            ms_e = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
            dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
            tot = torch.tensor([h2d, d2h], device="cuda", dtype=torch.int64)
            dist.all_reduce(tot)
            ms_step = float(ms_e.item()) / steps_e
            return {"value": n_total / (ms_step * 1e-3), "unit": bench.UNIT, "ms_per_step": ms_step,
                    "h2d_bytes_per_step": int(tot[0].item()) // steps_e,
                    "d2h_bytes_per_step": int(tot[1].item()) // steps_e,
                    "api": "wc_slab_step_peer_host per rank (pinned host AoS in and out; D2H fused "
                           "into the update kernel)" if fused else
                           "wc_upload_particles + wc_slab_step_peer + wc_download_particles per rank "
                           "(pinned host AoS; pack kernel + copy engine)"}

        routes = [timed_e2e(False)]
        if peer:
            routes.append(timed_e2e(True))
        routes.sort(key=lambda r: r["ms_per_step"])
        e2e = routes[0]
        if len(routes) > 1:
            e2e["other_route"] = {k: routes[1][k] for k in ("ms_per_step", "api")}
        del h_in, h_out

    # ---- the cuts are refreshed every k steps because the fluid moves along z.
    # One refresh in the measured run (not in the timed region: it happens every k >> 1 steps):
    # histogram all-reduce, all-to-all of the AoS records, fresh handles, neighbours re-attached,
    # then the run goes on; nothing may be lost and the counts must not get less even.
    rebalance = None
    if not args.no_rebalance:
        before = [None] * world
        dist.all_gather_object(before, int(b.num_particles))
        bin_size = float(np.float32(params["size"]) / np.float32(params["grid_res"]))
        barrier()
        t0 = time.time()
        new_cuts, b = slab.rebalance(b, rank, world, make_backend, bin_size, params["grid_res"])
        drv = slab.SlabDriver(b, rank, world)
        if peer:
            slab.attach_peers_ipc(b, rank, world)
        barrier()
        t_reb = time.time() - t0
        for _ in range(3):
            one_step()
        barrier()
        after = [None] * world
        dist.all_gather_object(after, int(b.num_particles))
        assert sum(after) == n_total, (after, n_total)
        rebalance = {"seconds": t_reb, "cuts_before": [int(c) for c in cuts],
                     "cuts_after": [int(c) for c in new_cuts], "particles_before": before,
                     "particles_after_3_more_steps": after,
                     "how": "slab.rebalance: layer-histogram all-reduce, all-to-all of the AoS "
                            "records, fresh handles, peers re-attached; outside the timed region"}

    per_stage = {k: v / n_stage for k, v in stage_ms.items()}
    stage_t = torch.tensor([per_stage[k] for k in capi.STAGES], device="cuda")
    dist.all_reduce(stage_t, op=dist.ReduceOp.MAX)
    per_stage = dict(zip(capi.STAGES, [float(x) for x in stage_t.tolist()]))
    counts = [None] * world
    dist.all_gather_object(counts, int(b.num_particles))
    b.close()
    torch.cuda.empty_cache()

    # ---- the same per-GPU load on ONE GPU, whole grid, in this very run: the denominator of
    # the weak-scaling efficiency (the N = 1 bench line is the 1M scene, another load)
    weak = None
    if rank == 0 and not args.no_weak_baseline:
        ppg = n_total // world
        wb = bench.measure_single_gpu(args, scenes.dam_break(ppg, seed=0), min(args.steps, 20), 3,
                                      local, with_e2e=False, sample_clocks=False)
        weak = {"particles": wb["particles"], "ms_per_step": wb["ms_per_step"],
                "value": wb["value"], "unit": bench.UNIT, "steps": wb["steps"],
                "stage_ms": wb["stage_ms"],
                "what": f"whole-grid dam break of {ppg} particles on rank 0's GPU alone, after "
                        f"the {world}-GPU run"}
    if rank == 0:
        peak, peak_src = bench.peak_hbm()
        dom = max(per_stage, key=per_stage.get)
        n_rank_max = max(counts)
        achieved = bench.ALGO_BYTES[dom] * n_rank_max / (per_stage[dom] * 1e-3) / 1e9
        step_gbs = bench.ALGO_BYTES_STEP * n_total / world / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": bench.METRIC, "value": value, "unit": bench.UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": bench.arm_config(args, world, (hist, cuts)),
            "particles_per_rank_end": counts,
            "stage_ms": per_stage,
            "roofline": {"bound": "hbm", "kernel": bench.KERNEL_NAMES[dom], "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": bench.recorded_traffic(bench.KERNEL_NAMES[dom], n_rank_max),
                         "algorithmic_bytes_per_particle": bench.ALGO_BYTES[dom],
                         "peak_source": peak_src, "ms_per_launch": per_stage[dom],
                         "note": "largest rank, per GPU"},
            "roofline_step": {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s",
                              "frac": step_gbs / peak, "note": "per GPU, whole step incl. exchange"},
            "clocks": clocks, "gpu_launches": int(launches) * world,
            "gpu_launches_per_step_per_rank": int(launches) / max(args.steps, 1),
        }
        if parity:
            line["slab_parity"] = parity["status"]
            line["slab_parity_detail"] = parity
        if weak:
            line[f"weak_baseline_{weak['particles'] // 1_000_000}m"] = weak
            line["efficiency_vs_weak_baseline"] = value / (world * weak["value"])
        if rebalance:
            line["rebalance"] = rebalance
        if e2e:
            line["e2e"] = e2e
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
