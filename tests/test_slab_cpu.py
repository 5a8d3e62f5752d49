"""N > 1 path on CPU: the slab decomposition / exchange logic of watercube_b200.slab with the
oracle standing in for the kernels (tests/slab_cpu_backend.py) -- in-process virtual ranks and
a real world_size-2 gloo run.  Bar: concatenating the ranks' buffers in z order reproduces the
undecomposed oracle run BIT-EXACTLY at every step (ghosts make neighbour sets identical and
the migrant ordering rule keeps the stable in-cell order)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.slab_cpu_backend import OracleSlabBackend
from watercube_b200 import scenes, slab

f32 = np.float32
FRAME_DT = 1.0 / 60.0


def make_scene():
    # a block that fills most of a small box, with velocities so particles cross slab faces
    sc = scenes.dam_break(6000, seed=3, size=0.5, grid_res=10)
    rng = np.random.default_rng(0)
    sc.particles[:, 4:7] = rng.uniform(-30, 30, (sc.n, 3)).astype(f32)
    return sc


def test_layer_of_matches_oracle_cells(oracle):
    sc = make_scene()
    sc.particles[:5, 2] = [-1.0, 0.0, 0.4999, 0.5, 7.0]
    p = oracle.default_params(num_particles=sc.n, size=sc.size, grid_res=sc.grid_res)
    d = oracle.derive(p)
    ids = oracle.cell_ids(sc.particles, d.bin_size, p.grid_res)
    np.testing.assert_array_equal(slab.layer_of(sc.particles[:, 2], d.bin_size, p.grid_res),
                                  ids // (p.grid_res ** 2))


def test_slab_cuts_balance_and_validity():
    hist = np.array([0, 0, 50, 100, 100, 100, 50, 0, 0, 0])
    for world in (1, 2, 3, 4, 8, 10):
        cuts = slab.slab_cuts(hist, world)
        assert cuts[0] == 0 and cuts[-1] == len(hist) and len(cuts) == world + 1
        assert all(b > a for a, b in zip(cuts, cuts[1:]))            # >= 1 layer per rank
    assert slab.slab_cuts(np.array([0, 0, 100, 100, 100, 100, 0, 0, 0, 0]), 2) == [0, 4, 10]
    assert slab.slab_cuts(np.array([10, 10, 10, 10, 10, 10, 10, 10, 10, 10]), 5) == [0, 2, 4, 6, 8, 10]
    with pytest.raises(ValueError):
        slab.slab_cuts(hist, 11)


def test_decompose_preserves_order_and_partitions(oracle):
    sc = make_scene()
    p = oracle.default_params(num_particles=sc.n, size=sc.size, grid_res=sc.grid_res)
    d = oracle.derive(p)
    cuts = [0, 3, 4, 10]
    parts = slab.decompose(sc.particles, cuts, d.bin_size, p.grid_res)
    assert sum(len(x) for x in parts) == sc.n
    lay = slab.layer_of(sc.particles[:, 2], d.bin_size, p.grid_res)
    for r, part in enumerate(parts):
        sel = (lay >= cuts[r]) & (lay < cuts[r + 1])
        np.testing.assert_array_equal(part, sc.particles[sel])


def reference_run(oracle, sc, p, steps):
    st = oracle.Stepper(sc.particles, p, nthreads=2)
    out = []
    for _ in range(steps):
        st.step(FRAME_DT)
        out.append((oracle.as_f32(st.buf1).copy(), oracle.as_f32(st.buf2).copy()))
    return out


@pytest.mark.parametrize("cuts", [[0, 10], [0, 4, 10], [0, 2, 3, 10], [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10]])
def test_virtual_ranks_equal_undecomposed_run(oracle, cuts):
    sc = make_scene()
    p = oracle.default_params(num_particles=sc.n, size=sc.size, grid_res=sc.grid_res)
    d = oracle.derive(p)
    steps = 6
    ref = reference_run(oracle, sc, p, steps)
    world = len(cuts) - 1
    parts = slab.decompose(sc.particles, cuts, d.bin_size, p.grid_res)
    drivers = []
    for r in range(world):
        b = OracleSlabBackend(oracle, p, cuts[r], cuts[r + 1], migrant_capacity=2000)
        b.upload(parts[r])
        drivers.append(slab.SlabDriver(b, r, world))
    migrated = 0
    for s in range(steps):
        slab.run_step_local(drivers, FRAME_DT)
        migrated += sum(dr.info["migrants_in_below"] + dr.info["migrants_in_above"] for dr in drivers)
        buf1 = np.concatenate([dr.backend.download(1) for dr in drivers])
        buf2 = np.concatenate([dr.backend.download(2) for dr in drivers])
        np.testing.assert_array_equal(buf2, ref[s][1], err_msg=f"sorted buffer, step {s}")
        np.testing.assert_array_equal(buf1, ref[s][0], err_msg=f"state buffer, step {s}")
        assert sum(dr.info["n_owned"] for dr in drivers) == sc.n
    if world > 1:
        assert migrated > 0                                       # the test exercises migration


def test_rebalance_virtual_ranks_continues_bit_identically(oracle):
    """Re-cutting the slabs mid-run (cuts refreshed every k steps) changes nothing in
    the results: 3 steps, re-cut, 3 more steps == 6 undecomposed steps, and the new cuts follow
    the moved fluid."""
    sc = make_scene()
    p = oracle.default_params(num_particles=sc.n, size=sc.size, grid_res=sc.grid_res)
    d = oracle.derive(p)
    steps, world = 6, 3
    ref = reference_run(oracle, sc, p, steps)
    cuts = [0, 1, 2, 10]                                          # deliberately lopsided
    parts = slab.decompose(sc.particles, cuts, d.bin_size, p.grid_res)
    make = lambda z0, z1: OracleSlabBackend(oracle, p, z0, z1, migrant_capacity=2000)
    backends = []
    for r in range(world):
        backends.append(make(cuts[r], cuts[r + 1]))
        backends[-1].upload(parts[r])
    for s in range(steps):
        if s == 3:
            before = [b.num_particles for b in backends]
            cuts, backends = slab.rebalance_local(backends, make, d.bin_size, p.grid_res)
            after = [b.num_particles for b in backends]
            assert sum(after) == sc.n and max(after) - min(after) < max(before) - min(before)
        drivers = [slab.SlabDriver(b, r, world) for r, b in enumerate(backends)]
        slab.run_step_local(drivers, FRAME_DT)
        buf1 = np.concatenate([b.download(1) for b in backends])
        np.testing.assert_array_equal(buf1, ref[s][0], err_msg=f"state buffer, step {s}")


# --------------------------------------------------------------------------- real gloo run
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, cuts, steps, out_dir):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import binding as ob

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = make_scene()
        p = ob.default_params(num_particles=sc.n, size=sc.size, grid_res=sc.grid_res)
        d = ob.derive(p)
        parts = slab.decompose(sc.particles, cuts, d.bin_size, p.grid_res)
        b = OracleSlabBackend(ob, p, cuts[rank], cuts[rank + 1], migrant_capacity=2000)
        b.upload(parts[rank])
        drv = slab.SlabDriver(b, rank, world)
        for s in range(steps):
            if s == 2:   # re-cut the slabs in the middle of the run (all-reduce + all-to-all)
                make = lambda z0, z1: OracleSlabBackend(ob, p, z0, z1, migrant_capacity=2000)
                new_cuts, b = slab.rebalance(b, rank, world, make, d.bin_size, p.grid_res)
                assert new_cuts[0] == 0 and new_cuts[-1] == p.grid_res
                drv = slab.SlabDriver(b, rank, world)
            slab.run_step(drv, FRAME_DT)
            np.save(os.path.join(out_dir, f"buf1_s{s}_r{rank}.npy"), b.download(1))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_process_gloo_run_equals_undecomposed_run(oracle, tmp_path):
    sc = make_scene()
    p = oracle.default_params(num_particles=sc.n, size=sc.size, grid_res=sc.grid_res)
    steps, cuts, world = 4, [0, 4, 10], 2
    ref = reference_run(oracle, sc, p, steps)
    mp.spawn(_gloo_worker, args=(world, _free_port(), cuts, steps, str(tmp_path)), nprocs=world,
             join=True)
    for s in range(steps):
        got = np.concatenate([np.load(tmp_path / f"buf1_s{s}_r{r}.npy") for r in range(world)])
        np.testing.assert_array_equal(got, ref[s][0], err_msg=f"step {s}")
