"""z-slab decomposition on the GPU: several slab handles (virtual ranks on one GPU, and real
ranks over NCCL when the box has >= 2 GPUs) must reproduce the whole-grid handle BIT-EXACTLY:
the per-particle accumulation order of the gather kernels does not depend on how targets are
grouped, and the migrant ordering rule keeps the stable in-cell order."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from watercube_b200 import scenes, slab

pytestmark = pytest.mark.gpu
f32 = np.float32
FRAME_DT = 1.0 / 60.0
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_scene(n=60000, seed=5):
    sc = scenes.dam_break(n, seed=seed)
    rng = np.random.default_rng(seed)
    sc.particles[:, 4:7] = rng.uniform(-30, 30, (sc.n, 3)).astype(f32)   # force migration
    return sc


def scene_params(sc):
    return dict(grid_res=sc.grid_res, size=sc.size, particle_radius=sc.particle_radius)


def whole_grid_run(capi, sc, steps, flags=0):
    out = []
    with capi.Fluid(num_particles=sc.n, flags=flags, **scene_params(sc)) as fl:
        fl.upload(sc.particles)
        for _ in range(steps):
            fl.step(FRAME_DT)
            out.append((fl.download(1), fl.download(2)))
    return out


@pytest.fixture(scope="module")
def capi():
    from watercube_b200 import capi as m

    m.lib()
    return m


@pytest.mark.parametrize("world", [1, 2, 3, 5])
@pytest.mark.parametrize("simple", [False, True], ids=["tiled", "simple"])
def test_virtual_ranks_on_one_gpu_equal_whole_grid(capi, world, simple):
    sc = make_scene()
    steps = 5
    flags = capi.FLAG_SIMPLE_KERNELS if simple else 0
    ref = whole_grid_run(capi, sc, steps, flags)
    d = capi.derive(capi.default_params(num_particles=sc.n, **scene_params(sc)))
    hist = np.bincount(slab.layer_of(sc.particles[:, 2], d.bin_size, sc.grid_res),
                       minlength=sc.grid_res)
    cuts = slab.slab_cuts(hist, world)
    parts = slab.decompose(sc.particles, cuts, d.bin_size, sc.grid_res)
    drivers = []
    for r in range(world):
        b = slab.CudaSlabBackend(scene_params(sc), cuts[r], cuts[r + 1], capacity=sc.n,
                                 ghost_capacity=sc.n, migrant_capacity=8192, flags=flags)
        b.upload(parts[r])
        drivers.append(slab.SlabDriver(b, r, world))
    migrated = 0
    for s in range(steps):
        slab.run_step_local(drivers, FRAME_DT)
        migrated += sum(dr.info["migrants_in_below"] + dr.info["migrants_in_above"]
                        for dr in drivers)
        assert sum(dr.info["n_owned"] for dr in drivers) == sc.n
        buf2 = np.concatenate([dr.backend.download(2) for dr in drivers])
        buf1 = np.concatenate([dr.backend.download(1) for dr in drivers])
        np.testing.assert_array_equal(buf2, ref[s][1], err_msg=f"sorted buffer, step {s}")
        np.testing.assert_array_equal(buf1, ref[s][0], err_msg=f"state buffer, step {s}")
    if world > 1:
        assert migrated > 0
    for dr in drivers:
        dr.backend.fluid.close()


@pytest.mark.parametrize("world", [1, 2, 4])
def test_peer_memory_exchange_virtual_ranks_equal_whole_grid(capi, world):
    """wc_slab_peer_attach: the phases push their messages into the neighbour's buffers and
    signal on the device; no host-side exchange.  Same bit-exactness as the NCCL path."""
    sc = make_scene(80000, seed=7)
    steps = 6
    ref = whole_grid_run(capi, sc, steps)
    d = capi.derive(capi.default_params(num_particles=sc.n, **scene_params(sc)))
    hist = np.bincount(slab.layer_of(sc.particles[:, 2], d.bin_size, sc.grid_res),
                       minlength=sc.grid_res)
    cuts = slab.slab_cuts(hist, world)
    parts = slab.decompose(sc.particles, cuts, d.bin_size, sc.grid_res)
    backends = []
    for r in range(world):
        b = slab.CudaSlabBackend(scene_params(sc), cuts[r], cuts[r + 1], capacity=sc.n,
                                 ghost_capacity=sc.n, migrant_capacity=8192)
        b.upload(parts[r])
        backends.append(b)
    slab.attach_peers_local(backends)
    migrated = 0
    for s in range(steps):
        slab.run_step_peer_local(backends, FRAME_DT)
        migrated += sum(b.info["migrants_in_below"] + b.info["migrants_in_above"] for b in backends)
        buf2 = np.concatenate([b.download(2) for b in backends])
        buf1 = np.concatenate([b.download(1) for b in backends])
        np.testing.assert_array_equal(buf2, ref[s][1], err_msg=f"sorted buffer, step {s}")
        np.testing.assert_array_equal(buf1, ref[s][0], err_msg=f"state buffer, step {s}")
    if world > 1:
        assert migrated > 0
    for b in backends:
        b.fluid.close()


def test_rebalance_mid_run_equals_whole_grid(capi):
    """Cuts refreshed in the middle of a run (slab.rebalance_local): fresh handles over the re-cut
    state, peers re-attached; every step before and after still equals the whole-grid run."""
    sc = make_scene(80000, seed=9)
    steps, world = 8, 3
    ref = whole_grid_run(capi, sc, steps)
    d = capi.derive(capi.default_params(num_particles=sc.n, **scene_params(sc)))
    cuts = [0, 3, 6, sc.grid_res]                                 # lopsided on purpose
    parts = slab.decompose(sc.particles, cuts, d.bin_size, sc.grid_res)
    make = lambda z0, z1: slab.CudaSlabBackend(scene_params(sc), z0, z1, capacity=sc.n,
                                               ghost_capacity=sc.n, migrant_capacity=8192)
    backends = []
    for r in range(world):
        backends.append(make(cuts[r], cuts[r + 1]))
        backends[-1].upload(parts[r])
    slab.attach_peers_local(backends)
    for s in range(steps):
        if s == 4:
            before = [b.num_particles for b in backends]
            cuts, backends = slab.rebalance_local(backends, make, d.bin_size, sc.grid_res)
            slab.attach_peers_local(backends)
            after = [b.num_particles for b in backends]
            assert sum(after) == sc.n and max(after) - min(after) < max(before) - min(before)
        slab.run_step_peer_local(backends, FRAME_DT)
        buf1 = np.concatenate([b.download(1) for b in backends])
        np.testing.assert_array_equal(buf1, ref[s][0], err_msg=f"state buffer, step {s}")
    for b in backends:
        b.close()


def _attached_backends(sc, world, d, **kw):
    hist = np.bincount(slab.layer_of(sc.particles[:, 2], d.bin_size, sc.grid_res),
                       minlength=sc.grid_res)
    cuts = slab.slab_cuts(hist, world)
    parts = slab.decompose(sc.particles, cuts, d.bin_size, sc.grid_res)
    backends = []
    for r in range(world):
        args = dict(capacity=sc.n, ghost_capacity=sc.n, migrant_capacity=8192)
        args.update(kw)
        b = slab.CudaSlabBackend(scene_params(sc), cuts[r], cuts[r + 1], **args)
        b.upload(parts[r])
        backends.append(b)
    slab.attach_peers_local(backends)
    return backends


@pytest.mark.parametrize("world", [2, 3])
def test_async_steps_one_host_thread_equal_whole_grid(capi, world):
    """wc_slab_step_peer(info = NULL): whole steps queued handle after handle by ONE host thread,
    several steps ahead, with no host wait anywhere inside -- what core::Fluid::devices and
    wc_headless --gpus N do.  The counts of a step exist on the device only; the host reads them
    when it downloads.  Same bits as the whole-grid run."""
    sc = make_scene(80000, seed=11)
    steps = 14   # far more launches than a stream's queue holds: the library bounds the run-ahead
    ref = whole_grid_run(capi, sc, steps)
    d = capi.derive(capi.default_params(num_particles=sc.n, **scene_params(sc)))
    backends = _attached_backends(sc, world, d)
    slab.run_steps_peer_async(backends, FRAME_DT, steps=3)         # three steps without a look
    buf1 = np.concatenate([b.download(1) for b in backends])
    np.testing.assert_array_equal(buf1, ref[2][0])
    slab.run_steps_peer_async(backends, FRAME_DT, steps=steps - 3)
    assert sum(b.num_particles for b in backends) == sc.n
    np.testing.assert_array_equal(np.concatenate([b.download(2) for b in backends]), ref[-1][1])
    np.testing.assert_array_equal(np.concatenate([b.download(1) for b in backends]), ref[-1][0])
    for b in backends:
        b.close()


def test_step_peer_host_equals_upload_step_download(capi):
    """wc_slab_step_peer_host (host AoS in and out per rank, the update kernel storing into the
    page-locked output itself) equals upload + step + download, chained over several steps."""
    import torch

    sc = make_scene(60000, seed=13)
    world, steps = 2, 3
    ref = whole_grid_run(capi, sc, steps)
    d = capi.derive(capi.default_params(num_particles=sc.n, **scene_params(sc)))
    backends = _attached_backends(sc, world, d)
    cap = sc.n
    h_in = [torch.empty((cap, 8), dtype=torch.float32, pin_memory=True) for _ in range(world)]
    h_out = [torch.full((cap, 8), float("nan"), dtype=torch.float32, pin_memory=True)
             for _ in range(world)]
    n_cur = []
    for r, b in enumerate(backends):
        cur = b.download(1)
        n_cur.append(cur.shape[0])
        h_in[r][:n_cur[r]].copy_(torch.from_numpy(cur))
    for s in range(steps):
        # one host thread, synchronous calls: queue every rank's step first (async), then the
        # host round trip of each rank would deadlock on its neighbour -- so ranks run in threads
        import threading

        infos = [None] * world

        def run(r):
            infos[r] = backends[r].fluid.slab_step_peer_host((h_in[r].data_ptr(), n_cur[r]),
                                                            h_out[r].data_ptr(), cap, FRAME_DT)

        ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        got = np.concatenate([h_out[r][:infos[r]["n_owned"]].numpy() for r in range(world)])
        np.testing.assert_array_equal(got, ref[s][0], err_msg=f"step {s}")
        n_cur = [infos[r]["n_owned"] for r in range(world)]
        h_in, h_out = h_out, h_in
    for b in backends:
        b.close()


def test_overflow_in_an_async_step_is_sticky_and_harmless(capi):
    """A capacity overflow inside an asynchronous step cannot be reported by that call: the step
    turns itself off on the device (no index derived from the counts is used), the error word
    stays set, and the host meets it at its next look."""
    sc = make_scene(40000, seed=15)
    d = capi.derive(capi.default_params(num_particles=sc.n, **scene_params(sc)))
    backends = _attached_backends(sc, 2, d, ghost_capacity=64)    # far too small a halo
    slab.run_steps_peer_async(backends, FRAME_DT, steps=2)
    with pytest.raises(capi.WcError):
        backends[0].fluid.slab_step_peer(FRAME_DT)                 # synchronous: reports
    for b in backends:
        b.close()


def test_slab_capacity_overflow_is_reported(capi):
    sc = make_scene(20000)
    b = slab.CudaSlabBackend(scene_params(sc), 0, sc.grid_res, capacity=sc.n, ghost_capacity=16,
                             migrant_capacity=16)
    b.upload(sc.particles)
    drv = slab.SlabDriver(b, 0, 1)
    slab.run_step_local([drv], FRAME_DT)           # one rank: nothing to overflow
    with pytest.raises(capi.WcError):
        b.fluid.step(FRAME_DT)                     # whole-grid call on a slab handle
    b.fluid.close()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("exchange", ["nccl", "peer"])
def test_multi_process_ranks_equal_whole_grid(capi, tmp_path, exchange):
    """One process per GPU: NCCL P2P exchange, and the CUDA-IPC peer-memory exchange."""
    import torch

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    steps = 4
    sc = make_scene(200000, seed=9)
    ref = whole_grid_run(capi, sc, steps)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port",
           str(_free_port()), os.path.join(ROOT, "tests", "slab_nccl_worker.py"), str(tmp_path),
           str(steps), exchange]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    for s in range(steps):
        got = np.concatenate([np.load(tmp_path / f"buf1_s{s}_r{r}.npy") for r in range(world)])
        np.testing.assert_array_equal(got, ref[s][0], err_msg=f"step {s}")
