"""z-slab decomposition of the SPH step across the GPUs of one node.

New work with no reference counterpart (the reference is single-GPU).  The cell index is
z-major (assets/sort/count.comp:33), so a rank that owns the z-layers [z_begin, z_end) owns a
contiguous range of bins and, after the sort, a contiguous slice of the particle array; its
first / last layer -- the halo its neighbours need -- are contiguous sub-slices and are sent
straight out of the sorted buffer (no pack kernel).

One step is a fixed sequence of compute phases (C-ABI ``wc_slab_*`` calls) separated by four
neighbour exchanges.  :class:`SlabDriver` expresses the sequence as a generator of exchange
requests, so the same logic runs over

* ``torch.distributed`` (NCCL on the GPUs, gloo in the CPU tests):  :func:`run_step`;
* several "virtual ranks" inside one process (tests on a single GPU):  :func:`run_step_local`.

The compute backend is pluggable: :class:`CudaSlabBackend` drives the native library;
tests plug a CPU stand-in (tests/slab_cpu_backend.py) to check the decomposition logic
without a GPU.  There is no CPU backend in the product.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np


# --------------------------------------------------------------------------- partitioning
def layer_of(z: np.ndarray, bin_size: float, grid_res: int) -> np.ndarray:
    """Global z-layer of a coordinate: count.comp:32 (fp32 divide, truncate, clamp)."""
    q = np.asarray(z, np.float32) / np.float32(bin_size)
    with np.errstate(invalid="ignore"):
        c = np.where(q >= 1.0, np.minimum(q, np.float32(grid_res - 1)), 0.0)
    return np.nan_to_num(c, nan=0.0).astype(np.int64).clip(0, grid_res - 1)


def slab_cuts(layer_hist: Sequence[int], world: int) -> List[int]:
    """z-cuts [z_0=0, ..., z_world=G] giving every rank >= 1 layer and roughly equal particle
    counts (a dam break is strongly imbalanced under equal-width slabs)."""
    hist = np.asarray(layer_hist, np.int64)
    G = len(hist)
    if world > G:
        raise ValueError(f"{world} slabs need at least {world} z-layers, grid has {G}")
    cum = np.concatenate([[0], np.cumsum(hist)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        z = int(np.searchsorted(cum, target, side="left"))
        if z > 0 and abs(cum[z - 1] - target) <= abs(cum[min(z, G)] - target):
            z -= 1
        z = max(z, cuts[-1] + 1)            # at least one layer per rank ...
        z = min(z, G - (world - r))         # ... and enough layers left for the others
        cuts.append(z)
    cuts.append(G)
    return cuts


def decompose(particles: np.ndarray, cuts: Sequence[int], bin_size: float, grid_res: int):
    """Split an [n, 8] AoS array by slab, preserving order inside each slab (so the ranks'
    inputs concatenate, per cell, to the single-GPU input order)."""
    lay = layer_of(particles[:, 2], bin_size, grid_res)
    rank_of = np.searchsorted(np.asarray(cuts[1:-1]), lay, side="right")
    return [np.ascontiguousarray(particles[rank_of == r]) for r in range(len(cuts) - 1)]


# --------------------------------------------------------------------------- re-balancing
# the cuts equalise particle counts from the per-layer histogram and are
# refreshed every k steps, because the fluid moves along z.  Concatenating the ranks' buffer 1
# (download(1)) in rank order IS the whole-grid state in the whole-grid order, and a cell
# belongs to exactly one old rank, so re-cutting that concatenation and handing every new rank
# its layers -- pieces taken in old-rank order -- continues the run bit-identically to one that
# was never re-cut.  Pending migrant messages are duplicates of particles still present in the
# senders' buffer 1; the fresh handles drop them.
def recut_plan(particles: np.ndarray, layer_hist_global: np.ndarray, world: int, bin_size: float,
               grid_res: int):
    """-> (new cuts, destination rank of every particle of this piece)."""
    cuts = slab_cuts(layer_hist_global, world)
    lay = layer_of(particles[:, 2], bin_size, grid_res)
    return cuts, np.searchsorted(np.asarray(cuts[1:-1]), lay, side="right")


def rebalance_local(backends: Sequence, make_backend, bin_size: float, grid_res: int):
    """Virtual ranks of one process: -> (new cuts, new backends).  make_backend(z0, z1) builds an
    empty backend for the layers [z0, z1); the old ones are closed."""
    world = len(backends)
    state = np.concatenate([b.download(1) for b in backends])
    hist = np.bincount(layer_of(state[:, 2], bin_size, grid_res), minlength=grid_res)
    cuts = slab_cuts(hist, world)
    parts = decompose(state, cuts, bin_size, grid_res)
    for b in backends:
        b.close()
    fresh = []
    for r in range(world):
        nb = make_backend(cuts[r], cuts[r + 1])
        nb.upload(parts[r])
        fresh.append(nb)
    return cuts, fresh


def rebalance(backend, rank: int, world: int, make_backend, bin_size: float, grid_res: int,
              device=None, group=None):
    """One process per rank (torch.distributed): every rank calls this at the same step.
    -> (new cuts, new backend).  The particles travel as one all-to-all of AoS records; the
    histogram as one all-reduce.  Infrequent (every k >> 1 steps), so it goes through
    download / upload rather than staying on the device."""
    import torch
    import torch.distributed as dist

    mine = backend.download(1)
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    hist = torch.from_numpy(np.bincount(layer_of(mine[:, 2], bin_size, grid_res),
                                        minlength=grid_res).astype(np.int64)).to(dev)
    dist.all_reduce(hist, group=group)
    cuts, dest = recut_plan(mine, hist.cpu().numpy(), world, bin_size, grid_res)
    order = np.argsort(dest, kind="stable")                  # by new owner, order kept inside
    send_counts = np.bincount(dest, minlength=world).astype(np.int64)
    sc = torch.from_numpy(send_counts).to(dev)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = rc.cpu().numpy()
    send = torch.from_numpy(np.ascontiguousarray(mine[order])).to(dev).reshape(-1)
    recv = torch.empty(int(recv_counts.sum()) * 8, dtype=torch.float32, device=dev)
    dist.all_to_all_single(recv, send, [int(c) * 8 for c in recv_counts],
                           [int(c) * 8 for c in send_counts], group=group)
    backend.close()
    nb = make_backend(cuts[rank], cuts[rank + 1])
    nb.upload(recv.cpu().numpy().reshape(-1, 8))             # pieces arrive in old-rank order
    return cuts, nb


# --------------------------------------------------------------------------- driver
@dataclass
class Xfer:
    """One neighbour transfer: `send` goes to, and `recv` comes from, the rank below
    (direction 0) or above (direction 1).  Tensors are flat views of the exact size."""
    direction: int
    send: object
    recv: object


class SlabDriver:
    """The per-step phase sequence of one rank."""

    def __init__(self, backend, rank: int, world: int):
        self.backend, self.rank, self.world = backend, rank, world
        self.info = None
        for d in (0, 1):
            if self.peer(d) is None:
                backend.clear_recv(d)

    def peer(self, direction: int):
        p = self.rank - 1 if direction == 0 else self.rank + 1
        return p if 0 <= p < self.world else None

    def step_phases(self, frame_dt: float):
        b = self.backend
        # 1. unpack last step's migrants, hash + count + scan of the owned layers
        b.sort_count()
        yield [Xfer(d, b.lc_send(d), b.lc_recv(d)) for d in (0, 1)]
        # 2. the only host sync of the step: this step's particle counts
        self.info = b.sync_info()
        if self.info["errors"]:
            raise RuntimeError(f"rank {self.rank}: slab capacity overflow / lost migrants: "
                               f"{self.info}")
        b.reorder()
        # 3. halo positions (density needs x,y,z of the neighbours' boundary layers)
        yield [Xfer(d, b.halo_send(d, "pos"), b.halo_recv(d, "pos")) for d in (0, 1)]
        b.density()
        # 4. halo density / pressure / velocity (the force needs rho_j, P_j, v_j)
        yield [Xfer(d, b.halo_send(d, k), b.halo_recv(d, k)) for d in (0, 1) for k in ("pos", "vel")]
        # 5. force + integrate, extraction of the particles that left the slab
        b.update(frame_dt)
        yield [Xfer(d, b.mig_out(d), b.mig_in(d)) for d in (0, 1)]


def run_step(driver: SlabDriver, frame_dt: float, group=None):
    """One step over torch.distributed (NCCL for CUDA tensors, gloo for CPU tensors)."""
    import torch.distributed as dist

    for xfers in driver.step_phases(frame_dt):
        ops = []
        for x in xfers:
            peer = driver.peer(x.direction)
            if peer is None:
                continue
            if x.send.numel():
                ops.append(dist.P2POp(dist.isend, x.send, peer, group))
            if x.recv.numel():
                ops.append(dist.P2POp(dist.irecv, x.recv, peer, group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()


# --------------------------------------------------------------------------- peer-memory mode
# With the neighbours attached (CudaSlabBackend.attach_peers_*), the native phases move the
# four messages themselves -- copies into the neighbour's mapped buffers plus device-side
# signals (include/wc_sph.h, wc_slab_peer_*) -- so a step is just the five phase calls.
def attach_peers_ipc(backend, rank: int, world: int, group=None):
    """One process per GPU: trade the IPC blobs over torch.distributed, open the neighbours."""
    import torch.distributed as dist

    blobs = [None] * world
    dist.all_gather_object(blobs, backend.fluid.slab_ipc_export(), group=group)
    for d, peer in ((0, rank - 1), (1, rank + 1)):
        if 0 <= peer < world:
            backend.fluid.slab_peer_open(d, blobs[peer])
        else:
            backend.clear_recv(d)
    dist.barrier(group=group)


def attach_peers_local(backends: Sequence["CudaSlabBackend"]):
    """Several slab handles in this process (virtual ranks, or one process driving many GPUs)."""
    for r, b in enumerate(backends):
        for d, peer in ((0, r - 1), (1, r + 1)):
            if 0 <= peer < len(backends):
                b.fluid.slab_peer_attach(d, backends[peer].fluid)
            else:
                b.clear_recv(d)


def run_step_peer(backend, frame_dt: float, wait: bool = True):
    """One step of one rank in peer-memory mode (every rank runs this; no host exchange): the
    five phases as ONE C-ABI call, wc_slab_step_peer.  wait=False only queues the step -- the
    host does not wait anywhere in it; capacity errors then surface at the next info read."""
    if hasattr(backend, "fluid"):
        backend.info = backend.fluid.slab_step_peer(frame_dt, wait=wait)  # raises on errors
        return
    backend.sort_count()
    info = backend.sync_info()
    if info["errors"]:
        raise RuntimeError(f"slab capacity overflow / lost migrants: {info}")
    backend.reorder()
    backend.density()
    backend.update(frame_dt)


def run_steps_peer_async(backends: Sequence["CudaSlabBackend"], frame_dt: float, steps: int = 1):
    """Several slab handles driven by ONE host thread (virtual ranks, or one process driving
    every GPU of the box): whole steps are queued handle after handle -- nothing in a step waits
    for the host, and a kernel that waits for a neighbour's signal only ever waits for work that
    is already queued or will be queued without anybody waiting for it."""
    for _ in range(steps):
        for b in backends:
            b.fluid.slab_step_peer(frame_dt, wait=False)


def run_step_peer_local(backends: Sequence["CudaSlabBackend"], frame_dt: float):
    """Virtual ranks of one process in peer-memory mode: phase by phase over all ranks, so every
    push a wait depends on has been queued before the host blocks in sync_info."""
    for b in backends:
        b.sort_count()
    for b in backends:
        if b.sync_info()["errors"]:
            raise RuntimeError(f"slab capacity overflow / lost migrants: {b.info}")
    for b in backends:
        b.reorder()
    for b in backends:
        b.density()
    for b in backends:
        b.update(frame_dt)


def run_step_local(drivers: Sequence[SlabDriver], frame_dt: float):
    """One step of several virtual ranks living in this process (lock-step, copies instead of
    messages).  Used to test the decomposition on a single GPU."""
    gens = [d.step_phases(frame_dt) for d in drivers]
    while True:
        batches = [next(g, None) for g in gens]
        if all(b is None for b in batches):
            return
        assert all(b is not None for b in batches)
        for d in drivers:
            d.backend.sync()
        for r, xfers in enumerate(batches):
            for direction in (0, 1):
                peer = drivers[r].peer(direction)
                if peer is None:
                    continue
                mine = [x for x in xfers if x.direction == direction]
                theirs = [x for x in batches[peer] if x.direction == 1 - direction]
                assert len(mine) == len(theirs)
                for a, b in zip(mine, theirs):
                    assert a.send.numel() == b.recv.numel(), (a.send.shape, b.recv.shape)
                    if a.send.numel():
                        b.recv.copy_(a.send)
        for d in drivers:
            d.backend.sync(after_copy=True)


# --------------------------------------------------------------------------- CUDA backend
class _CudaBuffer:
    """Zero-copy torch view of library-owned device memory (__cuda_array_interface__)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1",
                                         "data": (int(ptr), False), "version": 2}


def _view(ptr, nbytes, device):
    import torch

    if nbytes == 0:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_CudaBuffer(ptr, nbytes), device=device)


class CudaSlabBackend:
    """The native library in slab mode behind the driver's backend interface."""

    def __init__(self, scene_params: dict, z_begin: int, z_end: int, capacity: int,
                 ghost_capacity: int, migrant_capacity: int, device: int = 0, flags: int = 0,
                 stream=None, **step_kw):
        import torch

        from . import capi

        self.torch = torch
        self.device = torch.device("cuda", device)
        if stream is None:
            # run_step() issues the NCCL sends / receives on torch's current stream against
            # buffers the wc_slab_* kernels produce and consume: both must be ONE stream (a
            # private library stream would leave the two unordered)
            stream = torch.cuda.current_stream(self.device).cuda_stream
        self.fluid = capi.Fluid(num_particles=0, capacity=capacity, device=device, flags=flags,
                                stream=stream,
                                slab=(z_begin, z_end, ghost_capacity, migrant_capacity),
                                **scene_params, **step_kw)
        v = self.fluid.slab_view()
        self.first = int(v.owned_first)
        slots = capacity + 2 * ghost_capacity
        self._mig_out = [_view(v.mig_out[d], v.mig_bytes, self.device) for d in (0, 1)]
        self._mig_in = [_view(v.mig_in[d], v.mig_bytes, self.device) for d in (0, 1)]
        self._lc_send = [_view(v.lc_send[d], v.lc_bytes, self.device) for d in (0, 1)]
        self._lc_recv = [_view(v.lc_recv[d], v.lc_bytes, self.device) for d in (0, 1)]
        self._sorted = {"pos": _view(v.pos_rho_sorted, slots * 16, self.device).view(slots, 16),
                        "vel": _view(v.vel_pres_sorted, slots * 16, self.device).view(slots, 16)}
        self.info = None

    # -- data in / out
    def upload(self, particles: np.ndarray):
        self.fluid.upload(particles)

    def download(self, which=1) -> np.ndarray:
        return self.fluid.download(which)

    @property
    def num_particles(self) -> int:
        return self.fluid.num_particles

    # -- compute phases
    def sort_count(self):
        self.fluid.slab_sort_count()

    def sync_info(self) -> dict:
        self.info = self.fluid.slab_sync_info()
        return self.info

    def reorder(self):
        self.fluid.slab_reorder()

    def density(self):
        self.fluid.slab_density()

    def update(self, frame_dt):
        self.fluid.slab_update(frame_dt)

    def sync(self, after_copy=False):
        if after_copy:
            self.torch.cuda.synchronize(self.device)
        else:
            self.fluid.sync()

    def close(self):
        self.fluid.close()

    # -- exchange buffers
    def clear_recv(self, direction):
        self.fluid.slab_clear_recv(direction)

    def lc_send(self, d):
        return self._lc_send[d]

    def lc_recv(self, d):
        return self._lc_recv[d]

    def mig_out(self, d):
        return self._mig_out[d]

    def mig_in(self, d):
        return self._mig_in[d]

    def halo_send(self, d, kind):
        i, f = self.info, self.first
        lo, hi = (f, f + i["n_first"]) if d == 0 else (f + i["n_owned"] - i["n_last"],
                                                      f + i["n_owned"])
        return self._sorted[kind][lo:hi].reshape(-1)

    def halo_recv(self, d, kind):
        i, f = self.info, self.first
        lo, hi = (f - i["n_ghost_below"], f) if d == 0 else (
            f + i["n_owned"], f + i["n_owned"] + i["n_ghost_above"])
        return self._sorted[kind][lo:hi].reshape(-1)
