"""Small scenes for compute-sanitizer (tools/sanitize.sh): the whole-grid step -- including a
crowded cell (the radix reorder), a group whose cull over-reads past its slice, the pre-hash
of the second step and a third step with the extended physics switched on -- and three
virtual ranks with peer memory attached (remote stores + device-side signals)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from watercube_b200 import capi, scenes, slab  # noqa: E402

FRAME_DT = 1.0 / 60.0
which = sys.argv[1] if len(sys.argv) > 1 else "all"

if which in ("all", "whole"):
    sc = scenes.dam_break(20000, seed=3)
    sc.particles[:400, :3] = sc.particles[0, :3]          # 400 particles in one cell
    with capi.Fluid(num_particles=sc.n, grid_res=sc.grid_res, size=sc.size,
                    particle_radius=sc.particle_radius, flags=capi.FLAG_DEBUG_OUTPUTS) as fl:
        fl.upload(sc.particles)
        for _ in range(2):
            fl.step(FRAME_DT)
        fl.set_physics(capi.PHYS_WALL_PARTICLES | capi.PHYS_SURFACE_TENSION)
        fl.step(FRAME_DT)
        out = fl.download(1)
        fl.diagnose(1)
    assert np.isfinite(out).all()
    print("whole-grid ok", out.shape)

if which in ("all", "slab"):
    sc = scenes.dam_break(20000, seed=5)
    rng = np.random.default_rng(5)
    sc.particles[:, 4:7] = rng.uniform(-30, 30, (sc.n, 3)).astype(np.float32)
    kw = dict(grid_res=sc.grid_res, size=sc.size, particle_radius=sc.particle_radius)
    d = capi.derive(capi.default_params(num_particles=sc.n, **kw))
    hist = np.bincount(slab.layer_of(sc.particles[:, 2], d.bin_size, sc.grid_res),
                       minlength=sc.grid_res)
    cuts = slab.slab_cuts(hist, 3)
    parts = slab.decompose(sc.particles, cuts, d.bin_size, sc.grid_res)
    backends = []
    for r in range(3):
        b = slab.CudaSlabBackend(kw, cuts[r], cuts[r + 1], capacity=sc.n, ghost_capacity=sc.n,
                                 migrant_capacity=4096)
        b.upload(parts[r])
        backends.append(b)
    slab.attach_peers_local(backends)
    for _ in range(2):
        slab.run_step_peer_local(backends, FRAME_DT)
    n = sum(b.download(1).shape[0] for b in backends)
    assert n == sc.n, n
    for b in backends:
        b.close()
    print("slab ok", n)
