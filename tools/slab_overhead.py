"""Slab-mode overhead without any communication: ONE slab handle covering the whole grid
(no neighbours) against the whole-grid handle on the same scene, per-stage times (ms)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from watercube_b200 import capi, scenes, slab  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
steps = 10
sc = scenes.dam_break(n, seed=0)
kw = dict(grid_res=sc.grid_res, size=sc.size, particle_radius=sc.particle_radius)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)


def avg(times):
    return {k: round(float(np.mean([t[k] for t in times])), 4) for k in times[0]}


def whole_grid():
    with capi.Fluid(num_particles=sc.n, flags=capi.FLAG_STAGE_TIMING, stream=stream.cuda_stream, **kw) as fl:
        fl.upload(sc.particles)
        for _ in range(3):
            fl.step()
        ts = []
        for _ in range(steps):
            fl.step()
            ts.append(fl.stage_times())
        print("whole-grid", avg(ts), "sum", round(sum(avg(ts).values()), 4))


layer = int(np.bincount(slab.layer_of(sc.particles[:, 2], sc.size / sc.grid_res, sc.grid_res)).max())


def one_slab(tag, capacity, migrants):
    b = slab.CudaSlabBackend(kw, 0, sc.grid_res, capacity=capacity, ghost_capacity=int(layer * 1.5) + 1024,
                             migrant_capacity=migrants, flags=capi.FLAG_STAGE_TIMING,
                             stream=stream.cuda_stream)
    b.upload(sc.particles)
    for d in (0, 1):
        b.clear_recv(d)
    for _ in range(3):
        slab.run_step_peer(b, 1 / 60.0, wait=False)
    ts = []
    for _ in range(steps):
        slab.run_step_peer(b, 1 / 60.0, wait=False)
        ts.append(b.fluid.stage_times())
    print(tag, avg(ts), "sum", round(sum(avg(ts).values()), 4), "launches/step",
          b.fluid.launch_count() / (steps + 3))
    b.close()


# interleaved, so that a clock drift over the process would show as a trend instead of a gap
for _ in range(2):
    whole_grid()
    one_slab("one slab  ", int(sc.n * 1.25) + 4 * layer + 1024, max(layer // 2, 65536))
