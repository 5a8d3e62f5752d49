"""Cost of the extended physics (wc_physics) per stage: the same scene stepped with flags
0 (the reference's step), 1 (wall particles), 2 (surface tension), 3 (both).  Stage times, ms."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from watercube_b200 import capi, scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sc = scenes.dam_break(n, seed=0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
for flags in (0, 1, 2, 3, 0):
    with capi.Fluid(num_particles=sc.n, grid_res=sc.grid_res, size=sc.size,
                    particle_radius=sc.particle_radius, flags=capi.FLAG_STAGE_TIMING,
                    stream=stream.cuda_stream) as fl:
        fl.set_physics(flags)
        fl.upload(sc.particles)
        for _ in range(5):
            fl.step()
        ts = []
        for _ in range(20):
            fl.step()
            ts.append(fl.stage_times())
        avg = {k: round(float(np.mean([t[k] for t in ts])), 4) for k in ts[0]}
        print("flags", flags, avg, "sum", round(sum(avg.values()), 4))
