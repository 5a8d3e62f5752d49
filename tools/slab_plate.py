"""One rank's slab of the N-GPU dam break, ALONE on one GPU (no neighbours: its two z faces
become free surfaces): separates what the slab's SHAPE costs -- a thin plate with the big
scene's cross-section -- from what the exchange costs.  Stage times in ms.
    python tools/slab_plate.py [world] [rank]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from watercube_b200 import capi, slab, slab_bench  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n_total = 8_000_000 * world
params, cuts, hist, mine = slab_bench.make_rank_scene(n_total, rank, world)
layer = int(hist.max())
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
b = slab.CudaSlabBackend(params, cuts[rank], cuts[rank + 1], capacity=int(mine.shape[0] * 1.25) + 4 * layer + 1024,
                         ghost_capacity=int(layer * 1.5) + 1024, migrant_capacity=max(layer // 2, 65536),
                         flags=capi.FLAG_STAGE_TIMING, stream=stream.cuda_stream)
b.upload(mine)
for d in (0, 1):
    b.clear_recv(d)
for _ in range(3):
    slab.run_step_peer(b, 1 / 60.0, wait=False)
ts = []
for _ in range(10):
    slab.run_step_peer(b, 1 / 60.0, wait=False)
    ts.append(b.fluid.stage_times())
avg = {k: round(float(np.mean([t[k] for t in ts])), 4) for k in ts[0]}
print(f"world {world} rank {rank}: {mine.shape[0]} particles, layers {cuts[rank]}..{cuts[rank + 1]} of "
      f"G={params['grid_res']}", avg, "sum", round(sum(avg.values()), 4),
      "per 8M:", round(sum(avg.values()) * 8e6 / mine.shape[0], 4))
b.close()
