"""torchrun worker for tests/test_slab_gpu.py::test_multi_process_ranks_equal_whole_grid."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_slab_gpu import FRAME_DT, make_scene, scene_params  # noqa: E402
from watercube_b200 import capi, slab  # noqa: E402


def main():
    out_dir, steps = sys.argv[1], int(sys.argv[2])
    exchange = sys.argv[3] if len(sys.argv) > 3 else "nccl"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sc = make_scene(200000, seed=9)
    d = capi.derive(capi.default_params(num_particles=sc.n, **scene_params(sc)))
    hist = np.bincount(slab.layer_of(sc.particles[:, 2], d.bin_size, sc.grid_res),
                       minlength=sc.grid_res)
    cuts = slab.slab_cuts(hist, world)
    parts = slab.decompose(sc.particles, cuts, d.bin_size, sc.grid_res)
    b = slab.CudaSlabBackend(scene_params(sc), cuts[rank], cuts[rank + 1], capacity=sc.n,
                             ghost_capacity=sc.n, migrant_capacity=16384, device=local,
                             stream=stream.cuda_stream)
    b.upload(parts[rank])
    drv = slab.SlabDriver(b, rank, world)
    if exchange == "peer":
        slab.attach_peers_ipc(b, rank, world)
    for s in range(steps):
        if exchange == "peer":
            slab.run_step_peer(b, FRAME_DT)
        else:
            slab.run_step(drv, FRAME_DT)
        np.save(os.path.join(out_dir, f"buf1_s{s}_r{rank}.npy"), b.download(1))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
