"""Run under torchrun with >= 2 ranks (one GPU each): the film after the REAL NCCL reduce of the ranks'
fixed-point accumulators must equal the 1-GPU film bit for bit, for the spp-slice and the tile-slice
partition.  Rank 0 prints "NCCL_FILM_OK <world>" on success; any rank exits non-zero on a mismatch.
Used by tests/test_gpu_parity.py::test_nccl_reduced_film_equals_the_single_gpu_film and by hand:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/tools/nccl_film_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import raytracingrenderer_b200 as rtb  # noqa: E402
from raytracingrenderer_b200 import abi, distributed as D  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import refbvh
    scenes = [("synthetic", refbvh.random_scene(3, 64, 96, 64, True), 11),
              ("cornell-box", abi.FlatScene.load(os.path.join(ROOT, "tests", "golden", "cornell-box.rtbs")), 5)]
    for name, scene, spp in scenes:
        rt = rtb.RayTracer(local)
        rt.set_stream(torch.cuda.current_stream().cuda_stream)
        rt.init(scene)
        # the single-GPU film, computed by every rank on its own
        rt.render(spp, 0)
        full = rt.read_film().copy()
        acc = D.accum_tensor(rt)
        full_acc = acc.clone()
        for mode in ("spp", "tile"):
            rt.set_params(**D.partition_params(rank, world, mode))
            rt.clear()
            rt.render(spp, 0)
            D.reduce_film(rt, spp)           # dist.reduce(SUM, int64) over NCCL onto rank 0
            torch.cuda.synchronize()
            if rank == 0:
                if not torch.equal(acc, full_acc):
                    print("rank 0: %s/%s accumulators differ in %d entries" % (name, mode, int((acc != full_acc).sum())), flush=True)
                    sys.exit(3)
                film = rt.read_film()
                if film.tobytes() != full.tobytes() or rt.getSPP() != spp:
                    print("rank 0: %s/%s film differs" % (name, mode), flush=True)
                    sys.exit(4)
        rt.close()
    dist.barrier()
    if rank == 0:
        print("NCCL_FILM_OK %d" % world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
