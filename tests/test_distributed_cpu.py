"""World-size-2 (and 3) gloo runs of the multi-GPU host logic on CPU: the partition each rank
is given, and the film read-out collective, compose to the single-process film.  The renders
themselves are done by the oracle here (no GPU); the GPU-side partitions are checked against
the same property in tests/test_gpu_parity.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, synthetic_scene


def _worker(rank, world, mode, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import port as oracle_port
    from raytracingrenderer_b200 import distributed as D
    scene = synthetic_scene()
    spp = 5
    params = D.partition_params(rank, world, mode)
    film, st = oracle_port.Oracle(scene, **params).render(spp, threads=2)
    # the same fixed-point representation the GPU film uses (units of 2^-32)
    acc = torch.from_numpy(np.rint(film.astype(np.float64) * 2.0 ** 32).astype(np.int64))
    D.reduce_sum_(acc, dst=0)
    n = torch.tensor([st["samples"]])
    D.reduce_sum_(n, dst=0)
    if mode == "spp":
        assert st["samples"] == scene.width * scene.height * D.local_sample_count(0, spp, rank, world)
    if rank == 0:
        np.save(os.path.join(out_dir, "acc.npy"), acc.numpy())
        np.save(os.path.join(out_dir, "n.npy"), n.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, "spp"), (2, "tile"), (3, "tile")])
def test_partitioned_render_reduces_to_the_full_film(tmp_path, oracle_mod, world, mode):
    port = 29500 + (os.getpid() % 500) + world * 7 + (0 if mode == "spp" else 3)
    mp.spawn(_worker, args=(world, mode, port, str(tmp_path)), nprocs=world, join=True)
    scene = synthetic_scene()
    full, st = oracle_mod.Oracle(scene).render(5, threads=2)
    acc = np.load(os.path.join(str(tmp_path), "acc.npy"))
    n = np.load(os.path.join(str(tmp_path), "n.npy"))
    assert int(n[0]) == st["samples"]
    got = (acc.astype(np.float64) / 2.0 ** 32).reshape(full.shape)
    if mode == "tile":
        # disjoint pixels: every pixel comes from exactly one rank
        assert np.array_equal(np.rint(full.astype(np.float64) * 2.0 ** 32).astype(np.int64).ravel(), acc.ravel())
    else:
        assert np.allclose(got, full, rtol=1e-6, atol=1e-7)


def test_partition_params_and_sample_counts():
    from raytracingrenderer_b200 import abi, distributed as D
    assert D.partition_params(0, 1)["partition"] == abi.PART_NONE
    assert D.partition_params(3, 8)["partition"] == abi.PART_SPP and D.partition_params(3, 8)["part_rank"] == 3
    assert D.partition_params(1, 2, "tile")["partition"] == abi.PART_TILE
    for begin, count, world in ((0, 256, 8), (5, 13, 4), (7, 3, 8), (0, 1, 2)):
        assert sum(D.local_sample_count(begin, count, r, world) for r in range(world)) == count
