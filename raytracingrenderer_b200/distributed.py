"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink; gloo in the
CPU tests).  The path-tracing loop has no exchange step: every pixel sample is independent and
the scene is replicated, so ranks only meet at the film read-out:

  spp slice   rank r renders the global sample indices s with s % world == r   (default)
  tile slice  rank r renders the 32x32 tiles t (TILE_SIZE, RTBase/Renderer.h:18) with
              t % world == r

and the film is ONE sum-reduce of the per-rank fixed-point accumulators (int64, units of 2^-32;
see rtb_accum_device_ptr).  Integer addition is associative, so the reduced film is bit-identical
to the single-GPU film for both partitions, whatever the reduction tree NCCL picks.
"""
import torch
import torch.distributed as dist

from . import abi


def partition_params(rank, world, mode="spp"):
    """rtb_params fields for this rank."""
    if world <= 1:
        return dict(partition=abi.PART_NONE, part_rank=0, part_world=1)
    part = {"spp": abi.PART_SPP, "tile": abi.PART_TILE}[mode]
    return dict(partition=part, part_rank=int(rank), part_world=int(world))


def local_sample_count(spp_begin, spp_count, rank, world):
    """How many of the global sample indices [spp_begin, spp_begin+spp_count) an spp-slice rank renders."""
    if world <= 1:
        return spp_count
    first = spp_begin + (rank + world - (spp_begin % world)) % world
    end = spp_begin + spp_count
    return (end - 1 - first) // world + 1 if first < end else 0


class _CudaArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def accum_tensor(rt):
    """Zero-copy int64 torch view of a RayTracer's fixed-point film sums."""
    ptr, n = rt.accum_device_ptr()
    return torch.as_tensor(_CudaArray(ptr, n, "<i8"), device="cuda:%d" % rt.device)


def film_tensor(rt):
    """Zero-copy float32 torch view of a RayTracer's film (re-derived from the sums if stale)."""
    ptr, n = rt.film_device_ptr()
    return torch.as_tensor(_CudaArray(ptr, n, "<f4"), device="cuda:%d" % rt.device)


def reduce_sum_(t, dst=0, group=None):
    """In-place SUM of `t` over the ranks onto `dst` (no-op without a process group).
    Works for the int64 accumulators on NCCL and for CPU tensors on gloo."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return t


def reduce_film(rt, total_spp, dst=0, group=None):
    """The film read-out collective: sum the ranks' accumulators onto `dst`, then tell the
    context the global sample count (Film::SPP).  Returns the accumulator view."""
    acc = accum_tensor(rt)
    reduce_sum_(acc, dst=dst, group=group)
    rt.accum_device_ptr()      # marks the float film stale on every rank
    rt.set_spp(total_spp)
    return acc
