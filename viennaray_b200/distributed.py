"""Multi-GPU plumbing of the flux path: one process per GPU, rays sharded by
contiguous ray-index range, scene and BVH replicated, ONE all-reduce(sum) over
the fixed-point flux words and the TraceInfo counters.

The reference has no distributed component (single process, OpenMP threads,
rayTraceKernel.hpp:87-118); what makes the partition exact here is that a ray's
whole walk depends only on (seed, particle stream, ray index) -- the same
property the reference relies on across threads (rayTraceKernel.hpp:120-121)
-- and that flux is accumulated in integers, so the union of the shards is
bit-identical to the single-GPU job at any world size.
"""
import torch
import torch.distributed as dist


def shard_bounds(num_rays, rank, world):
    """[begin, end) of `rank`'s contiguous slice of ray indices 0..num_rays-1.
    Slices differ by at most one ray and tile the range exactly."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank %r outside world of %r" % (rank, world))
    base, rem = divmod(int(num_rays), world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def step_shard(step, rays_per_rank, rank, world):
    """bench.py's weak-scaling layout: step k of the job hands every rank its
    own `rays_per_rank` indices; steps and ranks never overlap."""
    begin = (step * world + rank) * int(rays_per_rank)
    return begin, begin + int(rays_per_rank)


def as_int64_tensor(device_ptr, num_words, device):
    """Zero-copy int64 view of the context's result words (vr_flux_device)."""
    iface = {"shape": (int(num_words),), "typestr": "<i8", "data": (int(device_ptr), False),
             "version": 3}
    holder = type("VrResultWords", (), {"__cuda_array_interface__": iface})()
    return torch.as_tensor(holder, device=device)


def all_reduce_words(words, group=None):
    """Sum of the fixed-point flux words + counters over all ranks, in place.
    `words` is an int64 tensor (uint64 sums reinterpret exactly: two's-complement
    addition is the same operation)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(words, op=dist.ReduceOp.SUM, group=group)
    return words
