"""Multi-GPU sharding of the elementwise path: one process per GPU, contiguous shards.

Every element is independent, so there is no data-path collective: rank r converts the
contiguous slice ``shard_bounds(n, world, r)`` of the flat index range and keeps its results
resident on its own GPU.  This is the reference's thread split (hpgeom/hpgeom.c:272-279:
``chunk = n / T``, the first ``n % T`` workers get one extra element) lifted from pthreads to
devices.  ``gather`` is the optional collective (NCCL all_gather over NVLink when the process
group is NCCL; gloo in the CPU tests).
"""
import os


def shard_bounds(n, world_size, rank):
    """[start, end) of `rank`'s contiguous shard; same split rule as hpgeom/hpgeom.c:272-279."""
    n, world_size, rank = int(n), int(world_size), int(rank)
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size/rank")
    chunk, rem = divmod(n, world_size)
    start = rank * chunk + min(rank, rem)
    return start, start + chunk + (1 if rank < rem else 0)


def env_rank():
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard(array, world_size=None, rank=None):
    """This rank's contiguous slice of a flat (leading-dimension) array."""
    r, _, w = env_rank()
    world_size = w if world_size is None else world_size
    rank = r if rank is None else rank
    lo, hi = shard_bounds(array.shape[0], world_size, rank)
    return array[lo:hi]


def gather(local, group=None):
    """Optional: all-gather variable-length shards along dim 0 -> the full result on every rank.

    The output is allocated ONCE at its final size and the collective writes straight into it: equal shards
    are one ``all_gather_into_tensor`` (NCCL ring / NVLS over NVLink for CUDA tensors); ragged shards (the
    reference split gives the first ``n % G`` ranks one extra row) are G broadcasts, each into its own slice of
    the output -- the same bytes on the wire, still no padded temporaries and no concatenation.  Not part of the
    hot path: results normally stay resident per GPU (an 8-GPU gather of 8e9 int64 pixels moves 56 GB into each
    GPU, an order of magnitude more time than computing them)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    local = local.contiguous()
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes_t = torch.empty(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(sizes_t, n_local, group=group)
    sizes = [int(v) for v in sizes_t.tolist()]
    out = torch.empty((sum(sizes),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if len(set(sizes)) == 1:
        if sizes[0]:
            dist.all_gather_into_tensor(out, local, group=group)
        return out
    start = 0
    for r, m in enumerate(sizes):
        part = out[start:start + m]
        if m:
            if r == rank:
                part.copy_(local)
            dist.broadcast(part, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
        start += m
    return out
