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
    """Optional: all-gather variable-length shards along dim 0 -> full result on every rank.

    Pads to the largest shard, runs ONE all_gather (NCCL over NVLink/NVSwitch for CUDA tensors),
    then trims.  Not part of the hot path: results normally stay resident per GPU."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    padded = local
    if local.shape[0] < m:
        pad = torch.zeros((m - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
