"""Multi-GPU plumbing of the hot path: proofs of a batch are independent, so rank r proves the index range
shard_range(r, world, total) with no exchange during proving; the only collective is the gather of the
fixed-size proof records at the end (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(rank, world, total):
    """contiguous, balanced partition of [0, total): the first total % world ranks get one extra proof"""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_records(local, total, group=None):
    """local: uint8 tensor [count_r, record_len] of this rank's proofs (count_r from shard_range).
    Returns the [total, record_len] tensor of all proofs in global index order on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    base, extra = divmod(total, world)
    width = base + (1 if extra else 0)
    padded = torch.zeros((width, local.shape[1]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * width, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = []
    for r in range(world):
        a, b = shard_range(r, world, total)
        parts.append(out[r * width: r * width + (b - a)])
    del rank
    return torch.cat(parts, dim=0)
