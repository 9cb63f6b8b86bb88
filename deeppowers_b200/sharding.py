"""Batch sharding across ranks: ciphertexts are independent, so rank r of W owns the
contiguous slice [start, stop) of the batch and no collective is needed on the data path.  The only
collective is the optional final gather of results to one rank (NCCL on GPUs, gloo in CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(batch, rank, world):
    """contiguous, balanced split: the first (batch % world) ranks get one extra ciphertext"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_results(local, batch, dst=0, group=None):
    """Gathers the per-rank result slices (tensor [n_local, ...]) into [batch, ...] on rank `dst`.
    Slices may be ragged (batch not divisible by world): they are padded to the largest slice for the
    collective and trimmed afterwards.  Returns the full tensor on dst, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    if local.shape[0] != sizes[rank]:
        raise ValueError("local slice has %d items, expected %d" % (local.shape[0], sizes[rank]))
    biggest = max(sizes)
    padded = local
    if local.shape[0] < biggest:
        pad = torch.zeros((biggest - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], 0)
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded.contiguous(), bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: sizes[r]] for r in range(world)], 0)
