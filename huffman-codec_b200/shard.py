"""Multi-GPU plumbing of the batched codec.

Files are independent (the FGK tree is built per file, reference src/transform.cpp:366,388), so a
batch shards over ranks with NO data-path collective: every rank runs the full pipeline on its own
contiguous slice of the batch.  The only exchange is one all-gather of the per-file output sizes
(a few KB), from which every rank derives the same global offsets table -- the index of the
concatenated output container.  Works with any torch.distributed backend (NCCL on GPUs, gloo on
CPU for the tests)."""
import torch
import torch.distributed as dist


def shard_range(nfiles, rank, world):
    """Contiguous, balanced slice [lo, hi) of the batch owned by `rank`."""
    base, extra = divmod(nfiles, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_sizes(local_sizes, nfiles, group=None):
    """All-gather the per-file output sizes of every rank's shard.

    local_sizes: 1-D int64 tensor (this rank's shard, on the backend's device).
    Returns a 1-D int64 tensor of nfiles entries in global file order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_sizes.clone()
    per = -(-nfiles // world)                      # shards differ by at most one file: pad to equal
    padded = torch.zeros(per, dtype=torch.int64, device=local_sizes.device)
    padded[: local_sizes.numel()] = local_sizes
    out = torch.empty(per * world, dtype=torch.int64, device=local_sizes.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(nfiles, r, world)
        parts.append(out[r * per: r * per + (hi - lo)])
    return torch.cat(parts)


def global_offsets(sizes, align=16):
    """Exclusive scan of the aligned sizes -> (offsets, total) of the concatenated container."""
    al = (sizes + (align - 1)) // align * align
    off = torch.cumsum(al, 0) - al
    return off, int(al.sum().item())
