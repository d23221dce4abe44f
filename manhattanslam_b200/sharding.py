"""Multi-GPU plumbing of the front-end: independent RGB-D frames are sharded one
contiguous chunk per rank; the only collective is one all-gather of the per-frame count table
(keypoints, new surfels, updated surfels) so that every rank knows every frame's output sizes.
torch.distributed is plumbing only (NCCL on the GPU box, gloo in the CPU tests)."""


def shard_frames(n_frames, rank, world):
    """Contiguous chunk [lo, hi) of this rank; chunk sizes differ by at most one frame."""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_counts(local_counts, n_frames, rank, world):
    """local_counts: (chunk, C) int32 tensor of this rank's frames -> (n_frames, C) table on every rank."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_counts
    C = local_counts.shape[1]
    chunk = (n_frames + world - 1) // world
    pad = torch.zeros((chunk, C), dtype=local_counts.dtype, device=local_counts.device)
    pad[:local_counts.shape[0]] = local_counts
    out = torch.empty((world * chunk, C), dtype=local_counts.dtype, device=local_counts.device)
    dist.all_gather_into_tensor(out, pad)
    rows = []
    for r in range(world):
        lo, hi = shard_frames(n_frames, r, world)
        rows.append(out[r * chunk:r * chunk + (hi - lo)])
    return torch.cat(rows, 0)
