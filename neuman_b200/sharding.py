"""Ray sharding across the GPUs of one box: every ray is independent, so rank r
renders a contiguous row-major pixel range and one all_gather reassembles the frame.  No other
collective is on the path."""
import torch


def shard_range(n_pixels, rank, world):
    """Contiguous, balanced split: returns (first_pixel, count) of `rank`."""
    base, rem = divmod(int(n_pixels), int(world))
    cnt = base + (1 if rank < rem else 0)
    p0 = rank * base + min(rank, rem)
    return p0, cnt


def gather_frame(local, n_pixels, rank, world, group=None):
    """local: [count_r, C] tensor of this rank's pixels (CUDA for nccl, CPU for gloo).
    Returns the full [n_pixels, C] frame on every rank (one all_gather of equal-sized, padded shards)."""
    import torch.distributed as dist
    if world == 1:
        return local
    C = local.shape[1]
    per = (n_pixels + world - 1) // world
    pad = torch.zeros(per, C, dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty(world * per, C, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = []
    for r in range(world):
        _, cnt = shard_range(n_pixels, r, world)
        parts.append(out[r * per:r * per + cnt])
    return torch.cat(parts, 0)
