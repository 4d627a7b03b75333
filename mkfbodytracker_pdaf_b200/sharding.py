"""Multi-GPU plumbing of the hot path: tracks are independent, so they are
block-partitioned over the ranks of one node with NO data-path collective; the only
communication is one final all_gather of the per-track summaries {pose[D], wsum, status}.

torch.distributed is used for the plumbing only (NCCL over NVLink on the GPU box, gloo in the
CPU tests)."""
from __future__ import annotations


def shard_tracks(total_tracks: int, world: int, rank: int):
    """contiguous block partition: returns (first_track, n_tracks) of `rank`; the first
    `total_tracks % world` ranks take one extra track.  A person's two arm filters share a
    track id, so they always land on the same rank."""
    if world < 1 or not (0 <= rank < world) or total_tracks < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(total_tracks, world)
    n = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n


def pack_summary(pose, wsum, status):
    """(T, D) pose, (T,) wsum, (T,) status -> (T, D+2) float64 rows"""
    import torch
    T, D = pose.shape
    out = torch.empty((T, D + 2), dtype=torch.float64, device=pose.device)
    out[:, :D] = pose
    out[:, D] = wsum
    out[:, D + 1] = status.to(torch.float64)
    return out


def gather_summaries(local, world: int, counts=None):
    """all_gather of per-track summary rows; `counts` = rows per rank when shards are ragged.
    Returns the (sum(counts), D+2) tensor in global track order on every rank."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    if counts is None or len(set(counts)) == 1:
        out = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    mx = max(counts)
    pad = torch.zeros((mx, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * mx, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * mx: r * mx + counts[r]] for r in range(world)], 0)
