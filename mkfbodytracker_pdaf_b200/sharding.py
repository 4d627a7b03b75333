"""Multi-GPU plumbing of the hot path: tracks are independent, so they are
block-partitioned over the ranks of one node with NO data-path collective; the only
communication is one final all_gather of the per-track summaries {pose[D], wsum, status}.

Two routes to that gather:
  * `Comm` + `gather_summaries_native`: the C ABI's own collective (mkf_batch_gather_summaries: rows packed on the device
    and gathered in place with ncclAllGather on the batch's stream) -- what a C++ host uses; torch.distributed only
    carries the 128-byte NCCL id from rank 0 to the others;
  * `pack_summary` + `gather_summaries`: the same rows through torch.distributed (gloo in the CPU tests)."""
from __future__ import annotations

import ctypes as C


class Comm:
    """one NCCL communicator per rank, created through the C ABI (mkf_comm_create)"""

    def __init__(self, world: int, rank: int, device: int, id_bytes: bytes):
        from . import _lib as L
        self._L = L
        h = C.c_void_p()
        buf = C.create_string_buffer(id_bytes, L.COMM_ID_BYTES)
        L.check(L.lib.mkf_comm_create(C.byref(h), world, rank, C.cast(buf, C.c_void_p), device))
        self._h = h
        self.world, self.rank, self.device = world, rank, device

    @staticmethod
    def unique_id() -> bytes:
        from . import _lib as L
        buf = C.create_string_buffer(L.COMM_ID_BYTES)
        L.check(L.lib.mkf_comm_unique_id(C.cast(buf, C.c_void_p)))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device: int):
        """rank 0 draws the NCCL id; the initialised torch.distributed group (any backend) broadcasts its 128 bytes"""
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(world, rank, device, box[0])

    def nccl_version(self) -> int:
        v = C.c_int(0)
        self._L.check(self._L.lib.mkf_comm_info(self._h, None, None, C.byref(v)))
        return v.value

    def close(self):
        if self._h:
            self._L.lib.mkf_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gather_summaries_native(batch, comm: Comm, out, rows_per_rank: int = 0):
    """mkf_batch_gather_summaries: `out` (world * rows_per_rank, D + 2) float64, device tensor (asynchronous on the
    batch's stream) or host array (complete on return)"""
    from . import _lib as L
    from .tracker import _addr
    addr, mem = _addr(out)
    L.check(L.lib.mkf_batch_gather_summaries(batch._h, comm._h, int(rows_per_rank), addr, mem))
    return out


def shard_tracks_native(total_tracks: int, world: int, rank: int):
    """mkf_shard_tracks (the C ABI's copy of shard_tracks)"""
    from . import _lib as L
    first, n = C.c_int64(0), C.c_int64(0)
    L.check(L.lib.mkf_shard_tracks(total_tracks, world, rank, C.byref(first), C.byref(n)))
    return first.value, n.value


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
