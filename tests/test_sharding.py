"""the N>1 path (track sharding + final gather of per-track summaries) on CPU: world_size 2, gloo"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mkfbodytracker_pdaf_b200.sharding import gather_summaries, pack_summary, shard_tracks


def test_shard_tracks_partitions_exactly():
    for total in (0, 1, 7, 4096, 1048576, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_tracks(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == total
            for (f0, n0), (f1, _) in zip(spans, spans[1:]):
                assert f1 == f0 + n0
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1
    with pytest.raises(ValueError):
        shard_tracks(10, 2, 2)


def test_c_abi_shard_tracks_matches_python():
    """mkf_shard_tracks (what a C++ host calls) is the same partition"""
    from mkfbodytracker_pdaf_b200.sharding import shard_tracks_native
    import mkfbodytracker_pdaf_b200 as mk
    for total in (0, 1, 7, 4096, 1048576, 1000003):
        for world in (1, 2, 3, 8):
            for r in range(world):
                assert shard_tracks_native(total, world, r) == shard_tracks(total, world, r)
    with pytest.raises(mk.MkfError):
        shard_tracks_native(10, 2, 2)


@pytest.mark.gpu
def test_native_gather_single_rank(left_arm):
    """mkf_comm_create / mkf_batch_gather_summaries with one rank (the -m gpu box has one GPU; bench.py --gpus N runs
    the same call over N ranks and cross-checks rows between ranks): rows = {pose, wsum, status}, padding rows zero,
    device and host destinations agree"""
    import mkfbodytracker_pdaf_b200 as mk
    from helpers import synth_frame, synth_u_init
    from mkfbodytracker_pdaf_b200.sharding import Comm, gather_summaries_native
    T, N, seed = 37, 120, 0x5EED0005
    b = mk.TrackBatch(left_arm.mk, T, N)
    b.reset(synth_u_init(seed, range(T)))
    for fr in range(2):
        b.update(*synth_frame(seed, list(range(T)), fr))
    comm = Comm(1, 0, 0, Comm.unique_id())
    assert comm.nccl_version() >= 20000
    D = left_arm.mk.D
    host = np.full((T + 3, D + 2), -1.0)
    gather_summaries_native(b, comm, host, rows_per_rank=T + 3)
    _, pose = b.estimate()
    d = b.download(state=False, cov=False)
    assert np.array_equal(host[:T, :D], pose) and np.array_equal(host[:T, D], d["wsum"])
    assert np.array_equal(host[:T, D + 1], d["status"].astype(np.float64)) and not host[T:].any()
    dev = torch.full((T, D + 2), -1.0, dtype=torch.float64, device="cuda:0")
    gather_summaries_native(b, comm, dev)
    b.sync()
    assert np.array_equal(dev.cpu().numpy(), host[:T])
    comm.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    import mkf_oracle as orc
    import mkfbodytracker_pdaf_b200 as mk
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = mk.Model.load(mk.LEFT_ARM_MODEL, mk.RIGHT_ARM_MODEL)
    a = m.arrays()
    om = orc.Model(a["means"], a["covs"], a["weights"], a["gamma"], a["pca_proj"], a["pca_mean"])
    first, n = shard_tracks(total, world, rank)
    # the sharded job: each rank filters its own block of tracks with the CPU oracle (host-side
    # stand-in for the per-GPU batch) and contributes {pose, wsum, status} rows
    secs, _, pose = orc.bench_tracks(om, 0, 15, 1, want_pose=True)  # exercise the empty shard path
    pose = np.zeros((n, m.D))
    wsum = np.zeros(n)
    for i in range(n):
        f = orc.Filter(om, 15)
        f.reset(u=orc.synth_u(5, first + i, 0xFFFFFFFFFFFF, 0x1003))
        r = f.update(orc.synth_meas(5, first + i, 0, -1, 1), orc.synth_u(5, first + i, 0, 0x1001),
                     orc.synth_u(5, first + i, 0, 0x1002))
        _, pose[i] = f.estimate()
        wsum[i] = r["wsum"]
    local = pack_summary(torch.from_numpy(pose), torch.from_numpy(wsum), torch.zeros(n, dtype=torch.int32))
    counts = [shard_tracks(total, world, r)[1] for r in range(world)]
    out = gather_summaries(local, world, counts)
    if rank == 0:
        q.put(out.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_matches_single_process():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import mkf_oracle as orc
    import mkfbodytracker_pdaf_b200 as mk
    total, world = 9, 2  # ragged: 5 + 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = mk.Model.load(mk.LEFT_ARM_MODEL, mk.RIGHT_ARM_MODEL)
    a = m.arrays()
    om = orc.Model(a["means"], a["covs"], a["weights"], a["gamma"], a["pca_proj"], a["pca_mean"])
    assert got.shape == (total, m.D + 2)
    for t in range(total):
        f = orc.Filter(om, 15)
        f.reset(u=orc.synth_u(5, t, 0xFFFFFFFFFFFF, 0x1003))
        r = f.update(orc.synth_meas(5, t, 0, -1, 1), orc.synth_u(5, t, 0, 0x1001), orc.synth_u(5, t, 0, 0x1002))
        _, pose = f.estimate()
        assert np.array_equal(got[t, : m.D], pose) and got[t, m.D] == r["wsum"] and got[t, m.D + 1] == 0
