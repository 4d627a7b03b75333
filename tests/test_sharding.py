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
