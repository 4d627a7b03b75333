#!/usr/bin/env python
"""Experiment: BASELINE config 2 (4096 tracks x 500 slots, shared column) as S independent sub-batches on S streams, so that
one sub-batch's bookkeeping kernels (k_frame_heads, k_runs_repair, k_resample_runs: latency chains) run beside another's
slot kernel.  Prints ms per frame of the whole 4096-track job for S = 1, 2, 3, 4."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mkfbodytracker_pdaf_b200 as mk

SEED = 0x5EED0002
T, N, F, W = 4096, 500, 200, 5
dev = torch.device("cuda", 0)
model = mk.Model.load(mk.LEFT_ARM_MODEL, mk.RIGHT_ARM_MODEL)
for S in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]:
    Ts = [T // S + (1 if i < T % S else 0) for i in range(S)]
    streams = [torch.cuda.Stream() for _ in range(S)]
    subs = []
    t0_ = 0
    for i in range(S):
        with torch.cuda.stream(streams[i]):
            b = mk.TrackBatch(model, Ts[i], N, device=0, stream=streams[i].cuda_stream)
            meas = torch.empty((F + W, Ts[i], 6), dtype=torch.float64, device=dev)
            ui = torch.empty((F + W, Ts[i]), dtype=torch.float64, device=dev)
            up = torch.empty((F + W, Ts[i]), dtype=torch.float64, device=dev)
            for f in range(F + W):
                b.synth_fill(SEED, t0_, f, 1, mk.MEAS_SHARED, meas[f], ui[f], up[f])
            u0 = torch.empty(Ts[i], dtype=torch.float64, device=dev)
            b.synth_fill(SEED, t0_, 0xFFFFFF, 1, mk.MEAS_SHARED, meas[0].clone(), u0, None)
            pose = torch.empty((Ts[i], model.D), dtype=torch.float64, device=dev)
            b.reset(u0)
            subs.append((b, meas, ui, up, pose))
        t0_ += Ts[i]
    torch.cuda.synchronize()

    def frame(f):
        for b, meas, ui, up, pose in subs:
            b.update(meas[f], ui[f], up[f])
            b.estimate_into(None, pose)

    for f in range(W):
        frame(f)
    torch.cuda.synchronize()
    reps = []
    for _ in range(3):
        t0 = time.perf_counter()
        for f in range(W, W + F):
            frame(f)
        t_issue = time.perf_counter()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        reps.append(((t1 - t0) / F * 1e3, (t_issue - t0) / F * 1e3))
    ms, issue = sorted(reps)[1]
    chk = float(sum(p[4][:, :2].sum() for p in subs))
    print(json.dumps({"sub_batches": S, "ms_per_frame": ms, "host_issue_ms_per_frame": issue,
                      "frame_updates_per_s": T / ms * 1e3, "pose_check": chk}), flush=True)
    for p in subs:
        p[0].close()
