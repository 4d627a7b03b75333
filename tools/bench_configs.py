#!/usr/bin/env python
"""Secondary measurements (bank mode, association-only variants, the reference operating point ...) beside the legs
bench.py itself carries for BASELINE.json configs 2-5 and the legacy pf2D filter, on one B200 (device-resident inputs, CUDA events on the launching stream).  Writes one JSON object per
line; results are summarised in BASELINE.md section 4.  Not the headline bench (that is bench.py)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import mkfbodytracker_pdaf_b200 as mk

PEAK = 6536.7
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)


def timed(fn, steps, warmup=3):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def rbpf(name, model, T, N, steps, per_slot, seed):
    b = mk.TrackBatch(model, T, N, 0, stream.cuda_stream)
    lay = mk.MEAS_PER_SLOT if per_slot else mk.MEAS_SHARED
    meas = torch.empty((T, 6, N) if per_slot else (T, 6), dtype=torch.float64, device=dev)
    ui = torch.empty(T, dtype=torch.float64, device=dev)
    up = torch.empty(T, dtype=torch.float64, device=dev)
    pose = torch.empty((T, model.D), dtype=torch.float64, device=dev)
    b.synth_fill(seed, 0, 0xFFFFFF, 1, mk.MEAS_SHARED, torch.empty((T, 6), dtype=torch.float64, device=dev), ui, None)
    b.reset(ui.clone())

    def step(f):
        b.synth_fill(seed, 0, f, 1, lay, meas, ui, up)  # per-slot inputs are too large to pre-generate
        b.update(meas, ui, up)
        b.estimate_into(None, pose)

    b.profile(0)
    ms = timed(step, steps)
    b.profile(steps)
    flagged = degenerate = 0
    for f in range(steps):
        step(100 + f)
        st = b.status()  # per-frame status (untimed loop): how often the closed-form resampler hands a track over
        flagged += int(((st & 0x3) != 0).sum())
        degenerate += int(((st & 0x4) != 0).sum())
    pr = b.profile_read()
    slot_ms = pr["ms_slot_update"] / pr["n"]
    out = dict(config=name, tracks=T, slots=N, ms_per_frame=ms, frame_updates_per_s=T / ms * 1e3,
               slot_updates_per_s=T * N / ms * 1e3, slot_kernel_ms=slot_ms,
               slot_kernel_gbs_algorithmic=T * N * 1500 / slot_ms / 1e6,
               slot_kernel_frac_of_measured_hbm=T * N * 1500 / slot_ms / 1e6 / PEAK,
               resample_ms=pr["ms_resample"] / pr["n"], bounds_ms=pr["ms_bounds"] / pr["n"],
               flagged_fallback_track_frames=flagged, degenerate_track_frames=degenerate, track_frames_checked=T * steps,
               includes="synthetic input generation kernel + update + estimate")
    print(json.dumps(out), flush=True)
    b.close()


def assoc(name, left, right, T, N, C, steps, seed):
    b0 = mk.TrackBatch(left, T, N, 0, stream.cuda_stream)
    b1 = mk.TrackBatch(right, T, N, 0, stream.cuda_stream)
    rng = np.random.default_rng(1)
    u0 = torch.tensor(rng.random(T), device=dev)
    b0.reset(u0)
    b1.reset(u0)
    cand = torch.empty((T, 2, 2, C), dtype=torch.float64, device=dev)
    cand[:, :, 0] = torch.rand((T, 2, C), dtype=torch.float64, device=dev) * 704 - 32
    cand[:, :, 1] = torch.rand((T, 2, C), dtype=torch.float64, device=dev) * 528 - 24
    cand[:, 0, 0, 0], cand[:, 0, 1, 0] = 388.0, 250.0
    cand[:, 1, 0, 0], cand[:, 1, 1, 0] = 248.0, 250.0
    L = torch.randint(0, 129, (T, 2, C), dtype=torch.uint8, device=dev)
    L[:, :, 0] = 220
    roi = torch.tensor([300.0, 51.0, 47.0, 47.0], dtype=torch.float64, device=dev).repeat(T, 1).contiguous()
    us = [torch.rand((T, 2), dtype=torch.float64, device=dev) for _ in range(3)]

    def only_assoc(f):
        mk.associate(b0, b1, cand, L, roi, us[0], None, None, do_update=False)

    def full(f):
        mk.associate(b0, b1, cand, L, roi, us[0], us[1], us[2], do_update=True)

    pose0 = torch.empty((T, left.D), dtype=torch.float64, device=dev)
    pose1 = torch.empty((T, right.D), dtype=torch.float64, device=dev)

    def tracker_frame(f):  # what a tracker loop runs per frame: associate + update both arms, then the output estimate
        mk.associate(b0, b1, cand, L, roi, us[0], us[1], us[2], do_update=True)
        b0.estimate_into(None, pose0)
        b1.estimate_into(None, pose1)

    ms_a = timed(only_assoc, steps)
    ms_f = timed(full, steps)
    ms_t = timed(tracker_frame, steps)
    st = b0.status() | b1.status()
    print(json.dumps(dict(config=name, persons=T, slots=N, candidates_per_hand=C, assoc_only_ms=ms_a,
                          assoc_only_person_frames_per_s=T / ms_a * 1e3,
                          candidate_weights_per_s=T * 2 * C / ms_a * 1e3, assoc_plus_update_ms=ms_f,
                          assoc_update_estimate_ms=ms_t,
                          frame_updates_per_s=2 * T / ms_f * 1e3, slot_updates_per_s=2 * T * N / ms_f * 1e3,
                          degenerate=int(((st & 0x24) != 0).sum()))), flush=True)
    b0.close()
    b1.close()


def pf2d(name, T, N, d, K, steps):
    rng = np.random.default_rng(2)
    means = rng.uniform(100, 400, (K, d))
    covs = []
    for _ in range(K):
        a = rng.standard_normal((d, d))
        covs.append(40 * (a @ a.T + d * np.eye(d)))
    covs = np.stack(covs)
    wts = rng.dirichlet(np.ones(K))
    p = mk.Pf2dBatch(T, N, means, covs, wts, 0, stream.cuda_stream)
    parts = torch.tensor(means[rng.integers(0, K, T * N)].reshape(T, N, d), device=dev)
    parts += torch.randn((T, N, d), dtype=torch.float64, device=dev) * 6
    p.set_particles(parts)
    meas = torch.tensor(np.tile(np.array([[250.0, 250.0], [250.0, 250.0]]), (T, 1, 1)), device=dev)
    meas[:, 0, 0] = parts[:, :, 6].mean(1)
    meas[:, 0, 1] = parts[:, :, 7].mean(1)
    meas[:, 1, 0] = parts[:, :, 0].mean(1)
    meas[:, 1, 1] = parts[:, :, 1].mean(1)
    u = torch.rand(T, dtype=torch.float64, device=dev)
    noise = torch.randn((T, N, d), dtype=torch.float64, device=dev)

    def step(f):
        p.update(meas, u, noise)

    ms = timed(step, steps)
    print(json.dumps(dict(config=name, filters=T, particles=N, dims=d, components=K, ms_per_update=ms,
                          particle_likelihoods_per_s=T * N / ms * 1e3, gbs_algorithmic_200B=T * N * 200 / ms / 1e6)),
          flush=True)
    p.close()


left = mk.Model.load(mk.LEFT_ARM_MODEL, mk.RIGHT_ARM_MODEL)
right = mk.Model.load(mk.RIGHT_ARM_MODEL, mk.RIGHT_ARM_MODEL)
which = sys.argv[1:] or ["2", "2lit", "2b", "3", "4", "5", "pf2d"]
if "2" in which:
    rbpf("config 2: 4096 tracks x 500 slots, shared column, alias INDEPENDENT", left, 4096, 500, 50, False, 0x5EED0002)
if "2lit" in which:
    prm = mk.default_params()
    prm.alias_mode = 1
    a = left.arrays()
    lit = mk.Model.from_arrays(a["means"], a["covs"], a["weights"], a["gamma"], a["pca_proj"], a["pca_mean"], prm)
    rbpf("config 2: 4096 tracks x 500 slots, shared column, alias CV_SHALLOW_LITERAL (quirk B3)", lit, 4096, 500, 50,
         False, 0x5EED0002)
    rbpf("config 4: 256 tracks x 65536 slots, per-slot columns, alias CV_SHALLOW_LITERAL", lit, 256, 65536, 10, True,
         0x5EED0004)
if "2b" in which:
    rbpf("config 2 (bank): 4096 tracks x 15 slots, shared column", left, 4096, 15, 50, False, 0x5EED0002)
if "3" in which:
    assoc("config 3: 16384 persons, 17 candidates/hand, N=500", left, right, 16384, 500, 17, 10, 0x5EED0003)
    assoc("config 3 (assoc-heavy): 16384 persons, 17 candidates/hand, N=15", left, right, 16384, 15, 17, 20, 0x5EED0003)
    assoc("reference operating point: 5000 candidates/hand, N=500, 256 persons", left, right, 256, 500, 5000, 10, 3)
if "4" in which:
    rbpf("config 4: 256 tracks x 65536 slots, per-slot columns", left, 256, 65536, 10, True, 0x5EED0004)
if "5" in which:
    rbpf("config 5 (1 GPU): 1048576 tracks x 15 slots, data23D, shared column", right, 1048576, 15, 20, False,
         0x5EED0005)
if "pf2d" in which:
    pf2d("legacy pf2D: 256 filters x 65536 particles, d=8, K=15", 256, 65536, 8, 15, 10)
    pf2d("legacy pf2D: 4096 filters x 500 particles, d=8, K=15", 4096, 500, 8, 15, 50)
