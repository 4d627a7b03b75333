#!/usr/bin/env python
"""Free-running soak of the association path (BASELINE config 3 inputs): persons x frames of
mkf_batch_associate(do_update) -- candidate gather + record sharing keyed on (parent record, component, candidate) --
against the CPU oracle on identical candidates and draws.  Counts gate / bin / parent mismatches and the worst
relative errors.  Output: one JSON line (committed under profiles/).

    python tools/soak_assoc.py PERSONS SLOTS CANDIDATES FRAMES"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import mkf_oracle as orc
import mkfbodytracker_pdaf_b200 as mk
from helpers import rel_err, rel_err_weights, synth_u_init

T, N, Cn, frames, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), 0x5EED0003
ml = mk.Model.load(mk.LEFT_ARM_MODEL, mk.RIGHT_ARM_MODEL)
mr = mk.Model.load(mk.RIGHT_ARM_MODEL, mk.RIGHT_ARM_MODEL)


def omodel(m):
    a = m.arrays()
    return orc.Model(a["means"], a["covs"], a["weights"], a["gamma"], a["pca_proj"], a["pca_mean"])


ol, orr = omodel(ml), omodel(mr)
tracks = list(range(T))
u0 = synth_u_init(seed, tracks)
fl = [orc.Filter(ol, N) for _ in tracks]
fr_ = [orc.Filter(orr, N) for _ in tracks]
for t in tracks:
    fl[t].reset(u=u0[t])
    fr_[t].reset(u=u0[t])
s = torch.cuda.Stream()
b0 = mk.TrackBatch(ml, T, N, stream=s.cuda_stream)
b1 = mk.TrackBatch(mr, T, N, stream=s.cuda_stream)
b0.reset(u0)
b1.reset(u0)
roi = np.tile(np.array([300.0, 51.0, 47.0, 47.0]), (T, 1))
rng = np.random.default_rng(5)
gate_m = bin_m = par_m = 0
worst = dict(aw=0.0, w=0.0, x=0.0, P=0.0)
shared = []
t0 = time.time()
for frame in range(frames):
    cand = np.zeros((T, 2, 2, Cn))
    Lv = np.zeros((T, 2, Cn), np.uint8)
    for t in tracks:
        for h in range(2):
            for c in range(Cn):
                cand[t, h, 0, c], cand[t, h, 1, c], Lv[t, h, c] = orc.synth_candidate(seed, t, frame, h, Cn, c)
    u_cand, u_ind, u_post = rng.random((T, 2)), rng.random((T, 2)), rng.random((T, 2))
    mk.associate(b0, b1, cand, Lv, roi, u_cand, u_ind, u_post)
    res = mk.assoc_results(b0, Cn)
    full = frame % 10 == 0 or frame == frames - 1
    d0, d1 = b0.download(state=full, cov=full), b1.download(state=full, cov=full)
    rec, ns = b0.shared_records()
    shared.append(rec / max(ns, 1))
    for t in tracks:
        want = orc.associate(fl[t], fr_[t], cand[t], Lv[t], roi[t], u_cand[t])
        gate_m += int((res["gate"][t] != want["gate"]).sum())
        bin_m += int((res["bins"][t] != want["bins"]).sum())
        worst["aw"] = max(worst["aw"], rel_err_weights(res["weights"][t], want["weights"]))
        for arm, (f, d) in enumerate(((fl[t], d0), (fr_[t], d1))):
            r = f.update(want["meas"][arm], u_ind[t, arm], u_post[t, arm])
            par_m += int((d["parents"][t] != r["parents"]).sum())
            worst["w"] = max(worst["w"], rel_err_weights(d["w_norm"][t], r["w_norm"]))
            if full:
                xo, Po = f.get_state()
                worst["x"] = max(worst["x"], rel_err(d["x"][t], xo))
                worst["P"] = max(worst["P"], rel_err(d["P"][t], Po))
print(json.dumps(dict(persons=T, slots=N, candidates_per_hand=Cn, frames=frames, gate_decisions=T * 2 * Cn * frames,
                      gate_mismatches=gate_m, candidate_bins=T * 2 * N * frames, bin_mismatches=bin_m,
                      resampled_indices=T * 2 * N * frames, index_mismatches=par_m, worst_rel_err=worst,
                      distinct_records_fraction_last=round(shared[-1], 4), seconds=round(time.time() - t0, 1))))
