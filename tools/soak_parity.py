#!/usr/bin/env python
"""Free-running soak: many tracks x many frames, GPU vs CPU oracle on identical inputs and draws.
Counts resampled-index / indicator mismatches, the worst relative errors and how often the exact
sequential fallback was needed.  Output: one JSON line (committed under profiles/)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np

import mkf_oracle as orc
import mkfbodytracker_pdaf_b200 as mk
from helpers import rel_err, rel_err_weights, synth_frame, synth_u_init

T, N, frames, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 0x5EED0002
alias = int(sys.argv[4]) if len(sys.argv) > 4 else 0
per_slot = len(sys.argv) > 5 and sys.argv[5] == "slot"
every = int(sys.argv[6]) if len(sys.argv) > 6 else 1  # read back every `every`-th frame only (the run-length pipeline
                                                      # then runs the frames in between without any per-slot array)
m0 = mk.Model.load(mk.LEFT_ARM_MODEL, mk.RIGHT_ARM_MODEL)
a = m0.arrays()
args = (a["means"], a["covs"], a["weights"], a["gamma"], a["pca_proj"], a["pca_mean"])
prm = mk.default_params()
prm.alias_mode = alias
m = mk.Model.from_arrays(*args, prm)
om = orc.Model(*args)
tracks = list(range(T))
u0 = synth_u_init(seed, tracks)
fs = [orc.Filter(om, N, alias_mode=alias) for _ in tracks]
for f, u in zip(fs, u0):
    f.reset(u=u)
b = mk.TrackBatch(m, T, N)
b.reset(u0)
mism = ind_mism = flagged = 0
worst = dict(w=0.0, x=0.0, P=0.0)
t0 = time.time()
for fr in range(frames):
    meas, ui, up = synth_frame(seed, tracks, fr, N if per_slot else None)
    res = [fs[t].update(meas[t], ui[t], up[t]) for t in tracks]
    b.update(meas, ui, up)
    if fr % every and fr != frames - 1:
        continue
    full = fr % 20 == 0 or fr == frames - 1
    d = b.download(state=full, cov=full)
    flagged += int(((d["status"] & 0x3) != 0).sum())
    for t in tracks:
        mism += int((d["parents"][t] != res[t]["parents"]).sum())
        ind_mism += int((d["indicators"][t] != res[t]["indicators"]).sum())
        worst["w"] = max(worst["w"], rel_err_weights(d["w_norm"][t], res[t]["w_norm"]))
        if full:
            xo, Po = fs[t].get_state()
            worst["x"] = max(worst["x"], rel_err(d["x"][t], xo))
            worst["P"] = max(worst["P"], rel_err(d["P"][t], Po))
compared = len([fr for fr in range(frames) if fr % every == 0 or fr == frames - 1])
print(json.dumps(dict(tracks=T, slots=N, frames=frames, alias_mode=alias, per_slot_columns=per_slot,
                      read_back_every=every, resamples=T * frames, resampled_indices_compared=T * compared * N,
                      index_mismatches=mism,
                      indicator_mismatches=ind_mism, fallback_flagged_track_frames=flagged, worst_rel_err=worst,
                      seconds=round(time.time() - t0, 1))))
