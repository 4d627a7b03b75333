#!/usr/bin/env python
"""Per-frame latency of the reference's node (PFTracker::callback, src/pfPose.cpp, N = 500 slots per arm, 5000
candidates per hand -- BASELINE config[0]'s operating point) in its two builds under oracle/_ref:

    python tools/bench_node.py ref      # pfPose.cpp on the reference's own KF_model / my_gmm / pf2DRao   (CPU)
    python tools/bench_node.py dropin   # the same pfPose.cpp on mkf_shims.hpp + libmkf_b200.so            (GPU)

One JSON line each.  Both run the node's host code (15x15 blur, candidate loops, cv::Mat glue) on one CPU thread, as the
reference does; only the ParticleFilter / my_gmm calls differ.  Test infrastructure: needs oracle/_ref.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import mkf_ref  # noqa: E402
import mkfbodytracker_pdaf_b200 as mk  # noqa: E402
from test_ref_tracker import ROI, likelihood_image  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "dropin"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rng = np.random.default_rng(1)
ticks0 = [int(v) for v in rng.integers(1, 2**62, 2)]
cls = mkf_ref.DropinTracker if which == "dropin" else mkf_ref.RefTracker
tr = cls(mk.MODEL_DIR, "data13D_PCA_100000_15_12.yml", "data23D_PCA_100000_15_12.yml", *ticks0)
L = tr._L
imgs = [likelihood_image(f, rng) for f in range(8)]
r = np.array([ROI[0], ROI[1], ROI[3], ROI[2]], np.uint32)
times = []
for f in range(frames + 5):
    like = imgs[f % 8]
    t = np.ascontiguousarray(rng.integers(1, 2**62, 6), np.int64)
    t0 = time.perf_counter()
    L.ref_tracker_callback(tr.h, like.ctypes.data_as(mkf_ref._u8p), 480, 640, 1, r.ctypes.data_as(mkf_ref._u32p),
                           t.ctypes.data_as(mkf_ref._i64p), 6)
    times.append(time.perf_counter() - t0)
times = np.array(times[5:]) * 1e3
e, _ = tr.pose(0)
print(json.dumps(dict(build=which, frames=frames, slots_per_arm=tr.N, candidates_per_hand=10 * tr.N,
                      ms_per_frame_median=float(np.median(times)), ms_per_frame_p90=float(np.percentile(times, 90)),
                      frames_per_s=float(1e3 / np.median(times)), left_hand_estimate=[float(e[0]), float(e[1])])))
