#!/usr/bin/env python
"""BASELINE config[0] golden trajectory FROM THE REFERENCE ITSELF: the reference's own node code
(src/pfPose.cpp, pf2DRao.cpp, my_gmm.cpp, KF_model.cpp compiled in place into oracle/_ref/libref.so against the
OpenCV/ROS stubs of oracle/cvshim) run on a deterministic synthetic 300-frame head/hands sequence with the
data13D/data23D_PCA_100000_15_12 models.  Needs /root/reference (authoring container).  Writes
tests/golden/config0_reference_node.npz: per-frame 22-D pose of both arms, TF translations, 2-D joints, and the
final particle means."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import mkf_ref  # noqa: E402
import mkfbodytracker_pdaf_b200 as mk  # noqa: E402
from test_ref_tracker import ROI, likelihood_image  # noqa: E402

SEED, FRAMES = 20261017, 300


def run(frames=FRAMES):
    rng = np.random.default_rng(SEED)
    ticks0 = [int(v) for v in rng.integers(1, 2**62, 2)]
    tr = mkf_ref.RefTracker(mk.MODEL_DIR, "data13D_PCA_100000_15_12.yml", "data23D_PCA_100000_15_12.yml", *ticks0)
    pose = np.zeros((frames, 2, 22))
    tf = np.zeros((frames, 10, 3))
    j2 = np.zeros((frames, 8, 2))
    for fr in range(frames):
        like = likelihood_image(fr, rng)
        ticks = [int(v) for v in rng.integers(1, 2**62, 6)]
        out = tr.callback(like, ROI, ticks)
        pose[fr, 0], _ = tr.pose(0)
        pose[fr, 1], _ = tr.pose(1)
        tf[fr], j2[fr] = out["tf"], out["joints2d"]
    x0, _ = tr.get_state(0)
    x1, _ = tr.get_state(1)
    return dict(pose=pose, tf=tf, joints2d=j2, x_final=np.stack([x0, x1]))


if __name__ == "__main__":
    assert mkf_ref.available(), "oracle/_ref/libref.so could not be built (needs /root/reference)"
    g = run()
    np.savez_compressed(os.path.join(HERE, "config0_reference_node.npz"), **g)
    print("written; final left hand estimate", g["pose"][-1, 0, :2], "right", g["pose"][-1, 1, :2])
