"""GPU (literal alias mode) against the reference's own node code (oracle/_ref: src/pfPose.cpp et al. on the
OpenCV/ROS shim), BASELINE config[0] style: per frame the reference draws its candidates and blurs its likelihood
image; the same candidates, likelihood samples, ROI and uniform draws go through mkf_batch_associate, and the
filter states, pose, 3-D joints, TF translations and 2-D joints must agree."""
import numpy as np
import pytest

import mkf_ref
import mkfbodytracker_pdaf_b200 as mk
from helpers import RTOL, rel_err
from mkfbodytracker_pdaf_b200 import _lib as L
from test_ref_tracker import ROI, likelihood_image

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not mkf_ref.available(), reason="oracle/_ref/libref.so not built")]


def literal(arm):
    a = arm.arrays
    p = mk.default_params()
    p.alias_mode = L.ALIAS_CV_SHALLOW_LITERAL
    return mk.Model.from_arrays(a["means"], a["covs"], a["weights"], a["gamma"], a["pca_proj"], a["pca_mean"], p)


def test_gpu_tracks_the_reference_node(left_arm, right_arm):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(7)
    ticks0 = [int(v) for v in rng.integers(1, 2**62, 2)]
    tr = mkf_ref.RefTracker(mk.MODEL_DIR, "data13D_PCA_100000_15_12.yml", "data23D_PCA_100000_15_12.yml", *ticks0)
    N = tr.N
    s = torch.cuda.Stream()
    b0 = mk.TrackBatch(literal(left_arm), 1, N, stream=s.cuda_stream)
    b1 = mk.TrackBatch(literal(right_arm), 1, N, stream=s.cuda_stream)
    b0.reset(np.array([mkf_ref.tick_to_u(ticks0[0], 0)]))
    b1.reset(np.array([mkf_ref.tick_to_u(ticks0[1], 0)]))
    roi = np.array([ROI], float)
    worst = dict(x=0.0, P=0.0, pose=0.0, p3=0.0, tf=0.0, j2=0.0)
    frames = 25
    for fr in range(frames):
        like = likelihood_image(fr, rng)
        ticks = [int(v) for v in rng.integers(1, 2**62, 6)]
        out = tr.callback(like, ROI, ticks)
        cands = np.stack(out["cands"])[None]  # (1, 2, 2, C)
        x, y = cands[0, :, 0], cands[0, :, 1]
        inside = (y > 0) & (y < 480) & (x > 0) & (x < 640)
        blurred = out["blurred"]
        Lv = np.where(inside, blurred[np.clip(y.astype(int), 0, 479), np.clip(x.astype(int), 0, 639)], 0).astype(np.uint8)[None]
        u = np.array([mkf_ref.tick_to_u(t, 0) for t in ticks])
        seeds = np.array([[[ticks[0], ticks[2], ticks[3]], [ticks[1], ticks[4], ticks[5]]]], np.uint64)
        mk.associate(b0, b1, cands, Lv, roi, u[None, 0:2], u[None, [2, 4]], u[None, [3, 5]], seeds=seeds)
        poses = []
        for arm, b in ((0, b0), (1, b1)):
            d = b.download()
            assert not d["status"].any()
            xr, Pr = tr.get_state(arm)
            worst["x"] = max(worst["x"], rel_err(d["x"][0], xr))
            worst["P"] = max(worst["P"], rel_err(d["P"][0], Pr))
            e_ref, p3_ref = tr.pose(arm)
            _, pose = b.estimate()
            worst["pose"] = max(worst["pose"], rel_err(pose[0], e_ref))
            worst["p3"] = max(worst["p3"], rel_err(b.pose3d()[0], p3_ref))
            poses.append(pose[0])
        tf, j2 = mk.skeleton(b0, b1)
        worst["tf"] = max(worst["tf"], float(np.abs(tf[0] - out["tf"]).max() / np.abs(out["tf"]).max()))
        worst["j2"] = max(worst["j2"], float(np.abs(j2[0] - out["joints2d"]).max() / np.abs(out["joints2d"]).max()))
        assert max(worst.values()) <= RTOL, (fr, worst)
    print("GPU vs reference node, worst relative errors over", frames, "frames:", worst)
    assert max(worst.values()) < 1e-8
