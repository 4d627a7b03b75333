"""BASELINE config[0]: the reference's OWN node code -- PFTracker's constructor, callback, getMeasurementProposal,
get3Dpose, publishTFtree and publish2Dpos (src/pfPose.cpp), compiled in place against oracle/cvshim -- driven with
synthetic image / likelihood / face-ROI messages, against the oracle restatement of the same path.  The candidates the
reference draws internally (cv::randn / cv::randu) and its blurred likelihood image are read back from the shim, so
both sides see identical inputs and identical uniform draws."""
import numpy as np
import pytest

import mkf_oracle as orc
import mkf_ref
import mkfbodytracker_pdaf_b200 as mk

pytestmark = pytest.mark.skipif(not mkf_ref.available(), reason="oracle/_ref/libref.so not built")

ROI = (300, 51, 47, 47)  # x, y, w, h


def likelihood_image(frame, rng):
    """skin-likelihood image: two hand blobs, a face blob, speckle, lots of exact zeros"""
    yy, xx = np.mgrid[0:480, 0:640]
    img = np.zeros((480, 640))
    hx = 388 + 60 * np.sin(2 * np.pi * frame / 75)
    hy = 250 + 70 * np.sin(2 * np.pi * frame / 50 + np.pi / 3)
    for cx, cy, amp, sd in ((hx, hy, 255, 14), (hx - 140, hy, 230, 14), (323, 74, 200, 20)):
        img += amp * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * sd * sd))
    img += (rng.random(img.shape) < 0.02) * rng.integers(20, 90, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def test_reference_node_frames_match_oracle(left_arm, right_arm):
    rng = np.random.default_rng(2024)
    ticks0 = [int(v) for v in rng.integers(1, 2**62, 2)]
    tr = mkf_ref.RefTracker(mk.MODEL_DIR, "data13D_PCA_100000_15_12.yml", "data23D_PCA_100000_15_12.yml", *ticks0)
    N = tr.N
    assert N == 500  # src/pfPose.cpp:57
    fL = orc.Filter(left_arm.orc, N, alias_mode=orc.ALIAS_CV_SHALLOW_LITERAL)
    fR = orc.Filter(right_arm.orc, N, alias_mode=orc.ALIAS_CV_SHALLOW_LITERAL)
    fL.reset(u=-1.0, seed=ticks0[0])
    fR.reset(u=-1.0, seed=ticks0[1])
    for arm, f in ((0, fL), (1, fR)):  # constructor: loadGaussian (incl. quirk B4) + resample + resetTracker
        xr, Pr = tr.get_state(arm)
        xo, Po = f.get_state()
        assert np.array_equal(xr, xo) and np.array_equal(Pr, Po)
    roi = np.array(ROI, float)
    frames = 6
    for fr in range(frames):
        like = likelihood_image(fr, rng)
        ticks = [int(v) for v in rng.integers(1, 2**62, 6)]
        has_face = fr != 3  # one frame without a face: the reference publishes zeros and re-initialises (init = false)
        out = tr.callback(like, ROI if has_face else None, ticks if has_face else [])
        if not has_face:
            assert out["cands"] is None and out["n_tf"] == 0 and np.all(out["joints2d"] == 0)
            continue
        Cn = 10 * N  # src/pfPose.cpp:216
        cands = np.stack(out["cands"])  # (2 hands, 2, C)
        assert cands.shape == (2, 2, Cn)
        first = fr == 0 or fr == 4  # uniform box on the first frame and after the face was lost
        if first:
            assert cands[:, 0].min() >= max(300 - 4 * 47, 0) and cands[:, 0].max() < min(300 + 5 * 47, 640)
            assert cands[:, 1].min() >= 51 + 47 and cands[:, 1].max() < min(51 + 7 * 47, 480)
        blurred = out["blurred"]
        assert blurred.max() > 150 and (blurred == 0).mean() > 0.2
        x, y = cands[:, 0], cands[:, 1]
        inside = (y > 0) & (y < 480) & (x > 0) & (x < 640)
        Lv = np.where(inside, blurred[np.clip(y.astype(int), 0, 479), np.clip(x.astype(int), 0, 639)], 0).astype(np.uint8)
        u = [mkf_ref.tick_to_u(t, 0) for t in ticks]
        a = orc.associate(fL, fR, cands, Lv, roi, np.array(u[:2]), seed_cand=[ticks[0], ticks[1]])
        assert a["status"] == 0
        rl = fL.update(a["meas"][0], u[2], u[3], seed_ind=ticks[2], seed_post=ticks[3])
        rr = fR.update(a["meas"][1], u[4], u[5], seed_ind=ticks[4], seed_post=ticks[5])
        assert rl["status"] == 0 and rr["status"] == 0
        poses = []
        for arm, f in ((0, fL), (1, fR)):
            xr, Pr = tr.get_state(arm)
            xo, Po = f.get_state()
            assert np.array_equal(xr, xo), f"frame {fr} arm {arm}: filter state differs from the reference node"
            assert np.array_equal(Pr, Po)
            e_ref, p3_ref = tr.pose(arm)
            _, e_orc = f.estimate()
            assert np.array_equal(e_ref, e_orc)
            assert np.allclose(p3_ref, orc.get3dpose(e_orc), rtol=1e-13, atol=1e-13)
            poses.append(e_orc)
        tf, j2 = orc.skeleton(poses[0], poses[1])
        assert out["n_tf"] == 9
        assert np.allclose(out["tf"], tf, rtol=1e-12, atol=1e-12)
        assert np.array_equal(out["joints2d"], j2)
        # sanity: the tracked hands sit on the likelihood blobs
        assert abs(poses[0][0] - (388 + 60 * np.sin(2 * np.pi * fr / 75))) < 60


def test_config0_golden_trajectory_of_the_reference_node():
    """the committed 300-frame trajectory of the reference node is reproduced bit for bit (first 30 frames)"""
    import os
    import sys
    gold_path = os.path.join(os.path.dirname(__file__), "golden", "config0_reference_node.npz")
    if not os.path.exists(gold_path):
        pytest.skip("golden/config0_reference_node.npz missing")
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden_reference_node as mg
    gold = np.load(gold_path)
    got = mg.run(frames=30)
    assert np.array_equal(got["pose"], gold["pose"][:30])
    assert np.array_equal(got["tf"], gold["tf"][:30]) and np.array_equal(got["joints2d"], gold["joints2d"][:30])
    # the node tracks the synthetic hands: left hand follows the moving blob, right hand its mirror
    fr = np.arange(300)
    hx = 388 + 60 * np.sin(2 * np.pi * fr / 75)
    hy = 250 + 70 * np.sin(2 * np.pi * fr / 50 + np.pi / 3)
    err = np.hypot(gold["pose"][20:, 0, 0] - hx[20:], gold["pose"][20:, 0, 1] - hy[20:])
    assert np.median(err) < 25.0
