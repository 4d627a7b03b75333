"""INTEGRATION.md section 2 carried out for real: the reference's UNMODIFIED node source (src/pfPose.cpp: PFTracker's
constructor, callback, getMeasurementProposal, get3Dpose, publishTFtree, publish2Dpos) compiled against
include/mkf_shims.hpp instead of the reference's pf2DRao.h / my_gmm.h / KF_model.h and linked to libmkf_b200.so
(oracle/Makefile target `dropin`, built where /root/reference exists; the GPU box uses the prebuilt file).  Every
ParticleFilter / my_gmm call the node makes therefore runs on the GPU.  The node is fed synthetic frames and compared,
frame by frame, with the CPU oracle given the candidates the node drew and the same clock ticks -- the oracle itself
being pinned bit for bit to the reference node on its own classes (tests/test_ref_tracker.py)."""
import numpy as np
import pytest

import mkf_oracle as orc
import mkf_ref
import mkfbodytracker_pdaf_b200 as mk
from helpers import RTOL, rel_err
from test_ref_tracker import ROI, likelihood_image

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not mkf_ref.dropin_available(), reason="oracle/_ref/libref_dropin.so not built")]


@pytest.mark.parametrize("alias", [orc.ALIAS_CV_SHALLOW_LITERAL, orc.ALIAS_INDEPENDENT])
def test_reference_node_source_on_the_shims_matches_oracle(left_arm, right_arm, alias):
    rng = np.random.default_rng(99)
    ticks0 = [int(v) for v in rng.integers(1, 2**62, 2)]
    tr = mkf_ref.DropinTracker(mk.MODEL_DIR, "data13D_PCA_100000_15_12.yml", "data23D_PCA_100000_15_12.yml", *ticks0,
                               alias_mode=alias)
    N = tr.N
    assert N == 500  # src/pfPose.cpp:57
    fL = orc.Filter(left_arm.orc, N, alias_mode=alias)
    fR = orc.Filter(right_arm.orc, N, alias_mode=alias)
    fL.reset(u=-1.0, seed=ticks0[0])
    fR.reset(u=-1.0, seed=ticks0[1])
    for arm, f in ((0, fL), (1, fR)):  # constructor: loadGaussian (incl. quirk B4) + resample + resetTracker
        xr, Pr = tr.get_state(arm)
        xo, Po = f.get_state()
        # same component per slot (same clock-seeded draw); values pass through the device's measurement-aligned
        # basis and back, hence 1e-12 rather than bit equality
        assert rel_err(xr, xo) < 1e-12 and rel_err(Pr, Po) < 1e-12
    roi = np.array(ROI, float)
    worst = dict(x=0.0, P=0.0, pose=0.0, p3=0.0, tf=0.0, j2=0.0)
    frames = 12
    for fr in range(frames):
        like = likelihood_image(fr, rng)
        ticks = [int(v) for v in rng.integers(1, 2**62, 6)]
        has_face = fr != 5  # one frame without a face: the node publishes zeros and re-initialises
        out = tr.callback(like, ROI if has_face else None, ticks if has_face else [])
        if not has_face:
            assert out["cands"] is None and out["n_tf"] == 0 and np.all(out["joints2d"] == 0)
            continue
        Cn = 10 * N
        cands = np.stack(out["cands"])  # (2 hands, 2, C): cv::randu box on (re)acquisition, else the shim's getSamples
        assert cands.shape == (2, 2, Cn)
        blurred = out["blurred"]
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
            worst["x"] = max(worst["x"], rel_err(xr, xo))
            worst["P"] = max(worst["P"], rel_err(Pr, Po))
            e_node, p3_node = tr.pose(arm)
            _, e_orc = f.estimate()
            worst["pose"] = max(worst["pose"], rel_err(e_node, e_orc))
            worst["p3"] = max(worst["p3"], rel_err(p3_node, orc.get3dpose(e_orc)))
            poses.append(e_orc)
        tf, j2 = orc.skeleton(poses[0], poses[1])
        assert out["n_tf"] == 9
        worst["tf"] = max(worst["tf"], float(np.abs(tf - out["tf"]).max() / np.abs(tf).max()))
        worst["j2"] = max(worst["j2"], float(np.abs(j2 - out["joints2d"]).max() / np.abs(j2).max()))
        assert max(worst.values()) <= RTOL, (fr, worst)
        if fr >= 8:  # the node tracks the synthetic hand
            assert abs(poses[0][0] - (388 + 60 * np.sin(2 * np.pi * fr / 75))) < 60
    print("reference node source on the GPU shims vs oracle, worst relative errors:", worst)
    assert max(worst.values()) < 1e-8
