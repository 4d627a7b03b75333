"""Output back-end (the step after the path): PFTracker::get3Dpose, publishTFtree joint
differences and publish2Dpos joints (src/pfPose.cpp:84-208)."""
import os

import numpy as np
import pytest

import mkf_oracle as orc
import mkfbodytracker_pdaf_b200 as mk
from helpers import RTOL, rel_err, synth_frame, synth_u_init


def np_get3dpose(e, K):
    e = np.asarray(e, float).copy()
    roll, pitch, yaw = e[16], e[17], e[15]
    R1 = np.array([[1, 0, 0], [0, np.cos(roll), -np.sin(roll)], [0, np.sin(roll), np.cos(roll)]])
    R2 = np.array([[np.cos(pitch), 0, np.sin(pitch)], [0, 1, 0], [-np.sin(pitch), 0, np.cos(pitch)]])
    R3 = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
    P = K @ np.hstack([R3 @ R2 @ R1, e[18:21, None]])
    pts = np.stack([[e[3 * k] * e[3 * k + 2], e[3 * k + 1] * e[3 * k + 2], e[3 * k + 2]] for k in range(5)])
    return np.linalg.inv(P[:, :3]) @ (pts - P[:, 3]).T


def test_oracle_get3dpose_matches_numpy(left_arm, rng):
    nm = left_arm.np
    for _ in range(5):
        e = nm.proj.T @ (nm.means[rng.integers(0, 15)] + rng.standard_normal(12)) + nm.pmean
        for K in (orc.KINECT_K, np.array([[660.326889, 0, 318.70589], [0, 660.857176, 240.784699], [0, 0, 1.0]])):
            got = orc.get3dpose(e, K)
            want = np_get3dpose(e, K)
            assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
    # reprojection: K [R|t] X = z (x, y, 1) for every joint
    e = nm.pmean.copy()
    X = orc.get3dpose(e)
    roll, pitch, yaw = e[16], e[17], e[15]
    assert abs(roll) < 1e-2 and abs(pitch) < 1e-2 and abs(yaw) < 1e-2
    # skeleton bookkeeping
    e2 = e + 0.5
    tf, j2 = orc.skeleton(e, e2)
    p1, p2 = orc.get3dpose(e), orc.get3dpose(e2)
    assert np.allclose(tf[0], [p1[0, 0] - p1[0, 1], p1[2, 0] - p1[2, 1], -p1[1, 0] + p1[1, 1]])
    assert np.allclose(tf[3], [p2[0, 1] - p2[0, 2], p2[2, 1] - p2[2, 2], -p2[1, 1] + p2[1, 2]])
    neck = 0.5 * (p1[:, 4] + p2[:, 4])
    head = 0.5 * (p1[:, 3] + p2[:, 3])
    assert np.allclose(tf[4], [p1[0, 2] - neck[0], p1[2, 2] - neck[2], -p1[1, 2] + neck[1]])
    assert np.allclose(tf[6], [neck[0] - head[0], neck[2] - head[2], -neck[1] + head[1]])
    assert np.allclose(tf[7], [head[0], head[2], -head[1]])
    assert np.allclose(tf[8], [-e[18], -e[20], e[19]]) and np.allclose(tf[9], [-e[16], -e[17], -e[15]])
    assert np.allclose(j2[0], e[:2]) and np.allclose(j2[1], e2[:2]) and np.allclose(j2[2], 0.5 * (e[9:11] + e2[9:11]))
    assert np.allclose(j2[4], e[3:5]) and np.allclose(j2[7], e2[6:8])


def test_camera_matrix_loader():
    K = mk.load_camera_matrix(os.path.join(mk.MODEL_DIR, "webcam_camera_matrix.yml"))
    assert np.array_equal(K, np.array([[660.326889, 0, 318.70589], [0, 660.857176, 240.784699], [0, 0, 1]]))  # cal.yml:4-7
    with pytest.raises(mk.MkfError):
        mk.load_camera_matrix("/nonexistent/cal.yml")


@pytest.mark.gpu
def test_pose3d_and_skeleton_on_device(left_arm, right_arm):
    torch = pytest.importorskip("torch")
    T, N, seed = 5, 300, 0x5EED0002
    s = torch.cuda.Stream()
    b0 = mk.TrackBatch(left_arm.mk, T, N, stream=s.cuda_stream)
    b1 = mk.TrackBatch(right_arm.mk, T, N, stream=s.cuda_stream)
    u0 = synth_u_init(seed, range(T))
    f0 = [orc.Filter(left_arm.orc, N) for _ in range(T)]
    f1 = [orc.Filter(right_arm.orc, N) for _ in range(T)]
    for b, fs in ((b0, f0), (b1, f1)):
        b.reset(u0)
        for f, u in zip(fs, u0):
            f.reset(u=u)
    for fr in range(3):
        meas, ui, up = synth_frame(seed, range(T), fr)
        for b, fs in ((b0, f0), (b1, f1)):
            b.update(meas, ui, up)
            for t in range(T):
                fs[t].update(meas[t], ui[t], up[t])
    Kcal = mk.load_camera_matrix(os.path.join(mk.MODEL_DIR, "webcam_camera_matrix.yml"))
    for K in (None, Kcal):
        p3 = b0.pose3d(K)
        tf, j2 = mk.skeleton(b0, b1, K)
        for t in range(T):
            _, e1 = f0[t].estimate()
            _, e2 = f1[t].estimate()
            Ko = orc.KINECT_K if K is None else K
            want = orc.get3dpose(e1, Ko)
            assert rel_err(p3[t], want) <= RTOL and rel_err(p3[t], want) < 1e-9
            wtf, wj2 = orc.skeleton(e1, e2, Ko)
            assert np.abs(tf[t] - wtf).max() <= 1e-9 * np.abs(wtf).max()
            assert np.abs(j2[t] - wj2).max() <= 1e-12 * np.abs(wj2).max()


@pytest.mark.gpu
def test_candidate_proposal_front_end(left_arm, right_arm):
    """getSamples / first-frame box / likelihood lookup (src/pf2DRao.cpp:85-103, src/pfPose.cpp:216-236,254),
    then the whole callback chain propose -> associate -> update -> estimate against the oracle"""
    torch = pytest.importorskip("torch")
    from mkfbodytracker_pdaf_b200 import propose
    T, N, Cn, seed = 4, 200, 64, 0x5EED0003
    s = torch.cuda.Stream()
    b0 = mk.TrackBatch(left_arm.mk, T, N, stream=s.cuda_stream)
    b1 = mk.TrackBatch(right_arm.mk, T, N, stream=s.cuda_stream)
    u0 = synth_u_init(seed, range(T))
    f0 = [orc.Filter(left_arm.orc, N) for _ in range(T)]
    f1 = [orc.Filter(right_arm.orc, N) for _ in range(T)]
    for b, fs in ((b0, f0), (b1, f1)):
        b.reset(u0)
        for f, u in zip(fs, u0):
            f.reset(u=u)
    rng = np.random.default_rng(9)
    like = rng.integers(0, 256, (T, 480, 640), dtype=np.uint8)
    like[:, ::7, :] = 0  # exact zeros exercise the L != 0 gate
    roi = np.tile(np.array([300.0, 51.0, 47.0, 47.0]), (T, 1))
    roi[1] = [600.0, 400.0, 60.0, 50.0]  # box clipped by the image border
    for frame in range(3):
        tracking = np.array([0, 1, 1, 0] if frame == 0 else [1, 1, 0, 1], np.uint8)
        xy, Lv = propose(b0, b1, Cn, roi, tracking, like, seed=seed, frame=frame, track0=10)
        for t in range(T):
            wxy, wL = orc.propose(f0[t], f1[t], Cn, roi[t], int(tracking[t]), like[t], seed, 10 + t, frame)
            if tracking[t]:
                assert np.abs(xy[t] - wxy).max() <= 1e-9
                sd = np.std(xy[t] - np.array([f0[t].estimate()[1][:2], f1[t].estimate()[1][:2]])[:, :, None])
                assert abs(sd - 0.8 * roi[t, 2]) < 0.25 * 0.8 * roi[t, 2]  # quirk B10: 0.8*scale is a std-dev
            else:
                assert np.array_equal(xy[t], wxy)
                x0, y0, w, h = roi[t]
                assert xy[t][:, 0].min() >= max(int(x0) - int(4 * w), 0) and xy[t][:, 0].max() < min(int(x0) + int(5 * w), 640)
                assert xy[t][:, 1].min() >= min(int(y0) + int(h), 480) and xy[t][:, 1].max() <= min(int(y0) + int(7 * h), 480)
            assert np.array_equal(Lv[t], wL)
        # shared-image variant
        xy1, Lv1 = propose(b0, b1, Cn, roi, tracking, like[0], seed=seed, frame=frame, track0=10)
        assert np.array_equal(xy1, xy)
        inside = (xy[:, :, 0] > 0) & (xy[:, :, 0] < 640) & (xy[:, :, 1] > 0) & (xy[:, :, 1] < 480)
        want = np.where(inside, like[0][np.clip(xy[:, :, 1].astype(int), 0, 479), np.clip(xy[:, :, 0].astype(int), 0, 639)], 0)
        assert np.array_equal(Lv1, want.astype(np.uint8))
        # the rest of the callback on both sides
        u_c, u_i, u_p = rng.random((T, 2)), rng.random((T, 2)), rng.random((T, 2))
        mk.associate(b0, b1, xy, Lv, roi, u_c, u_i, u_p)
        d0, d1 = b0.download(state=True, cov=False), b1.download(state=True, cov=False)
        for t in range(T):
            wxy, wL = orc.propose(f0[t], f1[t], Cn, roi[t], int(tracking[t]), like[t], seed, 10 + t, frame)
            a = orc.associate(f0[t], f1[t], xy[t], Lv[t], roi[t], u_c[t])
            for arm, (f, d) in enumerate(((f0[t], d0), (f1[t], d1))):
                r = f.update(a["meas"][arm], u_i[t, arm], u_p[t, arm])
                assert np.array_equal(d["parents"][t], r["parents"])
                xo, _ = f.get_state()
                assert rel_err(d["x"][t], xo) <= RTOL


def test_bench_reference_arm_reads_models_without_the_product():
    """bench.py's CPU legs parse the model YAMLs themselves (the reference arm must not map libmkf_b200.so); the
    arrays equal what the product's loader (cv::FileStorage semantics, src/pfPose.cpp:34-55) returns"""
    import importlib.util
    import os
    import mkfbodytracker_pdaf_b200 as mk
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    got = bench.model_arrays(bench.LEFT_YML, bench.RIGHT_YML)
    want = mk.Model.load(mk.LEFT_ARM_MODEL, mk.RIGHT_ARM_MODEL).arrays()
    for k in ("means", "covs", "weights", "gamma", "pca_proj", "pca_mean"):
        assert np.array_equal(np.asarray(got[k]).reshape(-1), np.asarray(want[k]).reshape(-1)), k
    src = open(os.path.join(root, "bench.py")).read()
    cpu_part = src[src.index("def read_opencv_yaml"):src.index("_JSON_FD = None")]
    assert "mkfbodytracker_pdaf_b200 as" not in cpu_part and "import mkfbodytracker" not in cpu_part
