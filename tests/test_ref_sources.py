"""oracle/_ref -- the reference's own KF_model.cpp / my_gmm.cpp / pf2DRao.cpp compiled in place against the
OpenCV-subset shim -- pins the oracle restatement's reading of the reference: same control flow, same
operation sequence, same cv::Mat aliasing.  Skipped when neither /root/reference nor a prebuilt
oracle/_ref/libref.so is present."""
import numpy as np
import pytest

import mkf_oracle as orc
import mkf_ref
from helpers import synth_frame

pytestmark = pytest.mark.skipif(not mkf_ref.available(), reason="oracle/_ref/libref.so not built")


def model_arrays(arm):
    a = arm.arrays
    return {k: a[k] for k in ("means", "covs", "weights", "gamma", "pca_proj", "pca_mean")}


def test_load_gaussian_members_bit_exact(left_arm):
    rf = mkf_ref.RefFilter(model_arrays(left_arm), 10)
    c = left_arm.orc.constants()
    for k in (0, 7, 14):
        m = rf.kf_members(k)
        assert np.array_equal(m["Q"], c["Q"][k]) and np.array_equal(m["B"], c["B"][k])
        assert np.array_equal(m["H"], c["H"]) and np.array_equal(m["BH"], c["BH"]) and np.array_equal(m["R"], c["R"])
        assert np.array_equal(m["F"], left_arm.arrays["gamma"][k] * np.eye(12))


def test_kf_predict_update_bit_exact(left_arm, rng):
    rf = mkf_ref.RefFilter(model_arrays(left_arm), 10)
    nm = left_arm.np
    for k in (1, 9):
        x = nm.means[k] + rng.standard_normal(12)
        P = nm.covs[k]
        xr, Pr = rf.kf_predict(k, x, P)
        xo, Po = left_arm.orc.kf_predict(k, x, P)
        assert np.array_equal(xr, xo) and np.array_equal(Pr, Po)
        z = nm.H @ xo + nm.BH + rng.standard_normal(6) * 9
        xr2, Pr2 = rf.kf_update(k, z, xr, Pr)
        xo2, Po2 = left_arm.orc.kf_update(k, z, xo, Po)
        assert np.array_equal(xr2, xo2) and np.array_equal(Pr2, Po2)


def test_chol_and_mvnpdf_bit_exact(left_arm, rng):
    rf = mkf_ref.RefFilter(model_arrays(left_arm), 10)
    for _ in range(5):
        a = rng.standard_normal((6, 6))
        S = 300 * (a @ a.T + 6 * np.eye(6))
        want, ok = orc.chol(S, orc.CHOL_CV24_LITERAL)
        assert ok and np.array_equal(rf.chol(S), want)
        x, u = rng.standard_normal(6) * 40, rng.standard_normal(6) * 5
        assert rf.mvnpdf(x, u, S) == orc.mvnpdf(x, u, S, orc.CHOL_CV24_LITERAL)[0]
    bad = S.copy()
    bad[2, 2] = -5.0  # cv::Cholesky fails: the partially factored clone comes back
    want, ok = orc.chol(bad, orc.CHOL_CV24_LITERAL)
    assert not ok and np.array_equal(rf.chol(bad), want)
    S2 = 37.6 * np.eye(2)
    assert rf.mvnpdf(np.array([3.0, -4.0]), np.zeros(2), S2) == orc.mvnpdf(np.array([3.0, -4.0]), np.zeros(2), S2)[0]


def test_resample_bit_exact(left_arm, rng):
    rf = mkf_ref.RefFilter(model_arrays(left_arm), 10)
    for L, N in ((15, 500), (500, 500), (5000, 500)):
        w = rng.lognormal(0, 3, L)
        w /= w.sum()
        tick = int(rng.integers(1, 2**62))
        got = rf.resample(w, N, tick)
        want, deg = orc.resample(w, N, -1.0, seed=tick)       # oracle drawing from cv::RNG(seed) itself
        assert deg == 0 and np.array_equal(got, want)
        want2, _ = orc.resample(w, N, mkf_ref.tick_to_u(tick, L))  # and with the injected draw
        assert np.array_equal(got, want2)
    got = rf.resample(np.zeros(15), 40, 77)                   # degenerate fallback
    want, deg = orc.resample(np.zeros(15), 40, 0.5, seed=77)
    assert deg == 1 and np.array_equal(got, want)


@pytest.mark.parametrize("N", [60, 500])
def test_particle_filter_frames_match_oracle_literal_alias_mode(left_arm, N, rng):
    """the reference's ParticleFilter::update chains duplicates in place (quirk B3): it must equal the oracle
    in CV_SHALLOW_LITERAL mode bit for bit, and differ from INDEPENDENT mode from frame 2 on"""
    rf = mkf_ref.RefFilter(model_arrays(left_arm), N)
    fl = orc.Filter(left_arm.orc, N, alias_mode=orc.ALIAS_CV_SHALLOW_LITERAL)
    fi = orc.Filter(left_arm.orc, N, alias_mode=orc.ALIAS_INDEPENDENT)
    tick0 = 123456789
    rf.reset(tick0)
    for f in (fl, fi):
        f.reset(u=-1.0, seed=tick0)
    x0, P0 = rf.get_state()
    xo, Po = fl.get_state()
    assert np.array_equal(x0, xo) and np.array_equal(P0, Po)
    differs = False
    for fr in range(6):
        meas = synth_frame(0x5EED0001, [0], fr, N, jitter=0)[0][0]
        t_ind, t_post = int(rng.integers(1, 2**62)), int(rng.integers(1, 2**62))
        rf.update(meas, t_ind, t_post)
        rl = fl.update(meas, -1.0, -1.0, seed_ind=t_ind, seed_post=t_post)
        fi.update(meas, -1.0, -1.0, seed_ind=t_ind, seed_post=t_post)
        assert rl["status"] == 0
        xr, Pr = rf.get_state()
        xo, Po = fl.get_state()
        assert np.array_equal(xr, xo), f"frame {fr}: state differs from the literal-alias oracle"
        assert np.array_equal(Pr, Po)
        assert np.array_equal(rf.estimate(), fl.estimate()[0])
        xi, _ = fi.get_state()
        if fr == 0:
            assert np.array_equal(xr, xi)  # frame 1: every slot still owns its buffers
        else:
            differs |= not np.allclose(xr, xi, rtol=1e-6, atol=1e-6)
    assert differs


def test_sample_prob_matches_oracle(left_arm):
    N = 100
    rf = mkf_ref.RefFilter(model_arrays(left_arm), N)
    f = orc.Filter(left_arm.orc, N, alias_mode=orc.ALIAS_CV_SHALLOW_LITERAL)
    rf.reset(42)
    f.reset(u=-1.0, seed=42)
    in1 = np.array([[380.0, 390.0, 100.0], [250.0, 260.0, 400.0]])
    in2 = np.array([[388.0, 10.0], [250.0, 20.0]])
    w1, w2 = rf.sample_prob(in1, in2, 47.0)
    _, pose = f.estimate()
    for pts, got in ((in1, w1), (in2, w2)):
        for c in range(pts.shape[1]):
            want = orc.mvnpdf(pts[:, c], pose[:2], 0.8 * 47.0 * np.eye(2))[0]
            assert abs(got[c] - want) <= 1e-13 * want
    m, s = rf.samples_mean_sd(20000, 47.0)  # quirk B10: 0.8*scale used as a standard deviation
    assert np.abs(m - pose[:2]).max() < 1.5 and np.abs(s - 0.8 * 47.0).max() < 1.0
