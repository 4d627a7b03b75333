"""The reference-named C++ host shims (include/mkf_shims.hpp: KF_model, my_gmm, state_params,
ParticleFilter) driven by tests/shim_driver.cpp the way pfPose.cpp drives the reference classes;
its trace is checked against the CPU oracle fed the same uniform draws."""
import os
import subprocess

import numpy as np
import pytest

import mkf_oracle as orc
import mkfbodytracker_pdaf_b200 as mk
from helpers import RTOL, synth_frame

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "shim_driver")


def test_shim_header_mirrors_reference_interfaces():
    """CPU check: every public name of src/KF_model.h, src/my_gmm.h, src/pf2DRao.h exists in the shim"""
    txt = open(os.path.join(ROOT, "include", "mkf_shims.hpp")).read()
    for name in ("class KF_model", "cv::Mat Q, R, F, B, H, BH", "void predict(cv::Mat& state, cv::Mat& cov)",
                 "void update(cv::Mat measurement, cv::Mat& state, cv::Mat& cov)", "class state_params",
                 "cv::Mat state;", "cv::Mat cov;", "double weight", "class my_gmm",
                 "void loadGaussian(cv::Mat u, cv::Mat s, cv::Mat& H, cv::Mat& m, double w, double g)",
                 "void resetTracker(std::vector<int> bins)", "std::vector<cv::Mat> mean;", "std::vector<cv::Mat> cov;",
                 "std::vector<double> weight;", "std::vector<KF_model> KFtracker;", "std::vector<state_params> tracks;",
                 "int nParticles", "class ParticleFilter", "ParticleFilter(int nParticles)",
                 "void update(cv::Mat measurement)", "cv::Mat getEstimator()",
                 "cv::Mat getSamples(cv::Mat H, cv::Mat M, int N, double scale)", "void getSampleProb(cv::Mat H, cv::Mat M",
                 "my_gmm gmm;", "std::vector<int> resample(std::vector<double> weights, int N)"):
        assert name in txt, name


@pytest.mark.gpu
def test_shim_driver_against_oracle(left_arm):
    if not os.path.exists(DRIVER):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests"), "shim_driver"], check=True)
    frames, N = 12, 500
    out = subprocess.run([DRIVER, mk.LEFT_ARM_MODEL, mk.RIGHT_ARM_MODEL, str(frames), str(N)], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = {}
    frames_out = []
    for ln in out.stdout.splitlines():
        tok = ln.split()
        if tok[0] == "FRAME":
            frames_out.append((int(tok[1]), float(tok[3]), float(tok[4]), np.array(tok[6:], float)))
        else:
            lines[tok[0]] = tok[1:]
    nm = left_arm.np
    a = left_arm.arrays
    # KF_model members as src/my_gmm.cpp:53-72 fills them
    kf = lines["KF0_Q00"]
    assert float(kf[0]) == a["Q"][0, 0, 0] and float(kf[2]) == a["B"][0, 0] and float(kf[4]) == a["H"][0, 0]
    assert float(kf[6]) == a["BH"][0] and float(kf[8]) == 100.0 and float(kf[10]) == a["gamma"][0]
    # standalone KF_model::predict / update
    xo, Po = left_arm.orc.kf_predict(3, nm.means[3], nm.covs[3])
    got = np.array(lines["KFPRED"], float)
    assert np.allclose(got[:12], xo, rtol=1e-9, atol=1e-9) and abs(got[12] - Po[0, 0]) <= 1e-9 * abs(Po[0, 0])
    assert abs(got[13] - Po[11, 2]) <= 1e-9 * np.abs(Po).max()
    z = orc.synth_meas(0x5EED0001, 0, 0, -1, 0)
    xu, Pu = left_arm.orc.kf_update(3, z, xo, Po)
    got = np.array(lines["KFUPD"], float)
    assert np.allclose(got[:12], xu, rtol=1e-9, atol=1e-8) and abs(got[12] - Pu[0, 0]) <= 1e-9 * np.abs(Pu).max()
    # the frame loop with the shim's own draws; the class shims default to the reference binary's shallow-copy
    # aliasing (quirk B3), the mode its own sources compute
    f = orc.Filter(left_arm.orc, N, alias_mode=orc.ALIAS_CV_SHALLOW_LITERAL)
    f.reset(u=float(lines["INIT_U"][0]))
    for fr, ui, up, xbar in frames_out:
        meas, _, _ = synth_frame(0x5EED0001, [0], fr, N, jitter=0)
        r = f.update(meas[0], ui, up)
        assert r["status"] == 0
        xb, _ = f.estimate()
        assert np.abs(xbar - xb).max() <= RTOL * np.abs(xb).max(), fr
    assert len(frames_out) == frames
    x_final, _ = f.get_state()
    assert np.abs(np.array(lines["TRACK0"], float) - x_final[0]).max() <= RTOL * np.abs(x_final[0]).max()
    # getSampleProb = 2-D isotropic density around the posterior hand estimate
    _, pose = f.estimate()
    pts = [(380, 250), (390, 260), (100, 400), (388, 250), (10, 20)]
    want = [orc.mvnpdf(np.array(p, float), pose[:2], 0.8 * 47.0 * np.eye(2))[0] for p in pts]
    got = np.array(lines["PROB"], float)
    assert np.allclose(got, want, rtol=1e-9, atol=0)
    sm = np.array(lines["SAMPLES_MEAN"], float)
    assert np.abs(sm - pose[:2]).max() < 4.0  # 2000 draws with sd 37.6
    assert lines["COPY"] == ["deep", "1", "shallow", "1"]
    assert lines["ERR"] == ["-1"]


def test_legacy_shim_header_mirrors_reference_interfaces():
    """CPU check: every public name of src/pf2D.h exists in the legacy shim"""
    txt = open(os.path.join(ROOT, "include", "mkf_shims_pf2d.hpp")).read()
    for name in ("class my_gmm", "void loadGaussian(cv::Mat u, cv::Mat s, double w)", "std::vector<cv::Mat> mean;",
                 "std::vector<cv::Mat> sigma_i;", "std::vector<double> det_s;", "std::vector<double> weight;", "int N;",
                 "class ParticleFilter", "ParticleFilter(int numParticles, int numDims, bool side1)", "ParticleFilter()",
                 "void predict()", "void update(cv::Mat measurement)", "cv::Mat getEstimator()", "my_gmm gmm;"):
        assert name in txt, name


@pytest.mark.gpu
@pytest.mark.parametrize("N,d", [(300, 8), (1000, 12)])
def test_legacy_shim_driver_against_oracle(N, d):
    """mkf_legacy::ParticleFilter (src/pf2D.h:25-51) driven from C++; the trace is replayed through the oracle's
    restatement of src/pf2D.cpp with the same particles, uniform and noise"""
    drv = os.path.join(ROOT, "tests", "shim_pf2d_driver")
    if not os.path.exists(drv):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests"), "shim_pf2d_driver"], check=True)
    frames = 3
    out = subprocess.run([drv, str(N), str(d), str(frames), "11"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr
    recs = []
    for ln in out.stdout.splitlines():
        tok = ln.split()
        if tok[0] in ("DET_S", "U", "RANGE"):
            recs.append((tok[0], np.array(tok[1:], float)))
        else:
            r, c = int(tok[1]), int(tok[2])
            recs.append((tok[0], np.array(tok[3:], float).reshape(r, c)))
    get = lambda tag: [v for t, v in recs if t == tag]
    means = np.concatenate(get("MEAN"))
    covs = np.stack(get("SIGMA"))
    K = means.shape[0]
    o = orc.Pf2d(N, means, covs, np.full(K, 1.0 / K))
    si, ds = o.gmm()
    # my_gmm::loadGaussian's public members (src/pf2D.cpp:28-37)
    assert np.allclose(np.stack(get("SIGMA_I")), si, rtol=1e-9, atol=0)
    assert np.allclose(np.concatenate(get("DET_S")), ds, rtol=1e-12)
    # constructor: column 6 on the half of the image `side` selects, others across the image (src/pf2D.cpp:58-70)
    rg = get("RANGE")[0]
    assert 321 <= rg[0] < 640 and 321 <= rg[1] < 640 and 1 <= rg[2] < 640 and 1 <= rg[3] < 480
    parts, meas, us, noise, after, est = (get(t) for t in ("PART", "MEAS", "U", "NOISE", "AFTER", "EST"))
    assert len(parts) == frames
    o.set_particles(parts[0])
    assert np.allclose(get("EST0")[0][0], o.estimate(), rtol=1e-12)
    for f in range(frames):
        o.set_particles(parts[f])  # teacher-forced per frame (float-expf ulp, quirk B12, may move an index)
        r = o.update(meas[f], float(us[f][0]), noise[f])
        po, wo = o.get()
        same = np.array_equal(after[f], po)
        # resampled + predicted particles: identical rows wherever the index agrees
        agree = (after[f] == po).all(axis=1).mean()
        assert agree >= 0.99, agree
        if same:
            eo = o.estimate()
            assert np.max(np.abs(est[f][0] - eo) / np.abs(eo)) <= 1e-5
        if f + 1 < frames:
            assert np.array_equal(parts[f + 1], after[f])
    # stand-alone predict(): N(0, 5) on the first eight dimensions only (src/pf2D.cpp:90-102)
    pred = get("PRED")[0]
    dlt = pred - after[-1]
    assert np.all(dlt[:, 8:] == 0) and 3.5 < dlt[:, :8].std() < 6.5


def test_shim_drivers_fail_loudly_without_a_gpu():
    """no CPU fallback behind the C++ shims either: without a device the drivers exit non-zero naming the reason"""
    if mk.device_count() > 0:
        pytest.skip("a GPU is present")
    for exe, args in (("shim_pf2d_driver", ["50", "8", "1", "3"]),):
        path = os.path.join(ROOT, "tests", exe)
        if not os.path.exists(path):
            subprocess.run(["make", "-C", os.path.join(ROOT, "tests"), exe], check=True)
        out = subprocess.run([path] + args, capture_output=True, text=True, timeout=60)
        assert out.returncode != 0 and "no CPU fallback" in out.stderr, out.stderr
