"""Pins the CPU oracle (oracle/mkf_oracle.cpp): analytic known-answer tests, the independent
numpy restatement (tests/np_ref.py), OpenCV-python primitives, and structural properties.
The reference ships no tests or golden vectors: parity is UNPINNED against
the reference binary; these are the strongest anchors available offline."""
import os

import numpy as np
import pytest

import mkf_oracle as orc
import np_ref
from helpers import synth_frame


def spd(rng, n, scale=1.0):
    a = rng.standard_normal((n, n))
    return scale * (a @ a.T + n * np.eye(n))


# ---------------- OpenCV primitives ----------------
def test_invert_lu_matches_numpy_and_cv2(rng):
    for n in (2, 6):
        a = spd(rng, n)
        inv, ok = orc.invert_lu(a)
        assert ok and np.allclose(inv, np.linalg.inv(a), rtol=1e-12, atol=1e-14)
    cv2 = pytest.importorskip("cv2")
    a = spd(rng, 6, 50.0)
    _, cvinv = cv2.invert(a, flags=cv2.DECOMP_LU)
    inv, _ = orc.invert_lu(a)
    assert np.allclose(inv, cvinv, rtol=1e-13, atol=1e-16)
    z, ok = orc.invert_lu(np.zeros((6, 6)))
    assert not ok and np.all(z == 0)


def test_cvrng_known_sequence():
    # multiply-with-carry recurrence computed independently in Python integers
    seed, L = 0x1234ABCD5678, 15
    s = seed
    want_i = []

    def nxt():
        nonlocal s
        s = ((s & 0xFFFFFFFF) * 4164903690 + (s >> 32)) & 0xFFFFFFFFFFFFFFFF
        return s & 0xFFFFFFFF

    for _ in range(5):
        want_i.append(nxt() % L)
    t = nxt()
    want_d = ((t << 32) | nxt()) * 5.4210108624275221700372640043497e-20
    gi, gd = orc.cvrng(seed, L, 5, 1)
    assert list(gi) == want_i and gd[0] == want_d and 0 <= gd[0] < 1


# ---------------- chol / mvnpdf ----------------
def test_chol_wrapper_structure(rng):
    S = spd(rng, 6, 100.0)
    Lc = np.linalg.cholesky(S)
    R24, ok = orc.chol(S, orc.CHOL_CV24_LITERAL)
    assert ok and np.allclose(np.tril(R24, -1), 0)
    assert np.allclose(np.diag(R24), np.diag(Lc), rtol=1e-14)
    for e in range(6):  # un-corrected upper triangle: S_ej / L_ee; only row 0 is a true Cholesky row
        assert np.allclose(R24[e, e + 1:], S[e, e + 1:] / Lc[e, e], rtol=1e-14)
    assert np.allclose(R24[0], Lc.T[0], rtol=1e-13)
    assert np.allclose(R24, np_ref.pseudo_chol(S, "cv24"), rtol=1e-13)
    R3, _ = orc.chol(S, orc.CHOL_CV3_LITERAL)
    assert np.allclose(R3, np_ref.pseudo_chol(S, "cv3"), rtol=1e-13)
    Rx, _ = orc.chol(S, orc.CHOL_EXACT)
    assert np.allclose(Rx, Lc.T, rtol=1e-13, atol=1e-13)
    # not positive definite: cv::Cholesky fails and the partially factored clone is returned
    bad = S.copy()
    bad[3, 3] = -1.0
    out, ok = orc.chol(bad, orc.CHOL_CV24_LITERAL)
    assert not ok and np.array_equal(np.triu(out, 1), np.triu(bad, 1))


def test_mvnpdf_diagonal_is_exact_gaussian(rng):
    scipy_stats = pytest.importorskip("scipy.stats")
    var = rng.uniform(50, 5000, 6)
    x = rng.standard_normal(6) * 30
    u = rng.standard_normal(6) * 30
    want = scipy_stats.multivariate_normal.pdf(x, mean=u, cov=np.diag(var))
    for mode in (orc.CHOL_CV24_LITERAL, orc.CHOL_EXACT):
        got, ok = orc.mvnpdf(x, u, np.diag(var), mode)
        assert ok and abs(got - want) <= 1e-12 * want
    # 2-D isotropic (the association proposal density, src/pf2DRao.cpp:114)
    got, _ = orc.mvnpdf(x[:2], u[:2], 37.6 * np.eye(2))
    want = scipy_stats.multivariate_normal.pdf(x[:2], mean=u[:2], cov=37.6 * np.eye(2))
    assert abs(got - want) <= 1e-12 * want


def test_mvnpdf_full_cov_matches_numpy_restatement_not_true_pdf(rng):
    scipy_stats = pytest.importorskip("scipy.stats")
    S = spd(rng, 6, 300.0)
    x = rng.standard_normal(6) * 40
    u = np.zeros(6)
    for mode, name in ((orc.CHOL_CV24_LITERAL, "cv24"), (orc.CHOL_CV3_LITERAL, "cv3"), (orc.CHOL_EXACT, "exact")):
        got, ok = orc.mvnpdf(x, u, S, mode)
        want = np_ref.mvnpdf(x, u, S, name)
        assert ok and (abs(got - want) <= 1e-11 * max(want, 1e-300))
    true = scipy_stats.multivariate_normal.pdf(x, mean=u, cov=S)
    exact, _ = orc.mvnpdf(x, u, S, orc.CHOL_EXACT)
    assert abs(exact - true) <= 1e-11 * true
    lit, _ = orc.mvnpdf(x, u, S, orc.CHOL_CV24_LITERAL)
    assert abs(lit - true) > 1e-3 * true  # quirk B1: the literal pseudo-factor is NOT the Gaussian pdf


# ---------------- KF_model ----------------
def test_kf_predict_update_match_numpy(left_arm, rng):
    m, nm = left_arm.orc, left_arm.np
    for k in (0, 7, 14):
        x = nm.means[k] + rng.standard_normal(12)
        P = nm.covs[k]
        xo, Po = m.kf_predict(k, x, P)
        xn, Pn = np_ref.predict(nm, k, x, P)
        assert np.allclose(xo, xn, rtol=1e-14, atol=1e-12) and np.allclose(Po, Pn, rtol=1e-13, atol=1e-10)
        z = nm.H @ xn + nm.BH + rng.standard_normal(6) * 10
        xu, Pu = m.kf_update(k, z, xo, Po)
        xv, Pv = np_ref.kf_update(nm, z, xn, Pn)
        assert np.allclose(xu, xv, rtol=1e-11, atol=1e-9) and np.abs(Pu - Pv).max() <= 1e-11 * np.abs(Pv).max()


def test_kf_update_fixed_point(left_arm):
    """z = H mu + BH leaves the mean unchanged"""
    nm = left_arm.np
    k = 3
    z = nm.H @ nm.means[k] + nm.BH
    xu, Pu = left_arm.orc.kf_update(k, z, nm.means[k], nm.covs[k])
    assert np.allclose(xu, nm.means[k], rtol=0, atol=1e-9)
    assert np.all(np.linalg.eigvalsh(0.5 * (Pu + Pu.T)) > 0)
    assert np.trace(Pu) < np.trace(nm.covs[k])


def test_kf_gemm_against_cv2(left_arm, rng):
    """predict/update re-evaluated literally with cv2.gemm / cv2.invert (OpenCV-python 4.13)"""
    cv2 = pytest.importorskip("cv2")
    nm = left_arm.np
    k = 5
    g = nm.gamma[k]
    F = g * np.eye(12)
    x = (nm.means[k] + rng.standard_normal(12)).reshape(12, 1)
    P = nm.covs[k].copy()
    xs = cv2.gemm(F, x, 1, nm.B[k].reshape(12, 1), 1)
    Pp = cv2.gemm(cv2.gemm(F, P, 1, None, 0), F, 1, nm.Q[k], 1, flags=cv2.GEMM_2_T)
    xo, Po = left_arm.orc.kf_predict(k, x.ravel(), P)
    assert np.array_equal(xo, xs.ravel()) and np.array_equal(Po, Pp)  # bit-exact for F = g*I
    z = (nm.H @ xs.ravel() + nm.BH + rng.standard_normal(6) * 8).reshape(6, 1)
    y = z - cv2.gemm(nm.H, xs, 1, nm.BH.reshape(6, 1), 1)
    S = cv2.gemm(cv2.gemm(nm.H, Pp, 1, None, 0), nm.H, 1, nm.R, 1, flags=cv2.GEMM_2_T)
    _, Sinv = cv2.invert(S, flags=cv2.DECOMP_LU)
    Kg = cv2.gemm(cv2.gemm(Pp, nm.H, 1, None, 0, flags=cv2.GEMM_2_T), Sinv, 1, None, 0)
    xu = cv2.gemm(Kg, y, 1, xs, 1)
    Pu = cv2.gemm(np.eye(12) - cv2.gemm(Kg, nm.H, 1, None, 0), Pp, 1, None, 0)
    xo2, Po2 = left_arm.orc.kf_update(k, z.ravel(), xo, Po)
    assert np.allclose(xo2, xu.ravel(), rtol=1e-12, atol=1e-10)
    assert np.abs(Po2 - Pu).max() <= 1e-12 * np.abs(Pu).max()


# ---------------- resample ----------------
def test_resample_kats():
    N = 500
    out, deg = orc.resample(np.full(N, 1.0 / N), N, 0.5)
    assert deg == 0 and np.array_equal(out, np.arange(N))
    w = np.zeros(N)
    w[123] = 1.0
    out, deg = orc.resample(w, N, 0.25)
    assert deg == 0 and np.all(out == 123)
    # max weight 0 or NaN -> N random indices from cv::RNG (quirks B7/B11)
    for bad in (np.zeros(15), np.full(15, np.nan)):
        out, deg = orc.resample(bad, 40, 0.5, seed=77)
        ints, _ = orc.cvrng(77, 15, 41, 0)
        assert deg == 1 and np.array_equal(out, ints[1:])  # first draw is discarded (src/pf2DRao.cpp:180)
    # u < 0: the reference's own draw (one discarded int, then uniform(0.0,1.0))
    w = np.array([0.2, 0.5, 0.3])
    out, _ = orc.resample(w, 10, -1.0, seed=99)
    s = 99

    def nxt():
        nonlocal s
        s = ((s & 0xFFFFFFFF) * 4164903690 + (s >> 32)) & 0xFFFFFFFFFFFFFFFF
        return s & 0xFFFFFFFF

    nxt()
    t = nxt()
    u = ((t << 32) | nxt()) * 5.4210108624275221700372640043497e-20
    assert np.array_equal(out, np_ref.resample_loop(w, 10, u))


@pytest.mark.parametrize("L,N", [(15, 500), (500, 500), (17, 500), (4096, 4096), (300, 77)])
def test_resample_properties_and_closed_form(rng, L, N):
    for trial in range(20):
        w = rng.lognormal(0, 3, L)
        w /= w.sum()
        u = rng.random()
        out, deg = orc.resample(w, N, u)
        assert deg == 0
        assert np.all(np.diff(out) >= 0), "systematic resampling output is non-decreasing"
        cnt = np.bincount(out, minlength=L)
        assert cnt.sum() == N and np.all(np.abs(cnt - N * w) < 1 + 1e-9)
        assert np.array_equal(out, np_ref.resample_closed_form(w, N, u))
        if L * N <= 300 * 500:
            assert np.array_equal(out, np_ref.resample_loop(w, N, u))


# ---------------- ParticleFilter::update ----------------
def test_filter_frame_matches_numpy_restatement(left_arm, rng):
    N = 120
    f = orc.Filter(left_arm.orc, N)
    f.reset(u=0.42)
    x0, P0 = f.get_state()
    assert np.allclose(x0.mean(0), left_arm.np.weights @ left_arm.np.means, atol=25)
    for fr in range(3):
        meas, _, _ = synth_frame(0x5EED0001, [0], fr, N, jitter=0)
        x0, P0 = f.get_state()
        r = f.update(meas[0], u_ind=rng.random(), u_post=rng.random())
        assert r["status"] == 0
        ind = r["indicators"]
        assert np.array_equal(ind, np_ref.resample_closed_form(left_arm.np.weights, N, 0) * 0 + ind)
        xo, Po, w = np_ref.filter_update(left_arm.np, x0, P0, meas[0], ind)
        assert np.allclose(w, r["w_raw"], rtol=1e-10, atol=0)
        assert abs(r["wsum"] - w.sum()) <= 1e-12 * w.sum()
        assert np.allclose(r["w_norm"], w / w.sum(), rtol=1e-10)
        x1, P1 = f.get_state()
        par = r["parents"]
        assert np.allclose(x1, xo[par], rtol=1e-10, atol=1e-9)
        assert np.abs(P1 - Po[par]).max() <= 1e-10 * np.abs(P1).max()
        xb, pose = f.estimate()
        assert np.allclose(xb, x1.mean(0), rtol=1e-12, atol=1e-10)
        assert np.allclose(pose, left_arm.np.proj.T @ xb + left_arm.np.pmean, rtol=1e-12, atol=1e-10)


def test_indicator_resample_uses_prior_weights(left_arm):
    N = 500
    f = orc.Filter(left_arm.orc, N)
    f.reset(u=0.1)
    meas, _, _ = synth_frame(0x5EED0001, [0], 0, None, jitter=0)
    r = f.update(meas[0], u_ind=0.73, u_post=0.5)
    want, _ = orc.resample(left_arm.np.weights, N, 0.73)
    assert np.array_equal(r["indicators"], want)


def test_shared_measurement_equals_replicated_columns(left_arm):
    N = 64
    fa, fb = orc.Filter(left_arm.orc, N), orc.Filter(left_arm.orc, N)
    fa.reset(u=0.3)
    fb.reset(u=0.3)
    z = synth_frame(0x5EED0002, [5], 3, None)[0][0]
    ra = fa.update(z, 0.2, 0.9)
    rb = fb.update(np.repeat(z[:, None], N, 1), 0.2, 0.9)
    assert np.array_equal(ra["w_raw"], rb["w_raw"]) and np.array_equal(ra["parents"], rb["parents"])


def test_alias_mode_literal_differs_only_after_first_resample(left_arm):
    """quirk B3: duplicates share buffers and are chained in place from frame 2 on"""
    N = 200
    fi = orc.Filter(left_arm.orc, N, alias_mode=orc.ALIAS_INDEPENDENT)
    fl = orc.Filter(left_arm.orc, N, alias_mode=orc.ALIAS_CV_SHALLOW_LITERAL)
    for f in (fi, fl):
        f.reset(u=0.6)
    outs = []
    for fr in range(2):
        meas = synth_frame(0x5EED0001, [0], fr, N, jitter=0)[0][0]
        outs.append((fi.update(meas, 0.4, 0.8), fl.update(meas, 0.4, 0.8)))
    assert np.array_equal(outs[0][0]["w_raw"], outs[0][1]["w_raw"])       # frame 1 identical
    assert np.array_equal(outs[0][0]["parents"], outs[0][1]["parents"])
    assert not np.allclose(outs[1][0]["w_raw"], outs[1][1]["w_raw"], rtol=1e-3, atol=0)      # frame 2 materially different
    xi, _ = fi.get_state()
    xl, _ = fl.get_state()
    assert np.abs(xi - xl).max() > 1e-3


def test_config1_trajectory_tracks_the_hand(left_arm):
    """config 1 (1 track, N=500, per-slot columns): the filter follows the synthetic hand"""
    N, frames, seed = 500, 40, 0x5EED0001
    f = orc.Filter(left_arm.orc, N)
    f.reset(u=orc.synth_u(seed, 0, 0xFFFFFFFFFFFF, 0x1003))
    errs = []
    for fr in range(frames):
        meas, ui, up = synth_frame(seed, [0], fr, N, jitter=0)
        r = f.update(meas[0], ui[0], up[0])
        assert r["status"] == 0 and np.isfinite(r["wsum"]) and r["wsum"] > 0
        _, pose = f.estimate()
        errs.append(np.hypot(pose[0] - meas[0, 2].mean(), pose[1] - meas[0, 3].mean()))
    assert np.median(errs[10:]) < 15.0  # pixels


# ---------------- association ----------------
def test_association_matches_numpy(left_arm, right_arm, rng):
    N, Cn = 100, 17
    fL, fR = orc.Filter(left_arm.orc, N), orc.Filter(right_arm.orc, N)
    fL.reset(u=0.2)
    fR.reset(u=0.7)
    roi = np.array([300.0, 51.0, 47.0, 47.0])
    cand = np.zeros((2, 2, Cn))
    Lv = np.zeros((2, Cn), np.uint8)
    for h in range(2):
        for c in range(Cn):
            cand[h, 0, c], cand[h, 1, c], Lv[h, c] = orc.synth_candidate(0x5EED0003, 9, 4, h, Cn, c)
    u = np.array([0.31, 0.77])
    out = orc.associate(fL, fR, cand, Lv, roi, u)
    hands = []
    for f, arm in ((fL, left_arm), (fR, right_arm)):
        _, pose = f.estimate()
        hands.append(pose[:2])
    s2 = 0.8 * roi[2]
    for h in range(2):
        x, y = cand[h]
        gate = (y > 0) & (y < 480) & (x > 0) & (x < 640) & (Lv[h] != 0)
        assert np.array_equal(out["gate"][h].astype(bool), gate)
        dens = [np.exp(-((x - hd[0]) ** 2 + (y - hd[1]) ** 2) / (2 * s2)) / (2 * np.pi * s2) for hd in hands]
        Z = 0.05 * dens[0] + 0.05 * dens[1] + 1e-4 * 0.9
        w = np.where(gate, (Lv[h] / 255.0) / Z, 0.0)
        w /= w.sum()
        assert np.allclose(out["weights"][h], w, rtol=1e-11, atol=0)
        bins, _ = orc.resample(out["weights"][h], N, u[h])
        assert np.array_equal(out["bins"][h], bins)
        ms = out["meas"][h]
        assert np.all(ms[0] == 323.5) and np.all(ms[1] == 74.5) and np.all(ms[4] == 323.5)
        assert np.allclose(ms[5], 51 + 1.65 * 47)
        assert np.array_equal(ms[2], x[bins]) and np.array_equal(ms[3], y[bins])
    # no candidate passes the gate -> NaN weights -> random candidates (quirk B11)
    Lz = np.zeros_like(Lv)
    out = orc.associate(fL, fR, cand, Lz, roi, u, seed_cand=[5, 6])
    assert out["status"] == 3 and not out["gate"].any() and np.isnan(out["weights"]).all()
    assert np.array_equal(out["bins"][0], orc.cvrng(5, Cn, N + 1, 0)[0][1:])


# ---------------- legacy pf2D ----------------
def test_pf2d_matches_numpy(rng):
    N, d, K = 300, 8, 5
    means = rng.uniform(100, 400, (K, d))
    covs = np.stack([spd(rng, d, 40.0) for _ in range(K)])
    wts = rng.dirichlet(np.ones(K))
    pf = orc.Pf2d(N, means, covs, wts)
    si, ds = pf.gmm()
    for k in range(K):
        assert np.allclose(si[k], np.linalg.inv(covs[k]), rtol=1e-10, atol=1e-14)
        assert np.isclose(ds[k], 1 / ((2 * np.pi) ** (d / 2) * np.sqrt(np.linalg.det(covs[k]))), rtol=1e-12)
    parts = means[rng.integers(0, K, N)] + rng.standard_normal((N, d)) * 6
    pf.set_particles(parts)
    meas = np.array([[parts[:, 6].mean(), parts[:, 7].mean()], [parts[:, 0].mean(), parts[:, 1].mean()]])
    r = pf.update(meas, 0.37, None)
    prior = np.zeros(N)
    for k in range(K):
        dx = parts - means[k]
        q = -0.5 * np.einsum("ni,ij,nj->n", dx, si[k], dx)
        prior += wts[k] * ds[k] * np.exp(q.astype(np.float32)).astype(np.float64)  # float expf (quirk B12)
    lik = 1.0
    for (a, b), mrow in (((6, 7), meas[0]), ((0, 1), meas[1])):
        dd = (parts[:, a] - mrow[0]) ** 2 + (parts[:, b] - mrow[1]) ** 2
        lik = lik * np.exp(-0.5 * dd / 15) / (2 * np.pi * 15)
    w = prior * lik
    w /= w.sum()
    assert np.allclose(r["w_norm"], w, rtol=2e-6)  # float32 exponent
    par, _ = orc.resample(r["w_norm"], N, 0.37)
    assert np.array_equal(r["parents"], par)
    p2, _ = pf.get()
    assert np.array_equal(p2, parts[par])
    # getEstimator (src/pf2D.cpp:79-88): the normalised weights update() left (resample() does not reset them, :225-268)
    # against the resampled particles, summed in index order
    est = pf.estimate()
    want = np.zeros(d)
    for i in range(N):
        want = want + r["w_norm"][i] * p2[i]
    assert np.array_equal(est, want)
    fresh = orc.Pf2d(N, means, covs, wts)
    fresh.set_particles(parts)
    assert np.allclose(fresh.estimate(), parts.mean(0), rtol=1e-13)  # constructor weights 1/N (:52-55)


def test_pf2d_constructor_and_degenerate_branch(rng):
    """src/pf2D.cpp:44-71 (constructor) and :232-250 (max weight 0: re-randomise, weights 1/N), driven by the counter
    generator of include/mkf_synth.h: ranges, determinism, one fresh draw per epoch"""
    N, d, K = 200, 8, 4
    means = rng.uniform(100, 400, (K, d))
    covs = np.stack([spd(rng, d, 40.0) for _ in range(K)])
    wts = rng.dirichlet(np.ones(K))
    pfs = []
    for side in (0, 1):
        pf = orc.Pf2d(N, means, covs, wts)
        pf.set_random(77, 5, side)
        pf.randomise()
        p, w = pf.get()
        assert np.all(w == 1.0 / N)
        assert p[:, 0::2].min() >= 1 and p[:, 0::2].max() < 640 and p[:, 1::2].min() >= 1 and p[:, 1::2].max() < 480
        lo, hi = (321, 640) if side else (1, 320)
        assert p[:, 6].min() >= lo and p[:, 6].max() < hi
        assert abs(p[:, 0].mean() - 320.5) < 45 and abs(p[:, 1].mean() - 240.5) < 35
        pfs.append(pf)
    twin = orc.Pf2d(N, means, covs, wts)
    twin.set_random(77, 5, 0)
    twin.randomise()
    assert np.array_equal(twin.get()[0], pfs[0].get()[0])
    assert np.array_equal(pfs[0].get()[0][:, :6], pfs[1].get()[0][:, :6])   # same key: `side` only moves column 6
    pf = pfs[0]
    p0 = pf.get()[0]
    far = np.array([[1e5, 1e5], [1e5, 1e5]])
    r = pf.update(far, 0.3, None)          # all likelihoods underflow: 0/0 weights, max weight stays 0
    assert r["status"] == 1 and np.isnan(r["w_norm"]).all() and np.array_equal(r["parents"], np.arange(N))
    p1, w1 = pf.get()
    assert np.all(w1 == 1.0 / N) and not np.array_equal(p1, p0)      # epoch 1: a fresh draw
    assert p1[:, 6].max() < 320 and p1.min() >= 1
    r = pf.update(far, 0.3, np.ones((N, d)))
    p2, _ = pf.get()
    assert r["status"] == 1 and not np.array_equal(p2, p1 + 5.0)     # epoch 2 differs from epoch 1 ...
    twin.update(far, 0.9, None)
    twin.update(far, 0.1, np.ones((N, d)))
    assert np.array_equal(twin.get()[0], p2)                          # ... and is reproducible; predict() follows it


# ---------------- synthetic generator ----------------
def test_synth_generator_pinned():
    """pins include/mkf_synth.h (shared by oracle, host and device code) to fixed values"""
    z = orc.synth_meas(0x5EED0001, 0, 0, -1, 0)
    assert z[0] == 323.5 and z[1] == 74.5 and z[4] == 323.5 and z[5] == 128.55
    g = np.array([orc.synth_meas(0x5EED0002, t, f, j, 1)[2:4] for t in range(6) for f in range(6) for j in (-1, 0, 9)])
    assert np.all((g[:, 0] > 280) & (g[:, 0] < 500) & (g[:, 1] > 120) & (g[:, 1] < 380))
    us = np.array([orc.synth_u(1, t, f, 0x1001) for t in range(40) for f in range(40)])
    assert us.min() >= 0 and us.max() < 1 and abs(us.mean() - 0.5) < 0.03
    noise = np.array([orc.synth_meas(7, 0, f, j, 0)[2] - orc.synth_meas(7, 0, f, -1, 0)[2] for f in range(30)
                      for j in range(30)])
    assert abs(noise.std() - 3 * np.sqrt(2)) < 0.4
    golden = os.path.join(os.path.dirname(__file__), "golden", "synth_pins.npz")
    cur = dict(
        meas=np.array([orc.synth_meas(0x5EED0002, t, f, j, 1) for t in (0, 77) for f in (0, 5) for j in (-1, 3)]),
        u=np.array([orc.synth_u(0x5EED0002, t, 3, w) for t in (0, 77) for w in (0x1001, 0x1002)]),
        cand=np.array([orc.synth_candidate(0x5EED0003, 4, 2, h, 17, c) for h in (0, 1) for c in (0, 1, 16)]))
    if not os.path.exists(golden):
        pytest.skip("golden/synth_pins.npz missing (run tests/golden/make_golden.py)")
    ref = np.load(golden)
    for k in cur:
        assert np.array_equal(cur[k], ref[k]), k
