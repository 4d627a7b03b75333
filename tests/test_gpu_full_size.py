"""BASELINE.json configs 3, 4, 5 at their full sizes: the oracle cannot run them in seconds, so parity is checked
through size-independent properties plus oracle spot-checks of a few tracks (config 2 full size lives in
test_gpu_parity.py::test_full_size_properties_config2)."""
import numpy as np
import pytest

import mkf_oracle as orc
import mkfbodytracker_pdaf_b200 as mk
from helpers import RTOL, rel_err, rel_err_weights, synth_frame, synth_u_init
from mkfbodytracker_pdaf_b200 import _lib as L_

pytestmark = pytest.mark.gpu


def check_resample_properties(par, wn, N, rows):
    assert par.min() >= 0 and par.max() < N
    assert np.all(np.diff(par, axis=1) >= 0)
    for t in rows:
        cnt = np.bincount(par[t], minlength=N)
        assert cnt.sum() == N and np.all(np.abs(cnt - N * wn[t]) < 1 + 1e-9)


def test_config5_full_size_one_million_tracks(right_arm):
    torch = pytest.importorskip("torch")
    seed, T, N = 0x5EED0005, 1 << 20, 15
    dev = torch.device("cuda:0")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        b = mk.TrackBatch(right_arm.mk, T, N, stream=s.cuda_stream)
        meas = torch.empty((T, 6), dtype=torch.float64, device=dev)
        ui = torch.empty(T, dtype=torch.float64, device=dev)
        up = torch.empty(T, dtype=torch.float64, device=dev)
        u0 = torch.empty(T, dtype=torch.float64, device=dev)
        b.synth_fill(seed, 0, 0xFFFFFF, 1, mk.MEAS_SHARED, meas, u0, None)
        b.reset(u0)
        frames = 3
        for fr in range(frames):
            b.synth_fill(seed, 0, fr, 1, mk.MEAS_SHARED, meas, ui, up)
            b.update(meas, ui, up)
        d = b.download(state=False, cov=False)
        xb, pose = b.estimate()
    assert not d["status"].any()
    assert np.allclose(d["w_norm"].sum(1), 1.0, rtol=0, atol=1e-12)
    check_resample_properties(d["parents"], d["w_norm"], N, range(0, T, 65537))
    assert np.all(np.diff(d["indicators"], axis=1) >= 0) and np.isfinite(pose).all()
    u0h = u0.cpu().numpy()
    for t in (0, 524287, T - 1):  # oracle spot-check, free-running
        f = orc.Filter(right_arm.orc, N)
        f.reset(u=u0h[t])
        for fr in range(frames):
            mz, a_, c_ = synth_frame(seed, [t], fr)
            r = f.update(mz[0], a_[0], c_[0])
        assert np.array_equal(d["parents"][t], r["parents"])
        assert rel_err_weights(d["w_norm"][t], r["w_norm"]) <= RTOL
        xo, po = f.estimate()
        assert rel_err(xb[t], xo) <= RTOL and rel_err(pose[t], po) <= RTOL


def test_config4_full_size(left_arm):
    torch = pytest.importorskip("torch")
    seed, T, N = 0x5EED0004, 256, 65536
    dev = torch.device("cuda:0")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        b = mk.TrackBatch(left_arm.mk, T, N, stream=s.cuda_stream)
        meas = torch.empty((T, 6, N), dtype=torch.float64, device=dev)
        ui = torch.empty(T, dtype=torch.float64, device=dev)
        up = torch.empty(T, dtype=torch.float64, device=dev)
        u0 = torch.tensor(synth_u_init(seed, range(T)), device=dev)
        b.reset(u0)
        frames = 2
        for fr in range(frames):
            b.synth_fill(seed, 0, fr, 0, mk.MEAS_PER_SLOT, meas, ui, up)
            b.update(meas, ui, up)
        par = torch.empty((T, N), dtype=torch.int32, device=dev)
        wsum = torch.empty(T, dtype=torch.float64, device=dev)
        status = torch.empty(T, dtype=torch.int32, device=dev)
        mk._lib.check(mk._lib.lib.mkf_batch_download(b._h, None, None, None, None, None, par.data_ptr(),
                                                     wsum.data_ptr(), status.data_ptr(), mk.MEM_DEVICE))
        xb, pose = b.estimate()
    par = par.cpu().numpy()
    assert not (status.cpu().numpy() & 0x4C).any() and (wsum.cpu().numpy() > 0).all()
    assert par.min() >= 0 and par.max() < N and np.all(np.diff(par, axis=1) >= 0)
    # oracle spot-check of one full 65 536-slot track, free-running over both frames
    t = 131
    f = orc.Filter(left_arm.orc, N)
    f.reset(u=float(u0[t].cpu()))
    for fr in range(frames):
        mz, a_, c_ = synth_frame(seed, [t], fr, N, jitter=0)
        r = f.update(mz[0], a_[0], c_[0])
    assert np.array_equal(par[t], r["parents"])
    xo, po = f.estimate()
    assert rel_err(xb[t], xo) <= RTOL and rel_err(pose[t], po) <= RTOL


def test_config3_full_size_gate_and_bins(left_arm, right_arm):
    torch = pytest.importorskip("torch")
    seed, T, N, Cn = 0x5EED0003, 16384, 15, 17
    s = torch.cuda.Stream()
    b0 = mk.TrackBatch(left_arm.mk, T, N, stream=s.cuda_stream)
    b1 = mk.TrackBatch(right_arm.mk, T, N, stream=s.cuda_stream)
    rng = np.random.default_rng(3)
    u0 = rng.random(T)
    b0.reset(u0)
    b1.reset(u0)
    cand = np.zeros((T, 2, 2, Cn))
    cand[:, :, 0] = rng.uniform(-32, 672, (T, 2, Cn))
    cand[:, :, 1] = rng.uniform(-24, 504, (T, 2, Cn))
    cand[:, 0, :, 0] = [388.0, 250.0]
    cand[:, 1, :, 0] = [248.0, 250.0]
    cand[::50, :, 0, 1] = 0.0     # exactly on the image border: the gate is strict (src/pfPose.cpp:251)
    cand[::75, :, 1, 2] = 480.0
    Lv = np.where(rng.random((T, 2, Cn)) < 0.5, 0, rng.integers(1, 129, (T, 2, Cn))).astype(np.uint8)
    Lv[:, :, 0] = 220
    roi = np.tile(np.array([300.0, 51.0, 47.0, 47.0]), (T, 1))
    u_c, u_i, u_p = rng.random((T, 2)), rng.random((T, 2)), rng.random((T, 2))
    mk.associate(b0, b1, cand, Lv, roi, u_c, u_i, u_p)
    res = mk.assoc_results(b0, Cn)
    x, y = cand[:, :, 0], cand[:, :, 1]
    gate = (y > 0) & (y < 480) & (x > 0) & (x < 640) & (Lv != 0)
    assert np.array_equal(res["gate"].astype(bool), gate), "gate decisions must be bit-exact at full size"
    w = res["weights"]
    assert np.all(w[~gate] == 0) and np.allclose(w.sum(2), 1.0, rtol=0, atol=1e-12)
    bins = res["bins"]
    assert np.all(np.diff(bins, axis=2) >= 0) and bins.min() >= 0 and bins.max() < Cn
    assert np.all(np.take_along_axis(gate, bins.astype(np.int64), axis=2)), "only gated candidates are drawn"
    fL, fR = orc.Filter(left_arm.orc, N), orc.Filter(right_arm.orc, N)
    for t in (0, 50, 8191, T - 1):  # oracle spot-checks
        fL.reset(u=u0[t])
        fR.reset(u=u0[t])
        want = orc.associate(fL, fR, cand[t], Lv[t], roi[t], u_c[t])
        assert np.array_equal(res["gate"][t], want["gate"]) and np.array_equal(bins[t], want["bins"])
        assert rel_err_weights(w[t], want["weights"]) <= RTOL


def test_config3_full_size_associate_and_update(left_arm, right_arm):
    """BASELINE config 3 at its real size: 16 384 persons x 500 slots per arm x 17 candidates per hand (1 detection +
    16 clutter), association FOLLOWED by both arm updates (src/pfPose.cpp:238-326), three free-running frames.
    Gate decisions are checked for every person and frame; six persons are replayed through the oracle from reset:
    candidate bins, resampled indices of both arms bit-exact, weights and estimates within 1e-4."""
    torch = pytest.importorskip("torch")
    T, N, Cn, frames = 16384, 500, 17, 3
    dev = torch.device("cuda:0")
    s = torch.cuda.Stream()
    spots = (0, 50, 75, 8191, 12345, T - 1)
    with torch.cuda.stream(s):
        b0 = mk.TrackBatch(left_arm.mk, T, N, stream=s.cuda_stream)
        b1 = mk.TrackBatch(right_arm.mk, T, N, stream=s.cuda_stream)
        rng = np.random.default_rng(33)
        u0 = rng.random(T)
        b0.reset(u0)
        b1.reset(u0)
        fl = {t: orc.Filter(left_arm.orc, N) for t in spots}
        fr_ = {t: orc.Filter(right_arm.orc, N) for t in spots}
        for t in spots:
            fl[t].reset(u=u0[t])
            fr_[t].reset(u=u0[t])
        roi = np.tile(np.array([300.0, 51.0, 47.0, 47.0]), (T, 1))
        par = [torch.empty((T, N), dtype=torch.int32, device=dev) for _ in range(2)]
        wn = [torch.empty((T, N), dtype=torch.float64, device=dev) for _ in range(2)]
        for frame in range(frames):
            cand = np.zeros((T, 2, 2, Cn))
            cand[:, :, 0] = rng.uniform(-32, 672, (T, 2, Cn))
            cand[:, :, 1] = rng.uniform(-24, 504, (T, 2, Cn))
            hx = 388.0 + 60.0 * np.sin(2 * np.pi * (frame / 75.0 + rng.random(T)))  # the detection: near the hand
            hy = 250.0 + 70.0 * np.sin(2 * np.pi * (frame / 50.0 + rng.random(T)))
            cand[:, 0, 0, 0], cand[:, 0, 1, 0] = hx + 3 * rng.standard_normal(T), hy + 3 * rng.standard_normal(T)
            cand[:, 1, 0, 0], cand[:, 1, 1, 0] = hx - 140 + 3 * rng.standard_normal(T), hy + 3 * rng.standard_normal(T)
            cand[::50, :, 0, 1] = 0.0     # exactly on the image border: the gate is strict (src/pfPose.cpp:251)
            cand[::75, :, 1, 2] = 480.0
            Lv = np.where(rng.random((T, 2, Cn)) < 0.5, 0, rng.integers(1, 129, (T, 2, Cn))).astype(np.uint8)
            Lv[:, :, 0] = rng.integers(200, 256, (T, 2))
            u_c, u_i, u_p = rng.random((T, 2)), rng.random((T, 2)), rng.random((T, 2))
            mk.associate(b0, b1, cand, Lv, roi, u_c, u_i, u_p, do_update=True)
            res = mk.assoc_results(b0, Cn)
            x, y = cand[:, :, 0], cand[:, :, 1]
            gate = (y > 0) & (y < 480) & (x > 0) & (x < 640) & (Lv != 0)
            assert np.array_equal(res["gate"].astype(bool), gate), "gate decisions must be bit-exact at full size"
            w = res["weights"]
            assert np.all(w[~gate] == 0) and np.allclose(w.sum(2), 1.0, rtol=0, atol=1e-12)
            bins = res["bins"]
            assert np.all(np.diff(bins, axis=2) >= 0) and bins.min() >= 0 and bins.max() < Cn
            assert np.all(np.take_along_axis(gate, bins.astype(np.int64), axis=2)), "only gated candidates are drawn"
            for arm, b in enumerate((b0, b1)):
                mk._lib.check(mk._lib.lib.mkf_batch_download(b._h, None, None, None, wn[arm].data_ptr(), None,
                                                             par[arm].data_ptr(), None, None, mk.MEM_DEVICE))
            est = [b.estimate() for b in (b0, b1)]
            st = b0.status() | b1.status()
            assert not (st & (L_.ST_POST_DEGENERATE | L_.ST_CHOL_FAIL | L_.ST_CAND_DEGENERATE)).any()
            for arm in range(2):
                p = par[arm].cpu().numpy()
                assert p.min() >= 0 and p.max() < N and np.all(np.diff(p, axis=1) >= 0)
            for t in spots:
                want = orc.associate(fl[t], fr_[t], cand[t], Lv[t], roi[t], u_c[t])
                assert np.array_equal(res["gate"][t], want["gate"]) and np.array_equal(bins[t], want["bins"]), (frame, t)
                assert rel_err_weights(w[t], want["weights"]) <= RTOL
                for arm, f in enumerate((fl[t], fr_[t])):
                    r = f.update(want["meas"][arm], u_i[t, arm], u_p[t, arm])
                    assert np.array_equal(par[arm][t].cpu().numpy(), r["parents"]), (frame, t, arm)
                    assert rel_err_weights(wn[arm][t].cpu().numpy(), r["w_norm"]) <= RTOL
                    xo, po = f.estimate()
                    assert rel_err(est[arm][0][t], xo) <= RTOL and rel_err(est[arm][1][t], po) <= RTOL


def test_legacy_pf2d_full_size():
    """config 4's size for the legacy plain filter (src/pf2D.cpp): 256 filters x 65 536 particles, d = 8, K = 15.
    Size-independent properties on every filter, the oracle on one (65 536 particles take it a second)."""
    rng = np.random.default_rng(4)
    T, N, d, K = 256, 65536, 8, 15
    means = rng.uniform(100, 400, (K, d))
    a = rng.standard_normal((K, d, d))
    covs = 40.0 * (a @ a.transpose(0, 2, 1) + d * np.eye(d))
    wts = rng.dirichlet(np.ones(K))
    pb = mk.Pf2dBatch(T, N, means, covs, wts)
    parts = means[rng.integers(0, K, (T, N))] + rng.standard_normal((T, N, d)) * 6
    pb.set_particles(parts)
    meas = np.stack([parts[:, :, 6].mean(1), parts[:, :, 7].mean(1), parts[:, :, 0].mean(1), parts[:, :, 1].mean(1)],
                    axis=1).reshape(T, 2, 2)
    u = rng.random(T)
    noise = rng.standard_normal((T, N, d))
    pb.update(meas, u, noise)
    p, w, par = pb.get()
    est = pb.estimate()
    assert np.allclose(w.sum(1), 1.0, rtol=0, atol=1e-10)
    check_resample_properties(par, w, N, range(0, T, 37))
    exp = np.take_along_axis(parts, par[:, :, None].astype(np.int64), axis=1)
    exp[:, :, :8] += noise[:, :, :8] * 5.0
    assert np.array_equal(p, exp)                       # particles.row(i) = old.row(parent_i) + N(0,5) (:90-102,:262)
    want = np.einsum("tn,tnd->td", w, p)
    assert np.max(np.abs(est - want) / np.abs(want)) <= 1e-11
    t = 101
    o = orc.Pf2d(N, means, covs, wts)
    o.set_particles(parts[t])
    r = o.update(meas[t], u[t], noise[t])
    assert rel_err_weights(w[t], r["w_norm"]) <= 1e-9    # float expf inside (quirk B12): include/mkf_expf.h on both sides
    assert np.array_equal(par[t], r["parents"])          # 65 536 resampled indices, bit-exact against the oracle
