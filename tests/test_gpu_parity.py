"""GPU parity tests proper: the CUDA path (through the C ABI, libmkf_b200.so) against the CPU oracle on
identical inputs and identical uniform draws.  Tolerances from BASELINE.json:north_star: resampled
indices bit-exact; means, covariances and weights within 1e-4 relative (helpers.RTOL)."""
import os

import numpy as np
import pytest

import mkf_oracle as orc
import mkfbodytracker_pdaf_b200 as mk
from helpers import RTOL, rel_err, rel_err_weights, synth_frame, synth_u_init
from mkfbodytracker_pdaf_b200 import _lib as L

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def oracle_filters(arm, T, N, u_init, chol=orc.CHOL_CV24_LITERAL):
    fs = [orc.Filter(arm.orc, N, chol_mode=chol) for _ in range(T)]
    for f, u in zip(fs, u_init):
        f.reset(u=u)
    return fs


def compare_frame(b, fs, res, check_state=True):
    """compare one frame of GPU batch b with oracle results res (list of dicts)"""
    d = b.download(state=check_state, cov=check_state)
    T = len(fs)
    stats = dict(idx_mismatch=0, ind_mismatch=0, w=0.0, x=0.0, P=0.0)
    for t in range(T):
        stats["ind_mismatch"] += int((d["indicators"][t] != res[t]["indicators"]).sum())
        stats["idx_mismatch"] += int((d["parents"][t] != res[t]["parents"]).sum())
        stats["w"] = max(stats["w"], rel_err_weights(d["w_raw"][t], res[t]["w_raw"]),
                         rel_err_weights(d["w_norm"][t], res[t]["w_norm"]),
                         abs(d["wsum"][t] - res[t]["wsum"]) / res[t]["wsum"])
        if check_state:
            xo, Po = fs[t].get_state()
            stats["x"] = max(stats["x"], rel_err(d["x"][t], xo))
            stats["P"] = max(stats["P"], rel_err(d["P"][t], Po))
    return stats, d


def assert_parity(stats):
    assert stats["ind_mismatch"] == 0, stats
    assert stats["idx_mismatch"] == 0, stats
    assert stats["w"] <= RTOL and stats["x"] <= RTOL and stats["P"] <= RTOL, stats


def test_reset_matches_oracle(left_arm):
    T, N = 7, 500
    u = np.linspace(0.03, 0.97, T)
    b = mk.TrackBatch(left_arm.mk, T, N)
    b.reset(u)
    d = b.download()
    fs = oracle_filters(left_arm, T, N, u)
    for t in range(T):
        xo, Po = fs[t].get_state()
        assert rel_err(d["x"][t], xo) <= 1e-12 and rel_err(d["P"][t], Po) <= 1e-12
        want, _ = orc.resample(left_arm.np.weights, N, u[t])
        assert np.array_equal(d["indicators"][t], want)
        assert np.array_equal(d["parents"][t], np.arange(N))


def test_upload_download_roundtrip(left_arm, rng):
    T, N, d = 3, 70, 12
    x = rng.standard_normal((T, N, d)) * 50
    A = rng.standard_normal((T, N, d, d))
    P = A @ A.transpose(0, 1, 3, 2) * 30 + np.eye(d)
    b = mk.TrackBatch(left_arm.mk, T, N)
    b.upload(x, P)
    out = b.download()
    assert rel_err(out["x"], x) <= 1e-13 and rel_err(out["P"], P) <= 1e-13


@pytest.mark.parametrize("N,shared", [(500, False), (500, True), (15, True), (33, False), (1200, True)])
def test_teacher_forced_frames(left_arm, N, shared):
    """T1: every frame starts from the oracle's state (uploaded), so errors cannot accumulate"""
    T, seed = 5, 0x5EED0002
    tracks = list(range(T))
    u0 = synth_u_init(seed, tracks)
    fs = oracle_filters(left_arm, T, N, u0)
    b = mk.TrackBatch(left_arm.mk, T, N)
    worst = dict(w=0, x=0, P=0)
    for fr in range(6):
        xs, Ps = zip(*[f.get_state() for f in fs])
        b.upload(np.stack(xs), np.stack(Ps))
        meas, ui, up = synth_frame(seed, tracks, fr, None if shared else N)
        res = [fs[t].update(meas[t], ui[t], up[t]) for t in range(T)]
        b.update(meas, ui, up)
        stats, d = compare_frame(b, fs, res)
        assert_parity(stats)
        assert not (d["status"] & (L.ST_POST_DEGENERATE | L.ST_CHOL_FAIL)).any()
        for k in worst:
            worst[k] = max(worst[k], stats[k])
    print(f"teacher-forced N={N} shared={shared}: worst rel err {worst}")
    assert worst["w"] < 1e-9 and worst["x"] < 1e-9 and worst["P"] < 1e-9  # expected ~1e-13; 1e-4 is the contract


def test_config1_free_running_300_frames(left_arm):
    """T2: BASELINE config 1 -- 1 track, N=500, 300 frames, per-slot columns, free-running"""
    seed, N, frames = 0x5EED0001, 500, 300
    u0 = synth_u_init(seed, [0])
    fs = oracle_filters(left_arm, 1, N, u0)
    b = mk.TrackBatch(left_arm.mk, 1, N)
    b.reset(u0)
    gold = np.load(os.path.join(GOLD, "config1_left.npz"))
    mism = 0
    worst = dict(w=0.0, x=0.0, P=0.0, pose=0.0)
    for fr in range(frames):
        meas, ui, up = synth_frame(seed, [0], fr, N, jitter=0)
        res = [fs[0].update(meas[0], ui[0], up[0])]
        b.update(meas, ui, up)
        check_state = fr % 25 == 0 or fr == frames - 1
        stats, d = compare_frame(b, fs, res, check_state)
        mism += stats["idx_mismatch"] + stats["ind_mismatch"]
        for k in ("w", "x", "P"):
            worst[k] = max(worst[k], stats[k])
        _, pose = b.estimate()
        worst["pose"] = max(worst["pose"], float(np.abs(pose[0] - gold["pose"][fr]).max() / np.abs(gold["pose"][fr]).max()))
        if f"parents_{fr}" in gold.files:
            assert np.array_equal(d["parents"][0], gold[f"parents_{fr}"])
            assert np.array_equal(d["indicators"][0], gold[f"indicators_{fr}"])
            assert rel_err_weights(d["w_norm"][0], gold[f"w_norm_{fr}"]) <= RTOL
    print(f"config 1 free-running: index mismatches {mism}, worst rel err {worst}")
    assert mism == 0
    assert max(worst.values()) <= RTOL
    out = b.download()
    assert rel_err(out["x"][0], gold["x_final"]) <= RTOL


def test_config2_small_free_running(left_arm):
    """config 2 at a size the oracle finishes in seconds: 48 tracks x N=500, shared column, free-running"""
    seed, T, N, frames = 0x5EED0002, 48, 500, 12
    tracks = list(range(100, 100 + T))
    u0 = synth_u_init(seed, tracks)
    fs = oracle_filters(left_arm, T, N, u0)
    b = mk.TrackBatch(left_arm.mk, T, N)
    b.reset(u0)
    for fr in range(frames):
        meas, ui, up = synth_frame(seed, tracks, fr)
        res = [fs[t].update(meas[t], ui[t], up[t]) for t in range(T)]
        b.update(meas, ui, up)
        stats, _ = compare_frame(b, fs, res, check_state=(fr == frames - 1))
        assert_parity(stats)
    xb, pose = b.estimate()
    for t in range(T):
        xo, po = fs[t].estimate()
        assert rel_err(xb[t], xo) <= RTOL and rel_err(pose[t], po) <= RTOL


def test_config5_bank_mode_right_arm(right_arm):
    """config 5 shape: data23D model, N=15 slots per track, many tracks, shared column"""
    seed, T, N, frames = 0x5EED0005, 300, 15, 10
    tracks = list(range(T))
    u0 = synth_u_init(seed, tracks)
    fs = oracle_filters(right_arm, T, N, u0)
    b = mk.TrackBatch(right_arm.mk, T, N)
    b.reset(u0)
    for fr in range(frames):
        meas, ui, up = synth_frame(seed, tracks, fr)
        res = [fs[t].update(meas[t], ui[t], up[t]) for t in range(T)]
        b.update(meas, ui, up)
        stats, _ = compare_frame(b, fs, res, check_state=(fr in (0, frames - 1)))
        assert_parity(stats)


def test_config4_large_n_one_track(left_arm):
    """config 4 shape (per-slot columns, N = 16384 on 2 tracks; multi-tile block resampler)"""
    seed, T, N = 0x5EED0004, 2, 16384
    u0 = synth_u_init(seed, [0, 1])
    fs = oracle_filters(left_arm, T, N, u0)
    b = mk.TrackBatch(left_arm.mk, T, N)
    b.reset(u0)
    rng = np.random.default_rng(3)
    for fr in range(3):
        hand = np.array([388.0 + 10 * fr, 250.0])
        meas = np.zeros((T, 6, N))
        meas[:, 0], meas[:, 1], meas[:, 4], meas[:, 5] = 323.5, 74.5, 323.5, 128.55
        meas[:, 2] = hand[0] + 3 * rng.standard_normal((T, N))
        meas[:, 3] = hand[1] + 3 * rng.standard_normal((T, N))
        ui, up = rng.random(T), rng.random(T)
        res = [fs[t].update(meas[t], ui[t], up[t]) for t in range(T)]
        b.update(meas, ui, up)
        stats, _ = compare_frame(b, fs, res, check_state=(fr == 2))
        assert_parity(stats)


@pytest.mark.parametrize("mode", [mk.CHOL_CV3_LITERAL, mk.CHOL_EXACT])
def test_other_chol_modes(left_arm, mode):
    a = left_arm.arrays
    p = mk.default_params()
    p.chol_mode = mode
    m = mk.Model.from_arrays(a["means"], a["covs"], a["weights"], a["gamma"], a["pca_proj"], a["pca_mean"], p)
    T, N, seed = 3, 200, 0x5EED0002
    u0 = synth_u_init(seed, range(T))
    fs = oracle_filters(left_arm, T, N, u0, chol=mode)
    b = mk.TrackBatch(m, T, N)
    b.reset(u0)
    meas, ui, up = synth_frame(seed, range(T), 0)
    res = [fs[t].update(meas[t], ui[t], up[t]) for t in range(T)]
    b.update(meas, ui, up)
    d = b.download()
    for t in range(T):
        if mode == mk.CHOL_CV3_LITERAL:
            # every weight underflows to 0 -> random-index fallback forever
            assert res[t]["status"] & 2 and d["status"][t] & L.ST_POST_DEGENERATE
            assert np.array_equal(d["parents"][t], res[t]["parents"])
        else:
            assert rel_err_weights(d["w_raw"][t], res[t]["w_raw"]) <= RTOL
            assert np.array_equal(d["parents"][t], res[t]["parents"])


def test_resample_kats_on_device(rng):
    N = 500
    out, deg = mk.resample(np.full(N, 1.0 / N), N, 0.5)
    assert deg == 0 and np.array_equal(out, np.arange(N))
    w = np.zeros(N)
    w[123] = 1.0
    out, deg = mk.resample(w, N, 0.25)
    assert deg == 0 and np.all(out == 123)
    for bad in (np.zeros(15), np.full(15, np.nan), np.zeros(300)):
        out, deg = mk.resample(bad, 400, 0.5, seed=77)
        want, wdeg = orc.resample(bad, 400, 0.5, seed=77)
        assert deg == 1 and wdeg == 1 and np.array_equal(out, want)
    out, _ = mk.resample(np.array([0.2, 0.5, 0.3]), 10, -1.0, seed=99)
    want, _ = orc.resample(np.array([0.2, 0.5, 0.3]), 10, -1.0, seed=99)
    assert np.array_equal(out, want)


@pytest.mark.parametrize("L_,N", [(15, 15), (15, 500), (500, 500), (17, 500), (5000, 500), (4096, 4096),
                                  (65536, 65536), (300, 77), (1000, 3000)])
def test_resample_random_weights_bit_exact(rng, L_, N):
    for trial in range(6):
        kind = trial % 3
        if kind == 0:
            w = rng.lognormal(0, 3, L_)
        elif kind == 1:
            w = rng.random(L_) ** 8
            w[rng.integers(0, L_, L_ // 3)] = 0.0  # exact zeros (underflowed weights)
        else:
            w = np.zeros(L_)
            w[rng.integers(0, L_, 3)] = rng.random(3) + 0.1  # collapsed: a few heavy parents
        w = w / w.sum()
        u = [rng.random(), 0.0, 1.0 - 2.0**-53][trial % 3] if trial >= 3 else rng.random()
        out, deg = mk.resample(w, N, u)
        want, wdeg = orc.resample(w, N, u)
        assert deg == wdeg == 0
        assert np.array_equal(out, want), f"L={L_} N={N} trial={trial}: {(out != want).sum()} mismatches"


@pytest.mark.parametrize("L_,N", [(500, 500), (4096, 4096), (65536, 65536), (5000, 500), (700, 1536), (64, 640)])
def test_resample_ties_take_the_literal_path_bit_exact(rng, L_, N):
    """thresholds that coincide with prefix sums (dyadic weights, u = 0 or a dyadic u) cannot be decided by the
    closed form: the track goes through the double-double second opinion and then the warp-run literal loop"""
    cases = []
    cases.append((np.full(L_, 1.0 / L_), 0.0))                       # every C_k is a threshold when N | L or L | N
    w = np.zeros(L_)
    w[:: max(1, L_ // 64)] = 1.0
    cases.append((w / w.sum(), 0.0))
    w = rng.integers(0, 8, L_).astype(np.float64)                     # small integers / power-of-two total: exact sums
    w[0] += 2.0 ** np.ceil(np.log2(w.sum() + 1)) - w.sum()
    cases.append((w / w.sum(), 0.5))
    cases.append((w / w.sum(), 0.0))
    for w, u in cases:
        out, deg = mk.resample(w, N, u)
        want, wdeg = orc.resample(w, N, u)
        assert deg == wdeg == 0
        assert np.array_equal(out, want), f"L={L_} N={N} u={u}: {(out != want).sum()} mismatches"


def test_resample_degenerate_and_tied_weight_vectors(rng):
    """uniform / comb / zero / NaN weight vectors at N = 1024 (block kernel): ties go to the literal loop, zero and
    NaN maxima to the cv::RNG branch"""
    N = 1024
    comb = np.zeros(N)
    comb[::16] = 1.0 / 64
    for w, u in ((np.full(N, 1.0 / N), 0.0), (comb, 0.0), (np.zeros(N), 0.3), (np.full(N, np.nan), 0.3)):
        seed = int(rng.integers(1, 2**62))
        out, deg = mk.resample(w, N, u, seed=seed)
        want, wdeg = orc.resample(w, N, u, seed=seed)
        assert deg == wdeg and np.array_equal(out, want)


def test_full_size_properties_config2(left_arm):
    """config 2 at BASELINE size (4096 x 500): size-independent properties instead of the oracle"""
    torch = pytest.importorskip("torch")
    seed, T, N = 0x5EED0002, 4096, 500
    b = mk.TrackBatch(left_arm.mk, T, N)
    dev = torch.device("cuda:0")
    meas = torch.empty((T, 6), dtype=torch.float64, device=dev)
    ui = torch.empty(T, dtype=torch.float64, device=dev)
    up = torch.empty(T, dtype=torch.float64, device=dev)
    u0 = torch.tensor(synth_u_init(seed, range(64)).repeat(T // 64), dtype=torch.float64, device=dev)
    b.reset(u0)
    for fr in range(4):
        b.synth_fill(seed, 0, fr, 1, mk.MEAS_SHARED, meas, ui, up)
        b.update(meas, ui, up)
    d = b.download(state=True, cov=False)
    # device-generated inputs equal the CPU generator bit for bit
    m_cpu, ui_cpu, up_cpu = synth_frame(seed, range(16), 3)
    assert np.array_equal(meas[:16].cpu().numpy(), m_cpu) and np.array_equal(ui[:16].cpu().numpy(), ui_cpu)
    assert np.array_equal(up[:16].cpu().numpy(), up_cpu)
    par, wn = d["parents"], d["w_norm"]
    assert np.all(np.diff(par, axis=1) >= 0)                       # sorted parents
    assert par.min() >= 0 and par.max() < N
    assert np.allclose(wn.sum(1), 1.0, rtol=0, atol=1e-12)          # normalised
    cnt = np.stack([np.bincount(par[t], minlength=N) for t in range(0, T, 37)])
    assert np.all(np.abs(cnt - N * wn[::37]) < 1 + 1e-9)            # systematic resampling property
    assert np.all(np.diff(d["indicators"], axis=1) >= 0) and d["indicators"].max() < 15
    assert np.isfinite(d["x"]).all() and not d["status"].any()
    # spot-check 3 tracks of the full batch against the oracle (free-running, same draws)
    for t in (0, 2049, 4095):
        f = orc.Filter(left_arm.orc, N)
        f.reset(u=float(u0[t].cpu()))
        for fr in range(4):
            mz, a_, c_ = synth_frame(seed, [t], fr)
            r = f.update(mz[0], a_[0], c_[0])
        assert np.array_equal(par[t], r["parents"])
        xo, _ = f.get_state()
        assert rel_err(d["x"][t], xo) <= RTOL


def literal_model(arm):
    a = arm.arrays
    p = mk.default_params()
    p.alias_mode = L.ALIAS_CV_SHALLOW_LITERAL
    return mk.Model.from_arrays(a["means"], a["covs"], a["weights"], a["gamma"], a["pca_proj"], a["pca_mean"], p)


@pytest.mark.parametrize("N,dyn", [(500, False), (37, False), (500, True), (1100, True)])
def test_literal_alias_mode_matches_oracle_and_reference_sources(left_arm, N, dyn, monkeypatch):
    """alias_mode = CV_SHALLOW_LITERAL (quirk B3): the GPU must equal the oracle in literal mode and -- through
    oracle/_ref -- the reference's own pf2DRao.cpp, free-running, indices bit-exact.  dyn: the dynamic chain walker
    (MKF_ALIAS_DYN=1, k_alias_runs + k_slot_update_chain_dyn), an A/B experiment that must give the same results."""
    import mkf_ref
    if dyn:
        monkeypatch.setenv("MKF_ALIAS_DYN", "1")
    else:
        monkeypatch.delenv("MKF_ALIAS_DYN", raising=False)
    have_ref = mkf_ref.available()
    T, frames, seed = 3, 10, 0x5EED0001
    rng = np.random.default_rng(17)
    tick0 = [int(x) for x in rng.integers(1, 2**62, T)]
    u0 = np.array([mkf_ref.tick_to_u(t, 15) for t in tick0])
    fs = [orc.Filter(left_arm.orc, N, alias_mode=orc.ALIAS_CV_SHALLOW_LITERAL) for _ in range(T)]
    fi = orc.Filter(left_arm.orc, N, alias_mode=orc.ALIAS_INDEPENDENT)
    for f, u in zip(fs + [fi], list(u0) + [u0[0]]):
        f.reset(u=u)
    rf = None
    if have_ref:
        a = left_arm.arrays
        rf = mkf_ref.RefFilter({k: a[k] for k in ("means", "covs", "weights", "gamma", "pca_proj", "pca_mean")}, N)
        rf.reset(tick0[0])
    b = mk.TrackBatch(literal_model(left_arm), T, N)
    b.reset(u0)
    differs_from_independent = False
    for fr in range(frames):
        meas, _, _ = synth_frame(seed, range(T), fr, N, jitter=0)
        ticks = rng.integers(1, 2**62, (T, 2))
        ui = np.array([mkf_ref.tick_to_u(t, 15) for t in ticks[:, 0]])
        up = np.array([mkf_ref.tick_to_u(t, N) for t in ticks[:, 1]])
        res = [fs[t].update(meas[t], ui[t], up[t]) for t in range(T)]
        fi.update(meas[0], ui[0], up[0])
        b.update(meas, ui, up)
        stats, d = compare_frame(b, fs, res, check_state=True)
        assert_parity(stats)
        assert stats["x"] < 1e-9 and stats["P"] < 1e-9 and stats["w"] < 1e-9
        if rf is not None:
            rf.update(meas[0], int(ticks[0, 0]), int(ticks[0, 1]))
            xr, Pr = rf.get_state()
            assert rel_err(d["x"][0], xr) <= RTOL and rel_err(d["P"][0], Pr) <= RTOL
            xb, _ = b.estimate()
            assert rel_err(xb[0], rf.estimate()) <= RTOL
        xi, _ = fi.get_state()
        if fr >= 1 and rel_err(d["x"][0], xi) > 1e-3:
            differs_from_independent = True
    assert differs_from_independent, "literal aliasing must change the result from frame 2 on"


@pytest.mark.parametrize("alias,N", [(0, 120), (1, 120), (0, 15)])  # N = 15: the cv::RNG branch inside k_frame_small
def test_degenerate_frame_then_recovery(left_arm, alias, N):
    """all weights underflow (measurement far away) -> cv::RNG random-index fallback (unsorted parents, seeded);
    the following frames must still match the oracle in both alias modes"""
    T = 4
    a = left_arm.arrays
    p = mk.default_params()
    p.alias_mode = alias
    m = mk.Model.from_arrays(a["means"], a["covs"], a["weights"], a["gamma"], a["pca_proj"], a["pca_mean"], p)
    u0 = np.array([0.11, 0.37, 0.62, 0.93])
    fs = [orc.Filter(left_arm.orc, N, alias_mode=alias) for _ in range(T)]
    for f, u in zip(fs, u0):
        f.reset(u=u)
    b = mk.TrackBatch(m, T, N)
    b.reset(u0)
    rng = np.random.default_rng(23)
    for fr in range(5):
        meas, ui, up = synth_frame(0x5EED0002, range(T), fr, N)
        if fr == 1:
            meas[1:3, 2:4, :] = 1e5  # tracks 1 and 2 lose the hand completely on this frame
        seeds = rng.integers(1, 2**62, (T, 2)).astype(np.uint64)
        res = [fs[t].update(meas[t], ui[t], up[t], seed_ind=int(seeds[t, 0]), seed_post=int(seeds[t, 1]))
               for t in range(T)]
        b.update(meas, ui, up, seeds=seeds)
        d = b.download()
        for t in range(T):
            assert np.array_equal(d["parents"][t], res[t]["parents"]), (fr, t)
            deg = bool(res[t]["status"] & 2)
            assert deg == bool(d["status"][t] & L.ST_POST_DEGENERATE)
            if fr == 1 and t in (1, 2):
                assert deg
            if t in (0, 3) or fr == 0:
                assert not deg
            if deg:
                assert res[t]["wsum"] == 0 and d["wsum"][t] == 0 and np.isnan(d["w_norm"][t]).all()
            else:
                assert rel_err_weights(d["w_norm"][t], res[t]["w_norm"]) <= RTOL
            xo, Po = fs[t].get_state()
            assert rel_err(d["x"][t], xo) <= RTOL and rel_err(d["P"][t], Po) <= RTOL


@pytest.mark.parametrize("N", [15, 64, 200])  # 15: k_frame_small + the repair kernel's serial tail; 200: the two-launch
                                              # record-sharing path (k_slot_update_heads_direct)
def test_cholesky_failure_branch_matches_oracle(left_arm, rng, N):
    """S not positive definite: cv::Cholesky fails, chol() returns the partially factored clone and the
    reference carries on with LU inverses (src/pf2DRao.cpp:37,52; src/KF_model.cpp:21)"""
    T = 2
    fs = oracle_filters(left_arm, T, N, [0.3, 0.8])
    H = left_arm.np.H
    xs, Ps = [], []
    for f in fs:
        x, P = f.get_state()
        for j in range(0, N, 3):  # every third slot gets an indefinite innovation covariance
            scale = [200.0, 3000.0, 40000.0][(j // 3) % 3]
            P[j] = P[j] - scale * (H.T @ np.diag(rng.uniform(0.5, 1.5, 6)) @ H)
        f.set_state(x, P)
        xs.append(x)
        Ps.append(P)
    b = mk.TrackBatch(left_arm.mk, T, N)
    b.upload(np.stack(xs), np.stack(Ps))
    meas, ui, up = synth_frame(0x5EED0002, range(T), 0, N)
    res = [fs[t].update(meas[t], ui[t], up[t]) for t in range(T)]
    b.update(meas, ui, up)
    d = b.download()
    n_fail = 0
    for t in range(T):
        assert res[t]["status"] & 4 and d["status"][t] & L.ST_CHOL_FAIL
        wo, wg = res[t]["w_raw"], d["w_raw"][t]
        assert np.array_equal(np.isnan(wo), np.isnan(wg))
        fin = np.isfinite(wo) & (wo > 1e-290) & (wo < 1e290)
        assert np.max(np.abs(wg[fin] - wo[fin]) / wo[fin]) <= 1e-6  # ill-conditioned by construction
        n_fail += int((~fin).sum())
        # NaN weights make wsum NaN -> random-index fallback (seed 1 on both sides): still bit-exact
        assert np.array_equal(d["parents"][t], res[t]["parents"])
        xo, Po = fs[t].get_state()
        assert np.array_equal(np.isnan(xo), np.isnan(d["x"][t]))
        ok = np.isfinite(Po).all(axis=(1, 2)) & np.isfinite(xo).all(axis=1)
        assert ok.sum() > N // 2
        assert rel_err(d["x"][t][ok], xo[ok]) <= RTOL and rel_err(d["P"][t][ok], Po[ok]) <= RTOL
    print("cholesky-failure slots with non-finite or extreme weights:", n_fail)
    # a second frame on top: the repaired tracks stored one record per slot, their parents are the unsorted random
    # indices of the NaN-weight fallback, and record sharing has to cope with both
    meas, ui, up = synth_frame(0x5EED0002, range(T), 1, N)
    res = [fs[t].update(meas[t], ui[t], up[t]) for t in range(T)]
    b.update(meas, ui, up)
    d = b.download()
    for t in range(T):
        wo, wg = res[t]["w_raw"], d["w_raw"][t]
        assert np.array_equal(np.isnan(wo), np.isnan(wg))
        assert np.array_equal(d["parents"][t], res[t]["parents"])
        xo, Po = fs[t].get_state()
        ok = np.isfinite(Po).all(axis=(1, 2)) & np.isfinite(xo).all(axis=1) & np.isfinite(d["x"][t]).all(axis=1)
        assert np.array_equal(np.isfinite(xo).all(axis=1), np.isfinite(d["x"][t]).all(axis=1))
        if ok.any():
            assert rel_err(d["x"][t][ok], xo[ok]) <= RTOL and rel_err(d["P"][t][ok], Po[ok]) <= RTOL


@pytest.mark.parametrize("K,d", [(25, 10), (35, 10), (26, 12)])
def test_other_model_shapes(K, d):
    """the file-name pattern data?3D_PCA_<samples>_<K>_<d>.yml implies other shapes (launch files use
    25_10, 26_10, 35_10; bodyTrackingBagCompare.launch:2-3, launch/ChaLearn.launch:7-8): synthetic models"""
    rng = np.random.default_rng(K * 100 + d)
    D = 22
    q, _ = np.linalg.qr(rng.standard_normal((D, D)))
    proj = q[:d].astype(np.float32).astype(np.float64)
    pmean = np.concatenate([[388, 281, 1.7, 390, 223, 1.8, 369, 158, 1.8, 324, 74.5, 1.8, 326, 128.6, 1.85],
                            np.zeros(D - 15)])[:D]
    means = rng.standard_normal((K, d)) * 40
    covs = np.zeros((K, d, d))
    for k in range(K):
        a_ = rng.standard_normal((d, d))
        covs[k] = 150.0 * (a_ @ a_.T / d + 0.05 * np.eye(d))
    wts = rng.dirichlet(np.ones(K))
    gam = rng.uniform(0.85, 0.99, K)
    m = mk.Model.from_arrays(means, covs, wts, gam, proj, pmean)
    om = orc.Model(means, covs, wts, gam, proj, pmean)
    T, N = 4, 300
    u0 = rng.random(T)
    fs = [orc.Filter(om, N) for _ in range(T)]
    for f, u in zip(fs, u0):
        f.reset(u=u)
    b = mk.TrackBatch(m, T, N)
    b.reset(u0)
    for fr in range(5):
        meas, ui, up = synth_frame(0x5EED0002, range(T), fr, N)
        res = [fs[t].update(meas[t], ui[t], up[t]) for t in range(T)]
        b.update(meas, ui, up)
        stats, _ = compare_frame(b, fs, res, check_state=True)
        assert_parity(stats)
    xb, pose = b.estimate()
    for t in range(T):
        xo, po = fs[t].estimate()
        assert rel_err(xb[t], xo) <= RTOL and rel_err(pose[t], po) <= RTOL



def test_indicator_bounds_all_shapes_and_edge_draws(left_arm):
    """K -> N indicator resample through reset(): every K group size, u at the edges, bit-exact vs the loop"""
    rng = np.random.default_rng(5)
    a = left_arm.arrays
    for K in (1, 2, 15, 16, 17, 33, 64):
        means = rng.standard_normal((K, 12)) * 30
        covs = np.tile(np.eye(12) * 50.0, (K, 1, 1))
        wts = rng.dirichlet(np.ones(K) * 0.3)
        m = mk.Model.from_arrays(means, covs, wts, np.full(K, 0.9), a["pca_proj"], a["pca_mean"])
        for N in (7, 500, 4099):
            u = np.concatenate([[0.0, 1.0 - 2.0**-53, 0.5], rng.random(13)])
            b = mk.TrackBatch(m, len(u), N)
            b.reset(u)
            ind = b.download(state=False, cov=False)["indicators"]
            for t in range(len(u)):
                want, _ = orc.resample(wts, N, u[t])
                assert np.array_equal(ind[t], want), (K, N, t)


@pytest.mark.parametrize("N", [10, 500, 130])
def test_indicator_draw_with_unnormalised_prior_wraps_like_the_loop(left_arm, N):
    """Prior weights that do not sum to 1 (my_gmm::loadGaussian takes any w, src/my_gmm.cpp:45-52): the literal loop of
    src/pf2DRao.cpp:198-207 walks past the last component and starts over at component 0 -- possibly several times --
    so the indicators are no longer monotone in the slot index.  ADVICE r1: w = [.2, .2, .1], N = 10, u = .5 must give
    0,0,1,1,2,0,0,1,1,2.  Checked through reset() and through full frame updates (shared column -> record sharing,
    per-slot columns -> k_slot_update) against the oracle."""
    rng = np.random.default_rng(17)
    a = left_arm.arrays
    for wts in (np.array([0.2, 0.2, 0.1]), np.array([0.05, 0.3, 0.02, 0.11]), np.array([0.7, 0.2999])):
        K = len(wts)
        means = a["means"][:K]
        covs = a["covs"][:K]
        gam = a["gamma"][:K]
        m = mk.Model.from_arrays(means, covs, wts, gam, a["pca_proj"], a["pca_mean"])
        om = orc.Model(means, covs, wts, gam, a["pca_proj"], a["pca_mean"])
        u = np.concatenate([[0.5, 0.0, 1.0 - 2.0**-53], rng.random(5)])
        T = len(u)
        b = mk.TrackBatch(m, T, N)
        b.reset(u)
        ind = b.download(state=False, cov=False)["indicators"]
        for t in range(T):
            want, _ = orc.resample(wts, N, u[t])
            assert np.array_equal(ind[t], want), (wts, N, t)
        if K == 3 and N == 10:
            assert list(ind[0]) == [0, 0, 1, 1, 2, 0, 0, 1, 1, 2]
        fs = [orc.Filter(om, N) for _ in range(T)]
        for f, uu in zip(fs, u):
            f.reset(u=uu)
        for fr in range(4):
            meas, ui, up = synth_frame(0x5EED0002, list(range(T)), fr, N if fr % 2 else None)
            res = [f.update(meas[t], ui[t], up[t]) for t, f in enumerate(fs)]
            b.update(meas, ui, up)
            stats, d = compare_frame(b, fs, res)
            assert_parity(stats)
        if wts.sum() < 0.99:
            assert (b.status() & L.ST_IND_WRAP).all()


def test_host_async_pipeline_matches_synchronous_calls(left_arm):
    """MKF_MEM_HOST_ASYNC (inputs copied on the library's copy stream, results returned on its output stream, no
    synchronisation until mkf_batch_sync) gives bit-identical poses to the synchronous host path, frame by frame,
    with many frames in flight"""
    torch = pytest.importorskip("torch")
    seed, T, N, frames = 0x5EED0002, 96, 500, 12
    tracks = list(range(T))
    u0 = synth_u_init(seed, tracks)
    ins = [synth_frame(seed, tracks, f) for f in range(frames)]
    b = mk.TrackBatch(left_arm.mk, T, N)
    b.reset(u0)
    want = []
    for m, ui, up in ins:
        b.update(m, ui, up)
        want.append(b.estimate()[1].copy())
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_in = [(pin(m), pin(ui), pin(up)) for m, ui, up in ins]
    h_pose = [torch.zeros((T, left_arm.mk.D), dtype=torch.float64).pin_memory() for _ in range(frames)]
    b.reset(u0)
    for f in range(frames):
        m, ui, up = h_in[f]
        b.update(m, ui, up, mem=mk.MEM_HOST_ASYNC)
        b.estimate_into(None, h_pose[f], mem=mk.MEM_HOST_ASYNC)
    b.sync()
    for f in range(frames):
        assert np.array_equal(h_pose[f].numpy(), want[f]), f"frame {f}"
    # and the synchronous path still works on the same batch afterwards
    b.update(*ins[0])
    assert np.isfinite(b.estimate()[1]).all()


def test_record_sharing_is_invisible(left_arm):
    """shared-measurement frames store identical children once (mkf_batch_shared_records < slots) while every per-slot
    output stays that of N independent slots: compared with the oracle slot by slot, then a per-slot-measurement frame
    (no sharing) on top of a shared one, then download / upload round trip of a shared state"""
    seed, T, N = 0x5EED0002, 5, 500
    tracks = list(range(T))
    u0 = synth_u_init(seed, tracks)
    fs = [orc.Filter(left_arm.orc, N) for _ in tracks]
    for t in tracks:
        fs[t].reset(u=u0[t])
    b = mk.TrackBatch(left_arm.mk, T, N)
    b.reset(u0)
    assert b.shared_records() == (T * N, T * N)
    for fr in range(5):
        m, ui, up = synth_frame(seed, tracks, fr)
        per_slot = fr == 3
        if per_slot:  # every slot its own column: nothing to share
            mm = np.repeat(m[:, :, None], N, axis=2) + np.random.default_rng(fr).normal(0, 2.0, (T, 6, N))
            b.update(mm, ui, up)
        else:
            b.update(m, ui, up)
        rec, slots = b.shared_records()
        # frame 0 starts from one record per slot (reset), so sharing shows from the second shared frame on
        assert slots == T * N and (rec == slots if (per_slot or fr == 0) else T <= rec < 0.9 * slots), (fr, rec, slots)
        d = b.download()
        for t in tracks:
            r = fs[t].update(mm[t] if per_slot else m[t], ui[t], up[t])
            assert np.array_equal(d["parents"][t], r["parents"]) and np.array_equal(d["indicators"][t], r["indicators"])
            assert rel_err_weights(d["w_norm"][t], r["w_norm"]) <= 1e-9
            xo, Po = fs[t].get_state()
            assert rel_err(d["x"][t], xo) <= 1e-9 and rel_err(d["P"][t], Po) <= 1e-9
    # a shared state survives download -> upload (upload stores one record per slot again)
    d = b.download()
    b2 = mk.TrackBatch(left_arm.mk, T, N)
    b2.reset(u0)
    b2.upload(d["x"], d["P"])
    d2 = b2.download()
    assert rel_err(d2["x"], d["x"]) <= 1e-12 and rel_err(d2["P"], d["P"]) <= 1e-12


@pytest.mark.parametrize("T,N", [(7, 333), (3, 1025), (70, 65), (2, 4099), (1, 61), (130, 500)])
def test_record_sharing_variants_agree(left_arm, T, N, monkeypatch):
    """the slot-update paths of a shared-measurement frame -- every slot computed (MKF_DEDUP=0), the single-launch
    record-sharing kernel (MKF_SHARE_SPLIT=0), the two-launch per-slot one (MKF_RUNS=0: k_share_keys +
    k_slot_update_heads_direct, weights per record read by the resampler) and the default run-length pipeline
    (mkf_runs.cuh: k_frame_heads + k_slot_update_heads_direct + k_resample_runs, per-slot views replayed on demand) --
    run the same arithmetic on the same Gaussians: every download agrees bit for bit over free-running frames, on shapes
    that leave ragged tails (slots not a multiple of 4 / 1024, tracks straddling chunks, a chunk with more than 128
    heads, N <= 64 falling back to the per-slot resampler).  Estimates: bit-identical among the per-slot paths; the
    run-length path sums multiplicity x mean (one rounding instead of m), so it agrees to 1e-13."""
    seed = 0x5EED0007
    tracks = list(range(T))
    u0 = synth_u_init(seed, tracks)

    def make(env):
        for k in ("MKF_DEDUP", "MKF_SHARE_SPLIT", "MKF_RUNS", "MKF_FUSED"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        b = mk.TrackBatch(left_arm.mk, T, N)  # the variant is fixed at creation
        b.reset(u0)
        return b

    variants = [make({"MKF_DEDUP": "0"}), make({"MKF_SHARE_SPLIT": "0"}), make({"MKF_RUNS": "0"}), make({})]
    keys = ("parents", "indicators", "x", "P", "w_raw", "w_norm", "wsum", "status")
    # the run-length frame as ONE launch (k_frame_fused, an experiment kept for A/B runs): MKF_FUSED is read per frame
    monkeypatch.setenv("MKF_FUSED", "1")
    fused = mk.TrackBatch(left_arm.mk, T, N)
    fused.reset(u0)
    for fr in range(3):
        m, ui, up = synth_frame(seed, tracks, fr)
        fused.update(m, ui, up)
    monkeypatch.delenv("MKF_FUSED")
    ref3 = mk.TrackBatch(left_arm.mk, T, N)
    ref3.reset(u0)
    for fr in range(3):
        m, ui, up = synth_frame(seed, tracks, fr)
        ref3.update(m, ui, up)
    d_f, d_r = fused.download(), ref3.download()
    for key in keys:
        assert np.array_equal(d_f[key], d_r[key]), ("fused", key)
    assert np.array_equal(fused.estimate()[1], ref3.estimate()[1])

    def check_estimates():
        e0 = variants[0].estimate()
        for b in variants[1:3]:
            e = b.estimate()
            assert np.array_equal(e0[0], e[0]) and np.array_equal(e0[1], e[1])
        e = variants[3].estimate()
        assert rel_err(e[0], e0[0]) <= 1e-13 and rel_err(e[1], e0[1]) <= 1e-13

    for fr in range(6):
        m, ui, up = synth_frame(seed, tracks, fr)
        ds = []
        for b in variants:
            b.update(m, ui, up)
            ds.append(b.download())
        for d in ds[1:]:
            for key in keys:
                assert np.array_equal(ds[0][key], d[key]), (fr, key)
        if fr >= 2:  # sharing is real on the sharing variants, absent on the first
            assert variants[0].shared_records()[0] == T * N
            assert variants[1].shared_records() == variants[2].shared_records()
            # the per-slot paths compute a group that straddles a 1024-slot chunk twice; the run list does not
            assert variants[3].shared_records()[0] <= variants[2].shared_records()[0]
            if N >= 4 * 15:
                assert variants[3].shared_records()[0] < T * N
        # estimates read the state through the record indices / the run list
        check_estimates()
    # frames with nothing read back in between: the run-length path never materialises a per-slot array here
    for fr in range(6, 14):
        m, ui, up = synth_frame(seed, tracks, fr)
        for b in variants:
            b.update(m, ui, up)
        check_estimates()
    ds = [b.download() for b in variants]
    for d in ds[1:]:
        for key in keys:
            assert np.array_equal(ds[0][key], d[key]), ("after 8 unobserved frames", key)
    # a per-slot-measurement frame on top (the run-length path hands its set over to the per-slot kernels), then back
    mm = np.repeat(m[:, :, None], N, axis=2) + np.random.default_rng(3).normal(0, 2.0, (T, 6, N))
    for fr, meas in ((14, mm), (15, None), (16, None)):
        m, ui, up = synth_frame(seed, tracks, fr)
        for b in variants:
            b.update(m if meas is None else meas, ui, up)
    ds = [b.download() for b in variants]
    for d in ds[1:]:
        for key in keys:
            assert np.array_equal(ds[0][key], d[key]), ("after the per-slot frame", key)


@pytest.mark.parametrize("T,N,arm", [(301, 15, "right"), (9, 16, "left"), (70, 9, "left"), (1, 12, "left"), (1027, 15, "left")])
def test_short_track_frame_variants_agree(left_arm, right_arm, T, N, arm, monkeypatch):
    """Short tracks (9 <= N <= 16; BASELINE config 5 / bank mode): the whole frame in ONE launch (k_frame_small: a half
    warp per track draws the indicators, updates the slots, resamples and estimates, mkf_frame_small.cuh) against the
    five-launch per-slot frame (MKF_SMALL_FUSED=0: k_indicator_bounds, k_slot_update, k_slot_update_repair,
    k_resample_small, k_estimate_small).  Same arithmetic in the same order, so EVERY download and the estimates agree
    bit for bit over free-running frames with shared and per-slot measurement columns, on track counts that leave
    half-empty warps and CTAs."""
    model = (right_arm if arm == "right" else left_arm).mk
    seed = 0x5EED0015
    tracks = list(range(T))
    u0 = synth_u_init(seed, tracks)

    def make(env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        b = mk.TrackBatch(model, T, N)  # the variant is fixed at creation
        b.reset(u0)
        return b

    ref, fus = make({"MKF_SMALL_FUSED": "0"}), make({"MKF_SMALL_FUSED": "1"})
    keys = ("parents", "indicators", "x", "P", "w_raw", "w_norm", "wsum", "status")
    for fr in range(10):
        m, ui, up = synth_frame(seed, tracks, fr, N if fr % 3 == 2 else None)
        ref.update(m, ui, up)
        fus.update(m, ui, up)
        if fr % 2 == 0 or fr == 9:
            er, ef = ref.estimate(), fus.estimate()
            assert np.array_equal(er[0], ef[0]) and np.array_equal(er[1], ef[1]), fr
        if fr in (0, 3, 4, 9):
            dr, df = ref.download(), fus.download()
            for key in keys:
                assert np.array_equal(dr[key], df[key]), (fr, key)
    # the output back-end reads the pose the frame left in the batch
    assert np.array_equal(ref.pose3d(), fus.pose3d())


def test_short_track_frame_is_two_launches_by_default(left_arm, monkeypatch):
    """A batch of short tracks takes k_frame_small by default (<= 16384 tracks): one frame = k_frame_small + the repair
    kernel, and the estimate is a copy out of the batch -- against four launches + the estimate kernel when forced off."""
    T, N, seed = 64, 15, 0x5EED0015
    tracks = list(range(T))
    u0 = synth_u_init(seed, tracks)
    counts = {}
    for label, env in (("default", None), ("off", "0")):
        if env is None:
            monkeypatch.delenv("MKF_SMALL_FUSED", raising=False)
        else:
            monkeypatch.setenv("MKF_SMALL_FUSED", env)
        b = mk.TrackBatch(left_arm.mk, T, N)
        b.reset(u0)
        m, ui, up = synth_frame(seed, tracks, 0)
        b.update(m, ui, up)  # (first call: function attributes, buffers)
        b.estimate()
        n0 = mk.launch_count()
        m, ui, up = synth_frame(seed, tracks, 1)
        b.update(m, ui, up)
        b.estimate()
        counts[label] = mk.launch_count() - n0
    assert counts["default"] == 2, counts
    assert counts["off"] >= 5, counts  # bounds, slot update, repair, resample + the estimate kernel


@pytest.mark.parametrize("K,d,N", [(25, 10, 14), (26, 12, 16), (31, 10, 9)])
def test_short_track_frame_other_model_shapes(K, d, N, monkeypatch):
    """k_frame_small with more components than lanes in a half warp (16 < K <= 32: two indicator boundaries per lane) and
    with d = 10: against the oracle and, bit for bit, against the five-launch frame"""
    rng = np.random.default_rng(K * 100 + d)
    D = 22
    q, _ = np.linalg.qr(rng.standard_normal((D, D)))
    proj = q[:d].astype(np.float32).astype(np.float64)
    pmean = np.concatenate([[388, 281, 1.7, 390, 223, 1.8, 369, 158, 1.8, 324, 74.5, 1.8, 326, 128.6, 1.85],
                            np.zeros(D - 15)])[:D]
    means = rng.standard_normal((K, d)) * 40
    covs = np.zeros((K, d, d))
    for k in range(K):
        a_ = rng.standard_normal((d, d))
        covs[k] = 150.0 * (a_ @ a_.T / d + 0.05 * np.eye(d))
    wts = rng.dirichlet(np.ones(K))
    gam = rng.uniform(0.85, 0.99, K)
    m = mk.Model.from_arrays(means, covs, wts, gam, proj, pmean)
    om = orc.Model(means, covs, wts, gam, proj, pmean)
    T = 37
    u0 = rng.random(T)
    fs = [orc.Filter(om, N) for _ in range(T)]
    for f, u in zip(fs, u0):
        f.reset(u=u)
    monkeypatch.setenv("MKF_SMALL_FUSED", "0")
    ref = mk.TrackBatch(m, T, N)
    monkeypatch.setenv("MKF_SMALL_FUSED", "1")
    fus = mk.TrackBatch(m, T, N)
    ref.reset(u0)
    fus.reset(u0)
    keys = ("parents", "indicators", "x", "P", "w_raw", "w_norm", "wsum", "status")
    for fr in range(6):
        meas, ui, up = synth_frame(0x5EED0002, range(T), fr, N if fr % 2 else None)
        res = [fs[t].update(meas[t], ui[t], up[t]) for t in range(T)]
        ref.update(meas, ui, up)
        fus.update(meas, ui, up)
        stats, _ = compare_frame(fus, fs, res, check_state=True)
        assert_parity(stats)
        dr, df = ref.download(), fus.download()
        for key in keys:
            assert np.array_equal(dr[key], df[key]), (fr, key)
        er, ef = ref.estimate(), fus.estimate()
        assert np.array_equal(er[0], ef[0]) and np.array_equal(er[1], ef[1]), fr
    xb, pose = fus.estimate()
    for t in range(T):
        xo, po = fs[t].estimate()
        assert rel_err(xb[t], xo) <= RTOL and rel_err(pose[t], po) <= RTOL
