"""GPU parity of the association ("PDAF") step and of the legacy pf2D filter against the oracle."""
import numpy as np
import pytest

import mkf_oracle as orc
import mkfbodytracker_pdaf_b200 as mk
from helpers import RTOL, rel_err, rel_err_weights, synth_frame, synth_u_init
from mkfbodytracker_pdaf_b200 import _lib as L

pytestmark = pytest.mark.gpu


def make_candidates(seed, tracks, frame, Cn):
    T = len(tracks)
    cand = np.zeros((T, 2, 2, Cn))
    Lv = np.zeros((T, 2, Cn), np.uint8)
    for i, t in enumerate(tracks):
        for h in range(2):
            for c in range(Cn):
                cand[i, h, 0, c], cand[i, h, 1, c], Lv[i, h, c] = orc.synth_candidate(seed, t, frame, h, Cn, c)
    return cand, Lv


@pytest.mark.parametrize("N,Cn", [(500, 17), (100, 5000), (15, 17)])
def test_association_config3_small(left_arm, right_arm, N, Cn):
    """config 3 at oracle-sized T: gate bits and bins bit-exact, weights / states within 1e-4"""
    seed, T, frames = 0x5EED0003, 6, 4
    if Cn > 1000:
        T, frames = 2, 2
    tracks = list(range(T))
    u0 = synth_u_init(seed, tracks)
    fl = [orc.Filter(left_arm.orc, N) for _ in tracks]
    fr_ = [orc.Filter(right_arm.orc, N) for _ in tracks]
    for t in tracks:
        fl[t].reset(u=u0[t])
        fr_[t].reset(u=u0[t])
    b0 = mk.TrackBatch(left_arm.mk, T, N)
    b1 = mk.TrackBatch(right_arm.mk, T, N, stream=None)
    # both arm batches must share a stream: create the second on the first's stream via torch-free path
    b1.close()
    torch = pytest.importorskip("torch")
    s = torch.cuda.Stream()
    b0 = mk.TrackBatch(left_arm.mk, T, N, stream=s.cuda_stream)
    b1 = mk.TrackBatch(right_arm.mk, T, N, stream=s.cuda_stream)
    b0.reset(u0)
    b1.reset(u0)
    roi = np.tile(np.array([300.0, 51.0, 47.0, 47.0]), (T, 1))
    rng = np.random.default_rng(5)
    for frame in range(frames):
        cand, Lv = make_candidates(seed, tracks, frame, Cn)
        u_cand = rng.random((T, 2))
        u_ind = rng.random((T, 2))
        u_post = rng.random((T, 2))
        mk.associate(b0, b1, cand, Lv, roi, u_cand, u_ind, u_post)
        res = mk.assoc_results(b0, Cn)
        d0, d1 = b0.download(), b1.download()
        for t in tracks:
            want = orc.associate(fl[t], fr_[t], cand[t], Lv[t], roi[t], u_cand[t])
            assert np.array_equal(res["gate"][t], want["gate"]), "gate decisions must be bit-exact"
            assert rel_err_weights(res["weights"][t], want["weights"]) <= RTOL
            assert np.array_equal(res["bins"][t], want["bins"]), "candidate bins must be bit-exact"
            for arm, (f, d) in enumerate(((fl[t], d0), (fr_[t], d1))):
                r = f.update(want["meas"][arm], u_ind[t, arm], u_post[t, arm])
                assert np.array_equal(d["parents"][t], r["parents"])
                assert rel_err_weights(d["w_norm"][t], r["w_norm"]) <= RTOL
                xo, Po = f.get_state()
                assert rel_err(d["x"][t], xo) <= RTOL and rel_err(d["P"][t], Po) <= RTOL
    assert not (d0["status"] | d1["status"]).any()


@pytest.mark.parametrize("N", [64, 200])  # 200: the warp-per-track candidate resampler (k_resample_warp)
def test_association_no_candidate_passes_gate(left_arm, right_arm, N):
    """quirk B11: all weights NaN -> random candidates from cv::RNG (seeded)"""
    torch = pytest.importorskip("torch")
    T, Cn = 3, 9
    s = torch.cuda.Stream()
    b0 = mk.TrackBatch(left_arm.mk, T, N, stream=s.cuda_stream)
    b1 = mk.TrackBatch(right_arm.mk, T, N, stream=s.cuda_stream)
    u0 = np.array([0.1, 0.5, 0.9])
    b0.reset(u0)
    b1.reset(u0)
    cand = np.random.default_rng(1).uniform(10, 400, (T, 2, 2, Cn))
    Lv = np.zeros((T, 2, Cn), np.uint8)
    Lv[1, 1, 3] = 200  # one hand of one person does have a valid candidate
    roi = np.tile(np.array([300.0, 51.0, 47.0, 47.0]), (T, 1))
    seeds = np.arange(1, T * 6 + 1, dtype=np.uint64).reshape(T, 2, 3)
    mk.associate(b0, b1, cand, Lv, roi, np.full((T, 2), 0.5), None, None, seeds=seeds, do_update=False)
    res = mk.assoc_results(b0, Cn)
    for t in range(T):
        for h in range(2):
            if (t, h) == (1, 1):
                assert np.all(res["bins"][t, h] == 3) and res["gate"][t, h].sum() == 1
                continue
            assert not res["gate"][t, h].any() and np.isnan(res["weights"][t, h]).all()
            want = orc.cvrng(int(seeds[t, h, 0]), Cn, N + 1, 0)[0][1:]
            assert np.array_equal(res["bins"][t, h], want)
    st0, st1 = b0.status(), b1.status()
    assert (st0 & L.ST_CAND_DEGENERATE).all() and (st1[[0, 2]] & L.ST_CAND_DEGENERATE).all()
    assert not (st1[1] & L.ST_CAND_DEGENERATE)


@pytest.mark.parametrize("N,Cn", [(500, 5000), (64, 640), (340, 17)])  # (340, 17): k_resample_warp, 340 = 20 x 17
def test_association_tied_candidate_weights_inside_a_batch(left_arm, right_arm, N, Cn):
    """persons whose candidates are all identical (equal weights) with u = 0 put every threshold on a prefix sum: their
    C -> N candidate resample must leave the closed form (status CAND_FALLBACK) inside a batch of ordinary persons and
    reproduce the literal loop bit for bit"""
    torch = pytest.importorskip("torch")
    seed, T = 0x5EED0003, 37
    tied = [1, 4, 33, 36]
    tracks = list(range(T))
    u0 = synth_u_init(seed, tracks)
    fl = [orc.Filter(left_arm.orc, N) for _ in tracks]
    fr_ = [orc.Filter(right_arm.orc, N) for _ in tracks]
    for t in tracks:
        fl[t].reset(u=u0[t])
        fr_[t].reset(u=u0[t])
    s = torch.cuda.Stream()
    b0 = mk.TrackBatch(left_arm.mk, T, N, stream=s.cuda_stream)
    b1 = mk.TrackBatch(right_arm.mk, T, N, stream=s.cuda_stream)
    b0.reset(u0)
    b1.reset(u0)
    roi = np.tile(np.array([300.0, 51.0, 47.0, 47.0]), (T, 1))
    rng = np.random.default_rng(11)
    cand = rng.uniform(40, 400, (T, 2, 2, Cn))
    Lv = rng.integers(1, 255, (T, 2, Cn)).astype(np.uint8)
    u_cand = rng.random((T, 2))
    for t in tied:
        cand[t] = cand[t, :, :, :1]
        Lv[t] = 128
        u_cand[t] = 0.0
    mk.associate(b0, b1, cand, Lv, roi, u_cand, None, None, do_update=False)
    res = mk.assoc_results(b0, Cn)
    for t in tracks:
        want = orc.associate(fl[t], fr_[t], cand[t], Lv[t], roi[t], u_cand[t])
        assert np.array_equal(res["gate"][t], want["gate"])
        assert rel_err_weights(res["weights"][t], want["weights"]) <= RTOL
        if t in tied:
            # with every threshold sitting on a prefix sum the outcome hinges on the last bit of the weights, and the
            # device's weight sum is a tree sum (the oracle's is sequential): the contract that can hold here is the
            # literal loop applied to the weights the device itself produced
            for h in range(2):
                lit, _ = orc.resample(res["weights"][t, h], N, 0.0)
                assert np.array_equal(res["bins"][t, h], lit), f"person {t} hand {h}"
        else:
            assert np.array_equal(res["bins"][t], want["bins"]), f"person {t}"
    st = b0.status() | b1.status()
    assert (st[tied] & L.ST_CAND_FALLBACK).all(), "tied persons were expected to take the literal loop"
    assert not (st[[t for t in tracks if t not in tied]] & L.ST_CAND_FALLBACK).any()


def spd(rng, n, scale):
    a = rng.standard_normal((n, n))
    return scale * (a @ a.T + n * np.eye(n))


@pytest.mark.parametrize("d,N,frames", [(8, 300, 60), (8, 5000, 12), (12, 257, 50), (10, 100, 50)])
def test_pf2d_matches_oracle(d, N, frames):
    """legacy plain particle filter (src/pf2D.cpp:148-268) FREE-RUNNING against the oracle: the device filter is never
    re-synchronised, and every resampled index of every frame must equal the oracle's (the float expf of
    src/pf2D.cpp:108 is the shared restatement of glibc's, include/mkf_expf.h)."""
    rng = np.random.default_rng(11)
    T, K = 3, 15
    means = rng.uniform(100, 400, (K, d))
    covs = np.stack([spd(rng, d, 40.0) for _ in range(K)])
    wts = rng.dirichlet(np.ones(K))
    pb = mk.Pf2dBatch(T, N, means, covs, wts)
    parts = means[rng.integers(0, K, (T, N))] + rng.standard_normal((T, N, d)) * 6
    pb.set_particles(parts)
    ofs = []
    for t in range(T):
        o = orc.Pf2d(N, means, covs, wts)
        o.set_particles(parts[t])
        ofs.append(o)
    # legacy getEstimator (src/pf2D.cpp:79-88) before any update: weights are the constructor's 1/N
    est0 = pb.estimate()
    for t in range(T):
        assert np.max(np.abs(est0[t] - ofs[t].estimate()) / np.abs(ofs[t].estimate())) <= 1e-12
    worst_w = 0.0
    for frame in range(frames):
        cur = np.stack([o.get()[0] for o in ofs])
        meas = np.stack([np.array([[c[:, 6].mean(), c[:, 7].mean()], [c[:, 0].mean(), c[:, 1].mean()]]) for c in cur])
        u = rng.random(T)
        noise = rng.standard_normal((T, N, d))
        pb.update(meas, u, noise)
        p, w, par = pb.get()
        for t in range(T):
            r = ofs[t].update(meas[t], u[t], noise[t])
            assert r["status"] == 0
            worst_w = max(worst_w, rel_err_weights(w[t], r["w_norm"]))
            assert np.array_equal(par[t], r["parents"]), f"frame {frame} filter {t}: resampled indices differ"
            assert np.array_equal(p[t], ofs[t].get()[0]), f"frame {frame} filter {t}: particles differ"
        # getEstimator after the update: the un-reset normalised weights against the resampled, predicted particles
        est = pb.estimate()
        for t in range(T):
            eo = ofs[t].estimate()
            assert np.max(np.abs(est[t] - eo) / np.abs(eo)) <= 1e-11
    print(f"pf2d d={d} N={N}: {frames} free-running frames, 0 index mismatches, worst weight error {worst_w:.1e}")
    assert worst_w <= 1e-9


def test_pf2d_constructor_and_degenerate_branch():
    """ParticleFilter(numParticles, numDims, side1) (src/pf2D.cpp:44-71) and the `mw == 0` branch of resample()
    (src/pf2D.cpp:232-250): particles re-drawn across the image, weights back to 1/N, then predict()."""
    rng = np.random.default_rng(5)
    T, N, d, K = 4, 333, 8, 6
    means = rng.uniform(100, 400, (K, d))
    covs = np.stack([spd(rng, d, 40.0) for _ in range(K)])
    wts = rng.dirichlet(np.ones(K))
    seed, track0 = 0xC0FFEE, 1000
    side = np.array([0, 1, 1, 0], np.uint8)
    pb = mk.Pf2dBatch(T, N, means, covs, wts)
    pb.set_random(seed, track0, side, 640, 480)
    pb.randomise()
    ofs = []
    for t in range(T):
        o = orc.Pf2d(N, means, covs, wts)
        o.set_random(seed, track0 + t, int(side[t]))
        o.randomise()
        ofs.append(o)
    p0, _, _ = pb.get()
    for t in range(T):
        assert np.array_equal(p0[t], ofs[t].get()[0])
        # the constructor's ranges: [1, 640) even columns, [1, 480) odd ones, column 6 in the half `side` selects
        assert p0[t][:, 0::2].min() >= 1 and p0[t][:, 0::2].max() < 640 and p0[t][:, 1::2].max() < 480
        lo, hi = (321, 640) if side[t] else (1, 320)
        assert p0[t][:, 6].min() >= lo and p0[t][:, 6].max() < hi
        assert np.max(np.abs(pb.estimate()[t] - ofs[t].estimate()) / np.abs(ofs[t].estimate())) <= 1e-12
    # particles drawn across the whole image sit far outside the prior: their weights live in the subnormal range, where
    # one ulp of libm's / CUDA's double exp decides between "tiny" and "exactly 0" -- not a regime with a parity
    # contract.  The filters are put near the components, then knocked out on two frames (every likelihood exactly 0)
    # and re-seeded.
    n_dead = 0
    for frame in range(8):
        if frame in (0, 5):
            near = means[rng.integers(0, K, (T, N))] + rng.standard_normal((T, N, d)) * 6
            pb.set_particles(near)
            for t in range(T):
                ofs[t].set_particles(near[t])
        cur = np.stack([o.get()[0] for o in ofs])
        meas = np.stack([np.array([[c[:, 6].mean(), c[:, 7].mean()], [c[:, 0].mean(), c[:, 1].mean()]]) for c in cur])
        if frame in (3, 6):
            meas[[1, 2]] += 1.0e5  # every likelihood underflows to 0: weight sum 0, weights NaN, max weight "0"
        u = rng.random(T)
        noise = rng.standard_normal((T, N, d))
        pb.update(meas, u, noise)
        p, w, par = pb.get()
        est = pb.estimate()
        for t in range(T):
            r = ofs[t].update(meas[t], u[t], noise[t])
            if frame in (3, 6) and t in (1, 2):
                assert r["status"] == 1
            if frame in (0, 1, 2, 5):
                assert r["status"] == 0
            n_dead += r["status"]
            assert np.array_equal(par[t], r["parents"]), (frame, t)
            assert np.array_equal(p[t], ofs[t].get()[0]), f"frame {frame} filter {t}"
            eo = ofs[t].estimate()
            assert np.max(np.abs(est[t] - eo) / np.abs(eo)) <= 1e-11
            if r["status"]:
                assert np.isnan(w[t]).all() and np.array_equal(par[t], np.arange(N))
            else:
                assert rel_err_weights(w[t], r["w_norm"]) <= 1e-9
    assert n_dead >= 4
