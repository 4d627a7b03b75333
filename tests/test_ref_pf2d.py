"""The oracle's legacy plain particle filter against the reference's OWN src/pf2D.cpp (compiled in place against
oracle/cvshim -> oracle/_ref/libref_pf2d.so): loadGaussian members, weights, resampled particles, predict, estimator,
free-running over many frames.  Runs where /root/reference exists or the prebuilt library travelled with the repo."""
import numpy as np
import pytest

import mkf_oracle as orc
import mkf_ref

pytestmark = pytest.mark.skipif(not mkf_ref.pf2d_available(), reason="oracle/_ref/libref_pf2d.so not built")


def spd(rng, n, scale):
    a = rng.standard_normal((n, n))
    return scale * (a @ a.T + n * np.eye(n))


def make(rng, N, K=6, d=8, side=1, seed=3):
    means = rng.uniform(100, 400, (K, d))
    covs = np.stack([spd(rng, d, 40.0) for _ in range(K)])
    wts = rng.dirichlet(np.ones(K))
    ref = mkf_ref.RefPf2d(N, d, side, means, covs, wts, rng_seed=seed)
    o = orc.Pf2d(N, means, covs, wts)
    o.set_noise_scaled(True)
    return ref, o, means, covs, wts


def test_load_gaussian_members_bit_exact(rng):
    """sigma_i = invert(s, DECOMP_CHOLESKY), det_s = 1 / (pow(2 pi, d/2) sqrt(cv::determinant(s)))  (src/pf2D.cpp:28-37)"""
    ref, o, *_ = make(rng, 50)
    si_r, ds_r = ref.gmm()
    si_o, ds_o = o.gmm()
    assert np.array_equal(si_r, si_o)
    assert np.array_equal(ds_r, ds_o)


def test_constructor_ranges(rng):
    """ParticleFilter(N, d, side1) (src/pf2D.cpp:44-71): uniform weights; columns uniform over the image, column 6 over
    the half `side` selects (the draws are cv::randu's: ranges only)"""
    for side in (0, 1):
        ref, *_ = make(rng, 400, side=side)
        p, w = ref.get()
        assert np.all(w == 1.0 / 400)
        assert p[:, 0::2].min() >= 1 and p[:, 0::2].max() < 640 and p[:, 1::2].min() >= 1 and p[:, 1::2].max() < 480
        lo, hi = (321, 640) if side else (1, 320)
        assert p[:, 6].min() >= lo and p[:, 6].max() < hi


@pytest.mark.parametrize("N", [300, 1000])
def test_update_free_running_bit_exact(rng, N):
    """ParticleFilter::update (src/pf2D.cpp:148-210) -- weights with the float expf, normalise, resample() with the C
    library's uniform, predict() -- free-running for 25 frames: the oracle, given the same uniform and the noise the
    reference's cv::randn calls returned, reproduces the reference's particles and weights BIT FOR BIT"""
    ref, o, means, covs, wts = make(rng, N)
    K, d = means.shape
    parts = means[rng.integers(0, K, N)] + rng.standard_normal((N, d)) * 6
    ref.set_particles(parts)
    o.set_particles(parts)
    assert np.array_equal(ref.estimate(), o.estimate())  # constructor weights 1/N
    for frame in range(25):
        cur, _ = ref.get()
        meas = np.array([[cur[:, 6].mean(), cur[:, 7].mean()], [cur[:, 0].mean(), cur[:, 1].mean()]])
        u, noise, deg = ref.update(meas, srand_seed=1000 + frame)
        r = o.update(meas, u, noise)
        assert not deg and r["status"] == 0
        pr, wr = ref.get()
        po, wo = o.get()
        assert np.array_equal(wr, wo), f"frame {frame}: normalised weights"
        assert np.array_equal(pr, po), f"frame {frame}: resampled + predicted particles"
        assert np.array_equal(ref.estimate(), o.estimate())


def test_degenerate_branch_structure(rng):
    """`mw == 0` (src/pf2D.cpp:232-250): every particle re-drawn across the image, weights back to 1/N, then predict().
    The draws themselves are cv::randu's in the reference and the counter generator's in the oracle: same structure."""
    N = 200
    ref, o, means, covs, wts = make(rng, N, side=0)
    K, d = means.shape
    parts = means[rng.integers(0, K, N)] + rng.standard_normal((N, d)) * 6
    ref.set_particles(parts)
    o.set_particles(parts)
    far = np.array([[1e5, 1e5], [1e5, 1e5]])
    u, noise, deg = ref.update(far, srand_seed=5)
    r = o.update(far, u, noise)
    assert deg and r["status"] == 1
    for p, w in (ref.get(), o.get()):
        assert np.all(w == 1.0 / N)
        base = p - noise  # what the re-randomisation drew, before predict() added its noise
        assert base[:, 0::2].min() >= 1 - 1e-9 and base[:, 0::2].max() < 640 and base[:, 1::2].max() < 480
        assert base[:, 6].max() < 320 + 1e-9
