"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

CHOL_CV24_LITERAL, CHOL_CV3_LITERAL, CHOL_EXACT = 0, 1, 2
ALIAS_INDEPENDENT, ALIAS_CV_SHALLOW_LITERAL = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)


def _ptr(a, typ):
    if a is None:
        return C.cast(None, typ)
    return a.ctypes.data_as(typ)


def build(force: bool = False) -> str:
    """Compile liboracle.so (and _ref/libref.so when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("mkf_oracle.cpp", "mkf_oracle.h")] + [
        os.path.join(_HERE, "..", "include", "mkf_synth.h")
    ]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(so):
        build()
    else:
        try:
            build()
        except Exception:
            pass  # no compiler on this box: use the prebuilt file
    L = C.CDLL(so)
    L.orc_model_create.restype = C.c_void_p
    L.orc_model_create.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
    L.orc_model_destroy.argtypes = [C.c_void_p]
    L.orc_model_get.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp]
    L.orc_filter_create.restype = C.c_void_p
    L.orc_filter_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.orc_filter_destroy.argtypes = [C.c_void_p]
    L.orc_filter_reset.argtypes = [C.c_void_p, C.c_double, C.c_uint64]
    L.orc_filter_update.argtypes = [C.c_void_p, _dp, C.c_double, C.c_uint64, C.c_double, C.c_uint64, _dp, _dp, _ip,
                                    _ip, _dp]
    L.orc_filter_update_shared.argtypes = L.orc_filter_update.argtypes
    L.orc_filter_get_state.argtypes = [C.c_void_p, _dp, _dp]
    L.orc_filter_set_state.argtypes = [C.c_void_p, _dp, _dp]
    L.orc_filter_estimate.argtypes = [C.c_void_p, _dp, _dp]
    L.orc_resample.argtypes = [_dp, C.c_int, C.c_int, C.c_double, C.c_uint64, _ip]
    L.orc_kf_predict.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    L.orc_kf_update.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
    L.orc_mvnpdf.restype = C.c_double
    L.orc_mvnpdf.argtypes = [C.c_int, _dp, _dp, _dp, C.c_int, C.POINTER(C.c_int)]
    L.orc_chol.argtypes = [C.c_int, _dp, _dp, C.c_int]
    L.orc_invert_lu.argtypes = [C.c_int, _dp, _dp]
    L.orc_cvrng.argtypes = [C.c_uint64, C.c_int, C.c_int, _ip, C.c_int, _dp]
    L.orc_associate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _dp, _bp, _dp, C.c_int, C.c_int, _dp, _u64p, _bp,
                                _dp, _ip, _dp]
    L.orc_propose.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _dp, C.c_int, _bp, C.c_int, C.c_int, C.c_uint64,
                              C.c_uint64, C.c_uint64, _dp, _bp]
    L.orc_get3dpose.argtypes = [_dp, _dp, _dp]
    L.orc_skeleton.argtypes = [_dp, _dp, _dp, _dp, _dp]
    L.orc_pf2d_create.restype = C.c_void_p
    L.orc_pf2d_create.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]
    L.orc_pf2d_destroy.argtypes = [C.c_void_p]
    L.orc_pf2d_set_particles.argtypes = [C.c_void_p, _dp]
    L.orc_pf2d_set_random.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int]
    L.orc_pf2d_set_random.restype = None
    L.orc_pf2d_randomise.argtypes = [C.c_void_p]
    L.orc_pf2d_randomise.restype = None
    L.orc_pf2d_set_noise_scaled.argtypes = [C.c_void_p, C.c_int]
    L.orc_pf2d_set_noise_scaled.restype = None
    L.orc_expf.argtypes = [C.c_float]
    L.orc_expf.restype = C.c_float
    L.orc_libm_expf.argtypes = [C.c_float]
    L.orc_libm_expf.restype = C.c_float
    L.orc_expf_compare.argtypes = [C.c_void_p, C.c_uint64]
    L.orc_expf_compare.restype = C.c_uint64
    L.orc_pf2d_get_particles.argtypes = [C.c_void_p, _dp, _dp]
    L.orc_pf2d_get_gmm.argtypes = [C.c_void_p, _dp, _dp]
    L.orc_pf2d_estimate.argtypes = [C.c_void_p, _dp]
    L.orc_pf2d_update.argtypes = [C.c_void_p, _dp, C.c_double, _dp, _dp, _ip]
    L.orc_bench_tracks.restype = C.c_double
    L.orc_bench_tracks.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int,
                                   C.c_int, C.c_int, _dp, C.POINTER(C.c_int)]
    L.orc_synth_meas.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int64, C.c_int, _dp]
    L.orc_synth_u.restype = C.c_double
    L.orc_synth_u.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
    L.orc_synth_candidate.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, _dp,
                                      _dp, _bp]
    _LIB = L
    return L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Model:
    """my_gmm::loadGaussian for every component (src/my_gmm.cpp:45-75)."""

    def __init__(self, means, covs, weights, gamma, pca_proj, pca_mean):
        L = lib()
        self.means = _f64(means)
        self.K, self.d = self.means.shape
        self.covs = _f64(covs).reshape(self.K, self.d, self.d)
        self.weights = _f64(weights).reshape(-1)
        self.gamma = _f64(gamma).reshape(-1)
        self.pca_proj = _f64(pca_proj)
        self.D = self.pca_proj.shape[1]
        self.pca_mean = _f64(pca_mean).reshape(-1)
        self.h = L.orc_model_create(self.K, self.d, self.D, _ptr(self.means, _dp), _ptr(self.covs, _dp),
                                    _ptr(self.weights, _dp), _ptr(self.gamma, _dp), _ptr(self.pca_proj, _dp),
                                    _ptr(self.pca_mean, _dp))
        if not self.h:
            raise ValueError("orc_model_create failed")

    def constants(self):
        H = np.zeros((6, self.d))
        BH = np.zeros(6)
        Q = np.zeros((self.K, self.d, self.d))
        B = np.zeros((self.K, self.d))
        R = np.zeros((6, 6))
        lib().orc_model_get(self.h, _ptr(H, _dp), _ptr(BH, _dp), _ptr(Q, _dp), _ptr(B, _dp), _ptr(R, _dp))
        return dict(H=H, BH=BH, Q=Q, B=B, R=R)

    def kf_predict(self, k, x, P):
        x = _f64(x).copy()
        P = _f64(P).copy()
        lib().orc_kf_predict(self.h, k, _ptr(x, _dp), _ptr(P, _dp))
        return x, P

    def kf_update(self, k, z, x, P):
        x = _f64(x).copy()
        P = _f64(P).copy()
        z = _f64(z)
        lib().orc_kf_update(self.h, k, _ptr(z, _dp), _ptr(x, _dp), _ptr(P, _dp))
        return x, P

    def __del__(self):
        try:
            if self.h:
                lib().orc_model_destroy(self.h)
                self.h = None
        except Exception:
            pass


class Filter:
    """ParticleFilter of src/pf2DRao.{h,cpp} (one arm / one track)."""

    def __init__(self, model: Model, N: int, chol_mode=CHOL_CV24_LITERAL, alias_mode=ALIAS_INDEPENDENT):
        self.model = model
        self.N = N
        self.h = lib().orc_filter_create(model.h, N, chol_mode, alias_mode)

    def reset(self, u=-1.0, seed=1):
        return lib().orc_filter_reset(self.h, float(u), int(seed))

    def update(self, meas, u_ind, u_post, seed_ind=1, seed_post=1):
        """meas: (6,) shared column or (6, N).  returns dict(status, w_raw, w_norm, indicators, parents, wsum)."""
        meas = _f64(meas)
        N = self.N
        w_raw = np.zeros(N)
        w_norm = np.zeros(N)
        ind = np.zeros(N, np.int32)
        par = np.zeros(N, np.int32)
        wsum = C.c_double(0)
        fn = lib().orc_filter_update_shared if meas.ndim == 1 else lib().orc_filter_update
        if meas.ndim == 2:
            assert meas.shape == (6, N)
        st = fn(self.h, _ptr(meas, _dp), float(u_ind), int(seed_ind), float(u_post), int(seed_post), _ptr(w_raw, _dp),
                _ptr(w_norm, _dp), _ptr(ind, _ip), _ptr(par, _ip), C.byref(wsum))
        return dict(status=st, w_raw=w_raw, w_norm=w_norm, indicators=ind, parents=par, wsum=wsum.value)

    def get_state(self):
        d = self.model.d
        x = np.zeros((self.N, d))
        P = np.zeros((self.N, d, d))
        lib().orc_filter_get_state(self.h, _ptr(x, _dp), _ptr(P, _dp))
        return x, P

    def set_state(self, x, P):
        x = _f64(x)
        P = _f64(P)
        lib().orc_filter_set_state(self.h, _ptr(x, _dp), _ptr(P, _dp))

    def estimate(self):
        xbar = np.zeros(self.model.d)
        pose = np.zeros(self.model.D)
        lib().orc_filter_estimate(self.h, _ptr(xbar, _dp), _ptr(pose, _dp))
        return xbar, pose

    def __del__(self):
        try:
            if self.h:
                lib().orc_filter_destroy(self.h)
                self.h = None
        except Exception:
            pass


def resample(w, N, u=-1.0, seed=1):
    w = _f64(w)
    out = np.zeros(N, np.int32)
    deg = lib().orc_resample(_ptr(w, _dp), len(w), N, float(u), int(seed), _ptr(out, _ip))
    return out, deg


def chol(a, mode=CHOL_CV24_LITERAL):
    a = _f64(a)
    n = a.shape[0]
    out = np.zeros((n, n))
    ok = lib().orc_chol(n, _ptr(a, _dp), _ptr(out, _dp), mode)
    return out, bool(ok)


def mvnpdf(x, u, sigma, mode=CHOL_CV24_LITERAL):
    x = _f64(x)
    u = _f64(u)
    sigma = _f64(sigma)
    ok = C.c_int(0)
    v = lib().orc_mvnpdf(len(x), _ptr(x, _dp), _ptr(u, _dp), _ptr(sigma, _dp), mode, C.byref(ok))
    return v, bool(ok.value)


def invert_lu(a):
    a = _f64(a)
    n = a.shape[0]
    out = np.zeros((n, n))
    ok = lib().orc_invert_lu(n, _ptr(a, _dp), _ptr(out, _dp))
    return out, bool(ok)


def cvrng(seed, L, n_int, n_dbl):
    oi = np.zeros(max(n_int, 1), np.int32)
    od = np.zeros(max(n_dbl, 1))
    lib().orc_cvrng(int(seed), L, n_int, _ptr(oi, _ip), n_dbl, _ptr(od, _dp))
    return oi[:n_int], od[:n_dbl]


def associate(armL: Filter, armR: Filter, cand_xy, cand_L, roi, u_cand, img_rows=480, img_cols=640, seed_cand=None):
    """cand_xy (2 hands, 2, C), cand_L (2, C) uint8, roi (4,), u_cand (2,)."""
    cand_xy = _f64(cand_xy)
    Cn = cand_xy.shape[2]
    cand_L = np.ascontiguousarray(cand_L, np.uint8)
    roi = _f64(roi)
    u_cand = _f64(u_cand)
    N = armL.N
    seeds = np.ascontiguousarray(seed_cand if seed_cand is not None else [1, 1], np.uint64)
    gate = np.zeros((2, Cn), np.uint8)
    w = np.zeros((2, Cn))
    bins = np.zeros((2, N), np.int32)
    meas = np.zeros((2, 6, N))
    st = lib().orc_associate(armL.h, armR.h, Cn, _ptr(cand_xy, _dp), _ptr(cand_L, _bp), _ptr(roi, _dp), img_rows,
                             img_cols, _ptr(u_cand, _dp), _ptr(seeds, _u64p), _ptr(gate, _bp), _ptr(w, _dp),
                             _ptr(bins, _ip), _ptr(meas, _dp))
    return dict(status=st, gate=gate, weights=w, bins=bins, meas=meas)


class Pf2d:
    """legacy plain particle filter of src/pf2D.{h,cpp}."""

    def __init__(self, N, means, covs, weights):
        means = _f64(means)
        self.K, self.d = means.shape
        self.N = N
        covs = _f64(covs)
        weights = _f64(weights)
        self.h = lib().orc_pf2d_create(N, self.d, self.K, _ptr(means, _dp), _ptr(covs, _dp), _ptr(weights, _dp))
        if not self.h:
            raise ValueError("orc_pf2d_create failed")

    def set_particles(self, p):
        p = _f64(p)
        assert p.shape == (self.N, self.d)
        lib().orc_pf2d_set_particles(self.h, _ptr(p, _dp))

    def set_random(self, seed, track=0, side=0, im_w=640, im_h=480):
        """parameters of the constructor / degenerate-branch randomisation (src/pf2D.cpp:44-71,232-250)"""
        lib().orc_pf2d_set_random(self.h, int(seed), int(track), int(side), int(im_w), int(im_h))

    def randomise(self):
        """the constructor's draw: particles across the image, weights 1/N"""
        lib().orc_pf2d_randomise(self.h)

    def set_noise_scaled(self, on=True):
        """update()'s `noise` then holds what predict() adds (already x 5), as the reference's cv::randn calls return it"""
        lib().orc_pf2d_set_noise_scaled(self.h, 1 if on else 0)

    def get(self):
        p = np.zeros((self.N, self.d))
        w = np.zeros(self.N)
        lib().orc_pf2d_get_particles(self.h, _ptr(p, _dp), _ptr(w, _dp))
        return p, w

    def estimate(self):
        est = np.zeros(self.d)
        lib().orc_pf2d_estimate(self.h, _ptr(est, _dp))
        return est

    def gmm(self):
        si = np.zeros((self.K, self.d, self.d))
        ds = np.zeros(self.K)
        lib().orc_pf2d_get_gmm(self.h, _ptr(si, _dp), _ptr(ds, _dp))
        return si, ds

    def update(self, meas, u, noise=None):
        meas = _f64(meas)
        wn = np.zeros(self.N)
        par = np.zeros(self.N, np.int32)
        nz = _f64(noise) if noise is not None else None
        st = lib().orc_pf2d_update(self.h, _ptr(meas, _dp), float(u), _ptr(nz, _dp), _ptr(wn, _dp), _ptr(par, _ip))
        return dict(status=st, w_norm=wn, parents=par)

    def __del__(self):
        try:
            if self.h:
                lib().orc_pf2d_destroy(self.h)
                self.h = None
        except Exception:
            pass


def bench_tracks(model: Model, T, N, frames, per_slot=False, seed=0x5EED0002, jitter=1, chol_mode=CHOL_CV24_LITERAL,
                 alias_mode=ALIAS_INDEPENDENT, threads=0, want_pose=False):
    pose = np.zeros((T, model.D)) if want_pose else None
    used = C.c_int(0)
    secs = lib().orc_bench_tracks(model.h, T, N, frames, int(per_slot), int(seed), int(jitter), chol_mode, alias_mode,
                                  threads, _ptr(pose, _dp), C.byref(used))
    return secs, used.value, pose


def synth_meas(seed, track, frame, slot=-1, jitter=1):
    z = np.zeros(6)
    lib().orc_synth_meas(int(seed), int(track), int(frame), int(slot), int(jitter), _ptr(z, _dp))
    return z


def synth_u(seed, track, frame, which):
    return lib().orc_synth_u(int(seed), int(track), int(frame), int(which))


def synth_candidate(seed, track, frame, hand, Cn, c, jitter=1):
    cx = C.c_double(0)
    cy = C.c_double(0)
    Lv = C.c_uint8(0)
    lib().orc_synth_candidate(int(seed), int(track), int(frame), hand, Cn, c, jitter, C.byref(cx), C.byref(cy),
                              C.byref(Lv))
    return cx.value, cy.value, Lv.value


KINECT_K = np.array([[525.0, 0.0, 319.5], [0.0, 525.0, 239.5], [0.0, 0.0, 1.0]])  # src/pfPose.cpp:101


def get3dpose(estimate, K=KINECT_K):
    e = _f64(estimate)
    K = _f64(K)
    out = np.zeros((3, 5))
    lib().orc_get3dpose(_ptr(e, _dp), _ptr(K, _dp), _ptr(out, _dp))
    return out


def skeleton(e1, e2, K=KINECT_K):
    e1, e2, K = _f64(e1), _f64(e2), _f64(K)
    tf = np.zeros((10, 3))
    j2 = np.zeros((8, 2))
    lib().orc_skeleton(_ptr(e1, _dp), _ptr(e2, _dp), _ptr(K, _dp), _ptr(tf, _dp), _ptr(j2, _dp))
    return tf, j2


def propose(armL: Filter, armR: Filter, Cn, roi, tracking, like, seed, track, frame):
    roi = _f64(roi)
    xy = np.zeros((2, 2, Cn))
    Lv = np.zeros((2, Cn), np.uint8) if like is not None else None
    rows, cols = (like.shape if like is not None else (480, 640))
    like_c = np.ascontiguousarray(like, np.uint8) if like is not None else None
    lib().orc_propose(armL.h, armR.h, Cn, _ptr(roi, _dp), int(tracking), _ptr(like_c, _bp), rows, cols, int(seed),
                      int(track), int(frame), _ptr(xy, _dp), _ptr(Lv, _bp))
    return xy, Lv


def expf(x):
    """include/mkf_expf.h: glibc's expf restated (what the oracle and the device use for src/pf2D.cpp:108)"""
    return float(lib().orc_expf(float(np.float32(x))))


def libm_expf(x):
    return float(lib().orc_libm_expf(float(np.float32(x))))


def expf_compare(bits):
    """bit-level mismatches between mkf_expf and the host libm's expf over float bit patterns `bits` (uint32)"""
    bits = np.ascontiguousarray(bits, dtype=np.uint32)
    return int(lib().orc_expf_compare(bits.ctypes.data, bits.size))
