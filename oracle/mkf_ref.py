"""ctypes binding of oracle/_ref/libref.so: the reference's OWN src/KF_model.cpp, src/my_gmm.cpp and
src/pf2DRao.cpp compiled in place against the OpenCV-subset shim oracle/cvshim (oracle/Makefile).

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py CPU legs).  The library is rebuilt only where
/root/reference exists; elsewhere (the GPU box) the prebuilt file that travelled with the snapshot is used.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref.so")
_LIB = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def available() -> bool:
    if os.path.isdir("/root/reference/src"):
        try:
            subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)
        except Exception:
            pass
    return os.path.exists(SO)


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise FileNotFoundError(SO)
        L = C.CDLL(SO)
        L.ref_pf_create.restype = C.c_void_p
        L.ref_pf_create.argtypes = [C.c_int]
        L.ref_pf_destroy.argtypes = [C.c_void_p]
        L.ref_pf_load_model.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
        L.ref_pf_get_kf.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
        L.ref_pf_reset.argtypes = [C.c_void_p, C.c_int64]
        L.ref_pf_update.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int64]
        L.ref_pf_get_state.argtypes = [C.c_void_p, _dp, _dp]
        L.ref_pf_estimate.argtypes = [C.c_void_p, _dp]
        L.ref_pf_resample.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, C.c_int64, _ip]
        L.ref_pf_chol.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.ref_pf_mvnpdf.restype = C.c_double
        L.ref_pf_mvnpdf.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
        L.ref_kf_predict.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.ref_kf_update.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
        L.ref_pf_sample_prob.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _dp, C.c_int, C.c_double,
                                         _dp, _dp]
        L.ref_pf_get_samples_mean.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_double, _dp, _dp]
        L.ref_bench_tracks.restype = C.c_double
        L.ref_bench_tracks.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int64, C.c_int,
                                       C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_int)]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(_dp)


def _f(a):
    return np.ascontiguousarray(a, np.float64)


class RefFilter:
    """the reference's ParticleFilter (src/pf2DRao.h:13-31) with its gmm loaded as src/pfPose.cpp:61-65 does"""

    def __init__(self, arrays, N):
        self.a = {k: _f(v) for k, v in arrays.items()}
        self.K, self.d = self.a["means"].shape
        self.D = self.a["pca_proj"].shape[1]
        self.N = N
        self.h = lib().ref_pf_create(N)
        a = self.a
        lib().ref_pf_load_model(self.h, self.K, self.d, self.D, _p(a["means"]), _p(a["covs"]), _p(a["weights"]),
                                _p(a["gamma"]), _p(a["pca_proj"]), _p(a["pca_mean"]))

    def kf_members(self, k):
        d = self.d
        out = dict(Q=np.zeros((d, d)), B=np.zeros(d), H=np.zeros((6, d)), BH=np.zeros(6), R=np.zeros((6, 6)),
                   F=np.zeros((d, d)))
        lib().ref_pf_get_kf(self.h, k, *[_p(out[n]) for n in ("Q", "B", "H", "BH", "R", "F")])
        return out

    def reset(self, tick):
        lib().ref_pf_reset(self.h, int(tick))

    def update(self, meas, tick_ind, tick_post):
        meas = _f(meas)
        assert meas.shape == (6, self.N)
        lib().ref_pf_update(self.h, _p(meas), int(tick_ind), int(tick_post))

    def get_state(self):
        x = np.zeros((self.N, self.d))
        P = np.zeros((self.N, self.d, self.d))
        lib().ref_pf_get_state(self.h, _p(x), _p(P))
        return x, P

    def estimate(self):
        xb = np.zeros(self.d)
        lib().ref_pf_estimate(self.h, _p(xb))
        return xb

    def resample(self, w, N, tick):
        w = _f(w)
        out = np.zeros(N, np.int32)
        lib().ref_pf_resample(self.h, _p(w), len(w), N, int(tick), out.ctypes.data_as(_ip))
        return out

    def chol(self, S):
        S = _f(S)
        out = np.zeros_like(S)
        lib().ref_pf_chol(self.h, S.shape[0], _p(S), _p(out))
        return out

    def mvnpdf(self, x, u, S):
        x, u, S = _f(x), _f(u), _f(S)
        return lib().ref_pf_mvnpdf(self.h, len(x), _p(x), _p(u), _p(S))

    def kf_predict(self, k, x, P):
        x, P = _f(x).copy(), _f(P).copy()
        lib().ref_kf_predict(self.h, k, _p(x), _p(P))
        return x, P

    def kf_update(self, k, z, x, P):
        z, x, P = _f(z), _f(x).copy(), _f(P).copy()
        lib().ref_kf_update(self.h, k, _p(z), _p(x), _p(P))
        return x, P

    def sample_prob(self, in1, in2, scale):
        in1, in2 = _f(in1), _f(in2)
        w1, w2 = np.zeros(in1.shape[1]), np.zeros(in2.shape[1])
        lib().ref_pf_sample_prob(self.h, self.d, self.D, _p(self.a["pca_proj"]), _p(self.a["pca_mean"]), _p(in1),
                                 in1.shape[1], _p(in2), in2.shape[1], float(scale), _p(w1), _p(w2))
        return w1, w2

    def samples_mean_sd(self, N, scale):
        m, s = np.zeros(2), np.zeros(2)
        lib().ref_pf_get_samples_mean(self.h, self.d, self.D, _p(self.a["pca_proj"]), _p(self.a["pca_mean"]), N,
                                      float(scale), _p(m), _p(s))
        return m, s

    def __del__(self):
        try:
            if self.h:
                lib().ref_pf_destroy(self.h)
                self.h = None
        except Exception:
            pass


def tick_to_u(tick, L):
    """the uniform draw cv::RNG(tick) produces inside resample(): one discarded uniform(0,L), then uniform(0.0,1.0)"""
    s = int(tick) & 0xFFFFFFFFFFFFFFFF
    if s == 0:
        s = 0xFFFFFFFF

    def nxt():
        nonlocal s
        s = ((s & 0xFFFFFFFF) * 4164903690 + (s >> 32)) & 0xFFFFFFFFFFFFFFFF
        return s & 0xFFFFFFFF

    nxt()
    t = nxt()
    return ((t << 32) | nxt()) * 5.4210108624275221700372640043497e-20


def bench_tracks(arrays, T, N, frames, per_slot=False, seed=0x5EED0002, jitter=1, threads=0):
    a = {k: _f(v) for k, v in arrays.items()}
    K, d = a["means"].shape
    D = a["pca_proj"].shape[1]
    used = C.c_int(0)
    secs = lib().ref_bench_tracks(K, d, D, _p(a["means"]), _p(a["covs"]), _p(a["weights"]), _p(a["gamma"]),
                                  _p(a["pca_proj"]), _p(a["pca_mean"]), T, N, frames, int(per_slot), int(seed),
                                  int(jitter), threads, C.byref(used))
    return secs, used.value


_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i64p = C.POINTER(C.c_int64)


def _bind_tracker(L):
    if getattr(L, "_tracker_bound", False):
        return
    L.ref_tracker_create.restype = C.c_void_p
    L.ref_tracker_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int64, C.c_int64]
    L.ref_tracker_destroy.argtypes = [C.c_void_p]
    L.ref_set_rng_seed.argtypes = [C.c_uint64]
    L.ref_tracker_num_particles.argtypes = [C.c_void_p]
    L.ref_random_log_get.argtypes = [C.c_int, _dp, C.c_int]
    L.ref_tracker_callback.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, _u32p, _i64p, C.c_int]
    L.ref_tracker_outputs.argtypes = [_dp, _dp, _u8p, C.c_int, C.c_int]
    L.ref_tracker_get_state.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    L.ref_tracker_pose.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    L._tracker_bound = True


class RefTracker:
    """the reference's PFTracker (src/pfPose.{h,cpp}) driven without ROS: one synthetic frame per call()"""

    def _load(self):
        return lib()

    def __init__(self, model_dir, left_file, right_file, tick1, tick2, rng_seed=0x12345678):
        L = self._L = self._load()
        _bind_tracker(L)
        L.ref_set_rng_seed(int(rng_seed))  # the generator behind the node's cv::randn / cv::randu draws
        self.h = L.ref_tracker_create(os.fsencode(model_dir), os.fsencode("/" + left_file),
                                      os.fsencode("/" + right_file), int(tick1), int(tick2))
        self.N = L.ref_tracker_num_particles(self.h)

    def callback(self, like, roi_xywh, ticks):
        """like: (rows, cols) uint8 raw likelihood image; roi (x, y, w, h) or None (no face); ticks: 6 ints.
        returns dict(cands=[hand][2, C], blurred=(rows, cols) uint8, joints2d (8,2), tf (10,3), n_tf)"""
        L = self._L
        like = np.ascontiguousarray(like, np.uint8)
        rows, cols = like.shape
        L.ref_random_log_clear()
        has = roi_xywh is not None
        r = np.zeros(4, np.uint32)
        if has:
            x, y, w, h = roi_xywh
            r[:] = [x, y, h, w]  # message order: x_offset, y_offset, height, width
        t = np.ascontiguousarray(ticks, np.int64)
        L.ref_tracker_callback(self.h, like.ctypes.data_as(_u8p), rows, cols, int(has), r.ctypes.data_as(_u32p),
                               t.ctypes.data_as(_i64p), len(t))
        j2 = np.zeros((8, 2))
        tf = np.zeros((10, 3))
        prob = np.zeros((rows, cols), np.uint8)
        ntf = L.ref_tracker_outputs(_p(j2), _p(tf), prob.ctypes.data_as(_u8p), rows, cols)
        logs = []
        for i in range(L.ref_random_log_count()):
            n = L.ref_random_log_get(i, None, 0)
            buf = np.zeros(n)
            L.ref_random_log_get(i, _p(buf), n)
            logs.append(buf)
        cands = None
        if len(logs) == 4:    # first frame after (re)acquiring the face: randu x1, x2, y1, y2
            cands = [np.stack([logs[0], logs[2]]), np.stack([logs[1], logs[3]])]
        elif len(logs) == 2:  # getSamples: interleaved (x, y) per arm
            cands = [np.stack([logs[0][0::2], logs[0][1::2]]), np.stack([logs[1][0::2], logs[1][1::2]])]
        return dict(cands=cands, blurred=prob, joints2d=j2, tf=tf, n_tf=ntf)

    def get_state(self, arm, d=12):
        x = np.zeros((self.N, d))
        P = np.zeros((self.N, d, d))
        self._L.ref_tracker_get_state(self.h, arm, _p(x), _p(P))
        return x, P

    def pose(self, arm, D=22):
        e = np.zeros(D)
        p3 = np.zeros((3, 5))
        self._L.ref_tracker_pose(self.h, arm, _p(e), _p(p3))
        return e, p3

    def __del__(self):
        try:
            if self.h:
                self._L.ref_tracker_destroy(self.h)
                self.h = None
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------------
# the drop-in build: the reference's unmodified src/pfPose.cpp on include/mkf_shims.hpp + libmkf_b200.so
# ------------------------------------------------------------------------------------------------------
SO_DROPIN = os.path.join(_HERE, "_ref", "libref_dropin.so")
_LIB_DROPIN = None


def dropin_available() -> bool:
    if os.path.isdir("/root/reference/src"):
        try:
            subprocess.run(["make", "-C", _HERE, "dropin"], check=True, capture_output=True)
        except Exception:
            pass
    return os.path.exists(SO_DROPIN)


class DropinTracker(RefTracker):
    """the same PFTracker source, but its ParticleFilter / my_gmm / KF_model are the drop-in shims: every filter
    operation runs on the GPU through libmkf_b200.so.  alias_mode: 0 independent slots, 1 the reference binary's
    shallow-copy behaviour (quirk B3)."""

    def __init__(self, *args, alias_mode=1, **kw):
        self._alias_mode = int(alias_mode)
        super().__init__(*args, **kw)

    def _load(self):
        global _LIB_DROPIN
        if _LIB_DROPIN is None:
            if not dropin_available():
                raise FileNotFoundError(SO_DROPIN)
            _LIB_DROPIN = C.CDLL(SO_DROPIN)
        _LIB_DROPIN.ref_dropin_configure.argtypes = [C.c_int]
        _LIB_DROPIN.ref_dropin_configure(self._alias_mode)
        return _LIB_DROPIN


# ---------------------------------------------------------------------------------------------------------------------
# the reference's LEGACY plain particle filter: its own src/pf2D.cpp compiled against the shim (libref_pf2d.so)
# ---------------------------------------------------------------------------------------------------------------------
SO_PF2D = os.path.join(_HERE, "_ref", "libref_pf2d.so")
_LIB_PF2D = None


def pf2d_available() -> bool:
    available()  # `make ref` builds both libraries when the reference sources are present
    return os.path.exists(SO_PF2D)


def lib_pf2d():
    global _LIB_PF2D
    if _LIB_PF2D is None:
        if not pf2d_available():
            raise FileNotFoundError(SO_PF2D)
        L = C.CDLL(SO_PF2D)
        L.refpf_create.restype = C.c_void_p
        L.refpf_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64]
        L.refpf_destroy.argtypes = [C.c_void_p]
        L.refpf_load_gaussian.argtypes = [C.c_void_p, _dp, _dp, C.c_double]
        L.refpf_get_gmm.argtypes = [C.c_void_p, _dp, _dp]
        L.refpf_set_particles.argtypes = [C.c_void_p, _dp]
        L.refpf_get.argtypes = [C.c_void_p, _dp, _dp]
        L.refpf_update.restype = C.c_double
        L.refpf_update.argtypes = [C.c_void_p, _dp, C.c_uint, _dp, C.POINTER(C.c_int)]
        L.refpf_estimate.argtypes = [C.c_void_p, _dp]
        _LIB_PF2D = L
    return _LIB_PF2D


class RefPf2d:
    """the reference's own legacy ParticleFilter(numParticles, numDims, side1) of src/pf2D.{h,cpp}"""

    def __init__(self, N, d, side, means, covs, weights, rng_seed=1):
        self.N, self.d, self.K = N, d, len(weights)
        self.h = lib_pf2d().refpf_create(N, d, int(side), int(rng_seed))
        for k in range(self.K):
            m, c = _f(means[k]), _f(covs[k])
            lib_pf2d().refpf_load_gaussian(self.h, _p(m), _p(c), float(weights[k]))

    def gmm(self):
        si = np.zeros((self.K, self.d, self.d))
        ds = np.zeros(self.K)
        lib_pf2d().refpf_get_gmm(self.h, _p(si), _p(ds))
        return si, ds

    def set_particles(self, x):
        x = _f(x)
        lib_pf2d().refpf_set_particles(self.h, _p(x))

    def get(self):
        x = np.zeros((self.N, self.d))
        w = np.zeros(self.N)
        lib_pf2d().refpf_get(self.h, _p(x), _p(w))
        return x, w

    def update(self, meas, srand_seed):
        """returns (u drawn by resample(), noise added by predict() (N x d), degenerate flag)"""
        meas = _f(meas)
        noise = np.zeros((self.N, self.d))
        deg = C.c_int(0)
        u = lib_pf2d().refpf_update(self.h, _p(meas), int(srand_seed), _p(noise), C.byref(deg))
        return u, noise, bool(deg.value)

    def estimate(self):
        e = np.zeros(self.d)
        lib_pf2d().refpf_estimate(self.h, _p(e))
        return e

    def __del__(self):
        try:
            if self.h:
                lib_pf2d().refpf_destroy(self.h)
                self.h = None
        except Exception:
            pass
