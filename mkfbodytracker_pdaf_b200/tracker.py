"""Thin host-side binding of the C ABI (include/mkf_b200.h) for the test-suite and bench.py.

The reference's own host interface is C++ (`KF_model`, `my_gmm`, `state_params`,
`ParticleFilter`, src/*.h); its drop-in mirror is include/mkf_shims.hpp.  This module only
wraps the same C entry points for Python callers: numpy arrays are host buffers, torch CUDA
tensors (or raw integer device addresses) are device buffers.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L

MODEL_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "models")
LEFT_ARM_MODEL = os.path.join(MODEL_DIR, "data13D_PCA_100000_15_12.yml")
RIGHT_ARM_MODEL = os.path.join(MODEL_DIR, "data23D_PCA_100000_15_12.yml")


def _addr(a):
    """(address, mem) of a numpy array (host), torch tensor (host/device) or None."""
    if a is None:
        return None, L.MEM_HOST
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data, L.MEM_HOST
    if hasattr(a, "data_ptr"):  # torch tensor
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return a.data_ptr(), (L.MEM_DEVICE if a.is_cuda else L.MEM_HOST)
    raise TypeError(f"unsupported buffer type {type(a)}")


def _same_mem(*arrs):
    mems = {(_addr(a)[1]) for a in arrs if a is not None}
    if len(mems) > 1:
        raise ValueError("all buffers of one call must live in the same memory space")
    return mems.pop() if mems else L.MEM_HOST


def _h(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Model:
    """One arm model: GMM prior + the per-component KF_model constants (src/my_gmm.cpp:45-75)."""

    def __init__(self, handle):
        self._h = handle
        K, d, D = C.c_int(), C.c_int(), C.c_int()
        L.check(L.lib.mkf_model_dims(self._h, C.byref(K), C.byref(d), C.byref(D)))
        self.K, self.d, self.D = K.value, d.value, D.value

    @classmethod
    def from_arrays(cls, means, covs, weights, gamma, pca_proj, pca_mean, params: L.Params | None = None):
        means = _h(means, np.float64)
        K, d = means.shape
        covs = _h(covs, np.float64).reshape(K * d, d)
        weights = _h(weights, np.float64).reshape(-1)
        gamma = _h(gamma, np.float64).reshape(-1)
        pca_proj = _h(pca_proj, np.float64)
        pca_mean = _h(pca_mean, np.float64).reshape(-1)
        D = pca_proj.shape[1]
        if pca_proj.shape[0] != d or pca_mean.size != D or weights.size != K or gamma.size < K:
            raise ValueError("inconsistent model array shapes")
        h = C.c_void_p()
        L.check(L.lib.mkf_model_create(C.byref(h), K, d, D, means.ctypes.data, covs.ctypes.data, weights.ctypes.data,
                                       gamma.ctypes.data, pca_proj.ctypes.data, pca_mean.ctypes.data,
                                       C.byref(params) if params is not None else None))
        return cls(h)

    @classmethod
    def load(cls, path, gamma_path=None, params: L.Params | None = None):
        """cv::FileStorage-style load (src/pfPose.cpp:34-55).  gamma_path reproduces quirk B4."""
        h = C.c_void_p()
        L.check(L.lib.mkf_model_load_yaml(C.byref(h), os.fsencode(path),
                                          os.fsencode(gamma_path) if gamma_path else None,
                                          C.byref(params) if params is not None else None))
        return cls(h)

    def save(self, path):
        """writes the model back in the reference's OpenCV-YAML-1.0 schema (src/pfPose.cpp:34-55 reads it)"""
        L.check(L.lib.mkf_model_save_yaml(self._h, os.fsencode(path)))

    def arrays(self):
        K, d, D = self.K, self.d, self.D
        out = dict(means=np.zeros((K, d)), covs=np.zeros((K, d, d)), weights=np.zeros(K), gamma=np.zeros(K),
                   pca_proj=np.zeros((d, D)), pca_mean=np.zeros(D), Q=np.zeros((K, d, d)), B=np.zeros((K, d)),
                   H=np.zeros((6, d)), BH=np.zeros(6))
        order = ["means", "covs", "weights", "gamma", "pca_proj", "pca_mean", "Q", "B", "H", "BH"]
        L.check(L.lib.mkf_model_get(self._h, *[out[k].ctypes.data for k in order]))
        return out

    def close(self):
        if self._h:
            L.lib.mkf_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TrackBatch:
    """T independent ParticleFilter instances (src/pf2DRao.h:13-31) of N slots on one GPU."""

    def __init__(self, model: Model, T: int, N: int, device: int = 0, stream: int | None = None):
        self.model, self.T, self.N, self.device = model, int(T), int(N), device
        h = C.c_void_p()
        L.check(L.lib.mkf_batch_create(C.byref(h), model._h, self.T, self.N, device, stream))
        self._h = h

    def reset(self, u_init):
        p, mem = _addr(u_init)
        L.check(L.lib.mkf_batch_reset(self._h, p, mem))

    def update(self, meas, u_ind, u_post, seeds=None, layout=None, mem=None):
        """ParticleFilter::update for all tracks.  meas: (T,6) shared or (T,6,N) per-slot."""
        if layout is None:
            layout = L.MEAS_SHARED if len(meas.shape) == 2 else L.MEAS_PER_SLOT
        mem = _same_mem(meas, u_ind, u_post, seeds) if mem is None else mem
        L.check(L.lib.mkf_batch_update(self._h, _addr(meas)[0], layout, _addr(u_ind)[0], _addr(u_post)[0],
                                       _addr(seeds)[0], mem))

    def estimate(self):
        xbar = np.zeros((self.T, self.model.d))
        pose = np.zeros((self.T, self.model.D))
        L.check(L.lib.mkf_batch_estimate(self._h, xbar.ctypes.data, pose.ctypes.data, L.MEM_HOST))
        return xbar, pose

    def estimate_into(self, xbar, pose, mem=None):
        mem = _same_mem(xbar, pose) if mem is None else mem
        L.check(L.lib.mkf_batch_estimate(self._h, _addr(xbar)[0], _addr(pose)[0], mem))

    def download(self, state=True, cov=True):
        T, N, d = self.T, self.N, self.model.d
        out = dict(w_raw=np.zeros((T, N)), w_norm=np.zeros((T, N)), indicators=np.zeros((T, N), np.int32),
                   parents=np.zeros((T, N), np.int32), wsum=np.zeros(T), status=np.zeros(T, np.uint32))
        x = np.zeros((T, N, d)) if state else None
        P = np.zeros((T, N, d, d)) if cov else None
        L.check(L.lib.mkf_batch_download(self._h, _addr(x)[0], _addr(P)[0], out["w_raw"].ctypes.data,
                                         out["w_norm"].ctypes.data, out["indicators"].ctypes.data,
                                         out["parents"].ctypes.data, out["wsum"].ctypes.data,
                                         out["status"].ctypes.data, L.MEM_HOST))
        out["x"], out["P"] = x, P
        return out

    def status(self):
        st = np.zeros(self.T, np.uint32)
        L.check(L.lib.mkf_batch_download(self._h, None, None, None, None, None, None, None, st.ctypes.data,
                                         L.MEM_HOST))
        return st

    def summary_into(self, wsum, status):
        """per-track wsum (float64) and status (32-bit) of the last update into caller buffers"""
        mem = _same_mem(wsum, status)
        L.check(L.lib.mkf_batch_download(self._h, None, None, None, None, None, None, _addr(wsum)[0],
                                         _addr(status)[0], mem))

    def upload(self, x, P):
        x = _h(x, np.float64).reshape(self.T, self.N, self.model.d)
        P = _h(P, np.float64).reshape(self.T, self.N, self.model.d, self.model.d)
        L.check(L.lib.mkf_batch_upload(self._h, x.ctypes.data, P.ctypes.data, L.MEM_HOST))

    def synth_fill(self, seed, track0, frame, jitter, layout, meas_dev, u_ind_dev, u_post_dev):
        L.check(L.lib.mkf_synth_fill(self._h, int(seed), int(track0), int(frame), int(jitter), layout,
                                     _addr(meas_dev)[0], _addr(u_ind_dev)[0], _addr(u_post_dev)[0]))

    def pose3d(self, Kcam=None):
        """PFTracker::get3Dpose (src/pfPose.cpp:93-127) of every track: (T, 3, 5)"""
        out = np.zeros((self.T, 3, 5))
        k = _h(Kcam, np.float64).reshape(9) if Kcam is not None else None
        L.check(L.lib.mkf_batch_pose3d(self._h, _addr(k)[0], out.ctypes.data, L.MEM_HOST))
        return out

    def sync(self):
        L.check(L.lib.mkf_batch_sync(self._h))

    def shared_records(self):
        """(records, slots): distinct Gaussians stored by the last update for how many slots (record sharing)"""
        r, n = C.c_int64(), C.c_int64()
        L.check(L.lib.mkf_batch_shared_records(self._h, C.byref(r), C.byref(n)))
        return r.value, n.value

    def heads_kernel(self):
        """name of the kernel that ran the distinct Gaussians of the last run-length frame ('' before the first one)"""
        buf = C.create_string_buffer(64)
        L.check(L.lib.mkf_batch_heads_kernel(self._h, buf, 64))
        return buf.value.decode()

    def join(self):
        """order the batch's stream after all MEM_HOST_ASYNC copies issued so far (no host synchronisation)"""
        L.check(L.lib.mkf_batch_join(self._h))

    def profile(self, max_updates, every=1):
        """arm per-stage CUDA-event timing of update(): up to max_updates samples, one every `every` updates"""
        L.check(L.lib.mkf_batch_profile_every(self._h, int(max_updates), int(every)))

    def profile_read(self):
        a, b_, c, n = C.c_double(), C.c_double(), C.c_double(), C.c_int()
        L.check(L.lib.mkf_batch_profile_read(self._h, C.byref(a), C.byref(b_), C.byref(c), C.byref(n)))
        return dict(ms_bounds=a.value, ms_slot_update=b_.value, ms_resample=c.value, n=n.value)

    def profile_read_stages(self):
        """summed milliseconds per kernel of the profiled updates: bounds, share keys, slot kernel, repair, resample
        (+ ms_slot_span / n_span: the slot kernel's own device span, %globaltimer)"""
        sp, ns = C.c_double(), C.c_int()
        L.check(L.lib.mkf_batch_profile_read_slot_span(self._h, C.byref(sp), C.byref(ns)))
        out = self._profile_read_stages()
        out["ms_slot_span"] = sp.value
        out["n_span"] = ns.value
        return out

    def _profile_read_stages(self):
        ms, n = (C.c_double * 5)(), C.c_int()
        L.check(L.lib.mkf_batch_profile_read_stages(self._h, ms, C.byref(n)))
        return dict(ms_bounds=ms[0], ms_share_keys=ms[1], ms_slot_kernel=ms[2], ms_repair=ms[3], ms_resample=ms[4],
                    n=n.value)

    def close(self):
        if self._h:
            L.lib.mkf_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def associate(arm0: TrackBatch, arm1: TrackBatch, cand_xy, cand_L, roi, u_cand, u_ind, u_post, seeds=None,
              do_update=True):
    """Association step of PFTracker::getMeasurementProposal (src/pfPose.cpp:238-326) for T persons."""
    C_ = cand_xy.shape[-1]
    mem = _same_mem(cand_xy, cand_L, roi, u_cand, u_ind, u_post, seeds)
    L.check(L.lib.mkf_batch_associate(arm0._h, arm1._h, C_, _addr(cand_xy)[0], _addr(cand_L)[0], _addr(roi)[0],
                                      _addr(u_cand)[0], _addr(u_ind)[0], _addr(u_post)[0], _addr(seeds)[0],
                                      1 if do_update else 0, mem))


def assoc_results(arm0: TrackBatch, C_: int):
    T, N = arm0.T, arm0.N
    gate = np.zeros((T, 2, C_), np.uint8)
    w = np.zeros((T, 2, C_))
    bins = np.zeros((T, 2, N), np.int32)
    L.check(L.lib.mkf_batch_assoc_results(arm0._h, gate.ctypes.data, w.ctypes.data, bins.ctypes.data, L.MEM_HOST))
    return dict(gate=gate, weights=w, bins=bins)


def propose(arm0: TrackBatch, arm1: TrackBatch, C_: int, roi, tracking=None, like=None, seed=0, frame=0, track0=0,
            cand_xy=None, cand_L=None):
    """candidate generation front-end (src/pfPose.cpp:216-236, src/pf2DRao.cpp:85-103): returns (cand_xy, cand_L)"""
    T = arm0.T
    if cand_xy is None:
        cand_xy = np.zeros((T, 2, 2, C_))
        cand_L = np.zeros((T, 2, C_), np.uint8) if like is not None else None
    n_img = 1 if like is None or len(like.shape) == 2 else like.shape[0]
    mem = _same_mem(roi, tracking, like, cand_xy, cand_L)
    L.check(L.lib.mkf_batch_propose(arm0._h, arm1._h, C_, _addr(roi)[0], _addr(tracking)[0], _addr(like)[0], n_img,
                                    int(seed), int(frame), int(track0), _addr(cand_xy)[0], _addr(cand_L)[0], mem))
    return cand_xy, cand_L


def skeleton(arm0: TrackBatch, arm1: TrackBatch, Kcam=None):
    """publishTFtree translations (+ camera Euler triple) and publish2Dpos joints (src/pfPose.cpp:129-208)"""
    tf = np.zeros((arm0.T, 10, 3))
    j2 = np.zeros((arm0.T, 8, 2))
    k = _h(Kcam, np.float64).reshape(9) if Kcam is not None else None
    L.check(L.lib.mkf_batch_skeleton(arm0._h, arm1._h, _addr(k)[0], tf.ctypes.data, j2.ctypes.data, L.MEM_HOST))
    return tf, j2


def load_camera_matrix(path):
    K = np.zeros(9)
    L.check(L.lib.mkf_load_camera_matrix(os.fsencode(path), K.ctypes.data))
    return K.reshape(3, 3)


def resample(w, N, u=-1.0, seed=1, device=0):
    """ParticleFilter::resample (src/pf2DRao.cpp:175-210) on the device."""
    w = _h(w, np.float64)
    out = np.zeros(N, np.int32)
    rc = L.check(L.lib.mkf_resample(w.ctypes.data, w.size, N, float(u), int(seed), out.ctypes.data, device))
    return out, rc


class Pf2dBatch:
    """T legacy plain particle filters (src/pf2D.{h,cpp})."""

    def __init__(self, T, N, means, covs, weights, device=0, stream=None):
        means = _h(means, np.float64)
        self.K, self.d = means.shape
        self.T, self.N = int(T), int(N)
        covs = _h(covs, np.float64)
        weights = _h(weights, np.float64)
        h = C.c_void_p()
        L.check(L.lib.mkf_pf2d_create(C.byref(h), self.T, self.N, self.d, self.K, means.ctypes.data, covs.ctypes.data,
                                      weights.ctypes.data, device, stream))
        self._h = h

    def set_particles(self, p):
        L.check(L.lib.mkf_pf2d_set_particles(self._h, _addr(p)[0], _addr(p)[1]))

    def set_random(self, seed, track0=0, side=None, im_w=640, im_h=480):
        """parameters of the constructor / degenerate-branch randomisation (src/pf2D.cpp:44-71,232-250);
        side: T flags (uint8) or None"""
        sp = None
        if side is not None:
            side = _h(side, np.uint8)
            assert side.size == self.T
            sp = side.ctypes.data
        L.check(L.lib.mkf_pf2d_set_random(self._h, int(seed), int(track0), sp, int(im_w), int(im_h)))

    def randomise(self):
        """the constructor's draw (src/pf2D.cpp:44-71): particles across the image, weights 1/N"""
        L.check(L.lib.mkf_pf2d_randomise(self._h))

    def profile(self, max_updates):
        L.check(L.lib.mkf_pf2d_profile(self._h, int(max_updates)))

    def profile_read(self):
        ms = (C.c_double * 3)()
        n = C.c_int(0)
        L.check(L.lib.mkf_pf2d_profile_read(self._h, ms, C.byref(n)))
        return dict(ms_weights=ms[0], ms_resample=ms[1], ms_predict=ms[2], n=n.value)

    def update(self, meas, u, noise=None):
        mem = _same_mem(meas, u, noise)
        L.check(L.lib.mkf_pf2d_update(self._h, _addr(meas)[0], _addr(u)[0], _addr(noise)[0], mem))

    def get(self):
        p = np.zeros((self.T, self.N, self.d))
        w = np.zeros((self.T, self.N))
        par = np.zeros((self.T, self.N), np.int32)
        L.check(L.lib.mkf_pf2d_get(self._h, p.ctypes.data, w.ctypes.data, par.ctypes.data, L.MEM_HOST))
        return p, w, par

    def estimate(self):
        """legacy getEstimator (src/pf2D.cpp:79-88): sum_i weights[i] * particles.row(i) per filter, T x d"""
        est = np.zeros((self.T, self.d))
        L.check(L.lib.mkf_pf2d_estimate(self._h, est.ctypes.data, L.MEM_HOST))
        return est

    def sync(self):
        L.check(L.lib.mkf_pf2d_sync(self._h))

    def close(self):
        if self._h:
            L.lib.mkf_pf2d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
