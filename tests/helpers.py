"""shared helpers of the parity tests: synthetic inputs through the oracle's copy of
include/mkf_synth.h, and comparison metrics with the tolerances of BASELINE.json:north_star."""
import numpy as np

import mkf_oracle as orc

RTOL = 1e-4  # north_star: state means, covariances and weights within 1e-4 relative

LANE_U_IND, LANE_U_POST, LANE_U_INIT, LANE_U_CAND = 0x1001, 0x1002, 0x1003, 0x1004
NO_FRAME = 0xFFFFFFFFFFFF


def synth_frame(seed, tracks, frame, N=None, jitter=1):
    """(meas, u_ind, u_post): meas (T,6) if N is None else (T,6,N)"""
    T = len(tracks)
    if N is None:
        meas = np.stack([orc.synth_meas(seed, t, frame, -1, jitter) for t in tracks])
    else:
        meas = np.zeros((T, 6, N))
        for i, t in enumerate(tracks):
            for j in range(N):
                meas[i, :, j] = orc.synth_meas(seed, t, frame, j, jitter)
    ui = np.array([orc.synth_u(seed, t, frame, LANE_U_IND) for t in tracks])
    up = np.array([orc.synth_u(seed, t, frame, LANE_U_POST) for t in tracks])
    return meas, ui, up


def synth_u_init(seed, tracks):
    return np.array([orc.synth_u(seed, t, NO_FRAME, LANE_U_INIT) for t in tracks])


def rel_err(a, b):
    """max |a-b| relative to the scale of b (per leading index), the metric the 1e-4 bound is applied to"""
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    if a.size == 0:
        return 0.0
    scale = np.abs(b).reshape(b.shape[0], -1).max(axis=1) if b.ndim > 1 else np.abs(b).max()
    scale = np.maximum(scale, 1e-300)
    diff = np.abs(a - b).reshape(b.shape[0], -1).max(axis=1) if b.ndim > 1 else np.abs(a - b).max()
    return float(np.max(diff / scale))


def rel_err_weights(a, b):
    """element-wise relative error of weights (only where the reference weight is not denormal-tiny)"""
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    m = b > 1e-290
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m]) / b[m]))
