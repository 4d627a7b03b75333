#!/usr/bin/env python
"""Host-visible latency of single C-ABI calls at the reference's operating point (1 person: N = 500 slots per arm,
5000 candidates per hand, host buffers) -- what the drop-in shims pay per ParticleFilter method call."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import mkfbodytracker_pdaf_b200 as mk  # noqa: E402
from mkfbodytracker_pdaf_b200 import _lib as L  # noqa: E402
import ctypes as C  # noqa: E402

N, Cn = 500, 5000
rng = np.random.default_rng(0)
left = mk.Model.load(os.path.join(mk.MODEL_DIR, "data13D_PCA_100000_15_12.yml"))
b = mk.TrackBatch(left, 1, N)
b.reset(np.array([0.3]))


def t(fn, n=200, warm=10):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e6


w = rng.random(Cn)
w /= w.sum()
meas = rng.normal(300, 30, (1, 6, N))
cand = rng.uniform(100, 400, (2, Cn))
out = np.zeros(Cn)
res = {}
res["mkf_resample 5000->500"] = t(lambda: mk.resample(w, N, 0.4))
res["mkf_resample 15->500"] = t(lambda: mk.resample(w[:15] / w[:15].sum(), N, 0.4))
res["mkf_batch_update per-slot T=1"] = t(lambda: b.update(meas, np.array([0.2]), np.array([0.7]), layout=mk.MEAS_PER_SLOT))
res["mkf_batch_estimate T=1"] = t(lambda: b.estimate())
res["mkf_batch_sample_prob 5000"] = t(lambda: L.check(L.lib.mkf_batch_sample_prob(b._h, 0, cand.ctypes.data_as(C.POINTER(C.c_double)), Cn, 47.0, out.ctypes.data_as(C.POINTER(C.c_double)))))
res["mkf_batch_download x,P"] = t(lambda: b.download())
print(json.dumps({k: round(v, 1) for k, v in res.items()}))
