#!/usr/bin/env python
"""Device timeline of the four kernels of a run-length frame in the pipelined headline loop (no events, no host
synchronisation inside): start / end of k_frame_heads, the heads' slot kernel, k_runs_repair, k_resample_runs from
%globaltimer stamps.  Needs the MKF_TIMELINE variant
(make -C mkfbodytracker_pdaf_b200/csrc variant NAME=timeline DEFS=-DMKF_TIMELINE; MKF_LIB_VARIANT=timeline)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mkfbodytracker_pdaf_b200 as mk

T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N, F = 500, 100
SEED = 0x5EED0002
dev = torch.device("cuda:0")
model = mk.Model.load(mk.LEFT_ARM_MODEL, mk.RIGHT_ARM_MODEL)
stream = torch.cuda.Stream()
batch = mk.TrackBatch(model, T, N, device=0, stream=stream.cuda_stream)
meas = torch.empty((F, T, 6), dtype=torch.float64, device=dev)
ui = torch.empty((F, T), dtype=torch.float64, device=dev)
up = torch.empty((F, T), dtype=torch.float64, device=dev)
for f in range(F):
    batch.synth_fill(SEED, 0, f, 1, mk.MEAS_SHARED, meas[f], ui[f], up[f])
u0 = torch.empty(T, dtype=torch.float64, device=dev)
batch.synth_fill(SEED, 0, 0xFFFFFF, 1, mk.MEAS_SHARED, meas[0].clone(), u0, None)
pose = torch.empty((T, model.D), dtype=torch.float64, device=dev)
batch.reset(u0)
lib = mk._lib.lib
lib.mkf_debug_timeline.argtypes = [C.c_void_p, C.c_int]
lib.mkf_debug_timeline.restype = C.c_int
for f in range(40):
    batch.update(meas[f], ui[f], up[f])
    batch.estimate_into(None, pose)
torch.cuda.synchronize()
buf = np.zeros((64, 4, 2), dtype=np.uint64)
assert lib.mkf_debug_timeline(buf.ctypes.data, 1) == 0, "not a MKF_TIMELINE build"
for f in range(40, F):
    batch.update(meas[f], ui[f], up[f])
    batch.estimate_into(None, pose)
torch.cuda.synchronize()
lib.mkf_debug_timeline(buf.ctypes.data, 0)
# frames 40..99 -> slots (frame & 63); order the 60 frames by the start of their first kernel
if os.environ.get("MKF_LIB_VARIANT") == "tlskew":  # -DMKF_TL_LATEST_START: the slot kernel's start entry is ~(latest CTA start)
    buf[:, 1, 0] = ~buf[:, 1, 0]
fr = [buf[i].astype(np.int64) for i in range(64) if buf[i, 0, 1] > 0 and buf[i, 3, 1] > 0]
fr.sort(key=lambda a: a[0, 0])
fr = fr[5:-2]
names = ["frame_heads", "slot_kernel", "repair", "resample"]
dur = {n: float(np.mean([a[k, 1] - a[k, 0] for a in fr])) / 1e3 for k, n in enumerate(names)}
gap = {f"{names[k]}->{names[k + 1]}": float(np.mean([a[k + 1, 0] - a[k, 1] for a in fr])) / 1e3 for k in range(3)}
gap["resample->next frame_heads"] = float(np.mean([fr[i + 1][0, 0] - fr[i][3, 1] for i in range(len(fr) - 1)])) / 1e3
period = float(np.mean([fr[i + 1][0, 0] - fr[i][0, 0] for i in range(len(fr) - 1)])) / 1e3
print(json.dumps({"variant": os.environ.get("MKF_HEADS_TMA", "1") + "/" + os.environ.get("MKF_HEADS_TMA_CFG", "default"),
                  "tracks": T, "frames": len(fr), "period_us": round(period, 2),
                  "kernel_us": {k: round(v, 2) for k, v in dur.items()}, "gap_us": {k: round(v, 2) for k, v in gap.items()}}))
