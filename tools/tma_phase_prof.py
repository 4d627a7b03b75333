#!/usr/bin/env python
"""Per-phase cycle breakdown of k_slot_update_heads_tma on the headline workload (needs the MKF_TMA_PROF variant:
make -C mkfbodytracker_pdaf_b200/csrc variant NAME=tmaprof DEFS=-DMKF_TMA_PROF; run with MKF_LIB_VARIANT=tmaprof)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mkfbodytracker_pdaf_b200 as mk

T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N, F = 500, 60
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
batch.reset(u0)
lib = mk._lib.lib
lib.mkf_debug_tma_prof.argtypes = [C.c_void_p, C.c_int]
lib.mkf_debug_tma_prof.restype = C.c_int
for f in range(30):
    batch.update(meas[f], ui[f], up[f])
torch.cuda.synchronize()
out = (C.c_ulonglong * 8)()
assert lib.mkf_debug_tma_prof(out, 1) == 0, "not a MKF_TMA_PROF build"
for f in range(30, F):
    batch.update(meas[f], ui[f], up[f])
torch.cuda.synchronize()
lib.mkf_debug_tma_prof(out, 0)
v = list(out)
steps = max(v[7], 1)
names = ["wait input stage", "next fetch issued", "arithmetic", "wait output stage", "regs->stage + stores",
         "wait stores read + loop end", "stage->regs"]
print(json.dumps({"cfg": os.environ.get("MKF_HEADS_TMA_CFG", "default"), "tracks": T, "warp_steps": steps,
                  "cycles_per_step": {n: round(v[i] / steps) for i, n in enumerate(names)},
                  "total": round(sum(v[:7]) / steps)}))
