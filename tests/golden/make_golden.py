#!/usr/bin/env python
"""Generates the committed golden fixtures from the CPU oracle (oracle/mkf_oracle.cpp).

  synth_pins.npz     fixed outputs of include/mkf_synth.h (shared CPU/GPU input generator)
  config1_left.npz   BASELINE.json config 1: data13D_PCA_100000_15_12.yml (gamma from the 23D file,
                     quirk B4), 1 track, N=500, 300 frames, per-slot measurement columns, seed
                     0x5EED0001, CV24_LITERAL / INDEPENDENT: per-frame pose, wsum, index checksums,
                     a few full frames and the final state.

The reference itself cannot run here (ROS + OpenCV C++ absent), so these vectors come from the
restatement: they pin regressions of the oracle and give the GPU tests a fixed target.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import mkf_oracle as orc  # noqa: E402
import mkfbodytracker_pdaf_b200 as mk  # noqa: E402
from helpers import synth_frame  # noqa: E402


def index_checksum(idx):
    idx = np.asarray(idx, np.int64)
    return int(((idx + 1) * (np.arange(idx.size, dtype=np.int64) * 2654435761 % 1000003 + 1)).sum() % (2**61 - 1))


def main():
    np.savez(os.path.join(HERE, "synth_pins.npz"),
             meas=np.array([orc.synth_meas(0x5EED0002, t, f, j, 1) for t in (0, 77) for f in (0, 5) for j in (-1, 3)]),
             u=np.array([orc.synth_u(0x5EED0002, t, 3, w) for t in (0, 77) for w in (0x1001, 0x1002)]),
             cand=np.array([orc.synth_candidate(0x5EED0003, 4, 2, h, 17, c) for h in (0, 1) for c in (0, 1, 16)]))
    m = mk.Model.load(mk.LEFT_ARM_MODEL, mk.RIGHT_ARM_MODEL)
    a = m.arrays()
    om = orc.Model(a["means"], a["covs"], a["weights"], a["gamma"], a["pca_proj"], a["pca_mean"])
    seed, N, frames = 0x5EED0001, 500, 300
    f = orc.Filter(om, N)
    f.reset(u=orc.synth_u(seed, 0, 0xFFFFFFFFFFFF, 0x1003))
    pose = np.zeros((frames, 22))
    wsum = np.zeros(frames)
    par_ck = np.zeros(frames, np.int64)
    ind_ck = np.zeros(frames, np.int64)
    keep = {}
    for fr in range(frames):
        meas, ui, up = synth_frame(seed, [0], fr, N, jitter=0)
        r = f.update(meas[0], ui[0], up[0])
        assert r["status"] == 0
        _, pose[fr] = f.estimate()
        wsum[fr] = r["wsum"]
        par_ck[fr] = index_checksum(r["parents"])
        ind_ck[fr] = index_checksum(r["indicators"])
        if fr in (0, 1, 2, 150, 299):
            keep[f"parents_{fr}"] = r["parents"]
            keep[f"indicators_{fr}"] = r["indicators"]
            keep[f"w_norm_{fr}"] = r["w_norm"]
    x, P = f.get_state()
    np.savez_compressed(os.path.join(HERE, "config1_left.npz"), pose=pose, wsum=wsum, par_ck=par_ck, ind_ck=ind_ck,
                        x_final=x, trP_final=np.trace(P, axis1=1, axis2=2), **keep)
    print("golden written; final hand estimate", pose[-1, :2], "wsum[-1]", wsum[-1])


if __name__ == "__main__":
    main()
