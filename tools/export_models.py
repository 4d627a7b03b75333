#!/usr/bin/env python
"""Re-emit the reference's GMM/PCA model files as OpenCV-YAML-1.0 fixtures.

Run in the authoring container only (needs /root/reference and cv2).  The numbers are read with
cv2.FileStorage (the same parser the reference uses, src/pfPose.cpp:34-55) and written with
shortest-round-trip decimal text so that every f64/f32 value is reproduced bit-for-bit; the
output keeps the reference's schema (means, covs, weights, pca_proj[f32], pca_mean[f32], gamma)
so that user files in the original format load unchanged.
"""
import os
import sys

import cv2
import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mkfbodytracker_pdaf_b200", "models")
KEYS = ["means", "covs", "weights", "pca_proj", "pca_mean", "gamma"]


def emit(f, name, a):
    dt = {np.dtype("float64"): "d", np.dtype("float32"): "f"}[a.dtype]
    f.write(f"{name}: !!opencv-matrix\n   rows: {a.shape[0]}\n   cols: {a.shape[1]}\n   dt: {dt}\n   data: [ ")
    flat = a.reshape(-1)
    txt = [repr(float(v)) if dt == "d" else repr(float(np.format_float_scientific(v, unique=True))) for v in flat]
    if dt == "f":
        txt = [np.format_float_scientific(v, unique=True) for v in flat]
    line = ""
    rows = []
    for t in txt:
        if len(line) + len(t) + 2 > 100:
            rows.append(line)
            line = ""
        line += t + ", "
    rows.append(line.rstrip(", "))
    f.write("\n       ".join(rows))
    f.write(" ]\n")


for fn in ("data13D_PCA_100000_15_12.yml", "data23D_PCA_100000_15_12.yml"):
    fs = cv2.FileStorage(os.path.join(REF, fn), cv2.FILE_STORAGE_READ)
    mats = {k: fs.getNode(k).mat() for k in KEYS}
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, fn), "w") as f:
        f.write("%YAML:1.0\n")
        f.write("# GMM/PCA arm model of mgb45/mkfbodytracker_pdaf (MIT licence), values re-emitted bit-exactly\n")
        f.write("# by tools/export_models.py; schema as read by the reference at src/pfPose.cpp:34-55.\n")
        for k in KEYS:
            emit(f, k, mats[k])
    # verify the round trip with the same parser
    fs2 = cv2.FileStorage(os.path.join(OUT, fn), cv2.FILE_STORAGE_READ)
    for k in KEYS:
        b = fs2.getNode(k).mat()
        assert b.dtype == mats[k].dtype and b.shape == mats[k].shape and np.array_equal(b, mats[k]), (fn, k)
    print("wrote", fn)
# the webcam camera matrix (cal.yml:4-7) lives in models/webcam_camera_matrix.yml (hand-written, 9 numbers)
