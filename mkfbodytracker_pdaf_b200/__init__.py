"""mkfbodytracker_pdaf_b200 -- B200-native (sm_100a) per-frame filtering hot path of
mgb45/mkfbodytracker_pdaf: GMM-KF bank (KF_model/my_gmm), Rao-Blackwellised particle filter
(pf2DRao), sample-based "PDAF" association (pfPose.cpp) and the legacy pf2D weighting.

The product is csrc/ (CUDA kernels + C ABI, include/mkf_b200.h) and include/mkf_shims.hpp
(the reference's C++ class interfaces over that ABI).  The Python modules are bindings only.
"""
from . import _lib
from ._lib import (ALIAS_INDEPENDENT, CHOL_CV24_LITERAL, CHOL_CV3_LITERAL, CHOL_EXACT, MEAS_PER_SLOT, MEAS_SHARED,
                   MEM_AUTO, MEM_DEVICE, MEM_HOST, MEM_HOST_ASYNC, MkfError, Params, default_params)
from .tracker import (LEFT_ARM_MODEL, MODEL_DIR, RIGHT_ARM_MODEL, Model, Pf2dBatch, TrackBatch, assoc_results,
                      associate, load_camera_matrix, propose, resample, skeleton)

__all__ = [
    "Model", "TrackBatch", "Pf2dBatch", "associate", "assoc_results", "resample", "skeleton", "load_camera_matrix", "propose", "Params", "default_params",
    "MkfError", "LEFT_ARM_MODEL", "RIGHT_ARM_MODEL", "MODEL_DIR", "MEAS_SHARED", "MEAS_PER_SLOT", "MEM_AUTO",
    "MEM_HOST", "MEM_DEVICE", "MEM_HOST_ASYNC", "CHOL_CV24_LITERAL", "CHOL_CV3_LITERAL", "CHOL_EXACT", "ALIAS_INDEPENDENT",
]


def device_count() -> int:
    return _lib.lib.mkf_device_count()


def launch_count() -> int:
    return int(_lib.lib.mkf_launch_count())
