"""ctypes loader of libmkf_b200.so (the C ABI of include/mkf_b200.h).

There is no Python or CPU fallback: if the shared library is missing this module raises at
import time, and every compute entry point returns MKF_E_CUDA without a device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MKF_LIB_VARIANT=cw4 loads an experiment build of the SAME sources (csrc/Makefile `variant`) for A/B runs
_VARIANT = os.environ.get("MKF_LIB_VARIANT", "")
LIB_PATH = os.path.join(_HERE, "libmkf_b200%s.so" % ("_" + _VARIANT if _VARIANT else ""))

OK = 0
E_INVALID, E_CUDA, E_NOMEM, E_IO, E_PARSE, E_UNSUPPORTED = -1, -2, -3, -4, -5, -6
CHOL_CV24_LITERAL, CHOL_CV3_LITERAL, CHOL_EXACT = 0, 1, 2
ALIAS_INDEPENDENT, ALIAS_CV_SHALLOW_LITERAL = 0, 1
MEM_AUTO, MEM_HOST, MEM_DEVICE, MEM_HOST_ASYNC = 0, 1, 2, 3
MEAS_SHARED, MEAS_PER_SLOT = 0, 1
ST_IND_FALLBACK, ST_POST_FALLBACK, ST_POST_DEGENERATE, ST_CHOL_FAIL = 0x1, 0x2, 0x4, 0x8
ST_CAND_FALLBACK, ST_CAND_DEGENERATE, ST_IND_WRAP = 0x10, 0x20, 0x40

# every symbol include/mkf_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "mkf_params_default", "mkf_last_error", "mkf_abi_version", "mkf_device_count", "mkf_model_create",
    "mkf_model_load_yaml", "mkf_model_save_yaml", "mkf_model_destroy", "mkf_model_dims", "mkf_model_get", "mkf_batch_create",
    "mkf_batch_destroy", "mkf_batch_sync", "mkf_batch_reset", "mkf_batch_update", "mkf_batch_estimate",
    "mkf_batch_associate", "mkf_batch_assoc_results", "mkf_batch_download", "mkf_batch_upload", "mkf_resample",
    "mkf_pf2d_create", "mkf_pf2d_destroy", "mkf_pf2d_set_particles", "mkf_pf2d_get", "mkf_pf2d_update",
    "mkf_pf2d_sync", "mkf_pf2d_estimate", "mkf_pf2d_set_random", "mkf_pf2d_randomise", "mkf_pf2d_profile", "mkf_pf2d_profile_read", "mkf_synth_fill", "mkf_launch_count", "mkf_batch_join", "mkf_batch_shared_records", "mkf_batch_heads_kernel", "mkf_batch_profile", "mkf_batch_profile_every", "mkf_batch_profile_read", "mkf_batch_profile_read_stages", "mkf_batch_profile_read_slot_span", "mkf_kf_apply", "mkf_batch_sample_prob", "mkf_batch_pose3d", "mkf_batch_skeleton", "mkf_load_camera_matrix", "mkf_batch_propose",
    "mkf_shard_tracks", "mkf_comm_unique_id", "mkf_comm_create", "mkf_comm_wrap", "mkf_comm_destroy", "mkf_comm_info",
    "mkf_batch_summaries", "mkf_batch_gather_summaries",
]


class Params(C.Structure):
    _fields_ = [
        ("chol_mode", C.c_int),
        ("alias_mode", C.c_int),
        ("meas_noise_var", C.c_double),
        ("assoc_pa", C.c_double),
        ("assoc_clutter", C.c_double),
        ("proposal_spread", C.c_double),
        ("neck_offset", C.c_double),
        ("img_rows", C.c_int),
        ("img_cols", C.c_int),
    ]


class MkfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmkf_b200 error {code}: {msg}")
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
        "`make -C mkfbodytracker_pdaf_b200/csrc` (there is no fallback implementation)")

lib = C.CDLL(LIB_PATH)

_vp, _dp, _ip, _bp = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p  # raw addresses: host or device

lib.mkf_params_default.argtypes = [C.POINTER(Params)]
lib.mkf_params_default.restype = None
lib.mkf_last_error.restype = C.c_char_p
lib.mkf_abi_version.restype = C.c_int
lib.mkf_device_count.restype = C.c_int
lib.mkf_model_create.argtypes = [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp,
                                 C.POINTER(Params)]
lib.mkf_model_load_yaml.argtypes = [C.POINTER(_vp), C.c_char_p, C.c_char_p, C.POINTER(Params)]
lib.mkf_model_save_yaml.argtypes = [_vp, C.c_char_p]
lib.mkf_model_destroy.argtypes = [_vp]
lib.mkf_model_destroy.restype = None
lib.mkf_model_dims.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
lib.mkf_model_get.argtypes = [_vp] + [_dp] * 10
lib.mkf_batch_create.argtypes = [C.POINTER(_vp), _vp, C.c_int64, C.c_int, C.c_int, _vp]
lib.mkf_batch_destroy.argtypes = [_vp]
lib.mkf_batch_destroy.restype = None
lib.mkf_batch_sync.argtypes = [_vp]
lib.mkf_batch_reset.argtypes = [_vp, _dp, C.c_int]
lib.mkf_batch_update.argtypes = [_vp, _dp, C.c_int, _dp, _dp, _vp, C.c_int]
lib.mkf_batch_estimate.argtypes = [_vp, _dp, _dp, C.c_int]
lib.mkf_batch_associate.argtypes = [_vp, _vp, C.c_int, _dp, _bp, _dp, _dp, _dp, _dp, _vp, C.c_int, C.c_int]
lib.mkf_batch_assoc_results.argtypes = [_vp, _bp, _dp, _ip, C.c_int]
lib.mkf_batch_download.argtypes = [_vp, _dp, _dp, _dp, _dp, _ip, _ip, _dp, _vp, C.c_int]
lib.mkf_batch_upload.argtypes = [_vp, _dp, _dp, C.c_int]
lib.mkf_resample.argtypes = [_dp, C.c_int, C.c_int, C.c_double, C.c_uint64, _ip, C.c_int]
lib.mkf_pf2d_create.argtypes = [C.POINTER(_vp), C.c_int64, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _vp]
lib.mkf_pf2d_destroy.argtypes = [_vp]
lib.mkf_pf2d_destroy.restype = None
lib.mkf_pf2d_set_particles.argtypes = [_vp, _dp, C.c_int]
lib.mkf_pf2d_get.argtypes = [_vp, _dp, _dp, _ip, C.c_int]
lib.mkf_pf2d_update.argtypes = [_vp, _dp, _dp, _dp, C.c_int]
lib.mkf_pf2d_sync.argtypes = [_vp]
lib.mkf_pf2d_set_random.argtypes = [_vp, C.c_uint64, C.c_int64, _vp, C.c_int, C.c_int]
lib.mkf_pf2d_randomise.argtypes = [_vp]
lib.mkf_pf2d_profile.argtypes = [_vp, C.c_int]
lib.mkf_pf2d_profile_read.argtypes = [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]
lib.mkf_pf2d_estimate.argtypes = [_vp, _dp, C.c_int]
lib.mkf_synth_fill.argtypes = [_vp, C.c_uint64, C.c_int64, C.c_uint64, C.c_int, C.c_int, _dp, _dp, _dp]
lib.mkf_launch_count.restype = C.c_uint64
lib.mkf_kf_apply.argtypes = [_vp, C.c_int, _ip, C.c_int, _dp, _dp, _dp, _dp, C.c_int]
lib.mkf_batch_sample_prob.argtypes = [_vp, C.c_int64, _dp, C.c_int, C.c_double, _dp]
lib.mkf_batch_propose.argtypes = [_vp, _vp, C.c_int, _dp, _bp, _bp, C.c_int, C.c_uint64, C.c_uint64, C.c_int64, _dp, _bp,
                                  C.c_int]
lib.mkf_batch_pose3d.argtypes = [_vp, _dp, _dp, C.c_int]
lib.mkf_batch_skeleton.argtypes = [_vp, _vp, _dp, _dp, _dp, C.c_int]
lib.mkf_load_camera_matrix.argtypes = [C.c_char_p, _dp]
lib.mkf_batch_join.argtypes = [_vp]
lib.mkf_batch_shared_records.argtypes = [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
lib.mkf_batch_heads_kernel.argtypes = [_vp, C.c_char_p, C.c_int]
lib.mkf_batch_profile.argtypes = [_vp, C.c_int]
lib.mkf_batch_profile_every.argtypes = [_vp, C.c_int, C.c_int]
lib.mkf_batch_profile_read_stages.argtypes = [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]
lib.mkf_batch_profile_read_slot_span.argtypes = [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]
lib.mkf_batch_profile_read.argtypes = [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_int)]
lib.mkf_shard_tracks.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
lib.mkf_comm_unique_id.argtypes = [_vp]
lib.mkf_comm_create.argtypes = [C.POINTER(_vp), C.c_int, C.c_int, _vp, C.c_int]
lib.mkf_comm_wrap.argtypes = [C.POINTER(_vp), _vp, C.c_int]
lib.mkf_comm_destroy.argtypes = [_vp]
lib.mkf_comm_destroy.restype = None
lib.mkf_comm_info.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
lib.mkf_batch_summaries.argtypes = [_vp, C.c_int64, _dp, C.c_int]
lib.mkf_batch_gather_summaries.argtypes = [_vp, _vp, C.c_int64, _dp, C.c_int]
COMM_ID_BYTES = 128


def check(rc: int) -> int:
    if rc < 0:
        raise MkfError(rc, lib.mkf_last_error().decode(errors="replace"))
    return rc


def default_params() -> Params:
    p = Params()
    lib.mkf_params_default(C.byref(p))
    return p
