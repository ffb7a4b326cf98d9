"""ctypes binding of libpcbridge.so (include/pcbridge.h).

There is no CPU fallback and no other backend: if the library is missing or a call fails the
caller gets an exception.  The library is built in-tree by pointcloud_bridge_b200/build.py
(nvcc, sm_100a) and is never JIT-compiled at import time on the GPU box.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCB_LIB_PATH") or os.path.join(_HERE, "libpcbridge.so")   # override: kernel experiments

_lib = None

_vp = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_f = ctypes.c_float
_d = ctypes.c_double

# name -> argtypes (all return int)
_SIGNATURES = {
    "pcb_fps_f32": [_vp, _i, _i, _vp, _i, _vp, _vp],
    "pcb_square_distance_f32": [_vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "pcb_ball_query_f32": [_vp, _vp, _i, _i, _i, _f, _i, _vp, _vp],
    "pcb_ball_query_multi_f32": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "pcb_gather_f32": [_vp, _vp, _i, _i, _i, _i64, _i, _vp, _vp, _vp],
    "pcb_gather_bwd_f32": [_vp, _vp, _i, _i, _i, _i64, _i, _vp, _vp],
    "pcb_group_points_f32": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "pcb_group_points_bwd_f32": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "pcb_group_points_bf16": [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "pcb_group_points_bwd_bf16": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "pcb_three_nn_f32": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "pcb_interpolate_f32": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "pcb_interpolate_bwd_f32": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "pcb_fp_concat_bf16": [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "pcb_fp_concat_bwd_bf16": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "pcb_wgrad_rows_bf16": [_vp, _vp, _i64, _i, _i, _i, _i, _vp, _i, _vp],
    "pcb_adam_flat_f32": [_vp, _vp, _vp, _vp, _i64, _vp, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp],
    "pcb_linear_rows_bf16": [_vp, _i64, _vp, _i64, _i64, _i, _i, _i, _vp, _i64, _vp],
    "pcb_linear_bias_act_rows_bf16": [_vp, _i64, _vp, _i64, _i64, _i, _i, _i, _vp, _i, _i, _f, _i, _vp, _i64, _vp],
    "pcb_linear_bn_stats_rows_bf16": [_vp, _i64, _vp, _i64, _i64, _i, _i, _i, _vp, _i64, _i, _f, _vp, _vp, _vp, _vp, _vp,
                                      _vp, _vp, _vp],
    "pcb_dgrad_bn_rows_bf16": [_vp, _i64, _vp, _i64, _i64, _i, _i, _i, _vp, _i64, _vp, _vp, _vp, _vp, _i, _i, _vp, _i64,
                               _vp, _vp, _vp, _vp, _vp, _vp],
    "pcb_bn_apply_rows": [_vp, _i, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _i64, _vp, _vp, _vp, _f, _vp, _vp, _vp,
                          _i, _f, _vp, _vp],
    "pcb_bn_pool_bwd_rows": [_vp, _i64, _vp, _vp, _vp, _i, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp],
    "pcb_bn_bwd_apply_rows": [_vp, _vp, _i, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp],
    "pcb_scene_window_count_f32": [_vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _d, _d, _d, _i, _vp, _vp],
    "pcb_scene_window_fill_f32": [_vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _d, _d, _d, _i, _vp, _vp, ctypes.c_uint, _vp, _vp],
    "pcb_scene_blocks_f32": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i64, _i, _d, _d, _d, ctypes.c_uint, _vp, _vp, _vp],
    "pcb_scene_vote": [_vp, _vp, _i64, _i64, _i, _vp, _vp],
    "pcb_scene_vote_argmax": [_vp, _i64, _i, _vp, _vp],
    "pcb_nll_rows_fwd": [_vp, _i, _vp, _vp, _i64, _i, _i, _vp, _vp],
    "pcb_nll_rows_bwd": [_vp, _i, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp],
    "pcb_eigvalsh3_f32": [_vp, _i64, _vp, _vp],
    "pcb_structure_rows_f32": [_vp, _vp, _i, _i, _i, _vp, _i, _f, _i, _i, _vp, _vp, _vp],
    "pcb_knn_f32": [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "pcb_knn_cdist_f32": [_vp, _i, _i, _i, _vp, _vp, _vp],
    "pcb_graph_feature_f32": [_vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "pcb_graph_feature_bwd_f32": [_vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "pcb_bn_fwd_rows": [_vp, _i, _i64, _i, _i, _i, _vp, _vp, _vp, _f, _f, _vp, _vp, _i, _vp, _vp, _vp, _i64, _vp, _vp, _vp],
    "pcb_sa_fused_bf16": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _f, _vp, _i, _vp],
    "pcb_bn_bwd_rows": [_vp, _i64, _vp, _vp, _i, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp],
}

EXPORTS = ["pcb_version", "pcb_error_string", "pcb_bn_work_floats", "pcb_nll_rows_blocks", "pcb_gemm_work_floats",
           "pcb_gemm_tickets", "pcb_gemm_max_groups", "pcb_bn_set_coop_sms", *_SIGNATURES]


class PcbError(RuntimeError):
    pass


def lib():
    """The loaded library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PcbError(
                f"{LIB_PATH} not found: build it with `python -m pointcloud_bridge_b200.build` "
                "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for the hot path.")
        l = ctypes.CDLL(LIB_PATH)
        l.pcb_version.restype = _i
        l.pcb_version.argtypes = []
        l.pcb_error_string.restype = ctypes.c_char_p
        l.pcb_error_string.argtypes = [_i]
        l.pcb_bn_work_floats.restype = _i64
        l.pcb_bn_work_floats.argtypes = [_i]
        l.pcb_nll_rows_blocks.restype = _i
        l.pcb_nll_rows_blocks.argtypes = [_i64]
        l.pcb_gemm_work_floats.restype = _i64
        l.pcb_gemm_work_floats.argtypes = [_i64, _i, _i]
        l.pcb_gemm_tickets.restype = _i
        l.pcb_gemm_tickets.argtypes = []
        l.pcb_gemm_max_groups.restype = _i
        l.pcb_gemm_max_groups.argtypes = []
        l.pcb_bn_set_coop_sms.restype = _i
        l.pcb_bn_set_coop_sms.argtypes = [_i]
        for name, args in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = _i
            fn.argtypes = args
        _lib = l
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().pcb_error_string(code).decode()
        raise PcbError(f"{what} failed: {msg} (code {code})")


# launch counter: how many kernels of ours were enqueued (bench.py reports it as gpu_launches)
_launches = 0


def count_launches(n: int = 1) -> None:
    global _launches
    _launches += n


def launches() -> int:
    return _launches
