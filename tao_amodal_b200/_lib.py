"""ctypes binding of the C ABI in include/ta_eval.h (libta_eval.so, built in-tree by
``__graft_entry__.build()``).

There is deliberately no fallback: if the shared library is missing or a call
fails, the product raises.  The reference has no FFI for this path (its hot
loops are Python methods, tao_amodal/evaluation/tao_amodal/eval.py:306-584),
so these prototypes are the binding a maintainer would add.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# TA_EVAL_LIB selects another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("TA_EVAL_LIB") or os.path.join(_HERE, "csrc", "libta_eval.so")

TA_OK = 0
TA_ERR_INVALID = -1
TA_ERR_CUDA = -2
TA_ERR_TOO_LARGE = -3
TA_ERR_ASSERT = -4
TA_ERR_NCCL = -5

IOU_MODES = {"3d_iou": 0, "avg_iou": 1, "imagenetvid": 2, "3d_iou_seq": 3}

EXPORTS = [
    "ta_abi_version", "ta_last_error", "ta_ctx_create", "ta_ctx_destroy", "ta_ctx_sm_count",
    "ta_ctx_launch_count", "ta_track_iou", "ta_box_iou", "ta_match_greedy", "ta_frame_eval",
    "ta_ctx_timing", "ta_ctx_timing_read", "ta_frame_eval_max_gt", "ta_frame_eval_max_dt", "ta_frame_eval_max_pairs", "ta_pr_accumulate", "ta_eval_plan_host", "ta_eval_plans_host", "ta_host_alloc", "ta_host_free", "ta_widen_u16", "ta_offsets_from_counts", "ta_gather_boxes",
    "ta_peer_window_create", "ta_peer_window_destroy", "ta_peer_window_ptr", "ta_peer_window_set_plan",
    "ta_peer_window_acquire", "ta_peer_window_put", "ta_peer_window_exchange", "ta_peer_window_check",
    "ta_rle_iou", "ta_frame_sched_bytes", "ta_frame_sched_build",
    "ta_exchange_unique_id", "ta_exchange_create", "ta_exchange_destroy", "ta_exchange_rank",
    "ta_exchange_world", "ta_exchange_gather", "ta_exchange_scatter", "ta_exchange_alltoallv",
    "ta_exchange_allreduce_sum", "ta_exchange_group_begin", "ta_exchange_group_end",
    "ta_widen_boxes", "ta_ctx_take_assert_count", "ta_ctx_debug_list_count", "ta_zero",
]


class RangeCfg(C.Structure):
    """struct ta_range_cfg."""
    _fields_ = [
        ("gt_a_lo", C.c_double), ("gt_a_hi", C.c_double),
        ("gt_b_lo", C.c_double), ("gt_b_hi", C.c_double),
        ("dt_a_lo", C.c_double), ("dt_a_hi", C.c_double),
        ("dt_b_lo", C.c_double), ("dt_b_hi", C.c_double),
        ("gt_hp_min", C.c_int32), ("gt_need_oof", C.c_int32),
    ]


class PlanHost(C.Structure):
    """struct ta_plan_host."""
    _fields_ = (
        [(n, C.c_int64) for n in ("n_groups", "n_dt", "n_gt", "n_dt_boxes", "n_gt_boxes",
                                  "n_big")]
        + [(n, C.c_int32) for n in ("n_cat", "n_cfg", "n_thr", "n_rec", "n_slots_max", "g_max",
                                    "iou_mode", "flags")]
        + [(n, C.c_void_p) for n in (
            "grp_dt_off", "grp_gt_off", "iou_off", "cat_dt_off", "grp_cat", "acc_perm", "big_list",
            "dt_box", "gt_box", "dt_trk_off", "gt_trk_off", "dt_slot", "gt_slot",
            "dt_attr_a", "dt_attr_b", "gt_attr_a", "gt_attr_b", "dt_flag", "gt_flag",
            "gt_hp", "iou_thrs", "rec_thrs", "cfgs", "dt_box_idx")]
        + [("dt_box_pool", C.c_int32), ("reserved_", C.c_int32)]
    )


class PeerCopy(C.Structure):
    """struct ta_peer_copy."""
    _fields_ = [("peer", C.c_int32), ("reserved_", C.c_int32), ("src_off", C.c_int64),
                ("bytes", C.c_int64), ("dst", C.c_void_p)]


class PeerSum(C.Structure):
    """struct ta_peer_sum."""
    _fields_ = [("off", C.c_int64), ("count", C.c_int64), ("dst", C.c_void_p)]


class HostOut(C.Structure):
    """struct ta_host_out."""
    _fields_ = [(n, C.c_void_p) for n in ("precision", "recall", "tp_cnt", "fp_cnt", "num_gt")]


class TaEvalError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("ta_eval error %d: %s" % (code, msg))
        self.code = code


_lib: Optional[C.CDLL] = None


def lib_available() -> bool:
    return os.path.exists(LIB_PATH)


def load() -> C.CDLL:
    """Load libta_eval.so; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "CUDA library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the evaluation kernels)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    P, I32, I64 = C.c_void_p, C.c_int32, C.c_int64
    lib.ta_abi_version.restype = C.c_int
    lib.ta_last_error.restype = C.c_char_p
    lib.ta_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.ta_ctx_destroy.argtypes = [P]
    lib.ta_ctx_sm_count.argtypes = [P]
    lib.ta_ctx_launch_count.argtypes = [P]
    lib.ta_ctx_timing.argtypes = [P, C.c_int]
    lib.ta_ctx_timing_read.argtypes = [P, C.c_char_p, C.c_int, P, P, C.c_int]
    lib.ta_ctx_launch_count.restype = I64
    lib.ta_track_iou.argtypes = [P, P, C.c_int, I64, P, P, P, P, P, P, P, P, I32, P, P]
    lib.ta_box_iou.argtypes = [P, P, I64, P, I64, P, P, P, P, P, P]
    lib.ta_match_greedy.argtypes = ([P, P, I64, P, I64, P, P, P, P, P, I32, P, I32, P,
                                     I64, P, P, P, I64, P, P, P, P, I32,
                                     P, P, P, P])
    lib.ta_frame_eval.argtypes = ([P, P, I64, P, P, P, P, P, I32, P, I32, P, I64, P,
                                   I64, P, P, I64, P, I32, P, P, I32, P, P, P, P, P, P])
    lib.ta_frame_sched_bytes.argtypes = [I64, I64, I64, I32, I32]
    lib.ta_frame_sched_bytes.restype = I64
    lib.ta_frame_sched_build.argtypes = [P, P, I64, P, P, P, I64, P, I64, P, P, I32, I32, P, P]
    lib.ta_exchange_unique_id.argtypes = [P, I32]
    lib.ta_exchange_create.argtypes = [P, I32, I32, P, C.POINTER(C.c_void_p)]
    lib.ta_exchange_destroy.argtypes = [P]
    lib.ta_exchange_rank.argtypes = [P]
    lib.ta_exchange_world.argtypes = [P]
    lib.ta_exchange_gather.argtypes = [P, P, I64, I32, P, P, P]
    lib.ta_exchange_scatter.argtypes = [P, P, I64, I32, P, P, P]
    lib.ta_exchange_alltoallv.argtypes = [P, P, P, P, P, P]
    lib.ta_exchange_allreduce_sum.argtypes = [P, P, P, I64, I32]
    lib.ta_widen_boxes.argtypes = [P, P, I64, P, P]
    lib.ta_ctx_take_assert_count.argtypes = [P, P, C.POINTER(C.c_int32)]
    lib.ta_ctx_debug_list_count.argtypes = [P, P, C.POINTER(C.c_int32)]
    lib.ta_zero.argtypes = [P, P, P, I64]
    lib.ta_exchange_group_begin.argtypes = [P]
    lib.ta_exchange_group_end.argtypes = [P]
    lib.ta_rle_iou.argtypes = [P, P, I64, P, I64, P, P, P, P, P, P, P, P, P, P, P, P]
    lib.ta_pr_accumulate.argtypes = [P, P, I32, P, P, I64, P, P, P, I32, I32, I32, P, P, P, P, P]
    lib.ta_eval_plan_host.argtypes = [P, C.POINTER(PlanHost), P, P, P, P, P,
                                      C.POINTER(I64), C.POINTER(I64)]
    lib.ta_eval_plans_host.argtypes = [I32, P, P, P, P, P]
    lib.ta_widen_u16.argtypes = [P, P, I64, P, P]
    lib.ta_offsets_from_counts.argtypes = [P, P, I64, P, P]
    lib.ta_gather_boxes.argtypes = [P, P, I64, P, P, P]
    lib.ta_peer_window_create.argtypes = [P, I64, P]
    lib.ta_peer_window_destroy.argtypes = [P]
    lib.ta_peer_window_ptr.argtypes = [P]
    lib.ta_peer_window_ptr.restype = C.c_void_p
    lib.ta_peer_window_set_plan.argtypes = [P, I32, P, I32, P]
    lib.ta_peer_window_acquire.argtypes = [P, P]
    lib.ta_peer_window_put.argtypes = [P, P, I64, P, I64]
    lib.ta_peer_window_exchange.argtypes = [P, P]
    lib.ta_peer_window_check.argtypes = [P, P, C.POINTER(I32)]
    lib.ta_host_alloc.argtypes = [C.c_size_t]
    lib.ta_host_alloc.restype = C.c_void_p
    lib.ta_host_free.argtypes = [P]
    lib.ta_host_free.restype = None
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("ta_last_error", "ta_ctx_launch_count", "ta_frame_sched_bytes",
                        "ta_host_alloc", "ta_host_free", "ta_peer_window_ptr"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != TA_OK:
        msg = load().ta_last_error().decode("utf-8", "replace")
        if rc == TA_ERR_ASSERT:
            raise AssertionError(msg)          # the reference asserts i <= u (eval.py:95)
        raise TaEvalError(rc, msg)
