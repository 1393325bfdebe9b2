"""ctypes loader for libmvtm.so, the C ABI declared in include/mvtm.h.

There is no fallback of any kind: if the shared library is missing or does not export every symbol of the
header, importing this module's `lib()` raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libmvtm.so")
MAX_VIEWS = 8


class MvtmConfig(C.Structure):
    _fields_ = [("num_topics", C.c_int32), ("num_views", C.c_int32), ("num_docs", C.c_int64),
                ("vocab_sizes", C.POINTER(C.c_int32)), ("seed", C.c_uint64), ("device", C.c_int32),
                ("flags", C.c_uint32), ("doc_id_base", C.c_int64), ("doc_id_stride", C.c_int64),
                ("warps_per_cta", C.c_int32), ("ring_depth", C.c_int32), ("max_ctas", C.c_int32)]


class MvtmSweepStats(C.Structure):
    _fields_ = [("tokens", C.c_int64), ("changed", C.c_int64), ("new_topic", C.c_int64), ("ms_total", C.c_double),
                ("ms_view", C.c_double * MAX_VIEWS), ("kernel_launches", C.c_int32),
                ("ring_depth", C.c_int32 * MAX_VIEWS), ("ring_locked", C.c_int32 * MAX_VIEWS)]


FLAG_DOC_ORDER = 1
FLAG_SINGLE_WARP = 2
FLAG_Q1_COMPAT = 4
FLAG_BETA_MALLET = 8
FLAG_REFERENCE_COMPAT = 12
OPT_P, OPT_DP, OPT_GAMMA, OPT_BETA, OPT_ALL = 1, 2, 4, 8, 15

_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64
STAT_REDUCER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_double), C.c_int64)
# name -> (restype, argtypes); must list every function include/mvtm.h declares (tests check this)
SIGNATURES = {
    "mvtm_create": (_i32, [C.POINTER(MvtmConfig), C.POINTER(_vp)]),
    "mvtm_destroy": (_i32, [_vp]),
    "mvtm_last_error": (C.c_char_p, [_vp]),
    "mvtm_add_view": (_i32, [_vp, _i32, _vp, _vp, _vp]),
    "mvtm_init_assignments": (_i32, [_vp]),
    "mvtm_set_assignments": (_i32, [_vp, _i32, _vp]),
    "mvtm_set_counts": (_i32, [_vp, _i32, _vp, _vp]),
    "mvtm_init_assignments_from_counts": (_i32, [_vp]),
    "mvtm_set_hyper": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32]),
    "mvtm_get_hyper": (_i32, [_vp, _vp, _vp, _vp, C.POINTER(_i32)]),
    "mvtm_sweep": (_i32, [_vp, _i32, _i32]),
    "mvtm_sweep_host": (_i32, [_vp, _i32, C.POINTER(_vp)]),
    "mvtm_set_host_mirror": (_i32, [_vp, _i32, _vp]),
    "mvtm_get_assignments": (_i32, [_vp, _i32, _vp]),
    "mvtm_get_counts": (_i32, [_vp, _i32, _vp, _vp]),
    "mvtm_doc_topic_hist": (_i32, [_vp, _i32, _vp, C.POINTER(_i32)]),
    "mvtm_loglik": (_i32, [_vp, _vp, _i32]),
    "mvtm_loglik_parts": (_i32, [_vp, _vp, _vp, _i32]),
    "mvtm_heldout_loglik": (_i32, [_vp, _i32, _vp, _vp, C.POINTER(C.c_double), C.POINTER(_i64)]),
    "mvtm_cond_probs": (_i32, [_vp, _i32, _i64, _i32, _vp, _vp]),
    "mvtm_cond_probs_ex": (_i32, [_vp, _i32, _i64, _i32, _vp, _i32, _vp, _i32, _vp]),
    "mvtm_check_invariants": (_i32, [_vp, C.POINTER(_i64)]),
    "mvtm_stats": (_i32, [_vp, C.POINTER(MvtmSweepStats)]),
    "mvtm_delta_begin": (_i32, [_vp]),
    "mvtm_delta_reset": (_i32, [_vp]),
    "mvtm_delta_export": (_i32, [_vp, _i32, C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_vp), C.POINTER(_i64)]),
    "mvtm_delta_import": (_i32, [_vp, _i32]),
    "mvtm_sum_exchange_buffers": (_i32, [_vp, _i32, C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_vp), C.POINTER(_i64)]),
    "mvtm_sum_exchange_finish": (_i32, [_vp, _i32, _i32]),
    "mvtm_row_stride": (_i32, [_vp, C.POINTER(_i32)]),
    "mvtm_sweep_view_async": (_i32, [_vp, _i32, _i32, _i32]),
    "mvtm_sweep_finish": (_i32, [_vp]),
    "mvtm_stream_wait_view": (_i32, [_vp, _i32, _vp]),
    "mvtm_view_wait_stream": (_i32, [_vp, _i32, _vp]),
    "mvtm_sum_exchange_finish_async": (_i32, [_vp, _i32, _i32, _vp, _i32]),
    "mvtm_comm_unique_id": (_i32, [_vp]),
    "mvtm_comm_init": (_i32, [_vp, _vp, _i32, _i32, _i32]),
    "mvtm_comm_destroy": (_i32, [_vp]),
    "mvtm_comm_info": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i64)]),
    "mvtm_sync_counts": (_i32, [_vp, _i32]),
    "mvtm_sweep_dist": (_i32, [_vp, _i32]),
    "mvtm_comm_drain": (_i32, [_vp]),
    "mvtm_sweep_host_dist": (_i32, [_vp, _i32, C.POINTER(_vp)]),
    "mvtm_comm_last_host_step": (_i32, [_vp, C.POINTER(_i32)]),
    "mvtm_loglik_dist": (_i32, [_vp, _vp, _i32]),
    "mvtm_scan_layout": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i32)]),
    "mvtm_optimize_hyper": (_i32, [_vp, _i32, C.c_uint32]),
    "mvtm_activate_topics": (_i32, [_vp]),
    "mvtm_set_stat_reducer": (_i32, [_vp, STAT_REDUCER, _vp]),
    "mvtm_p_statistics": (_i32, [_vp, _vp, _vp]),
    "mvtm_get_hyper_full": (_i32, [_vp] + [_vp] * 11),
    "mvtm_test_sampler": (_i32, [C.c_uint64, _i32, C.c_double, C.c_double, _i32, _vp]),
    "mvtm_test_learn_symmetric_concentration": (C.c_double, [_vp, _i32, _vp, _i32, _i32, C.c_double]),
    "mvtm_test_hyper_core": (_i32, [_i32, _i32, C.c_uint32] + [_vp] * 13 + [C.c_int64, _vp, _vp]),
    "mvtm_test_launch_shape": (_i32, [_i32, _i32, C.c_uint32, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i64), C.POINTER(_i32)]),
    "mvtm_build_info": (C.c_char_p, []),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(f"{SO_PATH} is missing: build it with `python -m mvtopicmodel_b200.build` "
                              "(nvcc, sm_100a). mvtopicmodel_b200 has no CPU or PyTorch fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)          # AttributeError if the symbol is not exported
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib
