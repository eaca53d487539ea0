"""ctypes binding of libreo_cuda.so (include/reo.h).  There is NO CPU fallback: if the CUDA library is
missing or cannot be loaded this module raises, loudly."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("REO_CUDA_LIB") or os.path.join(HERE, "libreo_cuda.so")  # override: tuning variants

REO_OK = 0
REO_ERR_DIM = -1
REO_ERR_ARG = -2
REO_ERR_CUDA = -3
REO_ERR_COMM = -4
REO_ERR_OOM = -5
REO_ERR_BOUNDS = -6
REO_ERR_UNSUPPORTED = -7
REO_ERR_STATE = -8

REO_I64, REO_F64, REO_I32, REO_F32 = 0, 1, 2, 3
REO_DATA_ON_DEVICE = 1
REO_OUT_PINNED = 2
REO_MAX_ITER_LOG = 256

# every symbol include/reo.h declares
SYMBOLS = [
    "reo_version", "reo_create", "reo_destroy", "reo_last_error", "reo_set_collective", "reo_comm_unique_id",
    "reo_comm_init_rank", "reo_host_alloc", "reo_host_free", "reo_threshold",
    "reo_identify_degs", "reo_iter_log", "reo_debug_pair_plan", "reo_stage", "reo_stage_info", "reo_pair_counts", "reo_tables", "reo_tables_delta",
    "reo_mccullagh", "reo_empirical_null", "reo_bh", "reo_sort_f64", "reo_pseudobulk", "reo_detect_counts", "reo_subset",
]


class ReoStats(C.Structure):
    _fields_ = [
        ("iters_done", C.c_int32),
        ("converged", C.c_int32),
        ("n_deg", C.c_int32 * REO_MAX_ITER_LOG),
        ("n_ref", C.c_int32 * REO_MAX_ITER_LOG),
        ("rank_bits", C.c_int32),
        ("sample_words", C.c_int32),
        ("compares", C.c_int64),
        ("ms_stage", C.c_double),
        ("ms_pairs", C.c_double),
        ("ms_stats", C.c_double),
        ("ms_total", C.c_double),
        ("ms_wall", C.c_double),
        ("pair_launches", C.c_int32),
        ("kernel_launches", C.c_int32),
        ("ordered_triples", C.c_int64),
        ("planes_per_word", C.c_double),
    ]


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64)

_lib = None


def load():
    """Load libreo_cuda.so; raise if it was not built (run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            f"{SO_PATH} is missing: build it with `make -C {os.path.join(HERE, 'csrc')}` "
            "(there is no CPU fallback for the REO path)")
    L = C.CDLL(SO_PATH)
    vp, i32, i64, u64, u32, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_uint32, C.c_double
    L.reo_version.restype = C.c_int
    L.reo_version.argtypes = []
    L.reo_create.restype = C.c_int
    L.reo_create.argtypes = [C.POINTER(vp), C.c_int, vp, u64, u32]
    L.reo_destroy.restype = C.c_int
    L.reo_destroy.argtypes = [vp]
    L.reo_last_error.restype = C.c_char_p
    L.reo_last_error.argtypes = [vp]
    L.reo_set_collective.restype = C.c_int
    L.reo_set_collective.argtypes = [vp, C.c_int, C.c_int, ALLGATHER_FN, vp]
    L.reo_comm_unique_id.restype = C.c_int
    L.reo_comm_unique_id.argtypes = [vp]
    L.reo_comm_init_rank.restype = C.c_int
    L.reo_comm_init_rank.argtypes = [vp, C.c_int, C.c_int, vp]
    L.reo_host_alloc.restype = vp
    L.reo_host_alloc.argtypes = [C.c_size_t]
    L.reo_host_free.restype = None
    L.reo_host_free.argtypes = [vp]
    L.reo_threshold.restype = C.c_int
    L.reo_threshold.argtypes = [C.c_int, dbl]
    L.reo_identify_degs.restype = C.c_int
    L.reo_identify_degs.argtypes = [vp, vp, C.c_int, i64, i64, i64, vp, i32, vp, dbl, dbl, dbl, vp, i32, i32, u32,
                                    vp, vp, vp, vp, C.POINTER(ReoStats)]
    L.reo_iter_log.restype = C.c_int
    L.reo_iter_log.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32), vp, vp, i32]
    L.reo_debug_pair_plan.restype = C.c_longlong
    L.reo_debug_pair_plan.argtypes = [i64, i64, i32, i32, i32, i32, i32, vp, i64, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.reo_stage.restype = C.c_int
    L.reo_stage.argtypes = [vp, vp, C.c_int, i64, i64, i64, vp, i32, u32]
    L.reo_stage_info.restype = C.c_int
    L.reo_stage_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.reo_pair_counts.restype = C.c_int
    L.reo_pair_counts.argtypes = [vp, i32, vp, i32, vp, i32, vp, vp]
    L.reo_tables.restype = C.c_int
    L.reo_tables.argtypes = [vp, i32, vp, dbl, vp, vp]
    L.reo_tables_delta.restype = C.c_int
    L.reo_tables_delta.argtypes = [vp, i32, vp, dbl, vp, vp, vp]
    L.reo_mccullagh.restype = C.c_int
    L.reo_mccullagh.argtypes = [vp, vp, i64, i32, vp]
    L.reo_empirical_null.restype = C.c_int
    L.reo_empirical_null.argtypes = [vp, vp, i64, vp, vp]
    L.reo_bh.restype = C.c_int
    L.reo_bh.argtypes = [vp, vp, i64, vp]
    L.reo_sort_f64.restype = C.c_int
    L.reo_sort_f64.argtypes = [vp, vp, i64, vp, vp]
    L.reo_pseudobulk.restype = C.c_int
    L.reo_pseudobulk.argtypes = [vp, vp, C.c_int, i64, i64, i64, vp, vp, i32, u32, vp, C.POINTER(vp)]
    L.reo_detect_counts.restype = C.c_int
    L.reo_detect_counts.argtypes = [vp, vp, C.c_int, i64, i64, i64, u32, vp, vp]
    L.reo_subset.restype = C.c_int
    L.reo_subset.argtypes = [vp, vp, C.c_int, i64, i64, i64, vp, i64, vp, i64, u32, vp, C.POINTER(vp)]
    _lib = L
    return L
