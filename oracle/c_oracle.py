"""
oracle/c_oracle.py -- TEST INFRASTRUCTURE ONLY: ctypes access to oracle/_build/libreo_oracle.so
(the plain-C restatement in oracle/reo_oracle.c; reference lines cited there).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libreo_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "reo_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.reo_oracle_u.restype = C.c_uint32
        L.reo_oracle_u.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
        L.reo_oracle_num_threads.restype = C.c_int
        L.reo_oracle_set_threads.restype = None
        L.reo_oracle_set_threads.argtypes = [C.c_int]
        L.reo_oracle_set_input_f32.restype = None
        L.reo_oracle_set_input_f32.argtypes = [C.c_int]
        L.reo_oracle_set_coin_mode.restype = None
        L.reo_oracle_set_coin_mode.argtypes = [C.c_int]
        L.reo_oracle_threshold.restype = C.c_int
        L.reo_oracle_threshold.argtypes = [C.c_int, C.c_double]
        L.reo_oracle_mccullagh.restype = None
        L.reo_oracle_mccullagh.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.reo_oracle_categories.restype = None
        L.reo_oracle_categories.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                                            C.c_void_p, C.c_uint64, C.c_int, C.c_int64, C.c_int64, C.c_void_p]
        L.reo_oracle_block_tables.restype = C.c_int64
        L.reo_oracle_block_tables.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                                              C.c_void_p, C.c_uint64, C.c_int, C.c_int64, C.c_int64,
                                              C.c_void_p, C.c_int64, C.c_void_p]
        L.reo_oracle_block_tables_idx.restype = C.c_int64
        L.reo_oracle_block_tables_idx.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                                                  C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                                  C.c_void_p, C.c_int64, C.c_void_p]
        L.reo_oracle_empirical_null.restype = C.c_double
        L.reo_oracle_empirical_null.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.reo_oracle_bh.restype = None
        L.reo_oracle_bh.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.reo_oracle_identify_degs.restype = C.c_int
        L.reo_oracle_identify_degs.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                                               C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_int,
                                               C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _colmajor_f64(data):
    """Julia layout: column-major r x c Float64.  A float32 matrix switches the oracle to Float32 differences (the
    reference evaluates abs(x - y) in the matrix' own element type, src:72); any other dtype switches it back."""
    a = np.asarray(data)
    lib().reo_oracle_set_input_f32(1 if a.dtype == np.float32 else 0)
    return np.asfortranarray(a, dtype=np.float64)


def num_threads() -> int:
    return int(lib().reo_oracle_num_threads())


def use_all_cores() -> int:
    """OpenMP threads = the cores this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().reo_oracle_set_threads(int(n))
    return num_threads()


def set_coin_mode(mode: int) -> None:
    """0: the product's XOR coin rule; 1: independent per-pair coins (study of the tie rule only, scripts/coin_study.py)."""
    lib().reo_oracle_set_coin_mode(int(mode))


def threshold(n: int, pval: float) -> int:
    return int(lib().reo_oracle_threshold(int(n), float(pval)))


def thresholds_for(gid, gnum, pval_reo):
    gid = np.asarray(gid)
    c = len(gid)
    thr = np.zeros((2, gnum), dtype=np.int32)
    for k in range(gnum):
        n1 = int((gid == k).sum())
        thr[0, k] = threshold(n1, pval_reo)
        thr[1, k] = threshold(c - n1, pval_reo)
    return thr


def mccullagh(mat):
    m = np.ascontiguousarray(np.asarray(mat, dtype=np.int64))
    out = np.zeros(5)
    lib().reo_oracle_mccullagh(_p(m), int(m.shape[0]), _p(out))
    return tuple(out)


def categories(data, gid, gnum, thr, seed=0, k=0, i0=0, i1=None):
    d = _colmajor_f64(data)
    r, c = d.shape
    i1 = r if i1 is None else i1
    gid = np.ascontiguousarray(gid, dtype=np.int32)
    thr_cm = np.asfortranarray(np.asarray(thr, dtype=np.int32))
    cat = np.zeros((i1 - i0, r), dtype=np.uint8)
    lib().reo_oracle_categories(_p(d), r, c, r, _p(gid), int(gnum), _p(thr_cm), int(seed), int(k), i0, i1, _p(cat))
    return cat


def block_tables(data, gid, gnum, thr, cols, seed=0, k=0, i0=0, i1=None):
    d = _colmajor_f64(data)
    r, c = d.shape
    i1 = r if i1 is None else i1
    gid = np.ascontiguousarray(gid, dtype=np.int32)
    thr_cm = np.asfortranarray(np.asarray(thr, dtype=np.int32))
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    tab = np.zeros((i1 - i0, 9), dtype=np.int32)
    n = lib().reo_oracle_block_tables(_p(d), r, c, r, _p(gid), int(gnum), _p(thr_cm), int(seed), int(k), i0, i1,
                                      _p(cols), len(cols), _p(tab))
    return tab, int(n)


def block_tables_idx(sub, gidx, gid, gnum, thr, rows_local, cols_local, seed=0, k=0):
    """Tables of sub-matrix rows against sub-matrix columns; local row i is global gene gidx[i] (ascending)."""
    d = _colmajor_f64(sub)
    r, c = d.shape
    gid = np.ascontiguousarray(gid, dtype=np.int32)
    gidx = np.ascontiguousarray(gidx, dtype=np.int32)
    assert np.all(np.diff(gidx) > 0)
    thr_cm = np.asfortranarray(np.asarray(thr, dtype=np.int32))
    rows = np.ascontiguousarray(rows_local, dtype=np.int32)
    cols = np.ascontiguousarray(cols_local, dtype=np.int32)
    tab = np.zeros((len(rows), 9), dtype=np.int32)
    lib().reo_oracle_block_tables_idx(_p(d), r, c, r, _p(gid), int(gnum), _p(thr_cm), int(seed), int(k), _p(gidx),
                                      _p(rows), len(rows), _p(cols), len(cols), _p(tab))
    return tab


def empirical_null(d1):
    d1 = np.ascontiguousarray(d1, dtype=np.float64)
    p = np.zeros_like(d1)
    se = lib().reo_oracle_empirical_null(_p(d1), len(d1), _p(p))
    return float(se), p


def bh(p):
    p = np.ascontiguousarray(p, dtype=np.float64)
    q = np.zeros_like(p)
    lib().reo_oracle_bh(_p(p), len(p), _p(q))
    return q


def identify_degs(data, gid, gnum, thr, pval_deg, padj_deg, ref_mask, n_iter, n_conv, seed=0):
    d = _colmajor_f64(data)
    r, c = d.shape
    K = 1 if gnum == 2 else gnum
    gid = np.ascontiguousarray(gid, dtype=np.int32)
    thr_cm = np.asfortranarray(np.asarray(thr, dtype=np.int32))
    ref = np.ascontiguousarray(ref_mask, dtype=np.uint8)
    result = np.zeros((K, r, 15))
    updown = np.zeros((K, r), dtype=np.int8)
    final_ref = np.zeros((K, r), dtype=np.uint8)
    iters = np.zeros(K, dtype=np.int32)
    deg_log = np.zeros((K, max(n_iter, 1)), dtype=np.int32)
    rc = lib().reo_oracle_identify_degs(_p(d), r, c, r, _p(gid), int(gnum), _p(thr_cm), float(pval_deg),
                                        float(padj_deg), _p(ref), int(n_iter), int(n_conv), int(seed), _p(result),
                                        _p(updown), _p(final_ref), _p(iters), _p(deg_log))
    if rc != 0:
        raise ValueError(f"reo_oracle_identify_degs failed: {rc}")
    return dict(result=result, updown=updown, final_ref=final_ref, iters=[int(v) for v in iters], deg_log=deg_log)
