"""
Host-side mirror of the reference's interface for the REO path, on top of the C ABI (include/reo.h).

Reference (pathint/RankCompV3.jl, src/RankCompV3.jl):
    get_major_reo_lower_count   81-92
    McCullagh_test              225-259   (exported, src:23)
    identify_degs               339-438
    reoa                        536-685   (exported, src:21; only 652-662 is re-pointed at the library)
Names, argument meaning and error behaviour follow the reference; Julia exceptions map to Python
ones (DimensionMismatch -> ValueError, ArgumentError -> ValueError, BoundsError -> IndexError).
Everything numeric is computed by libreo_cuda.so on the GPU; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass, field

import numpy as np

from . import _lib as L

HEADER = ["pval", "padj", "n11", "n12", "n13", "n21", "n22", "n23", "n31", "n32", "n33",
          "Δ1", "Δ2", "se", "z1", "up_down"]  # src:665


class _PinnedBlock:
    """A page-locked host block from reo_host_alloc; numpy arrays built on it keep it alive through .base, and
    it goes back to the pool (not to the driver: cudaFreeHost costs more than the copy it saves) when dropped."""

    __slots__ = ("ptr", "nbytes", "__array_interface__")

    def __init__(self, ptr, nbytes):
        self.ptr, self.nbytes = ptr, nbytes
        self.__array_interface__ = {"data": (ptr, False), "shape": (nbytes,), "typestr": "|u1", "version": 3}

    def __del__(self):
        global _pinned_pooled
        try:
            if _pinned_pooled + self.nbytes <= _PINNED_POOL_CAP:
                _pinned_pool.setdefault(self.nbytes, []).append(self.ptr)
                _pinned_pooled += self.nbytes
            else:
                L.load().reo_host_free(self.ptr)
        except Exception:  # interpreter shutdown
            pass


_pinned_pool: dict = {}
_pinned_pooled = 0                 # bytes parked in the pool
_PINNED_POOL_CAP = 1 << 28         # beyond 256 MiB idle blocks go back to the driver
_PINNED_ROUND = 1 << 16


def _pinned_bytes(n):
    """n bytes (uint8 array) of page-locked memory; sizes are rounded up to 64 KiB so blocks recycle across calls."""
    nbytes = max((n + _PINNED_ROUND - 1) // _PINNED_ROUND, 1) * _PINNED_ROUND
    global _pinned_pooled
    free = _pinned_pool.get(nbytes)
    if free:
        ptr = free.pop()
        _pinned_pooled -= nbytes
    else:
        ptr = L.load().reo_host_alloc(nbytes)
    if not ptr:
        raise MemoryError(f"reo_host_alloc({nbytes}) failed")
    return np.asarray(_PinnedBlock(ptr, nbytes))


def _pinned_empty(shape, dtype):
    """np.empty(shape, dtype) in page-locked memory."""
    dt = np.dtype(dtype)
    n = math.prod(shape) * dt.itemsize
    return _pinned_bytes(n)[:n].view(dt).reshape(shape)


class ReoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libreo_cuda error {code}: {msg}")
        self.code = code


def _raise(code: int, msg: str):
    if code == L.REO_ERR_DIM:
        raise ValueError("DimensionMismatch: " + msg)
    if code == L.REO_ERR_ARG:
        raise ValueError("ArgumentError: " + msg)
    if code == L.REO_ERR_BOUNDS:
        raise IndexError(msg)
    if code == L.REO_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if code == L.REO_ERR_OOM:
        raise MemoryError(msg)
    raise ReoError(code, msg)


def _ptr(a):
    return None if a is None else a.ctypes.data   # plain address: c_void_p argtypes take ints


_DT = {np.dtype(np.int64): L.REO_I64, np.dtype(np.float64): L.REO_F64,
       np.dtype(np.int32): L.REO_I32, np.dtype(np.float32): L.REO_F32}


def _as_colmajor(data):
    """Julia layout (column-major r x c) of a numpy matrix, in a dtype the ABI takes."""
    a = np.asarray(data)
    if a.ndim != 2:
        raise ValueError("DimensionMismatch: 'data' must be a matrix")
    if a.dtype not in _DT:
        a = a.astype(np.int64 if np.issubdtype(a.dtype, np.integer) or a.dtype == bool else np.float64)
    return np.asfortranarray(a)


def group_levels(group):
    """unique(group) in order of first appearance (src:353) -> (levels, 0-based level id per sample)."""
    levels, ids, seen = [], np.empty(len(group), dtype=np.int32), {}
    for s, g in enumerate(group):
        if g not in seen:
            seen[g] = len(levels)
            levels.append(g)
        ids[s] = seen[g]
    return levels, ids


@dataclass
class DeviceMatrix:
    """A column-major r x c expression matrix already resident in HBM (REO_DATA_ON_DEVICE)."""
    ptr: int
    dtype: int
    r: int
    c: int
    ld: int
    keepalive: object = None


@dataclass
class DegResult:
    result: np.ndarray        # [K, r, 15]
    updown: np.ndarray        # [K, r] int8
    final_ref: np.ndarray     # [K, r] uint8
    iters: list
    stats: dict = field(default_factory=dict)


class Reo:
    """One libreo_cuda handle (one GPU)."""

    def __init__(self, device=0, seed: int = 0):
        """device: one CUDA device index, or a list of indices (single process driving several GPUs: gene-row
        tiles sharded over them, tables all-gathered with NCCL inside the library)."""
        self._lib = L.load()
        self._h = C.c_void_p()
        dl = [int(d) for d in device] if isinstance(device, (list, tuple)) else [int(device)]
        devs = (C.c_int * len(dl))(*dl)
        rc = self._lib.reo_create(C.byref(self._h), len(dl), devs, C.c_uint64(int(seed) & (2**64 - 1)), 0)
        if rc != 0:
            _raise(rc, (self._lib.reo_last_error(None) or b"").decode())
        self._cb = None
        self.device = dl[0]
        self.devices = dl

    # -- plumbing ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.reo_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            _raise(rc, (self._lib.reo_last_error(self._h) or b"").decode())

    def set_collective(self, rank: int, world: int, allgather=None):
        """allgather(dev_ptr:int, bytes_per_rank:int) -> None, all-gathers in place (see reo.h)."""
        if allgather is None:
            cb = C.cast(None, L.ALLGATHER_FN)
        else:
            def _tramp(_ctx, dev_ptr, nbytes):
                try:
                    allgather(int(dev_ptr), int(nbytes))
                    return 0
                except Exception:  # never let an exception cross the C boundary
                    import traceback
                    traceback.print_exc()
                    return 1
            cb = L.ALLGATHER_FN(_tramp)
        self._cb = cb
        self._check(self._lib.reo_set_collective(self._h, int(rank), int(world), cb, None))

    def comm_init(self, rank: int, world: int, unique_id: bytes):
        """One process per GPU: join the NCCL communicator identified by `unique_id` (128 bytes from
        `nccl_unique_id()` on rank 0, shared over any host channel)."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self._lib.reo_comm_init_rank(self._h, int(rank), int(world), buf))

    @staticmethod
    def _matrix_args(data):
        if isinstance(data, DeviceMatrix):
            return C.c_void_p(data.ptr), data.dtype, data.r, data.c, data.ld, L.REO_DATA_ON_DEVICE, data
        a = _as_colmajor(data)
        r, c = a.shape
        return _ptr(a), _DT[a.dtype], r, c, r, 0, a

    # -- the whole path -------------------------------------------------------------------------
    def identify_degs(self, data, group_id, gnum, ref_mask, pval_reo=0.01, pval_deg=1.0, padj_deg=0.05,
                      n_iter=128, n_conv=5, thresholds=None) -> DegResult:
        p, dt, r, c, ld, flags, keep = self._matrix_args(data)
        gid = np.ascontiguousarray(group_id, dtype=np.int32)
        if len(gid) != c:
            raise ValueError("DimensionMismatch: 'data' and 'group' do not have compatiable sizes")  # src:355
        ref = np.asarray(ref_mask)
        if ref.dtype == np.bool_ and ref.flags.c_contiguous:
            ref = ref.view(np.uint8)             # a BitVector-like mask: its bytes are already 0/1
        else:
            ref = np.ascontiguousarray(ref != 0, dtype=np.uint8)
        if len(ref) != r:
            raise ValueError("DimensionMismatch: 'ref_gene' and 'data' do not have compatiable sizes")
        K = 1 if gnum == 2 else max(int(gnum), 1)
        thr = None if thresholds is None else np.asfortranarray(np.asarray(thresholds, dtype=np.int32))
        # outputs live in ONE recycled page-locked block: the library copies device->host straight into it
        nres = K * 15 * r * 8
        blk = _pinned_bytes(nres + 2 * K * r)
        result = blk[:nres].view(np.float64).reshape(K, 15, r)   # column-major r x 15 per k, filled by the library
        updown = blk[nres:nres + K * r].view(np.int8).reshape(K, r)
        final_ref = blk[nres + K * r:nres + 2 * K * r].reshape(K, r)
        flags |= L.REO_OUT_PINNED
        iters = np.zeros(K, dtype=np.int32)
        st = L.ReoStats()
        rc = self._lib.reo_identify_degs(self._h, p, dt, r, c, ld, _ptr(gid), int(gnum), _ptr(thr), float(pval_reo),
                                         float(pval_deg), float(padj_deg), _ptr(ref), int(n_iter), int(n_conv), flags,
                                         _ptr(result), _ptr(updown), _ptr(final_ref), _ptr(iters), C.byref(st))
        del keep
        self._check(rc)
        ne = min(int(st.iters_done), L.REO_MAX_ITER_LOG)
        stats = dict(iters_done=int(st.iters_done), converged=int(st.converged), n_deg=list(st.n_deg[:ne]),
                     n_ref=list(st.n_ref[:ne]), rank_bits=int(st.rank_bits), sample_words=int(st.sample_words),
                     compares=int(st.compares), ms_stage=st.ms_stage, ms_pairs=st.ms_pairs, ms_stats=st.ms_stats,
                     ms_total=st.ms_total, ms_wall=st.ms_wall, pair_launches=int(st.pair_launches),
                     kernel_launches=int(st.kernel_launches), ordered_triples=int(st.ordered_triples),
                     planes_per_word=float(st.planes_per_word))
        # [K, r, 15] view of the library's column-major output (no copy)
        return DegResult(result.transpose(0, 2, 1), updown, final_ref, [int(v) for v in iters], stats)

    def iter_log(self, k: int):
        """Iteration log of level k of the last identify_degs: dict(iters_done, converged, n_deg[], n_ref[]) -- what the
        reference prints per level (src:418-420, 432-435)."""
        it, cv = C.c_int32(), C.c_int32()
        nd = np.zeros(L.REO_MAX_ITER_LOG, dtype=np.int32)
        nr = np.zeros(L.REO_MAX_ITER_LOG, dtype=np.int32)
        self._check(self._lib.reo_iter_log(self._h, int(k), C.byref(it), C.byref(cv), _ptr(nd), _ptr(nr), len(nd)))
        n = min(it.value, len(nd))
        return dict(iters_done=it.value, converged=cv.value, n_deg=nd[:n].tolist(), n_ref=nr[:n].tolist())

    # -- stage-level entry points -----------------------------------------------------------------
    def stage(self, data, group_id, gnum):
        p, dt, r, c, ld, flags, keep = self._matrix_args(data)
        gid = np.ascontiguousarray(group_id, dtype=np.int32)
        if len(gid) != c:
            raise ValueError("DimensionMismatch: 'data' and 'group' do not have compatiable sizes")
        self._check(self._lib.reo_stage(self._h, p, dt, r, c, ld, _ptr(gid), int(gnum), flags))
        self._r = r
        del keep
        b, w, t = C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self._lib.reo_stage_info(self._h, C.byref(b), C.byref(w), C.byref(t)))
        return dict(rank_bits=b.value, sample_words=w.value, gene_tiles=t.value)

    def pair_counts(self, k, rows, cols):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        nre = np.zeros((len(rows), len(cols)), dtype=np.int32)
        rest = np.zeros_like(nre)
        self._check(self._lib.reo_pair_counts(self._h, int(k), _ptr(rows), len(rows), _ptr(cols), len(cols),
                                              _ptr(nre), _ptr(rest)))
        return nre, rest

    def tables(self, k, mask, thresholds=None, pval_reo=0.01, mask_to=None):
        mask = np.ascontiguousarray(np.asarray(mask) != 0, dtype=np.uint8)
        thr = None if thresholds is None else np.asfortranarray(np.asarray(thresholds, dtype=np.int32))
        table = np.zeros((len(mask), 9), dtype=np.int32)
        if mask_to is None:
            rc = self._lib.reo_tables(self._h, int(k), _ptr(thr), float(pval_reo), _ptr(mask), _ptr(table))
        else:
            m2 = np.ascontiguousarray(np.asarray(mask_to) != 0, dtype=np.uint8)
            rc = self._lib.reo_tables_delta(self._h, int(k), _ptr(thr), float(pval_reo), _ptr(mask), _ptr(m2),
                                            _ptr(table))
        self._check(rc)
        return table

    def mccullagh(self, tables):
        t = np.ascontiguousarray(tables, dtype=np.int64)
        if t.ndim == 2:
            t = t[None]
        n, k, k2 = t.shape
        if k != k2:
            raise ValueError("DimensionMismatch: input matrix 'mat' should be a square matrix.")  # src:227
        out = np.zeros((n, 5))
        self._check(self._lib.reo_mccullagh(self._h, _ptr(t), n, k, _ptr(out)))
        return out

    def empirical_null(self, delta1):
        d = np.ascontiguousarray(delta1, dtype=np.float64)
        p = np.zeros_like(d)
        se = np.zeros(1)
        self._check(self._lib.reo_empirical_null(self._h, _ptr(d), len(d), _ptr(p), _ptr(se)))
        return float(se[0]), p

    def bh(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        q = np.zeros_like(p)
        self._check(self._lib.reo_bh(self._h, _ptr(p), len(p), _ptr(q)))
        return q

    # -- the steps right before the path ------------------------------------------
    def pseudobulk(self, data, profiles, to_host=True):
        """profiles: list of cell-index arrays (one per pseudo-bulk profile, summed in the given order), src:56-67.
        Returns (host matrix r x P or None, DeviceMatrix of the result kept in HBM)."""
        p, dt, r, c, ld, flags, keep = self._matrix_args(data)
        ptr = np.zeros(len(profiles) + 1, dtype=np.int32)
        ptr[1:] = np.cumsum([len(x) for x in profiles])
        cells = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.int32) for x in profiles])
                                     if len(profiles) else np.zeros(0, np.int32), dtype=np.int32)
        is_int = dt in (L.REO_I64, L.REO_I32)
        out = np.empty((len(profiles), r), dtype=np.int64 if is_int else np.float64) if to_host else None
        dev = C.c_void_p()
        self._check(self._lib.reo_pseudobulk(self._h, p, dt, r, c, ld, _ptr(ptr), _ptr(cells), len(profiles), flags,
                                             _ptr(out), C.byref(dev)))
        del keep
        dm = DeviceMatrix(dev.value, L.REO_I64 if is_int else L.REO_F64, r, len(profiles), r, keepalive=self)
        return (out.T if out is not None else None), dm

    def detect_counts(self, data):
        """src:618, 626 -> (detected genes per cell [c], detecting cells per gene [r])."""
        p, dt, r, c, ld, flags, keep = self._matrix_args(data)
        per_cell = np.zeros(c, dtype=np.int32)
        per_gene = np.zeros(r, dtype=np.int32)
        self._check(self._lib.reo_detect_counts(self._h, p, dt, r, c, ld, flags, _ptr(per_cell), _ptr(per_gene)))
        del keep
        return per_cell, per_gene

    def subset(self, data, genes, cells, to_host=True):
        """data[genes, cells] (src:624-628) -> (host matrix or None, DeviceMatrix)."""
        p, dt, r, c, ld, flags, keep = self._matrix_args(data)
        genes = np.ascontiguousarray(genes, dtype=np.int32)
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        npdt = {v: k for k, v in _DT.items()}[dt]
        out = np.empty((len(cells), len(genes)), dtype=npdt) if to_host else None
        dev = C.c_void_p()
        self._check(self._lib.reo_subset(self._h, p, dt, r, c, ld, _ptr(genes), len(genes), _ptr(cells), len(cells), flags,
                                         _ptr(out), C.byref(dev)))
        del keep
        dm = DeviceMatrix(dev.value, dt, len(genes), len(cells), len(genes), keepalive=self)
        return (out.T if out is not None else None), dm

    def sort(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        s = np.zeros_like(x)
        perm = np.zeros(len(x), dtype=np.int32)
        self._check(self._lib.reo_sort_f64(self._h, _ptr(x), len(x), _ptr(s), _ptr(perm)))
        return s, perm


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = L.load().reo_comm_unique_id(buf)
    if rc != 0:
        _raise(rc, (L.load().reo_last_error(None) or b"").decode())
    return buf.raw


# ---- module-level default handle ------------------------------------------------------------------
_default: dict = {}


def default_handle(seed: int = 0) -> Reo:
    """One cached handle per (device, tie seed): a second call with another seed gets its own coins."""
    key = (int(os.environ.get("LOCAL_RANK", "0")), int(seed))
    if key not in _default:
        _default[key] = Reo(key[0], key[1])
    return _default[key]


# ---- reference-named functions ----------------------------------------------------------------------
def get_major_reo_lower_count(sample_size: int, pval_threshold: float = 0.01) -> int:
    """src:81-92."""
    return int(L.load().reo_threshold(int(sample_size), float(pval_threshold)))


def McCullagh_test(mat, handle: Reo | None = None):
    """src:225-259 -> (pval, Δ1, Δ2, se, z1)."""
    m = np.asarray(mat)
    if m.ndim != 2 or m.shape[0] != m.shape[1]:
        raise ValueError("DimensionMismatch: input matrix 'mat' should be a square matrix.")
    out = (handle or default_handle()).mccullagh(m)[0]
    return tuple(float(v) for v in out)


def identify_degs(data, group, gene_names, pval_reo, pval_deg, padj_deg, ref_gene, n_iter, n_conv,
                  handle: Reo | None = None, return_raw: bool = False):
    """
    src:339-438, same positional arguments.  Returns the reference's `res`: an object matrix
    r x (1 + 16K): gene names, then per k the 15 result columns and the "up"/"down"/"no change" column.
    """
    a = data if isinstance(data, DeviceMatrix) else np.asarray(data)
    c = a.c if isinstance(a, DeviceMatrix) else a.shape[1]
    if c != len(group):
        raise ValueError("DimensionMismatch: 'data' and 'group' do not have compatiable sizes")
    levels, gid = group_levels(list(group))
    if len(levels) < 2:
        raise ValueError("DimensionMismatch: Only 1 level in 'group1, at least 2 levels!")
    h = handle or default_handle()
    out = h.identify_degs(a, gid, len(levels), ref_gene, pval_reo, pval_deg, padj_deg, n_iter, n_conv)
    if return_raw:
        return out
    r = out.result.shape[1]
    K = out.result.shape[0]
    res = np.empty((r, 1 + 16 * K), dtype=object)
    res[:, 0] = list(gene_names)
    names = {1: "up", -1: "down", 0: "no change"}
    for k in range(K):
        res[:, 1 + 16 * k:16 + 16 * k] = out.result[k]
        res[:, 16 + 16 * k] = [names[int(v)] for v in out.updown[k]]
    return res
