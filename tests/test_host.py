"""CPU suite: the C-ABI library loads and exports every symbol of include/reo.h; host-side logic."""
import ctypes
import os
import re

import numpy as np
import pytest
from conftest import ROOT


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg._lib.load()
    header = open(os.path.join(ROOT, "include", "reo.h")).read()
    declared = set(re.findall(r"\b(reo_[a-z0-9_]+)\s*\(", header)) - {"reo_allgather_fn"}
    assert declared == set(pkg._lib.SYMBOLS), declared ^ set(pkg._lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.reo_version() == 201


def test_no_gpu_fails_loudly(pkg):
    """Without a CUDA device the product path must raise, not fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(Exception) as e:
        pkg.Reo(0)
    assert "CUDA" in str(e.value) or "cuda" in str(e.value)


def test_product_does_not_import_oracle():
    """The product path (package + headers) never imports, links or executes anything under oracle/."""
    pdir = os.path.join(ROOT, "rankcompv3.jl_b200")
    for dp, _, files in os.walk(pdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                lines = open(os.path.join(dp, f)).read().splitlines()
                code = [l for l in lines if not l.lstrip().startswith(("//", "#", "*", "/*"))]
                for l in code:
                    assert "reo_oracle" not in l and "c_oracle" not in l and "load_oracle" not in l, (f, l)


def test_group_levels(pkg):
    lv, gid = pkg.api.group_levels(["b", "a", "b", "c", "a"])
    assert lv == ["b", "a", "c"] and gid.tolist() == [0, 1, 0, 2, 1]


def test_dimension_errors_before_gpu(pkg):
    data = np.zeros((20, 4), dtype=np.int64)
    with pytest.raises(ValueError, match="DimensionMismatch"):
        pkg.identify_degs(data, ["a", "b", "a"], list("x" * 20), 0.01, 1.0, 0.05, np.ones(20, bool), 4, 1)
    with pytest.raises(ValueError, match="DimensionMismatch"):
        pkg.identify_degs(data, ["a"] * 4, list("x" * 20), 0.01, 1.0, 0.05, np.ones(20, bool), 4, 1)


def test_pseudobulk_group(pkg):
    """src:56-67: chunks of ceil(c/n_pseudo) cells, row sums, every cell used exactly once."""
    rng = np.random.default_rng(0)
    x = rng.integers(0, 10, size=(30, 103))
    pb = pkg.pseudobulk_group(x, 10, np.random.default_rng(1))
    assert pb.shape == (30, 10)  # ceil(103/10) = 11 cells per profile -> 10 profiles
    assert np.array_equal(pb.sum(axis=1), x.sum(axis=1))
    pb2 = pkg.pseudobulk_group(x, 50, np.random.default_rng(1))
    assert pb2.shape[1] == 35  # ceil(103/50) = 3 -> 35 profiles: fewer than n_pseudo


def test_synth_shapes(pkg):
    d, g, de = pkg.synth.bulk(300, 6, 7, seed=1)
    assert d.shape == (300, 13) and d.dtype == np.int64 and len(g) == 13 and de.sum() == 30
    tf = pkg.synth.tie_free(d)
    assert all(len(set(tf[:, s])) == 300 for s in range(13))
    m = pkg.synth.reference_mask(de, 50)
    assert m.sum() == 50 and not (m & de).any()


def test_pair_plan_covers_every_tile_pair_once(pkg):
    """The library's partition (reo_debug_pair_plan = csrc/reo_pairs2.cu on the host): over all ranks every needed
    (row tile, column tile) is evaluated exactly once, in every shape, at any rank count."""
    from importlib import import_module
    dist = import_module(pkg.__name__ + ".dist")
    shapes = ((30000, 30000, 625, 8), (20000, 3000, 7, 13), (20000, 15000, 7, 13), (700, 700, 1, 10), (130, 70, 3, 5),
              # few reference columns against all genes (signed updates): supertiles narrower than they are tall
              (30000, 64, 625, 8), (30000, 128, 625, 8), (30000, 448, 625, 8), (20000, 500, 7, 13), (5000, 200, 40, 9))
    for r, ncols, W, NP in shapes:
        for world in (1, 3, 8):
            seen = {}
            sizes = []
            for rank in range(world):
                trip, nsym, ntr, ntc = dist.pair_plan(r, ncols, W, NP, rank, world)
                sizes.append(len(trip))
                for I, J, fl in trip:
                    assert (int(I), int(J)) not in seen
                    seen[(int(I), int(J))] = int(fl)
            nt = -(-r // 64)
            tc = -(-ncols // 64)
            if nsym == 0:
                want = {(i, j): 1 for i in range(nt) for j in range(tc)}
            else:
                want = {(i, j): (3 if j > i else 1) for i in range(nsym) for j in range(i, nsym)}
                n_rest = -(-(r - ncols) // 64)
                want.update({(i, j): 1 for i in range(ntr - n_rest, ntr) for j in range(ntc)})
            assert seen == want, (r, ncols, world)
            if sum(sizes) >= 4000:   # balance matters (and is only meaningful) on large tile spaces; edge half pairs count 1
                assert max(sizes) - min(sizes) <= 0.06 * sum(sizes) / world, (r, ncols, world, sizes)


def test_pinned_output_pool_recycles_and_caps(pkg, monkeypatch):
    """The Python wrapper's page-locked output blocks are recycled by size and handed back to the driver beyond the
    pool cap (allocator stubbed: no GPU needed)."""
    api = pkg.api

    class Stub:
        def __init__(self):
            self.allocs, self.freed, self.keep = 0, [], []

        def reo_host_alloc(self, nbytes):
            self.allocs += 1
            b = (ctypes.c_char * nbytes)()
            self.keep.append(b)
            return ctypes.addressof(b)

        def reo_host_free(self, p):
            self.freed.append(p)

    st = Stub()
    monkeypatch.setattr(api.L, "load", lambda: st)
    monkeypatch.setattr(api, "_pinned_pool", {})
    monkeypatch.setattr(api, "_pinned_pooled", 0)
    a = api._pinned_empty((3, 5), np.float64)
    a[:] = 1.5
    p0 = a.ctypes.data
    v = a.T                      # views keep the block alive
    del a
    assert v[0, 0] == 1.5 and st.allocs == 1 and not api._pinned_pool
    del v
    b = api._pinned_empty((15,), np.float64)
    assert b.ctypes.data == p0 and st.allocs == 1     # recycled, not re-allocated
    monkeypatch.setattr(api, "_PINNED_POOL_CAP", 0)
    del b
    assert st.freed == [p0] and api._pinned_pooled == 0




def test_host_narrowing_copies(pkg):
    """csrc/reo_host.cpp (host code, no GPU): the copy threads narrow a chunk of the caller's matrix to u16 only if every
    value is an integer in 0..65535; anything else must be reported so that the chunk travels raw."""
    lib = pkg._lib.load()
    rng = np.random.default_rng(5)
    n = 4099                                                    # not a multiple of any vector width
    good = rng.integers(0, 65536, n)
    good[:3] = (0, 65535, 1)
    cases = {
        "reo_host_narrow_i64": (np.int64, [65536, -1, 1 << 40, -(1 << 62)]),
        "reo_host_narrow_i32": (np.int32, [65536, -1, -(1 << 31)]),
        "reo_host_narrow_f64": (np.float64, [65536.0, -1.0, 0.5, 1e-300, 1e20, 2.0 ** 52, np.nan, np.inf, -np.inf, 3.0000000000000004]),
        "reo_host_narrow_f32": (np.float32, [65536.0, -1.0, 0.5, 1e-30, 1e20, 2.0 ** 23, np.nan, np.inf, 2.5]),
    }
    for name, (dt, bads) in cases.items():
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        src = good.astype(dt)
        if dt in (np.float64, np.float32):
            src[7] = -0.0                                        # minus zero is the integer 0
        out = np.zeros(n, dtype=np.uint16)
        assert fn(src.ctypes.data, out.ctypes.data, n) == 1, name
        assert np.array_equal(out, src.astype(np.int64).astype(np.uint16)), name
        for pos in (0, n // 2, n - 1):
            for bad in bads:
                b = src.copy()
                b[pos] = bad
                assert fn(b.ctypes.data, out.ctypes.data, n) == 0, (name, bad, pos)
