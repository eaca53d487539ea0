"""GPU suite (-m gpu): parity of the CUDA path, called through the C ABI, against the CPU oracle.
Bars: pair counts, classes and contingency tables bit-exact; p-values and statistics within 1e-12 relative;
identical up/down/no-change calls and iteration counts."""
import json
import os

import numpy as np
import pytest
from conftest import GOLDEN, small_case

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    e[(a == b)] = 0.0
    return float(np.nanmax(e)) if e.size else 0.0


def stat_close(a, b, atol=4e-15):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    bad = np.abs(a - b) > RTOL * np.abs(b) + atol
    assert not bad.any(), (a[bad][:5], b[bad][:5])


# ---- K1 + pair counts -------------------------------------------------------------------------
@pytest.mark.parametrize("seed,r,n1,n2,n3,scale", [
    (1, 70, 5, 5, 0, 4), (2, 200, 33, 31, 0, 8), (3, 130, 64, 1, 0, 2), (4, 90, 7, 40, 9, 8), (5, 65, 100, 100, 0, 16)])
def test_pair_counts_bit_exact(reo, oracle, seed, r, n1, n2, n3, scale):
    data, group = small_case(seed, r, n1, n2, scale=scale, n3=n3)
    levels, gid = oracle.group_levels(group)
    gnum = len(levels)
    info = reo.stage(data, gid, gnum)
    assert info["gene_tiles"] == -(-r // 64)
    rng = np.random.default_rng(seed)
    rows = rng.choice(r, min(r, 40), replace=False)
    cols = rng.choice(r, min(r, 50), replace=False)
    want = oracle.greater_counts(data, gid, gnum, rows, cols, seed=7)
    for k in range(gnum if gnum > 2 else 1):
        nre, rest = reo.pair_counts(k, rows, cols)
        assert np.array_equal(nre, want[k])
        assert np.array_equal(rest, want.sum(axis=0) - want[k])


@pytest.mark.parametrize("dtype", [np.int64, np.float64, np.int32, np.float32])
def test_dtypes_and_leading_dimension(reo, pkg, oracle, dtype):
    data, group = small_case(9, 100, 10, 12)
    levels, gid = oracle.group_levels(group)
    rows = np.arange(0, 100, 3)
    want = oracle.greater_counts(data, gid, 2, rows, rows, seed=7)
    reo.stage(data.astype(dtype), gid, 2)
    nre, rest = reo.pair_counts(0, rows, rows)
    assert np.array_equal(nre, want[0]) and np.array_equal(rest, want[1])


def test_wide_value_range_uses_sort_fallback(reo, oracle):
    """Columns whose value range exceeds the shared-memory bitmap go through the sort-based rank kernel."""
    data, group = small_case(12, 300, 9, 8)
    data = data.astype(np.int64)
    data[::7, :] *= 3_000_000          # range > 1.3 M in every column
    data[5, 3] = -(2 ** 40)
    levels, gid = oracle.group_levels(group)
    reo.stage(data, gid, 2)
    rows = np.arange(0, 300, 5)
    want = oracle.greater_counts(data, gid, 2, rows, rows, seed=7)
    nre, rest = reo.pair_counts(0, rows, rows)
    assert np.array_equal(nre, want[0]) and np.array_equal(rest, want[1])


def float_case(seed, r, n1, n2, n3=0):
    """log-scale expression with many values inside the 0.1 tie band (non-transitive ties, src:72)."""
    data, group = small_case(seed, r, n1, n2, scale=1, n3=n3)
    rng = np.random.default_rng(seed)
    x = np.log1p(data) + rng.normal(0, 0.04, data.shape)
    x[rng.random(data.shape) < 0.1] = 0.0
    return np.round(x, 3), group


@pytest.mark.parametrize("seed,r,n1,n2,n3,dtype", [(1, 90, 7, 9, 0, np.float64), (2, 150, 40, 33, 0, np.float64),
                                                   (3, 70, 12, 30, 9, np.float64), (4, 100, 20, 20, 0, np.float32)])
def test_float_path_pair_counts_and_tables(reo, oracle, coracle, seed, r, n1, n2, n3, dtype):
    """non-integral input goes through the raw-value FP64 compare path, bit-exact vs the oracle."""
    data, group = float_case(seed, r, n1, n2, n3)
    data = data.astype(dtype)
    levels, gid = oracle.group_levels(group)
    gnum = len(levels)
    info = reo.stage(data, gid, gnum)
    assert info["rank_bits"] == 0  # float mode: no rank planes
    rows = np.arange(0, r, 2)
    cols = np.arange(1, r, 3)
    want = oracle.greater_counts(data.astype(np.float64), gid, gnum, rows, cols, seed=7)
    for k in range(gnum if gnum > 2 else 1):
        nre, rest = reo.pair_counts(k, rows, cols)
        assert np.array_equal(nre, want[k]) and np.array_equal(rest, want.sum(axis=0) - want[k])
    thr = coracle.thresholds_for(gid, gnum, 0.01)
    mask = np.random.default_rng(seed).random(r) < 0.5
    for k in range(gnum if gnum > 2 else 1):
        tab, _ = coracle.block_tables(data.astype(np.float64), gid, gnum, thr, np.nonzero(mask)[0], seed=7, k=k)
        assert np.array_equal(reo.tables(k, mask, thresholds=thr), tab)
    # all genes as references + incremental update
    full = np.ones(r, bool)
    tab, _ = coracle.block_tables(data.astype(np.float64), gid, gnum, thr, np.arange(r), seed=7)
    assert np.array_equal(reo.tables(0, mask, thresholds=thr, mask_to=full), tab)


def test_float_path_identify_degs(reo, oracle, coracle):
    data, group = float_case(8, 400, 25, 30)
    levels, gid = oracle.group_levels(group)
    ref = np.arange(400) % 4 == 0
    thr = coracle.thresholds_for(gid, 2, 0.01)
    want = coracle.identify_degs(data, gid, 2, thr, 1.0, 0.05, ref, 128, 5, seed=7)
    out = reo.identify_degs(data, gid, 2, ref, 0.01, 1.0, 0.05, 128, 5)
    check_full(out, want)


# ---- K2 tables ----------------------------------------------------------------------------------
@pytest.mark.parametrize("seed,r,n1,n2,frac", [(1, 70, 5, 5, 0.3), (2, 333, 33, 31, 0.2), (3, 200, 70, 3, 1.0),
                                               (4, 129, 20, 45, 0.02), (6, 1000, 12, 12, 0.5)])
def test_tables_bit_exact_and_delta(reo, oracle, coracle, seed, r, n1, n2, frac):
    data, group = small_case(seed, r, n1, n2)
    levels, gid = oracle.group_levels(group)
    thr = coracle.thresholds_for(gid, 2, 0.01)
    rng = np.random.default_rng(seed)
    mask = rng.random(r) < frac
    if frac >= 1.0:
        mask[:] = True
    reo.stage(data, gid, 2)
    got = reo.tables(0, mask, thresholds=thr)
    want, _ = coracle.block_tables(data, gid, 2, thr, np.nonzero(mask)[0], seed=7)
    assert np.array_equal(got, want)
    assert np.array_equal(got.sum(axis=1), mask.sum() - mask.astype(int))  # gene i never counts itself
    # thresholds computed by the library from pval_reo
    assert np.array_equal(reo.tables(0, mask, pval_reo=0.01), want)
    # incremental path: build for `mask`, signed update over the symmetric difference to `mask2`
    mask2 = mask.copy()
    flip = rng.choice(r, max(r // 20, 1), replace=False)
    mask2[flip] = ~mask2[flip]
    want2, _ = coracle.block_tables(data, gid, 2, thr, np.nonzero(mask2)[0], seed=7)
    assert np.array_equal(reo.tables(0, mask, thresholds=thr, mask_to=mask2), want2)


def test_tables_three_groups_one_vs_rest(reo, oracle, coracle):
    data, group = small_case(21, 150, 9, 40, n3=33)
    levels, gid = oracle.group_levels(group)
    thr = coracle.thresholds_for(gid, 3, 0.01)
    mask = np.random.default_rng(0).random(150) < 0.4
    reo.stage(data, gid, 3)
    for k in range(3):
        want, _ = coracle.block_tables(data, gid, 3, thr, np.nonzero(mask)[0], seed=7, k=k)
        assert np.array_equal(reo.tables(k, mask, thresholds=thr), want), k


def test_tables_mirror_property_all_genes(reo, oracle, coracle):
    """With every gene a reference, each unordered pair contributes q to one gene and 10-q to the other
    (src:385-386), so the column sums of the table are palindromic."""
    data, group = small_case(8, 500, 40, 50)
    levels, gid = oracle.group_levels(group)
    reo.stage(data, gid, 2)
    tab = reo.tables(0, np.ones(500, bool))
    s = tab.sum(axis=0)
    assert np.array_equal(s, s[::-1]) and tab.sum() == 500 * 499


# ---- K3..K6 ----------------------------------------------------------------------------------------
def test_mccullagh_kat_on_device(reo, pkg):
    kat = json.load(open(os.path.join(GOLDEN, "kat_mccullagh.json")))
    got = pkg.McCullagh_test(np.array(kat["mat"]), handle=reo)
    # sqrt/div/mul/add are IEEE on the device; log and erfc may differ from libm in the last ulp
    assert rel_err(got, kat["expected"]) <= RTOL
    with pytest.raises(ValueError, match="square"):
        pkg.McCullagh_test(np.zeros((3, 4), dtype=int), handle=reo)


def test_mccullagh_random_tables(reo, coracle):
    rng = np.random.default_rng(5)
    t = rng.integers(0, 3000, size=(2000, 3, 3))
    t[:50, 0, 1] = t[:50, 1, 0] = t[:50, 1, 2] = t[:50, 2, 1] = 0  # a = b = c: singular in exact arithmetic
    t[50:60] = 0
    got = reo.mccullagh(t)
    want = np.array([coracle.mccullagh(x) for x in t])
    stat_close(got[:, 1:], want[:, 1:])
    assert rel_err(got[:, 0], want[:, 0]) <= 1e-11  # the test's own p (overwritten in the pipeline, src:415)


@pytest.mark.parametrize("n", [11, 30, 2047, 2049, 19999, 40000])
def test_sort_empirical_null_bh(reo, oracle, coracle, n):
    rng = np.random.default_rng(n)
    d = rng.normal(0, 1.5, n)
    d[rng.integers(0, n, n // 7)] = 0.0
    d[rng.integers(0, n, 3)] = -0.0
    s, perm = reo.sort(d)
    assert np.array_equal(s, np.sort(d)) and np.array_equal(d[perm], s)
    assert np.all((np.diff(s) > 0) | (np.diff(perm) > 0) | (np.signbit(s[:-1]) & ~np.signbit(s[1:])))  # stable
    se_w, p_w = coracle.empirical_null(d)
    se_g, p_g = reo.empirical_null(d)
    assert se_g == se_w, (se_g, se_w)  # same reduction tree, IEEE ops: bit-exact
    assert rel_err(p_g, p_w) <= RTOL
    q_g = reo.bh(p_w)
    assert np.array_equal(q_g, coracle.bh(p_w))


def test_small_r_bounds_error(reo):
    with pytest.raises(IndexError):
        reo.empirical_null(np.zeros(10))


# ---- the whole path -------------------------------------------------------------------------------
def check_full(out, want):
    assert out.iters == want["iters"]
    assert np.array_equal(out.result[:, :, 2:11], want["result"][:, :, 2:11]), "contingency tables"
    assert np.array_equal(out.updown, want["updown"]), "up/down/no-change calls"
    assert np.array_equal(out.final_ref, want["final_ref"])
    assert rel_err(out.result[:, :, 0], want["result"][:, :, 0]) <= RTOL  # pval
    assert rel_err(out.result[:, :, 1], want["result"][:, :, 1]) <= RTOL  # padj
    # d1 d2 se z1: relative 1e-12, with an absolute floor of a few ulp of O(1): d1 = w1.log-odds can cancel
    # exactly in one libm and to ~1e-17 in the other (CUDA log vs glibc log differ in the last ulp)
    stat_close(out.result[:, :, 11:15], want["result"][:, :, 11:15])


@pytest.mark.parametrize("seed,r,n1,n2,n3,nref", [(1, 150, 7, 9, 0, 40), (2, 600, 33, 40, 0, 100),
                                                  (3, 300, 6, 5, 7, 80), (4, 1200, 20, 20, 0, 300)])
def test_identify_degs_matches_oracle(reo, oracle, coracle, seed, r, n1, n2, n3, nref):
    data, group = small_case(seed, r, n1, n2, n3=n3)
    levels, gid = oracle.group_levels(group)
    gnum = len(levels)
    ref = np.zeros(r, bool)
    ref[np.random.default_rng(seed).choice(r, nref, replace=False)] = True
    thr = coracle.thresholds_for(gid, gnum, 0.01)
    want = coracle.identify_degs(data, gid, gnum, thr, 1.0, 0.05, ref, 128, 5, seed=7)
    out = reo.identify_degs(data, gid, gnum, ref, 0.01, 1.0, 0.05, 128, 5)
    check_full(out, want)
    assert out.stats["n_deg"] == [int(v) for v in want["deg_log"][-1 if gnum > 2 else 0][:out.iters[-1]]]
    # n_iter cap and explicit thresholds
    want2 = coracle.identify_degs(data, gid, gnum, thr, 0.5, 0.1, ref, 2, 0, seed=7)
    out2 = reo.identify_degs(data, gid, gnum, ref, 0.01, 0.5, 0.1, 2, 0, thresholds=thr)
    check_full(out2, want2)


def test_words_with_an_empty_top_plane_skip_it(reo, oracle, coracle):
    """One sample with 300 distinct values sets B = 9 for the whole matrix; every other sample has a handful, so all words
    but the first have empty top planes and the pair kernel stops their borrow chains one plane early
    (reo_stats.planes_per_word < B + 1).  Results must not change: the oracle knows nothing about planes."""
    rng = np.random.default_rng(123)
    r, n1, n2 = 300, 70, 60
    mu = np.exp(rng.normal(0.3, 0.8, r))
    data = rng.poisson(mu[:, None] * np.ones((1, n1 + n2))).astype(np.int64)
    data[:25, n1:] += 2
    data[:, 5] = rng.permutation(r)                       # the one wide sample (first word of the first group)
    gid = np.array([0] * n1 + [1] * n2, dtype=np.int32)
    ref = np.arange(r) % 3 != 0
    thr = coracle.thresholds_for(gid, 2, 0.01)
    want = coracle.identify_degs(data, gid, 2, thr, 1.0, 0.05, ref, 128, 5, seed=7)
    out = reo.identify_degs(data, gid, 2, ref, 0.01, 1.0, 0.05, 128, 5)
    check_full(out, want)
    assert out.stats["rank_bits"] == 9
    assert out.stats["sample_words"] == 5
    assert abs(out.stats["planes_per_word"] - (10 + 4 * 9) / 5) < 1e-9
    # all-genes build (symmetric sweep) on the same staged planes
    tab, _ = coracle.block_tables(data, gid, 2, thr, np.arange(r), seed=7)
    assert np.array_equal(reo.tables(0, np.ones(r, bool), thresholds=thr), tab)


def test_identify_degs_reference_signature(reo, pkg, oracle, coracle):
    """The reference-named wrapper returns the r x (1+16K) matrix of src:430/437."""
    data, group = small_case(5, 120, 8, 8)
    names = [f"g{i}" for i in range(120)]
    ref = np.arange(120) % 3 == 0
    res = pkg.identify_degs(data, group, names, 0.01, 1.0, 0.05, ref, 128, 5, handle=reo)
    assert res.shape == (120, 17) and list(res[:, 0]) == names
    assert set(res[:, 16]) <= {"up", "down", "no change"}
    levels, gid = oracle.group_levels(group)
    want = coracle.identify_degs(data, gid, 2, coracle.thresholds_for(gid, 2, 0.01), 1.0, 0.05, ref, 128, 5, seed=7)
    assert np.array_equal(res[:, 3:12].astype(np.int64), want["result"][0][:, 2:11].astype(np.int64))


def test_golden_small_case_gpu(reo):
    g = np.load(os.path.join(GOLDEN, "small_case.npz"))
    out = reo.identify_degs(g["data"], g["gid"], 2, g["ref"], 0.01, 1.0, 0.05, 128, 5)
    check_full(out, dict(result=g["result"], updown=g["updown"], final_ref=g["final_ref"], iters=g["iters"].tolist()))


def test_golden_bundled_data_gpu(reo):
    """Config 1: the reference's bundled test data (19999 x (5+5)), seeded 3000-gene reference mask."""
    g = np.load(os.path.join(GOLDEN, "bundled_c1.npz"))
    out = reo.identify_degs(g["data"], g["gid"], 2, g["ref"], 0.01, 1.0, 0.05, 128, 5)
    assert out.iters == g["iters"].tolist()
    assert np.array_equal(out.result[0][:, 2:11].astype(np.int32), g["tables"])
    assert np.array_equal(out.updown[0], g["updown"]) and np.array_equal(out.final_ref[0], g["final_ref"])
    assert rel_err(out.result[0][:, 0], g["pval"]) <= RTOL and rel_err(out.result[0][:, 1], g["padj"]) <= RTOL
    stat_close(out.result[0][:, 11:15], g["stat"])


def test_tie_free_input_is_seed_independent(pkg, oracle, coracle):
    """On tie-free data no coin is consulted: any seed gives the same result (= the real reference's)."""
    data, group, is_de = pkg.synth.bulk(400, 12, 14, seed=2)
    tf = pkg.synth.tie_free(data)
    levels, gid = oracle.group_levels(group)
    ref = pkg.synth.reference_mask(is_de, 80)
    outs = []
    for seed in (1, 99):
        with pkg.Reo(0, seed=seed) as h:
            outs.append(h.identify_degs(tf, gid, 2, ref, 0.01, 1.0, 0.05, 128, 5))
    assert np.array_equal(outs[0].result, outs[1].result)
    want = coracle.identify_degs(tf, gid, 2, coracle.thresholds_for(gid, 2, 0.01), 1.0, 0.05, ref, 128, 5, seed=12345)
    check_full(outs[0], want)


def test_full_size_properties_config2(reo, pkg, oracle, coracle):
    """BASELINE config 2 at full size (20k genes x 100 vs 100): size-independent properties + a sampled
    row block against the oracle."""
    data, group, is_de = pkg.synth.bulk(20000, 100, 100)
    levels, gid = oracle.group_levels(group)
    ref = pkg.synth.reference_mask(is_de, 3000)
    out = reo.identify_degs(data, gid, 2, ref, 0.01, 1.0, 0.05, 128, 5)
    tab = out.result[0][:, 2:11].astype(np.int64)
    fr = out.final_ref[0].astype(bool)
    assert np.array_equal(tab.sum(axis=1), fr.sum() - fr.astype(int))
    thr = coracle.thresholds_for(gid, 2, 0.01)
    cols = np.nonzero(fr)[0]
    want, _ = coracle.block_tables(data, gid, 2, thr, cols, seed=7, i0=4000, i1=4032)
    assert np.array_equal(tab[4000:4032], want)
    # statistics recomputed by the oracle from the device tables
    res = out.result[0]
    se_w, p_w = coracle.empirical_null(res[:, 11])
    assert rel_err(res[:, 0], p_w) <= RTOL and rel_err(res[:, 1], coracle.bh(p_w)) <= RTOL
    sig = (res[:, 0] <= 1.0) & (res[:, 1] <= 0.05)
    assert np.array_equal(out.updown[0], np.where(sig & (res[:, 14] > 0), 1, np.where(sig & (res[:, 14] < 0), -1, 0)))
    # the planted DE genes are recovered
    called = out.updown[0] != 0
    assert (called & is_de).sum() > 0.7 * is_de.sum() and (called & ~is_de).sum() < 0.1 * called.sum() + 50
    assert out.stats["kernel_launches"] > 0 and out.stats["compares"] > 0


# ---- several GPUs in one process (the Julia deployment shape) ------------------------------------
def test_single_process_multi_gpu_matches_single_gpu(reo, pkg, oracle, coracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    data, group = small_case(31, 700, 40, 45)
    levels, gid = oracle.group_levels(group)
    ref = np.arange(700) % 5 == 0
    thr = coracle.thresholds_for(gid, 2, 0.01)
    want = coracle.identify_degs(data, gid, 2, thr, 1.0, 0.05, ref, 128, 5, seed=7)
    ndev = min(torch.cuda.device_count(), 4)
    for shard_min in ("0", None):   # "0": K1 (copy + rank + bit-planes) sharded over the devices as well
        if shard_min is None:
            os.environ.pop("REO_K1_SHARD_MIN", None)
        else:
            os.environ["REO_K1_SHARD_MIN"] = shard_min
        try:
            with pkg.Reo(list(range(ndev)), seed=pkg.synth.TIE_SEED) as h:
                out = h.identify_degs(data, gid, 2, ref, 0.01, 1.0, 0.05, 128, 5)
                check_full(out, want)
                h.stage(data, gid, 2)
                mask = np.arange(700) % 3 != 0
                tab, _ = coracle.block_tables(data, gid, 2, thr, np.nonzero(mask)[0], seed=7)
                assert np.array_equal(h.tables(0, mask, thresholds=thr), tab)
                # float path and three levels through the sharded staging
                fdata, fgroup = float_case(5, 300, 20, 25, 9)
                flev, fgid = oracle.group_levels(fgroup)
                fthr = coracle.thresholds_for(fgid, 3, 0.01)
                fref = np.arange(300) % 4 == 0
                fwant = coracle.identify_degs(fdata, fgid, 3, fthr, 1.0, 0.05, fref, 128, 5, seed=7)
                check_full(h.identify_degs(fdata, fgid, 3, fref, 0.01, 1.0, 0.05, 128, 5), fwant)
                # fewer sample words than devices: 10 + 12 samples share ONE word, every other device stages nothing
                sdata, sgroup = small_case(77, 300, 10, 12)
                slev, sgid = oracle.group_levels(sgroup)
                sthr = coracle.thresholds_for(sgid, 2, 0.01)
                sref = np.arange(300) % 3 == 0
                swant = coracle.identify_degs(sdata, sgid, 2, sthr, 1.0, 0.05, sref, 128, 5, seed=7)
                check_full(h.identify_degs(sdata, sgid, 2, sref, 0.01, 1.0, 0.05, 128, 5), swant)
        finally:
            os.environ.pop("REO_K1_SHARD_MIN", None)


def test_iter_log_per_level_and_input_memory_kinds(reo, pkg, oracle, coracle):
    """Three levels: reo_iter_log returns every level's log (what the Julia shim prints, src:418-420, 432-435), and the
    result does not depend on where the input lives: pageable numpy memory (pinned bounce pipeline), page-locked
    memory (direct DMA) or HBM."""
    import torch
    data, group = small_case(21, 400, 12, 20, n3=17)
    levels, gid = oracle.group_levels(group)
    ref = np.arange(400) % 3 == 0
    thr = coracle.thresholds_for(gid, 3, 0.01)
    want = coracle.identify_degs(data, gid, 3, thr, 1.0, 0.05, ref, 128, 5, seed=7)
    out = reo.identify_degs(data, gid, 3, ref, 0.01, 1.0, 0.05, 128, 5)                      # pageable
    check_full(out, want)
    for k in range(3):
        log = reo.iter_log(k)
        assert log["iters_done"] == want["iters"][k] == len(log["n_deg"])
        assert log["n_deg"] == [int(v) for v in want["deg_log"][k][:log["iters_done"]]]
    pinned = torch.from_numpy(np.ascontiguousarray(data.T)).pin_memory()                     # [c, r] == r x c column-major
    out_p = reo.identify_degs(pinned.numpy().T, gid, 3, ref, 0.01, 1.0, 0.05, 128, 5)
    dev = pinned.cuda()
    dm = pkg.DeviceMatrix(dev.data_ptr(), pkg._lib.REO_I64, 400, data.shape[1], 400, keepalive=dev)
    out_d = reo.identify_degs(dm, gid, 3, ref, 0.01, 1.0, 0.05, 128, 5)
    for o in (out_p, out_d):
        assert np.array_equal(o.result, out.result) and np.array_equal(o.updown, out.updown)


def test_pageable_input_is_narrowed_chunk_by_chunk(reo, oracle, coracle):
    """Pageable host matrices are narrowed to u16 by the copy threads, chunk by chunk (csrc/reo_host.cpp): a chunk with a
    value above 65535 travels raw, and a matrix with a non-integral value in one chunk only is staged again without
    narrowing (float path).  Every variant must give the tables of the same matrix in page-locked memory, and the
    oracle's."""
    import torch
    rng = np.random.default_rng(77)
    r, n1, n2 = 1200, 1500, 1500                     # 2 copy chunks of <= 1628 columns
    mu = np.exp(rng.normal(0.5, 1.2, r))
    base = rng.poisson(mu[:, None] * np.ones((1, n1 + n2))).astype(np.int64)
    base[:40, n1:] += 3
    gid = np.array([0] * n1 + [1] * n2, dtype=np.int32)
    thr = coracle.thresholds_for(gid, 2, 0.01)
    mask = np.zeros(r, bool)
    mask[rng.choice(r, 24, replace=False)] = True

    def tables_of(mat):
        reo.stage(mat, gid, 2)
        return reo.tables(0, mask, thresholds=thr)

    def pinned_copy(mat):
        t = torch.from_numpy(np.ascontiguousarray(mat.T)).pin_memory()
        return t, t.numpy().T

    cases = []
    wide = base.copy(); wide[17, 2500] = 70000      # second chunk cannot be narrowed
    cases += [base, wide, base.astype(np.int32), base.astype(np.float64), base.astype(np.float32)]
    frac = base.astype(np.float64); frac[33, 2900] += 0.25      # non-integral value in the second chunk only
    cases.append(frac)
    for mat in cases:
        got = tables_of(np.asfortranarray(mat))                 # pageable
        keep, pm = pinned_copy(mat)
        assert np.array_equal(got, tables_of(pm))
        want, _ = coracle.block_tables(mat, gid, 2, thr, np.nonzero(mask)[0], seed=7)
        assert np.array_equal(got, want)


def test_more_than_65535_samples_per_group_wide_counters(reo, oracle, coracle):
    """> 65 535 sample slots in a group: the pair kernel's packed 16-bit counters do not fit and the 32-bit variant runs."""
    rng = np.random.default_rng(9)
    r, n1, n2 = 70, 66000, 300
    mu = np.exp(rng.normal(1.0, 1.0, r))
    data = rng.poisson(mu[:, None] * np.ones((1, n1 + n2))).astype(np.int32)
    data[:10, n1:] += 2
    gid = np.array([0] * n1 + [1] * n2, dtype=np.int32)
    thr = coracle.thresholds_for(gid, 2, 0.01)
    mask = np.arange(r) % 5 != 0
    reo.stage(data, gid, 2)
    tab, _ = coracle.block_tables(data, gid, 2, thr, np.nonzero(mask)[0], seed=7)
    assert np.array_equal(reo.tables(0, mask, thresholds=thr), tab)
    tab_all, _ = coracle.block_tables(data, gid, 2, thr, np.arange(r), seed=7)
    assert np.array_equal(reo.tables(0, np.ones(r, bool), thresholds=thr), tab_all)


def test_float32_differences_are_rounded_to_float32(reo, oracle, coracle):
    """Matrix{Float32}: Julia evaluates abs(x - y) in Float32 and compares the result with the Float64 literal 0.1
    (src:72), so 0.07f0 - (-0.03f0) = 0.1f0 is NOT a tie although the exact difference is below 0.1.  Values on a 0.01
    grid put thousands of pairs on the edge of the band."""
    rng = np.random.default_rng(12)
    r, n1, n2 = 90, 21, 19
    data = (rng.integers(-20, 21, size=(r, n1 + n2)) / 100.0).astype(np.float32)
    data[:12, n1:] += np.float32(0.2)
    gid = np.array([0] * n1 + [1] * n2, dtype=np.int32)
    rows = np.arange(r)
    want32 = oracle.greater_counts(data, gid, 2, rows, rows, seed=7)
    want64 = oracle.greater_counts(data.astype(np.float64), gid, 2, rows, rows, seed=7)
    assert not np.array_equal(want32, want64)          # the two semantics really differ on this matrix
    reo.stage(data, gid, 2)
    nre, rest = reo.pair_counts(0, rows, rows)
    assert np.array_equal(nre, want32[0]) and np.array_equal(rest, want32[1])
    thr = coracle.thresholds_for(gid, 2, 0.01)
    mask = np.arange(r) % 4 != 0
    tab, _ = coracle.block_tables(data, gid, 2, thr, np.nonzero(mask)[0], seed=7)          # C oracle: float32 semantics too
    assert np.array_equal(reo.tables(0, mask, thresholds=thr), tab)
    reo.stage(data.astype(np.float64), gid, 2)           # the same values as Matrix{Float64}: exact differences
    nre64, rest64 = reo.pair_counts(0, rows, rows)
    assert np.array_equal(nre64, want64[0]) and np.array_equal(rest64, want64[1])


def test_subset_and_detect_more_than_65535_cells(reo):
    """gridDim.y is capped at 65535: the kernels next to the path stride over the cell dimension."""
    rng = np.random.default_rng(3)
    data = rng.integers(0, 3, size=(40, 70001)).astype(np.int32)
    genes = np.arange(0, 40, 3)
    cells = np.arange(70001)[::-1].copy()
    out, _ = reo.subset(data, genes, cells)
    assert np.array_equal(out, data[np.ix_(genes, cells)])
    per_cell, per_gene = reo.detect_counts(data)
    assert np.array_equal(per_cell, (data > 0).sum(axis=0)) and np.array_equal(per_gene, (data > 0).sum(axis=1))


def test_no_threshold_is_an_error(reo):
    """pval_reo >= 1: src:85 finds no count -- the reference throws, the library reports a bad argument."""
    data, group = small_case(5, 60, 6, 7)
    gid = np.array([0 if g == "a" else 1 for g in group], dtype=np.int32)
    with pytest.raises(ValueError):
        reo.identify_degs(data, gid, 2, np.ones(60, bool), 1.0, 1.0, 0.05, 4, 1)


# ---- the UNMODIFIED reference's own outputs (tie-free inputs), when a maintainer with Julia has dumped them -------
def test_reference_julia_fixtures(reo, oracle):
    """tests/golden/julia_out/<case>.tsv = RankCompV3.identify_degs run by julia/dump_reference_fixture.jl on the inputs
    of scripts/make_tiefree_inputs.py.  There is no Julia in the build image: missing fixtures are reported as xfail."""
    import julia_fixtures as jf
    have = [c for c in jf.cases() if jf.load_reference_output(c) is not None]
    if not have:
        pytest.xfail("reference fixtures missing: run julia/dump_reference_fixture.jl where Julia is available")
    for case in have:
        genes, data, group, ref, (pval_reo, pval_deg, padj_deg, n_iter, n_conv) = jf.load_input(case)
        levels, gid = oracle.group_levels(group)
        out = reo.identify_degs(data, gid, len(levels), ref, pval_reo, pval_deg, padj_deg, n_iter, n_conv)
        want_result, want_updown = jf.load_reference_output(case)
        jf.compare(out.result, out.updown, want_result, want_updown)


# ---- one process per GPU (the torchrun deployment shape): NCCL inside the library ------------------------
def test_one_process_per_gpu_matches_oracle(tmp_path):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import mp_worker
    mp.spawn(mp_worker.run, args=(world, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ok_{q}")) for q in range(world))


# ---- next to the path: pseudo-bulk, detection filters, subsetting ---------------
@pytest.mark.parametrize("dtype", [np.int64, np.int32, np.float64, np.float32])
def test_pseudobulk_detect_subset(reo, pkg, dtype):
    rng = np.random.default_rng(3)
    r, c, n_pseudo = 333, 207, 10
    x = rng.poisson(0.6, size=(r, c)).astype(dtype)
    if np.issubdtype(dtype, np.floating):
        x = (x * 1.25).astype(dtype)
    # src:60-63: shuffle, partition into chunks of ceil(c/n_pseudo)
    perm = rng.permutation(c)
    cp = -(-c // n_pseudo)
    profiles = [perm[i:i + cp] for i in range(0, c, cp)]
    host, dm = reo.pseudobulk(x, profiles)
    want = np.stack([x[:, p].astype(np.float64 if np.issubdtype(dtype, np.floating) else np.int64).sum(axis=1)
                     for p in profiles], axis=1)
    if np.issubdtype(dtype, np.floating):
        assert np.allclose(host, want, rtol=1e-13)
        seq = np.stack([np.add.reduce(x[:, p].astype(np.float64), axis=1) for p in profiles], axis=1)
        assert host.shape == seq.shape
    else:
        assert np.array_equal(host, want)
    assert dm.r == r and dm.c == len(profiles)
    per_cell, per_gene = reo.detect_counts(x)
    assert np.array_equal(per_cell, (x > 0).sum(axis=0)) and np.array_equal(per_gene, (x > 0).sum(axis=1))
    genes = np.nonzero(per_gene > 40)[0]
    cells = np.nonzero(per_cell > 120)[0]
    sub, dms = reo.subset(x, genes, cells)
    assert np.array_equal(sub, x[np.ix_(genes, cells)])
    # chained on the device: pseudo-bulk output -> identify_degs without leaving HBM
    if dtype == np.int64:
        gid = np.array([0] * (len(profiles) // 2) + [1] * (len(profiles) - len(profiles) // 2), dtype=np.int32)
        ref = np.arange(r) % 3 == 0
        a = reo.identify_degs(dm, gid, 2, ref, 0.01, 1.0, 0.05, 16, 2)
        b = reo.identify_degs(host, gid, 2, ref, 0.01, 1.0, 0.05, 16, 2)
        assert np.array_equal(a.result, b.result) and np.array_equal(a.updown, b.updown)


def test_reoa_driver_end_to_end(reo, pkg, tmp_path):
    """reoa() mirror (src:536-685): files in, TSVs out, same calls as a direct identify_degs."""
    import pandas as pd
    data, group, is_de = pkg.synth.bulk(600, 9, 11, seed=4)
    data[5, :] = 0  # an all-zero gene row is dropped by the > min_features filter (src:626)
    cols = [f"S{i}" for i in range(20)]
    expr = pd.DataFrame(data, columns=cols)
    expr.insert(0, "gene_name", [f"G{i}" for i in range(600)])
    expr.to_csv(tmp_path / "fn_expr.txt", sep="\t", index=False)
    pd.DataFrame({"sample_name": cols, "group": group}).to_csv(tmp_path / "fn_meta.txt", sep="\t", index=False)
    out = pkg.reoa("fn_expr.txt", "fn_meta.txt", use_hk_genes="no", ref_gene_max=150, work_dir=str(tmp_path), seed=3,
                   handle=reo)
    assert list(out.columns) == ["gene_name", "group1_vs_group2"] and len(out) == 599
    assert set(out["group1_vs_group2"]) <= {"up", "down", "no change"}
    for f in ("fn_expr_group1_group2_result.tsv", "fn_expr_df_expr.tsv", "fn_expr_df_meta.tsv", "fn_expr_gene_up_down.tsv"):
        assert (tmp_path / f).exists(), f
    res = pd.read_csv(tmp_path / "fn_expr_group1_group2_result.tsv", sep="\t")
    assert list(res.columns) == ["genename"] + pkg.api.HEADER and len(res) == 599
    called = out["group1_vs_group2"].to_numpy() != "no change"
    keep = np.ones(600, bool); keep[5] = False
    assert (called & is_de[keep]).sum() > 0.5 * called.sum()
    # pseudo-bulk mode (src:608-612): 20 cells -> 2 x 3 profiles
    out2 = pkg.reoa("fn_expr.txt", "fn_meta.txt", use_hk_genes="no", ref_gene_max=150, work_dir=str(tmp_path), seed=3,
                    n_pseudo=3, handle=reo, write_files=False)
    assert len(out2) <= 600 and list(out2.columns)[1] == "group1_vs_group2"
    with pytest.raises(ValueError, match="ArgumentError"):
        pkg.reoa("nope.txt", "fn_meta.txt", work_dir=str(tmp_path), handle=reo)


@pytest.mark.parametrize("n1,n2", [(600, 700), (640, 1000), (1500, 33)])
def test_many_samples_compare_path(reo, oracle, coracle, n1, n2):
    """> ~1000 samples: the class lookup tables do not fit and the kernel classifies by comparisons."""
    data, group = small_case(41, 140, n1, n2, scale=3)
    levels, gid = oracle.group_levels(group)
    thr = coracle.thresholds_for(gid, 2, 0.01)
    reo.stage(data, gid, 2)
    mask = np.arange(140) % 3 != 1
    tab, _ = coracle.block_tables(data, gid, 2, thr, np.nonzero(mask)[0], seed=7)
    assert np.array_equal(reo.tables(0, mask, thresholds=thr), tab)
    rows = np.arange(0, 140, 7)
    want = oracle.greater_counts(data, gid, 2, rows, rows, seed=7)
    nre, rest = reo.pair_counts(0, rows, rows)
    assert np.array_equal(nre, want[0]) and np.array_equal(rest, want[1])


def test_full_size_properties_config4(reo, pkg, oracle, coracle):
    """BASELINE config 4 at full size (30k genes x 10k vs 10k cells, 3000 initial references): properties plus
    a row block checked against the oracle through a sub-matrix with global gene indices."""
    import torch
    dev, group, is_de = pkg.synth.scrna_torch(30000, 10000, 10000)
    r, c = 30000, 20000
    levels, gid = oracle.group_levels(group)
    ref = pkg.synth.reference_mask(is_de, 3000)
    dm = pkg.DeviceMatrix(dev.data_ptr(), pkg._lib.REO_I64, r, c, r, keepalive=dev)
    out = reo.identify_degs(dm, gid, 2, ref, 0.01, 1.0, 0.05, 128, 5)
    tab = out.result[0][:, 2:11].astype(np.int64)
    fr = out.final_ref[0].astype(bool)
    assert np.array_equal(tab.sum(axis=1), fr.sum() - fr.astype(int))
    assert out.stats["compares"] >= 30000 * 3000 * 20000
    # oracle on a sub-matrix: 24 row genes + 400 of the final reference genes, global indices kept
    rng = np.random.default_rng(0)
    rows_g = np.sort(rng.choice(r, 24, replace=False))
    cols_g = np.sort(rng.choice(np.nonzero(fr)[0], 400, replace=False))
    genes = np.union1d(rows_g, cols_g)
    sub = dev[:, torch.as_tensor(genes, device=dev.device)].cpu().numpy().T.astype(np.int64)   # [genes, c]
    thr = coracle.thresholds_for(gid, 2, 0.01)
    want = coracle.block_tables_idx(sub, genes, gid, 2, thr, np.searchsorted(genes, rows_g), np.searchsorted(genes, cols_g),
                                    seed=7)
    mask = np.zeros(r, bool); mask[cols_g] = True
    got = reo.tables(0, mask, thresholds=thr)   # the matrix staged by identify_degs is still resident
    assert np.array_equal(got[rows_g], want)
    res = out.result[0]
    se_w, p_w = coracle.empirical_null(res[:, 11])
    assert rel_err(res[:, 0], p_w) <= RTOL and rel_err(res[:, 1], coracle.bh(p_w)) <= RTOL
    called = out.updown[0] != 0
    assert called.sum() > 0 and (called & is_de).sum() > 0.5 * called.sum()
    print("config4:", out.stats)


def test_more_than_65535_genes_u32_ranks(reo, oracle, coracle):
    """> 65 535 genes: dense ranks need more than 16 bits (u32 rank buffer, runtime plane count)."""
    r = 70000
    rng = np.random.default_rng(5)
    data = rng.permutation(r * 13).reshape(r, 13).astype(np.int64)   # tie-free: every column has 70 000 distinct levels
    data[::50] //= 50                                                # ... except 2 % of the genes, which collide
    group = ["a"] * 6 + ["b"] * 7
    levels, gid = oracle.group_levels(group)
    info = reo.stage(data, gid, 2)
    assert info["rank_bits"] == 17 and info["gene_tiles"] == -(-r // 64)
    thr = coracle.thresholds_for(gid, 2, 0.01)
    mask = rng.random(r) < 0.05
    cols = np.nonzero(mask)[0]
    got = reo.tables(0, mask, thresholds=thr)
    for i0 in (0, 33333, 69960):
        want, _ = coracle.block_tables(data, gid, 2, thr, cols, seed=7, i0=i0, i1=i0 + 40)
        assert np.array_equal(got[i0:i0 + 40], want)
    assert np.array_equal(got.sum(axis=1), mask.sum() - mask.astype(int))
    out = reo.identify_degs(data, gid, 2, mask, 0.01, 1.0, 0.05, 3, 5)
    assert out.iters[0] >= 1 and out.result.shape == (1, r, 15)


# ---- edge cases ------------------------------------------------------------------------------------
def test_edge_cases_match_oracle(reo, oracle, coracle):
    rng = np.random.default_rng(9)
    gid = np.array([0] * 6 + [1] * 7, dtype=np.int32)
    thr = coracle.thresholds_for(gid, 2, 0.01)
    cases = {
        "all ties (constant matrix)": np.full((60, 13), 5, dtype=np.int64),
        "two values only": rng.integers(0, 2, size=(60, 13)).astype(np.int64),
        "negative counts": rng.integers(-50, 50, size=(60, 13)).astype(np.int64),
        "tie-free": np.stack([rng.permutation(60) for _ in range(13)], axis=1).astype(np.int64),
    }
    for name, data in cases.items():
        for ref in (np.ones(60, bool), np.zeros(60, bool), np.arange(60) == 7, np.arange(60) % 2 == 0):
            want = coracle.identify_degs(data, gid, 2, thr, 1.0, 0.05, ref, 8, 1, seed=7)
            out = reo.identify_degs(data, gid, 2, ref, 0.01, 1.0, 0.05, 8, 1)
            check_full(out, want)
    # one sample in a group, interleaved columns, explicit leading dimension (ld > r through a Fortran-order slice)
    big = np.asfortranarray(rng.integers(0, 30, size=(90, 21)).astype(np.int64))
    view = big[:77, :]                      # column-major view with ld = 90 -> the wrapper copies to ld = r
    gid2 = np.array([1] + [0] * 20, dtype=np.int32)[rng.permutation(21)]
    gid2 = np.where(gid2 == gid2[0], 0, 1).astype(np.int32)   # level ids in order of first appearance
    thr2 = coracle.thresholds_for(gid2, 2, 0.01)
    ref = np.arange(77) % 3 == 0
    want = coracle.identify_degs(view, gid2, 2, thr2, 1.0, 0.05, ref, 8, 1, seed=7)
    check_full(reo.identify_degs(view, gid2, 2, ref, 0.01, 1.0, 0.05, 8, 1), want)


def test_error_behaviour_matches_reference(reo, pkg):
    data = np.random.default_rng(0).integers(0, 9, size=(10, 8)).astype(np.int64)
    gid = np.array([0, 0, 0, 0, 1, 1, 1, 1], dtype=np.int32)
    with pytest.raises(IndexError):                       # r <= 10: BoundsError at src:411
        reo.identify_degs(data, gid, 2, np.ones(10, bool))
    data = np.random.default_rng(0).integers(0, 9, size=(30, 8)).astype(np.int64)
    with pytest.raises(ValueError, match="DimensionMismatch"):   # src:355
        reo.identify_degs(data, gid[:7], 2, np.ones(30, bool))
    with pytest.raises(ValueError, match="DimensionMismatch"):   # src:356: one level only
        reo.identify_degs(data, np.zeros(8, np.int32), 1, np.ones(30, bool))
    with pytest.raises(ValueError, match="DimensionMismatch"):   # a declared level without samples
        reo.identify_degs(data, np.zeros(8, np.int32), 2, np.ones(30, bool))
    out = reo.identify_degs(data, gid, 2, np.ones(30, bool), n_iter=0)   # while i_iter < 0: never entered -> zeros
    assert out.iters == [0] and not out.result.any() and not out.updown.any()
    with pytest.raises(ValueError, match="ArgumentError"):
        pkg.Reo(9999)                                      # bad device index: loud failure, no fallback


def test_c_driver_example(tmp_path):
    """examples/reo_driver.c: the ABI from plain C (no Python in the process)."""
    import subprocess
    from conftest import ROOT
    exe = str(tmp_path / "reo_driver")
    lib_dir = os.path.join(ROOT, "rankcompv3.jl_b200")
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "reo_driver.c"), "-o", exe, "-L" + lib_dir, "-lreo_cuda",
                           "-Wl,-rpath," + lib_dir, "-Wl,--allow-shlib-undefined"])
    out = subprocess.run([exe, "3000", "20", "24"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "Convergence threshold is reached" in out.stdout or "iteration" in out.stdout
    last = [l for l in out.stdout.splitlines() if l.startswith("genes")][0].split()
    called_up, planted_hit = int(last[7]), int(last[11])
    assert called_up > 100 and planted_hit > 100, out.stdout


@pytest.mark.gpu
def test_randomised_shape_sweep():
    """120 random shapes / group layouts / dtypes / masks (scripts/fuzz_parity.py) against the C oracle."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "scripts", "fuzz_parity.py"), "120", "11"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]


@pytest.mark.gpu
def test_pageable_and_pinned_outputs_agree(reo, pkg):
    """REO_OUT_PINNED (what the Python wrapper uses: direct device->host copies into page-locked outputs) and the
    default path (pageable caller buffers, staged inside the library) return the same bytes; n_iter = 0 too."""
    import ctypes as C
    L = pkg._lib
    rng = np.random.default_rng(5)
    r, c = 700, 24
    data = np.asfortranarray(rng.poisson(6.0, size=(r, c)).astype(np.int64))
    gid = np.array([0] * 12 + [1] * 12, dtype=np.int32)
    ref = (rng.random(r) < 0.3).astype(np.uint8)
    for n_iter in (128, 0):
        want = reo.identify_degs(data, gid, 2, ref, 0.01, 1.0, 0.05, n_iter, 2)
        res = np.full((1, 15, r), -7.0)
        ud = np.full((1, r), 9, dtype=np.int8)
        fr = np.full((1, r), 9, dtype=np.uint8)
        it = np.zeros(1, dtype=np.int32)
        rc = reo._lib.reo_identify_degs(reo._h, data.ctypes.data_as(C.c_void_p), L.REO_I64, r, c, r,
                                        gid.ctypes.data_as(C.c_void_p), 2, None, 0.01, 1.0, 0.05,
                                        ref.ctypes.data_as(C.c_void_p), n_iter, 2, 0, res.ctypes.data_as(C.c_void_p),
                                        ud.ctypes.data_as(C.c_void_p), fr.ctypes.data_as(C.c_void_p),
                                        it.ctypes.data_as(C.c_void_p), None)
        assert rc == 0
        assert np.array_equal(res.transpose(0, 2, 1), want.result, equal_nan=True)
        assert np.array_equal(ud, want.updown) and np.array_equal(fr, want.final_ref)
        assert [int(it[0])] == want.iters
    # results of earlier calls stay valid while newer ones are alive (distinct pinned blocks)
    a = reo.identify_degs(data, gid, 2, ref, 0.01, 1.0, 0.05, 128, 2)
    keep = a.result.copy()
    b = reo.identify_degs(data[::-1].copy(order="F"), gid, 2, ref, 0.01, 1.0, 0.05, 128, 2)
    assert np.array_equal(a.result, keep, equal_nan=True) and b.result.ctypes.data != a.result.ctypes.data
