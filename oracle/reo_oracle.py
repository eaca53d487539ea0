"""
oracle/reo_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

A CPU (numpy) restatement of the REO hot path of pathint/RankCompV3.jl:

    is_greater                  src/RankCompV3.jl:71-77
    get_major_reo_lower_count   src/RankCompV3.jl:81-92
    McCullagh_test              src/RankCompV3.jl:225-259
    identify_degs               src/RankCompV3.jl:339-438

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module, and only as the checker.

PARITY PIN STATUS
-----------------
* McCullagh_test is pinned by the reference's single known-answer vector (src:206-222,
  test/McCullagh_test.jl:38-39): see tests/test_oracle_kat.py.
* Everything else (thresholds, pair counts, classes, tables, empirical-null p, BH, the
  iteration loop) is **parity unpinned** by the reference's own tests: the reference has no
  test-suite and Julia is not available in this image, so the real implementation cannot be
  run.  Build-side pins: a second independent restatement in plain C (oracle/reo_oracle.c),
  exact big-integer binomial sums for the thresholds, scipy's BH, and structural invariants.

RANDOM TIES
-----------
The reference flips `rand(Bool)` when |x-y| < 0.1 (src:72-73), once per unordered pair per
sample, and mirrors the result to the other gene (src:385-386).  That is not reproducible, so
oracle and device share a deterministic counter-based rule that is equal in distribution for
every gene's table:

    u(i,s)      = top bit of a 2-round 32-bit mixer of (seed, gene index i, sample index s)
    coin(i,j,s) = u(i,s) XOR u(j,s) XOR [i < j]          (0-based indices into `data`)
    is_greater(i,j,s) on a tie  :=  coin(i,j,s) == 1

coin(j,i,s) = NOT coin(i,j,s), so the reference's mirror property holds exactly, and for a fixed
gene i the coins over all partners j and samples s are independent fair coins.
On tie-free inputs no coin is consulted and the result equals the reference's.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

_M32 = 0xFFFFFFFF
_K_GENE = 0x9E3779B1
_K_SAMP = 0x85EBCA77

INVSQRT2 = 0.7071067811865476  # IrrationalConstants.invsqrt2 rounded to Float64


# --------------------------------------------------------------------------------------
# tie coins
# --------------------------------------------------------------------------------------
def mix32(x):
    """lowbias32 integer mixer on uint32 arrays (wraps mod 2^32)."""
    x = np.asarray(x, dtype=np.uint64) & _M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & _M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & _M32
    x ^= x >> np.uint64(16)
    return x


def coin_bits(seed: int, r: int, c: int) -> np.ndarray:
    """u(i,s) for i<r, s<c as a uint8 r x c array (see module docstring)."""
    lo = np.uint64(seed & _M32)
    hi = np.uint64((seed >> 32) & _M32)
    i = np.arange(r, dtype=np.uint64)[:, None]
    s = np.arange(c, dtype=np.uint64)[None, :]
    h = mix32(lo ^ ((i * np.uint64(_K_GENE)) & _M32))
    h = mix32((h + ((s * np.uint64(_K_SAMP)) & _M32) + hi) & _M32)
    return (h >> np.uint64(31)).astype(np.uint8)


# --------------------------------------------------------------------------------------
# a-2: get_major_reo_lower_count  (src:81-92), exact rational arithmetic
# --------------------------------------------------------------------------------------
def major_reo_lower_count(n: int, pval: float = 0.01) -> int:
    """
    src:81-92.  pvalue(Binomial(n, 1/2), x) (HypothesisTests 0.10.11, discrete two-sided) is
    min(1, 2*min(ccdf(x-1), cdf(x))).  Evaluated here with exact integers, so the result is the
    mathematically exact threshold (the reference's goes through a Float64 beta_inc).
    """
    n = int(n)
    alpha = Fraction(pval)
    two_n = 1 << n
    # pval_min = pvalue(Binomial(n), 0) = min(1, 2*min(P[X>=0], P[X<=0])) = min(1, 2*2^-n)
    pmin = min(Fraction(1), Fraction(2, two_n))
    if not (pmin < alpha):
        return n  # warn path, src:88-90
    cum = 0
    comb = 1  # C(n,0)
    total = two_n
    for x in range(0, n // 2 + 1):
        cum += comb  # sum_{k<=x} C(n,k)
        cdf = Fraction(cum, total)
        ccdf_xm1 = Fraction(total - (cum - comb), total)  # P[X >= x]
        p = min(Fraction(1), 2 * min(ccdf_xm1, cdf))
        if p > alpha:
            return n - x + 1  # -(x+1) + 2 + n
        comb = comb * (n - x) // (x + 1)
    # findfirst returned nothing: reference would throw (cannot happen for alpha < 1)
    raise ValueError("no threshold found")


# --------------------------------------------------------------------------------------
# a-5: McCullagh_test  (src:225-259), general k x k
# --------------------------------------------------------------------------------------
def mccullagh_NR(mat):
    """N (symmetric (k-1)x(k-1), Int), n = diag(N), R -- src:229-240."""
    mat = np.asarray(mat, dtype=np.int64)
    k = mat.shape[0]
    if mat.shape[0] != mat.shape[1]:
        raise ValueError("input matrix 'mat' should be a square matrix.")
    N = np.zeros((k - 1, k - 1), dtype=np.int64)
    for i in range(1, k):
        for j in range(i, k):
            v = mat[0:i, j:].sum() + mat[j:, 0:i].sum()
            N[i - 1, j - 1] = N[j - 1, i - 1] = v
    R = np.array([mat[0:i, i:].sum() for i in range(1, k)], dtype=np.int64)
    return N, np.diag(N).copy(), R


def _lu_getrf(A):
    """Unblocked partial-pivot LU in Float64, reciprocal scaling (LAPACK dgetf2 order)."""
    A = A.astype(np.float64).copy()
    m = A.shape[0]
    piv = list(range(m))
    sign = 1.0
    for j in range(m):
        p = j + int(np.argmax(np.abs(A[j:, j])))  # first max, like idamax
        piv[j] = p
        if A[p, j] != 0.0:
            if p != j:
                A[[j, p], :] = A[[p, j], :]
                sign = -sign
            rinv = 1.0 / A[j, j]
            for i in range(j + 1, m):
                A[i, j] = A[i, j] * rinv
        for jj in range(j + 1, m):  # rank-1 update, column by column
            for i in range(j + 1, m):
                A[i, jj] = A[i, jj] - A[i, j] * A[j, jj]
    return A, piv, sign


def _getri(LU, piv):
    """inv from LU, LAPACK dtrti2 + dgetri (unblocked) operation order, no FMA."""
    A = LU.copy()
    m = A.shape[0]
    # dtrti2, upper, non-unit
    for j in range(m):
        A[j, j] = 1.0 / A[j, j]
        ajj = -A[j, j]
        # x = A[0:j, j];  x := T * x with T = already inverted leading block (dtrmv U,N,N)
        for jj in range(j):
            if A[jj, j] != 0.0:
                temp = A[jj, j]
                for i in range(jj):
                    A[i, j] = A[i, j] + temp * A[i, jj]
                A[jj, j] = A[jj, j] * A[jj, jj]
        for i in range(j):
            A[i, j] = ajj * A[i, j]
    # dgetri unblocked: solve inv(A)*L = inv(U)
    work = np.zeros(m)
    for j in range(m - 2, -1, -1):
        for i in range(j + 1, m):
            work[i] = A[i, j]
            A[i, j] = 0.0
        for kk in range(j + 1, m):  # dgemv('N', alpha=-1): y += (-work[k]) * A[:,k]
            temp = -work[kk]
            if temp != 0.0:
                for i in range(m):
                    A[i, j] = A[i, j] + temp * A[i, kk]
    for j in range(m - 2, -1, -1):
        jp = piv[j]
        if jp != j:
            A[:, [j, jp]] = A[:, [jp, j]]
    return A


def mccullagh_test(mat, variant: str = "lapack"):
    """
    src:225-259.  Returns (pval, d1, d2, se, z1).

    variant="lapack": follows the operation order of Julia 1.7's det/inv on a dense Float64 copy
    of N (diagonal N -> exact integer det + reciprocal diagonal; otherwise partial-pivot LU,
    getri), no FMA.  This is the definition the device kernel is held to.
    variant="closed": the algebraic closed form for k=3, for cross-checking.
    """
    N, n, R = mccullagh_NR(mat)
    m = N.shape[0]
    eps = np.finfo(np.float64).eps
    nf = n.astype(np.float64)
    Rf = R.astype(np.float64)
    if variant == "closed":
        if m != 2:
            raise ValueError("closed form is for 3x3 tables")
        a, b, c = int(N[0, 0]), int(N[0, 1]), int(N[1, 1])
        det = a * c - b * b
        if abs(det) <= eps:
            return (1.0, 0.0, 0.0, 0.0, 0.0)
        w2 = np.array([c * (a - b) / det, a * (c - b) / det], dtype=np.float64)
        nu = det / (float(a) * float(c) * float(a + c - 2 * b))
        w1 = np.array([(a - b) / (a + c - 2 * b), (c - b) / (a + c - 2 * b)], dtype=np.float64)
    else:
        is_tri = bool(np.all(np.triu(N, 1) == 0))  # symmetric -> diagonal
        if is_tri:
            det = 1
            for i in range(m):
                det *= int(N[i, i])
            if abs(det) <= eps:
                return (1.0, 0.0, 0.0, 0.0, 0.0)
            Ni = np.diag(1.0 / nf)
        else:
            LU, piv, sign = _lu_getrf(N)
            det = sign
            for i in range(m):
                det = det * LU[i, i]
            if abs(det) <= eps:
                return (1.0, 0.0, 0.0, 0.0, 0.0)
            Ni = _getri(LU, piv)
        # w2 = Ni*n  (column sweep like gemv 'N')
        w2 = np.zeros(m)
        for kk in range(m):
            for i in range(m):
                w2[i] = w2[i] + nf[kk] * Ni[i, kk]
        s = 0.0
        for i in range(m):
            s = s + nf[i] * w2[i]
        nu = 1.0 / s
        w1 = np.array([(nf[i] * w2[i]) * nu for i in range(m)])
    lg = np.array([math.log((Rf[i] + 0.5) / ((nf[i] - Rf[i]) + 0.5)) for i in range(m)])
    d1 = 0.0
    for i in range(m):
        d1 = d1 + w1[i] * lg[i]
    sa = 0.0
    sb = 0.0
    for i in range(m):
        sa = sa + w2[i] * Rf[i]
        sb = sb + w2[i] * (nf[i] - Rf[i])
    d2 = math.log((0.5 + sa) / (0.5 + sb))
    v1 = 4 * (1 + 0.25 * d1 ** 2) * nu
    v2 = 4 * (1 + 0.25 * d2 ** 2) * nu
    se = math.sqrt((v1 + v2) * 0.5)
    z1 = d1 / se
    pval = normal_two_sided_p(z1)
    return (pval, d1, d2, se, z1)


def normal_two_sided_p(z: float) -> float:
    """pvalue(Normal(0,1), z; tail=:both) = min(1, 2*min(cdf, ccdf)), cdf = erfc(-z*invsqrt2)/2."""
    if math.isnan(z):
        return float("nan")
    cdf = math.erfc(-z * INVSQRT2) / 2
    ccdf = math.erfc(z * INVSQRT2) / 2
    return min(1.0, 2 * min(cdf, ccdf))


# --------------------------------------------------------------------------------------
# a-6: empirical null + BH (src:409-416)
# --------------------------------------------------------------------------------------
def pairwise_sum(a, f=lambda x: x, blk: int = 1024) -> float:
    """Julia Base.mapreduce_impl(f, +, a, 1, n, 1024) with a sequential base case."""
    a = np.asarray(a, dtype=np.float64)

    def rec(lo, hi):  # inclusive, 0-based
        if lo == hi:
            return float(f(a[lo]))
        if hi - lo < blk:
            v = float(f(a[lo])) + float(f(a[lo + 1]))
            for i in range(lo + 2, hi + 1):
                v = v + float(f(a[i]))
            return v
        mid = lo + ((hi - lo) >> 1)
        return rec(lo, mid) + rec(mid + 1, hi)

    return rec(0, len(a) - 1)


def julia_std(v) -> float:
    """Statistics.std (corrected, two-pass, pairwise)."""
    v = np.asarray(v, dtype=np.float64)
    m = len(v)
    mean = pairwise_sum(v) / m
    ss = pairwise_sum(v, lambda x: (x - mean) * (x - mean))
    return math.sqrt(ss / (m - 1))


def trim_bounds(r: int):
    """round(Int, r*0.05), round(Int, r*0.95): 1-based inclusive (src:411)."""
    lo = int(np.rint(r * 0.05))
    hi = int(np.rint(r * 0.95))
    return lo, hi


def empirical_null_p(delta1):
    """src:409-412 -> (se_emp, pval[r])."""
    d = np.asarray(delta1, dtype=np.float64)
    r = len(d)
    srt = np.sort(d, kind="stable")
    lo, hi = trim_bounds(r)
    if lo < 1:
        raise IndexError("BoundsError: r <= 10 (src:411)")
    se = julia_std(srt[lo - 1:hi])
    p = np.empty(r)
    for i in range(r):
        if se == 0.0:
            z = 0.0 if d[i] == 0.0 else math.copysign(math.inf, d[i])
        else:
            z = (d[i] - 0.0) / se
        p[i] = normal_two_sided_p(z)
    return se, p


def bh_adjust(p):
    """MultipleTesting.adjust(p, BenjaminiHochberg()) -- SURVEY Appendix A item 6."""
    p = np.asarray(p, dtype=np.float64)
    n = len(p)
    if n <= 1:
        return p.copy()
    o = np.argsort(p, kind="stable")
    q = p[o].copy()
    for m in range(1, n + 1):
        q[m - 1] = q[m - 1] * (n / m)
    for m in range(n - 1, 0, -1):
        q[m - 1] = min(q[m], q[m - 1])
    out = np.empty(n)
    out[o] = q
    return np.minimum(out, 1.0)


# --------------------------------------------------------------------------------------
# a-1/a-3: pair counts and classes (src:366-392)
# --------------------------------------------------------------------------------------
def group_levels(group):
    """unique(group) in order of first appearance -> (levels, group_id[c])."""
    levels = []
    gid = np.empty(len(group), dtype=np.int32)
    for s, g in enumerate(group):
        if g not in levels:
            levels.append(g)
        gid[s] = levels.index(g)
    return levels, gid


def greater_counts(data, gid, gnum, rows, cols, seed=0, u=None):
    """
    nre[k, a, b] = #{s in level k : is_greater(data[rows[a], s], data[cols[b], s])}
    with the module's tie rule.  (src:372-373 for one pair.)
    """
    data = np.asarray(data)
    r, c = data.shape
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    if u is None:
        u = coin_bits(seed, r, c)
    # Matrix{Float32} input: Julia subtracts in Float32 and only then compares with the Float64 literal 0.1 (src:72)
    f32 = data.dtype == np.float32
    X = data[rows] if f32 else data[rows].astype(np.float64)
    Y = data[cols] if f32 else data[cols].astype(np.float64)
    uX = u[rows]
    uY = u[cols]
    lt_idx = (rows[:, None] < cols[None, :])
    out = np.zeros((gnum, len(rows), len(cols)), dtype=np.int64)
    for k in range(gnum):
        sel = np.nonzero(gid == k)[0]
        for s in sel:
            x = X[:, s][:, None]
            y = Y[:, s][None, :]
            tie = np.abs(x - y).astype(np.float64) < 0.1   # (x - y) is rounded to Float32 first when the input is
            coin = (uX[:, s][:, None] ^ uY[:, s][None, :]) ^ lt_idx
            out[k] += np.where(tie, coin.astype(bool), x > y)
    return out


def classify(nre_k, not_k, n1, n2, thr1, thr2):
    """src:376-377: (ic, it) -> q = 3*(ic-1)+it in 1..9."""
    ic = np.where(nre_k >= thr1, 3, np.where((n1 - nre_k) >= thr1, 1, 2))
    it = np.where(not_k >= thr2, 3, np.where((n2 - not_k) >= thr2, 1, 2))
    return 3 * (ic - 1) + it


def pair_categories(data, gid, gnum, thresholds, seed=0, rows=None, cols=None, block=256):
    """
    Category q (1..9) of gene rows[a] versus gene cols[b], for every level k
    (k = 0 only when gnum == 2, src:387-389).  0 on the diagonal (i == j).
    Returns uint8 [K, len(rows), len(cols)].
    """
    data = np.asarray(data)
    r, c = data.shape
    rows = np.arange(r) if rows is None else np.asarray(rows)
    cols = np.arange(r) if cols is None else np.asarray(cols)
    K = 1 if gnum == 2 else gnum
    gs1 = np.array([(gid == k).sum() for k in range(gnum)], dtype=np.int64)
    gs2 = c - gs1
    u = coin_bits(seed, r, c)
    out = np.zeros((K, len(rows), len(cols)), dtype=np.uint8)
    for a0 in range(0, len(rows), block):
        rb = rows[a0:a0 + block]
        nre = greater_counts(data, gid, gnum, rb, cols, seed, u)
        tot = nre.sum(axis=0)
        for k in range(K):
            q = classify(nre[k], tot - nre[k], gs1[k], gs2[k], thresholds[0][k], thresholds[1][k])
            q = np.where(rb[:, None] == cols[None, :], 0, q)
            out[k, a0:a0 + block] = q
    return out


def tables_from_categories(cat_k, ref_mask_cols):
    """src:403: 9-bin count of categories over reference columns -> int64 [rows, 9]."""
    sel = cat_k[:, np.asarray(ref_mask_cols, dtype=bool)]
    out = np.zeros((cat_k.shape[0], 9), dtype=np.int64)
    for q in range(1, 10):
        out[:, q - 1] = (sel == q).sum(axis=1)
    return out


# --------------------------------------------------------------------------------------
# identify_degs (src:339-438)
# --------------------------------------------------------------------------------------
def identify_degs(data, group, pval_reo=0.01, pval_deg=1.0, padj_deg=0.05, ref_gene=None,
                  n_iter=128, n_conv=5, seed=0, variant="lapack", thresholds=None):
    """
    Returns a dict:
      result   float64 [K, r, 15]  (pval padj n11..n33 d1 d2 se z1, src:398/405/665)
      updown   int8 [K, r]         (+1 up, -1 down, 0 no change, src:426-429)
      final_ref uint8 [K, r]       reference mask used by the last evaluation
      iters    list[K] of evaluations performed
      log      per-k list of (n_deg, n_nondeg) per evaluation (src:418)
      thresholds int64 [2, gnum]
    """
    data = np.asarray(data)
    r, c = data.shape
    levels, gid = group_levels(list(group))
    gnum = len(levels)
    if c != len(group):
        raise ValueError("DimensionMismatch: 'data' and 'group' do not have compatiable sizes")
    if gnum < 2:
        raise ValueError("DimensionMismatch: Only 1 level in 'group', at least 2 levels!")
    gs1 = np.array([(gid == k).sum() for k in range(gnum)], dtype=np.int64)
    gs2 = c - gs1
    if thresholds is None:
        thresholds = np.array([[major_reo_lower_count(int(v), pval_reo) for v in gs1],
                               [major_reo_lower_count(int(v), pval_reo) for v in gs2]], dtype=np.int64)
    ref0 = np.asarray(ref_gene, dtype=bool)
    cat = pair_categories(data, gid, gnum, thresholds, seed)
    K = cat.shape[0]
    result = np.zeros((K, r, 15))
    updown = np.zeros((K, r), dtype=np.int8)
    final_ref = np.zeros((K, r), dtype=np.uint8)
    iters, logs = [], []
    for k in range(K):
        ref = ref0.copy()
        i_iter = 0
        res = np.zeros((r, 15))
        log = []
        n_eval = 0
        while i_iter < n_iter:
            tab = tables_from_categories(cat[k], ref)
            for i in range(r):
                t = tab[i]
                _, d1, d2, se, z1 = mccullagh_test(t.reshape(3, 3), variant)
                res[i, 2:11] = t
                res[i, 11:15] = (d1, d2, se, z1)
            _, pval = empirical_null_p(res[:, 11])
            padj = bh_adjust(pval)
            res[:, 0] = pval
            res[:, 1] = padj
            inds = ~((pval <= pval_deg) & (padj <= padj_deg))
            n_eval += 1
            final_ref[k] = ref
            log.append((int(r - inds.sum()), int(inds.sum())))
            if abs(int(ref.sum()) - int(inds.sum())) < n_conv:
                break
            i_iter += 1
            ref = inds
        sig = (res[:, 0] <= pval_deg) & (res[:, 1] <= padj_deg)
        updown[k][(res[:, 14] > 0) & sig] = 1
        updown[k][(res[:, 14] < 0) & sig] = -1
        result[k] = res
        iters.append(n_eval)
        logs.append(log)
    return dict(result=result, updown=updown, final_ref=final_ref, iters=iters, log=logs,
                thresholds=thresholds, levels=levels)
