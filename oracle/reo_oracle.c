/*
 * oracle/reo_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
 *
 * Second, independent CPU restatement (plain C + OpenMP) of the REO hot path of
 * pathint/RankCompV3.jl, file src/RankCompV3.jl:
 *     is_greater                 lines 71-77
 *     get_major_reo_lower_count  lines 81-92
 *     McCullagh_test             lines 225-259
 *     identify_degs              lines 339-438
 * It works on the raw expression values (no ranks, no bit-planes) so that it shares no
 * algorithmic shortcut with the CUDA path.  It is also the timed CPU baseline ("port").
 *
 * PARITY PIN STATUS: McCullagh_test is pinned by the reference's known-answer vector
 * (src:206-222); everything else is "parity unpinned" by the reference's own tests (it has
 * none, and Julia is not available here) -- pinned instead against oracle/reo_oracle.py,
 * exact binomial sums, scipy BH and structural invariants (tests/).
 *
 * Tie rule (the reference's is rand(Bool), src:72-73): see oracle/reo_oracle.py docstring.
 *     u(i,s) = mix32(mix32(seed_lo ^ i*0x9E3779B1) + s*0x85EBCA77 + seed_hi) >> 31
 *     coin(i,j,s) = u(i,s) ^ u(j,s) ^ [i<j]
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define REO_INVSQRT2 0.7071067811865476

static inline uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}

uint32_t reo_oracle_u(uint64_t seed, uint32_t i, uint32_t s) {
    uint32_t lo = (uint32_t)seed, hi = (uint32_t)(seed >> 32);
    uint32_t h = mix32(lo ^ (i * 0x9E3779B1u));
    h = mix32(h + s * 0x85EBCA77u + hi);
    return h >> 31;
}

/* launchers such as torchrun export OMP_NUM_THREADS=1; the timing harness asks for all host cores explicitly */
void reo_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int reo_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---------------------------------------------------------------- thresholds, src:81-92 */
static long double log_binom_pmf(int n, int k) {
    return lgammal((long double)n + 1) - lgammal((long double)k + 1) - lgammal((long double)(n - k) + 1)
           - (long double)n * logl(2.0L);
}
/* P[X <= x] for X ~ Binomial(n, 1/2), x <= n/2, summed downwards from x (terms decay) */
static long double binom_cdf_half(int n, int x) {
    long double lp = log_binom_pmf(n, x);
    long double term = 1.0L, sum = 0.0L;
    for (int k = x; k >= 0; --k) {
        sum += term;
        if (term < 1e-25L * sum) break;
        term *= (long double)k / (long double)(n - k + 1); /* pmf(k-1)/pmf(k) */
    }
    return expl(lp) * sum;
}
int reo_oracle_threshold(int n, double pval) {
    /* pval_min = pvalue(Binomial(n), 0) = min(1, 2 * 2^-n) */
    long double pmin = ldexpl(1.0L, 1 - n);
    if (pmin > 1.0L) pmin = 1.0L;
    if (!(pmin < (long double)pval)) return n; /* warn path */
    for (int x = 0; x <= n / 2; ++x) {
        long double cdf = binom_cdf_half(n, x);
        /* for x <= n/2 : P[X >= x] >= P[X <= x], so the inner min is the cdf */
        long double ccdf = 1.0L - (x > 0 ? binom_cdf_half(n, x - 1) : 0.0L);
        long double m = cdf < ccdf ? cdf : ccdf;
        long double p = 2.0L * m;
        if (p > 1.0L) p = 1.0L;
        if (p > (long double)pval) return n - x + 1;
    }
    return -1;
}

/* ---------------------------------------------------------------- McCullagh, src:225-259 */
#define MCC_MAXM 15
/* mat: k x k row-major int64.  out: pval d1 d2 se z1.  LAPACK operation order, no FMA. */
void reo_oracle_mccullagh(const int64_t* mat, int k, double* out) {
    int m = k - 1;
    int64_t N[MCC_MAXM][MCC_MAXM];
    int64_t R[MCC_MAXM];
    double A[MCC_MAXM][MCC_MAXM], nf[MCC_MAXM], Rf[MCC_MAXM], w2[MCC_MAXM], w1[MCC_MAXM];
    for (int i = 1; i < k; ++i)
        for (int j = i; j < k; ++j) {
            int64_t v = 0;
            for (int a = 0; a < i; ++a) for (int b = j; b < k; ++b) v += mat[a * k + b];
            for (int a = j; a < k; ++a) for (int b = 0; b < i; ++b) v += mat[a * k + b];
            N[i - 1][j - 1] = N[j - 1][i - 1] = v;
        }
    for (int i = 1; i < k; ++i) {
        int64_t v = 0;
        for (int a = 0; a < i; ++a) for (int b = i; b < k; ++b) v += mat[a * k + b];
        R[i - 1] = v;
    }
    for (int i = 0; i < m; ++i) { nf[i] = (double)N[i][i]; Rf[i] = (double)R[i]; }
    out[0] = 1.0; out[1] = out[2] = out[3] = out[4] = 0.0;
    int diag = 1;
    for (int i = 0; i < m; ++i) for (int j = i + 1; j < m; ++j) if (N[i][j] != 0) diag = 0;
    const double eps = 2.220446049250313e-16;
    if (diag) {
        /* det(UpperTriangular) = integer product of the diagonal; inv = reciprocal diagonal */
        int zero = 0;
        for (int i = 0; i < m; ++i) if (N[i][i] == 0) zero = 1;
        if (zero) return;
        for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) A[i][j] = (i == j) ? 1.0 / nf[i] : 0.0;
    } else {
        int piv[MCC_MAXM];
        double sign = 1.0;
        for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) A[i][j] = (double)N[i][j];
        for (int j = 0; j < m; ++j) { /* dgetf2 */
            int p = j; double best = fabs(A[j][j]);
            for (int i = j + 1; i < m; ++i) if (fabs(A[i][j]) > best) { best = fabs(A[i][j]); p = i; }
            piv[j] = p;
            if (A[p][j] != 0.0) {
                if (p != j) { for (int c = 0; c < m; ++c) { double t = A[j][c]; A[j][c] = A[p][c]; A[p][c] = t; } sign = -sign; }
                double rinv = 1.0 / A[j][j];
                for (int i = j + 1; i < m; ++i) A[i][j] = A[i][j] * rinv;
            }
            for (int jj = j + 1; jj < m; ++jj)
                for (int i = j + 1; i < m; ++i) A[i][jj] = A[i][jj] - A[i][j] * A[j][jj];
        }
        double det = sign;
        for (int i = 0; i < m; ++i) det = det * A[i][i];
        if (fabs(det) <= eps) return;
        for (int j = 0; j < m; ++j) { /* dtrti2 */
            A[j][j] = 1.0 / A[j][j];
            double ajj = -A[j][j];
            for (int jj = 0; jj < j; ++jj) {
                if (A[jj][j] != 0.0) {
                    double temp = A[jj][j];
                    for (int i = 0; i < jj; ++i) A[i][j] = A[i][j] + temp * A[i][jj];
                    A[jj][j] = A[jj][j] * A[jj][jj];
                }
            }
            for (int i = 0; i < j; ++i) A[i][j] = ajj * A[i][j];
        }
        double work[MCC_MAXM];
        for (int j = m - 2; j >= 0; --j) { /* dgetri */
            for (int i = j + 1; i < m; ++i) { work[i] = A[i][j]; A[i][j] = 0.0; }
            for (int kk = j + 1; kk < m; ++kk) {
                double temp = -work[kk];
                if (temp != 0.0) for (int i = 0; i < m; ++i) A[i][j] = A[i][j] + temp * A[i][kk];
            }
        }
        for (int j = m - 2; j >= 0; --j) {
            int jp = piv[j];
            if (jp != j) for (int i = 0; i < m; ++i) { double t = A[i][j]; A[i][j] = A[i][jp]; A[i][jp] = t; }
        }
    }
    for (int i = 0; i < m; ++i) w2[i] = 0.0;
    for (int kk = 0; kk < m; ++kk) for (int i = 0; i < m; ++i) w2[i] = w2[i] + nf[kk] * A[i][kk];
    double s = 0.0;
    for (int i = 0; i < m; ++i) s = s + nf[i] * w2[i];
    double nu = 1.0 / s;
    for (int i = 0; i < m; ++i) w1[i] = (nf[i] * w2[i]) * nu;
    double d1 = 0.0, sa = 0.0, sb = 0.0;
    for (int i = 0; i < m; ++i) d1 = d1 + w1[i] * log((Rf[i] + 0.5) / ((nf[i] - Rf[i]) + 0.5));
    for (int i = 0; i < m; ++i) { sa = sa + w2[i] * Rf[i]; sb = sb + w2[i] * (nf[i] - Rf[i]); }
    double d2 = log((0.5 + sa) / (0.5 + sb));
    double v1 = 4 * (1 + 0.25 * (d1 * d1)) * nu;
    double v2 = 4 * (1 + 0.25 * (d2 * d2)) * nu;
    double se = sqrt((v1 + v2) * 0.5);
    double z1 = d1 / se;
    double cdf = erfc(-z1 * REO_INVSQRT2) / 2, ccdf = erfc(z1 * REO_INVSQRT2) / 2;
    double p = 2 * (cdf < ccdf ? cdf : ccdf);
    out[0] = p < 1.0 ? p : 1.0; out[1] = d1; out[2] = d2; out[3] = se; out[4] = z1;
}

/* ---------------------------------------------------------------- pair classes, src:366-392 */
typedef struct {
    int64_t r, c;
    int gnum;
    const double* xt;   /* gene-major copy: xt[i*c + s] */
    const uint8_t* ut;  /* coins, same layout */
    const int32_t* gid; /* level of each sample */
    uint64_t seed;
    const int32_t* gidx; /* global gene index of local row i (NULL: identity) */
} reo_ctx;

/* Tie-coin variants (study only; the product and every parity test use mode 0):
 *   0  coin(i,j,s) = u(i,s) ^ u(j,s) ^ [i<j]          -- G bits of entropy per sample, free in the bit-sliced kernel
 *   1  coin(i,j,s) = h(seed, min(i,j), max(i,j), s)    -- an independent fair coin per (unordered pair, sample), which is
 *                     what the reference's rand(Bool) (src:73) draws; mirrored for (j,i) as src:385-386 requires */
/* Element type of the caller's matrix: Julia evaluates abs(x - y) in the matrix' own type before comparing with the
 * Float64 literal 0.1 (src:72), so for Matrix{Float32} the difference is rounded to Float32 first. */
static int g_input_f32 = 0;
void reo_oracle_set_input_f32(int on) { g_input_f32 = on; }
static int g_coin_mode = 0;
void reo_oracle_set_coin_mode(int mode) { g_coin_mode = mode; }
static inline uint32_t pair_coin(uint64_t seed, uint32_t lo, uint32_t hi, uint32_t s) {
    uint32_t h = mix32(((uint32_t)seed) ^ (lo * 0x9E3779B1u));
    h = mix32(h + hi * 0x85EBCA77u + (uint32_t)(seed >> 32));
    h = mix32(h ^ (s * 0xC2B2AE35u));
    return h >> 31;
}

/* nre[g] for g < gnum: count of samples of level g where gene i "is greater" than gene j */
static inline void pair_counts(const reo_ctx* x, int64_t i, int64_t j, int64_t* nre) {
    const double* a = x->xt + i * x->c;
    const double* b = x->xt + j * x->c;
    const uint8_t* ua = x->ut + i * x->c;
    const uint8_t* ub = x->ut + j * x->c;
    const uint8_t o = (uint8_t)(i < j);
    for (int g = 0; g < x->gnum; ++g) nre[g] = 0;
    for (int64_t s = 0; s < x->c; ++s) {
        double d = g_input_f32 ? (double)((float)a[s] - (float)b[s]) : a[s] - b[s];
        int gt;
        if (fabs(d) < 0.1) {                         /* src:72-73, deterministic coin */
            if (g_coin_mode == 0) gt = ua[s] ^ ub[s] ^ o;
            else {
                const uint32_t gi = (uint32_t)(x->gidx ? x->gidx[i] : i), gj = (uint32_t)(x->gidx ? x->gidx[j] : j);
                const uint32_t cbit = pair_coin(x->seed, gi < gj ? gi : gj, gi < gj ? gj : gi, (uint32_t)s);
                gt = (int)(o ? cbit : (cbit ^ 1u));
            }
        }
        else gt = a[s] > b[s];                       /* src:75 */
        nre[x->gid[s]] += gt;
    }
}
static inline int classify(int64_t nre, int64_t not_, int64_t n1, int64_t n2, int64_t t1, int64_t t2) {
    int ic = nre >= t1 ? 3 : ((n1 - nre) >= t1 ? 1 : 2);     /* src:376 */
    int it = not_ >= t2 ? 3 : ((n2 - not_) >= t2 ? 1 : 2);   /* src:377 */
    return 3 * (ic - 1) + it;
}

static void make_ctx_idx(reo_ctx* x, const double* data, int64_t r, int64_t c, int64_t ld, const int32_t* gid,
                         int gnum, uint64_t seed, const int32_t* gidx, double** xt_out, uint8_t** ut_out) {
    double* xt = (double*)malloc(sizeof(double) * (size_t)r * (size_t)c);
    uint8_t* ut = (uint8_t*)malloc((size_t)r * (size_t)c);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < r; ++i)
        for (int64_t s = 0; s < c; ++s) {
            xt[i * c + s] = data[i + ld * s];
            ut[i * c + s] = (uint8_t)reo_oracle_u(seed, (uint32_t)(gidx ? gidx[i] : i), (uint32_t)s);
        }
    x->r = r; x->c = c; x->gnum = gnum; x->xt = xt; x->ut = ut; x->gid = gid; x->seed = seed; x->gidx = gidx;
    *xt_out = xt; *ut_out = ut;
}

static void make_ctx(reo_ctx* x, const double* data, int64_t r, int64_t c, int64_t ld, const int32_t* gid,
                     int gnum, uint64_t seed, double** xt_out, uint8_t** ut_out) {
    double* xt = (double*)malloc(sizeof(double) * (size_t)r * (size_t)c);
    uint8_t* ut = (uint8_t*)malloc((size_t)r * (size_t)c);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < r; ++i)
        for (int64_t s = 0; s < c; ++s) {
            xt[i * c + s] = data[i + ld * s];
            ut[i * c + s] = (uint8_t)reo_oracle_u(seed, (uint32_t)i, (uint32_t)s);
        }
    x->r = r; x->c = c; x->gnum = gnum; x->xt = xt; x->ut = ut; x->gid = gid; x->seed = seed; x->gidx = NULL;
    *xt_out = xt; *ut_out = ut;
}

/*
 * Categories for a block of rows [i0,i1) against every gene, level k_sel (0-based).
 * cat[(i-i0)*r + j] in 1..9, 0 on the diagonal.  thresholds: 2 x gnum column-major
 * (thr[0 + 2k] = threshold[1,k], thr[1 + 2k] = threshold[2,k], src:362).
 */
void reo_oracle_categories(const double* data, int64_t r, int64_t c, int64_t ld, const int32_t* gid, int gnum,
                           const int32_t* thr, uint64_t seed, int k_sel, int64_t i0, int64_t i1, uint8_t* cat) {
    reo_ctx x; double* xt; uint8_t* ut;
    make_ctx(&x, data, r, c, ld, gid, gnum, seed, &xt, &ut);
    int64_t n1 = 0;
    for (int64_t s = 0; s < c; ++s) n1 += (gid[s] == k_sel);
    int64_t n2 = c - n1;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = i0; i < i1; ++i) {
        int64_t nre[64];
        for (int64_t j = 0; j < r; ++j) {
            if (i == j) { cat[(i - i0) * r + j] = 0; continue; }
            pair_counts(&x, i, j, nre);
            int64_t tot = 0;
            for (int g = 0; g < gnum; ++g) tot += nre[g];
            cat[(i - i0) * r + j] =
                (uint8_t)classify(nre[k_sel], tot - nre[k_sel], n1, n2, thr[0 + 2 * k_sel], thr[1 + 2 * k_sel]);
        }
    }
    free(xt); free(ut);
}

/*
 * Timed CPU-baseline kernel: 3x3 tables of rows [i0,i1) against the genes listed in cols[ncols]
 * (the reference mask), level k_sel.  table[(i-i0)*9 + q-1].  Returns compares executed.
 */
int64_t reo_oracle_block_tables(const double* data, int64_t r, int64_t c, int64_t ld, const int32_t* gid, int gnum,
                                const int32_t* thr, uint64_t seed, int k_sel, int64_t i0, int64_t i1,
                                const int32_t* cols, int64_t ncols, int32_t* table) {
    reo_ctx x; double* xt; uint8_t* ut;
    make_ctx(&x, data, r, c, ld, gid, gnum, seed, &xt, &ut);
    int64_t n1 = 0;
    for (int64_t s = 0; s < c; ++s) n1 += (gid[s] == k_sel);
    int64_t n2 = c - n1;
    memset(table, 0, sizeof(int32_t) * 9 * (size_t)(i1 - i0));
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = i0; i < i1; ++i) {
        int64_t nre[64];
        int32_t* t = table + (i - i0) * 9;
        for (int64_t jj = 0; jj < ncols; ++jj) {
            int64_t j = cols[jj];
            if (i == j) continue;
            pair_counts(&x, i, j, nre);
            int64_t tot = 0;
            for (int g = 0; g < gnum; ++g) tot += nre[g];
            t[classify(nre[k_sel], tot - nre[k_sel], n1, n2, thr[0 + 2 * k_sel], thr[1 + 2 * k_sel]) - 1] += 1;
        }
    }
    free(xt); free(ut);
    return (i1 - i0) * ncols * c;
}

/*
 * Same as reo_oracle_block_tables for a SUB-MATRIX of a larger problem: local row i is global gene gidx[i]
 * (coins u(gidx[i], s) and the orientation [gidx[i] < gidx[j]] use the global indices; gidx must be ascending).
 * Lets tests check row blocks of matrices too large to hand to the oracle whole.
 */
int64_t reo_oracle_block_tables_idx(const double* data, int64_t r, int64_t c, int64_t ld, const int32_t* gid, int gnum,
                                    const int32_t* thr, uint64_t seed, int k_sel, const int32_t* gidx,
                                    const int32_t* rows, int64_t nrows, const int32_t* cols, int64_t ncols,
                                    int32_t* table) {
    reo_ctx x; double* xt; uint8_t* ut;
    make_ctx_idx(&x, data, r, c, ld, gid, gnum, seed, gidx, &xt, &ut);
    int64_t n1 = 0;
    for (int64_t s = 0; s < c; ++s) n1 += (gid[s] == k_sel);
    int64_t n2 = c - n1;
    memset(table, 0, sizeof(int32_t) * 9 * (size_t)nrows);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t a = 0; a < nrows; ++a) {
        int64_t nre[64];
        const int64_t i = rows[a];
        int32_t* t = table + a * 9;
        for (int64_t jj = 0; jj < ncols; ++jj) {
            const int64_t j = cols[jj];
            if (i == j) continue;
            pair_counts(&x, i, j, nre);   /* local order == global order because gidx ascends */
            int64_t tot = 0;
            for (int g = 0; g < gnum; ++g) tot += nre[g];
            t[classify(nre[k_sel], tot - nre[k_sel], n1, n2, thr[0 + 2 * k_sel], thr[1 + 2 * k_sel]) - 1] += 1;
        }
    }
    free(xt); free(ut);
    return nrows * ncols * c;
}

/* ---------------------------------------------------------------- empirical null + BH */
static double pairwise_sum(const double* a, int64_t lo, int64_t hi, int sq, double mean) {
#define F(v) (sq ? ((v) - mean) * ((v) - mean) : (v))
    if (lo == hi) return F(a[lo]);
    if (hi - lo < 1024) {
        double v = F(a[lo]) + F(a[lo + 1]);
        for (int64_t i = lo + 2; i <= hi; ++i) v = v + F(a[i]);
        return v;
    }
    int64_t mid = lo + ((hi - lo) >> 1);
    double v1 = pairwise_sum(a, lo, mid, sq, mean);
    double v2 = pairwise_sum(a, mid + 1, hi, sq, mean);
    return v1 + v2;
#undef F
}
typedef struct { double v; int64_t i; } vi_t;
static int cmp_vi(const void* a, const void* b) {
    const vi_t* x = (const vi_t*)a; const vi_t* y = (const vi_t*)b;
    if (x->v < y->v) return -1;
    if (x->v > y->v) return 1;
    return (x->i > y->i) - (x->i < y->i);
}
static double two_sided_p(double z) {
    double cdf = erfc(-z * REO_INVSQRT2) / 2, ccdf = erfc(z * REO_INVSQRT2) / 2;
    double p = 2 * (cdf < ccdf ? cdf : ccdf);
    return p < 1.0 ? p : 1.0;
}
/* src:409-412.  Returns se_emp; pval[r].  r must be > 10. */
double reo_oracle_empirical_null(const double* d1, int64_t r, double* pval) {
    vi_t* v = (vi_t*)malloc(sizeof(vi_t) * (size_t)r);
    double* s = (double*)malloc(sizeof(double) * (size_t)r);
    for (int64_t i = 0; i < r; ++i) { v[i].v = d1[i]; v[i].i = i; }
    qsort(v, (size_t)r, sizeof(vi_t), cmp_vi);
    for (int64_t i = 0; i < r; ++i) s[i] = v[i].v;
    int64_t lo = (int64_t)nearbyint((double)r * 0.05), hi = (int64_t)nearbyint((double)r * 0.95);
    int64_t m = hi - lo + 1;
    double mean = pairwise_sum(s, lo - 1, hi - 1, 0, 0.0) / (double)m;
    double ss = pairwise_sum(s, lo - 1, hi - 1, 1, mean);
    double se = sqrt(ss / (double)(m - 1));
    for (int64_t i = 0; i < r; ++i) {
        double z;
        if (se == 0.0) z = d1[i] == 0.0 ? 0.0 : copysign(INFINITY, d1[i]);
        else z = (d1[i] - 0.0) / se;
        pval[i] = two_sided_p(z);
    }
    free(v); free(s);
    return se;
}
/* MultipleTesting BenjaminiHochberg */
void reo_oracle_bh(const double* p, int64_t n, double* padj) {
    if (n <= 1) { for (int64_t i = 0; i < n; ++i) padj[i] = p[i]; return; }
    vi_t* v = (vi_t*)malloc(sizeof(vi_t) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) { v[i].v = p[i]; v[i].i = i; }
    qsort(v, (size_t)n, sizeof(vi_t), cmp_vi);
    for (int64_t m = 1; m <= n; ++m) v[m - 1].v = v[m - 1].v * ((double)n / (double)m);
    for (int64_t m = n - 1; m >= 1; --m) if (v[m].v < v[m - 1].v) v[m - 1].v = v[m].v;
    for (int64_t m = 0; m < n; ++m) padj[v[m].i] = v[m].v < 1.0 ? v[m].v : 1.0;
    free(v);
}

/* ---------------------------------------------------------------- identify_degs, src:339-438 */
/*
 * result: K x r x 15 (k-major, then gene, then column); updown K x r; final_ref K x r; iters K;
 * deg_log K x n_iter (number of DEGs per evaluation, -1 padded).  K = 1 if gnum == 2 else gnum.
 * Returns 0, or -1 on bad dimensions.
 */
int reo_oracle_identify_degs(const double* data, int64_t r, int64_t c, int64_t ld, const int32_t* gid, int gnum,
                             const int32_t* thr, double pval_deg, double padj_deg, const uint8_t* ref_mask,
                             int n_iter, int n_conv, uint64_t seed, double* result, int8_t* updown,
                             uint8_t* final_ref, int32_t* iters, int32_t* deg_log) {
    if (gnum < 2 || r <= 10) return -1;
    int K = gnum == 2 ? 1 : gnum;
    reo_ctx x; double* xt; uint8_t* ut;
    make_ctx(&x, data, r, c, ld, gid, gnum, seed, &xt, &ut);
    uint8_t* cat = (uint8_t*)malloc((size_t)r * (size_t)r);
    uint8_t* ref = (uint8_t*)malloc((size_t)r);
    uint8_t* inds = (uint8_t*)malloc((size_t)r);
    double* pv = (double*)malloc(sizeof(double) * (size_t)r);
    double* pa = (double*)malloc(sizeof(double) * (size_t)r);
    double* dl = (double*)malloc(sizeof(double) * (size_t)r);
    for (int k = 0; k < K; ++k) {
        int64_t n1 = 0;
        for (int64_t s = 0; s < c; ++s) n1 += (gid[s] == k);
        int64_t n2 = c - n1;
        /* phase A: each unordered pair once, mirrored (src:366-386) */
#pragma omp parallel for schedule(dynamic, 8)
        for (int64_t i = 0; i < r; ++i) {
            int64_t nre[64];
            cat[i * r + i] = 0;
            for (int64_t j = i + 1; j < r; ++j) {
                pair_counts(&x, i, j, nre);
                int64_t tot = 0;
                for (int g = 0; g < gnum; ++g) tot += nre[g];
                int q = classify(nre[k], tot - nre[k], n1, n2, thr[0 + 2 * k], thr[1 + 2 * k]);
                cat[i * r + j] = (uint8_t)q;
                cat[j * r + i] = (uint8_t)(10 - q); /* 3*(3-ic) + (4-it), src:386 */
            }
        }
        /* phase B (src:396-430) */
        double* res = result + (size_t)k * (size_t)r * 15;
        memset(res, 0, sizeof(double) * (size_t)r * 15);
        memcpy(ref, ref_mask, (size_t)r);
        int i_iter = 0, n_eval = 0;
        for (int e = 0; e < n_iter; ++e) deg_log[k * n_iter + e] = -1;
        while (i_iter < n_iter) {
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < r; ++i) {
                int64_t t[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                const uint8_t* ci = cat + i * r;
                for (int64_t j = 0; j < r; ++j) if (ref[j] && ci[j]) t[ci[j] - 1] += 1;
                double o[5];
                reo_oracle_mccullagh(t, 3, o);
                double* ri = res + i * 15;
                for (int q = 0; q < 9; ++q) ri[2 + q] = (double)t[q];
                ri[11] = o[1]; ri[12] = o[2]; ri[13] = o[3]; ri[14] = o[4];
                dl[i] = o[1];
            }
            reo_oracle_empirical_null(dl, r, pv);
            reo_oracle_bh(pv, r, pa);
            int64_t nref = 0, nind = 0;
            for (int64_t i = 0; i < r; ++i) {
                res[i * 15 + 0] = pv[i]; res[i * 15 + 1] = pa[i];
                inds[i] = !((pv[i] <= pval_deg) && (pa[i] <= padj_deg));
                nref += ref[i]; nind += inds[i];
            }
            memcpy(final_ref + (size_t)k * r, ref, (size_t)r);
            deg_log[k * n_iter + n_eval] = (int32_t)(r - nind);
            n_eval += 1;
            if (llabs(nref - nind) < n_conv) break;
            i_iter += 1;
            memcpy(ref, inds, (size_t)r);
        }
        for (int64_t i = 0; i < r; ++i) {
            int sig = (res[i * 15] <= pval_deg) && (res[i * 15 + 1] <= padj_deg);
            double z = res[i * 15 + 14];
            updown[(size_t)k * r + i] = (int8_t)((sig && z > 0) ? 1 : ((sig && z < 0) ? -1 : 0));
        }
        iters[k] = n_eval;
    }
    free(cat); free(ref); free(inds); free(pv); free(pa); free(dl); free(xt); free(ut);
    return 0;
}
