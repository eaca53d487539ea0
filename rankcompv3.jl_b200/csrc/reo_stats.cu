// reo_stats.cu -- K3..K6: per-gene McCullagh test, empirical null, Benjamini-Hochberg, reference-mask
// update.  All FP64, on the device, compiled with -fmad=false so that the operation order below is
// exactly the order of the restatement in oracle/reo_oracle.c (which follows Julia's LAPACK path
// and reproduces the reference's known-answer vector bit for bit).
//
// Reference: src/RankCompV3.jl:225-259 (McCullagh_test), 409-417 (sort, trimmed std, p, BH, mask),
// 426-429 (up/down).  These are O(r) latency-bound kernels (r <= 65535 genes).
#include <math.h>

#include "reo_internal.cuh"

#define REO_INVSQRT2 0.7071067811865476
#define MCC_MAXM 8  // tables up to 9 x 9

// ------------------------------------------------------------------------------------------------
// McCullagh_test, src:225-259, general k x k (k <= 9), LAPACK operation order (dgetf2 / dtrti2 /
// dgetri unblocked, reciprocal pivot scaling, no FMA).  mat row-major.
// ------------------------------------------------------------------------------------------------
__device__ void mccullagh_device(const long long* mat, int k, double* out) {
    const int m = k - 1;
    long long N[MCC_MAXM][MCC_MAXM];
    double A[MCC_MAXM][MCC_MAXM], nf[MCC_MAXM], Rf[MCC_MAXM], w2[MCC_MAXM], w1[MCC_MAXM], work[MCC_MAXM];
    int piv[MCC_MAXM];
    for (int i = 1; i < k; ++i)
        for (int j = i; j < k; ++j) {
            long long v = 0;
            for (int a = 0; a < i; ++a) for (int b = j; b < k; ++b) v += mat[a * k + b];
            for (int a = j; a < k; ++a) for (int b = 0; b < i; ++b) v += mat[a * k + b];
            N[i - 1][j - 1] = v; N[j - 1][i - 1] = v;
        }
    for (int i = 1; i < k; ++i) {
        long long v = 0;
        for (int a = 0; a < i; ++a) for (int b = i; b < k; ++b) v += mat[a * k + b];
        Rf[i - 1] = (double)v;
    }
    for (int i = 0; i < m; ++i) nf[i] = (double)N[i][i];
    out[0] = 1.0; out[1] = 0.0; out[2] = 0.0; out[3] = 0.0; out[4] = 0.0;
    bool diag = true;
    for (int i = 0; i < m; ++i) for (int j = i + 1; j < m; ++j) if (N[i][j] != 0) diag = false;
    const double eps = 2.220446049250313e-16;
    if (diag) {
        for (int i = 0; i < m; ++i) if (N[i][i] == 0) return;
        for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) A[i][j] = (i == j) ? 1.0 / nf[i] : 0.0;
    } else {
        double sign = 1.0;
        for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) A[i][j] = (double)N[i][j];
        for (int j = 0; j < m; ++j) {
            int p = j; double best = fabs(A[j][j]);
            for (int i = j + 1; i < m; ++i) if (fabs(A[i][j]) > best) { best = fabs(A[i][j]); p = i; }
            piv[j] = p;
            if (A[p][j] != 0.0) {
                if (p != j) { for (int c = 0; c < m; ++c) { double t = A[j][c]; A[j][c] = A[p][c]; A[p][c] = t; } sign = -sign; }
                const double rinv = 1.0 / A[j][j];
                for (int i = j + 1; i < m; ++i) A[i][j] = A[i][j] * rinv;
            }
            for (int jj = j + 1; jj < m; ++jj)
                for (int i = j + 1; i < m; ++i) A[i][jj] = A[i][jj] - A[i][j] * A[j][jj];
        }
        double det = sign;
        for (int i = 0; i < m; ++i) det = det * A[i][i];
        if (fabs(det) <= eps) return;
        for (int j = 0; j < m; ++j) {
            A[j][j] = 1.0 / A[j][j];
            const double ajj = -A[j][j];
            for (int jj = 0; jj < j; ++jj) {
                if (A[jj][j] != 0.0) {
                    const double temp = A[jj][j];
                    for (int i = 0; i < jj; ++i) A[i][j] = A[i][j] + temp * A[i][jj];
                    A[jj][j] = A[jj][j] * A[jj][jj];
                }
            }
            for (int i = 0; i < j; ++i) A[i][j] = ajj * A[i][j];
        }
        for (int j = m - 2; j >= 0; --j) {
            for (int i = j + 1; i < m; ++i) { work[i] = A[i][j]; A[i][j] = 0.0; }
            for (int kk = j + 1; kk < m; ++kk) {
                const double temp = -work[kk];
                if (temp != 0.0) for (int i = 0; i < m; ++i) A[i][j] = A[i][j] + temp * A[i][kk];
            }
        }
        for (int j = m - 2; j >= 0; --j) {
            const int jp = piv[j];
            if (jp != j) for (int i = 0; i < m; ++i) { double t = A[i][j]; A[i][j] = A[i][jp]; A[i][jp] = t; }
        }
    }
    for (int i = 0; i < m; ++i) w2[i] = 0.0;
    for (int kk = 0; kk < m; ++kk) for (int i = 0; i < m; ++i) w2[i] = w2[i] + nf[kk] * A[i][kk];
    double s = 0.0;
    for (int i = 0; i < m; ++i) s = s + nf[i] * w2[i];
    const double nu = 1.0 / s;
    for (int i = 0; i < m; ++i) w1[i] = (nf[i] * w2[i]) * nu;
    double d1 = 0.0, sa = 0.0, sb = 0.0;
    for (int i = 0; i < m; ++i) d1 = d1 + w1[i] * log((Rf[i] + 0.5) / ((nf[i] - Rf[i]) + 0.5));
    for (int i = 0; i < m; ++i) { sa = sa + w2[i] * Rf[i]; sb = sb + w2[i] * (nf[i] - Rf[i]); }
    const double d2 = log((0.5 + sa) / (0.5 + sb));
    const double v1 = 4 * (1 + 0.25 * (d1 * d1)) * nu;
    const double v2 = 4 * (1 + 0.25 * (d2 * d2)) * nu;
    const double se = sqrt((v1 + v2) * 0.5);
    const double z1 = d1 / se;
    const double cdf = erfc(-z1 * REO_INVSQRT2) / 2, ccdf = erfc(z1 * REO_INVSQRT2) / 2;
    const double p = 2 * (cdf < ccdf ? cdf : ccdf);
    out[0] = p < 1.0 ? p : 1.0; out[1] = d1; out[2] = d2; out[3] = se; out[4] = z1;
}

// per gene: 3x3 table -> result columns 2..14 (col-major r x 15); src:403-405
__global__ void mccullagh_tables_kernel(const int32_t* __restrict__ table, int64_t r, double* __restrict__ result) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r) return;
    long long mat[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) mat[q] = table[i * 9 + q];
    double o[5];
    mccullagh_device(mat, 3, o);
#pragma unroll
    for (int q = 0; q < 9; ++q) result[i + r * (2 + q)] = (double)mat[q];
    result[i + r * 11] = o[1]; result[i + r * 12] = o[2]; result[i + r * 13] = o[3]; result[i + r * 14] = o[4];
}
cudaError_t reo_launch_mccullagh_tables(const int32_t* table, int64_t r, double* result, cudaStream_t st) {
    mccullagh_tables_kernel<<<(unsigned)((r + 127) / 128), 128, 0, st>>>(table, r, result);
    return cudaGetLastError();
}

__global__ void mccullagh_kxk_kernel(const long long* __restrict__ tables, int64_t n, int k, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long mat[81];
    for (int q = 0; q < k * k; ++q) mat[q] = tables[i * k * k + q];
    double o[5];
    mccullagh_device(mat, k, o);
    for (int q = 0; q < 5; ++q) out[i * 5 + q] = o[q];
}
cudaError_t reo_launch_mccullagh_kxk(const int64_t* tables, int64_t n, int k, double* out, cudaStream_t st) {
    if (k < 2 || k > MCC_MAXM + 1) return cudaErrorInvalidValue;
    mccullagh_kxk_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>((const long long*)tables, n, k, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// stable ascending sort of doubles: chunk bitonic sort in shared memory + cross-chunk ranking.
// Total order = (Julia isless key, original index): -0.0 < 0.0, NaN last.
// ------------------------------------------------------------------------------------------------
#define SORT_CHUNK 1024
#define SORT_THREADS 1024

__device__ __forceinline__ unsigned long long f64_key(double x) {
    if (x != x) return ~0ull;
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ bool kv_less(unsigned long long ka, uint32_t ia, unsigned long long kb, uint32_t ib) {
    return ka < kb || (ka == kb && ia < ib);
}

// One element per thread; bitonic network with warp shuffles for partner distances < 32 and one
// double-buffered shared-memory exchange (single barrier) for the larger ones.
__global__ void __launch_bounds__(SORT_THREADS)
sort_chunks_kernel(const double* __restrict__ x, int64_t n, unsigned long long* __restrict__ keys,
                   uint32_t* __restrict__ idx, int32_t* __restrict__ pos) {
    __shared__ unsigned long long sk[2][SORT_CHUNK];
    __shared__ uint32_t si[2][SORT_CHUNK];
    const int t = threadIdx.x;
    const int64_t g = (int64_t)blockIdx.x * SORT_CHUNK + t;
    unsigned long long k = g < n ? f64_key(x[g]) : ~0ull;
    uint32_t i = g < n ? (uint32_t)g : 0xffffffffu;
    int buf = 0;
    for (int kk = 2; kk <= SORT_CHUNK; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            unsigned long long ok; uint32_t oi;
            if (j >= 32) {
                sk[buf][t] = k; si[buf][t] = i;
                __syncthreads();
                ok = sk[buf][t ^ j]; oi = si[buf][t ^ j];
                buf ^= 1;
            } else {
                ok = __shfl_xor_sync(0xffffffffu, k, j); oi = __shfl_xor_sync(0xffffffffu, i, j);
            }
            const bool asc = (t & kk) == 0, lower = (t & j) == 0;
            const bool other_less = kv_less(ok, oi, k, i);
            // the lower index of an ascending pair keeps the minimum, the upper the maximum (reversed if descending)
            if ((lower == asc) == other_less) { k = ok; i = oi; }
        }
    }
    keys[g] = k; idx[g] = i; pos[g] = t;
}

// grid (A, B): every element of chunk A counts the elements of chunk B below it (binary search in shared memory);
// the last CTA to finish a chunk A scatters that chunk to its final positions.
__global__ void __launch_bounds__(SORT_THREADS)
sort_cross_kernel(const double* __restrict__ x, const unsigned long long* __restrict__ keys,
                  const uint32_t* __restrict__ idx, int32_t* __restrict__ pos, unsigned int* __restrict__ cnt,
                  double* __restrict__ sorted, int32_t* __restrict__ perm) {
    __shared__ unsigned long long sk[SORT_CHUNK];
    __shared__ uint32_t si[SORT_CHUNK];
    __shared__ int last;
    const int A = blockIdx.x, B = blockIdx.y, t = threadIdx.x;
    if (A == B) return;
    sk[t] = keys[(int64_t)B * SORT_CHUNK + t]; si[t] = idx[(int64_t)B * SORT_CHUNK + t];
    const unsigned long long ke = keys[(int64_t)A * SORT_CHUNK + t];
    const uint32_t ie = idx[(int64_t)A * SORT_CHUNK + t];
    __syncthreads();
    if (ie != 0xffffffffu) {
        int a = 0, b = SORT_CHUNK;
        while (a < b) { const int m = (a + b) >> 1; if (kv_less(sk[m], si[m], ke, ie)) a = m + 1; else b = m; }
        if (a) atomicAdd(&pos[(int64_t)A * SORT_CHUNK + t], a);
    }
    __threadfence();
    __syncthreads();
    if (t == 0) {
        last = (atomicAdd(&cnt[A], 1u) == gridDim.y - 2);    // gridDim.y - 1 CTAs work on chunk A
        if (last) cnt[A] = 0u;
    }
    __syncthreads();
    if (!last || ie == 0xffffffffu) return;
    __threadfence();
    const int32_t p = *(reinterpret_cast<volatile int32_t*>(pos) + (int64_t)A * SORT_CHUNK + t);
    sorted[p] = x[ie];
    if (perm) perm[p] = (int32_t)ie;
}

// single-chunk case: the chunk order is the final order
__global__ void sort_scatter_kernel(const double* __restrict__ x, int64_t total, const uint32_t* __restrict__ idx,
                                    const int32_t* __restrict__ pos, double* __restrict__ sorted,
                                    int32_t* __restrict__ perm) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const uint32_t ie = idx[e];
    if (ie == 0xffffffffu) return;
    const int32_t p = pos[e];
    sorted[p] = x[ie];
    if (perm) perm[p] = (int32_t)ie;
}

// make sure the workspace holds n elements (allocation must not happen while a stream is being captured)
cudaError_t reo_sort_reserve(ReoSortWs& ws, int64_t n, cudaStream_t st) {
    const int nchunks = (int)((n + SORT_CHUNK - 1) / SORT_CHUNK);
    const int64_t need = (int64_t)nchunks * SORT_CHUNK;
    if (ws.cap < need) {
        if (ws.keys) cudaFree(ws.keys);
        if (ws.idx) cudaFree(ws.idx);
        if (ws.pos) cudaFree(ws.pos);
        if (ws.cnt) cudaFree(ws.cnt);
        ws.keys = nullptr; ws.idx = nullptr; ws.pos = nullptr; ws.cnt = nullptr; ws.cap = 0;
        cudaError_t e = cudaMalloc(&ws.keys, need * sizeof(unsigned long long));
        if (e != cudaSuccess) return e;
        e = cudaMalloc(&ws.idx, need * sizeof(uint32_t));
        if (e != cudaSuccess) return e;
        e = cudaMalloc(&ws.pos, need * sizeof(int32_t));
        if (e != cudaSuccess) return e;
        e = cudaMalloc(&ws.cnt, nchunks * sizeof(unsigned int));
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(ws.cnt, 0, nchunks * sizeof(unsigned int), st);   // every launch leaves the counters at zero
        if (e != cudaSuccess) return e;
        ws.cap = need;
    }
    return cudaSuccess;
}

cudaError_t reo_launch_sort_f64(const double* x, int64_t n, double* sorted, int32_t* perm, ReoSortWs& ws,
                                cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int nchunks = (int)((n + SORT_CHUNK - 1) / SORT_CHUNK);
    const int64_t need = (int64_t)nchunks * SORT_CHUNK;
    const cudaError_t e = reo_sort_reserve(ws, n, st);
    if (e != cudaSuccess) return e;
    sort_chunks_kernel<<<nchunks, SORT_THREADS, 0, st>>>(x, n, ws.keys, ws.idx, ws.pos);
    if (nchunks > 1) {
        dim3 grid(nchunks, nchunks);
        sort_cross_kernel<<<grid, SORT_THREADS, 0, st>>>(x, ws.keys, ws.idx, ws.pos, ws.cnt, sorted, perm);
    } else {
        sort_scatter_kernel<<<(unsigned)((need + 255) / 256), 256, 0, st>>>(x, need, ws.idx, ws.pos, sorted, perm);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// trimmed std, src:411: std(sorted[lo:hi]) with Julia's pairwise reduction tree (block 1024,
// sequential base case), two-pass, n-1 denominator.  One CTA; leaves are summed in parallel and
// combined in the exact tree order by one thread.
// ------------------------------------------------------------------------------------------------
#define STD_MAX_LEAVES 256

struct StdFrame { int64_t lo, hi; int state; double v1; };

// Post-order evaluation of Julia's pairwise tree without device recursion (explicit frame stack):
// value(lo,hi) = leaf | value(lo,mid) + value(mid+1,hi).  Leaves are consumed left to right.
__device__ double std_combine(int64_t lo0, int64_t hi0, const double* leaf, StdFrame* fr) {
    int sp = 0, next = 0;
    double ret = 0.0;
    fr[sp].lo = lo0; fr[sp].hi = hi0; fr[sp].state = 0; fr[sp].v1 = 0.0; ++sp;
    while (sp > 0) {
        StdFrame& f = fr[sp - 1];
        const int64_t mid = f.lo + ((f.hi - f.lo) >> 1);
        if (f.state == 0) {
            if (f.lo == f.hi || f.hi - f.lo < 1024) { ret = leaf[next++]; --sp; }
            else { f.state = 1; fr[sp].lo = f.lo; fr[sp].hi = mid; fr[sp].state = 0; fr[sp].v1 = 0.0; ++sp; }
        } else if (f.state == 1) {
            f.v1 = ret; f.state = 2;
            fr[sp].lo = mid + 1; fr[sp].hi = f.hi; fr[sp].state = 0; fr[sp].v1 = 0.0; ++sp;
        } else {
            ret = f.v1 + ret; --sp;
        }
    }
    return ret;
}

// bounds of leaf number `want` (left to right) of the tree over [lo0, hi0]; returns the leaf count
__device__ int std_leaf_bounds(int64_t lo0, int64_t hi0, int want, int64_t* lo_out, int64_t* hi_out, StdFrame* fr) {
    int sp = 0, nl = 0;
    fr[0].lo = lo0; fr[0].hi = hi0; sp = 1;
    while (sp > 0) {
        --sp;
        const int64_t lo = fr[sp].lo, hi = fr[sp].hi;
        if (lo == hi || hi - lo < 1024) { if (nl == want) { *lo_out = lo; *hi_out = hi; } ++nl; }
        else {
            const int64_t mid = lo + ((hi - lo) >> 1);
            fr[sp].lo = mid + 1; fr[sp].hi = hi; ++sp;   // right is popped after left
            fr[sp].lo = lo; fr[sp].hi = mid; ++sp;
        }
    }
    return nl;
}

// Three small launches, no inter-CTA waiting (the first version spun on a flag between its two passes):
//   std_leaf_kernel<0>  one warp (CTA of 32) per leaf: sum of the leaf's values                      -> ws[leaf]
//   std_leaf_kernel<1>  every CTA first combines ws[] in tree order into the mean (redundantly: <= 256 adds), then
//                       sums its leaf's squared deviations                                          -> ws[256 + leaf]
//   std_final_kernel    combines those in tree order                                                -> se
// A leaf is summed strictly left to right -- f(a1) + f(a2), then + f(a_i) -- 32 values at a time: every lane loads one,
// the values are handed round with shuffles (off the dependency chain) and every lane carries the same running sum, so
// the serial DADD chain is the only dependency.  Leaves run on different SMs: the FP64 pipe of one SM is not shared.
__device__ __forceinline__ double std_leaf_sum(const double* __restrict__ sorted, int64_t lo, int cnt, int lane, bool sq, double mean) {
    double v = 0.0;
    for (int c0 = 0; c0 < cnt; c0 += 32) {
        double x = 0.0;
        if (c0 + lane < cnt) {
            x = sorted[lo + c0 + lane];
            if (sq) { const double d = x - mean; x = d * d; }
        }
        const int n = min(32, cnt - c0);
        if (c0 == 0) {                                      // the first element starts the sum (no 0.0 + a1)
            v = __shfl_sync(0xffffffffu, x, 0);
            for (int k = 1; k < n; ++k) v = v + __shfl_sync(0xffffffffu, x, k);
        } else if (n == 32) {
#pragma unroll
            for (int k = 0; k < 32; ++k) v = v + __shfl_sync(0xffffffffu, x, k);
        } else {
            for (int k = 0; k < n; ++k) v = v + __shfl_sync(0xffffffffu, x, k);
        }
    }
    return v;
}

template <int PASS>
__global__ void __launch_bounds__(32)
std_leaf_kernel(const double* __restrict__ sorted, int64_t lo0, int64_t hi0, double* __restrict__ ws) {
    __shared__ StdFrame frames[64];
    __shared__ double buf[STD_MAX_LEAVES];
    __shared__ int64_t b_lo, b_hi;
    __shared__ double mean_s;
    const int lane = threadIdx.x;
    if (PASS == 1) for (int l = lane; l < (int)gridDim.x; l += 32) buf[l] = ws[l];
    __syncwarp();
    if (lane == 0) {
        std_leaf_bounds(lo0, hi0, blockIdx.x, &b_lo, &b_hi, frames);
        if (PASS == 1) mean_s = std_combine(lo0, hi0, buf, frames) / (double)(hi0 - lo0 + 1);
    }
    __syncwarp();
    const double v = std_leaf_sum(sorted, b_lo, (int)(b_hi - b_lo + 1), lane, PASS == 1, PASS == 1 ? mean_s : 0.0);
    if (lane == 0) ws[PASS * STD_MAX_LEAVES + blockIdx.x] = v;
}

__global__ void __launch_bounds__(32)
std_final_kernel(int64_t lo0, int64_t hi0, int nleaf, const double* __restrict__ ws, double* __restrict__ se_out) {
    __shared__ StdFrame frames[64];
    __shared__ double buf[STD_MAX_LEAVES];
    for (int l = threadIdx.x; l < nleaf; l += 32) buf[l] = ws[STD_MAX_LEAVES + l];
    __syncwarp();
    if (threadIdx.x == 0) *se_out = sqrt(std_combine(lo0, hi0, buf, frames) / (double)(hi0 - lo0));
}

static int std_count_leaves(int64_t lo, int64_t hi) {
    if (lo == hi || hi - lo < 1024) return 1;
    const int64_t mid = lo + ((hi - lo) >> 1);
    return std_count_leaves(lo, mid) + std_count_leaves(mid + 1, hi);
}

cudaError_t reo_launch_trimmed_std(const double* sorted, int64_t n, double* se_out, double* ws, cudaStream_t st) {
    // round(Int, r*0.05) : round(Int, r*0.95), 1-based inclusive, round-half-even on the FP64 product
    const int64_t lo = (int64_t)nearbyint((double)n * 0.05), hi = (int64_t)nearbyint((double)n * 0.95);
    if (lo < 1 || hi > n || hi < lo) return cudaErrorInvalidValue;
    const int nleaf = std_count_leaves(lo - 1, hi - 1);
    if (nleaf > STD_MAX_LEAVES) return cudaErrorInvalidValue;
    if (!ws) return cudaErrorInvalidValue;            // 2 * STD_MAX_LEAVES doubles
    std_leaf_kernel<0><<<nleaf, 32, 0, st>>>(sorted, lo - 1, hi - 1, ws);
    std_leaf_kernel<1><<<nleaf, 32, 0, st>>>(sorted, lo - 1, hi - 1, ws);
    std_final_kernel<<<1, 32, 0, st>>>(lo - 1, hi - 1, nleaf, ws, se_out);
    return cudaGetLastError();
}

// src:412: pvalue(Normal(0, se), d1; tail = :both)
__device__ __forceinline__ double null_p(double d, double se) {
    double z;
    if (se == 0.0) z = (d == 0.0) ? 0.0 : copysign(INFINITY, d);
    else z = (d - 0.0) / se;
    const double cdf = erfc(-z * REO_INVSQRT2) / 2, ccdf = erfc(z * REO_INVSQRT2) / 2;
    const double p = 2 * (cdf < ccdf ? cdf : ccdf);
    return p < 1.0 ? p : 1.0;
}
__global__ void null_pvals_kernel(const double* __restrict__ d1, int64_t n, const double* __restrict__ se_p,
                                  double* __restrict__ pval) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pval[i] = null_p(d1[i], *se_p);
}
cudaError_t reo_launch_null_pvals(const double* d1, int64_t n, const double* se, double* pval, cudaStream_t st) {
    null_pvals_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d1, n, se, pval);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Ascending order of the empirical-null p-values WITHOUT a second sort: p is a non-increasing function of |d1|,
// so it is the merge, by descending magnitude, of the negative run of the sorted d1 (already descending in
// magnitude) with the reversed non-negative run.  Ties in p get equal adjusted values whatever their order.
// sorted: d1 ascending, perm1: its permutation.  Outputs sorted_p (ascending) and perm2.  With se_p the p-value
// of every gene is computed here from its sorted d1 (bit-identical to d1[g]) and also stored to pval[g] (src:412,
// 415); without it pval is an input.
// ------------------------------------------------------------------------------------------------
__global__ void p_order_kernel(const double* __restrict__ sorted, const int32_t* __restrict__ perm1, int64_t n,
                               double* __restrict__ pval, const double* __restrict__ se_p,
                               double* __restrict__ sorted_p, int32_t* __restrict__ perm2) {
    __shared__ int64_t nv_s, m_s;
    if (threadIdx.x == 0) {
        // NaNs sort last (Julia isless): they keep their places and the merge runs over the first nv elements
        int64_t lo = 0, hi = n;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (sorted[mid] == sorted[mid]) lo = mid + 1; else hi = mid; }
        nv_s = lo;
        // m = number of leading negative elements
        hi = lo; lo = 0;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (sorted[mid] < 0.0) lo = mid + 1; else hi = mid; }
        m_s = lo;
    }
    __syncthreads();
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int64_t nv = nv_s;
    const int64_t m = m_s;         // A = sorted[0..m) (magnitudes descending), B = sorted[nv-1 .. m] (descending)
    const double v = sorted[e];
    const double mag = fabs(v);
    int64_t pos;
    if (v != v) {
        pos = e;                   // NaNs (sorted last) stay last
    } else if (e < m) {
        // A[e]: e elements of A before it + elements of B with magnitude strictly greater
        int64_t a = m, b = nv;     // B region ascending in value: count of sorted[k] > mag for k in [m, nv)
        while (a < b) { const int64_t mid = (a + b) >> 1; if (sorted[mid] > mag) b = mid; else a = mid + 1; }
        pos = e + (nv - a);
    } else {
        // B element at reversed index j = n-1-e: j elements of B before it + elements of A with magnitude >= mag
        int64_t a = 0, b = m;      // A ascending in value (descending magnitude): count of -sorted[k] >= mag
        while (a < b) { const int64_t mid = (a + b) >> 1; if (-sorted[mid] >= mag) a = mid + 1; else b = mid; }
        pos = (nv - 1 - e) + a;
    }
    const int32_t g = perm1[e];
    double p;
    if (se_p) { p = null_p(v, *se_p); pval[g] = p; }
    else p = pval[g];
    sorted_p[pos] = p;
    perm2[pos] = g;
}
cudaError_t reo_launch_p_order(const double* sorted, const int32_t* perm1, int64_t n, double* pval, const double* se,
                               double* sorted_p, int32_t* perm2, cudaStream_t st) {
    p_order_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sorted, perm1, n, pval, se, sorted_p, perm2);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Benjamini-Hochberg, src:413 (MultipleTesting 0.5.1): q_(m) = p_(m) * (n/m), reverse running
// minimum, min(.,1), un-permute.
// ------------------------------------------------------------------------------------------------
#define BH_THREADS 1024
// q_(m) = min(q_(m+1), p_(m) * (n/m)): reverse running minimum (min is exact, so it may be formed in any order).
// bh_scan: every CTA forms the suffix minimum inside its 1024 elements (scratch) and the minimum of the whole
// segment (tot[]).  bh_finish: adds the minimum of the segments to the right, clamps, un-permutes and, when asked,
// writes src:417's mask.  The scattered stores are spread over all CTAs.
__global__ void __launch_bounds__(BH_THREADS)
bh_scan_kernel(const double* __restrict__ sp, int64_t n, double* __restrict__ scratch, double* __restrict__ tot) {
    __shared__ double wmin[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t i = (int64_t)blockIdx.x * BH_THREADS + tid;
    double v = INFINITY;
    if (i < n) v = sp[i] * ((double)n / (double)(i + 1));
    for (int o = 1; o < 32; o <<= 1) {                            // inclusive suffix minimum over lanes >= lane
        const double t = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v = t < v ? t : v;
    }
    if (lane == 0) wmin[wid] = v;
    __syncthreads();
    double right = INFINITY;                                      // warps to the right inside this CTA
    for (int w = wid + 1; w < 32; ++w) { const double t = wmin[w]; right = t < right ? t : right; }
    v = right < v ? right : v;
    if (i < n) scratch[i] = v;
    if (tid == 0) tot[blockIdx.x] = v;
}

__global__ void __launch_bounds__(BH_THREADS)
bh_finish_kernel(const double* __restrict__ sp, const int32_t* __restrict__ perm, int64_t n,
                 const double* __restrict__ scratch, const double* __restrict__ tot, double* __restrict__ padj,
                 uint8_t* __restrict__ mask_new, double pval_deg, double padj_deg) {
    __shared__ double wmin[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double right = INFINITY;                                      // minimum of the segments to the right
    for (int b = blockIdx.x + 1 + tid; b < (int)gridDim.x; b += BH_THREADS) { const double t = tot[b]; right = t < right ? t : right; }
    for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, right, o); right = t < right ? t : right; }
    if (lane == 0) wmin[wid] = right;
    __syncthreads();
    right = wmin[lane];
    for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, right, o); right = t < right ? t : right; }
    const int64_t i = (int64_t)blockIdx.x * BH_THREADS + tid;
    if (i >= n) return;
    double v = scratch[i];
    v = right < v ? right : v;
    v = v < 1.0 ? v : 1.0;
    const int32_t g = perm[i];
    padj[g] = v;
    // inds = .!((pval .<= pval_deg) .&& (padj .<= padj_deg)), src:417 (sp[i] is gene g's p-value)
    if (mask_new) mask_new[g] = !((sp[i] <= pval_deg) && (v <= padj_deg));
}
// ws: n + ceil(n/1024) doubles of scratch
cudaError_t reo_launch_bh(const double* sorted_p, const int32_t* perm, int64_t n, double* padj, double* ws,
                          uint8_t* mask_new, double pval_deg, double padj_deg, cudaStream_t st) {
    if (!ws) return cudaErrorInvalidValue;
    const unsigned nb = (unsigned)((n + BH_THREADS - 1) / BH_THREADS);
    bh_scan_kernel<<<nb, BH_THREADS, 0, st>>>(sorted_p, n, ws, ws + n);
    bh_finish_kernel<<<nb, BH_THREADS, 0, st>>>(sorted_p, perm, n, ws, ws + n, padj, mask_new, pval_deg, padj_deg);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// mask update, src:417-424, and ordered compaction of masks into column lists
// ------------------------------------------------------------------------------------------------
#define MK_THREADS 1024
__device__ __forceinline__ int mk_block_scan(int v, int* red, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) red[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = red[lane];
        int winc = w;
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
        red[lane] = winc - w;
        if (lane == 31) red[32] = winc;
    }
    __syncthreads();
    const int res = red[wid] + inc - v;
    *total = red[32];
    __syncthreads();
    return res;
}

// inds = .!((pval .<= pval_deg) .&& (padj .<= padj_deg)), src:417
__global__ void inds_kernel(const double* __restrict__ pval, const double* __restrict__ padj, int64_t r,
                            double pval_deg, double padj_deg, uint8_t* __restrict__ mask_new) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r) return;
    mask_new[i] = !((pval[i] <= pval_deg) && (padj[i] <= padj_deg));
}

// counts[0] = sum(old), counts[1] = sum(new), counts[2] = |old xor new|; ascending signed list of changes
__global__ void __launch_bounds__(MK_THREADS)
mask_diff_kernel(int64_t r, const uint8_t* __restrict__ mask_old, const uint8_t* __restrict__ mask_new,
                 int32_t* __restrict__ counts, int32_t* __restrict__ changed_gene, int8_t* __restrict__ changed_sign) {
    __shared__ int red[40];
    const int tid = threadIdx.x;
    const int64_t per = (r + MK_THREADS - 1) / MK_THREADS;
    const int64_t lo = (int64_t)tid * per, hi = (lo + per < r) ? lo + per : r;
    int n_old = 0, n_new = 0, n_chg = 0;
    for (int64_t i = lo; i < hi; ++i) {
        const uint8_t nd = mask_new[i] != 0, od = mask_old[i] != 0;
        n_old += od; n_new += nd; n_chg += (od != nd);
    }
    int t_old, t_new, t_chg;
    mk_block_scan(n_old, red, &t_old);
    mk_block_scan(n_new, red, &t_new);
    int pos = mk_block_scan(n_chg, red, &t_chg);
    for (int64_t i = lo; i < hi; ++i) {
        const uint8_t nd = mask_new[i] != 0, od = mask_old[i] != 0;
        if (od != nd) { changed_gene[pos] = (int32_t)i; changed_sign[pos] = nd ? 1 : -1; ++pos; }
    }
    const int padded = (t_chg + 2 * REO_TILE - 1) / (2 * REO_TILE) * (2 * REO_TILE);
    for (int i = t_chg + tid; i < padded; i += MK_THREADS) { changed_gene[i] = -1; changed_sign[i] = 0; }
    if (tid == 0) { counts[0] = t_old; counts[1] = t_new; counts[2] = t_chg; }
}
cudaError_t reo_launch_inds(const double* pval, const double* padj, int64_t r, double pval_deg, double padj_deg,
                            uint8_t* mask_new, cudaStream_t st) {
    inds_kernel<<<(unsigned)((r + 255) / 256), 256, 0, st>>>(pval, padj, r, pval_deg, padj_deg, mask_new);
    return cudaGetLastError();
}
cudaError_t reo_launch_mask_diff(int64_t r, const uint8_t* mask_old, const uint8_t* mask_new, int32_t* counts,
                                 int32_t* changed_gene, int8_t* changed_sign, cudaStream_t st) {
    mask_diff_kernel<<<1, MK_THREADS, 0, st>>>(r, mask_old, mask_new, counts, changed_gene, changed_sign);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(MK_THREADS)
mask_to_list_kernel(const uint8_t* __restrict__ mask, int64_t r, int32_t* __restrict__ list, int32_t* __restrict__ count) {
    __shared__ int red[40];
    const int tid = threadIdx.x;
    const int64_t per = (r + MK_THREADS - 1) / MK_THREADS;
    const int64_t lo = (int64_t)tid * per, hi = (lo + per < r) ? lo + per : r;
    int n = 0;
    for (int64_t i = lo; i < hi; ++i) n += mask[i] != 0;
    int total;
    int pos = mk_block_scan(n, red, &total);
    for (int64_t i = lo; i < hi; ++i) if (mask[i]) list[pos++] = (int32_t)i;
    const int padded = (total + 2 * REO_TILE - 1) / (2 * REO_TILE) * (2 * REO_TILE);
    for (int i = total + tid; i < padded; i += MK_THREADS) list[i] = -1;
    if (tid == 0) *count = total;
}
cudaError_t reo_launch_mask_to_list(const uint8_t* mask, int64_t r, int32_t* list, int32_t* count, cudaStream_t st) {
    mask_to_list_kernel<<<1, MK_THREADS, 0, st>>>(mask, r, list, count);
    return cudaGetLastError();
}

// Lists for the symmetric pair kernel: C = { i : in_c(i) } ascending, pads (-1) up to a whole number of T-tile
// blocks, then N = the other genes ascending, pads up to `cap`.  in_c(i) = m_new ? (m_old[i] != m_new[i]) : m_old[i];
// sign = +1 for a gene that enters the reference set (or a plain member), -1 for one that leaves, 0 for N and pads.
// counts_out[0] = |C|, counts_out[1] = |N|.
__global__ void __launch_bounds__(MK_THREADS)
sym_lists_kernel(int64_t r, const uint8_t* __restrict__ m_old, const uint8_t* __restrict__ m_new, int T,
                 int32_t* __restrict__ gene, int8_t* __restrict__ sign, int32_t* __restrict__ counts_out, int64_t cap) {
    __shared__ int red[40];
    const int tid = threadIdx.x;
    const int64_t per = (r + MK_THREADS - 1) / MK_THREADS;
    const int64_t lo = (int64_t)tid * per, hi = (lo + per < r) ? lo + per : r;
    int n_c = 0;
    for (int64_t i = lo; i < hi; ++i) n_c += m_new ? ((m_old[i] != 0) != (m_new[i] != 0)) : (m_old[i] != 0);
    const int n_n = (hi > lo ? (int)(hi - lo) : 0) - n_c;
    int t_c, t_n;
    int pc = mk_block_scan(n_c, red, &t_c);
    int pn = mk_block_scan(n_n, red, &t_n);
    const int64_t blk = (int64_t)T * REO_TILE;
    const int64_t off_n = (t_c + blk - 1) / blk * blk;
    for (int64_t i = lo; i < hi; ++i) {
        const bool od = m_old[i] != 0;
        const bool in_c = m_new ? (od != (m_new[i] != 0)) : od;
        if (in_c) { gene[pc] = (int32_t)i; sign[pc] = m_new ? (od ? -1 : 1) : 1; ++pc; }
        else { gene[off_n + pn] = (int32_t)i; sign[off_n + pn] = 0; ++pn; }
    }
    for (int64_t i = t_c + tid; i < off_n; i += MK_THREADS) { gene[i] = -1; sign[i] = 0; }
    for (int64_t i = off_n + t_n + tid; i < cap; i += MK_THREADS) { gene[i] = -1; sign[i] = 0; }
    if (tid == 0) { counts_out[0] = t_c; counts_out[1] = t_n; }
}
cudaError_t reo_launch_sym_lists(int64_t r, const uint8_t* m_old, const uint8_t* m_new, int T, int32_t* gene, int8_t* sign,
                                 int32_t* counts_out, int64_t cap, cudaStream_t st) {
    sym_lists_kernel<<<1, MK_THREADS, 0, st>>>(r, m_old, m_new, T, gene, sign, counts_out, cap);
    return cudaGetLastError();
}

// out[i] = sum over q of all[q][i]   (tables of `world` ranks gathered by a user collective)
__global__ void sum_slices_kernel(const int32_t* __restrict__ all, int world, int64_t n, int32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t v = 0;
    for (int q = 0; q < world; ++q) v += all[(int64_t)q * n + i];
    out[i] = v;
}
cudaError_t reo_launch_sum_slices(const int32_t* all, int world, int64_t n, int32_t* out, cudaStream_t st) {
    sum_slices_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(all, world, n, out);
    return cudaGetLastError();
}

// src:426-429
__global__ void updown_kernel(const double* __restrict__ result, int64_t r, double pval_deg, double padj_deg,
                              int8_t* __restrict__ updown) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r) return;
    const bool sig = (result[i] <= pval_deg) && (result[i + r] <= padj_deg);
    const double z = result[i + r * 14];
    updown[i] = (int8_t)((sig && z > 0) ? 1 : ((sig && z < 0) ? -1 : 0));
}
cudaError_t reo_launch_updown(const double* result, int64_t r, double pval_deg, double padj_deg, int8_t* updown,
                              cudaStream_t st) {
    updown_kernel<<<(unsigned)((r + 255) / 256), 256, 0, st>>>(result, r, pval_deg, padj_deg, updown);
    return cudaGetLastError();
}
