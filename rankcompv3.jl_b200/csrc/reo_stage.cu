// reo_stage.cu -- K1: staging of the expression matrix for the pair kernel.
//
// Reference semantics being prepared (src/RankCompV3.jl): the pair loop only ever evaluates
// is_greater(data[i,s], data[j,s]) (src:71-77, 372) for two genes in the SAME sample, and for
// integer-valued data the tie band |x-y| < 0.1 is exactly x == y.  So each sample column can be
// replaced by its dense rank over genes (ties share a rank) without changing any REO.  Ranks are
// then bit-sliced: 32 samples of one group level per 32-bit word, one word per rank bit, so the
// pair kernel compares 32 samples per LOP3.  Plane 0 carries the tie-coin bits u(i,s).
//
// Kernels: rank_columns_kernel (bitmap dense rank, one CTA per sample), rank_fallback_kernel
// (sort-based dense rank for columns whose value range exceeds the bitmap), bitplanes_kernel
// (32 x B bit transpose + coin plane), gather_panel_kernel (compacts reference-gene columns).
// All are HBM/L2-bound streaming passes; algorithmic bytes are stated in DESIGN.md.
#include <limits.h>

#include <algorithm>

#include "reo_internal.cuh"

#define RK_THREADS 1024                 // both tiers and the fallback
#define BM_WORDS 40960                  // wide tier: bitmap words in shared memory (1,310,720 distinct values), 1 CTA per SM
#define BM_WORDS_S 2048                 // small tier: value range < 65536 (counts of most samples), 2 CTAs per SM
#define RK_STASH_MAX_R 51200            // small tier keeps (v - min) as u16 in shared memory up to this many genes
#define BM_GROUP 8                      // words per prefix entry

template <typename T>
__device__ __forceinline__ bool to_ll(T v, long long& out);
template <>
__device__ __forceinline__ bool to_ll<long long>(long long v, long long& out) { out = v; return true; }
template <>
__device__ __forceinline__ bool to_ll<int>(int v, long long& out) { out = v; return true; }
template <>
__device__ __forceinline__ bool to_ll<unsigned short>(unsigned short v, long long& out) { out = v; return true; }
template <>
__device__ __forceinline__ bool to_ll<double>(double v, long long& out) {
    if (!(fabs(v) < 4.0e18) || v != rint(v)) return false;
    out = (long long)v;
    return true;
}
template <>
__device__ __forceinline__ bool to_ll<float>(float v, long long& out) {
    if (!(fabsf(v) < 4.0e18f) || v != rintf(v)) return false;
    out = (long long)v;
    return true;
}

__device__ __forceinline__ long long warp_min_ll(long long v) {
    for (int o = 16; o > 0; o >>= 1) { long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t < v ? t : v; }
    return v;
}
__device__ __forceinline__ long long warp_max_ll(long long v) {
    for (int o = 16; o > 0; o >>= 1) { long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
}

// exclusive block scan of one int per thread (blockDim.x threads); returns the exclusive prefix,
// *total receives the block total.  red: >= 33 ints of shared memory.
__device__ __forceinline__ int block_excl_scan(int v, int* red, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) red[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0;
        int winc = w;
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
        red[lane] = winc - w;
        if (lane == 31) red[32] = winc;
    }
    __syncthreads();
    int res = red[wid] + inc - v;
    *total = red[32];
    __syncthreads();
    return res;
}

// One CTA per sample column.  NTHR threads, BMW bitmap words: columns whose value range does not fit the bitmap are
// appended to over_list (their list positions) for the next tier.  `cols` (optional) lists the positions to process.
// The column is read from DRAM once: with 2 x 148 resident CTAs the columns in flight (2 x 148 x 8r bytes, 71 MB at
// 30k genes) stay in L2 for the second pass, and STASH keeps (v - min) as u16 in shared memory for the third.
template <typename T, typename RT, int NTHR, int BMW, bool STASH>
__global__ void __launch_bounds__(NTHR, (BMW == BM_WORDS) ? 1 : 2)
rank_columns_kernel(const T* __restrict__ data, int64_t r, int64_t ld, int64_t col0, const int32_t* __restrict__ cols,
                    const int32_t* __restrict__ src_col, const int32_t* __restrict__ sample_id,
                    const int32_t* __restrict__ slot_of_sample, RT* __restrict__ ranks, int64_t rpad,
                    int* max_distinct, int* flags, int* over_count, int32_t* over_list) {
    constexpr bool ZB = (BMW == BM_WORDS_S);                 // small tier: bitmap indexed by the value itself
    constexpr int BMP = BMW / BM_GROUP;                      // prefix entries
    constexpr int BM_ITEMS = (BMP + NTHR - 1) / NTHR;
    constexpr int RK_THREADS_L = NTHR;
    extern __shared__ uint32_t sm[];
    uint32_t* bm = sm;                   // [BMW]
    uint32_t* pre = sm + BMW;            // [BMP]
    int* red = (int*)(pre + BMP);        // [40]
    long long* redl = (long long*)(red + 40);  // [64]
    uint16_t* stash = (uint16_t*)(redl + 64);  // [r] when STASH

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t j = cols ? (int64_t)cols[blockIdx.x] : col0 + blockIdx.x;   // position in this rank's column list
    const int64_t s = sample_id[j];               // original sample index (slot lookup, fallback list)
    const T* __restrict__ col = data + ld * (int64_t)src_col[j];

    long long mn = 0;
    int nwords;
    if (ZB) {
        // small tier, ONE pass over global memory: counts are non-negative and small, so the bitmap is indexed by the
        // value itself (no minimum to subtract: the dense rank does not depend on the offset).  A value outside
        // [0, 32 * BMW) sends the column to the next tier, a non-integral one flags the matrix.
        for (int w = tid; w < BMW; w += RK_THREADS_L) bm[w] = 0u;
        __syncthreads();
        int bad = 0, oor = 0;
        for (int64_t g = tid; g < r; g += RK_THREADS_L) {
            long long v;
            if (!to_ll<T>(col[g], v)) { bad = 1; continue; }
            if ((unsigned long long)v >= (unsigned long long)BMW * 32ull) { oor = 1; continue; }
            const uint32_t k = (uint32_t)v;
            if (STASH) stash[g] = (uint16_t)k;
            const uint32_t bit = 1u << (k & 31);
            // sparse counts put most genes on the same few values: test first, the atomic is the rare case
            if (!(((volatile uint32_t*)bm)[k >> 5] & bit)) atomicOr(&bm[k >> 5], bit);
        }
        const int any_bad = __syncthreads_or(bad);      // (the result is a truth value, not a bitwise OR)
        const int any_oor = __syncthreads_or(oor);
        if (any_bad) {  // non-integral value: rank compression is not valid under the 0.1 tie band
            if (tid == 0) atomicExch(&flags[0], 1);
            return;
        }
        if (any_oor) {  // next tier
            if (tid == 0) { int k = atomicAdd(over_count, 1); over_list[k] = (int32_t)j; }
            return;
        }
        nwords = BMW;
    } else {
    // phase 1: min / max / integrality
    long long mx = LLONG_MIN;
    mn = LLONG_MAX;
    int bad = 0;
    for (int64_t g = tid; g < r; g += RK_THREADS_L) {
        long long v;
        if (!to_ll<T>(col[g], v)) bad = 1;
        else { mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
    }
    mn = warp_min_ll(mn); mx = warp_max_ll(mx);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) { redl[wid] = mn; redl[32 + wid] = mx; red[wid] = bad; }
    __syncthreads();
    if (wid == 0) {
        const bool have = lane < (NTHR >> 5);
        mn = warp_min_ll(have ? redl[lane] : LLONG_MAX); mx = warp_max_ll(have ? redl[32 + lane] : LLONG_MIN);
        bad = __any_sync(0xffffffffu, have ? red[lane] : 0);
        if (lane == 0) { redl[0] = mn; redl[32] = mx; red[0] = bad; }
    }
    __syncthreads();
    mn = redl[0]; mx = redl[32]; bad = red[0];
    __syncthreads();
    if (bad) {  // non-integral value: rank compression is not valid under the 0.1 tie band
        if (tid == 0) atomicExch(&flags[0], 1);
        return;
    }
    const unsigned long long range = (unsigned long long)mx - (unsigned long long)mn;
    if (range >= (unsigned long long)BMW * 32ull) {   // next tier
        if (tid == 0) { int k = atomicAdd(over_count, 1); over_list[k] = (int32_t)j; }
        return;
    }
    // phase 2: presence bitmap of (v - min)
    nwords = (int)(range >> 5) + 1;
    {
        const int ng = (nwords + BM_GROUP - 1) / BM_GROUP;
        for (int w = tid; w < ng * BM_GROUP; w += RK_THREADS_L) bm[w] = 0u;
    }
    __syncthreads();
    for (int64_t g = tid; g < r; g += RK_THREADS_L) {
        long long v; to_ll<T>(col[g], v);
        const uint32_t k = (uint32_t)((unsigned long long)v - (unsigned long long)mn);
        if (STASH) stash[g] = (uint16_t)k;
        const uint32_t bit = 1u << (k & 31);
        // sparse counts put most genes on the same few values: test first, the atomic is the rare case
        if (!(((volatile uint32_t*)bm)[k >> 5] & bit)) atomicOr(&bm[k >> 5], bit);
    }
    __syncthreads();
    }
    const int ngroups = (nwords + BM_GROUP - 1) / BM_GROUP;
    // phase 3: exclusive prefix of popcounts per group of BM_GROUP words
    int loc[BM_ITEMS];
    int sum = 0;
#pragma unroll
    for (int q = 0; q < BM_ITEMS; ++q) {
        const int grp = tid * BM_ITEMS + q;
        int cnt = 0;
        if (grp < ngroups) {
#pragma unroll
            for (int w = 0; w < BM_GROUP; ++w) cnt += __popc(bm[grp * BM_GROUP + w]);
        }
        loc[q] = sum; sum += cnt;
    }
    int total;
    const int base = block_excl_scan(sum, red, &total);
#pragma unroll
    for (int q = 0; q < BM_ITEMS; ++q) {
        const int grp = tid * BM_ITEMS + q;
        if (grp < ngroups) pre[grp] = (uint32_t)(base + loc[q]);
    }
    if (tid == 0) atomicMax(max_distinct, total);
    __syncthreads();
    // phase 4: dense rank lookup
    const int64_t slot = slot_of_sample[s];
    RT* __restrict__ out = ranks + slot * rpad;
    for (int64_t g = tid; g < r; g += RK_THREADS_L) {
        uint32_t k;
        if (STASH) k = stash[g];
        else { long long v; to_ll<T>(col[g], v); k = (uint32_t)((unsigned long long)v - (unsigned long long)mn); }
        const uint32_t wq = k >> 5;
        uint32_t rk = pre[wq / BM_GROUP];
        for (uint32_t w = (wq / BM_GROUP) * BM_GROUP; w < wq; ++w) rk += __popc(bm[w]);
        rk += __popc(bm[wq] & ((1u << (k & 31)) - 1u));
        out[g] = (RT)rk;
    }
}

// ---- small tier, specialised: values in [0, 65536), up to RK_STASH_MAX_R genes, u16 ranks ---------------------------
// The common case (counts; single-cell data in particular).  One CTA per sample column, ONE pass over global memory:
// two genes per thread and iteration, the values go into a presence bitmap indexed by the value itself and are parked
// as u16 pairs in shared memory; a per-word prefix of the bitmap popcounts then turns every parked value into its dense
// rank with two shared-memory reads, and the ranks leave as u16 pairs.  32-bit index arithmetic throughout.
template <typename T>
__global__ void __launch_bounds__(RK_THREADS, 2)
rank_small_kernel(const T* __restrict__ data, int r, int64_t ld, int64_t col0, const int32_t* __restrict__ src_col,
                  const int32_t* __restrict__ sample_id, const int32_t* __restrict__ slot_of_sample,
                  uint16_t* __restrict__ ranks, int64_t rpad, int* max_distinct, int* flags, int* over_count,
                  int32_t* over_list) {
    extern __shared__ uint32_t sm[];
    uint32_t* bm = sm;                                   // [BM_WORDS_S] presence bitmap of the values
    uint32_t* prew = sm + BM_WORDS_S;                    // [BM_WORDS_S] distinct values below word w
    int* red = (int*)(prew + BM_WORDS_S);                // [40]
    uint32_t* stash = (uint32_t*)(red + 40);             // [(r + 1) / 2] two parked values per word
    const int tid = threadIdx.x;
    const int64_t j = col0 + blockIdx.x;                 // position in this rank's column list
    const T* __restrict__ col = data + ld * (int64_t)src_col[j];
    for (int w = tid; w < BM_WORDS_S; w += RK_THREADS) bm[w] = 0u;
    __syncthreads();
    const int npair = (r + 1) >> 1;
    int bad = 0, oor = 0;
    for (int i = tid; i < npair; i += RK_THREADS) {
        const int g = 2 * i;
        const bool two = g + 1 < r;
        long long v0 = 0, v1 = 0;
        if (!to_ll<T>(col[g], v0)) { bad = 1; v0 = 0; }
        if (two && !to_ll<T>(col[g + 1], v1)) { bad = 1; v1 = 0; }
        if (((unsigned long long)v0 | (unsigned long long)v1) >= (unsigned long long)BM_WORDS_S * 32ull) { oor = 1; continue; }
        const uint32_t k0 = (uint32_t)v0, k1 = (uint32_t)v1;
        stash[i] = k0 | (k1 << 16);
        // sparse counts put most genes on the same few values: test first, the atomic is the rare case
        const uint32_t b0 = 1u << (k0 & 31), b1 = 1u << (k1 & 31);
        if (!(((volatile uint32_t*)bm)[k0 >> 5] & b0)) atomicOr(&bm[k0 >> 5], b0);
        if (two && !(((volatile uint32_t*)bm)[k1 >> 5] & b1)) atomicOr(&bm[k1 >> 5], b1);
    }
    const int any_bad = __syncthreads_or(bad);
    const int any_oor = __syncthreads_or(oor);
    if (any_bad) {  // non-integral value: rank compression is not valid under the 0.1 tie band
        if (tid == 0) atomicExch(&flags[0], 1);
        return;
    }
    if (any_oor) {  // next tier
        if (tid == 0) { int k = atomicAdd(over_count, 1); over_list[k] = (int32_t)j; }
        return;
    }
    {   // exclusive prefix of the popcounts, two bitmap words per thread (BM_WORDS_S == 2 * RK_THREADS)
        const int c0 = __popc(bm[2 * tid]), c1 = __popc(bm[2 * tid + 1]);
        int total;
        const int base = block_excl_scan(c0 + c1, red, &total);
        prew[2 * tid] = (uint32_t)base;
        prew[2 * tid + 1] = (uint32_t)(base + c0);
        if (tid == 0) atomicMax(max_distinct, total);
    }
    __syncthreads();
    const int64_t slot = slot_of_sample[sample_id[j]];
    uint32_t* __restrict__ out = reinterpret_cast<uint32_t*>(ranks + slot * rpad);   // rpad is a multiple of 64
    for (int i = tid; i < npair; i += RK_THREADS) {
        const uint32_t st = stash[i];
        const uint32_t k0 = st & 0xffffu, k1 = st >> 16;
        const uint32_t r0 = prew[k0 >> 5] + (uint32_t)__popc(bm[k0 >> 5] & ((1u << (k0 & 31)) - 1u));
        const uint32_t r1 = prew[k1 >> 5] + (uint32_t)__popc(bm[k1 >> 5] & ((1u << (k1 & 31)) - 1u));
        out[i] = r0 | (r1 << 16);
    }
}
static_assert(BM_WORDS_S == 2 * RK_THREADS, "rank_small_kernel scans two bitmap words per thread");

// ---- fallback: sort-based dense rank in a global-memory scratch (rare: value range > bitmap) ----
template <typename T, typename RT>
__global__ void __launch_bounds__(RK_THREADS, 1)
rank_fallback_kernel(const T* __restrict__ data, int64_t r, int64_t ld, const int32_t* __restrict__ list,
                     const int32_t* __restrict__ src_col, const int32_t* __restrict__ sample_id,
                     const int32_t* __restrict__ slot_of_sample, RT* __restrict__ ranks, int64_t rpad,
                     int* max_distinct, unsigned long long* scratch_keys, uint32_t* scratch_rank, int64_t n) {
    __shared__ int red[40];
    const int tid = threadIdx.x;
    const int64_t j = list[blockIdx.x];
    const int64_t s = sample_id[j];
    const T* __restrict__ col = data + ld * (int64_t)src_col[j];
    unsigned long long* keys = scratch_keys + (int64_t)blockIdx.x * n;
    uint32_t* dr = scratch_rank + (int64_t)blockIdx.x * n;
    for (int64_t g = tid; g < n; g += RK_THREADS) {
        unsigned long long k = ~0ull;
        if (g < r) { long long v; to_ll<T>(col[g], v); k = (unsigned long long)v ^ 0x8000000000000000ull; }
        keys[g] = k;
    }
    __syncthreads();
    for (int64_t k = 2; k <= n; k <<= 1) {
        for (int64_t j = k >> 1; j > 0; j >>= 1) {
            for (int64_t i = tid; i < n; i += RK_THREADS) {
                const int64_t l = i ^ j;
                if (l > i) {
                    const bool asc = (i & k) == 0;
                    const unsigned long long a = keys[i], b = keys[l];
                    if ((a > b) == asc) { keys[i] = b; keys[l] = a; }
                }
            }
            __syncthreads();
        }
    }
    // dense rank of each sorted position: (number of distinct keys <= it) - 1
    const int64_t per = (r + RK_THREADS - 1) / RK_THREADS;
    const int64_t lo = (int64_t)tid * per, hi = (lo + per < r) ? lo + per : r;
    int cnt = 0;
    for (int64_t i = lo; i < hi; ++i) cnt += (i == 0 || keys[i] != keys[i - 1]);
    int total;
    int base = block_excl_scan(cnt, red, &total);
    for (int64_t i = lo; i < hi; ++i) {
        base += (i == 0 || keys[i] != keys[i - 1]);
        dr[i] = (uint32_t)(base - 1);
    }
    if (tid == 0) atomicMax(max_distinct, total);
    __syncthreads();
    const int64_t slot = slot_of_sample[s];
    RT* __restrict__ out = ranks + slot * rpad;
    for (int64_t g = tid; g < r; g += RK_THREADS) {
        long long v; to_ll<T>(col[g], v);
        const unsigned long long key = (unsigned long long)v ^ 0x8000000000000000ull;
        int64_t a = 0, b = r;  // lower_bound
        while (a < b) { const int64_t m = (a + b) >> 1; if (keys[m] < key) a = m + 1; else b = m; }
        out[g] = (RT)dr[a];
    }
}

static size_t rank_smem_bytes(int bmw) { return (size_t)(bmw + bmw / BM_GROUP + 40) * 4 + 64 * 8 + 16; }

// tier 0 (small bitmap, 2 CTAs/SM) over list positions [col0, col0+ncols): overflow -> wide_list / flags[3]
// tier 1 (wide bitmap, 1 CTA/SM) over wide_list: overflow -> fallback_list / flags[1]
template <typename T, typename RT>
static cudaError_t launch_rank_t(const void* data, int64_t r, int64_t ld, int64_t col0, int ncols, const int32_t* cols,
                                 bool wide, const int32_t* src_col, const int32_t* sample_id,
                                 const int32_t* slot_of_sample, void* ranks, int64_t rpad, int* max_distinct, int* flags,
                                 int* over_count, int32_t* over_list, cudaStream_t st) {
    if (ncols <= 0) return cudaSuccess;
    cudaError_t e;
    if (wide) {
        const size_t smem = rank_smem_bytes(BM_WORDS);
        e = cudaFuncSetAttribute(rank_columns_kernel<T, RT, RK_THREADS, BM_WORDS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        rank_columns_kernel<T, RT, RK_THREADS, BM_WORDS, false><<<ncols, RK_THREADS, smem, st>>>(
            (const T*)data, r, ld, col0, cols, src_col, sample_id, slot_of_sample, (RT*)ranks, rpad, max_distinct, flags,
            over_count, over_list);
    } else if (r <= RK_STASH_MAX_R && sizeof(RT) == 2 && cols == nullptr) {
        const size_t smem = (size_t)(2 * BM_WORDS_S + 40) * 4 + (size_t)((r + 1) / 2) * 4 + 16;
        e = cudaFuncSetAttribute(rank_small_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        rank_small_kernel<T><<<ncols, RK_THREADS, smem, st>>>((const T*)data, (int)r, ld, col0, src_col, sample_id, slot_of_sample,
                                                              (uint16_t*)ranks, rpad, max_distinct, flags, over_count, over_list);
    } else if (r <= RK_STASH_MAX_R) {
        const size_t smem = rank_smem_bytes(BM_WORDS_S) + (size_t)r * 2 + 16;
        e = cudaFuncSetAttribute(rank_columns_kernel<T, RT, RK_THREADS, BM_WORDS_S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        rank_columns_kernel<T, RT, RK_THREADS, BM_WORDS_S, true><<<ncols, RK_THREADS, smem, st>>>(
            (const T*)data, r, ld, col0, cols, src_col, sample_id, slot_of_sample, (RT*)ranks, rpad, max_distinct, flags,
            over_count, over_list);
    } else {
        const size_t smem = rank_smem_bytes(BM_WORDS_S);
        rank_columns_kernel<T, RT, RK_THREADS, BM_WORDS_S, false><<<ncols, RK_THREADS, smem, st>>>(
            (const T*)data, r, ld, col0, cols, src_col, sample_id, slot_of_sample, (RT*)ranks, rpad, max_distinct, flags,
            over_count, over_list);
    }
    return cudaGetLastError();
}

cudaError_t reo_launch_rank_columns(const void* data, int dtype, int64_t r, int64_t ld, int64_t col0, int ncols,
                                    const int32_t* cols, int wide, const int32_t* src_col, const int32_t* sample_id,
                                    const int32_t* slot_of_sample, void* ranks, int rank_bytes, int64_t rpad,
                                    int* max_distinct, int* flags, int* over_count, int32_t* over_list, cudaStream_t st) {
#define LAUNCH_RK(T)                                                                                                   \
    return rank_bytes == 2                                                                                             \
               ? launch_rank_t<T, uint16_t>(data, r, ld, col0, ncols, cols, wide != 0, src_col, sample_id, slot_of_sample, \
                                            ranks, rpad, max_distinct, flags, over_count, over_list, st)              \
               : launch_rank_t<T, uint32_t>(data, r, ld, col0, ncols, cols, wide != 0, src_col, sample_id, slot_of_sample, \
                                            ranks, rpad, max_distinct, flags, over_count, over_list, st);
    switch (dtype) {
        case REO_I64: LAUNCH_RK(long long)
        case REO_F64: LAUNCH_RK(double)
        case REO_I32: LAUNCH_RK(int)
        case REO_F32: LAUNCH_RK(float)
        case REO_U16_STAGED: LAUNCH_RK(unsigned short)
        default: return cudaErrorInvalidValue;
    }
#undef LAUNCH_RK
}

cudaError_t reo_launch_rank_fallback(const void* data, int dtype, int64_t r, int64_t ld, const int32_t* fallback_list,
                                     int nfb, const int32_t* src_col, const int32_t* sample_id,
                                     const int32_t* slot_of_sample, void* ranks, int rank_bytes, int64_t rpad,
                                     int* max_distinct, unsigned long long* scratch_keys, uint32_t* scratch_rank,
                                     int64_t rpow2, cudaStream_t st) {
#define LAUNCH_FB2(T, RT)                                                                                             \
    rank_fallback_kernel<T, RT><<<nfb, RK_THREADS, 0, st>>>((const T*)data, r, ld, fallback_list, src_col, sample_id, slot_of_sample, \
                                                            (RT*)ranks, rpad, max_distinct, scratch_keys, scratch_rank, rpow2);
#define LAUNCH_FB(T) if (rank_bytes == 2) { LAUNCH_FB2(T, uint16_t) } else { LAUNCH_FB2(T, uint32_t) }
    switch (dtype) {
        case REO_I64: LAUNCH_FB(long long); break;
        case REO_F64: LAUNCH_FB(double); break;
        case REO_I32: LAUNCH_FB(int); break;
        case REO_F32: LAUNCH_FB(float); break;
        default: return cudaErrorInvalidValue;
    }
#undef LAUNCH_FB
#undef LAUNCH_FB2
    return cudaGetLastError();
}

// ---- bit-plane transpose ------------------------------------------------------------------------
// 32 x 32 bit-matrix transpose across the lanes of a warp (5 butterfly steps): on entry lane s holds row s, on exit
// lane b holds column b, i.e. bit s of the result is bit b of lane s's input.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane) {
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t m = (j == 16) ? 0x0000FFFFu : (j == 8) ? 0x00FF00FFu : (j == 4) ? 0x0F0F0F0Fu
                                                               : (j == 2) ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) ? ((x & ~m) | ((y & ~m) >> j)) : ((x & m) | ((y & m) << j));
    }
    return x;
}

// CTA = one sample word (32 slots) x GC = 256 * PER genes (8 tiles of u16 ranks, 4 tiles of u32 ranks), 256 threads.
// The 32 x GC ranks are loaded with genes along the lanes (1 KB contiguous per slot row) and parked in shared memory;
// a warp then takes GC / 8 genes with lane = sample slot: it packs GPW = 32 / FW genes into a word per lane -- FW-bit
// fields [rank : coin], FW >= planes -- and transposes the 32 x 32 bits across its lanes, so that lane (gene k, plane p)
// ends up with the finished 32-sample word.  Cost per (gene, sample): the coin hash (7 integer ops) + ~5 instructions
// for unpacking, packing and transposing (the first version spent 3 instructions per plane on ballots).  Output goes
// through shared memory: 256-byte coalesced stores per (tile, plane).
template <typename RT, int FW>
__global__ void __launch_bounds__(256)
bitplanes_kernel(const RT* __restrict__ ranks, int64_t rpad, int64_t r,
                 const int32_t* __restrict__ sample_of_slot, int w_lo, int w_n, int w_stride, int NP, uint32_t seed_lo,
                 uint32_t seed_hi, uint32_t* __restrict__ planes) {
    // words [w_lo, w_lo + w_n) of the staged order are written at local index (w - w_lo) with w_stride words per tile
    constexpr int PER = 4 / sizeof(RT);                 // ranks per 32-bit word
    constexpr int GC = 256 * PER;                       // genes per CTA
    constexpr int TPC = GC / REO_TILE;                  // gene tiles per CTA
    constexpr int GW = GC / 8;                          // genes per warp
    constexpr int LDW = 256 + 1;                        // row stride in words, odd: conflict-free column reads
    constexpr int GPW = 32 / FW;                        // genes per transposed word
    constexpr int OS = REO_TILE + 8;                    // plane stride of the output staging
    extern __shared__ uint32_t bp_sm[];
    uint32_t* tile = bp_sm;                             // [32 slots][LDW]
    uint32_t* ghash = tile + 32 * LDW;                  // [GC] inner hash of the coin, per gene
    uint32_t* outs = ghash + GC;                        // [TPC][NP][OS]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t gbase = (int64_t)blockIdx.x * GC;
    for (int i = tid; i < GC; i += 256) ghash[i] = reo_mix32(seed_lo ^ ((uint32_t)(gbase + i) * 0x9E3779B1u));
    for (int wl = blockIdx.y; wl < w_n; wl += gridDim.y) {   // gridDim.y is capped at 65535
        __syncthreads();
        const int so = sample_of_slot[(int64_t)(w_lo + wl) * 32 + lane];      // this lane's sample (pack phase), -1 = pad slot
        const int64_t g0 = gbase + (int64_t)tid * PER;
#pragma unroll 4
        for (int row = 0; row < 32; ++row) {
            const int so_r = __shfl_sync(0xffffffffu, so, row);
            uint32_t v = 0u;
            if (so_r >= 0 && g0 < rpad)
                v = *reinterpret_cast<const uint32_t*>(ranks + ((int64_t)(w_lo + wl) * 32 + row) * rpad + g0);
            tile[row * LDW + tid] = v;
        }
        __syncthreads();
        const uint32_t sterm = (uint32_t)so * 0x85EBCA77u + seed_hi;
        const uint32_t slotmask = __ballot_sync(0xffffffffu, so >= 0);
        const int myp = lane % FW, myk = lane / FW;     // after the transpose: this lane's plane and gene in the group
        const uint32_t* trow = tile + lane * LDW;
#pragma unroll 2
        for (int grp = 0; grp < GW / GPW; ++grp) {   // (full unrolling measured slower: 2.64 vs 2.47 ms of staging on 30k x 20k)
            uint32_t x = 0u;
#pragma unroll
            for (int k = 0; k < GPW; ++k) {
                const int gi = wid * GW + grp * GPW + k;                      // gene within the CTA
                uint32_t rk = trow[gi / PER];
                if (PER == 2) rk = (gi & 1) ? (rk >> 16) : (rk & 0xffffu);
                const uint32_t coin = reo_mix32_top(ghash[gi] + sterm);
                x |= ((rk << 1) | coin) << (k * FW);
            }
            x = warp_transpose32(x, lane) & slotmask;                        // pad slots contribute nothing
            const int gq = wid * GW + grp * GPW + myk;                        // this lane's gene within the CTA
            if (gbase + gq >= r) x = 0u;                                      // genes >= r are not ranked
            if (myp < NP) outs[((gq >> 6) * NP + myp) * OS + (gq & 63)] = x;
        }
        __syncthreads();
        for (int idx = tid; idx < TPC * NP * REO_TILE; idx += 256) {
            const int tp = idx / REO_TILE, l = idx - tp * REO_TILE;           // tp = tile-in-CTA * NP + plane
            const int tl = tp / NP;
            const int64_t t = (int64_t)blockIdx.x * TPC + tl;
            if (t * REO_TILE < rpad)
                planes[((size_t)t * w_stride + wl) * NP * REO_TILE + (size_t)(tp - tl * NP) * REO_TILE + l] = outs[tp * OS + l];
        }
    }
}

cudaError_t reo_launch_bitplanes(const void* ranks, int rank_bytes, int64_t rpad, int64_t r, const int32_t* sample_of_slot,
                                 int NT, int w_lo, int w_n, int w_stride, int NP, uint32_t seed_lo, uint32_t seed_hi,
                                 uint32_t* planes, cudaStream_t st) {
    if (w_n <= 0) return cudaSuccess;
    const int per = 4 / rank_bytes, gc = 256 * per, tpc = gc / REO_TILE;
    dim3 grid((NT + tpc - 1) / tpc, std::min(w_n, 65535)), block(256);
    const size_t smem = ((size_t)32 * 257 + gc + (size_t)tpc * NP * (REO_TILE + 8)) * sizeof(uint32_t);
#define LAUNCH_BP(RT, FW)                                                                                              \
    do {                                                                                                               \
        cudaError_t e_ = cudaFuncSetAttribute(bitplanes_kernel<RT, FW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e_ != cudaSuccess) return e_;                                                                              \
        bitplanes_kernel<RT, FW><<<grid, block, smem, st>>>((const RT*)ranks, rpad, r, sample_of_slot, w_lo, w_n, w_stride, NP, \
                                                            seed_lo, seed_hi, planes);                                 \
    } while (0)
#define LAUNCH_BPN(RT)                                                                                                 \
    if (NP <= 8) LAUNCH_BP(RT, 8); else if (NP <= 16) LAUNCH_BP(RT, 16); else LAUNCH_BP(RT, 32)
    if (rank_bytes == 2) { LAUNCH_BPN(uint16_t); } else { LAUNCH_BPN(uint32_t); }
#undef LAUNCH_BPN
#undef LAUNCH_BP
    return cudaGetLastError();
}

// ---- column panel gather ------------------------------------------------------------------------
// panel[tc][w][p][l] = planes[col_gene/64][w][p][col_gene%64]; pad columns (col_gene < 0) -> 0.
__global__ void __launch_bounds__(256)
gather_panel_kernel(const uint32_t* __restrict__ planes, int W, int NP, const int32_t* __restrict__ col_gene,
                    uint32_t* __restrict__ panel) {
    const int tc = blockIdx.x, l = threadIdx.x & 63;
    const int g = col_gene[tc * REO_TILE + l];
    const int total = W * NP;
    for (int wp = blockIdx.y * 4 + (threadIdx.x >> 6); wp < total; wp += gridDim.y * 4) {
        uint32_t v = 0u;
        if (g >= 0) v = planes[((size_t)(g >> 6) * total + wp) * REO_TILE + (g & 63)];
        panel[((size_t)tc * total + wp) * REO_TILE + l] = v;
    }
}

cudaError_t reo_launch_gather_panel(const uint32_t* planes, int W, int NP, const int32_t* col_gene, int ntc,
                                    uint32_t* panel, cudaStream_t st) {
    if (ntc <= 0) return cudaSuccess;
    int gy = (W * NP + 3) / 4;
    if (gy > 64) gy = 64;
    dim3 grid(ntc, gy);
    gather_panel_kernel<<<grid, 256, 0, st>>>(planes, W, NP, col_gene, panel);
    return cudaGetLastError();
}


// ---- float path staging ----------------------------------------------------------
// planes[t][w] = [64 coin words][32 samples][64 genes] FP64 raw values (pad slots / pad genes = 0.0)
template <typename T>
__global__ void __launch_bounds__(256)
fstage_kernel(const T* __restrict__ data, int64_t r, int64_t ld, const int32_t* __restrict__ sample_of_slot,
              const int32_t* __restrict__ col_of_sample, int w_lo, int w_n, int w_stride, uint32_t seed_lo, uint32_t seed_hi,
              uint32_t* __restrict__ planes) {
    const int t = blockIdx.x, l = threadIdx.x, y = threadIdx.y;
    const int64_t g = (int64_t)t * REO_TILE + l;
    for (int wl = blockIdx.y; wl < w_n; wl += gridDim.y) {      // gridDim.y is capped at 65535
    const int w = w_lo + wl;
    uint32_t* base = planes + ((size_t)t * w_stride + wl) * REO_FLT_OPWORDS;
    double* vals = reinterpret_cast<double*>(base + REO_TILE);
    for (int sidx = y; sidx < 32; sidx += 4) {
        const int so = sample_of_slot[(int64_t)w * 32 + sidx];
        double v = 0.0;
        if (so >= 0 && g < r) v = (double)data[(int64_t)col_of_sample[so] * ld + g];
        vals[sidx * REO_TILE + l] = v;
    }
    if (y == 0) {
        uint32_t coin = 0u;
        if (g < r) {
            for (int sidx = 0; sidx < 32; ++sidx) {
                const int so = sample_of_slot[(int64_t)w * 32 + sidx];
                if (so >= 0) coin |= reo_coin_u(seed_lo, seed_hi, (uint32_t)g, (uint32_t)so) << sidx;
            }
        }
        base[l] = coin;
    }
    }
}

cudaError_t reo_launch_fstage(const void* data, int dtype, int64_t r, int64_t ld, const int32_t* sample_of_slot,
                              const int32_t* col_of_sample, int NT, int w_lo, int w_n, int w_stride, uint32_t seed_lo,
                              uint32_t seed_hi, uint32_t* planes, cudaStream_t st) {
    if (w_n <= 0) return cudaSuccess;
    dim3 grid(NT, std::min(w_n, 65535)), block(REO_TILE, 4);
    switch (dtype) {
        case REO_F64: fstage_kernel<double><<<grid, block, 0, st>>>((const double*)data, r, ld, sample_of_slot, col_of_sample, w_lo, w_n, w_stride, seed_lo, seed_hi, planes); break;
        case REO_F32: fstage_kernel<float><<<grid, block, 0, st>>>((const float*)data, r, ld, sample_of_slot, col_of_sample, w_lo, w_n, w_stride, seed_lo, seed_hi, planes); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// float panel gather: panel[tc][w] = [coin words of the listed genes][32][64] values of the listed genes
__global__ void __launch_bounds__(256)
gather_panel_flt_kernel(const uint32_t* __restrict__ planes, int W, const int32_t* __restrict__ col_gene,
                        uint32_t* __restrict__ panel) {
    const int tc = blockIdx.x, l = threadIdx.x & 63, y = threadIdx.x >> 6;
    const int g = col_gene[tc * REO_TILE + l];
    for (int w = blockIdx.y; w < W; w += gridDim.y) {            // gridDim.y is capped at 65535
    const uint32_t* src = g >= 0 ? planes + ((size_t)(g >> 6) * W + w) * REO_FLT_OPWORDS : nullptr;
    uint32_t* dst = panel + ((size_t)tc * W + w) * REO_FLT_OPWORDS;
    if (y == 0) dst[l] = src ? src[g & 63] : 0u;
    const double* sv = src ? reinterpret_cast<const double*>(src + REO_TILE) : nullptr;
    double* dv = reinterpret_cast<double*>(dst + REO_TILE);
    for (int sidx = y; sidx < 32; sidx += 4) dv[sidx * REO_TILE + l] = sv ? sv[sidx * REO_TILE + (g & 63)] : 0.0;
    }
}

cudaError_t reo_launch_gather_panel_flt(const uint32_t* planes, int W, const int32_t* col_gene, int ntc, uint32_t* panel,
                                        cudaStream_t st) {
    if (ntc <= 0) return cudaSuccess;
    dim3 grid(ntc, std::min(W, 65535));
    gather_panel_flt_kernel<<<grid, 256, 0, st>>>(planes, W, col_gene, panel);
    return cudaGetLastError();
}


// ---- K1 sharded over ranks: gathered[q][t][wl][wb] (rank q staged words q*wq .. ) -> planes[t][w][wb] ----
__global__ void __launch_bounds__(256)
unshard_planes_kernel(const uint4* __restrict__ gathered, uint4* __restrict__ planes, int NT, int W, int wq, int wb4) {
    // one CTA per (tile, word); wb4 = operand words / 4 (16-byte units)
    const int t = blockIdx.x;
    for (int w = blockIdx.y; w < W; w += gridDim.y) {            // gridDim.y is capped at 65535
        const int q = w / wq, wl = w - q * wq;
        const uint4* src = gathered + (((size_t)q * NT + t) * wq + wl) * wb4;
        uint4* dst = planes + ((size_t)t * W + w) * wb4;
        for (int i = threadIdx.x; i < wb4; i += blockDim.x) dst[i] = src[i];
    }
}
// ---- planes in use per sample word --------------------------------------------------------------
// word_np[w] = 1 + the highest rank plane of word w in which any gene has a bit (1 = the coin plane only).  The dense
// ranks of a sample need only as many bits as it has distinct values, and B is the maximum over ALL samples: the 32
// samples of most words need fewer planes (single-cell counts: 6 bits for all but a few cells, B = 7), and a borrow-chain
// step over an all-zero plane is the identity, so the pair kernel skips it.  One CTA per word, top plane first.
__global__ void __launch_bounds__(256)
word_planes_kernel(const uint32_t* __restrict__ planes, int NT, int W, int NP, uint8_t* __restrict__ word_np, int* __restrict__ run_sum) {
    const int w = blockIdx.x;
    int np = 1;
    for (int p = NP - 1; p >= 1; --p) {
        uint32_t any = 0u;
        for (int i = threadIdx.x; i < NT * REO_TILE; i += 256) {
            const int t = i >> 6, l = i & 63;
            any |= planes[(((size_t)t * W + w) * NP + p) * REO_TILE + l];
        }
        if (__syncthreads_or(any != 0u)) { np = p + 1; break; }
    }
    if (threadIdx.x == 0) {
        word_np[w] = (uint8_t)np;
        atomicAdd(run_sum, (np < NP && NP > 2) ? NP - 1 : NP);   // planes the pair kernel will run for this word
    }
}

cudaError_t reo_launch_word_planes(const uint32_t* planes, int NT, int W, int NP, uint8_t* word_np, int* run_sum, cudaStream_t st) {
    if (W <= 0) return cudaSuccess;
    word_planes_kernel<<<W, 256, 0, st>>>(planes, NT, W, NP, word_np, run_sum);
    return cudaGetLastError();
}

cudaError_t reo_launch_unshard_planes(const uint32_t* gathered, uint32_t* planes, int NT, int W, int wq, int wb,
                                      cudaStream_t st) {
    dim3 grid(NT, std::min(W, 65535));
    unshard_planes_kernel<<<grid, 256, 0, st>>>((const uint4*)gathered, (uint4*)planes, NT, W, wq, wb / 4);
    return cudaGetLastError();
}
