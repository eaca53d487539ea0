// reo_pairs.cu -- K2: the pair-count / stable-REO class / per-gene 9-bin table kernel.
//
// Reference semantics (src/RankCompV3.jl): for every gene i and every reference gene j != i
//   nre  = #{ s in group k   : is_greater(x[i,s], x[j,s]) }                       (src:372-373)
//   rest = #{ s not in group : is_greater(x[i,s], x[j,s]) }                       (src:374)
//   ic = nre  >= thr1 ? 3 : (n1 - nre  >= thr1 ? 1 : 2)                            (src:376)
//   it = rest >= thr2 ? 3 : (n2 - rest >= thr2 ? 1 : 2)                            (src:377)
//   table[i][3*(ic-1)+it] += 1                                                     (src:385, 403)
// The G x G category matrix of the reference (R, src:363) is never materialised.
//
// B200 mapping.  Samples are bit-sliced (32 per word, see reo_stage.cu).  For one gene pair and one
// word, "x > y" over 32 samples is the borrow of y - x rippled through the B rank planes:
//     borrow' = (x & ~y) | (~(x ^ y) & borrow)        -- ONE LOP3 (LUT 0xB2) per plane
// seeded with borrow0 = u_i ^ u_j ^ [i<j], the tie coin, so that equal ranks resolve to the coin
// and the reference's mirror property (src:385-386) holds bit-exactly.  popc() of the final borrow
// is the per-thread integer accumulator.  Cost: (B+1) LOP3 + 1 POPC + 1 IADD per 32 ordered
// (i, j, sample) triples; the ALU (LOP3) pipe is the roofline, no tensor cores.
//
// Tiling: CTA = 64 row genes x 64 column genes, 256 threads, 4x4 pairs per thread (16 independent
// borrow chains per thread hide the 4-cycle ALU latency).  Operand tiles ([plane][64 genes] words)
// are streamed by cp.async.bulk (UBLKCP) into a 3-stage shared-memory ring guarded by mbarriers;
// each thread reads its 4 row words and 4 column words per plane with two LDS.128 (bank-conflict
// free: 8 distinct 16 B row chunks + 4 broadcast column chunks per warp).
// Work items (row tile x chunk of column tiles) are handed out by an atomic counter to a
// persistent grid of 2 CTAs per SM; each item accumulates a 64 x 9 table in shared memory and
// flushes it with integer atomics (order independent, exact).
#include "reo_internal.cuh"

#define PK_THREADS 256
#define PK_KW 4        // sample words per pipeline stage
#define PK_NS 3        // pipeline stages

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint32_t lop3_b2(uint32_t x, uint32_t y, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xB2;" : "=r"(d) : "r"(x), "r"(y), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t lop3_xor3(uint32_t x, uint32_t y, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(x), "r"(y), "r"(c));
    return d;
}

// NPT > 0: planes known at compile time (fully unrolled chain); NPT == 0: runtime p.NP.
template <int NPT>
__global__ void __launch_bounds__(PK_THREADS, 2) reo_pair_kernel(const ReoPairParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int NP = NPT > 0 ? NPT : p.NP;
    const int op_words = NP * REO_TILE;                 // words of one operand tile for one sample word
    const uint32_t op_bytes = (uint32_t)op_words * 4u;
    const int stage_words = 2 * PK_KW * op_words;       // rows then columns
    uint32_t* stages = reinterpret_cast<uint32_t*>(smem_raw);
    int32_t* tab_s = reinterpret_cast<int32_t*>(stages + (size_t)PK_NS * stage_words);  // [64][9]
    uint64_t* full = reinterpret_cast<uint64_t*>(tab_s + REO_TILE * 9 + 2);             // 8-byte aligned below
    full = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(full) + 7) & ~uintptr_t(7));
    int* item_s = reinterpret_cast<int*>(full + PK_NS);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int ty = (warp >> 2) * 8 + (lane >> 2);   // 0..15 : rows ty*4 .. ty*4+3
    const int tx = (warp & 3) * 4 + (lane & 3);     // 0..15 : cols tx*4 .. tx*4+3

    if (tid == 0) {
        for (int s = 0; s < PK_NS; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int nchunks = (p.W + PK_KW - 1) / PK_KW;
    const size_t word_stride = (size_t)p.NP * REO_TILE;          // allocation stride (runtime NP)
    const size_t tile_stride = (size_t)p.W * word_stride;
    uint32_t phase_bits = 0;   // one parity bit per stage
    int issued_total = 0;      // only meaningful in tid 0: steps issued so far (for stage rotation)
    int consumed_total = 0;

    for (;;) {
        if (tid == 0) *item_s = (int)atomicAdd(p.counter, 1u);
        for (int i = tid; i < REO_TILE * 9; i += PK_THREADS) tab_s[i] = 0;
        __syncthreads();
        const int item = *item_s;
        const int nitems = (p.t1 - p.t0) * p.njchunks;
        if (item >= nitems) break;
        const int ti = p.t0 + item / p.njchunks;
        const int jbeg = (item % p.njchunks) * p.jchunk;
        const int jend = min(jbeg + p.jchunk, p.ntc);
        const int nsteps = (jend - jbeg) * nchunks;

        // producer: issue step `st` of this item into the next ring slot
        auto issue = [&](int st) {
            const int J = jbeg + st / nchunks;
            const int ch = st % nchunks;
            const int w0 = ch * PK_KW;
            const int nw = min(PK_KW, p.W - w0);
            const int slot = issued_total % PK_NS;
            uint32_t* dst = stages + (size_t)slot * stage_words;
            mbar_expect_tx(&full[slot], (uint32_t)nw * 2u * op_bytes);
            for (int kk = 0; kk < nw; ++kk) {
                const int w = p.word_order[w0 + kk];
                bulk_g2s(dst + kk * op_words, p.row_planes + (size_t)ti * tile_stride + (size_t)w * word_stride,
                         op_bytes, &full[slot]);
                bulk_g2s(dst + (PK_KW + kk) * op_words,
                         p.col_planes + (size_t)J * tile_stride + (size_t)w * word_stride, op_bytes, &full[slot]);
            }
            issued_total++;
        };
        if (tid == 0) {
            const int pre = min(PK_NS, nsteps);
            for (int st = 0; st < pre; ++st) issue(st);
        }

        const int gi0 = ti * REO_TILE + ty * 4;
        int st = 0;
        for (int J = jbeg; J < jend; ++J) {
            // per-pair orientation masks: all-ones where row gene index < column gene index
            uint32_t omask[4][4];
            int gj[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) gj[b] = p.col_gene[J * REO_TILE + tx * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) omask[a][b] = (gi0 + a < gj[b]) ? 0xffffffffu : 0u;
            uint32_t acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = 0u;
            uint32_t icpack = 0u;

            for (int ch = 0; ch < nchunks; ++ch, ++st) {
                const int slot = consumed_total % PK_NS;
                mbar_wait(&full[slot], (phase_bits >> slot) & 1u);
                phase_bits ^= (1u << slot);
                const uint32_t* srow = stages + (size_t)slot * stage_words;
                const uint32_t* scol = srow + PK_KW * op_words;
                const int w0 = ch * PK_KW;
                const int nw = min(PK_KW, p.W - w0);
                for (int kk = 0; kk < nw; ++kk) {
                    if (w0 + kk == p.WA) {  // group A finished: classify ic, restart the counters
#pragma unroll
                        for (int a = 0; a < 4; ++a)
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                const int nre = (int)acc[a][b] - (int)(omask[a][b] & (uint32_t)p.padA);
                                const uint32_t ic = nre >= p.thrA ? 2u : ((p.nA - nre) >= p.thrA ? 0u : 1u);
                                icpack |= ic << (2 * (a * 4 + b));
                                acc[a][b] = 0u;
                            }
                    }
                    const uint32_t* xr = srow + kk * op_words + ty * 4;
                    const uint32_t* yc = scol + kk * op_words + tx * 4;
                    uint32_t bor[4][4];
                    {
                        const uint4 xv = *reinterpret_cast<const uint4*>(xr);
                        const uint4 yv = *reinterpret_cast<const uint4*>(yc);
                        const uint32_t x[4] = {xv.x, xv.y, xv.z, xv.w};
                        const uint32_t y[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
                        for (int a = 0; a < 4; ++a)
#pragma unroll
                            for (int b = 0; b < 4; ++b) bor[a][b] = lop3_xor3(x[a], y[b], omask[a][b]);
                    }
                    if (NPT > 0) {
#pragma unroll
                        for (int pl = 1; pl < (NPT > 0 ? NPT : 1); ++pl) {
                            const uint4 xv = *reinterpret_cast<const uint4*>(xr + pl * REO_TILE);
                            const uint4 yv = *reinterpret_cast<const uint4*>(yc + pl * REO_TILE);
                            const uint32_t x[4] = {xv.x, xv.y, xv.z, xv.w};
                            const uint32_t y[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
                            for (int a = 0; a < 4; ++a)
#pragma unroll
                                for (int b = 0; b < 4; ++b) bor[a][b] = lop3_b2(x[a], y[b], bor[a][b]);
                        }
                    } else {
#pragma unroll 2
                        for (int pl = 1; pl < NP; ++pl) {
                            const uint4 xv = *reinterpret_cast<const uint4*>(xr + pl * REO_TILE);
                            const uint4 yv = *reinterpret_cast<const uint4*>(yc + pl * REO_TILE);
                            const uint32_t x[4] = {xv.x, xv.y, xv.z, xv.w};
                            const uint32_t y[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
                            for (int a = 0; a < 4; ++a)
#pragma unroll
                                for (int b = 0; b < 4; ++b) bor[a][b] = lop3_b2(x[a], y[b], bor[a][b]);
                        }
                    }
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[a][b] += __popc(bor[a][b]);
                }
                consumed_total++;
                __syncthreads();  // every thread is done with this ring slot
                if (tid == 0 && st + PK_NS < nsteps) issue(st + PK_NS);
            }
            // group B finished: classify it, add to the shared table
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int gi = gi0 + a;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int rest = (int)acc[a][b] - (int)(omask[a][b] & (uint32_t)p.padB);
                    const int it = rest >= p.thrB ? 2 : ((p.nB - rest) >= p.thrB ? 0 : 1);
                    const int ic = (int)((icpack >> (2 * (a * 4 + b))) & 3u);
                    if (gj[b] >= 0 && gi != gj[b] && gi < p.r) {
                        const int sg = p.col_sign ? (int)p.col_sign[J * REO_TILE + tx * 4 + b] : 1;
                        atomicAdd(&tab_s[(ty * 4 + a) * 9 + ic * 3 + it], sg);
                    }
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < REO_TILE * 9; i += PK_THREADS) {
            const int v = tab_s[i];
            const int gi = ti * REO_TILE + i / 9;
            if (v != 0 && gi < p.r) atomicAdd(&p.table[(size_t)gi * 9 + (i % 9)], v);
        }
        __syncthreads();
    }
}

static size_t pair_smem_bytes(int NP) {
    return (size_t)PK_NS * 2 * PK_KW * NP * REO_TILE * 4 + (REO_TILE * 9 + 2) * 4 + 8 + PK_NS * 8 + 16;
}

template <int NPT>
static cudaError_t launch_np(const ReoPairParams& p, int num_sms, cudaStream_t st) {
    const size_t smem = pair_smem_bytes(p.NP);
    cudaError_t e = cudaFuncSetAttribute(reo_pair_kernel<NPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int nitems = (p.t1 - p.t0) * p.njchunks;
    int grid = 2 * num_sms;
    if (grid > nitems) grid = nitems;
    if (grid < 1) return cudaSuccess;
    reo_pair_kernel<NPT><<<grid, PK_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t reo_launch_pairs(const ReoPairParams& p, int num_sms, cudaStream_t st) {
    if (p.t1 <= p.t0 || p.ntc <= 0) return cudaSuccess;
    switch (p.NP) {
#define CASE_NP(n) case n: return launch_np<n>(p, num_sms, st);
        CASE_NP(2) CASE_NP(3) CASE_NP(4) CASE_NP(5) CASE_NP(6) CASE_NP(7) CASE_NP(8) CASE_NP(9) CASE_NP(10)
        CASE_NP(11) CASE_NP(12) CASE_NP(13) CASE_NP(14) CASE_NP(15) CASE_NP(16) CASE_NP(17)
#undef CASE_NP
        default: return launch_np<0>(p, num_sms, st);
    }
}

// ---- small-block debug/parity kernel: one thread per (row, col) pair, straight from the planes ----
__global__ void pair_counts_small_kernel(const uint32_t* __restrict__ planes, int W, int NP,
                                         const int32_t* __restrict__ word_order, int WA,
                                         const int32_t* __restrict__ rows, int nrows,
                                         const int32_t* __restrict__ cols, int ncols, int32_t* nre, int32_t* rest,
                                         int padA, int padB) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nrows * ncols) return;
    const int gi = rows[idx / ncols], gj = cols[idx % ncols];
    const uint32_t om = gi < gj ? 0xffffffffu : 0u;
    const size_t ws = (size_t)NP * REO_TILE, ts = (size_t)W * ws;
    const uint32_t* pi = planes + (size_t)(gi >> 6) * ts + (gi & 63);
    const uint32_t* pj = planes + (size_t)(gj >> 6) * ts + (gj & 63);
    int a = 0, b = 0;
    for (int k = 0; k < W; ++k) {
        const int w = word_order[k];
        const uint32_t* xi = pi + (size_t)w * ws;
        const uint32_t* yj = pj + (size_t)w * ws;
        uint32_t bor = xi[0] ^ yj[0] ^ om;
        for (int pl = 1; pl < NP; ++pl) {
            const uint32_t x = xi[(size_t)pl * REO_TILE], y = yj[(size_t)pl * REO_TILE];
            bor = (x & ~y) | (~(x ^ y) & bor);
        }
        if (k < WA) a += __popc(bor); else b += __popc(bor);
    }
    nre[idx] = a - (int)(om & (uint32_t)padA);
    rest[idx] = b - (int)(om & (uint32_t)padB);
}

cudaError_t reo_launch_pair_counts_small(const ReoStaged& S, const int32_t* word_order, int WA, const int32_t* rows,
                                         int nrows, const int32_t* cols, int ncols, int32_t* nre, int32_t* rest,
                                         int padA, int padB, cudaStream_t st) {
    const int n = nrows * ncols;
    if (n <= 0) return cudaSuccess;
    pair_counts_small_kernel<<<(n + 127) / 128, 128, 0, st>>>(S.planes, S.W, S.NP, word_order, WA, rows, nrows, cols,
                                                              ncols, nre, rest, padA, padB);
    return cudaGetLastError();
}
