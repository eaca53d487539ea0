// reo_pairs2.cu -- K2, second generation: the pair-count / stable-REO class / per-gene 9-bin table kernel.
//
// Reference semantics (src/RankCompV3.jl): for every gene pair i < j (src:366-372, visited ONCE)
//   nre  = #{ s in group k   : is_greater(x[i,s], x[j,s]) }                       (src:372-373)
//   rest = #{ s not in group : is_greater(x[i,s], x[j,s]) }                       (src:374)
//   ic = nre  >= thr1 ? 3 : (n1 - nre  >= thr1 ? 1 : 2),  it likewise             (src:376-377)
//   q = 3*(ic-1)+it is recorded for gene i and the mirror 10 - q for gene j       (src:385-386)
//   table[g][q] = #{ reference genes j : category of (g, j) is q }                 (src:403)
// The G x G category matrix (R, src:363) is never materialised.
//
// What is new against reo_pairs.cu (which stays for the raw-FP64 variant):
//  * mirror property used: inside the symmetric region (row panel == column panel) only tiles I <= J are evaluated
//    and every evaluated pair updates BOTH genes -- half the is_greater work of an all-ordered-pairs sweep, exactly
//    the reference's own visit count;
//  * warp specialisation: 8 consumer warps (64 x 128 CTA tile, 4 x 8 pairs per thread) + 1 producer warp that owns
//    the work counter and streams operand tiles with cp.async.bulk (UBLKCP) through an NS-stage shared-memory ring
//    guarded by full/empty mbarriers; consumers never issue copies, gene ids / signs of a tile travel with its first
//    stage, and a slot is released per consumer warp;
//  * work item = T x T tile block; row AND column tables of a block live in shared memory and are flushed with
//    integer atomics once per item (two named barriers per item, none inside);
//  * items are ordered by supertiles (SS x SS blocks) so that concurrently running CTAs share operand tiles in L2,
//    and supertiles are dealt round-robin to the ranks of a multi-GPU job (tables are summed across ranks).
// Arithmetic is unchanged: bit-sliced borrow chain, ONE LOP3 (0xB2) per rank plane per 32 samples, tie coin as
// plane 0, POPC + IMAD accumulation, lookup-table classification.  Bound: the ALU (LOP3) pipe; no tensor cores.
#include <stdlib.h>

#include <algorithm>

#include "reo_internal.cuh"
#include "reo_ptx.cuh"

#define P2_CWARPS 8
#define P2_CONSUMERS (P2_CWARPS * 32)
#define P2_THREADS (P2_CONSUMERS + 32)
#define P2_CTAS_PER_SM 2
#ifndef P2_SMEM_BUDGET
#define P2_SMEM_BUDGET (111 * 1024)   // per CTA; 2 CTAs + 2 x 1 KB reserved <= 228 KB per SM
#endif
#define P2_LUT_MAX_WORDS 4096
#define P2_MAX_NS 8

#define P2F_FIRST_J 1
#define P2F_LAST_J 2
#define P2F_LAST_ITEM 4
#define P2F_TERM 8
#define P2F_ROW0 16     // pairs with the first column tile update their ROW gene
#define P2F_COL0 32     // ... and their COLUMN gene (symmetric region, J > I)
#define P2F_ROW1 64     // same for the second column tile of the step
#define P2F_COL1 128

// what travels with a ring stage: gene ids / signs of the tile pair (valid on its first stage) and the stage header
struct __align__(128) P2Aux {
    int32_t cgene[128];
    int32_t rgene[64];
    int8_t csgn[128];
    int8_t rsgn[64];
    int I, J, w0, nw, flags, il, jl, rbase, cbase, pad[7];
};
static_assert(sizeof(P2Aux) == 1024, "P2Aux must stay 1 KB");

__device__ __forceinline__ void p2_consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(P2_CONSUMERS) : "memory"); }

// work item n of this rank -> block (bi, bj); false: nothing to do for this index
__device__ __forceinline__ bool p2_decode(const ReoPair2Params& p, int n, int& bi, int& bj) {
    const int ss2 = p.SS * p.SS;
    const int k = n / ss2, li = n - k * ss2;
    const long long s = (long long)k * p.world + p.rank;
    if (s >= p.NSUP) return false;
    const int di = li / p.SS, dj = li - di * p.SS;
    if (s < p.tri) {   // symmetric region: supertile row SI holds supertile columns SI .. Ms-1
        int si = 0;
        long long off = 0;
        while (off + (p.Ms - si) <= s) { off += p.Ms - si; ++si; }
        bi = si * p.SS + di;
        bj = (si + (int)(s - off)) * p.SS + dj;
        return bi < p.NBs && bj < p.NBc && bj >= bi;
    }
    const long long s2 = s - p.tri;
    bi = p.NBs + (int)(s2 / p.Mc) * p.SS + di;
    bj = (int)(s2 % p.Mc) * p.SS + dj;
    return bi < p.NBr && bj < p.NBc;
}

// NPT > 0: planes known at compile time (fully unrolled chain); NPT == 0: runtime p.NP.
// LUT: classification through shared-memory lookup tables (accumulators are shared-memory addresses).
template <int NPT, bool LUT>
__global__ void __launch_bounds__(P2_THREADS, P2_CTAS_PER_SM) reo_pair2_kernel(const ReoPair2Params p) {
    constexpr int NB = 8;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int NP = NPT > 0 ? NPT : p.NP;
    const int KW = p.KW, NS = p.NS, T = p.T;
    const int op_words = NP * REO_TILE;                  // words of one operand tile for one sample word
    const uint32_t op_bytes = (uint32_t)op_words * 4u;
    const int stage_words = 3 * KW * op_words;           // rows, first column tile, second column tile
    const int tabw = T * REO_TILE;                       // genes per table row (one row per bin)
    uint32_t* stages = reinterpret_cast<uint32_t*>(smem_raw);
    P2Aux* aux = reinterpret_cast<P2Aux*>(stages + (size_t)NS * stage_words);
    int32_t* tabR = reinterpret_cast<int32_t*>(aux + NS);            // [9][tabw] rows of the block
    int32_t* tabC = tabR + 9 * tabw;                                 // [9][tabw] columns of the block
    uint64_t* full = reinterpret_cast<uint64_t*>(tabC + 9 * tabw);
    uint64_t* empty = full + NS;
    uint32_t* lut = reinterpret_cast<uint32_t*>(empty + NS);
    // lut layout (words), o = tie-coin orientation of the pair (0: i>j, 1: i<j), SZA/SZB = slots + 1:
    //   lutA[o][v]         at o*SZA + v              -> shared-memory ADDRESS of lutB[o][ic*SZB + 0]
    //   lutB[o][ic*SZB+v]  at 2*SZA + o*3*SZB + ...  -> (3*ic + it) * binstride (byte offset of the bin row)
    const int SZA = p.lutSZA, SZB = p.lutSZB;
    const uint32_t binstride = (uint32_t)tabw * 4u;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], P2_CWARPS); }
        mbar_fence_init();
    }
    for (int i = tid; i < 18 * tabw; i += P2_THREADS) tabR[i] = 0;
    if (LUT) {
        const uint32_t lutB0 = smem_u32(lut) + (uint32_t)(2 * SZA * 4);
        for (int i = tid; i < 2 * SZA; i += P2_THREADS) {
            const int o = i / SZA, v = i - o * SZA;
            lut[i] = lutB0 + (uint32_t)(o * 3 * SZB * 4) + reo_class(v - o * p.padA, p.nA, p.thrA) * (uint32_t)(SZB * 4);
        }
        for (int i = tid; i < 6 * SZB; i += P2_THREADS) {
            const int o = i / (3 * SZB), rem = i - o * 3 * SZB;
            const int ic = rem / SZB, v = rem - ic * SZB;
            lut[2 * SZA + i] = (3u * ic + reo_class(v - o * p.padB, p.nB, p.thrB)) * binstride;
        }
    }
    __syncthreads();

    // =============================== producer warp ===============================
    if (warp == P2_CWARPS) {
        if (lane != 0) return;
        const int nchunks = (p.W + KW - 1) / KW;
        const size_t word_stride = (size_t)p.NP * REO_TILE;
        const size_t tile_stride = (size_t)p.W * word_stride;
        // k-th word of the two-group order -> staged word
        auto word_of = [&](int k) -> int {
            if (k < p.WA) return p.segA0 + k;
            k -= p.WA;
            if (p.mixed) { if (k == 0) return p.mixedW; k -= 1; }
            return k < p.segB0len ? p.segB0 + k : p.segB1 + (k - p.segB0len);
        };
        int slot = 0;
        uint32_t ephase = 1u;       // parity that means "slot is free": passes at once on the first round
        const int nsymp = p.NBs * T;
        for (;;) {
            const int n = (int)atomicAdd(p.counter, 1u);
            if (n >= p.nitems) break;
            int bi, bj;
            if (!p2_decode(p, n, bi, bj)) continue;
            const bool symrow = bi < p.NBs;
            const int I0 = bi * T, J0 = bj * T;
            const int Iend = symrow ? min(I0 + T, p.nsym) : min(I0 + T, p.ntr);
            const int Jend = min(J0 + T, p.ntc);
            const int jphi = (Jend + 1) >> 1;
            int left = 0;            // tile pairs (64 x 128) of this item
            for (int I = I0; I < Iend; ++I) {
                int jplo = J0 >> 1;
                if (symrow) jplo = max(jplo, I >> 1);
                left += max(0, jphi - jplo);
            }
            if (left == 0) continue;
            (void)nsymp;
            for (int I = I0; I < Iend; ++I) {
                int jplo = J0 >> 1;
                if (symrow) jplo = max(jplo, I >> 1);
                for (int jp = jplo; jp < jphi; ++jp) {
                    const int Ja = 2 * jp, Jb = Ja + 1;
                    const bool hasB = Jb < p.ntc;
                    int tf = 0;
                    if (symrow) {
                        if (Ja >= I) tf |= P2F_ROW0 | (Ja > I ? P2F_COL0 : 0);
                        if (hasB && Jb >= I) tf |= P2F_ROW1 | (Jb > I ? P2F_COL1 : 0);
                    } else {
                        tf |= P2F_ROW0 | (hasB ? P2F_ROW1 : 0);
                    }
                    --left;
                    const uint32_t* rbase = p.row_planes + (size_t)I * tile_stride;
                    const uint32_t* cbase = p.col_planes + (size_t)Ja * tile_stride;
                    for (int ch = 0; ch < nchunks; ++ch) {
                        mbar_wait(&empty[slot], ephase);
                        P2Aux* A = &aux[slot];
                        const int w0 = ch * KW;
                        const int nw = min(KW, p.W - w0);
                        const bool first = ch == 0, last = ch == nchunks - 1;
                        A->I = I; A->J = Ja; A->w0 = w0; A->nw = nw;
                        A->il = I - I0; A->jl = Ja - J0; A->rbase = I0 * REO_TILE; A->cbase = J0 * REO_TILE;
                        A->flags = tf | (first ? P2F_FIRST_J : 0) | (last ? P2F_LAST_J : 0) |
                                   ((last && left == 0) ? P2F_LAST_ITEM : 0);
                        uint32_t bytes = (uint32_t)nw * op_bytes * (hasB ? 3u : 2u);
                        if (first) bytes += 256u + 512u + (p.col_sign ? 128u : 0u) + (p.row_sign ? 64u : 0u);
                        mbar_expect_tx(&full[slot], bytes);
                        if (first) {
                            bulk_g2s(A->rgene, p.row_gene + (size_t)I * REO_TILE, 256u, &full[slot]);
                            bulk_g2s(A->cgene, p.col_gene + (size_t)Ja * REO_TILE, 512u, &full[slot]);
                            if (p.col_sign) bulk_g2s(A->csgn, p.col_sign + (size_t)Ja * REO_TILE, 128u, &full[slot]);
                            if (p.row_sign) bulk_g2s(A->rsgn, p.row_sign + (size_t)I * REO_TILE, 64u, &full[slot]);
                        }
                        uint32_t* dst = stages + (size_t)slot * stage_words;
                        // one bulk copy per operand per run of consecutive staged words
                        int kk = 0;
                        while (kk < nw) {
                            const int w = word_of(w0 + kk);
                            int run = 1;
                            while (kk + run < nw && word_of(w0 + kk + run) == w + run) ++run;
                            const uint32_t rb = (uint32_t)run * op_bytes;
                            bulk_g2s(dst + kk * op_words, rbase + (size_t)w * word_stride, rb, &full[slot]);
                            bulk_g2s(dst + (KW + kk) * op_words, cbase + (size_t)w * word_stride, rb, &full[slot]);
                            if (hasB)
                                bulk_g2s(dst + (2 * KW + kk) * op_words, cbase + tile_stride + (size_t)w * word_stride, rb, &full[slot]);
                            kk += run;
                        }
                        if (++slot == NS) { slot = 0; ephase ^= 1u; }
                    }
                }
            }
        }
        // no more work: one terminating stage
        mbar_wait(&empty[slot], ephase);
        aux[slot].flags = P2F_TERM;
        mbar_arrive(&full[slot]);
        return;
    }

    // =============================== consumer warps ===============================
    // warp = 8 x 4 threads: 32 rows x (16 + 16) columns; thread = rows ty*4..+3, columns tx*4..+3 of both column tiles
    const int ty = (warp >> 2) * 8 + (lane >> 2);
    const int tx = (warp & 3) * 4 + (lane & 3);
    uint32_t acc[4][NB];       // 4 * count [+ carried class offset], or lookup-table addresses (LUT)
    uint32_t om = 0u;          // tie-coin orientation [i<j] of this thread's pairs (all-ones / zero) ...
    uint32_t obits = 0u;       // ... and per pair (bit a*NB+b), used only when the 4 x 8 block is not uniform
    bool uniform = true;       // one orientation, every gene real, no self pair
    uint32_t csg0 = 0x01010101u, csg1 = 0x01010101u, rsg = 0x01010101u;   // packed int8 signs
    int curI = 0, curJ = 0;
    uint32_t rowoff = 0u, coloff = 0u;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[a][b] = 0u;
    const uint32_t four = p.one << 2;
    const uint32_t lut_s = smem_u32(lut);
    const uint32_t tabR_s = smem_u32(tabR), tabC_s = smem_u32(tabC);

    // group-A class of a pair -> new accumulator value (count restarts at 0, class carried along)
    auto flushA = [&](uint32_t acc4, int idx) -> uint32_t {
        if (LUT) return lds_u32(acc4);
        const uint32_t o = (obits >> idx) & 1u;
        return reo_class((int)(acc4 >> 2) - (int)(o * (uint32_t)p.padA), p.nA, p.thrA) << 28;
    };
    // (class of A, count of B) -> byte offset of the table bin row
    auto binB = [&](uint32_t acc4, int idx) -> uint32_t {
        if (LUT) return lds_u32(acc4);
        const uint32_t o = (obits >> idx) & 1u;
        const uint32_t it = reo_class((int)((acc4 & 0x0fffffffu) >> 2) - (int)(o * (uint32_t)p.padB), p.nB, p.thrB);
        return (3u * (acc4 >> 28) + it) * binstride;
    };

    int slot = 0;
    uint32_t phase = 0u;
    for (;;) {
        mbar_wait(&full[slot], phase);
        const P2Aux* A = &aux[slot];
        const int flags = A->flags;
        if (flags & P2F_TERM) break;
        const int nw = A->nw, w0 = A->w0;
        if (flags & P2F_FIRST_J) {
            const int4 r4 = *reinterpret_cast<const int4*>(A->rgene + ty * 4);
            const int4 c4 = *reinterpret_cast<const int4*>(A->cgene + tx * 4);
            const int4 d4 = *reinterpret_cast<const int4*>(A->cgene + REO_TILE + tx * 4);
            const int gi[4] = {r4.x, r4.y, r4.z, r4.w};
            const int gj[NB] = {c4.x, c4.y, c4.z, c4.w, d4.x, d4.y, d4.z, d4.w};
            if (p.col_sign) {
                csg0 = *reinterpret_cast<const uint32_t*>(A->csgn + tx * 4);
                csg1 = *reinterpret_cast<const uint32_t*>(A->csgn + REO_TILE + tx * 4);
            }
            if (p.row_sign) rsg = *reinterpret_cast<const uint32_t*>(A->rsgn + ty * 4);
            curI = A->I; curJ = A->J;
            rowoff = (uint32_t)(A->il * REO_TILE + ty * 4) * 4u;
            coloff = (uint32_t)(A->jl * REO_TILE + tx * 4) * 4u;
            // lists ascend inside a region and pads (-1) come last, so the end points decide
            const bool real = (gi[0] >= 0) && (gi[3] >= 0) && (gj[0] >= 0) && (gj[3] >= 0) && (gj[4] >= 0) && (gj[NB - 1] >= 0);
            const bool all_lt = real && (gi[3] < gj[0]);        // every row gene below every column gene
            const bool all_ge = real && (gi[0] > gj[NB - 1]);   // every row gene above every column gene
            uniform = all_lt || all_ge;
            om = all_lt ? 0xffffffffu : 0u;
            obits = all_lt ? 0xffffffffu : 0u;
            if (!uniform) {
                obits = 0u;
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < NB; ++b) obits |= (uint32_t)(gi[a] < gj[b]) << (a * NB + b);
                om = (obits & 1u) ? 0xffffffffu : 0u;
            }
            if (LUT) {
                const uint32_t a0 = lut_s + (om & (uint32_t)(SZA * 4));
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < NB; ++b) acc[a][b] = a0;
                if (!uniform) {
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < NB; ++b) acc[a][b] = lut_s + ((obits >> (a * NB + b)) & 1u) * (uint32_t)(SZA * 4);
                }
            } else {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < NB; ++b) acc[a][b] = 0u;
            }
        }
        const uint32_t* srow = stages + (size_t)slot * stage_words;
        const uint32_t* scol = srow + KW * op_words;
        for (int kk = 0; kk < nw; ++kk) {
            const bool boundary = (w0 + kk == p.WA);   // first word that is not a pure group-A word
            if (boundary && !p.mixed) {
                // group A finished: classify ic, restart the counters with the class carried along
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < NB; ++b) acc[a][b] = flushA(acc[a][b], a * NB + b);
            }
            const uint32_t* xr = srow + kk * op_words + ty * 4;
            const uint32_t* yc = scol + kk * op_words + tx * 4;
            const uint32_t* yc2 = yc + KW * op_words;
            uint32_t bor[4][NB];
            {
                const uint4 xv = *reinterpret_cast<const uint4*>(xr);
                const uint4 yv = *reinterpret_cast<const uint4*>(yc);
                const uint4 zv = *reinterpret_cast<const uint4*>(yc2);
                const uint32_t x[4] = {xv.x, xv.y, xv.z, xv.w};
                const uint32_t y[NB] = {yv.x, yv.y, yv.z, yv.w, zv.x, zv.y, zv.z, zv.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < NB; ++b) bor[a][b] = lop3_xor3(x[a], y[b], om);
            }
            if (!uniform) {   // rare: flip the coin seed of the pairs whose orientation differs from pair (0,0)
                uint32_t ob = obits;
                asm volatile("" : "+r"(ob));   // keep the mask arithmetic inside this branch
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < NB; ++b) bor[a][b] ^= (0u - (((ob >> (a * NB + b)) ^ ob) & 1u));
            }
            if (NPT > 0) {
#pragma unroll
                for (int pl = 1; pl < (NPT > 0 ? NPT : 1); ++pl) {
                    const uint4 xv = *reinterpret_cast<const uint4*>(xr + pl * REO_TILE);
                    const uint4 yv = *reinterpret_cast<const uint4*>(yc + pl * REO_TILE);
                    const uint4 zv = *reinterpret_cast<const uint4*>(yc2 + pl * REO_TILE);
                    const uint32_t x[4] = {xv.x, xv.y, xv.z, xv.w};
                    const uint32_t y[NB] = {yv.x, yv.y, yv.z, yv.w, zv.x, zv.y, zv.z, zv.w};
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < NB; ++b) bor[a][b] = lop3_b2(x[a], y[b], bor[a][b]);
                }
            } else {
#pragma unroll 2
                for (int pl = 1; pl < NP; ++pl) {
                    const uint4 xv = *reinterpret_cast<const uint4*>(xr + pl * REO_TILE);
                    const uint4 yv = *reinterpret_cast<const uint4*>(yc + pl * REO_TILE);
                    const uint4 zv = *reinterpret_cast<const uint4*>(yc2 + pl * REO_TILE);
                    const uint32_t x[4] = {xv.x, xv.y, xv.z, xv.w};
                    const uint32_t y[NB] = {yv.x, yv.y, yv.z, yv.w, zv.x, zv.y, zv.z, zv.w};
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < NB; ++b) bor[a][b] = lop3_b2(x[a], y[b], bor[a][b]);
                }
            }
            if (boundary && p.mixed) {
                // the word shared by the tails of both groups: count group A's slots, classify, then
                // count group B's slots (the masks select real samples only: no pad slots are counted)
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        const uint32_t fullA = mad_acc((uint32_t)__popc(bor[a][b] & p.maskA), four, acc[a][b]);
                        acc[a][b] = mad_acc((uint32_t)__popc(bor[a][b] & p.maskB), four, flushA(fullA, a * NB + b));
                    }
            } else {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < NB; ++b) acc[a][b] = mad_acc((uint32_t)__popc(bor[a][b]), four, acc[a][b]);
            }
        }
        if (flags & P2F_LAST_J) {
            // group B finished: look up the bin; the pair adds sign(j) to bin q of its row gene and, in the symmetric
            // region, sign(i) to the mirrored bin (8 - q, 0-based; src:385-386) of its column gene
            const uint32_t mirror = 8u * binstride;
            if (uniform) {
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const bool rowupd = flags & (b < 4 ? P2F_ROW0 : P2F_ROW1);
                    const bool colupd = flags & (b < 4 ? P2F_COL0 : P2F_COL1);
                    const int sgc = (int)(int8_t)((b < 4 ? csg0 : csg1) >> (8 * (b & 3)));
                    const uint32_t caddr = tabC_s + coloff + (uint32_t)((b >> 2) * REO_TILE * 4 + (b & 3) * 4) + mirror;
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const uint32_t off = binB(acc[a][b], a * NB + b);
                        if (rowupd) red_shared_add(tabR_s + rowoff + off + (uint32_t)(a * 4), sgc);
                        if (colupd) red_shared_add(caddr - off, (int)(int8_t)(rsg >> (8 * a)));
                    }
                }
            } else {   // pad genes, self pairs, mixed orientation: matrix edges and the diagonal only
                int gi[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) gi[a] = p.row_gene[(size_t)curI * REO_TILE + ty * 4 + a];
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const bool rowupd = flags & (b < 4 ? P2F_ROW0 : P2F_ROW1);
                    const bool colupd = flags & (b < 4 ? P2F_COL0 : P2F_COL1);
                    const int sgc = (int)(int8_t)((b < 4 ? csg0 : csg1) >> (8 * (b & 3)));
                    const uint32_t caddr = tabC_s + coloff + (uint32_t)((b >> 2) * REO_TILE * 4 + (b & 3) * 4) + mirror;
                    const int gjb = (rowupd || colupd) ? p.col_gene[(size_t)(curJ + (b >> 2)) * REO_TILE + tx * 4 + (b & 3)] : -1;
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const uint32_t off = binB(acc[a][b], a * NB + b);
                        if (gjb >= 0 && gi[a] >= 0 && gi[a] != gjb) {
                            if (rowupd) red_shared_add(tabR_s + rowoff + off + (uint32_t)(a * 4), sgc);
                            if (colupd) red_shared_add(caddr - off, (int)(int8_t)(rsg >> (8 * a)));
                        }
                    }
                }
            }
        }
        const bool last_item = flags & P2F_LAST_ITEM;
        const int rbase = A->rbase, cbase = A->cbase;
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);      // this warp is done with the slot (and its header)
        if (last_item) {
            p2_consumer_bar();
            for (int i = tid; i < 9 * tabw; i += P2_CONSUMERS) {
                const int bin = i / tabw, pos = i - bin * tabw;
                const int v = tabR[i];
                if (v != 0) {
                    const int g = p.row_gene[rbase + pos];
                    if (g >= 0) atomicAdd(&p.table[(size_t)g * 9 + bin], v);
                    tabR[i] = 0;
                }
                const int u = tabC[i];
                if (u != 0) {
                    const int g = p.col_gene[cbase + pos];
                    if (g >= 0) atomicAdd(&p.table[(size_t)g * 9 + bin], u);
                    tabC[i] = 0;
                }
            }
            p2_consumer_bar();
        }
        if (++slot == NS) { slot = 0; phase ^= 1u; }
    }
}

// ---- host side ----------------------------------------------------------------------------------
// block edge: one item should carry enough chain work to amortise its table flush and the tail of the launch
int reo_pairs2_block_edge(int W, int NP) {
    static const int forced = getenv("REO_P2_T") ? atoi(getenv("REO_P2_T")) : 0;
    if (forced == 2 || forced == 4 || forced == 8) return forced;
    const long long work = (long long)W * NP;     // LOP3 per pair
    if (work >= 512) return 2;
    if (work >= 64) return 4;
    return 8;
}

template <int NPT, bool LUT>
static cudaError_t launch_p2(const ReoPair2Params& p, size_t smem, int num_sms, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(reo_pair2_kernel<NPT, LUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int grid = P2_CTAS_PER_SM * num_sms;
    if (grid > p.nitems) grid = p.nitems;
    if (grid < 1) return cudaSuccess;
    reo_pair2_kernel<NPT, LUT><<<grid, P2_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

template <int NPT>
static cudaError_t launch_p2_np(const ReoPair2Params& p, size_t smem, int num_sms, cudaStream_t st) {
    return p.use_lut ? launch_p2<NPT, true>(p, smem, num_sms, st) : launch_p2<NPT, false>(p, smem, num_sms, st);
}

// Fills the geometry (blocks, supertiles, ring) from ntr / ntc / nsym / T / W / NP / rank / world and launches.
cudaError_t reo_launch_pairs2(ReoPair2Params p, int num_sms, cudaStream_t st) {
    if (p.ntr <= 0 || p.ntc <= 0) return cudaSuccess;
    const int T = p.T;
    p.NBs = (p.nsym + T - 1) / T;
    const int nsymp = p.NBs * T;
    p.NBr = p.nsym > 0 ? p.NBs + (std::max(p.ntr - nsymp, 0) + T - 1) / T : (p.ntr + T - 1) / T;
    p.NBc = (p.ntc + T - 1) / T;
    // supertile edge: ~24 tiles (both operand sets of the CTAs in flight stay in L2); smaller when several ranks share
    // the supertiles round-robin, and never so large that a rank is left with only a handful of them
    static const int forced_ss = getenv("REO_P2_SS") ? atoi(getenv("REO_P2_SS")) : 0;
    int SS = std::max(1, (p.world > 1 ? 12 : 24) / T);
    if (forced_ss > 0) SS = forced_ss;
    for (;;) {
        p.SS = SS;
        p.Ms = (p.NBs + SS - 1) / SS;
        p.Mc = (p.NBc + SS - 1) / SS;
        const int Mr = (p.NBr - p.NBs + SS - 1) / SS;
        p.tri = (long long)p.Ms * (p.Ms + 1) / 2;
        p.NSUP = p.tri + (long long)Mr * p.Mc;
        if (SS == 1 || forced_ss > 0 || p.NSUP >= 48LL * p.world) break;
        SS = SS > 2 ? SS / 2 : 1;
    }
    const long long mine = p.NSUP > p.rank ? (p.NSUP - p.rank + p.world - 1) / p.world : 0;
    p.nitems = (int)std::min<long long>(mine * p.SS * p.SS, 0x7fffffff);
    if (p.nitems <= 0) return cudaSuccess;
    // lookup tables
    const int sza = p.nA + p.padA + 1, szb = p.nB + p.padB + 1;
    p.lutSZA = sza; p.lutSZB = szb;
    p.use_lut = (2 * sza + 6 * szb) <= P2_LUT_MAX_WORDS;
    const int lut_words = p.use_lut ? 2 * sza + 6 * szb : 0;
    // ring: what is left of the budget after tables, headers and lookup tables
    const size_t fixed = (size_t)18 * T * REO_TILE * 4 + (size_t)lut_words * 4 + 2 * P2_MAX_NS * 8 + 64;
    const size_t perword = (size_t)3 * p.NP * REO_TILE * 4;
    const size_t avail = (size_t)P2_SMEM_BUDGET - fixed;
    static const int forced_kw = getenv("REO_P2_KW") ? atoi(getenv("REO_P2_KW")) : 0;
    static const int forced_ns = getenv("REO_P2_NS") ? atoi(getenv("REO_P2_NS")) : 0;
    int KW = (int)std::min<size_t>(3, avail / (4 * (perword + 256)));
    if (forced_kw > 0) KW = forced_kw;
    KW = std::max(1, std::min(KW, p.W));
    {   // spread the words evenly over the steps of one tile pair
        const int nsteps = (p.W + KW - 1) / KW;
        KW = (p.W + nsteps - 1) / nsteps;
    }
    int NS = (int)(avail / ((size_t)KW * perword + sizeof(P2Aux)));
    if (forced_ns > 0) NS = std::min(NS, forced_ns);
    NS = std::min(NS, P2_MAX_NS);
    if (NS < 2) return cudaErrorInvalidConfiguration;
    p.KW = KW; p.NS = NS;
    p.one = 1u;
    const size_t smem = (size_t)NS * KW * perword + (size_t)NS * sizeof(P2Aux) + fixed;
    switch (p.NP) {
#define CASE_NP(n) case n: return launch_p2_np<n>(p, smem, num_sms, st);
        CASE_NP(2) CASE_NP(3) CASE_NP(4) CASE_NP(5) CASE_NP(6) CASE_NP(7) CASE_NP(8) CASE_NP(9) CASE_NP(10)
        CASE_NP(11) CASE_NP(12) CASE_NP(13) CASE_NP(14) CASE_NP(15) CASE_NP(16) CASE_NP(17)
#undef CASE_NP
        default: return launch_p2_np<0>(p, smem, num_sms, st);
    }
}
