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
//    and the items of every supertile are dealt round-robin to the ranks of a multi-GPU job (tables are summed
//    across ranks);
//  * the tie-coin orientation [i<j] of a thread's 4 x 8 pairs is one register in all but a few blocks; a WARP votes
//    for one of three variants of the word loop (one orientation / per column via PRMT / per pair), so that the
//    blocks the sorted column list crosses -- one column group of every tile pair of a small update -- cost 2 %
//    instead of a second pass through the loop (profiles/r02_pair_kernel_trace.md).
// Arithmetic is unchanged: bit-sliced borrow chain, ONE LOP3 (0xB2) per rank plane per 32 samples, tie coin as
// plane 0, POPC + IMAD accumulation, lookup-table classification.  Bound: the ALU (LOP3) pipe; no tensor cores.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

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
    int flags, nw, kb, w0;       // first 16 bytes: what every stage needs (kb: index in the stage of the first word
                                 // that is not purely of group A, -1 if none)
    int I, J, il, jl, rbase, cbase;
    int skip;                    // bit k: word k of the stage has an empty top plane (its chain stops one plane early)
    int pad[5];
};
static_assert(sizeof(P2Aux) == 1024, "P2Aux must stay 1 KB");

__device__ __forceinline__ void p2_consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(P2_CONSUMERS) : "memory"); }

// work item n of this rank -> block (bi, bj) and row part `sub` of the block; false: nothing to do for this index.
// Items are numbered supertile by supertile (L2 locality of the CTAs in flight) and dealt round-robin to the ranks:
// every rank gets the same number of items of every supertile (+-1), whatever the shape of the tile space.
__host__ __device__ __forceinline__ bool p2_decode(const ReoPair2Params& p, int n, int& bi, int& bj, int& sub) {
    const long long g = (long long)n * p.world + p.rank;
    const int per_sup = p.SS * p.SSc * p.RS;
    const long long s = g / per_sup;
    if (s >= p.NSUP) return false;
    // inside a supertile the items are visited in a scrambled order (multiplication by a unit modulo their number):
    // the valid items of a diagonal supertile then spread evenly over the ranks
    const int li0 = (int)(((g - s * per_sup) * p.perm_mul + s) % per_sup);
    const int li = li0 / p.RS;
    sub = li0 - li * p.RS;
    const int di = li / p.SSc, dj = li - di * p.SSc;
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
    bj = (int)(s2 % p.Mc) * p.SSc + dj;
    return bi < p.NBr && bj < p.NBc;
}

// Tile pairs (64 rows x 128 columns) of work item n, in processing order: f(I, B0, Ja, J0, hasB, flags, last) with
// I = row tile, B0 / J0 = first row / column tile of the item's block, Ja = first column tile of the pair, hasB = the
// second column tile exists, flags = P2F_ROW0 | COL0 | ROW1 | COL1 (which genes the pairs update), last = last pair of
// the item.  Used by the kernel's producer warp AND by reo_pairs2_plan (host, tests): one enumeration, one truth.
template <class F>
__host__ __device__ __forceinline__ void p2_for_each_tile_pair(const ReoPair2Params& p, int n, F&& f) {
    int bi, bj, sub;
    if (!p2_decode(p, n, bi, bj, sub)) return;
    const int T = p.T;
    const bool symrow = bi < p.NBs;
    const int TI = T / p.RS;                       // row tiles of one item
    const int B0 = bi * T, J0 = bj * T;            // first row / column tile of the block
    const int I0 = B0 + sub * TI;
    const int Ilim = symrow ? p.nsym : p.ntr;
    const int Iend = I0 + TI < Ilim ? I0 + TI : Ilim;
    const int Jend = J0 + T < p.ntc ? J0 + T : p.ntc;
    const int jphi = (Jend + 1) >> 1;
    int left = 0;                                  // tile pairs of this item
    for (int I = I0; I < Iend; ++I) {
        int jplo = J0 >> 1;
        if (symrow && (I >> 1) > jplo) jplo = I >> 1;
        if (jphi > jplo) left += jphi - jplo;
    }
    for (int I = I0; I < Iend; ++I) {
        int jplo = J0 >> 1;
        if (symrow && (I >> 1) > jplo) jplo = I >> 1;
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
            f(I, B0, Ja, J0, hasB, tf, left == 0);
        }
    }
}

// NPT > 0: planes known at compile time (fully unrolled chain); NPT == 0: runtime p.NP.
// MODE: how a pair's counts are kept and classified --
//   P2_WIDE   32-bit counters, compare-based classification (any number of samples);
//   P2_LUT    classification through shared-memory lookup tables, accumulators ARE table addresses (few samples:
//             the per-tile epilogue costs loads, not ALU instructions);
//   P2_PACKED 16-bit counters, two per register, compare-based classification (up to 65535 sample slots per group).
#define P2_WIDE 0
#define P2_LUT 1
#define P2_PACKED 2
template <int NPT, int MODE>
__global__ void __launch_bounds__(P2_THREADS, P2_CTAS_PER_SM) reo_pair2_kernel(const ReoPair2Params p) {
    constexpr int NB = 8;
    constexpr bool LUT = MODE == P2_LUT, PACKED = MODE == P2_PACKED;
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
        for (;;) {
            const int n = (int)atomicAdd(p.counter, 1u);
            if (n >= p.nitems) break;
            p2_for_each_tile_pair(p, n, [&](int I, int B0, int Ja, int J0, bool hasB, int tf, bool last_pair) {
                    const uint32_t* rbase = p.row_planes + (size_t)I * tile_stride;
                    const uint32_t* cbase = p.col_planes + (size_t)Ja * tile_stride;
                    for (int ch = 0; ch < nchunks; ++ch) {
                        mbar_wait_parked(&empty[slot], ephase);
                        P2Aux* A = &aux[slot];
                        const int w0 = ch * KW;
                        const int nw = min(KW, p.W - w0);
                        const bool first = ch == 0, last = ch == nchunks - 1;
                        A->I = I; A->J = Ja; A->w0 = w0; A->nw = nw;
                        A->kb = (p.WA >= w0 && p.WA < w0 + nw) ? p.WA - w0 : -1;
                        A->il = I - B0; A->jl = Ja - J0; A->rbase = B0 * REO_TILE; A->cbase = J0 * REO_TILE;
                        A->flags = tf | (first ? P2F_FIRST_J : 0) | (last ? P2F_LAST_J : 0) |
                                   ((last && last_pair) ? P2F_LAST_ITEM : 0);
                        {   // words of this stage whose top plane is empty (header fields are published by the arrive below)
                            int skip = 0;
                            if (p.word_np)
                                for (int q = 0; q < nw; ++q) skip |= (int)(p.word_np[word_of(w0 + q)] < NP) << q;
                            A->skip = skip;
                        }
                        uint32_t bytes = (uint32_t)nw * op_bytes * (hasB ? 3u : 2u);
                        // gene ids travel with the first stage of a tile pair (orientation of the tie coin) and, together
                        // with the signs, with its last stage (table update): consumers keep none of them in registers
                        if (first || last) bytes += 256u + 512u;
                        if (last) bytes += (p.col_sign ? 128u : 0u) + (p.row_sign ? 64u : 0u);
                        mbar_expect_tx(&full[slot], bytes);
                        if (first || last) {
                            bulk_g2s(A->rgene, p.row_gene + (size_t)I * REO_TILE, 256u, &full[slot]);
                            bulk_g2s(A->cgene, p.col_gene + (size_t)Ja * REO_TILE, 512u, &full[slot]);
                        }
                        if (last) {
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
            });
        }
        // no more work: one terminating stage
        mbar_wait_parked(&empty[slot], ephase);
        aux[slot].flags = P2F_TERM;
        mbar_arrive(&full[slot]);
        return;
    }

    // =============================== consumer warps ===============================
    // warp = 8 x 4 threads: 32 rows x (16 + 16) columns; thread = rows ty*4..+3, columns tx*4..+3 of both column tiles
    const int ty = (warp >> 2) * 8 + (lane >> 2);
    const int tx = (warp & 3) * 4 + (lane & 3);
    // Accumulators.  WIDE: 4 * count, class of group A carried in bits 28..31.  LUT: shared-memory addresses (see
    // above).  PACKED: two 16-bit counts per register (column b in the low half, column b + 4 in the high half) and
    // the classes of group A in two bit fields -- the 16 registers this frees hold the next plane's operands.
    constexpr int NACC = PACKED ? 4 : NB;
    uint32_t acc[4][NACC];
    uint32_t cls_lo = 0u, cls_hi = 0u;   // PACKED: 2 bits per pair (a*4 + b), columns 0..3 / 4..7
    uint32_t om = 0u;          // tie-coin orientation [i<j] of this thread's pairs (all-ones / zero) ...
    uint32_t obits = 0u;       // ... and per pair (bit a*NB+b), used only when the 4 x 8 block is not uniform
    bool uniform = true;       // one orientation, every gene real, no self pair
    int wpath = 0;             // word loop of this warp: 0 one orientation, 1 per column, 2 per pair
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < NACC; ++b) acc[a][b] = 0u;
    const uint32_t kmul = PACKED ? p.one : (p.one << 2);      // IMAD multipliers (kernel parameters: stay on the FMA pipe)
    const uint32_t kmul_hi = PACKED ? (p.one << 16) : kmul;
    const uint32_t lut_s = smem_u32(lut);
    const uint32_t tabR_s = smem_u32(tabR), tabC_s = smem_u32(tabC);

    auto count_of = [&](int a, int b) -> int {   // PACKED / WIDE: plain count of pair (a, b)
        if (PACKED) return (int)(b < 4 ? (acc[a][b & 3] & 0xffffu) : (acc[a][b & 3] >> 16));
        return (int)((acc[a][b % NACC] & 0x0fffffffu) >> 2);
    };
    // group A finished for every pair: classify ic, restart the counters with the class carried along
    auto flushA_all = [&]() {
        if (LUT) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < NACC; ++b) acc[a][b] = lds_u32(acc[a][b]);
        } else {
            cls_lo = 0u; cls_hi = 0u;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const uint32_t o = (obits >> (a * NB + b)) & 1u;
                    const uint32_t cl = reo_class(count_of(a, b) - (int)(o * (uint32_t)p.padA), p.nA, p.thrA);
                    if (PACKED) { if (b < 4) cls_lo |= cl << (2 * (a * 4 + b)); else cls_hi |= cl << (2 * (a * 4 + (b & 3))); }
                    else acc[a][b % NACC] = cl << 28;
                }
            if (PACKED) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < NACC; ++b) acc[a][b] = 0u;
            }
        }
    };
    // (class of A, count of B) -> byte offset of the table bin row
    auto binB = [&](int a, int b) -> uint32_t {
        if (LUT) return lds_u32(acc[a][b % NACC]);
        const uint32_t o = (obits >> (a * NB + b)) & 1u;
        const uint32_t it = reo_class(count_of(a, b) - (int)(o * (uint32_t)p.padB), p.nB, p.thrB);
        uint32_t ic;
        if (PACKED) ic = ((b < 4 ? cls_lo : cls_hi) >> (2 * (a * 4 + (b & 3)))) & 3u;
        else ic = acc[a][b % NACC] >> 28;
        return (3u * ic + it) * binstride;
    };
    auto add_counts = [&](const uint32_t (&bor)[4][NB], uint32_t mask, bool masked) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const uint32_t pc = (uint32_t)__popc(masked ? (bor[a][b] & mask) : bor[a][b]);
                acc[a][b % NACC] = mad_acc(pc, (PACKED && b >= 4) ? kmul_hi : kmul, acc[a][b % NACC]);
            }
    };

#ifdef REO_P2_TRACE
    // -DREO_P2_TRACE: every consumer warp prints where its time went (one line per warp of a CTA that did work):
    // hardware warp slot, stages, items, time spent waiting for operands, time in the kernel, word-loop variant
    unsigned long long tr_t0, tr_first = 0ull, tr_items = 0ull, tr_wait = 0ull; int tr_nst = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_t0));
#endif
    int slot = 0;
    uint32_t phase = 0u;
    for (;;) {
#ifdef REO_P2_TRACE
        unsigned long long tr_w0, tr_w1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_w0));
#endif
        mbar_wait_parked(&full[slot], phase);
#ifdef REO_P2_TRACE
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_w1));
        if (tr_first == 0ull) tr_first = tr_w1;
        else tr_wait += tr_w1 - tr_w0;
        ++tr_nst;
#endif
        const P2Aux* A = &aux[slot];
        const int4 hdr = *reinterpret_cast<const int4*>(&A->flags);
        const int flags = hdr.x;
        if (flags & P2F_TERM) break;
        const int nw = hdr.y, kb = hdr.z;
        if (flags & P2F_FIRST_J) {
            const int4 r4 = *reinterpret_cast<const int4*>(A->rgene + ty * 4);
            const int4 c4 = *reinterpret_cast<const int4*>(A->cgene + tx * 4);
            const int4 d4 = *reinterpret_cast<const int4*>(A->cgene + REO_TILE + tx * 4);
            const int gi[4] = {r4.x, r4.y, r4.z, r4.w};
            const int gj[NB] = {c4.x, c4.y, c4.z, c4.w, d4.x, d4.y, d4.z, d4.w};
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
            // the word loop is chosen per warp (no divergence): one orientation for every lane's block; per column;
            // per pair
            {
                const bool colok = obits == (obits & 0xffu) * 0x01010101u;
                wpath = __all_sync(0xffffffffu, uniform) ? 0 : (__all_sync(0xffffffffu, colok) ? 1 : 2);
            }
            if (LUT) {
                const uint32_t a0 = lut_s + (om & (uint32_t)(SZA * 4));
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < NACC; ++b) acc[a][b] = a0;
                if (!uniform) {
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < NACC; ++b) acc[a][b] = lut_s + ((obits >> (a * NB + b)) & 1u) * (uint32_t)(SZA * 4);
                }
            } else {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < NACC; ++b) acc[a][b] = 0u;
            }
        }
        // operands of the stage: rows, first column tile, second column tile; this thread's 16-byte chunks.  The
        // loads run one plane ahead of the LOP3 chain (software pipeline: shared-memory latency never shows).
        uint32_t xa = smem_u32(stages + (size_t)slot * stage_words) + (uint32_t)ty * 16u;
        uint32_t ya = smem_u32(stages + (size_t)slot * stage_words + KW * op_words) + (uint32_t)tx * 16u;
        const uint32_t zoff = (uint32_t)(KW * op_words) * 4u;
        // word of this stage that is the first not purely of group A (the class of group A is taken there); the
        // plain words before and after it run in a loop without any test
        const bool has_b = kb >= 0 && kb < nw;
        const uint32_t skipmask = (uint32_t)A->skip;
        // UNI: one tie-coin orientation for the whole 4 x 8 block (all but the blocks on the diagonal / matrix edges)
        auto words = [&](auto uni_c) {
            constexpr bool UNI = decltype(uni_c)::value;
            // column flags of the orientation in the top bit of each byte (columns 0..3 / 4..7), for PRMT
            const uint32_t cfA = ((obits & 1u) << 7) | ((obits & 2u) << 14) | ((obits & 4u) << 21) | ((obits & 8u) << 28);
            const uint32_t cfB = ((obits & 16u) << 3) | ((obits & 32u) << 10) | ((obits & 64u) << 17) | ((obits & 128u) << 24);
            uint4 xn = lds_v4(xa), yn = lds_v4(ya), zn = lds_v4(ya + zoff);
            // borrow chain of one word over all planes; leaves the operands of the next word's plane 0 in xn/yn/zn
            // (npl_c: planes to run through -- NPT, or NPT - 1 for a word whose top plane is empty)
            auto chain = [&](uint32_t (&bor)[4][NB], auto npl_c) {
                constexpr int NPL = decltype(npl_c)::value;
                {
                    const uint32_t x[4] = {xn.x, xn.y, xn.z, xn.w};
                    const uint32_t y[NB] = {yn.x, yn.y, yn.z, yn.w, zn.x, zn.y, zn.z, zn.w};
                    xn = lds_v4(xa + 256u); yn = lds_v4(ya + 256u); zn = lds_v4(ya + zoff + 256u);   // plane 1
                    if (UNI) {
#pragma unroll
                        for (int b = 0; b < NB; ++b)
#pragma unroll
                            for (int a = 0; a < 4; ++a) bor[a][b] = lop3_xor3(x[a], y[b], om);
                    } else if (wpath == 1) {
                        // the orientation depends on the column only (a block the column list's order crosses):
                        // one PRMT per column replicates its flag into a full word
#pragma unroll
                        for (int b = 0; b < NB; ++b) {
                            const uint32_t d = prmt_sign(b < 4 ? cfA : cfB, 0x8888u + 0x1111u * (uint32_t)(b & 3));
#pragma unroll
                            for (int a = 0; a < 4; ++a) bor[a][b] = lop3_xor3(x[a], y[b], d);
                        }
                    } else {
                        // any orientation per pair (diagonal blocks, blocks that hold a pad or the row gene itself)
#pragma unroll
                        for (int b = 0; b < NB; ++b)
#pragma unroll
                            for (int a = 0; a < 4; ++a) {
                                const uint32_t d = (uint32_t)((int)(obits << (31 - (a * NB + b))) >> 31);
                                bor[a][b] = lop3_xor3(x[a], y[b], d);
                            }
                    }
                }
                if (NPT > 0) {
#pragma unroll
                    for (int pl = 1; pl < (NPL > 0 ? NPL : 1); ++pl) {
                        const uint32_t x[4] = {xn.x, xn.y, xn.z, xn.w};
                        const uint32_t y[NB] = {yn.x, yn.y, yn.z, yn.w, zn.x, zn.y, zn.z, zn.w};
                        // next plane, or plane 0 of the next word (one word = op_bytes further; past the last word of
                        // the stage this reads shared memory that is simply not used)
                        const uint32_t nx = (pl + 1 < NPL) ? (uint32_t)(pl + 1) * 256u : op_bytes;
                        xn = lds_v4(xa + nx); yn = lds_v4(ya + nx); zn = lds_v4(ya + zoff + nx);
#pragma unroll
                        for (int b = 0; b < NB; ++b)
#pragma unroll
                            for (int a = 0; a < 4; ++a) bor[a][b] = lop3_b2(x[a], y[b], bor[a][b]);
                    }
                } else {
#pragma unroll 2
                    for (int pl = 1; pl < NP; ++pl) {
                        const uint32_t x[4] = {xn.x, xn.y, xn.z, xn.w};
                        const uint32_t y[NB] = {yn.x, yn.y, yn.z, yn.w, zn.x, zn.y, zn.z, zn.w};
                        const uint32_t nx = (pl + 1 < NP) ? (uint32_t)(pl + 1) * 256u : op_bytes;
                        xn = lds_v4(xa + nx); yn = lds_v4(ya + nx); zn = lds_v4(ya + zoff + nx);
#pragma unroll
                        for (int b = 0; b < NB; ++b)
#pragma unroll
                            for (int a = 0; a < 4; ++a) bor[a][b] = lop3_b2(x[a], y[b], bor[a][b]);
                    }
                }
                xa += op_bytes; ya += op_bytes;
            };
            constexpr std::integral_constant<int, NPT> full_c{};
            int done = 0;
            for (;;) {
                const int end = (has_b && done <= kb) ? kb : nw;
                for (; done < end; ++done) {          // plain words
                    uint32_t bor[4][NB];
                    // an all-zero plane leaves the borrow as it is: a word whose samples need one rank bit less than
                    // the widest sample of the matrix stops one plane early (most words of single-cell data)
                    if (UNI && NPT > 2 && ((skipmask >> done) & 1u)) chain(bor, std::integral_constant<int, (NPT > 2 ? NPT - 1 : NPT)>{});
                    else chain(bor, full_c);
                    add_counts(bor, 0u, false);
                }
                if (done >= nw) break;
                // the boundary word
                uint32_t bor[4][NB];
                if (!p.mixed) {
                    flushA_all();
                    chain(bor, full_c);
                    add_counts(bor, 0u, false);
                } else {
                    // the word shared by the tails of both groups: count group A's slots, classify, then
                    // count group B's slots (the masks select real samples only: no pad slots are counted)
                    chain(bor, full_c);
                    add_counts(bor, p.maskA, true);
                    flushA_all();
                    add_counts(bor, p.maskB, true);
                }
                ++done;
            }
        };
        if (wpath == 0) words(std::true_type{}); else words(std::false_type{});
        if (flags & P2F_LAST_J) {
            // group B finished: look up the bin; the pair adds sign(j) to bin q of its row gene and, in the symmetric
            // region, sign(i) to the mirrored bin (8 - q, 0-based; src:385-386) of its column gene.  Signs, gene ids
            // and the tile's place in the block tables come with this (the tile pair's last) stage.
            const uint32_t mirror = 8u * binstride;
            uint32_t csg0 = 0x01010101u, csg1 = 0x01010101u, rsg = 0x01010101u;   // packed int8 signs
            if (p.col_sign) {
                csg0 = *reinterpret_cast<const uint32_t*>(A->csgn + tx * 4);
                csg1 = *reinterpret_cast<const uint32_t*>(A->csgn + REO_TILE + tx * 4);
            }
            if (p.row_sign) rsg = *reinterpret_cast<const uint32_t*>(A->rsgn + ty * 4);
            const uint32_t rowoff = (uint32_t)(A->il * REO_TILE + ty * 4) * 4u;
            const uint32_t coloff = (uint32_t)(A->jl * REO_TILE + tx * 4) * 4u;
            if (uniform) {
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const bool rowupd = flags & (b < 4 ? P2F_ROW0 : P2F_ROW1);
                    const bool colupd = flags & (b < 4 ? P2F_COL0 : P2F_COL1);
                    const int sgc = (int)(int8_t)((b < 4 ? csg0 : csg1) >> (8 * (b & 3)));
                    const uint32_t caddr = tabC_s + coloff + (uint32_t)((b >> 2) * REO_TILE * 4 + (b & 3) * 4) + mirror;
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const uint32_t off = binB(a, b);
                        if (rowupd) red_shared_add(tabR_s + rowoff + off + (uint32_t)(a * 4), sgc);
                        if (colupd) red_shared_add(caddr - off, (int)(int8_t)(rsg >> (8 * a)));
                    }
                }
            } else {   // pad genes, self pairs, mixed orientation: matrix edges and the diagonal only
                const int4 r4 = *reinterpret_cast<const int4*>(A->rgene + ty * 4);
                const int gi[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const bool rowupd = flags & (b < 4 ? P2F_ROW0 : P2F_ROW1);
                    const bool colupd = flags & (b < 4 ? P2F_COL0 : P2F_COL1);
                    const int sgc = (int)(int8_t)((b < 4 ? csg0 : csg1) >> (8 * (b & 3)));
                    const uint32_t caddr = tabC_s + coloff + (uint32_t)((b >> 2) * REO_TILE * 4 + (b & 3) * 4) + mirror;
                    const int gjb = A->cgene[(b >> 2) * REO_TILE + tx * 4 + (b & 3)];
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const uint32_t off = binB(a, b);
                        if (gjb >= 0 && gi[a] >= 0 && gi[a] != gjb) {
                            if (rowupd) red_shared_add(tabR_s + rowoff + off + (uint32_t)(a * 4), sgc);
                            if (colupd) red_shared_add(caddr - off, (int)(int8_t)(rsg >> (8 * a)));
                        }
                    }
                }
            }
        }
        const bool last_item = flags & P2F_LAST_ITEM;
        int rbase = 0, cbase = 0;
        if (last_item) { rbase = A->rbase; cbase = A->cbase; }
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
#ifdef REO_P2_TRACE
        if (last_item) ++tr_items;
#endif
        if (++slot == NS) { slot = 0; phase ^= 1u; }
    }
#ifdef REO_P2_TRACE
    if (lane == 0 && tr_nst > 1) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        unsigned smid, wid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        printf("[p2 trace] cta %d warp %d sm %u slot %u first %.1f stages %d items %llu wait %.1f total %.1f path %d\n", blockIdx.x,
               warp, smid, wid, (double)(tr_first - tr_t0) * 1e-3, tr_nst, tr_items, (double)tr_wait * 1e-3,
               (double)(t1 - tr_t0) * 1e-3, wpath);
    }
#endif
}

// ---- host side ----------------------------------------------------------------------------------
// block edge: one item should carry enough chain work to amortise its table flush and the tail of the launch
int reo_pairs2_block_edge(int W, int NP) {
    static const int forced = getenv("REO_P2_T") ? atoi(getenv("REO_P2_T")) : 0;
    if (forced == 2 || forced == 4 || forced == 8) return forced;
    const long long work = (long long)W * NP;     // LOP3 per pair
    // measured on the 20k x 200 bulk workload (W = 7, 13 planes): T = 2 -> 2.03 ms, 4 -> 2.28 ms, 8 -> 2.94 ms of pair
    // kernels: small items balance the tail of a launch better than large ones save table flushes
    return work >= 32 ? 2 : 4;
}

template <int NPT, int MODE>
static cudaError_t launch_p2(const ReoPair2Params& p, size_t smem, int num_sms, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(reo_pair2_kernel<NPT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int grid = P2_CTAS_PER_SM * num_sms;
    if (grid > p.nitems) grid = p.nitems;
    if (grid < 1) return cudaSuccess;
    reo_pair2_kernel<NPT, MODE><<<grid, P2_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

template <int NPT>
static cudaError_t launch_p2_np(const ReoPair2Params& p, size_t smem, int num_sms, cudaStream_t st) {
    if (p.use_lut) return launch_p2<NPT, P2_LUT>(p, smem, num_sms, st);
    static const bool no_packed = getenv("REO_P2_NO_PACKED") != nullptr;
    if (!no_packed && p.nA + p.padA <= 65535 && p.nB + p.padB <= 65535) return launch_p2<NPT, P2_PACKED>(p, smem, num_sms, st);
    return launch_p2<NPT, P2_WIDE>(p, smem, num_sms, st);
}

// Geometry of a launch (blocks, supertiles, items of this rank) from ntr / ntc / nsym / T / W / NP / rank / world.
static void p2_geometry(ReoPair2Params& p) {
    const int T = p.T;
    p.NBs = (p.nsym + T - 1) / T;
    const int nsymp = p.NBs * T;
    p.NBr = p.nsym > 0 ? p.NBs + (std::max(p.ntr - nsymp, 0) + T - 1) / T : (p.ntr + T - 1) / T;
    p.NBc = (p.ntc + T - 1) / T;
    // supertile edge: ~24 tiles (both operand sets of the CTAs in flight stay in L2)
    static const int forced_ss = getenv("REO_P2_SS") ? atoi(getenv("REO_P2_SS")) : 0;
    int SS = std::max(1, 24 / T);
    if (forced_ss > 0) SS = forced_ss;
    for (;;) {
        p.SS = SS;
        // no symmetric region: a supertile is not wider than the column blocks there are (few columns: no item index
        // without work, so a short launch gets exactly one CTA per item)
        p.SSc = p.nsym > 0 ? SS : std::min(SS, p.NBc);
        p.Ms = (p.NBs + SS - 1) / SS;
        p.Mc = (p.NBc + p.SSc - 1) / p.SSc;
        const int Mr = (p.NBr - p.NBs + SS - 1) / SS;
        p.tri = (long long)p.Ms * (p.Ms + 1) / 2;
        p.NSUP = p.tri + (long long)Mr * p.Mc;
        if (SS == 1 || forced_ss > 0 || p.NSUP >= 8) break;
        SS = SS > 2 ? SS / 2 : 1;
    }
    // long tile pairs (thousands of sample words): one row tile per item, so that the tail of a launch stays short
    static const int forced_rs = getenv("REO_P2_RS") ? atoi(getenv("REO_P2_RS")) : 0;
    p.RS = ((long long)p.W * p.NP >= 1024) ? T : 1;
    if (forced_rs > 0 && T % forced_rs == 0) p.RS = forced_rs;
    {
        const int per_sup = p.SS * p.SSc * p.RS;
        int m = 7;
        while (std::__gcd(m, per_sup) != 1) m += 2;
        p.perm_mul = per_sup > 1 ? m % per_sup : 1;
        if (p.perm_mul == 0) p.perm_mul = 1;
    }
    const long long total_items = p.NSUP * p.SS * p.SSc * p.RS;
    const long long mine = total_items > p.rank ? (total_items - p.rank + p.world - 1) / p.world : 0;
    p.nitems = (int)std::min<long long>(mine, 0x7fffffff);
}

// Host-side replay of what the producer warps of rank p.rank will do: every (row tile, column tile, update flags) this
// rank evaluates, through the SAME p2_decode / p2_for_each_tile_pair as the kernel.  out: 3 ints per half tile pair.
// For the CPU tests of the multi-rank partition (no GPU needed).
long long reo_pairs2_plan(ReoPair2Params p, int32_t* out, long long cap) {
    if (p.ntr <= 0 || p.ntc <= 0) return 0;
    p2_geometry(p);
    long long n_out = 0;
    for (int n = 0; n < p.nitems; ++n)
        p2_for_each_tile_pair(p, n, [&](int I, int, int Ja, int, bool hasB, int tf, bool) {
            for (int half = 0; half < (hasB ? 2 : 1); ++half) {
                const int rowf = half ? P2F_ROW1 : P2F_ROW0, colf = half ? P2F_COL1 : P2F_COL0;
                if (!(tf & (rowf | colf))) continue;
                if (out && n_out < cap) { out[3 * n_out] = I; out[3 * n_out + 1] = Ja + half; out[3 * n_out + 2] = ((tf & rowf) ? 1 : 0) | ((tf & colf) ? 2 : 0); }
                ++n_out;
            }
        });
    return n_out;
}

// Fills the geometry and the ring, and launches.
cudaError_t reo_launch_pairs2(ReoPair2Params p, int num_sms, cudaStream_t st) {
    if (p.ntr <= 0 || p.ntc <= 0) return cudaSuccess;
    p2_geometry(p);
    if (p.nitems <= 0) return cudaSuccess;
    const int T = p.T;
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
    // long stages amortise the hand-over between stages (measured: 2 stages of 8 words beat 4 of 4)
    int KW = (int)std::min<size_t>(8, (avail - 2 * sizeof(P2Aux)) / (2 * perword));
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
