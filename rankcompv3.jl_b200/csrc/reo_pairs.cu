// reo_pairs.cu -- K2, FIRST GENERATION: the pair-count / stable-REO class / per-gene 9-bin table kernel of round 1.
// Since round 2 the rank path runs reo_pairs2.cu (warp-specialised, symmetric sweep); this kernel is kept for the
// raw-FP64 variant (FLT: non-integral expression values), for the small-block parity kernel at the end of
// the file, and -- with run-time plane count only -- for A/B runs of the rank path (REO_PAIRS_V1=1).
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
// feeds the per-thread integer accumulator through an IMAD (FMA pipe), so the ALU pipe carries
// almost nothing but the LOP3 chains: (B+1) LOP3 + 1 POPC + 1 IMAD per 32 ordered (i, j, sample)
// triples.  The ALU (LOP3) pipe is the roofline; no tensor cores.
//
// Tiling: CTA = 64 row genes x 64 column genes, 256 threads, 4x4 pairs per thread (16 independent
// borrow chains per thread hide the 4-cycle ALU latency); -DPK_NB=8 builds the 4x8 variant (two column
// tiles per step).  Operand tiles ([plane][64 genes] words) are streamed by cp.async.bulk (UBLKCP) into
// a 2-slot shared-memory ring (up to 8 sample words per slot) guarded by mbarriers with transaction
// counts; the warp that is LAST to finish a slot refills it, so no warp ever waits on an "empty"
// barrier and there is no CTA-wide barrier in the loop.  Each thread reads its 4 row words and 4 column
// words per plane with two LDS.128 (bank-conflict free).
// Classification: accumulators count in units of 4, i.e. they ARE byte offsets into small
// shared-memory lookup tables that turn (count of group A) -> class offset and
// (class of A, count of group B) -> table bin, so the per-pair epilogue costs loads, not ALU ops
// (compare-based fallback when the tables would not fit, i.e. thousands of samples per group,
// where the epilogue is negligible anyway).
// Work items (row tile x chunk of column tiles) are handed out by an atomic counter to a
// persistent grid of 3 CTAs per SM; each item accumulates a 9 x 64 table in shared memory and
// flushes it with integer atomics (order independent, exact).
#include "reo_internal.cuh"
#include "reo_ptx.cuh"

#define PK_THREADS 256
#define PK_WARPS (PK_THREADS / 32)
#ifndef PK_NS
#define PK_NS 2              // ring slots
#endif
#ifndef PK_CTAS_PER_SM
#define PK_CTAS_PER_SM 3    // resident CTAs per SM (<= 85 registers per thread)
#endif
#ifndef PK_MAX_KW
#define PK_MAX_KW 8          // sample words per slot (upper bound)
#endif
#ifndef PK_SMEM_BUDGET
#define PK_SMEM_BUDGET (70 * 1024)
#endif
#define PK_LUT_MAX_WORDS 4096   // lookup-table words (16 KB) above which the compare path is used
// Columns per thread: 4 (CTA tile 64 x 64, 4x4 pairs per thread) or 8 (CTA tile 64 x 128: two column tiles per step,
// 4x8 pairs per thread -- 12 operand words per 32 LOP3 instead of 8 per 16, at 2 CTAs per SM).
#ifndef PK_NB
#define PK_NB 4
#endif
#if PK_NB == 8
#define PK_CTAS_NB 2
#ifndef PK_SMEM_BUDGET_NB
#define PK_SMEM_BUDGET_NB (104 * 1024)
#endif
#else
#define PK_CTAS_NB PK_CTAS_PER_SM
#define PK_SMEM_BUDGET_NB PK_SMEM_BUDGET
#endif

#define PKF_FIRST_J 1
#define PKF_LAST_J 2
#define PKF_LAST_ITEM 4
#define PKF_TERM 8
#define PKF_SECOND 16     // the step carries a second column tile (J + 1)

struct PkMeta { int ti, J, w0, nw, flags, pad0, pad1, pad2; };
struct PkProd { int count, step, nsteps, ti, J, ch, done, jend; };   // producer state (shared memory)

__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(PK_THREADS) : "memory"); }

// NPT > 0: planes known at compile time (fully unrolled chain); NPT == 0: runtime p.NP.
// LUT: classification through the shared-memory lookup tables (accumulators are shared-memory addresses).
// FLT: non-integral expression values: the 0.1 tie band of is_greater (src:72) is not
//      transitive, so ranks cannot be used; an operand word is [64 coin words][32 samples][64 genes] FP64
//      and the 32 "is greater" bits of a word come from FP64 compares on the raw values instead of the
//      LOP3 chain.  Everything around it (ring, classification, tables) is shared with the rank path.
template <int NPT, bool LUT, bool FLT, int NB>
__global__ void __launch_bounds__(PK_THREADS, FLT ? 2 : (NB == 8 ? 2 : PK_CTAS_PER_SM)) reo_pair_kernel(const ReoPairParams p) {
    constexpr int NCT = NB / 4;                          // column tiles per step
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int NP = NPT > 0 ? NPT : p.NP;
    const int KW = p.KW;
    const int op_words = FLT ? REO_FLT_OPWORDS : NP * REO_TILE;   // words of one operand tile for one sample word
    const uint32_t op_bytes = (uint32_t)op_words * 4u;
    const int stage_words = (1 + NCT) * KW * op_words;  // rows then columns (then the second column tile)
    uint32_t* stages = reinterpret_cast<uint32_t*>(smem_raw);
    int32_t* tab_s = reinterpret_cast<int32_t*>(stages + (size_t)PK_NS * stage_words);      // [9][64]
    PkMeta* metas = reinterpret_cast<PkMeta*>(tab_s + REO_TILE * 9);                         // 32-byte entries
    uint64_t* full = reinterpret_cast<uint64_t*>(metas + PK_NS);
    PkProd* pst = reinterpret_cast<PkProd*>(full + PK_NS);
    int* slot_cnt = reinterpret_cast<int*>(pst + 1);   // warps done with each ring slot
    uint32_t* lut = reinterpret_cast<uint32_t*>(slot_cnt + PK_NS);
    // lut layout (words), o = tie-coin orientation of the pair (0: i>j, 1: i<j), SZA/SZB = slots + 1:
    //   lutA[o][v]         at o*SZA + v              -> shared-memory ADDRESS of lutB[o][ic*SZB + 0]
    //   lutB[o][ic*SZB+v]  at 2*SZA + o*3*SZB + ...  -> (3*ic + it) * 256 (byte offset of the bin row)
    // An accumulator starts as the address of lutA[o][0] and grows by 4 per counted sample, so
    // "classify group A" is one LDS [acc] and "find the bin" is one more.
    const int SZA = p.lutSZA, SZB = p.lutSZB;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
#ifndef PK_WARP_SHAPE
#define PK_WARP_SHAPE 0
#endif
#if PK_WARP_SHAPE == 0      // warp = 8 x 4 threads: 32 rows x 16 columns
    const int ty = (warp >> 2) * 8 + (lane >> 2);   // 0..15 : rows ty*4 .. ty*4+3
    const int tx = (warp & 3) * 4 + (lane & 3);     // 0..15 : cols tx*4 .. tx*4+3
#elif PK_WARP_SHAPE == 1    // warp = 4 x 8 threads: 16 rows x 32 columns
    const int ty = (warp >> 1) * 4 + (lane >> 3);
    const int tx = (warp & 1) * 8 + (lane & 7);
#elif PK_WARP_SHAPE == 2    // warp = 2 x 16 threads: 8 rows x 64 columns
    const int ty = warp * 2 + (lane >> 4);
    const int tx = lane & 15;
#else                       // warp = 16 x 2 threads: 64 rows x 8 columns
    const int ty = (warp >> 3) * 16 + (lane >> 1);
    const int tx = (warp & 7) * 2 + (lane & 1);
#endif

    if (tid == 0) {
        for (int s = 0; s < PK_NS; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < REO_TILE * 9; i += PK_THREADS) tab_s[i] = 0;
    if (LUT) {
        const uint32_t lutB0 = smem_u32(lut) + (uint32_t)(2 * SZA * 4);
        for (int i = tid; i < 2 * SZA; i += PK_THREADS) {
            const int o = i / SZA, v = i - o * SZA;
            lut[i] = lutB0 + (uint32_t)(o * 3 * SZB * 4) + reo_class(v - o * p.padA, p.nA, p.thrA) * (uint32_t)(SZB * 4);
        }
        for (int i = tid; i < 6 * SZB; i += PK_THREADS) {
            const int o = i / (3 * SZB), rem = i - o * 3 * SZB;
            const int ic = rem / SZB, v = rem - ic * SZB;
            lut[2 * SZA + i] = (3u * ic + reo_class(v - o * p.padB, p.nB, p.thrB)) * 256u;
        }
    }
    __syncthreads();

    const int nchunks = (p.W + KW - 1) / KW;
    const size_t word_stride = FLT ? (size_t)REO_FLT_OPWORDS : (size_t)p.NP * REO_TILE;
    const size_t tile_stride = (size_t)p.W * word_stride;
    const int nitems = (p.t1 - p.t0) * p.njchunks;

    // ---- producer: run by whichever warp is LAST to finish a ring slot (so nobody waits on an "empty"
    //      barrier); its state lives in shared memory.  k-th word of the two-group order -> staged word:
    auto word_of = [&](int k) -> int {
        if (k < p.WA) return p.segA0 + k;
        k -= p.WA;
        if (p.mixed) { if (k == 0) return p.mixedW; k -= 1; }
        return k < p.segB0len ? p.segB0 + k : p.segB1 + (k - p.segB0len);
    };
    auto produce_one = [&]() {
        if (pst->done) return;
        const int pr_count = pst->count;
        const int slot = pr_count % PK_NS;
        PkMeta& m = metas[slot];
        if (pst->step == pst->nsteps) {
            const int item = (int)atomicAdd(p.counter, 1u);
            if (item >= nitems) {
                m.flags = PKF_TERM;
                mbar_arrive(&full[slot]);
                pst->done = 1; pst->count = pr_count + 1;
                return;
            }
            const int it_row = item / p.njchunks;
            const int jbeg = (item - it_row * p.njchunks) * p.jchunk;
            const int jend = min(jbeg + p.jchunk, p.ntc);
            pst->ti = p.t0 + it_row;
            pst->J = jbeg; pst->ch = 0; pst->jend = jend;
            pst->nsteps = ((jend - jbeg + NCT - 1) / NCT) * nchunks;
            pst->step = 0;
        }
        const int st = pst->step, ti = pst->ti;
        const int J = pst->J, ch = pst->ch;      // (column tile, word chunk) advance without divisions
        if (ch + 1 == nchunks) { pst->ch = 0; pst->J = J + NCT; } else pst->ch = ch + 1;
        const bool second = (NCT == 2) && (J + 1 < pst->jend);
        const int w0 = ch * KW;
        const int nw = min(KW, p.W - w0);
        m.ti = ti; m.J = J; m.w0 = w0; m.nw = nw;
        m.flags = (ch == 0 ? PKF_FIRST_J : 0) | (ch == nchunks - 1 ? PKF_LAST_J : 0) |
                  (st == pst->nsteps - 1 ? PKF_LAST_ITEM : 0) | (second ? PKF_SECOND : 0);
        uint32_t* dst = stages + (size_t)slot * stage_words;
        mbar_expect_tx(&full[slot], (uint32_t)nw * (second ? 3u : 2u) * op_bytes);
        const uint32_t* rbase = p.row_planes + (size_t)ti * tile_stride;
        const uint32_t* cbase = p.col_planes + (size_t)J * tile_stride;
        // one bulk copy per run of consecutive staged words (the whole step when the order is contiguous)
        int kk = 0;
        while (kk < nw) {
            const int w = word_of(w0 + kk);
            int run = 1;
            while (kk + run < nw && word_of(w0 + kk + run) == w + run) ++run;
            bulk_g2s(dst + kk * op_words, rbase + (size_t)w * word_stride, (uint32_t)run * op_bytes, &full[slot]);
            bulk_g2s(dst + (KW + kk) * op_words, cbase + (size_t)w * word_stride, (uint32_t)run * op_bytes, &full[slot]);
            if (second)
                bulk_g2s(dst + (2 * KW + kk) * op_words, cbase + tile_stride + (size_t)w * word_stride, (uint32_t)run * op_bytes, &full[slot]);
            kk += run;
        }
        pst->step = st + 1; pst->count = pr_count + 1;
    };
    if (tid == 0) {
        pst->count = 0; pst->step = 0; pst->nsteps = 0; pst->ti = 0; pst->J = 0; pst->ch = 0; pst->done = 0; pst->jend = 0;
        for (int s = 0; s < PK_NS; ++s) { slot_cnt[s] = 0; produce_one(); }
    }
    __syncthreads();

    // ---- consumers (all 8 warps) ----
    // Accumulators count in units of 4 (acc = 4 * count [+ carried class offset]).
    uint32_t acc[4][NB];
    int gj[NB];
    int sgn[NB];
    uint32_t om = 0u;          // tie-coin orientation [i<j] of this thread's pairs (all-ones / zero) ...
    uint32_t obits = 0u;       // ... and per pair (bit a*NB+b), used only when the 4xNB block straddles i == j
    bool uniform = true, all_valid = false;
    int gi0 = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[a][b] = 0u;
#pragma unroll
    for (int b = 0; b < NB; ++b) { gj[b] = -1; sgn[b] = 1; }
    const uint32_t four = p.one << 2;
    const uint32_t lut_s = smem_u32(lut);
    const uint32_t tab_row_s = smem_u32(tab_s) + (uint32_t)(ty * 4) * 4u;   // + bin*256 + a*4

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
        return (3u * (acc4 >> 28) + it) * 256u;
    };

    for (int cc = 0;; ++cc) {
        const int slot = cc % PK_NS;
        mbar_wait(&full[slot], (uint32_t)(cc / PK_NS) & 1u);
        PkMeta m;   // snapshot now: the slot's metadata is rewritten as soon as the last warp releases it
        {
            const volatile PkMeta* vm = &metas[slot];
            m.ti = vm->ti; m.J = vm->J; m.w0 = vm->w0; m.nw = vm->nw; m.flags = vm->flags;
        }
        if (m.flags & PKF_TERM) break;
        if (m.flags & PKF_FIRST_J) {
            gi0 = m.ti * REO_TILE + ty * 4;
            const int4 g4 = *reinterpret_cast<const int4*>(p.col_gene + m.J * REO_TILE + tx * 4);
            gj[0] = g4.x; gj[1] = g4.y; gj[2] = g4.z; gj[3] = g4.w;
            if (p.col_sign) {
                const char4 s4 = *reinterpret_cast<const char4*>(p.col_sign + m.J * REO_TILE + tx * 4);
                sgn[0] = s4.x; sgn[1] = s4.y; sgn[2] = s4.z; sgn[3] = s4.w;
            }
            if (NB == 8) {   // second column tile of the step (absent at the end of an odd chunk: pad columns)
                if (m.flags & PKF_SECOND) {
                    const int4 h4 = *reinterpret_cast<const int4*>(p.col_gene + (m.J + 1) * REO_TILE + tx * 4);
                    gj[NB - 4] = h4.x; gj[NB - 3] = h4.y; gj[NB - 2] = h4.z; gj[NB - 1] = h4.w;
                    if (p.col_sign) {
                        const char4 s4 = *reinterpret_cast<const char4*>(p.col_sign + (m.J + 1) * REO_TILE + tx * 4);
                        sgn[NB - 4] = s4.x; sgn[NB - 3] = s4.y; sgn[NB - 2] = s4.z; sgn[NB - 1] = s4.w;
                    }
                } else {
                    gj[NB - 4] = -1; gj[NB - 3] = -1; gj[NB - 2] = -1; gj[NB - 1] = -1;
                }
            }
            // columns ascend.  Orientation [i<j] is uniform over the 4xNB block unless it straddles i == j.
            const bool cols_ok = (gj[0] >= 0) && (gj[NB - 1] >= 0);
            const bool all_lt = cols_ok && (gi0 + 3 < gj[0]);     // every row index below every column index
            const bool all_ge = cols_ok && (gi0 > gj[NB - 1]);    // every row index above every column index
            uniform = all_lt || all_ge;
            all_valid = uniform && (gi0 + 3 < p.r);
            om = all_lt ? 0xffffffffu : 0u;
            obits = all_lt ? (NB == 8 ? 0xffffffffu : 0xffffu) : 0u;
            if (!uniform) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < NB; ++b) obits |= (uint32_t)(gi0 + a < gj[b]) << (a * NB + b);
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
        for (int kk = 0; kk < m.nw; ++kk) {
            const bool boundary = (m.w0 + kk == p.WA);   // first word that is not a pure group-A word
            if (boundary && !p.mixed) {
                // group A finished: classify ic, restart the counters with the class carried along
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < NB; ++b) acc[a][b] = flushA(acc[a][b], a * NB + b);
            }
            const uint32_t* xr = srow + kk * op_words + ty * 4;
            const uint32_t* yc = scol + kk * op_words + tx * 4;
            const uint32_t* yc2 = yc + KW * op_words;     // second column tile (NB == 8)
            uint32_t bor[4][NB];
            if (FLT) {
                // raw-value path: per sample d = x - y; "greater" iff d >= 0.1, tie iff -0.1 < d < 0.1
                // (== abs(x - y) < 0.1 ? coin : x > y, src:72-76), 32 samples -> two bit masks per pair
                uint32_t gtm[4][4], lem[4][4];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) { gtm[a][b] = 0u; lem[a][b] = 0u; }
                const bool f32 = p.flt == 2;
                const double* xd = reinterpret_cast<const double*>(srow + kk * op_words + REO_TILE) + ty * 4;
                const double* yd = reinterpret_cast<const double*>(scol + kk * op_words + REO_TILE) + tx * 4;
#pragma unroll 2
                for (int sidx = 0; sidx < 32; ++sidx) {
                    const double2 x01 = *reinterpret_cast<const double2*>(xd + sidx * REO_TILE);
                    const double2 x23 = *reinterpret_cast<const double2*>(xd + sidx * REO_TILE + 2);
                    const double2 y01 = *reinterpret_cast<const double2*>(yd + sidx * REO_TILE);
                    const double2 y23 = *reinterpret_cast<const double2*>(yd + sidx * REO_TILE + 2);
                    const double x[4] = {x01.x, x01.y, x23.x, x23.y};
                    const double y[4] = {y01.x, y01.y, y23.x, y23.y};
                    const uint32_t bit = 1u << sidx;
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            // Matrix{Float32}: Julia rounds x - y to Float32 before comparing with the Float64 literal 0.1
                            const double d = f32 ? (double)(__double2float_rn(x[a]) - __double2float_rn(y[b])) : x[a] - y[b];
                            if (d >= 0.1) gtm[a][b] |= bit;
                            if (d > -0.1) lem[a][b] |= bit;
                        }
                }
                const uint4 xv = *reinterpret_cast<const uint4*>(xr);
                const uint4 yv = *reinterpret_cast<const uint4*>(yc);
                const uint32_t xc[4] = {xv.x, xv.y, xv.z, xv.w};
                const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {     // the raw-value path is instantiated with NB == 4 only
                        uint32_t coin = xc[a] ^ yw[b] ^ om;
                        if (!uniform) coin ^= (0u - (((obits >> (a * NB + b)) ^ obits) & 1u));
                        bor[a][b] = gtm[a][b] | (lem[a][b] & ~gtm[a][b] & coin);
                    }
            } else {
                {
                    const uint4 xv = *reinterpret_cast<const uint4*>(xr);
                    const uint4 yv = *reinterpret_cast<const uint4*>(yc);
                    const uint32_t x[4] = {xv.x, xv.y, xv.z, xv.w};
                    uint32_t y[NB];
                    y[0] = yv.x; y[1] = yv.y; y[2] = yv.z; y[3] = yv.w;
                    if (NB == 8) {
                        const uint4 zv = *reinterpret_cast<const uint4*>(yc2);
                        y[NB - 4] = zv.x; y[NB - 3] = zv.y; y[NB - 2] = zv.z; y[NB - 1] = zv.w;
                    }
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
                        const uint32_t x[4] = {xv.x, xv.y, xv.z, xv.w};
                        uint32_t y[NB];
                        y[0] = yv.x; y[1] = yv.y; y[2] = yv.z; y[3] = yv.w;
                        if (NB == 8) {
                            const uint4 zv = *reinterpret_cast<const uint4*>(yc2 + pl * REO_TILE);
                            y[NB - 4] = zv.x; y[NB - 3] = zv.y; y[NB - 2] = zv.z; y[NB - 1] = zv.w;
                        }
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
                        const uint32_t x[4] = {xv.x, xv.y, xv.z, xv.w};
                        uint32_t y[NB];
                        y[0] = yv.x; y[1] = yv.y; y[2] = yv.z; y[3] = yv.w;
                        if (NB == 8) {
                            const uint4 zv = *reinterpret_cast<const uint4*>(yc2 + pl * REO_TILE);
                            y[NB - 4] = zv.x; y[NB - 3] = zv.y; y[NB - 2] = zv.z; y[NB - 1] = zv.w;
                        }
#pragma unroll
                        for (int a = 0; a < 4; ++a)
#pragma unroll
                            for (int b = 0; b < NB; ++b) bor[a][b] = lop3_b2(x[a], y[b], bor[a][b]);
                    }
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
        if (m.flags & PKF_LAST_J) {
            // group B finished: look up the bin, add to the shared table
            if (all_valid) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        const uint32_t addr = tab_row_s + binB(acc[a][b], a * NB + b) + (uint32_t)(a * 4);
                        asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(sgn[b]) : "memory");
                    }
            } else {   // pad columns, rows beyond r, self pairs: only on the matrix edges and the diagonal
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int gi = gi0 + a;
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        const uint32_t off = binB(acc[a][b], a * NB + b);
                        if (gj[b] >= 0 && gi != gj[b] && gi < p.r) {
                            const uint32_t addr = tab_row_s + off + (uint32_t)(a * 4);
                            asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(sgn[b]) : "memory");
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            // the last warp to finish this slot refills it (3 steps ahead of its consumers)
            if (atomicAdd(&slot_cnt[slot], 1) == PK_WARPS - 1) {
                slot_cnt[slot] = 0;
                produce_one();
            }
        }
        if (m.flags & PKF_LAST_ITEM) {
            consumer_bar();
            for (int i = tid; i < REO_TILE * 9; i += PK_THREADS) {
                const int v = tab_s[i];
                const int gi = m.ti * REO_TILE + (i & 63);
                if (v != 0 && gi < p.r) atomicAdd(&p.table[(size_t)gi * 9 + (i >> 6)], v);
                tab_s[i] = 0;
            }
            consumer_bar();
        }
    }
}

static void pair_lut_sizes(const ReoPairParams& p, int* sza, int* szb, int* use) {
    // slots whose borrow bits can be counted for group A / group B (real samples + pad slots) + 1
    const int a = p.nA + p.padA + 1, b = p.nB + p.padB + 1;
    *sza = a; *szb = b;
    *use = (2 * a + 6 * b) <= PK_LUT_MAX_WORDS;
}
// nct = column tiles per step (1, or 2 for the 4x8 register tile)
static int pair_kw(int NP, int W, int lut_words, int nct) {
    int kw = ((nct == 2 ? PK_SMEM_BUDGET_NB : PK_SMEM_BUDGET) - lut_words * 4) / (PK_NS * (1 + nct) * NP * REO_TILE * 4);
    if (kw > PK_MAX_KW) kw = PK_MAX_KW;
    if (kw > W) kw = W;
    if (kw < 1) kw = 1;
    const int nsteps = (W + kw - 1) / kw;   // spread the words evenly over the steps of one column tile
    return (W + nsteps - 1) / nsteps;
}
static size_t pair_smem_bytes(int NP, int KW, int lut_words, int nct) {
    return (size_t)PK_NS * (1 + nct) * KW * NP * REO_TILE * 4 + REO_TILE * 9 * 4 + PK_NS * sizeof(PkMeta) + PK_NS * 8 +
           sizeof(PkProd) + PK_NS * 4 + (size_t)lut_words * 4 + 16;
}

template <int NPT, bool LUT, bool FLT = false>
static cudaError_t launch_np_lut(const ReoPairParams& p, int lut_words, int num_sms, cudaStream_t st) {
    constexpr int NB = FLT ? 4 : PK_NB;
    const size_t smem = pair_smem_bytes(FLT ? REO_FLT_OPWORDS / REO_TILE : p.NP, p.KW, lut_words, NB / 4);
    cudaError_t e = cudaFuncSetAttribute(reo_pair_kernel<NPT, LUT, FLT, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int nitems = (p.t1 - p.t0) * p.njchunks;
    int grid = (NB == 8 ? 2 : PK_CTAS_PER_SM) * num_sms;
    if (grid > nitems) grid = nitems;
    if (grid < 1) return cudaSuccess;
    reo_pair_kernel<NPT, LUT, FLT, NB><<<grid, PK_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

// column tiles one step of the rank-path kernel consumes: chunks of column tiles should be multiples of it
int reo_pairs_col_tiles_per_step(bool flt) { return flt ? 1 : PK_NB / 4; }

template <int NPT>
static cudaError_t launch_np(ReoPairParams p, int num_sms, cudaStream_t st) {
    pair_lut_sizes(p, &p.lutSZA, &p.lutSZB, &p.use_lut);
    const int lut_words = p.use_lut ? 2 * p.lutSZA + 6 * p.lutSZB : 0;
    p.KW = pair_kw(p.NP, p.W, lut_words, PK_NB / 4);
    p.one = 1u;
    return p.use_lut ? launch_np_lut<NPT, true>(p, lut_words, num_sms, st)
                     : launch_np_lut<NPT, false>(p, lut_words, num_sms, st);
}

cudaError_t reo_launch_pairs(const ReoPairParams& p, int num_sms, cudaStream_t st) {
    if (p.t1 <= p.t0 || p.ntc <= 0) return cudaSuccess;
    if (p.flt) {   // raw FP64 values: operand word = REO_FLT_OPWORDS words
        ReoPairParams q = p;
        pair_lut_sizes(q, &q.lutSZA, &q.lutSZB, &q.use_lut);
        const int lut_words = q.use_lut ? 2 * q.lutSZA + 6 * q.lutSZB : 0;
        q.KW = pair_kw(REO_FLT_OPWORDS / REO_TILE, q.W, lut_words, 1);
        q.one = 1u;
        return q.use_lut ? launch_np_lut<0, true, true>(q, lut_words, num_sms, st)
                         : launch_np_lut<0, false, true>(q, lut_words, num_sms, st);
    }
    return launch_np<0>(p, num_sms, st);   // run-time plane count (the specialised chains live in reo_pairs2.cu)
}

// ---- small-block debug/parity kernel: one thread per (row, col) pair, straight from the planes ----
__global__ void pair_counts_small_kernel(const uint32_t* __restrict__ planes, int W, int NP,
                                         const int32_t* __restrict__ word_order, int WA,
                                         const int32_t* __restrict__ rows, int nrows,
                                         const int32_t* __restrict__ cols, int ncols, int32_t* nre, int32_t* rest,
                                         int padA, int padB, int mixed, uint32_t maskA, uint32_t maskB, int f32) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nrows * ncols) return;
    const int gi = rows[idx / ncols], gj = cols[idx % ncols];
    const uint32_t om = gi < gj ? 0xffffffffu : 0u;
    if (NP == 0) {   // float path: [64 coin words][32][64] FP64 per (tile, word)
        const size_t ws = REO_FLT_OPWORDS, ts = (size_t)W * ws;
        int a = 0, b = 0;
        for (int k = 0; k < W; ++k) {
            const int w = word_order[k];
            const uint32_t* bi = planes + (size_t)(gi >> 6) * ts + (size_t)w * ws;
            const uint32_t* bj = planes + (size_t)(gj >> 6) * ts + (size_t)w * ws;
            const double* xi = reinterpret_cast<const double*>(bi + REO_TILE) + (gi & 63);
            const double* yj = reinterpret_cast<const double*>(bj + REO_TILE) + (gj & 63);
            const uint32_t coin = bi[gi & 63] ^ bj[gj & 63] ^ om;
            uint32_t bor = 0u;
            for (int sidx = 0; sidx < 32; ++sidx) {
                const double xv = xi[sidx * REO_TILE], yv = yj[sidx * REO_TILE];
                const double d = f32 ? (double)(__double2float_rn(xv) - __double2float_rn(yv)) : xv - yv;
                const uint32_t gt = d >= 0.1, tie = (d > -0.1) && !(d >= 0.1);
                bor |= (gt | (tie & (coin >> sidx))) << sidx;
            }
            if (k == WA && mixed) { a += __popc(bor & maskA); b += __popc(bor & maskB); }
            else if (k < WA) a += __popc(bor);
            else b += __popc(bor);
        }
        nre[idx] = a - (int)(om & (uint32_t)padA);
        rest[idx] = b - (int)(om & (uint32_t)padB);
        return;
    }
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
        if (k == WA && mixed) { a += __popc(bor & maskA); b += __popc(bor & maskB); }
        else if (k < WA) a += __popc(bor);
        else b += __popc(bor);
    }
    nre[idx] = a - (int)(om & (uint32_t)padA);
    rest[idx] = b - (int)(om & (uint32_t)padB);
}

cudaError_t reo_launch_pair_counts_small(const ReoStaged& S, const int32_t* word_order, int WA, const int32_t* rows,
                                         int nrows, const int32_t* cols, int ncols, int32_t* nre, int32_t* rest,
                                         int padA, int padB, int mixed, uint32_t maskA, uint32_t maskB, cudaStream_t st) {
    const int n = nrows * ncols;
    if (n <= 0) return cudaSuccess;
    pair_counts_small_kernel<<<(n + 127) / 128, 128, 0, st>>>(S.planes, S.W, S.flt ? 0 : S.NP, word_order, WA, rows, nrows, cols,
                                                              ncols, nre, rest, padA, padB, mixed, maskA, maskB, S.flt_f32 ? 1 : 0);
    return cudaGetLastError();
}
