// reo_internal.cuh -- shared declarations of libreo_cuda.so (not part of the C ABI).
//
// Data layout in HBM:
//   planes[t][w][p][l]  uint32, t = gene tile (64 genes), w = sample word (32 samples of ONE group
//   level, levels padded to a multiple of 32), p = bit-plane (p = 0: tie-coin bits u(i,s);
//   p = 1..B: bits of the per-sample dense rank, LSB first), l = gene within the tile.
//   Bit `b` of a word is sample slot 32*w + b.  Pad slots and pad genes are all-zero.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "reo.h"

#define REO_TILE 64            // genes per tile
#define REO_MAX_BITS 20        // rank bits (u16 ranks up to 65535 genes, u32 ranks beyond)
#define REO_MAX_PLANES (REO_MAX_BITS + 1)
// float path: one operand word = 64 coin words + 32 samples x 64 genes FP64
#define REO_FLT_OPWORDS (REO_TILE + 32 * REO_TILE * 2)

// ---- tie coin: must match oracle/reo_oracle.{py,c} bit for bit --------------------------------
__host__ __device__ __forceinline__ uint32_t reo_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
// bit 31 of reo_mix32(x): the mixer's last xor-shift (x ^= x >> 16) cannot change the top bit
__host__ __device__ __forceinline__ uint32_t reo_mix32_top(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu;
    return x >> 31;
}
__host__ __device__ __forceinline__ uint32_t reo_coin_u(uint32_t seed_lo, uint32_t seed_hi, uint32_t gene,
                                                        uint32_t sample) {
    uint32_t h = reo_mix32(seed_lo ^ (gene * 0x9E3779B1u));
    h = reo_mix32(h + sample * 0x85EBCA77u + seed_hi);
    return h >> 31;
}

// ---- staged matrix ------------------------------------------------------------------------------
struct ReoStaged {
    int64_t r = 0, c = 0;
    int gnum = 0;
    int NT = 0;        // gene tiles
    int W = 0;         // sample words (all levels)
    int B = 0;         // rank bits
    int NP = 0;        // planes = B + 1
    uint32_t* planes = nullptr;      // [NT][W][NP][64]; float path: [NT][W][REO_FLT_OPWORDS]
    uint8_t* word_np = nullptr;      // [W] planes of each sample word that hold anything (coin plane + rank bits its samples need)
    double planes_per_word = 0.0;    // mean planes the pair kernel runs per word (NP when nothing is skipped)
    bool flt = false;                // non-integral input: raw FP64 values instead of rank planes
    bool flt_f32 = false;            // ... that came from a Float32 matrix: differences are rounded to Float32 (src:72 in Julia)
    std::vector<int> lev_word0;      // first word of each level
    std::vector<int> lev_words;      // words of each level
    std::vector<int> lev_n;          // real samples of each level
    // two-level inputs only: the tails (n % 32) of both levels share one word when they fit
    int mixed_word = -1;             // index of that word, -1 if none
    int mixed_rem[2] = {0, 0};       // level 0 occupies bits [0, rem0), level 1 bits [rem0, rem0 + rem1)
    bool valid = false;
    size_t word_stride() const { return flt ? (size_t)REO_FLT_OPWORDS : (size_t)NP * REO_TILE; }
    size_t tile_stride() const { return (size_t)W * word_stride(); }
};

// ---- pair kernel parameters ---------------------------------------------------------------------
struct ReoPairParams {
    const uint32_t* row_planes;  // [NT][W][NP][64]
    const uint32_t* col_planes;  // [NTc][W][NP][64] (panel of compacted columns, may alias row_planes)
    const int32_t* col_gene;     // [NTc*64] gene index of each panel column, -1 = pad
    const int8_t* col_sign;      // [NTc*64] +1/-1 (nullptr -> +1)
    // two-group word order without a table: k < WA -> segA0 + k; then the mixed word (if any); then the
    // other levels' words: segB0 .. segB0 + segB0len - 1, then segB1 ...
    int segA0, mixedW, segB0, segB0len, segB1;
    int32_t* table;              // [.. ][9] Int32, += sign per (row gene, category)
    unsigned int* counter;       // dynamic work counter (zeroed before launch)
    int W, WA, NP;
    int r;                       // real genes
    int t0, t1;                  // row tile range [t0, t1)
    int ntc;                     // column tiles in the panel
    int jchunk, njchunks;        // column tiles per work item, items per row tile
    int nA, nB, padA, padB, thrA, thrB;
    int mixed;                   // 1: word_order[WA] holds the tails of BOTH groups (maskA / maskB select them)
    uint32_t maskA, maskB;
    int KW;                      // sample words per pipeline slot (set by the launcher)
    int use_lut, lutSZA, lutSZB; // class lookup tables in shared memory (set by the launcher)
    int flt;                     // 1: planes hold raw FP64 values (REO_FLT_OPWORDS words per operand word); 2: the same,
                                 //    from a Float32 matrix (x - y is rounded to Float32 before the 0.1 test)
    unsigned int one;            // == 1 (see mad_acc in reo_pairs.cu)
};

// ---- pair kernel v2 (reo_pairs2.cu): warp-specialised, symmetric ------------------------------------
// Row panel and column panel are both [tiles][W][NP][64] with a gene id (and optional sign) per panel position.
// The first `nsym` row tiles ARE the column panel (same genes, same order): there every unordered pair is
// evaluated once (tiles I <= J) and updates both genes (src:385-386: q for gene i, 10 - q for gene j).  Row tiles
// from nsymp = round_up(nsym, T) on are ordinary rows (one-sided update).  nsym == 0: one-sided everywhere.
struct ReoPair2Params {
    const uint32_t* row_planes;
    const uint32_t* col_planes;
    const int32_t* row_gene;     // [ntr*64] gene of each row position, -1 = pad
    const int32_t* col_gene;     // [round_up(ntc, 2)*64] gene of each column position, -1 = pad
    const int8_t* row_sign;      // sign of the row gene AS A COLUMN (symmetric region); nullptr -> +1
    const int8_t* col_sign;      // [round_up(ntc, 2)*64]; nullptr -> +1
    int32_t* table;              // [r][9] Int32, this rank's partial sums
    unsigned int* counter;       // dynamic work counter (zeroed before launch)
    const uint8_t* word_np;      // [W] planes in use per staged word (nullptr: all NP); a word whose top plane is empty skips it
    int ntr, ntc, nsym;          // row tiles, column tiles, symmetric row tiles
    int T;                       // block edge in tiles (even): one work item = T row tiles x T column tiles
    int SS, SSc;                 // supertile edge in blocks (L2 locality): rows, columns (SSc < SS only without a symmetric region)
    int perm_mul;                // unit modulo SS*SS*RS that scrambles the item order inside a supertile
    int RS;                      // row parts per block: one work item = T / RS row tiles x T column tiles
    int NBs, NBr, NBc;           // blocks: symmetric rows, all rows, columns
    int Ms, Mc;                  // supertile rows of the symmetric region, supertile columns
    long long tri, NSUP;         // supertiles in the triangle, supertiles in total
    int nitems;                  // work items of this rank (every world-th item; invalid ones are skipped)
    int rank, world;
    int segA0, mixedW, segB0, segB0len, segB1;
    int W, WA, NP;
    int nA, nB, padA, padB, thrA, thrB;
    int mixed;
    uint32_t maskA, maskB;
    int KW, NS;                  // sample words per ring stage, ring stages
    int use_lut, lutSZA, lutSZB;
    unsigned int one;
};
cudaError_t reo_launch_pairs2(ReoPair2Params p, int num_sms, cudaStream_t st);
int reo_pairs2_block_edge(int W, int NP);   // T for a staged matrix
// host replay of a rank's tile pairs: (row tile, column tile, 1 = updates row genes | 2 = updates column genes) triples
long long reo_pairs2_plan(ReoPair2Params p, int32_t* out, long long cap);

// kernels / launchers implemented in the .cu files
struct ReoDev;  // per-device state (reo_api.cu)

cudaError_t reo_launch_pairs(const ReoPairParams& p, int num_sms, cudaStream_t st);
int reo_pairs_col_tiles_per_step(bool flt);   // 1, or 2 when the pair kernel is built with the 4x8 register tile
cudaError_t reo_launch_pair_counts_small(const ReoStaged& S, const int32_t* word_order, int WA, const int32_t* rows,
                                         int nrows, const int32_t* cols, int ncols, int32_t* nre, int32_t* rest,
                                         int padA, int padB, int mixed, uint32_t maskA, uint32_t maskB,
                                         cudaStream_t st);

// staging (reo_stage.cu).  Columns are addressed through lists so that a rank can stage only its share:
// list position j -> data column src_col[j], original sample sample_id[j].
// Two bitmap tiers: wide == 0 -> 256 threads, range < 65536, 8 CTAs/SM; wide != 0 -> 1024 threads, range < 1.3 M,
// 1 CTA/SM.  cols (optional) lists the positions to process; columns that do not fit go to over_list/over_count.
#define REO_U16_STAGED 100   // internal element type: a chunk of the input narrowed to u16 by the host (reo_host.cpp)
cudaError_t reo_launch_rank_columns(const void* data, int dtype, int64_t r, int64_t ld, int64_t col0, int ncols,
                                    const int32_t* cols, int wide, const int32_t* src_col, const int32_t* sample_id,
                                    const int32_t* slot_of_sample, void* ranks, int rank_bytes, int64_t rpad,
                                    int* max_distinct, int* flags /*[0]=nonintegral*/, int* over_count,
                                    int32_t* over_list, cudaStream_t st);
cudaError_t reo_launch_rank_fallback(const void* data, int dtype, int64_t r, int64_t ld, const int32_t* fallback_list,
                                     int nfb, const int32_t* src_col, const int32_t* sample_id,
                                     const int32_t* slot_of_sample, void* ranks, int rank_bytes, int64_t rpad,
                                     int* max_distinct, unsigned long long* scratch_keys, uint32_t* scratch_rank,
                                     int64_t rpow2, cudaStream_t st);
cudaError_t reo_launch_bitplanes(const void* ranks, int rank_bytes, int64_t rpad, int64_t r, const int32_t* sample_of_slot,
                                 int NT, int w_lo, int w_n, int w_stride, int NP, uint32_t seed_lo, uint32_t seed_hi,
                                 uint32_t* planes, cudaStream_t st);
cudaError_t reo_launch_word_planes(const uint32_t* planes, int NT, int W, int NP, uint8_t* word_np, int* run_sum, cudaStream_t st);
cudaError_t reo_launch_unshard_planes(const uint32_t* gathered, uint32_t* planes, int NT, int W, int wq, int wb,
                                      cudaStream_t st);
cudaError_t reo_launch_gather_panel(const uint32_t* planes, int W, int NP, const int32_t* col_gene, int ntc,
                                    uint32_t* panel, cudaStream_t st);

cudaError_t reo_launch_fstage(const void* data, int dtype, int64_t r, int64_t ld, const int32_t* sample_of_slot,
                              const int32_t* col_of_sample, int NT, int w_lo, int w_n, int w_stride, uint32_t seed_lo,
                              uint32_t seed_hi, uint32_t* planes, cudaStream_t st);
cudaError_t reo_launch_gather_panel_flt(const uint32_t* planes, int W, const int32_t* col_gene, int ntc, uint32_t* panel,
                                        cudaStream_t st);

// pre-processing next to the path (reo_prep.cu)
cudaError_t reo_launch_pseudobulk(const void* data, int dtype, int64_t r, int64_t ld, const int32_t* cell_ptr,
                                  const int32_t* cell_list, int nprofiles, void* out, cudaStream_t st);
cudaError_t reo_launch_detect_counts(const void* data, int dtype, int64_t r, int64_t c, int64_t ld, int32_t* per_cell,
                                     int32_t* per_gene, cudaStream_t st);
cudaError_t reo_launch_subset(const void* data, int dtype, int64_t ld, const int32_t* gene_list, int64_t r2,
                              const int32_t* cell_list, int64_t c2, void* out, cudaStream_t st);

// statistics (reo_stats.cu)
cudaError_t reo_launch_mccullagh_tables(const int32_t* table, int64_t r, double* result /*col-major r x 15*/,
                                        cudaStream_t st);
cudaError_t reo_launch_mccullagh_kxk(const int64_t* tables, int64_t n, int k, double* out, cudaStream_t st);
struct ReoSortWs {  // workspace for reo_sort
    unsigned long long* keys = nullptr;  // chunk-sorted keys
    uint32_t* idx = nullptr;
    int32_t* pos = nullptr;
    unsigned int* cnt = nullptr;         // per-chunk arrival counters (zero between launches)
    int64_t cap = 0;
};
cudaError_t reo_sort_reserve(ReoSortWs& ws, int64_t n, cudaStream_t st);
cudaError_t reo_launch_sort_f64(const double* x, int64_t n, double* sorted, int32_t* perm, ReoSortWs& ws,
                                cudaStream_t st);
cudaError_t reo_launch_trimmed_std(const double* sorted, int64_t n, double* se_out, double* leaf_ws, cudaStream_t st);
cudaError_t reo_launch_null_pvals(const double* d1, int64_t n, const double* se, double* pval, cudaStream_t st);
cudaError_t reo_launch_p_order(const double* sorted, const int32_t* perm1, int64_t n, double* pval, const double* se,
                               double* sorted_p, int32_t* perm2, cudaStream_t st);
cudaError_t reo_launch_bh(const double* sorted_p, const int32_t* perm, int64_t n, double* padj, double* ws,
                          uint8_t* mask_new /*nullable: also write src:417's inds*/, double pval_deg, double padj_deg,
                          cudaStream_t st);
// inds = !(p<=pd && q<=qd) (src:417)
cudaError_t reo_launch_inds(const double* pval, const double* padj, int64_t r, double pval_deg, double padj_deg,
                            uint8_t* mask_new, cudaStream_t st);
// counts[0]=sum(old) counts[1]=sum(new) counts[2]=n_changed; changed_gene/changed_sign = ascending list over
// old xor new (padded to a multiple of 64 with -1/0)
cudaError_t reo_launch_mask_diff(int64_t r, const uint8_t* mask_old, const uint8_t* mask_new, int32_t* counts,
                                 int32_t* changed_gene, int8_t* changed_sign, cudaStream_t st);
// compaction of a mask into an ascending column list (padded to a multiple of 64 with -1); counts[0]=n
cudaError_t reo_launch_mask_to_list(const uint8_t* mask, int64_t r, int32_t* list, int32_t* count, cudaStream_t st);
// [C ascending | pad to a T-tile block | N ascending | pad to cap] with signs (see reo_stats.cu)
cudaError_t reo_launch_sym_lists(int64_t r, const uint8_t* m_old, const uint8_t* m_new, int T, int32_t* gene, int8_t* sign,
                                 int32_t* counts_out, int64_t cap, cudaStream_t st);
cudaError_t reo_launch_sum_slices(const int32_t* all, int world, int64_t n, int32_t* out, cudaStream_t st);
cudaError_t reo_launch_updown(const double* result, int64_t r, double pval_deg, double padj_deg, int8_t* updown,
                              cudaStream_t st);
