/*
 * reo.h -- C ABI of libreo_cuda.so: the B200-native REO core behind RankCompV3.jl's
 *          `identify_degs` (reference: src/RankCompV3.jl:339-438, called from reoa() at
 *          src/RankCompV3.jl:652-662).
 *
 * The reference has no FFI layer; the seam this library replaces is the Julia call
 *     identify_degs(Matrix(df_expr), meta_group.Group, gene_names, pval_reo, pval_deg, padj_deg,
 *                   ref_gene_vec, n_iter, n_conv)                       (src:652-662, sig. 339-350)
 * INTEGRATION.md shows the `ccall` shim a maintainer would add.  Plain pointers and sizes only;
 * no exceptions or aborts cross this boundary: every entry point returns a reo_status and
 * leaves caller outputs untouched on failure.  A handle is re-entrant per handle, owns its CUDA
 * streams, and never synchronises the legacy default stream.
 *
 * Conventions
 *  - `data` is column-major r x c with leading dimension ld (Julia `Matrix`): data[i + ld*s]
 *    is gene i in sample s.  Sample s belongs to level group_id[s] (0-based, levels numbered
 *    in order of first appearance, i.e. Julia's `unique(group)`, src:353).
 *  - K = 1 when gnum == 2 (src:387-389, 431-434), else gnum (one-vs-rest per level).
 *  - `thresholds` is the 2 x gnum column-major Int32 matrix of src:362
 *    (thresholds[0+2k] for level k, thresholds[1+2k] for the rest); NULL -> computed from pval_reo.
 *  - Ties (|x-y| < 0.1, src:72) are broken by the deterministic coin documented in DESIGN.md
 *    (`seed` of reo_create), which replaces the reference's rand(Bool) (src:73).
 */
#ifndef REO_H
#define REO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REO_VERSION 201 /* 0.2.0 */

typedef struct reo_handle_s* reo_handle_t;

typedef enum {
    REO_OK = 0,
    REO_ERR_DIM = -1,         /* DimensionMismatch of src:355-356 (c != length(group), < 2 levels) */
    REO_ERR_ARG = -2,         /* bad argument (NULL pointer, unknown dtype, negative size, ...)    */
    REO_ERR_CUDA = -3,        /* CUDA runtime error, see reo_last_error                            */
    REO_ERR_COMM = -4,        /* collective (NCCL / user all-gather callback) failed               */
    REO_ERR_OOM = -5,         /* host or device allocation failed                                  */
    REO_ERR_BOUNDS = -6,      /* r <= 10: the reference's BoundsError at src:411                   */
    REO_ERR_UNSUPPORTED = -7, /* input outside what this build supports (see message)              */
    REO_ERR_STATE = -8        /* stage-level call without a staged matrix                          */
} reo_status;

typedef enum { REO_I64 = 0, REO_F64 = 1, REO_I32 = 2, REO_F32 = 3 } reo_dtype;

/* reo_create flags */
#define REO_FLAG_NONE 0u
/* reo_identify_degs / reo_stage flags */
#define REO_DATA_ON_DEVICE 1u /* `data` is a device pointer on the handle's first device */
/* reo_identify_degs: result, updown and final_ref are page-locked host memory (reo_host_alloc, cudaHostAlloc or
 * cudaHostRegister).  The device->host copies then land in them directly (no staging copy on the host) and
 * overlap the last evaluation; the price: after a FAILED call they may hold intermediate values. */
#define REO_OUT_PINNED 2u

#define REO_MAX_ITER_LOG 256

/* Out-struct for logging: lets the Julia shim print the reference's @info lines (src:418-420). */
typedef struct {
    int32_t iters_done;                 /* evaluations performed for the last k (every k: reo_iter_log) */
    int32_t converged;                  /* 1 if the n_conv criterion stopped the loop (src:419-422) */
    int32_t n_deg[REO_MAX_ITER_LOG];    /* # DEGs per evaluation (src:418)                          */
    int32_t n_ref[REO_MAX_ITER_LOG];    /* size of the reference set used by each evaluation        */
    int32_t rank_bits;                  /* bit-planes per rank (B); planes staged = B + 1           */
    int32_t sample_words;               /* 32-sample words staged (all levels, padded)              */
    int64_t compares;                   /* ordered (gene, ref gene, sample) triples evaluated       */
    double ms_stage;                    /* H2D + rank + bit-plane staging                           */
    double ms_pairs;                    /* pair-count/class/table kernels (all iterations)          */
    double ms_stats;                    /* McCullagh + empirical null + BH + mask kernels           */
    double ms_total;                    /* whole call, device timeline                              */
    double ms_wall;                     /* whole call, host wall clock                              */
    int32_t pair_launches;              /* pair-kernel launches                                     */
    int32_t kernel_launches;            /* all kernel launches of this call                         */
    int64_t ordered_triples;            /* (row gene, column gene, sample) triples the evaluated pairs stand
                                           for: rows x columns x samples per table build (W_ord); `compares`
                                           counts each mirrored pair once, as the reference does (src:366-372) */
    double planes_per_word;             /* bit planes the pair kernel's borrow chain runs per 32-sample word, averaged
                                           over the staged words (<= rank_bits + 1: a word whose samples need fewer
                                           rank bits than the widest sample skips its empty top plane)          */
} reo_stats;

/* Multi-process sharding hook: all-gather `bytes_per_rank` bytes per rank, in place, inside the
 * device buffer `dev_buf` (rank q's slice at offset q*bytes_per_rank).  The library has
 * synchronised its stream before the call and expects the result to be complete on return.   */
typedef int (*reo_allgather_fn)(void* ctx, void* dev_buf, uint64_t bytes_per_rank);

int reo_version(void);

/* ndev >= 1 devices driven by this (single) process; devs == NULL -> 0..ndev-1.  With ndev > 1 gene-row
 * tiles are sharded over the devices (one host thread each inside every call) and the per-gene tables are
 * all-gathered with NCCL (ncclCommInitAll).  seed keys the tie coins.  The reference has no handle. */
int reo_create(reo_handle_t* out, int ndev, const int* devs, uint64_t seed, uint32_t flags);
int reo_destroy(reo_handle_t h);
const char* reo_last_error(reo_handle_t h); /* h may be NULL: last create error */

/* One-process-per-GPU mode (torchrun): this handle is rank `rank` of `world`; gene-row tiles
 * are sharded by rank and tables exchanged through `fn` (e.g. torch.distributed all_gather). */
int reo_set_collective(reo_handle_t h, int rank, int world, reo_allgather_fn fn, void* ctx);

/* One process per GPU with NCCL inside the library: rank 0 calls reo_comm_unique_id (128 bytes), every
 * rank receives a copy over any host channel and calls reo_comm_init_rank.  Tables are then all-gathered
 * with ncclAllGather on the handle's own stream (no host synchronisation, NVLink/NVSwitch transport). */
int reo_comm_unique_id(void* out128);
int reo_comm_init_rank(reo_handle_t h, int rank, int world, const void* id128);

/* Page-locked host memory for REO_OUT_PINNED outputs (and for inputs: pinned inputs are copied to the device at
 * full PCIe rate).  NULL on failure.  The memory outlives handles; free it with reo_host_free. */
void* reo_host_alloc(size_t bytes);
void reo_host_free(void* p);

/* get_major_reo_lower_count(sample_size, pval_threshold), src:81-92.  Host arithmetic. */
int reo_threshold(int sample_size, double pval_reo);

/*
 * identify_degs, src:339-438: the whole path.  result: column-major r x 15 x K doubles (a Julia
 * Array{Float64,3}(r, 15, K)): result[i + r*col + r*15*k], the 15 columns of src:398/405/665
 * (pval padj n11 n12 n13 n21 n22 n23 n31 n32 n33 d1 d2 se z1).
 * updown: K x r (+1 "up", -1 "down", 0 "no change", src:426-429).  final_ref: K x r, the mask
 * used by the last evaluation (may be NULL).  iters_done: K ints (may be NULL).  stats may be NULL.
 */
int reo_identify_degs(reo_handle_t h, const void* data, int dtype, int64_t r, int64_t c, int64_t ld,
                      const int32_t* group_id, int32_t gnum, const int32_t* thresholds, double pval_reo,
                      double pval_deg, double padj_deg, const uint8_t* ref_mask, int32_t n_iter, int32_t n_conv,
                      uint32_t flags, double* result, int8_t* updown, uint8_t* final_ref, int32_t* iters_done,
                      reo_stats* stats);

/* Iteration log of level k (0 <= k < K) of the last reo_identify_degs on this handle: evaluations performed,
 * whether the n_conv criterion stopped the loop (src:419-422), and per evaluation the number of DEGs (src:418) and
 * the size of the reference set it used; at most `cap` entries are written.  Lets the Julia shim print the
 * reference's per-level @info lines (src:418-420, 432-435).  Any output pointer may be NULL. */
int reo_iter_log(reo_handle_t h, int32_t k, int32_t* iters_done, int32_t* converged, int32_t* n_deg, int32_t* n_ref,
                 int32_t cap);

/* Test hook, host only (no GPU): the (row tile, column tile, update flags) triples rank `rank` of `world` evaluates for
 * a table build over `ncols` of `r` genes, replayed from the library's own partition code (csrc/reo_pairs2.cu).  See the
 * definition in csrc/reo_api.cu; used by tests/test_dist_gloo.py to check the multi-rank partition on CPU. */
long long reo_debug_pair_plan(int64_t r, int64_t ncols, int32_t sample_words, int32_t planes, int32_t rank, int32_t world,
                              int32_t mode, int32_t* out, int64_t cap, int32_t* nsym_tiles, int32_t* ntr, int32_t* ntc);

/* ---- stage-level entry points (parity tests, benches, profilers) ------------------------- */

/* K1: copy, dense-rank per sample, bit-slice, coin plane.  The staged matrix stays resident in
 * the handle until the next reo_stage / reo_identify_degs / reo_destroy.  src:351-362 + 372. */
int reo_stage(reo_handle_t h, const void* data, int dtype, int64_t r, int64_t c, int64_t ld,
              const int32_t* group_id, int32_t gnum, uint32_t flags);
int reo_stage_info(reo_handle_t h, int32_t* rank_bits, int32_t* sample_words, int32_t* gene_tiles);

/* src:372-374 for a small block: for level k, nre[a*ncols+b] = # level-k samples where gene
 * rows[a] "is greater" than gene cols[b], rest[...] likewise over all other samples.        */
int reo_pair_counts(reo_handle_t h, int32_t k, const int32_t* rows, int32_t nrows, const int32_t* cols,
                    int32_t ncols, int32_t* nre, int32_t* rest);

/* src:366-392 + 403: 3x3 reversal tables of every gene against the genes with mask != 0, level k;
 * table: r x 9 Int32 row-major (n11 n12 n13 n21 ... n33).  thresholds as above (NULL -> pval_reo). */
int reo_tables(reo_handle_t h, int32_t k, const int32_t* thresholds, double pval_reo, const uint8_t* mask,
               int32_t* table);
/* Same result through the incremental path: build for mask_from, then apply the signed update
 * over the symmetric difference to reach mask_to. */
int reo_tables_delta(reo_handle_t h, int32_t k, const int32_t* thresholds, double pval_reo,
                     const uint8_t* mask_from, const uint8_t* mask_to, int32_t* table);

/* McCullagh_test, src:225-259, on the device: n tables of k x k Int64 (row-major) ->
 * n x 5 doubles (pval d1 d2 se z1). */
int reo_mccullagh(reo_handle_t h, const int64_t* tables, int64_t n, int32_t k, double* out);

/* src:409-412: sort, trimmed std, two-sided normal p.  n > 10. */
int reo_empirical_null(reo_handle_t h, const double* delta1, int64_t n, double* pval, double* se);
/* src:413: Benjamini-Hochberg adjustment on the device. */
int reo_bh(reo_handle_t h, const double* p, int64_t n, double* padj);
/* ascending stable sort of doubles on the device (used by the two above); perm may be NULL. */
int reo_sort_f64(reo_handle_t h, const double* x, int64_t n, double* sorted, int32_t* perm);

/* ---- the steps right before the path, HBM-bound streaming kernels ------- */

/* pseudobulk_group, src:56-67 / 608-612: out[:, p] = sum of data[:, s] over s in cell_list[cell_ptr[p] .. cell_ptr[p+1])
 * in that order (the random shuffle + partition of src:62 stays a host decision).  out: r x nprofiles column-major,
 * Int64 for integer input, Float64 for float input.  out_host may be NULL; *out_dev (may be NULL) receives a
 * handle-owned device pointer valid until the next reo_pseudobulk -- pass it to reo_identify_degs with
 * REO_DATA_ON_DEVICE so the pseudo-bulk matrix never leaves HBM. */
int reo_pseudobulk(reo_handle_t h, const void* data, int dtype, int64_t r, int64_t c, int64_t ld, const int32_t* cell_ptr,
                   const int32_t* cell_list, int32_t nprofiles, uint32_t flags, void* out_host, void** out_dev);
/* src:618, 626: per_cell[s] = #{i : data[i,s] > 0}, per_gene[i] = #{s : data[i,s] > 0}. */
int reo_detect_counts(reo_handle_t h, const void* data, int dtype, int64_t r, int64_t c, int64_t ld, uint32_t flags,
                      int32_t* per_cell, int32_t* per_gene);
/* src:624-628: out = data[gene_list, cell_list] (r2 x c2 column-major, same dtype); out_host / out_dev as above. */
int reo_subset(reo_handle_t h, const void* data, int dtype, int64_t r, int64_t c, int64_t ld, const int32_t* gene_list,
               int64_t r2, const int32_t* cell_list, int64_t c2, uint32_t flags, void* out_host, void** out_dev);

#ifdef __cplusplus
}
#endif
#endif /* REO_H */
