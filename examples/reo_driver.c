/*
 * examples/reo_driver.c -- plain-C caller of the libreo_cuda.so ABI (include/reo.h): what any FFI (Julia ccall,
 * ctypes, cgo ...) binds.  Also the natural target for ncu / cuda-gdb runs that should not involve Python.
 *
 *   gcc -O2 -Iinclude examples/reo_driver.c -o examples/reo_driver -Lrankcompv3.jl_b200 -lreo_cuda \
 *       -Wl,-rpath,$PWD/rankcompv3.jl_b200
 *   ./examples/reo_driver [genes] [n1] [n2]
 *
 * Builds a synthetic count matrix (column-major, Int64, like a Julia Matrix{Int64}), calls reo_identify_degs
 * (the replacement of identify_degs, src/RankCompV3.jl:339-438) and prints the log a Julia shim would print.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "reo.h"

static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 11); }

int main(int argc, char** argv) {
    const int64_t r = argc > 1 ? atoll(argv[1]) : 2000;
    const int64_t n1 = argc > 2 ? atoll(argv[2]) : 20, n2 = argc > 3 ? atoll(argv[3]) : 24, c = n1 + n2;
    int64_t* data = (int64_t*)malloc(sizeof(int64_t) * (size_t)r * (size_t)c);
    int32_t* gid = (int32_t*)malloc(sizeof(int32_t) * (size_t)c);
    uint8_t* ref = (uint8_t*)calloc((size_t)r, 1);
    for (int64_t s = 0; s < c; ++s) gid[s] = s < n1 ? 0 : 1;
    for (int64_t i = 0; i < r; ++i) {
        const uint32_t base = 1 + rnd() % 200;
        const int de = (i % 10) == 0;                    /* every 10th gene is 3x higher in group 2 */
        for (int64_t s = 0; s < c; ++s) data[i + r * s] = (int64_t)((rnd() % (2 * base)) * ((de && s >= n1) ? 3 : 1));
        ref[i] = (i % 10) == 5;                          /* reference genes: non-DE */
    }
    reo_handle_t h = NULL;
    int dev = 0;
    int rc = reo_create(&h, 1, &dev, 7, REO_FLAG_NONE);
    if (rc != REO_OK) { fprintf(stderr, "reo_create: %d %s\n", rc, reo_last_error(NULL)); return 1; }
    double* result = (double*)malloc(sizeof(double) * (size_t)r * 15);
    int8_t* updown = (int8_t*)malloc((size_t)r);
    uint8_t* final_ref = (uint8_t*)malloc((size_t)r);
    int32_t iters = 0;
    reo_stats st;
    rc = reo_identify_degs(h, data, REO_I64, r, c, r, gid, 2, NULL, 0.01, 1.0, 0.05, ref, 128, 5, 0, result, updown,
                           final_ref, &iters, &st);
    if (rc != REO_OK) { fprintf(stderr, "reo_identify_degs: %d %s\n", rc, reo_last_error(h)); return 1; }
    for (int e = 0; e < st.iters_done && e < REO_MAX_ITER_LOG; ++e)
        printf("INFO: iteration %d,  # DEGs %d, # non-DEGs %lld\n", e, st.n_deg[e], (long long)(r - st.n_deg[e]));
    if (st.converged) printf("INFO: Convergence threshold is reached\n");
    int up = 0, down = 0, hit = 0;
    for (int64_t i = 0; i < r; ++i) { up += updown[i] == 1; down += updown[i] == -1; hit += (updown[i] != 0) && (i % 10 == 0); }
    printf("genes %lld samples %lld evaluations %d up %d down %d planted-and-called %d\n", (long long)r, (long long)c,
           iters, up, down, hit);
    printf("rank bits %d, sample words %d, comparisons %.3e, device ms: stage %.3f pairs %.3f total %.3f\n",
           st.rank_bits, st.sample_words, (double)st.compares, st.ms_stage, st.ms_pairs, st.ms_total);
    reo_destroy(h);
    free(data); free(gid); free(ref); free(result); free(updown); free(final_ref);
    return 0;
}
