// reo_prep.cu -- the steps immediately before the REO core in reoa(), as HBM-bound
// streaming kernels over the raw column-major expression matrix:
//   pseudobulk_kernel     src/RankCompV3.jl:56-67, 608-612: row sums over the cells of each pseudo-bulk profile
//                         (the random shuffle/partition of src:62 is a host decision and arrives as a cell -> profile map)
//   detect_counts_kernel  src:618, 626: number of detected (> 0) genes per cell and of detecting cells per gene
//   subset_kernel         src:624-628: gather of the kept genes x kept cells into a compact column-major matrix
// Algorithmic bytes: one read of the r x c matrix (+ the small outputs).
#include <algorithm>

#include "reo_internal.cuh"

template <typename T> struct AccT { typedef long long type; };
template <> struct AccT<double> { typedef double type; };
template <> struct AccT<float> { typedef double type; };

// grid (ceil(r/256), nprofiles); cells of profile p are cell_list[cell_ptr[p] .. cell_ptr[p+1])
template <typename T>
__global__ void __launch_bounds__(256)
pseudobulk_kernel(const T* __restrict__ data, int64_t r, int64_t ld, const int32_t* __restrict__ cell_ptr,
                  const int32_t* __restrict__ cell_list, int nprofiles, typename AccT<T>::type* __restrict__ out) {
    typedef typename AccT<T>::type A;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= r) return;
    for (int p = blockIdx.y; p < nprofiles; p += gridDim.y) {   // gridDim.y is capped at 65535
    A acc = 0;
    const int b = cell_ptr[p], e = cell_ptr[p + 1];
    int k = b;
    for (; k + 4 <= e; k += 4) {  // 4 independent coalesced loads in flight
        const A v0 = (A)data[(int64_t)cell_list[k] * ld + g], v1 = (A)data[(int64_t)cell_list[k + 1] * ld + g];
        const A v2 = (A)data[(int64_t)cell_list[k + 2] * ld + g], v3 = (A)data[(int64_t)cell_list[k + 3] * ld + g];
        acc = acc + v0; acc = acc + v1; acc = acc + v2; acc = acc + v3;   // left-to-right, like sum() over the cells
    }
    for (; k < e; ++k) acc = acc + (A)data[(int64_t)cell_list[k] * ld + g];
    out[(int64_t)p * r + g] = acc;
    }
}

cudaError_t reo_launch_pseudobulk(const void* data, int dtype, int64_t r, int64_t ld, const int32_t* cell_ptr,
                                  const int32_t* cell_list, int nprofiles, void* out, cudaStream_t st) {
    if (nprofiles <= 0 || r <= 0) return cudaSuccess;
    dim3 grid((unsigned)((r + 255) / 256), (unsigned)std::min(nprofiles, 65535));
    switch (dtype) {
        case REO_I64: pseudobulk_kernel<long long><<<grid, 256, 0, st>>>((const long long*)data, r, ld, cell_ptr, cell_list, nprofiles, (long long*)out); break;
        case REO_I32: pseudobulk_kernel<int><<<grid, 256, 0, st>>>((const int*)data, r, ld, cell_ptr, cell_list, nprofiles, (long long*)out); break;
        case REO_F64: pseudobulk_kernel<double><<<grid, 256, 0, st>>>((const double*)data, r, ld, cell_ptr, cell_list, nprofiles, (double*)out); break;
        case REO_F32: pseudobulk_kernel<float><<<grid, 256, 0, st>>>((const float*)data, r, ld, cell_ptr, cell_list, nprofiles, (double*)out); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// grid (ceil(r/256), ceil(c/32)): each thread walks 32 cells of its gene; per-cell counts through a block reduction
template <typename T>
__global__ void __launch_bounds__(256)
detect_counts_kernel(const T* __restrict__ data, int64_t r, int64_t c, int64_t ld, int32_t* __restrict__ per_cell,
                     int32_t* __restrict__ per_gene) {
    __shared__ int cell_cnt[32];
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int mine = 0;
    const unsigned lane = threadIdx.x & 31;
    for (int64_t s0 = (int64_t)blockIdx.y * 32; s0 < c; s0 += (int64_t)gridDim.y * 32) {   // gridDim.y <= 65535
    __syncthreads();
    if (threadIdx.x < 32) cell_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int k = 0; k < 32; ++k) {
        const int64_t s = s0 + k;
        const bool det = (s < c && g < r) ? (data[s * ld + g] > (T)0) : false;
        mine += det;
        const unsigned m = __ballot_sync(0xffffffffu, det);
        if (lane == 0 && m) atomicAdd(&cell_cnt[k], __popc(m));
    }
    __syncthreads();
    if (threadIdx.x < 32 && s0 + threadIdx.x < c && cell_cnt[threadIdx.x]) atomicAdd(&per_cell[s0 + threadIdx.x], cell_cnt[threadIdx.x]);
    }
    if (g < r && mine) atomicAdd(&per_gene[g], mine);
}

cudaError_t reo_launch_detect_counts(const void* data, int dtype, int64_t r, int64_t c, int64_t ld, int32_t* per_cell,
                                     int32_t* per_gene, cudaStream_t st) {
    dim3 grid((unsigned)((r + 255) / 256), (unsigned)std::min<int64_t>((c + 31) / 32, 65535));
    switch (dtype) {
        case REO_I64: detect_counts_kernel<long long><<<grid, 256, 0, st>>>((const long long*)data, r, c, ld, per_cell, per_gene); break;
        case REO_I32: detect_counts_kernel<int><<<grid, 256, 0, st>>>((const int*)data, r, c, ld, per_cell, per_gene); break;
        case REO_F64: detect_counts_kernel<double><<<grid, 256, 0, st>>>((const double*)data, r, c, ld, per_cell, per_gene); break;
        case REO_F32: detect_counts_kernel<float><<<grid, 256, 0, st>>>((const float*)data, r, c, ld, per_cell, per_gene); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// out[i + r2 * s] = data[gene_list[i] + ld * cell_list[s]]
template <typename T>
__global__ void __launch_bounds__(256)
subset_kernel(const T* __restrict__ data, int64_t ld, const int32_t* __restrict__ gene_list, int64_t r2,
              const int32_t* __restrict__ cell_list, int64_t c2, T* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r2) return;
    const int64_t gsrc = gene_list[i];
    for (int64_t s = blockIdx.y; s < c2; s += gridDim.y)        // gridDim.y is capped at 65535, c2 is not
        out[i + r2 * s] = data[gsrc + ld * (int64_t)cell_list[s]];
}

cudaError_t reo_launch_subset(const void* data, int dtype, int64_t ld, const int32_t* gene_list, int64_t r2,
                              const int32_t* cell_list, int64_t c2, void* out, cudaStream_t st) {
    if (r2 <= 0 || c2 <= 0) return cudaSuccess;
    dim3 grid((unsigned)((r2 + 255) / 256), (unsigned)std::min<int64_t>(c2, 65535));
    switch (dtype) {
        case REO_I64: subset_kernel<long long><<<grid, 256, 0, st>>>((const long long*)data, ld, gene_list, r2, cell_list, c2, (long long*)out); break;
        case REO_I32: subset_kernel<int><<<grid, 256, 0, st>>>((const int*)data, ld, gene_list, r2, cell_list, c2, (int*)out); break;
        case REO_F64: subset_kernel<double><<<grid, 256, 0, st>>>((const double*)data, ld, gene_list, r2, cell_list, c2, (double*)out); break;
        case REO_F32: subset_kernel<float><<<grid, 256, 0, st>>>((const float*)data, ld, gene_list, r2, cell_list, c2, (float*)out); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
