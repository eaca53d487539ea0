// Micro-benchmark of the pair kernel's steady-state inner loop WITHOUT any pipeline / barrier around it: a 4 x NBT
// register tile of borrow chains over NP planes read from shared memory (LDS.128, one plane ahead), POPC + IMAD per
// word.  Reports LOP3 lane-ops per clock per SM (the ALU pipe issues 64) against resident warps per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/ubench_chain scripts/ubench_chain.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint4 lds_v4(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lop3_b2(uint32_t x, uint32_t y, uint32_t c) { uint32_t d; asm("lop3.b32 %0, %1, %2, %3, 0xB2;" : "=r"(d) : "r"(x), "r"(y), "r"(c)); return d; }
__device__ __forceinline__ uint32_t mad_acc(uint32_t pc, uint32_t k, uint32_t acc) { uint32_t d; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(pc), "r"(k), "r"(acc)); return d; }

// PF: operands loaded one plane ahead; PACK: two 16-bit counters per register (NBT == 8)
template <int NBT, int NP, bool PF, bool PACK>
__global__ void __launch_bounds__(256) chain_kernel(uint32_t* out, int words, uint32_t one) {
    extern __shared__ __align__(16) uint32_t sm[];   // [words_in_smem][3][NP][64]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 8 * 3 * NP * 64; i += 256) sm[i] = i * 2654435761u;
    __syncthreads();
    const int ty = (warp >> 2) * 8 + (lane >> 2), tx = (warp & 3) * 4 + (lane & 3);
    constexpr int NACC = PACK ? NBT / 2 : NBT;
    uint32_t acc[4][NACC];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < NACC; ++b) acc[a][b] = 0;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    const uint32_t k1 = PACK ? one : one << 2, k2 = PACK ? one << 16 : k1;
    for (int w = 0; w < words; ++w) {
        const uint32_t wa = base + (uint32_t)((w & 7) * 3 * NP * 256);
        const uint32_t xa = wa + ty * 16, ya = wa + NP * 256 + tx * 16, za = wa + 2 * NP * 256 + tx * 16;
        uint32_t bor[4][NBT];
        uint4 xn = lds_v4(xa), yn = lds_v4(ya), zn = NBT == 8 ? lds_v4(za) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int pl = 0; pl < NP; ++pl) {
            uint4 xv, yv, zv;
            if (PF) {
                xv = xn; yv = yn; zv = zn;
                if (pl + 1 < NP) { xn = lds_v4(xa + (pl + 1) * 256); yn = lds_v4(ya + (pl + 1) * 256); if (NBT == 8) zn = lds_v4(za + (pl + 1) * 256); }
            } else {
                xv = lds_v4(xa + pl * 256); yv = lds_v4(ya + pl * 256); if (NBT == 8) zv = lds_v4(za + pl * 256);
            }
            const uint32_t x[4] = {xv.x, xv.y, xv.z, xv.w};
            const uint32_t y[8] = {yv.x, yv.y, yv.z, yv.w, zv.x, zv.y, zv.z, zv.w};
#pragma unroll
            for (int b = 0; b < NBT; ++b)
#pragma unroll
                for (int a = 0; a < 4; ++a) bor[a][b] = lop3_b2(x[a], y[b], pl == 0 ? one : bor[a][b]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < NBT; ++b) acc[a][b % NACC] = mad_acc(__popc(bor[a][b]), (PACK && b >= NACC) ? k2 : k1, acc[a][b % NACC]);
    }
    uint32_t r = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < NACC; ++b) r ^= acc[a][b];
    out[blockIdx.x * 256 + tid] = r;
}

template <int NBT, int NP, bool PF, bool PACK>
void run(const char* name, int ctas_per_sm, uint32_t* out, int sms, double mhz) {
    auto kern = chain_kernel<NBT, NP, PF, PACK>;
    // shared memory sized so that exactly ctas_per_sm CTAs are resident
    const int need = 8 * 3 * NP * 256;                       // the kernel's own footprint
    const int smem = (227 * 1024) / ctas_per_sm - 1024;
    if (smem < need) { printf("%-28s needs %d KB of shared memory: %d CTAs/SM do not fit\n", name, need / 1024, ctas_per_sm); return; }
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem);
    if (occ < 1) { printf("%-28s does not fit\n", name); return; }
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    const int words = 20000, grid = sms * occ;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    kern<<<grid, 256, smem>>>(out, words / 10, 1u); cudaDeviceSynchronize();
    cudaEventRecord(a); kern<<<grid, 256, smem>>>(out, words, 1u); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double lop3 = (double)grid * 256 * words * NP * 4 * NBT;
    printf("%-28s regs %3d  CTAs/SM %d (warps/SMSP %d)  %8.3f ms  LOP3 %5.1f lanes/clk/SM = %.3f of 64\n", name, fa.numRegs, occ, occ * 2,
           ms, lop3 / (ms * 1e-3) / (mhz * 1e6) / sms, lop3 / (ms * 1e-3) / (mhz * 1e6) / sms / 64.0);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    uint32_t* out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    printf("%s, %d SMs, %.0f MHz\n", p.name, p.multiProcessorCount, mhz);
    for (int c = 1; c <= 4; ++c) run<8, 8, true, true>("4x8 NP=8 prefetch packed", c, out, p.multiProcessorCount, mhz);
    for (int c = 1; c <= 3; ++c) run<8, 8, true, false>("4x8 NP=8 prefetch wide", c, out, p.multiProcessorCount, mhz);
    for (int c = 1; c <= 3; ++c) run<8, 8, false, true>("4x8 NP=8 no-prefetch packed", c, out, p.multiProcessorCount, mhz);
    for (int c = 1; c <= 4; ++c) run<4, 8, true, false>("4x4 NP=8 prefetch wide", c, out, p.multiProcessorCount, mhz);
    for (int c = 1; c <= 3; ++c) run<8, 13, true, true>("4x8 NP=13 prefetch packed", c, out, p.multiProcessorCount, mhz);
    for (int c = 1; c <= 4; ++c) run<4, 13, true, false>("4x4 NP=13 prefetch wide", c, out, p.multiProcessorCount, mhz);
    for (int c = 1; c <= 3; ++c) run<8, 4, true, true>("4x8 NP=4 prefetch packed", c, out, p.multiProcessorCount, mhz);
    return 0;
}
