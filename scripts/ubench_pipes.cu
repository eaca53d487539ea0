// Micro-benchmark: issue rates of LOP3, POPC and IMAD on sm_100a and how they share pipes (evidence for DESIGN.md 10).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/ubench_pipes scripts/ubench_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096
#define U 8

template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned* out, unsigned seed) {
    unsigned x[U], y[U], acc[U];
#pragma unroll
    for (int i = 0; i < U; ++i) { x[i] = seed * (threadIdx.x + 1) + i; y[i] = seed ^ (i * 0x9E3779B1u); acc[i] = i; }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            if (MODE == 0 || MODE == 3 || MODE == 4)      // LOP3 (dependent chain per i, U chains in flight)
                asm volatile("lop3.b32 %0, %1, %2, %0, 0xB2;" : "+r"(acc[i]) : "r"(x[i]), "r"(y[i]));
            if (MODE == 3) {                              // 13 more LOP3: the pair kernel's ratio 14 : 1 : 1
#pragma unroll
                for (int q = 0; q < 13; ++q) asm volatile("lop3.b32 %0, %1, %2, %0, 0xB2;" : "+r"(acc[i]) : "r"(x[i]), "r"(y[i]));
            }
            if (MODE == 1)                                // POPC only (dependent chain per i)
                asm volatile("popc.b32 %0, %0;" : "+r"(acc[i]));
            if (MODE == 3 || MODE == 5) {                 // POPC feeding an IMAD accumulate, as in the pair kernel
                unsigned c;
                asm volatile("popc.b32 %0, %1;" : "=r"(c) : "r"(MODE == 3 ? acc[i] : x[i]));
                asm volatile("mad.lo.u32 %0, %1, 4, %0;" : "+r"(x[i]) : "r"(c));
            }
            if (MODE == 2 || MODE == 4)                   // IMAD
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(y[i]), "r"(acc[i]));
        }
    }
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < U; ++i) r ^= acc[i] ^ x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, double instr_per_iter, unsigned* out, int sms, double mhz) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = sms * 8;
    k<MODE><<<grid, 256>>>(out, 12345u); cudaDeviceSynchronize();
    cudaEventRecord(a); k<MODE><<<grid, 256>>>(out, 12345u); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double thread_instr = (double)grid * 256 * ITER * U * instr_per_iter;
    const double per_clk_sm = thread_instr / (ms * 1e-3) / (mhz * 1e6) / sms;
    printf("%-34s %8.3f ms  %7.1f thread-instr/clk/SM\n", name, ms, per_clk_sm);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    unsigned* out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    printf("%s, %d SMs, %.0f MHz (attribute; rates assume this clock)\n", p.name, p.multiProcessorCount, mhz);
    run<0>("LOP3 only", 1, out, p.multiProcessorCount, mhz);
    run<1>("POPC only", 1, out, p.multiProcessorCount, mhz);
    run<2>("IMAD only", 1, out, p.multiProcessorCount, mhz);
    run<4>("LOP3 + IMAD (1:1)", 2, out, p.multiProcessorCount, mhz);
    run<5>("POPC + IMAD (1:1)", 2, out, p.multiProcessorCount, mhz);
    run<3>("14 LOP3 + POPC + IMAD (kernel mix)", 16, out, p.multiProcessorCount, mhz);
    return 0;
}
