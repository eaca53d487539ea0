// reo_host.cpp -- host-side helpers of the pageable-input pipeline (plain C++, built by g++; no CUDA here).
//
// A pageable input matrix (a Julia Matrix) has to be touched by the CPU anyway on its way into the pinned bounce buffers
// (reo_api.cu, do_stage).  Count matrices hold small non-negative integers in 8-byte elements, so the copy NARROWS them
// to u16 on the fly: the host writes a quarter of the bytes, the DMA engine reads a quarter from host memory and PCIe
// carries a quarter.  A slice with any value outside 0..65535 (or any non-integral value) reports failure and the chunk
// is copied raw instead; the device ranks u16 slices exactly as it ranks the original type (dense ranks depend on the
// values only).  Measured on the B200 host (16 cores, 8 copy threads): memcpy 52.8 GB/s of input, narrowing Int64 72.8.
#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define REO_CLONES __attribute__((target_clones("avx2", "default")))
#else
#define REO_CLONES
#endif

extern "C" {

REO_CLONES int reo_host_narrow_i64(const int64_t* src, uint16_t* dst, size_t n) {
    uint64_t acc = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint64_t v = (uint64_t)src[i];
        acc |= v;
        dst[i] = (uint16_t)v;
    }
    return (acc >> 16) == 0;
}

REO_CLONES int reo_host_narrow_i32(const int32_t* src, uint16_t* dst, size_t n) {
    uint32_t acc = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t v = (uint32_t)src[i];
        acc |= v;
        dst[i] = (uint16_t)v;
    }
    return (acc >> 16) == 0;
}

// v + 2^52 holds v in the low mantissa bits exactly when v is an integer in [0, 2^51); anything else (fraction, negative,
// huge, NaN, infinity) fails the round trip or leaves high bits set
REO_CLONES int reo_host_narrow_f64(const double* src, uint16_t* dst, size_t n) {
    uint64_t acc = 0;
    for (size_t i = 0; i < n; ++i) {
        const double v = src[i];
        const double t = v + 4503599627370496.0;
        uint64_t b;
        memcpy(&b, &t, 8);
        const uint64_t k = b - 0x4330000000000000ull;
        acc |= k | ((t - 4503599627370496.0) != v ? ~0ull : 0ull);
        dst[i] = (uint16_t)k;
    }
    return (acc >> 16) == 0;
}

REO_CLONES int reo_host_narrow_f32(const float* src, uint16_t* dst, size_t n) {
    uint32_t acc = 0;
    for (size_t i = 0; i < n; ++i) {
        const float v = src[i];
        const float t = v + 8388608.0f;
        uint32_t b;
        memcpy(&b, &t, 4);
        const uint32_t k = b - 0x4B000000u;
        acc |= k | ((t - 8388608.0f) != v ? ~0u : 0u);
        dst[i] = (uint16_t)k;
    }
    return (acc >> 16) == 0;
}

}  // extern "C"
