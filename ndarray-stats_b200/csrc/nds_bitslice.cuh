// Device helpers shared by the bit-sliced selection kernels (nds_short.cu, nds_cols.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace nds {
namespace {

// (a & mask) | (b & ~mask) as ONE LOP3 (the compiler otherwise emits two, one per constant)
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b, uint32_t mask) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(mask));
    return d;
}

// 16 words holding 32 16-bit fields (word k: field of slot 2k in the low half, slot 2k+1 in the high half)
// -> 16 plane words: word b, bit p = bit b of the field of slot 2*(p&15) + (p>>4).
// Butterfly levels 8 (byte permutes), 4, 2, 1 (shift + bit-select): 16 + 3 * 32 = 112 instructions.
__device__ __forceinline__ void transpose16(uint32_t (&w)[16]) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        uint32_t a = w[r], b = w[r + 8];
        w[r] = __byte_perm(a, b, 0x6240);
        w[r + 8] = __byte_perm(a, b, 0x7351);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        if (r & 4) continue;
        uint32_t a = w[r], b = w[r + 4];
        w[r] = bitsel(a, b << 4, 0x0F0F0F0Fu);
        w[r + 4] = bitsel(a >> 4, b, 0x0F0F0F0Fu);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        if (r & 2) continue;
        uint32_t a = w[r], b = w[r + 2];
        w[r] = bitsel(a, b << 2, 0x33333333u);
        w[r + 2] = bitsel(a >> 2, b, 0x33333333u);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        if (r & 1) continue;
        uint32_t a = w[r], b = w[r + 1];
        w[r] = bitsel(a, b << 1, 0x55555555u);
        w[r + 1] = bitsel(a >> 1, b, 0x55555555u);
    }
}

template <typename T> struct RawBits;
template <> struct RawBits<float> { using U = uint32_t; };
template <> struct RawBits<double> { using U = uint64_t; };
template <> struct RawBits<int32_t> { using U = uint32_t; };
template <> struct RawBits<uint32_t> { using U = uint32_t; };
template <> struct RawBits<int64_t> { using U = uint64_t; };
template <> struct RawBits<uint64_t> { using U = uint64_t; };

template <typename T> __device__ __forceinline__ T from_raw(typename RawBits<T>::U r) {
    T x;
    memcpy(&x, &r, sizeof x);
    return x;
}
template <typename T> __device__ __forceinline__ typename RawBits<T>::U to_raw(T x) {
    typename RawBits<T>::U r;
    memcpy(&r, &x, sizeof x);
    return r;
}


}  // namespace
}  // namespace nds
