// Synthetic inputs of the benchmark configurations: a counter-based splitmix64 stream, so that the
// device-resident input of a timed run and the host copy the CPU reference arm works on hold the SAME bytes
// (tests/test_synth.py restates the stream in numpy and compares).  Element i of a stream with seed s is a function
// of z = mix(s + i) only; a rank that owns elements [first, first + n) of a larger array passes first_index = first.
#include "nds_internal.h"

namespace nds {
namespace {

__host__ __device__ inline unsigned long long splitmix64(unsigned long long x) {
    unsigned long long z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <int KIND> struct SynthElem;
template <> struct SynthElem<NDS_SYNTH_F64_UNIFORM> {  // (z >> 11) * 2^-53: uniform [0, 1), D1 of SURVEY §8d
    using T = double;
    __device__ static T make(unsigned long long seed, unsigned long long i) {
        return (double)(splitmix64(seed + i) >> 11) * 0x1.0p-53;
    }
};
template <> struct SynthElem<NDS_SYNTH_F32_UNIFORM> {  // (z >> 40) * 2^-24
    using T = float;
    __device__ static T make(unsigned long long seed, unsigned long long i) {
        return (float)(splitmix64(seed + i) >> 40) * 0x1.0p-24f;
    }
};
template <> struct SynthElem<NDS_SYNTH_F32_UNIFORM_NAN10> {  // the same, NaN where mix(0x9E37 + seed + i) % 10 == 0
    using T = float;
    __device__ static T make(unsigned long long seed, unsigned long long i) {
        if (splitmix64(0x9E37ull + seed + i) % 10ull == 0ull) return __uint_as_float(0x7FC00000u);
        return (float)(splitmix64(seed + i) >> 40) * 0x1.0p-24f;
    }
};
template <> struct SynthElem<NDS_SYNTH_U32_16VALUES> {  // v[k] = k * 0x11111111, element = v[z & 15]
    using T = uint32_t;
    __device__ static T make(unsigned long long seed, unsigned long long i) {
        return (uint32_t)(splitmix64(seed + i) & 15ull) * 0x11111111u;
    }
};
template <> struct SynthElem<NDS_SYNTH_U32_16VALUES_ADV> {  // adversarial: v[k] = 0x80000000 + k (ties differ in the last digit)
    using T = uint32_t;
    __device__ static T make(unsigned long long seed, unsigned long long i) {
        return 0x80000000u + (uint32_t)(splitmix64(seed + i) & 15ull);
    }
};
template <> struct SynthElem<NDS_SYNTH_I64_16VALUES> {  // ((z & 15) - 8) * 2^40
    using T = int64_t;
    __device__ static T make(unsigned long long seed, unsigned long long i) {
        return ((int64_t)(splitmix64(seed + i) & 15ull) - 8) * ((int64_t)1 << 40);
    }
};

template <int KIND>
__global__ void __launch_bounds__(256) synth_fill_kernel(typename SynthElem<KIND>::T *dst, unsigned long long n,
                                                         unsigned long long seed, unsigned long long first) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = SynthElem<KIND>::make(seed, first + i);
}

template <int KIND> cudaError_t synth_launch(nds_ctx *ctx, void *dst, unsigned long long n, unsigned long long seed,
                                             unsigned long long first) {
    if (n == 0) return cudaSuccess;
    const unsigned long long blocks = (n + 255) / 256;
    const unsigned grid = (unsigned)(blocks < (unsigned long long)ctx->sm_count * 16 ? blocks : (unsigned long long)ctx->sm_count * 16);
    synth_fill_kernel<KIND><<<grid, 256, 0, ctx->stream>>>((typename SynthElem<KIND>::T *)dst, n, seed, first);
    ctx->launches += 1;
    return cudaGetLastError();
}

}  // namespace
}  // namespace nds

extern "C" nds_status nds_synth_fill(nds_ctx *ctx, nds_synth_kind kind, void *dst_device, size_t n, uint64_t seed,
                                     uint64_t first_index) {
    using namespace nds;
    if (!ctx) return NDS_BAD_ARGUMENT;
    if (n && !dst_device) return NDS_BAD_ARGUMENT;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return NDS_CUDA_ERROR;
    cudaError_t e;
    switch (kind) {
    case NDS_SYNTH_F64_UNIFORM: e = synth_launch<NDS_SYNTH_F64_UNIFORM>(ctx, dst_device, n, seed, first_index); break;
    case NDS_SYNTH_F32_UNIFORM: e = synth_launch<NDS_SYNTH_F32_UNIFORM>(ctx, dst_device, n, seed, first_index); break;
    case NDS_SYNTH_F32_UNIFORM_NAN10: e = synth_launch<NDS_SYNTH_F32_UNIFORM_NAN10>(ctx, dst_device, n, seed, first_index); break;
    case NDS_SYNTH_U32_16VALUES: e = synth_launch<NDS_SYNTH_U32_16VALUES>(ctx, dst_device, n, seed, first_index); break;
    case NDS_SYNTH_U32_16VALUES_ADV: e = synth_launch<NDS_SYNTH_U32_16VALUES_ADV>(ctx, dst_device, n, seed, first_index); break;
    case NDS_SYNTH_I64_16VALUES: e = synth_launch<NDS_SYNTH_I64_16VALUES>(ctx, dst_device, n, seed, first_index); break;
    default: ctx->last_error = "unknown synthetic stream"; return NDS_BAD_ARGUMENT;
    }
    if (e != cudaSuccess) {
        ctx->last_error = std::string("nds_synth_fill: ") + cudaGetErrorString(e);
        return NDS_CUDA_ERROR;
    }
    return NDS_OK;
}
