// Read-only HBM bandwidth probes (B200): how fast can ONE pass over a large array be read, by which mechanism?
//   ldg     grid-stride 16-byte loads, unrolled (LDG.E.128)
//   bulk    1-D TMA bulk copies (cp.async.bulk, UBLKCP) into a per-warp shared-memory ring, data discarded
//   ldgsts  16-byte cp.async (LDGSTS) into a per-warp ring, data discarded
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bw_probe bw_probe.cu ; run: ./bw_probe [GiB]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(512) k_ldg(const uint4 *__restrict__ p, size_t n16, unsigned *sink) {
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n16; i += 8 * stride) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    for (; i < n16; i += stride) { uint4 v = __ldg(p + i); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) *sink = acc;
}

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(s32(bar)), "r"(parity) : "memory");
}

// every warp streams its own contiguous range through S stages of B bytes
template <int STAGES>
__global__ void __launch_bounds__(512) k_bulk(const unsigned char *__restrict__ p, size_t bytes, unsigned blk, unsigned *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    unsigned char *ring = smem + (size_t)warp * STAGES * blk;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)W * STAGES * blk) + warp * STAGES;
    if (lane == 0) for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bars + s)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const size_t nblk = bytes / blk, nw = (size_t)gridDim.x * W, me = (size_t)blockIdx.x * W + warp;
    size_t nxt = me;
    unsigned acc = 0;
    auto issue = [&](int s) {
        if (nxt < nblk && lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bars + s)), "r"(blk) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(ring + (size_t)s * blk)), "l"(p + nxt * blk), "r"(blk), "r"(s32(bars + s)) : "memory");
        }
        nxt += nw;
    };
    for (int s = 0; s < STAGES; ++s) issue(s);
    unsigned par = 0;
    int s = 0;
    for (size_t b = me; b < nblk; b += nw) {
        mbar_wait(bars + s, (par >> s) & 1u);
        par ^= 1u << s;
        acc ^= *reinterpret_cast<volatile unsigned *>(ring + (size_t)s * blk + lane * 4);
        __syncwarp();
        issue(s);
        s = (s + 1 == STAGES) ? 0 : s + 1;
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int STAGES>
__global__ void __launch_bounds__(512) k_ldgsts(const unsigned char *__restrict__ p, size_t bytes, unsigned *sink) {
    constexpr unsigned BLK = 4096;  // 32 lanes x 8 x 16 B
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    unsigned char *ring = smem + (size_t)warp * STAGES * BLK;
    const size_t nblk = bytes / BLK, nw = (size_t)gridDim.x * W, me = (size_t)blockIdx.x * W + warp;
    size_t nxt = me;
    unsigned acc = 0;
    auto issue = [&](int s) {
        if (nxt < nblk) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(ring + (size_t)s * BLK + (j * 32 + lane) * 16)), "l"(p + nxt * BLK + (j * 32 + lane) * 16) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        nxt += nw;
    };
    for (int s = 0; s < STAGES; ++s) issue(s);
    int s = 0;
    for (size_t b = me; b < nblk; b += nw) {
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
        __syncwarp();
        acc ^= *reinterpret_cast<volatile unsigned *>(ring + (size_t)s * BLK + lane * 4);
        __syncwarp();
        issue(s);
        s = (s + 1 == STAGES) ? 0 : s + 1;
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <typename F> float timeit(F f, int reps = 10) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char **argv) {
    const double gib = argc > 1 ? atof(argv[1]) : 2.0;
    const size_t bytes = (size_t)(gib * (1ull << 30)) / 65536 * 65536;
    unsigned char *d; unsigned *sink;
    CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(d, 1, bytes));
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    printf("bytes %zu  SMs %d\n", bytes, sm);
    for (int mult : {1, 2, 4, 8}) {
        float ms = timeit([&] { k_ldg<<<sm * mult, 512>>>((const uint4 *)d, bytes / 16, sink); });
        printf("ldg      grid %4d x 512            %.4f ms  %.0f GB/s\n", sm * mult, ms, bytes / ms / 1e6);
    }
    auto bulk = [&](auto kern, int stages, unsigned blk, int warps) {
        size_t smem = (size_t)warps * stages * blk + (size_t)warps * stages * 8 + 128;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        float ms = timeit([&] { kern<<<sm, warps * 32, smem>>>(d, bytes, blk, sink); });
        printf("bulk     %2d warps x %d stages x %5u B  (%3zu KB in flight/SM)  %.4f ms  %.0f GB/s\n", warps, stages, blk, (size_t)warps * stages * blk / 1024, ms, bytes / ms / 1e6);
    };
    bulk(k_bulk<1>, 1, 4096, 16);
    bulk(k_bulk<2>, 2, 4096, 16);
    bulk(k_bulk<3>, 3, 4096, 16);
    bulk(k_bulk<2>, 2, 8192, 8);
    bulk(k_bulk<3>, 3, 8192, 8);
    bulk(k_bulk<4>, 4, 8192, 4);
    bulk(k_bulk<2>, 2, 16384, 4);
    bulk(k_bulk<4>, 4, 16384, 2);
    bulk(k_bulk<4>, 4, 2048, 16);
    auto lg = [&](auto kern, int stages, int warps) {
        size_t smem = (size_t)warps * stages * 4096 + 128;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        float ms = timeit([&] { kern<<<sm, warps * 32, smem>>>(d, bytes, sink); });
        printf("ldgsts   %2d warps x %d stages x  4096 B  (%3zu KB in flight/SM)  %.4f ms  %.0f GB/s\n", warps, stages, (size_t)warps * stages * 4, ms, bytes / ms / 1e6);
    };
    lg(k_ldgsts<1>, 1, 16);
    lg(k_ldgsts<2>, 2, 16);
    lg(k_ldgsts<3>, 3, 16);
    lg(k_ldgsts<4>, 4, 8);
    return 0;
}
