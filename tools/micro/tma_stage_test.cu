// Stand-alone repro of the per-warp double-buffered TMA ray staging (wavefront.cuh phase_trace_staged). nvcc -arch=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
constexpr int R = 32, BUF = R * 36;
template <int MODE>
__global__ void k(const float4* org, const float4* dir, const int* flags, int n_primary, const float4* sorg, const float4* sdir, int n_shadow, int* fetch, float* out) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[16 * 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* stage = smem + warp * 2 * BUF;
    const uint32_t bar0 = smem_u32(bars + warp * 2);
    const int n_items = n_primary + n_shadow;
    auto issue = [&](int k_, int s0) {
        if (lane != 0) return;
        unsigned char* buf = stage + k_ * BUF;
        const uint32_t bar = bar0 + 8u * k_;
        const int s1 = min(s0 + R, n_items), p1 = min(s1, n_primary);
        const int np = max(p1 - s0, 0), h0 = max(s0, n_primary), ns = max(s1 - h0, 0), npf = (np + 3) & ~3;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(np * 32 + (MODE & 1 ? 0 : npf * 4) + ns * 32)) : "memory");
        if (np > 0) { bulk_g2s(buf, org + s0, np * 16u, bar); bulk_g2s(buf + R * 16, dir + s0, np * 16u, bar); if (!(MODE & 1)) bulk_g2s(buf + R * 32, flags + s0, npf * 4u, bar); }
        if (ns > 0) { bulk_g2s(buf + (h0 - s0) * 16, sorg + (h0 - n_primary), ns * 16u, bar); bulk_g2s(buf + R * 16 + (h0 - s0) * 16, sdir + (h0 - n_primary), ns * 16u, bar); }
    };
    int nb0 = 0, pend = 0, cbuf = 1; unsigned parity = 0;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        nb0 = atomicAdd(fetch, R);
    }
    __syncwarp();
    nb0 = __shfl_sync(0xffffffffu, nb0, 0);
    if (nb0 < n_items) { issue(0, nb0); if (lane == 0) pend = atomicAdd(fetch, R); }
    float acc = 0;
    while (nb0 < n_items) {
        cbuf ^= 1;
        const uint32_t bar = bar0 + 8u * cbuf, par = (parity >> cbuf) & 1u;
        uint32_t done = 0;
        while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(par) : "memory");
        parity ^= 1u << cbuf;
        const int b0 = nb0, we = min(nb0 + R, n_items);
        __syncwarp();
        nb0 = __shfl_sync(0xffffffffu, pend, 0);
        if (nb0 < n_items) { issue(cbuf ^ 1, nb0); if (lane == 0) pend = atomicAdd(fetch, R); }
        const int item = b0 + lane;
        if (item < we) {
            const unsigned char* buf = stage + cbuf * BUF;
            const float4 o = reinterpret_cast<const float4*>(buf)[lane], d = reinterpret_cast<const float4*>(buf + R * 16)[lane];
            const int f = item < n_primary && !(MODE & 1) ? reinterpret_cast<const int*>(buf + R * 32)[lane] : 0;
            out[item] = o.x + d.y + (float)f;
        }
        __syncwarp();
    }
}
int main() {
    const int np = 100003, ns = 50001, cap = 1 << 18;
    float4 *org, *dir, *sorg, *sdir; int *flags, *fetch; float* out;
    cudaMalloc(&org, cap * 16); cudaMalloc(&dir, cap * 16); cudaMalloc(&sorg, cap * 16); cudaMalloc(&sdir, cap * 16); cudaMalloc(&flags, cap * 4); cudaMalloc(&fetch, 4); cudaMalloc(&out, cap * 4 * 2);
    float4* h = new float4[cap]; for (int i = 0; i < cap; ++i) h[i] = make_float4((float)i, (float)(2 * i), 0, 0);
    cudaMemcpy(org, h, cap * 16, cudaMemcpyHostToDevice); cudaMemcpy(dir, h, cap * 16, cudaMemcpyHostToDevice);
    for (int i = 0; i < cap; ++i) h[i] = make_float4((float)(i + 1000000), (float)(3 * i), 0, 0);
    cudaMemcpy(sorg, h, cap * 16, cudaMemcpyHostToDevice); cudaMemcpy(sdir, h, cap * 16, cudaMemcpyHostToDevice);
    int* hf = new int[cap]; for (int i = 0; i < cap; ++i) hf[i] = i % 7; cudaMemcpy(flags, hf, cap * 4, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(fetch, 0, 4); cudaMemset(out, 0, cap * 8);
        const int smem = 12 * 2 * BUF;
        if (mode == 0) { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<0><<<296, 384, smem>>>(org, dir, flags, np, sorg, sdir, ns, fetch, out); }
        else { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<1><<<296, 384, smem>>>(org, dir, flags, np, sorg, sdir, ns, fetch, out); }
        cudaError_t e = cudaDeviceSynchronize();
        float* ho = new float[np + ns]; cudaMemcpy(ho, out, (np + ns) * 4, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int i = 0; i < np + ns; ++i) { const float want = i < np ? (float)i + (float)(2 * i) + (mode ? 0 : i % 7) : (float)(i - np + 1000000) + (float)(3 * (i - np)); if (ho[i] != want) ++bad; }
        printf("mode %d: %s, mismatches %ld\n", mode, cudaGetErrorString(e), bad);
    }
    return 0;
}
