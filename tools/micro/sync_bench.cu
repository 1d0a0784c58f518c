// Micro-benchmark (GPU box): cost of one grid-wide barrier and of the per-phase bookkeeping of k_wavefront.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned long long gns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// own barrier: one arrival atomic per CTA, generation flag polled by thread 0
__device__ __forceinline__ void my_sync(unsigned int* bar, unsigned int nblocks, unsigned int& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++gen;
        __threadfence();
        const unsigned int old = atomicAdd(bar, 1u);
        if (old == gen * nblocks - 1) { atomicExch(bar + 32, gen); }
        else { while (*(volatile unsigned int*)(bar + 32) < gen) {} }
        __threadfence();
    }
    __syncthreads();
}

__global__ void k(int mode, int iters, unsigned int* bar, int* counter, unsigned long long* out) {
    cg::grid_group grid = cg::this_grid();
    unsigned int gen = 0;
    const bool boss = blockIdx.x == 0 && threadIdx.x == 0;
    unsigned long long t0 = 0;
    grid.sync();
    if (boss) t0 = gns();
    for (int i = 0; i < iters; ++i) {
        if (mode == 1 || mode == 3) { if ((threadIdx.x & 31) == 0) atomicAdd(counter, 32); }           // every warp fetches once
        if (mode == 4) { if (threadIdx.x == 0) atomicAdd(counter, 32); }                                  // every CTA fetches once
        if (mode == 2 || mode == 3) my_sync(bar, gridDim.x, gen); else grid.sync();
    }
    if (boss) out[0] = gns() - t0;
}

int main() {
    int dev = 0; cudaSetDevice(dev);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    unsigned int* bar; int* counter; unsigned long long* out;
    cudaMalloc(&bar, 4096); cudaMalloc(&counter, 4096); cudaMallocManaged(&out, 64);
    for (int bps = 1; bps <= 3; ++bps) {
        for (int mode = 0; mode <= 4; ++mode) {
            cudaMemset(bar, 0, 4096); cudaMemset(counter, 0, 4096);
            int iters = 200, grid = p.multiProcessorCount * bps;
            void* args[] = {&mode, &iters, &bar, &counter, &out};
            cudaError_t e = cudaLaunchCooperativeKernel((const void*)k, dim3(grid), dim3(256), args, 0, 0);
            cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("launch failed %s\n", cudaGetErrorString(e)); continue; }
            const char* names[] = {"cg grid.sync", "warp atomics + cg sync", "own barrier", "warp atomics + own barrier", "CTA atomics + cg sync"};
            printf("CTAs/SM %d  %-28s %.2f us per iteration\n", bps, names[mode], out[0] / 1e3 / iters);
        }
    }
    return 0;
}
