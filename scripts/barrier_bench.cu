// Micro-benchmark of grid-barrier implementations on a co-resident grid (diagnostic tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/barrier_bench scripts/barrier_bench.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_volatile(const unsigned *p) { unsigned v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// mode 0: acq_rel atomic + last arriver publishes generation + acquire polling (current)
// mode 1: release reduction on a monotonic counter, relaxed polling of the counter, acquire fence once
// mode 2: like 1 but polling with ld.volatile
// mode 3: cooperative groups grid.sync()
// mode 4: like 1, but without any memory ordering at all (lower bound: relaxed atomic + relaxed poll)
// `stores`: number of 4-byte global stores each thread issues before arriving (models the step's payload)
__global__ void bench(unsigned *bar, float *payload, long long *out, int iters, int mode, int stores) {
    cg::grid_group grid = cg::this_grid();
    const unsigned G = gridDim.x;
    unsigned gen = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int s = 0; s < stores; ++s) payload[((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8 + s] = (float)it;
        if (mode == 3) { grid.sync(); continue; }
        __syncthreads();
        gen += 1;
        if (threadIdx.x == 0) {
            if (mode == 0) {
                unsigned prev;
                asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(bar) : "memory");
                if (prev + 1u == G * gen) st_release(bar + 32, gen);
                else while ((int)(ld_acquire(bar + 32) - gen) < 0) {}
            } else if (mode == 1 || mode == 2) {
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
                const unsigned target = G * gen;
                if (mode == 1) { while ((int)(ld_relaxed(bar) - target) < 0) {} }
                else { while ((int)(ld_volatile(bar) - target) < 0) {} }
                asm volatile("fence.acquire.gpu;" ::: "memory");
            } else {
                asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
                const unsigned target = G * gen;
                while ((int)(ld_relaxed(bar) - target) < 0) {}
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
    unsigned *bar; float *payload; long long *out;
    cudaMalloc(&bar, 1024); cudaMalloc(&payload, 148 * 1024 * 8 * 4); cudaMalloc(&out, 148 * 8);
    const int iters = 2000;
    for (int G : {100, 148}) for (int threads : {512}) for (int stores : {0, 2}) for (int mode = 0; mode < 5; ++mode) {
        cudaMemset(bar, 0, 1024);
        void *args[] = {&bar, &payload, &out, (void *)&iters, &mode, &stores};
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        cudaError_t e = cudaLaunchCooperativeKernel((void *)bench, dim3(G), dim3(threads), args, 0, 0);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("G=%d threads=%d stores=%d mode=%d: %s  %.3f us per barrier\n", G, threads, stores, mode, cudaGetErrorString(e), 1e3 * ms / iters);
    }
    return 0;
}
