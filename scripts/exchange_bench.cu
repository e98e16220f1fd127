// Micro-benchmark of per-timestep grid exchange primitives on a co-resident grid (diagnostic tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/exchange_bench scripts/exchange_bench.cu
// mode 0: counter barrier (release red, relaxed poll of the counter, acquire fence) + one ld.cg of
//         exchanged data after it (what the round-1 fused kernel does per step)
// mode 1: tagged 64-bit messages, all-to-all: every CTA stores ONE header word {step tag | payload}
//         into hdr[parity][cta] (8-byte stride), warp 0 of every CTA polls all G headers until their
//         tag is the step's; no atomics, no fences
// mode 2: like 1 with the headers 128 bytes apart
// mode 3: like 1, and every CTA also stores `nent` tagged entry words that every reader fetches
// mode 4: per-sample words, every CTA adds to 128 words (one lane per sample), 128 lanes poll their
//         own word until G arrivals are in (arrival count and data in one word, relaxed atomics)
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long ld_relaxed64(const unsigned long long *p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_relaxed64(unsigned long long *p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

__global__ void bench(unsigned *bar, unsigned long long *hdr, unsigned long long *ent, unsigned *data, long long *out, int iters, int mode, int nent, unsigned long long *inbox) {
    const unsigned G = gridDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ unsigned s_sum;
    unsigned gen = 0, acc = 0;
    long long t0 = clock64();
    const int hs = mode == 2 ? 16 : 1;  // header stride in u64
    for (int it = 0; it < iters; ++it) {
        __syncthreads();
        gen += 1;
        if (mode == 0) {
            if (tid < 4) atomicAdd(data + (it % 3) * 128 + tid, 1u);  // a few exchange atomics
            __syncthreads();
            if (tid == 0) {
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
                const unsigned target = G * gen;
                while ((int)(ld_relaxed(bar) - target) < 0) {}
                asm volatile("fence.acquire.gpu;" ::: "memory");
            }
            __syncthreads();
            if (tid < 128) acc += __ldcg(data + (it % 3) * 128 + tid);
        } else if (mode >= 1 && mode <= 3) {
            const int par = it & 1;
            const unsigned long long tag = (unsigned long long)(gen & 0xffu) << 56;
            if (mode == 3 && tid < nent) st_relaxed64(ent + ((size_t)par * G + blockIdx.x) * 64 + tid, tag | (unsigned)tid);
            if (tid == 0) st_relaxed64(hdr + ((size_t)par * 160 + blockIdx.x) * hs, tag | (mode == 3 ? (unsigned)nent : 0u));
            if (warp == 0) {
                for (unsigned c = lane; c < G; c += 32) {
                    unsigned long long h;
                    do { h = ld_relaxed64(hdr + ((size_t)par * 160 + c) * hs); } while ((h >> 56) != (gen & 0xffu));
                    const int ne = (int)(h & 0xffffu);
                    for (int e = 0; e < ne; ++e) {
                        unsigned long long w;
                        do { w = ld_relaxed64(ent + ((size_t)par * G + c) * 64 + e); } while ((w >> 56) != (gen & 0xffu));
                        acc += (unsigned)w;
                    }
                }
            }
        } else if (mode == 5 || mode == 6 || mode == 7) {
            // parallel polling: every lane keeps the loads of all its headers in flight together
            const int par = it & 1;
            const unsigned long long tag = (unsigned long long)(gen & 0xffu) << 56;
            if (mode == 6 && tid < nent) st_relaxed64(ent + ((size_t)par * G + blockIdx.x) * 64 + tid, tag | (unsigned)tid);
            if (tid == 0) st_relaxed64(hdr + ((size_t)par * 160 + blockIdx.x), tag | (mode == 6 && (blockIdx.x % 10) == (it % 10) ? (unsigned)nent : 0u));
            if (warp == 0) {
                unsigned long long h[5], h2[5];
                unsigned need = 0;
                #pragma unroll
                for (int k = 0; k < 5; ++k) if (lane + 32 * k < G) need |= 1u << k;
                while (need) {
                    #pragma unroll
                    for (int k = 0; k < 5; ++k) if ((need >> k) & 1u) h[k] = ld_relaxed64(hdr + ((size_t)par * 160 + lane + 32 * k));
                    if (mode == 7) {
                        { const long long s0 = clock64(); while (clock64() - s0 < 300) {} }
                        #pragma unroll
                        for (int k = 0; k < 5; ++k) if ((need >> k) & 1u) h2[k] = ld_relaxed64(hdr + ((size_t)par * 160 + lane + 32 * k));
                    }
                    #pragma unroll
                    for (int k = 0; k < 5; ++k) if (((need >> k) & 1u) && (h[k] >> 56) == (gen & 0xffu)) need &= ~(1u << k);
                    if (mode == 7) {
                        #pragma unroll
                        for (int k = 0; k < 5; ++k) if (((need >> k) & 1u) && (h2[k] >> 56) == (gen & 0xffu)) { need &= ~(1u << k); h[k] = h2[k]; }
                    }
                }
                #pragma unroll 1
                for (int k = 0; k < 5; ++k) {
                    const unsigned c = lane + 32 * k;
                    if (c >= G) break;
                    const int ne = (int)(h[k] & 0xffffu);
                    for (int e = 0; e < ne; ++e) {
                        unsigned long long w;
                        do { w = ld_relaxed64(ent + ((size_t)par * G + c) * 64 + e); } while ((w >> 56) != (gen & 0xffu));
                        acc += (unsigned)w;
                    }
                }
            }
        } else if (mode == 8 || mode == 9) {
            // personalised all-to-all (the NCCL LL idea): every (destination, source) pair owns one 32-byte
            // sector {header, 3 inline entries}, every word tagged; a CTA stores its message G times (one
            // copy per destination) and polls only its own inbox — no line is read by more than one CTA
            const int par = it & 1;
            const unsigned long long tag = (unsigned long long)(gen & 0xffu) << 56;
            const int ne = (mode == 9 && (blockIdx.x % 10) == (it % 10)) ? nent : 0;
            if (tid < (int)G) {
                unsigned long long *slot = inbox + (((size_t)par * 160 + tid) * 160 + blockIdx.x) * 4;
                for (int e = 0; e < ne; ++e) st_relaxed64(slot + 1 + e, tag | (unsigned)e);
                st_relaxed64(slot, tag | (unsigned)ne);
            }
            if (warp == 0) {
                unsigned long long h[5], e0[5], e1[5], e2[5];
                unsigned need = 0;
                #pragma unroll
                for (int k = 0; k < 5; ++k) if (lane + 32 * k < G) need |= 1u << k;
                while (need) {
                    #pragma unroll
                    for (int k = 0; k < 5; ++k) if ((need >> k) & 1u) {
                        const unsigned long long *slot = inbox + (((size_t)par * 160 + blockIdx.x) * 160 + lane + 32 * k) * 4;
                        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(h[k]), "=l"(e0[k]) : "l"(slot) : "memory");
                        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(e1[k]), "=l"(e2[k]) : "l"(slot + 2) : "memory");
                    }
                    #pragma unroll
                    for (int k = 0; k < 5; ++k) if (((need >> k) & 1u) && (h[k] >> 56) == (gen & 0xffu)) {
                        const int n_ = (int)(h[k] & 0xffu);
                        bool ok = true;
                        if (n_ > 0) ok = ok && (e0[k] >> 56) == (gen & 0xffu);
                        if (n_ > 1) ok = ok && (e1[k] >> 56) == (gen & 0xffu);
                        if (n_ > 2) ok = ok && (e2[k] >> 56) == (gen & 0xffu);
                        if (ok) { need &= ~(1u << k); acc += (unsigned)(e0[k] + e1[k]); }
                    }
                }
            }
        } else if (mode == 4) {
            if (tid < 128) {
                unsigned *w = data + (size_t)tid * 32;  // 128-byte stride
                asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(w) : "memory");
                const unsigned target = G * gen;
                while ((int)(ld_relaxed(w) - target) < 0) {}
            }
        }
    }
    __syncthreads();
    if (tid == 0) { out[blockIdx.x] = clock64() - t0; s_sum = acc; }
    if (acc == 0xdeadbeefu) out[0] = 0;
}

int main() {
    unsigned *bar, *data; unsigned long long *hdr, *ent; long long *out;
    cudaMalloc(&bar, 1024); cudaMalloc(&data, 128 * 128 * 4); cudaMalloc(&out, 160 * 8);
    cudaMalloc(&hdr, 2 * 160 * 16 * 8); cudaMalloc(&ent, 2 * 160 * 64 * 8);
    unsigned long long *inbox; cudaMalloc(&inbox, 2 * 160 * 160 * 32); cudaMemset(inbox, 0, 2 * 160 * 160 * 32);
    const int iters = 4000;
    for (int G : {100, 134, 148}) for (int threads : {384}) for (int mode : {0, 4, 8, 9}) for (int nent : {0, 2}) {
        if (nent && mode != 9) continue;
        if (mode == 9 && !nent) continue;
        cudaMemset(bar, 0, 1024); cudaMemset(data, 0, 128 * 128 * 4); cudaMemset(hdr, 0, 2 * 160 * 16 * 8); cudaMemset(ent, 0, 2 * 160 * 64 * 8);
        void *args[] = {&bar, &hdr, &ent, &data, &out, (void *)&iters, &mode, &nent, &inbox};
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        cudaError_t e = cudaLaunchCooperativeKernel((void *)bench, dim3(G), dim3(threads), args, 0, 0);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("G=%d threads=%d mode=%d nent=%d: %s  %.3f us per step\n", G, threads, mode, nent, cudaGetErrorString(e), 1e3 * ms / iters);
    }
    return 0;
}
