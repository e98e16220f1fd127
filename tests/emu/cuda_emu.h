// cuda_emu.h — TEST INFRASTRUCTURE: a small CPU emulation of the CUDA execution model, just enough to run the generic
// window kernel's REAL source (bindsnet_b200/csrc/snn_generic.cu, snn_phases.cuh, snn_common.cuh, snn_api.cu) on the
// host, so that the CPU test tier can check the kernel's logic — work decomposition, indexing, summation orders —
// bit for bit against the oracle without a GPU.  It is not a product path and is never loaded by bindsnet_b200.
//
// Model: one OS thread per CTA; the CTA's threads are cooperatively scheduled fibers (ucontext).  A warp collective
// (__shfl*_sync, __ballot_sync, __any_sync, __syncwarp) is a rendezvous of the warp's 32 fibers with a double-buffered
// exchange slot; __syncthreads a rendezvous of the CTA's fibers; atomics and the grid barrier use the host's atomics, so
// several CTAs really run concurrently.  What this does NOT model: memory ordering weaker than the host's, L1 coherence,
// timing, divergence rules — only full-mask collectives reached by all lanes are supported (all the kernel uses).
#pragma once

#include <sched.h>
#include <ucontext.h>

#include <cmath>
#include <functional>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define __builtin_assume(x) ((void)0)
#define __isShared(p) (true)

struct emu_uint3 { unsigned int x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) uint4 { unsigned int x, y, z, w; };
struct alignas(8) uint2 { unsigned int x, y; };
struct uchar4 { unsigned char x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }
struct dim3 {
    unsigned int x, y, z;
    dim3(unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace emu {

constexpr int WARP = 32;

struct WarpState {
    int arrived = 0;
    unsigned gen = 0;
    unsigned long long slot[2][WARP];
};

struct Cta;
struct Fiber {
    ucontext_t ctx;
    char *stack = nullptr;
    emu_uint3 tid{0, 0, 0};
    int lane = 0, warp = 0;
    unsigned wseq = 0;   // warp collectives executed so far (selects the exchange buffer)
    unsigned cseq = 0;   // CTA collectives executed so far
    unsigned oseq = 0;   // __syncthreads_or calls executed so far (selects the OR accumulator)
    size_t static_off = 0;   // statically declared __shared__ objects handed out so far (identical in every thread)
    bool done = false;
    Cta *cta = nullptr;
};

struct Cta {
    ucontext_t main_ctx;
    Fiber *fibers = nullptr;
    WarpState *warps = nullptr;
    int nthreads = 0, nwarps = 0, live = 0, current = 0;
    int c_arrived = 0;
    unsigned c_gen = 0;
    int c_or[2] = {0, 0};
    struct { int arrived = 0; unsigned gen = 0; } named[16];   // bar.sync id, count
    emu_uint3 bidx{0, 0, 0};
    dim3 bdim, gdim;
    float *dyn_smem = nullptr;
    alignas(16) unsigned char static_smem[16384];  // the kernels' statically declared __shared__ objects
    int s_abort = 0;
    unsigned long long rng = 1;
    void (*entry)(void *) = nullptr;
    void *arg = nullptr;
};

extern thread_local Cta *tls_cta;
extern thread_local Fiber *tls_cur;

void yield();                                   // switch to the next runnable fiber of this CTA
void run_grid(int grid, int block, size_t dyn_smem_bytes, void (*entry)(void *), void *arg);
// independent CTAs (no grid-wide synchronisation): a 2-D grid executed by a small pool of host threads
void run_grid_independent(int gx, int gy, int block, size_t dyn_smem_bytes, void (*entry)(void *), void *arg);

// a kernel without grid-wide synchronisation: `body` is what each CUDA thread executes (SNN_LAUNCH)
void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()> &body);
// storage of a statically sized __shared__ array (SNN_SHARED): every thread of the CTA walks the same declarations in the
// same order, so a per-thread running offset names the same bytes in all of them
inline void *static_shared(size_t bytes) {
    Fiber *f = tls_cur;
    const size_t off = (f->static_off + 15) & ~(size_t)15;
    f->static_off = off + bytes;
    if (f->static_off > sizeof(f->cta->static_smem)) { fprintf(stderr, "emu: static shared memory exhausted\n"); abort(); }
    return f->cta->static_smem + off;
}

// mbarrier + bulk copy (cp.async.bulk ... mbarrier::complete_tx): the copy happens AT ISSUE — the earliest moment the
// hardware could overwrite the destination, i.e. the adversarial case for a missing synchronisation before the issue.
struct Mbar { uint32_t phase; int32_t pend; };   // pend: outstanding bytes (+ 2^30 while the phase's arrival is outstanding)
static_assert(sizeof(Mbar) == 8, "an mbarrier object is 64 bits");
constexpr int32_t MBAR_ARRIVAL = 1 << 30;
inline void mbar_settle(Mbar *m) { if (m->pend == 0) { ++m->phase; m->pend = MBAR_ARRIVAL; } }

inline void warp_rendezvous() {
    Fiber *f = tls_cur;
    WarpState &w = f->cta->warps[f->warp];
    const unsigned g = w.gen;
    if (++w.arrived == WARP) { w.arrived = 0; ++w.gen; }
    else while (w.gen == g) yield();
}

inline void cta_rendezvous() {
    Cta *c = tls_cta;
    const unsigned g = c->c_gen;
    if (++c->c_arrived == c->nthreads) { c->c_arrived = 0; ++c->c_gen; }
    else while (c->c_gen == g) yield();
}

// bar.sync id, count: a barrier among the `count` threads that name it
inline void named_barrier(int id, int count) {
    auto &nb = tls_cta->named[id & 15];
    const unsigned g = nb.gen;
    if (++nb.arrived == count) { nb.arrived = 0; ++nb.gen; }
    else while (nb.gen == g) yield();
}

template <class T> inline unsigned long long to_bits(T v) { unsigned long long b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

// every lane deposits `v`; returns the whole warp's values (valid until the lane's next collective)
template <class T> inline const unsigned long long *warp_exchange(T v) {
    Fiber *f = tls_cur;
    WarpState &w = f->cta->warps[f->warp];
    const unsigned buf = f->wseq++ & 1u;
    w.slot[buf][f->lane] = to_bits(v);
    warp_rendezvous();
    return w.slot[buf];
}

}  // namespace emu

#define threadIdx (emu::tls_cur->tid)
#define blockIdx (emu::tls_cta->bidx)
#define blockDim (emu::tls_cta->bdim)
#define gridDim (emu::tls_cta->gdim)

// ---- warp / CTA collectives (full mask only)
inline void __syncwarp(unsigned = 0xffffffffu) { emu::tls_cur->wseq++; emu::warp_rendezvous(); }
inline void __syncthreads() { emu::tls_cur->cseq++; emu::cta_rendezvous(); }
inline int __syncthreads_or(int pred) {
    emu::Cta *c = emu::tls_cta;
    emu::tls_cur->cseq++;
    // two accumulators used alternately by successive OR-barriers; the last arriver of one clears the other, so that it is
    // clean for the next OR-barrier however many plain __syncthreads lie in between
    const unsigned buf = emu::tls_cur->oseq++ & 1u;
    if (pred) c->c_or[buf] = 1;
    const unsigned g = c->c_gen;
    if (++c->c_arrived == c->nthreads) { c->c_or[buf ^ 1u] = 0; c->c_arrived = 0; ++c->c_gen; }
    else while (c->c_gen == g) emu::yield();
    return c->c_or[buf];
}
template <class T> inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu::from_bits<T>(emu::warp_exchange(v)[src & 31]); }
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned delta, int = 32) {
    const int lane = emu::tls_cur->lane;
    const unsigned long long *s = emu::warp_exchange(v);
    return lane >= (int)delta ? emu::from_bits<T>(s[lane - (int)delta]) : v;
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int mask, int = 32) {
    const int lane = emu::tls_cur->lane;
    return emu::from_bits<T>(emu::warp_exchange(v)[(lane ^ mask) & 31]);
}
inline unsigned __ballot_sync(unsigned, int pred) {
    const unsigned long long *s = emu::warp_exchange<int>(pred ? 1 : 0);
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= (unsigned)(s[l] & 1ull) << l;
    return m;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }

// ---- scalar intrinsics
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { return (unsigned)((((unsigned long long)hi << 32) | lo) >> (s & 31u)); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline float __expf(float x) { return expf(x); }
inline float __logf(float x) { return logf(x); }
inline long long clock64() { static thread_local long long c = 0; return c += 64; }
inline void __nanosleep(unsigned) { emu::yield(); sched_yield(); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <class T> inline T __ldcg(const T *p) { return *(const volatile T *)p; }
inline uint4 __ldcg(const uint4 *p) { return *p; }
inline float4 __ldcg(const float4 *p) { return *p; }
template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline T __ldcs(const T *p) { return *p; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline unsigned __vcmpeq4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int k = 0; k < 4; ++k)
        if (((a >> (8 * k)) & 0xffu) == ((b >> (8 * k)) & 0xffu)) r |= 0xffu << (8 * k);
    return r;
}
inline unsigned __vadd4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int k = 0; k < 4; ++k) r |= ((((a >> (8 * k)) & 0xffu) + ((b >> (8 * k)) & 0xffu)) & 0xffu) << (8 * k);
    return r;
}
inline unsigned __vcmpne4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int k = 0; k < 4; ++k)
        if (((a >> (8 * k)) & 0xffu) != ((b >> (8 * k)) & 0xffu)) r |= 0xffu << (8 * k);
    return r;
}

inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline float fminf_(float a, float b) { return a < b ? a : b; }

// ---- atomics on "global" memory (shared by the CTAs' OS threads)
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicOr(int *p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicExch(unsigned *p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

// ---- the few runtime calls the generic path makes (host memory stands in for device memory)
typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0, cudaErrorLaunchOutOfResources = 701, cudaErrorMemoryAllocation = 2 };
enum { cudaDevAttrMultiProcessorCount = 16 };
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { *v = 148; return cudaSuccess; }
template <class F> inline cudaError_t cudaLaunchCooperativeKernel(F, dim3, dim3, void **, size_t, cudaStream_t) { return 1; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyDeviceToHost = 2 };
inline cudaError_t cudaMemcpy(void *d, const void *s_, size_t n, cudaMemcpyKind) { memcpy(d, s_, n); return cudaSuccess; }
