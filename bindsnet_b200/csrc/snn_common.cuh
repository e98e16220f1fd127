// snn_common.cuh — device-side helpers shared by the window kernels.
//
// Arithmetic contract: every floating-point operation below is a single IEEE fp32 rounding, in
// the order the reference's ATen ops apply them (the library is compiled with --fmad=false, so
// the compiler never contracts a*b+c).  Where the reference leaves a summation order to ATen
// the kernels sum in ascending index order, like oracle/snn_oracle.c, so kernel and oracle
// agree bit for bit.
#pragma once

#ifdef SNN_EMU   // tests/emu: the same sources compiled for the host on a small CUDA-model emulation (test infrastructure)
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "../../include/snn_b200.h"

// Kernel launches and statically sized __shared__ arrays of the small kernels (snn_ops.cu, snn_encode.cu, snn_readout.cu)
// go through these two macros so that tests/emu can run the same sources on its CUDA-model emulation.
#ifdef SNN_EMU
#define SNN_LAUNCH(kernel, grid, block, smem, stream, ...) emu::launch((grid), (block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })
#define SNN_SHARED(type, name, count) type *name = (type *)emu::static_shared(sizeof(type) * (size_t)(count))
#define SNN_SHARED2(type, name, rows, cols) type(*name)[cols] = (type(*)[cols])emu::static_shared(sizeof(type) * (size_t)(rows) * (cols))
#define SNN_DYN_SHARED(type, name) type *name = (type *)emu::tls_cta->dyn_smem
#else
#define SNN_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define SNN_SHARED(type, name, count) __shared__ type name[count]
#define SNN_SHARED2(type, name, rows, cols) __shared__ type name[rows][cols]
#define SNN_DYN_SHARED(type, name) extern __shared__ type name[]
#endif

#define SNN_TILE 32          // neurons (columns) per work item: one warp lane per column
#define SNN_GEN_THREADS 256  // generic kernel: 8 warps per CTA
#define SNN_GEN_WARPS (SNN_GEN_THREADS / 32)

struct DevLayer {
    snn_layer_t L;
    uint32_t *bits;             // [2][B][nw]  bit-packed spikes, slot t&1 holds s(t)
    uint32_t *candbits;         // [B][nw]     DC one_spike: threshold crossers of this step
    unsigned long long *keys;   // [2][B]      DC one_spike: arg-max tie-break keys
    float *xpub;                // [2][B][n]   trace published for STDP readers (slot t&1), or NULL
    float *thdec;               // [2][n]      DC: the decayed adaptive threshold step t used (slot t&1)
    int32_t *thcnt;             // [3][n]      DC: threshold crossers of step t summed over the batch (slot t%3)
    uint32_t *anyf;             // [3][B]      wide source layers (nw > 32) of dense connections: non-zero iff the sample spiked
                                //             in step t (slot t%3) — lets a gather skip an all-zero bit row without reading it
    int32_t nw;                 // ceil(n / 32)
    int32_t item0;              // first work-item index of this layer
};

// State of an MSTDP rule, double-buffered: step t reads slot (t + T) & 1 and writes the other one, so
// that no thread overwrites a value another one still needs in the same step; slot 0 is the caller's
// tensors (snn_conn_t::p_plus ...), slot 1 lives in the workspace.
struct DevMstdp {
    float *pp[2], *pm[2], *el[2];
    uint8_t *sp[2], *st[2];
};

struct DevNet {
    int32_t n_layers, n_conns, learning, T, B, normalize, total_items, any_one_spike;
    int32_t any_mask;             // some connection carries a mask (Network.run(..., masks=...))
    int32_t one_step;             // Network.run(one_step=True): layers in insertion order, inputs from current spikes
    uint32_t seed, step_offset;
    int32_t nch, cs;              // phases 1 / 2: sample chunks per tile, samples per chunk
    int32_t p3_total;             // phase 3: number of work units
    int32_t p3_first[SNN_MAX_CONNS], p3_rc[SNN_MAX_CONNS];   // first unit / row chunks per tile of every connection
    int32_t *err;               // device error flags (may be NULL)
    long long *prof;            // profiling only (env SNN_B200_GPROF): [grid][8] phase cycles of thread 0
    unsigned int *bar;          // [0] arrival count, [32] generation, [64] abort
    DevLayer layers[SNN_MAX_LAYERS];
    snn_conn_t conns[SNN_MAX_CONNS];
    DevMstdp mst[SNN_MAX_CONNS];
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
#ifdef SNN_EMU
    return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#else
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void st_release_u32(unsigned int *p, unsigned int v) {
#ifdef SNN_EMU
    __atomic_store_n(p, v, __ATOMIC_RELEASE);
#else
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}

// Grid barrier on ONE monotonic arrival counter in global memory: a release reduction to arrive (no return value:
// one L2 transaction), relaxed polling of the same word with a short back-off until `gen * nblocks` arrivals are in,
// then one acquire fence.  `gen` is the caller's barrier count (a register, identical in every CTA).  Requires all
// CTAs of the grid to be co-resident (cooperative launch).  A time-out (~2 s) raises SNN_ERR_BARRIER and makes every
// CTA leave the time loop instead of hanging the device: the CTA that gives up adds 2^30 to the counter, which
// releases every present and future wait and is recognised as "abort" by whoever reads it.  The pollers back off
// (nanosleep): in the generic kernel many CTAs wait here while others still stream state and weights through L2.
__device__ __forceinline__ bool grid_barrier(unsigned int *bar, unsigned int nblocks, int32_t *err, unsigned int &gen) {
#ifdef SNN_EMU
    int &s_abort = emu::tls_cta->s_abort;
#else
    __shared__ int s_abort;
#endif
    ++gen;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int target = gen * nblocks;
#ifdef SNN_EMU
        __atomic_fetch_add(bar, 1u, __ATOMIC_RELEASE);
#else
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
#endif
        int ab = 0;
        const long long t0 = clock64();
        unsigned int ns = 20;
        for (;;) {
            unsigned int v;
#ifdef SNN_EMU
            v = __atomic_load_n(bar, __ATOMIC_ACQUIRE);
#else
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
#endif
            if ((int)(v - target) >= 0) { ab = (v - target) >= 0x20000000u; break; }
            __nanosleep(ns);
            if (ns < 160) ns += 20;
            if (clock64() - t0 > 4000000000LL) {
                if (err) atomicOr(err, SNN_ERR_BARRIER);
                atomicAdd(bar, 0x40000000u);
                ab = 1;
                break;
            }
        }
#ifdef SNN_EMU
        __atomic_thread_fence(__ATOMIC_SEQ_CST);
#else
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
        s_abort = ab;
    }
    __syncthreads();
    return s_abort == 0;
}

__device__ __forceinline__ float clampf(float w, float lo, float hi) {
    w = w < lo ? lo : w;
    w = w > hi ? hi : w;
    return w;
}

// Nodes.forward trace update (nodes.py:96-103): decay, then set / add on spike.
__device__ __forceinline__ float trace_step(float x, bool s, float decay, float scale, int additive) {
    x = x * decay;
    if (additive) x = x + scale * (s ? 1.0f : 0.0f);
    else if (s) x = scale;
    return x;
}

// LIFNodes.forward (nodes.py:500-529).  `xin` is masked in place like the reference does.
__device__ __forceinline__ bool lif_step(const snn_layer_t &L, float &v, float &rc, float &xin) {
    v = L.decay * (v - L.rest) + L.rest;
    if (rc > 0.0f) xin = 0.0f;
    rc = rc - L.dt;
    v = v + xin;
    const bool s = v >= L.thresh;
    if (s) { rc = L.refrac; v = L.reset; }
    if (L.has_lbound && v < L.lbound) v = L.lbound;
    return s;
}

// IFNodes.forward (nodes.py:377-394): no leak, the gate is taken before the refractory decrement, x stays unmasked.
__device__ __forceinline__ bool if_step(const snn_layer_t &L, float &v, float &rc, float xin) {
    const float gate = rc <= 0.0f ? 1.0f : 0.0f;
    v = v + gate * xin;
    rc = rc - L.dt;
    const bool s = v >= L.thresh;
    if (s) { rc = L.refrac; v = L.reset; }
    if (L.has_lbound && v < L.lbound) v = L.lbound;
    return s;
}

// CurrentLIFNodes.forward (nodes.py:770-791): decaying synaptic current `ic`, gate taken after the decrement.
__device__ __forceinline__ bool clif_step(const snn_layer_t &L, float &v, float &rc, float &ic, float xin) {
    v = L.decay * (v - L.rest) + L.rest;
    ic = ic * L.i_decay;
    rc = rc - L.dt;
    ic = ic + xin;
    const float gate = rc <= 0.0f ? 1.0f : 0.0f;
    v = v + gate * ic;
    const bool s = v >= L.thresh;
    if (s) { rc = L.refrac; v = L.reset; }
    if (L.has_lbound && v < L.lbound) v = L.lbound;
    return s;
}

// BoostedLIFNodes.forward (nodes.py:620-647): v *= decay, x masked in place while refractory, reset to 0.
__device__ __forceinline__ bool boosted_step(const snn_layer_t &L, float &v, float &rc, float &xin) {
    v = v * L.decay;
    if (rc > 0.0f) xin = 0.0f;
    rc = rc - L.dt;
    v = v + xin;
    const bool s = v >= L.thresh;
    if (s) { rc = L.refrac; v = 0.0f; }
    return s;
}

// DiehlAndCookNodes.forward up to the threshold test (nodes.py:1077-1092); `theta` is the
// already decayed adaptive threshold of the neuron.  Returns the candidate flag.
__device__ __forceinline__ bool dc_step(const snn_layer_t &L, float &v, float &rc, float xin, float theta) {
    v = L.decay * (v - L.rest) + L.rest;
    const float gate = rc <= 0.0f ? 1.0f : 0.0f;
    v = v + gate * xin;
    rc = rc - L.dt;
    const bool s = v >= (L.thresh + theta);
    if (s) { rc = L.refrac; v = L.reset; }
    return s;
}

// One STDP-family update of a single synapse, in the reference's order: pre term, post term,
// weight decay, clamp (learning.py:87-104,390-420,626-653,1110-1136; MCC_learning.py:86-110,224-302).
// U / V are the batch-reduced outer products of this step for this synapse (0 if untouched).
__device__ __forceinline__ float apply_rule(const snn_conn_t &C, float w, float U, bool pre_t, float V, bool post_t) {
    if (C.rule == SNN_RULE_WDEP_POSTPRE) {
        float upd = 0.0f;
        if (C.nu0 != 0.0f) upd = upd - (C.nu0 * (pre_t ? U : 0.0f)) * (w - C.wmin);
        if (C.nu1 != 0.0f) upd = upd + (C.nu1 * (post_t ? V : 0.0f)) * (C.wmax - w);
        w = w + upd;
    } else if (C.rule == SNN_RULE_MCC_POSTPRE) {
        if (pre_t) w = w - U * C.dt_scale;
        if (post_t) w = w + V * C.dt_scale;
    } else if (C.rule == SNN_RULE_POSTPRE) {
        if (pre_t) w = w - U;
        if (post_t) w = w + V;
    } else if (C.rule == SNN_RULE_HEBBIAN) {   // learning.py:1124-1134: U / V are the plain reduced sums
        if (C.nu0 != 0.0f) w = w + C.nu0 * (pre_t ? U : 0.0f);
        if (C.nu1 != 0.0f) w = w + C.nu1 * (post_t ? V : 0.0f);
    }
    if (C.weight_decay != 0.0f) w = w * C.weight_decay;
    if (C.has_clamp) w = clampf(w, C.wmin, C.wmax);
    return w;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
