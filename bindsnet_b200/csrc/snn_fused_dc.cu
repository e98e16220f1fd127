// snn_fused_dc.cu — fused persistent window kernel for the DiehlAndCook2015 graph
//     X (Input) --W, learned--> Ae (DiehlAndCookNodes) --exc*I--> Ai (LIFNodes) --(-inh)(1-I)--> Ae
// (reference wiring: bindsnet/models/models.py:94-244; per-step semantics: SURVEY.md App. A).
//
// One cooperative grid, ONE grid barrier per timestep, the whole T-step window in one launch.
//
// Partition: CTA c owns TJ consecutive neurons j of Ae AND the same TJ neurons of Ai (the
// Ae->Ai matrix is diagonal, so that pairing is CTA-local) for all B samples.  It keeps
//   * W[:, tile]  (P x TJ fp32; 50 KB at P=784, TJ=16) in SHARED MEMORY for the whole window:
//     the spike-gather reads it, STDP + clamp rewrite it, normalize() finishes on it; HBM sees
//     it once in and once out,
//   * v, refrac_count, x of its Ae neurons and v, refrac_count of its Ai neurons in REGISTERS
//     (thread = one sample x 4 neurons), theta[tile] in shared memory.
// Per step the grid exchanges only: per sample the arg-max one_spike key (atomicMax) and the
// number of Ai spikes (atomicAdd) — the Ai->Ae matrix is constant off-diagonal, so lateral
// inhibition needs just that count — plus the input-trace rows of the few winners.
// The step's input spikes are prepared once per window by a pre-pass in two forms — per sample
// an ascending list of spiking pixels (for the gather) and per pixel a bit mask over samples
// (for the STDP pre term) — and staged a step ahead into shared memory with cp.async.bulk
// (TMA bulk copy) + mbarrier.
//
// Loop iteration t = [finalise step t-1: exchange results, winner, Ae trace, STDP on the tile]
//                    [step t: gather, Ae/Ai update, candidates -> atomics] [grid barrier].
//
// Arithmetic and summation orders are those of snn_phases.cuh / oracle/snn_oracle.c, so the
// result is bit-identical to the generic kernel and to the oracle.
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "snn_common.cuh"

int snn_verify_structure(const snn_conn_t &C, int n, int32_t *err, cudaStream_t stream);

namespace {

constexpr int XR = 8;        // input-trace rows staged per CTA per step (samples with a candidate)
constexpr int EV_CAP = 32;   // staged events per sample and step; longer lists take the slow path

struct FusedParams {
    snn_layer_t X, E, I;      // Input, DiehlAndCookNodes (Ae), LIFNodes (Ai)
    snn_conn_t C;             // X -> Ae
    float exc, inh_neg;       // diag value of Ae->Ai, off-diag value of Ai->Ae
    int32_t T, B, P, n, learning, normalize;
    int32_t SW;               // words per sample row of inS
    int32_t SB;               // bytes of one event-list block
    int32_t liE;              // index of Ae in the user's layer list (enters the tie-break hash)
    uint32_t o_W, o_tx, o_ev, o_inT, o_xrow, o_rep, o_xown, o_theta, o_live, o_misc;  // smem_layout() byte offsets
    int32_t dbg;              // profiling only (env SNN_B200_DEBUG): 1 no barrier wait, 2 skip STDP,
                              // 4 skip gather, 8 skip trace publish, 16 skip staging — results invalid
    uint32_t seed, step_offset;
    uint32_t *inS;            // [T+1][B][SW]  slot t = spikes of step t-1: bit i of sample b
    uint32_t *inT;            // [T+1][P][BW]  same spikes: bit b of pixel i
    unsigned char *evS;       // [T+1][SB]     same spikes as lists: u16 count[B], dense flag, pad to 16 B,
                              //               then u16 idx[B][EV_CAP] ascending, padded with P
    unsigned long long *win;  // [3][B] arg-max keys, slot t % 3
    unsigned int *sisum;      // [3][B] Ai spike counts, slot t % 3 (slot 2 = step -1)
    float *xpub;              // [3][B][P] published input traces, slot t % 3
    float *delta_w, *delta_theta;   // snn_run_opts_t: write the window's weight / theta change instead of the new values
    unsigned int *bar;        // [0] arrivals (monotonic), [32] generation
    int *dense;               // [T+1] slot holds a sample whose event list overflowed EV_CAP
    int32_t *err;
    long long *prof;          // profiling only (env SNN_B200_PROF): [grid][NPROF] phase cycles of thread 0
};
constexpr int NPROF = 16;

#ifdef SNN_EMU   // tests/emu: mbarrier / bulk copy / polling loads on the CUDA-model emulation (test infrastructure)
__device__ __forceinline__ void mbar_init(uint64_t *bar, int) { emu::Mbar *m = (emu::Mbar *)bar; m->phase = 0; m->pend = emu::MBAR_ARRIVAL; }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { emu::Mbar *m = (emu::Mbar *)bar; m->pend += (int32_t)bytes - emu::MBAR_ARRIVAL; emu::mbar_settle(m); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    const bool ok = (((const emu::Mbar *)bar)->phase & 1u) != parity;
    if (!ok) emu::yield();
    return ok;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    memcpy(dst, src, bytes);
    emu::Mbar *m = (emu::Mbar *)bar; m->pend -= (int32_t)bytes; emu::mbar_settle(m);
}
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int *p) {
    const unsigned int v = __atomic_load_n(p, __ATOMIC_ACQUIRE);
    emu::yield(); sched_yield();
    return v;
}
#else
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Grid barrier, one L2 round trip to arrive and a polling load to leave: a monotonic arrival
// counter (acq_rel atomic; the last arriver of generation g publishes g) and a generation word.
// All cross-CTA data is read with ld.global.cg, so no L1 invalidation is needed.
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
#endif

// The grid barrier is ONE monotonic arrival counter: a release reduction to arrive, relaxed polling
// of the same word until nblocks * generation arrivals are in, then one acquire fence
// (scripts/barrier_bench.cu: 1.2 us on B200 against 1.9 us for a counter + generation-word scheme
// and 1.3 us for cooperative groups' grid.sync; 0.85 us is the no-ordering floor).  It is split
// into arrive / wait inside the time loop.  All cross-CTA data is read with ld.global.cg, so no
// L1 invalidation is needed.

struct Misc {  // small per-step scratch (lives in shared memory)
    uint64_t mbar[2];
    uint32_t wl[4];         // first 4 winners of the step being finalised: column << 16 | sample << 8 | staged row slot (0xff: none)
    int cnt[2][32];         // candidates per column (theta update), double-buffered by step parity
    uint32_t wmask[32][8];  // winners of the step being finalised: per column, bit mask over samples
    uint32_t nz4[8][8];     // per column group: samples with a non-zero Ae trace in that group
    int nwl;                // number of winners of the step being finalised
    int nlive;              // live (sample, column group) pairs, listed in live[]
    int denseflag[2];       // staged slot (by buffer) holds a sample whose event list overflowed EV_CAP
    int ncand[2];           // samples with a candidate in this tile (by step parity)
    int candb[2][XR];       // ... the first XR of them: their input-trace rows get staged
    uint32_t candgrp[2];    // column groups holding a candidate (by step parity)
    uint32_t colwin;        // bit j: column j has a winner in the step being finalised
    int abort;              // barrier time-out: leave the time loop
    long long lc[8];        // late-pass statistics (profiling variant only)
    long long pc[16];       // phase timers of thread 0 (profiling variant only)
};

static_assert(offsetof(Misc, wl) % 16 == 0 && offsetof(Misc, nz4) % 16 == 0 && offsetof(Misc, wmask) % 16 == 0, "Misc: 16-byte rows");
struct SmemLayout { size_t W, tx, ev, inT, xrow, rep, xown, theta, live, misc, total; };

__host__ __device__ inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline int ev_count_bytes(int B) { return (int)al16(2 * (size_t)(B + 8)); }  // u16 count[B], then [B] = dense flag
__host__ __device__ inline int ev_block_bytes(int B) { return ev_count_bytes(B) + 2 * B * EV_CAP; }
__host__ __device__ inline SmemLayout smem_layout(int P, int TJ, int B, int BW, int n, int own) {
    SmemLayout L;
    size_t o = 0;
    L.W = o; o += al16(sizeof(float) * (size_t)(P + 1) * (TJ + 4));  // row stride TJ+4 (column passes
                                                                        // hit 8 banks, not 2); + one zero row
    L.tx = o; o += al16(sizeof(float) * (size_t)B * TJ);
    L.ev = o; o += 2 * al16((size_t)ev_block_bytes(B));
    L.inT = o; o += al16(sizeof(uint32_t) * 2 * (size_t)P * BW);
    L.xrow = o; o += al16(sizeof(float) * (size_t)XR * P);
    L.rep = o; o += al16(sizeof(float) * (size_t)(n + 1));
    L.xown = o; o += al16(sizeof(float) * (size_t)own * P);
    L.theta = o; o += al16(sizeof(float) * 32);
    L.live = o; o += al16(sizeof(uint16_t) * (size_t)B * (TJ / 4));
    L.misc = o; o += al16(sizeof(Misc));
    L.total = o;
    return L;
}

// Constants of the STDP passes, kept in shared memory so that the passes can live out of line.
// The kernel is bound by instruction fetch as much as by anything else (measured: a code path that
// is not resident in the 32 KB L1.5 instruction cache runs at ~7 cycles per instruction), so the
// late (rare) STDP pass and the early (every step) pass share ONE body, stdp_rows.
struct PassCtx {
    float *W, *tx;
    const float *xrow;
    const uint32_t *inT;
    const unsigned char *evb;
    const uint16_t *live;
    Misc *M;
    int P, B, evblk, cntb;
    int pre_on, wdep, reduce_mean, has_clamp;
    float Bf, dts, weight_decay, wmin, wmax, nu0, nu1;
};

// The context lives in shared memory and its pointers point into shared memory; loaded back from there they are
// generic pointers to the compiler (LD/ST through the generic path, 64-bit address arithmetic).  These hints
// restore the address space: LDS/STS with 32-bit addresses.
template <class T> __device__ __forceinline__ T *sh(T *p) { __builtin_assume(__isShared((const void *)p)); return p; }
__device__ __forceinline__ PassCtx load_ctx(const PassCtx *cx) {
    PassCtx c = *sh(cx);
    c.W = sh(c.W); c.tx = sh(c.tx); c.xrow = sh(c.xrow); c.inT = sh(c.inT); c.evb = sh(c.evb); c.live = sh(c.live); c.M = sh(c.M);
    return c;
}

// STDP of one step on the column groups selected by `groups` (bit per group), thread = input row i
// (MCC_learning.py:234-299, 86-110; learning.py:641-651 for the weight-dependent pre term).
//   pre term: for every selected group whose live samples (non-zero Ae trace, Misc::nz4) spiked at
//     pixel i:  w - U*dt  on the group's 4 columns, U = sum of the traces in ascending sample order;
//   columns without a pending post term (not in `colwin`) are finished here: decay, clamp;
//   post term: for each of the `nwl` winners (Misc::wl, distinct columns, staged rows):
//     w + x_pre[b,i]*nu1*dt, decay, clamp on that one column.
// `full`: every row of the selected groups is decayed / clamped (first update of a window, weight
// decay).  Per column the operation order is the reference's: pre, post, decay, clamp.
template <int TJ, int BW>
__device__ __noinline__ void stdp_rows(const PassCtx *cx, int sb, uint32_t groups, int nwl, uint32_t colwin, int full) {
    constexpr int WS = TJ + 4;
    // the context lives in shared memory: read it ONCE into registers (the stores to W below would
    // otherwise force every field to be reloaded per row)
    const PassCtx c_ = load_ctx(cx);
    const int P = c_.P;
    const Misc &M = *c_.M;
    const uint4 *cT = (const uint4 *)(c_.inT + sb * P * BW);
    #pragma unroll 1
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        float *wrow = c_.W + i * WS;
        uint32_t q[BW];
        {
            const uint4 q0 = cT[i * (BW / 4)];
            q[0] = q0.x; q[1] = q0.y; q[2] = q0.z; q[3] = q0.w;
            if (BW == 8) { const uint4 q1 = cT[i * (BW / 4) + 1]; q[BW - 4] = q1.x; q[BW - 3] = q1.y; q[BW - 2] = q1.z; q[BW - 1] = q1.w; }
        }
        uint32_t anyq = 0;
        #pragma unroll
        for (int g = 0; g < BW; ++g) anyq |= q[g];
        if (!c_.pre_on) anyq = 0;
        if (anyq | (uint32_t)full) {
            #pragma unroll 1
            for (uint32_t lg = groups; lg; lg &= lg - 1) {
                const int c4 = __ffs(lg) - 1;
                uint32_t a[BW], anya = 0;
                {
                    const uint4 z0 = *(const uint4 *)&M.nz4[c4][0];
                    a[0] = q[0] & z0.x; a[1] = q[1] & z0.y; a[2] = q[2] & z0.z; a[3] = q[3] & z0.w;
                    if (BW == 8) {
                        const uint4 z1 = *(const uint4 *)&M.nz4[c4][4];
                        a[BW - 4] = q[BW - 4] & z1.x; a[BW - 3] = q[BW - 3] & z1.y; a[BW - 2] = q[BW - 2] & z1.z; a[BW - 1] = q[BW - 1] & z1.w;
                    }
                }
                #pragma unroll
                for (int g = 0; g < BW; ++g) anya |= a[g];
                if (!c_.pre_on) anya = 0;
                if (!(anya | (uint32_t)full)) continue;
                float U[4] = {0.f, 0.f, 0.f, 0.f};
                if (anya) {
                    #pragma unroll
                    for (int g = 0; g < BW; ++g) {
                        uint32_t mm = a[g];
                        while (mm) {  // live samples with a spike at pixel i, ascending
                            const int bb = g * 32 + __ffs(mm) - 1;
                            mm &= mm - 1;
                            const float4 t4 = *(const float4 *)(c_.tx + bb * TJ + 4 * c4);
                            U[0] = U[0] + t4.x; U[1] = U[1] + t4.y; U[2] = U[2] + t4.z; U[3] = U[3] + t4.w;
                        }
                    }
                    if (c_.reduce_mean) { U[0] = U[0] / c_.Bf; U[1] = U[1] / c_.Bf; U[2] = U[2] / c_.Bf; U[3] = U[3] / c_.Bf; }
                }
                const uint32_t gwin = (colwin >> (4 * c4)) & 0xFu;  // these columns finish in the post step
                float *wp = wrow + 4 * c4;
                const float4 w4 = *(const float4 *)wp;
                float wv[4] = {w4.x, w4.y, w4.z, w4.w};
                #pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float w = wv[c];
                    if (!c_.wdep) {
                        // PostPre family: w - U*dt (x * 1.0f is exact, so the classic rule's missing dt
                        // factor is dts = 1)
                        if (anya) w = w - U[c] * c_.dts;
                    } else {
                        // WeightDependentPostPre, pre term only (learning.py:641-644, 651)
                        float upd = 0.0f;
                        if (c_.nu0 != 0.0f) upd = upd - (c_.nu0 * (anya ? U[c] : 0.0f)) * (w - c_.wmin);
                        if (c_.nu1 != 0.0f) upd = upd + (c_.nu1 * 0.0f) * (c_.wmax - w);
                        w = w + upd;
                    }
                    if (!((gwin >> c) & 1u)) {
                        if (c_.weight_decay != 0.0f) w = w * c_.weight_decay;
                        if (c_.has_clamp) w = clampf(w, c_.wmin, c_.wmax);
                    }
                    wv[c] = w;
                }
                *(float4 *)wp = make_float4(wv[0], wv[1], wv[2], wv[3]);
            }
        }
        #pragma unroll 1
        for (int k = 0; k < nwl; ++k) {
            const uint32_t e = M.wl[k];  // column << 16 | sample << 8 | staged row slot
            float *wp = wrow + (e >> 16);
            float w = *wp;
            const float V = 0.0f + c_.xrow[(e & 0xffu) * P + i] * c_.nu1;
            w = w + V * c_.dts;
            if (c_.weight_decay != 0.0f) w = w * c_.weight_decay;
            if (c_.has_clamp) w = clampf(w, c_.wmin, c_.wmax);
            *wp = w;
        }
    }
}

// Early STDP of one step (pre term only) in list form: the work items are (live (sample, group)
// pair, event of that sample) — far fewer than rows x groups while few pairs are live.  Several
// samples can spike at the same pixel: the item whose sample is the LOWEST live one at that pixel
// owns the row (no atomics), sums all of them in ascending order and updates the group's 4 columns:
// w - U*dt, decay, clamp (MCC_learning.py:234-263, 86-110).  Threads tid0 < 0 (warp 0, busy arriving
// at the grid barrier) do not take part.
constexpr int EVH = 16;  // list slots enumerated per pair and round
template <int TJ, int BW>
__device__ __noinline__ void stdp_list(const PassCtx *cx, int sb, uint32_t groups, uint32_t colwin, int tid0, int nthr0) {
    constexpr int CG = TJ / 4, WS = TJ + 4;
    const PassCtx c_ = load_ctx(cx);
    const int P = c_.P;
    const Misc &M = *c_.M;
    const uint16_t *ec = (const uint16_t *)(c_.evb + sb * c_.evblk);
    const uint16_t *el = (const uint16_t *)(c_.evb + sb * c_.evblk + c_.cntb);
    const uint4 *cT = (const uint4 *)(c_.inT + sb * P * BW);
    const int total = M.nlive * EVH;
    if (tid0 < 0) return;
    #pragma unroll 1
    for (int idx = tid0; idx < total; idx += nthr0) {
        const int lp = c_.live[idx / EVH];
        const int bb = lp / CG, c4 = lp % CG;
        if (!((groups >> c4) & 1u)) continue;
        const int cnt = min((int)ec[bb], EV_CAP);
        #pragma unroll 1
        for (int k = idx % EVH; k < cnt; k += EVH) {
            const int i = el[bb * EV_CAP + k];
            uint32_t a[BW];
            {
                const uint4 q0 = cT[i * (BW / 4)];
                const uint4 z0 = *(const uint4 *)&M.nz4[c4][0];
                a[0] = q0.x & z0.x; a[1] = q0.y & z0.y; a[2] = q0.z & z0.z; a[3] = q0.w & z0.w;
                if (BW == 8) {
                    const uint4 q1 = cT[i * (BW / 4) + 1];
                    const uint4 z1 = *(const uint4 *)&M.nz4[c4][4];
                    a[BW - 4] = q1.x & z1.x; a[BW - 3] = q1.y & z1.y; a[BW - 2] = q1.z & z1.z; a[BW - 1] = q1.w & z1.w;
                }
            }
            // owner of row i in this group = the lowest live sample spiking at pixel i
            uint32_t lower = 0;  // any live spiking sample below bb?
            #pragma unroll
            for (int g = 0; g < BW; ++g) {
                const uint32_t below = g < (bb >> 5) ? 0xffffffffu : (g == (bb >> 5) ? ((1u << (bb & 31)) - 1u) : 0u);
                lower |= a[g] & below;
            }
            if (lower) continue;
            float U0 = 0.f, U1 = 0.f, U2 = 0.f, U3 = 0.f;
            #pragma unroll
            for (int g = 0; g < BW; ++g) {
                uint32_t mm = a[g];
                while (mm) {
                    const int b2 = g * 32 + __ffs(mm) - 1;
                    mm &= mm - 1;
                    const float4 t4 = *(const float4 *)(c_.tx + b2 * TJ + 4 * c4);
                    U0 = U0 + t4.x; U1 = U1 + t4.y; U2 = U2 + t4.z; U3 = U3 + t4.w;
                }
            }
            if (c_.reduce_mean) { U0 = U0 / c_.Bf; U1 = U1 / c_.Bf; U2 = U2 / c_.Bf; U3 = U3 / c_.Bf; }
            float *wp = c_.W + i * WS + 4 * c4;
            const float4 w4 = *(const float4 *)wp;
            float wv[4] = {w4.x, w4.y, w4.z, w4.w};
            const float Uv[4] = {U0, U1, U2, U3};
            const uint32_t gwin = (colwin >> (4 * c4)) & 0xFu;  // columns with a pending post term finish there
            #pragma unroll
            for (int c = 0; c < 4; ++c) {
                float w = wv[c];
                if (!c_.wdep) {
                    w = w - Uv[c] * c_.dts;
                } else {  // WeightDependentPostPre, pre term only (learning.py:641-644, 651)
                    float upd = 0.0f;
                    if (c_.nu0 != 0.0f) upd = upd - (c_.nu0 * Uv[c]) * (w - c_.wmin);
                    if (c_.nu1 != 0.0f) upd = upd + (c_.nu1 * 0.0f) * (c_.wmax - w);
                    w = w + upd;
                }
                if (!((gwin >> c) & 1u)) {
                    if (c_.weight_decay != 0.0f) w = w * c_.weight_decay;
                    if (c_.has_clamp) w = clampf(w, c_.wmin, c_.wmax);
                }
                wv[c] = w;
            }
            *(float4 *)wp = make_float4(wv[0], wv[1], wv[2], wv[3]);
        }
    }
}

// Late STDP, general form (rare: more than 4 winners in the tile, two winners in one column, a
// winner whose input-trace row is not staged, WeightDependentPostPre, mean reduction, or a full
// pass): one row loop per candidate column group, pre and post term of a column applied together.
template <int TJ, int BW>
__device__ __noinline__ void late_generic_fn(const PassCtx *cx, const snn_conn_t *Cp, int sb, int ppar, uint32_t lategrp, uint32_t colwin,
                                             int full, const float *xsrc) {
    constexpr int WS = TJ + 4;
    const PassCtx c_ = load_ctx(cx);
    const snn_conn_t &C = *Cp;
    const Misc &M = *c_.M;
    const int P = c_.P, tid = threadIdx.x, nthr = blockDim.x;
    const bool wdep = c_.wdep != 0, pre_on = c_.pre_on != 0;
    const float Bf = c_.Bf, dts = c_.dts;
    float *W = c_.W;
    const float *tx = c_.tx, *xrow = c_.xrow;
    const uint4 *cTl = (const uint4 *)(c_.inT + sb * P * BW);
    const int ns = min(M.ncand[ppar], XR);
    #pragma unroll 1
    for (uint32_t lg = lategrp; lg; lg &= lg - 1) {
        const int c4 = __ffs(lg) - 1;
        const uint32_t gwin = (colwin >> (4 * c4)) & 0xFu;
        // staged trace-row offset of each winner (ascending sample order) per winner column
        int nwin[4] = {0, 0, 0, 0}, wrow[4][2];
        bool generic = false;
        #pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            if (!((gwin >> c) & 1u)) continue;
            #pragma unroll 1
            for (int g = 0; g < BW; ++g) {
                uint32_t mm = M.wmask[4 * c4 + c][g];
                while (mm) {
                    const int bb = g * 32 + __ffs(mm) - 1;
                    mm &= mm - 1;
                    int sl = -1;
                    for (int q = 0; q < ns; ++q) if (M.candb[ppar][q] == bb) sl = q;
                    if (nwin[c] < 2 && sl >= 0) wrow[c][nwin[c]] = sl * P; else generic = true;
                    ++nwin[c];
                }
            }
        }
        const uint4 z0 = pre_on ? *(const uint4 *)&M.nz4[c4][0] : make_uint4(0, 0, 0, 0);
        #pragma unroll 1
        for (int i = tid; i < P; i += nthr) {
            uint32_t m[BW];
            const uint4 q0 = cTl[i * (BW / 4)];
            m[0] = q0.x & z0.x; m[1] = q0.y & z0.y; m[2] = q0.z & z0.z; m[3] = q0.w & z0.w;
            uint32_t anym = m[0] | m[1] | m[2] | m[3];
            if (BW == 8) {
                const uint4 q1 = cTl[i * (BW / 4) + 1];
                const uint4 z1 = pre_on ? *(const uint4 *)&M.nz4[c4][4] : make_uint4(0, 0, 0, 0);
                m[BW - 4] = q1.x & z1.x; m[BW - 3] = q1.y & z1.y; m[BW - 2] = q1.z & z1.z; m[BW - 1] = q1.w & z1.w;
                anym |= m[BW - 4] | m[BW - 3] | m[BW - 2] | m[BW - 1];
            }
            const bool pre_t = anym != 0u;
            if (!(pre_t || gwin || full)) continue;
            float U[4] = {0.f, 0.f, 0.f, 0.f};
            if (pre_t) {
                #pragma unroll 1
                for (int g = 0; g < BW; ++g) {
                    uint32_t mm = m[g];
                    while (mm) {
                        const int bb = g * 32 + __ffs(mm) - 1;
                        mm &= mm - 1;
                        const float4 t4 = *(const float4 *)(tx + bb * TJ + 4 * c4);
                        U[0] = U[0] + t4.x; U[1] = U[1] + t4.y; U[2] = U[2] + t4.z; U[3] = U[3] + t4.w;
                    }
                }
                if (C.reduction == SNN_REDUCE_MEAN) { U[0] = U[0] / Bf; U[1] = U[1] / Bf; U[2] = U[2] / Bf; U[3] = U[3] / Bf; }
            }
            float *wp = W + i * WS + 4 * c4;
            const float4 w4 = *(const float4 *)wp;
            float wv[4] = {w4.x, w4.y, w4.z, w4.w};
            #pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const bool post_t = (gwin >> c) & 1u;
                float V = 0.0f;
                if (post_t) {
                    if (!generic) {
                        V = V + xrow[wrow[c][0] + i] * (wdep ? 1.0f : C.nu1);
                        if (nwin[c] > 1) V = V + xrow[wrow[c][1] + i] * (wdep ? 1.0f : C.nu1);
                    } else {  // more winners than staged rows: read them from L2 (rare)
                        for (int g = 0; g < BW; ++g) {
                            uint32_t mm = M.wmask[4 * c4 + c][g];
                            while (mm) {
                                const int bb = g * 32 + __ffs(mm) - 1;
                                mm &= mm - 1;
                                V = V + __ldcg(xsrc + (size_t)bb * P + i) * (wdep ? 1.0f : C.nu1);
                            }
                        }
                    }
                    if (C.reduction == SNN_REDUCE_MEAN) V = V / Bf;
                }
                if (!wdep) {
                    float w = wv[c];
                    if (pre_t) w = w - U[c] * dts;
                    if (post_t) w = w + V * dts;
                    if (C.weight_decay != 0.0f) w = w * C.weight_decay;
                    if (C.has_clamp) w = clampf(w, C.wmin, C.wmax);
                    wv[c] = w;
                } else {
                    wv[c] = apply_rule(C, wv[c], U[c], pre_t, V, post_t);
                }
            }
            *(float4 *)wp = make_float4(wv[0], wv[1], wv[2], wv[3]);
        }
    }
}

// Spike-gather of a sample whose event list overflowed EV_CAP: walk its bit row in global memory
// (rare; out of line to keep the hot loop small).
__device__ __noinline__ float4 gather_dense(const uint32_t *row, int SW, const float *Wc, int WS) {
    Wc = sh(Wc);
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
    for (int w = 0; w < SW; ++w) {
        uint32_t word = __ldg(row + w);
        while (word) {
            const int i = w * 32 + __ffs(word) - 1;
            word &= word - 1;
            const float4 r4 = *(const float4 *)(Wc + i * WS);
            p0 = p0 + r4.x; p1 = p1 + r4.y; p2 = p2 + r4.z; p3 = p3 + r4.w;
        }
    }
    return make_float4(p0, p1, p2, p3);
}

// TJ: neurons per CTA (4 per thread); BW: 32-bit words of a per-pixel sample mask (4 -> B <= 128,
// 8 -> B <= 256).  Threads = B * TJ/4 <= 32 * BW * TJ/4.
//
// Software pipeline of one timestep t (the grid barrier is split into arrive / wait so that its
// L2 round trips overlap the bulk of the STDP work):
//   wait(t-1) | late(t-1): exchange read, winners, traces + STDP of column groups that held a
//   candidate | D(t): gather, Ae/Ai update, candidates -> atomics | arrive(t) | early(t): traces +
//   STDP pre term of the column groups WITHOUT a candidate (their step-t state is already final),
//   input trace of step t+1 published
// VAR bit 0 (lean): the rarely used options — additive traces, voltage lower bounds, voltage
// monitors, WeightDependentPostPre, mean reduction — are compiled out (match() proves them off),
// which keeps the per-step code inside the instruction cache.  VAR bit 1: phase timers compiled in.
template <int TJ, int BW, int VAR>
__global__ void __launch_bounds__((8 * BW * TJ < 1024 ? 8 * BW * TJ : 1024), 1)
snn_dc_fused_window(const __grid_constant__ FusedParams Q0) {
    constexpr bool LEAN = (VAR & 1) != 0, PROFV = (VAR & 2) != 0;
    // a local copy whose option fields are compile-time constants in the lean variant
    FusedParams Q = Q0;
    if (LEAN) {
        Q.X.traces_additive = 0; Q.E.traces_additive = 0; Q.E.has_lbound = 0; Q.I.has_lbound = 0;
        Q.E.rec_v = nullptr; Q.I.rec_v = nullptr; Q.C.reduction = SNN_REDUCE_SUM;
        if (Q.C.rule == SNN_RULE_WDEP_POSTPRE) Q.C.rule = SNN_RULE_POSTPRE;
    }
    constexpr int CG = TJ / 4;  // float4 column groups = lanes that share one sample
    constexpr int WS = TJ + 4;  // row stride of the W tile in shared memory (floats)
#ifdef SNN_EMU
    unsigned char *smem = (unsigned char *)emu::tls_cta->dyn_smem;
#else
    extern __shared__ __align__(16) unsigned char smem[];
#endif
    const int B = Q.B, P = Q.P, n = Q.n, T = Q.T;
    const unsigned int G = gridDim.x;
    const int own = (B + (int)G - 1) / (int)G;
    // shared-memory carve-up: byte offsets computed by the host (smem_layout) and passed as kernel
    // parameters, so that a pointer the compiler chooses to rematerialise costs one add
    float *W = (float *)(smem + Q.o_W);
    float *tx = (float *)(smem + Q.o_tx);
    unsigned char *evb = smem + Q.o_ev;
    uint32_t *inT = (uint32_t *)(smem + Q.o_inT);
    float *xrow = (float *)(smem + Q.o_xrow);
    float *rep = (float *)(smem + Q.o_rep);
    float *xown = (float *)(smem + Q.o_xown);
    float *theta_s = (float *)(smem + Q.o_theta);
    uint16_t *live = (uint16_t *)(smem + Q.o_live);
    Misc &M = *(Misc *)(smem + Q.o_misc);
#ifdef SNN_EMU
    PassCtx &s_cx = *(PassCtx *)emu::tls_cta->static_smem;
#else
    __shared__ PassCtx s_cx;
#endif

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int b = tid / CG, cg = tid % CG;   // state ownership: sample b, neurons jc..jc+3
    const bool act = b < B;
    const int j0 = blockIdx.x * TJ;
    const int jc = j0 + 4 * cg;
    const snn_layer_t &E = Q.E, &I = Q.I, &X = Q.X;
    const snn_conn_t &C = Q.C;
    const bool stdp = C.rule >= SNN_RULE_POSTPRE;
    const bool wdep = C.rule == SNN_RULE_WDEP_POSTPRE;
    const bool pre_on = stdp && C.nu0 != 0.0f, post_on = stdp && C.nu1 != 0.0f;
    const bool decay_on = C.weight_decay != 0.0f && C.weight_decay != 1.0f;
    const bool update_on = Q.learning && C.rule != SNN_RULE_NONE && (stdp || decay_on) && !(Q.dbg & 2);
    const bool stage_on = update_on && post_on && X.traces;
    const float Bf = (float)B;
    const float dts = C.rule == SNN_RULE_MCC_POSTPRE ? C.dt_scale : 1.0f;
    const int evblk = (int)al16((size_t)Q.SB);  // stride between the two staged list blocks
    const int cntb = ev_count_bytes(B);         // bytes of the count array inside a block
    float *Wc = W + 4 * cg;                     // my 4 columns of row 0
    // an Ai neuron at rest with no input stays bitwise at rest (decay*(rest-rest)+rest == rest)
    const bool ai_rest_ok = I.rest < I.thresh && !(I.has_lbound && I.rest < I.lbound);
    unsigned int gen = 0;
    long long *pc = M.pc;  // phase timers of thread 0 (profiling variant only)
    #pragma unroll
    for (int k = 0; k < NPROF; ++k) if (PROFV && tid == 0) pc[k] = 0;
    long long pt = clock64();
    #define PROF(k) { if (PROFV && Q.prof && tid == 0) { const long long now_ = clock64(); pc[k] += now_ - pt; pt = now_; } }

    // ---- prologue: W tile, theta, inhibition table, owned input traces, state registers ----
    #pragma unroll 1
    for (int idx = tid; idx < (P + 1) * TJ; idx += nthr) {
        const int i = idx / TJ, jj = idx - i * TJ;
        W[i * WS + jj] = (i < P && j0 + jj < n) ? C.w[(size_t)i * n + j0 + jj] : 0.0f;
    }
    #pragma unroll 1
    for (int jj = tid; jj < 32; jj += nthr) theta_s[jj] = (jj < TJ && j0 + jj < n) ? E.theta[j0 + jj] : 0.0f;
    if (tid == 0) {
        // rep[m] = m-fold sequential sum of the Ai->Ae weight: what the reference's dense sum
        // over k of sI[b,k] * w_ie[k,j] evaluates to when m inhibitory neurons (other than j) spike
        float a = 0.0f;
        rep[0] = 0.0f;
        #pragma unroll 1
        for (int m = 1; m <= n; ++m) { a = a + Q.inh_neg; rep[m] = a; }
        mbar_init(&M.mbar[0], 1);
        mbar_init(&M.mbar[1], 1);
#ifndef SNN_EMU
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
        M.ncand[0] = M.ncand[1] = 0; M.candgrp[0] = M.candgrp[1] = 0; M.colwin = 0; M.abort = 0; M.nwl = 0; M.nlive = 0;
        M.denseflag[0] = Q.dense[0]; M.denseflag[1] = T >= 1 ? Q.dense[1] : 0;
        for (int k = 0; k < 8; ++k) M.lc[k] = 0;
        s_cx.W = W; s_cx.tx = tx; s_cx.xrow = xrow; s_cx.inT = inT; s_cx.evb = evb; s_cx.live = live; s_cx.M = &M;
        s_cx.P = P; s_cx.B = B; s_cx.evblk = evblk; s_cx.cntb = cntb;
        s_cx.pre_on = pre_on; s_cx.wdep = wdep; s_cx.reduce_mean = C.reduction == SNN_REDUCE_MEAN; s_cx.has_clamp = C.has_clamp;
        s_cx.Bf = Bf; s_cx.dts = dts; s_cx.weight_decay = C.weight_decay; s_cx.wmin = C.wmin; s_cx.wmax = C.wmax;
        s_cx.nu0 = C.nu0; s_cx.nu1 = C.nu1;
    }
    #pragma unroll 1
    for (int k = tid; k < 64; k += nthr) { (&M.nz4[0][0])[k] = 0; (&M.cnt[0][0])[k] = 0; }
    #pragma unroll 1
    for (int k = tid; k < 32 * 8; k += nthr) (&M.wmask[0][0])[k] = 0;
    if (X.traces)
        for (int o = 0; o < own; ++o) {
            const int bo = blockIdx.x + o * (int)G;
            if (bo < B)
                #pragma unroll 1
                for (int i = tid; i < P; i += nthr) xown[o * P + i] = X.x[(size_t)bo * P + i];
        }

    float vE[4], rE[4], xE[4], vI[4], rI[4];
    uint32_t sEprev = 0, sIprev = 0, candE = 0;  // 4-bit masks over my neurons
    uint32_t cE01 = 0, cE23 = 0, cI01 = 0, cI23 = 0;  // spike counters of my neurons, 16 bits each
    #pragma unroll
    for (int c = 0; c < 4; ++c) {
        const bool ok = act && jc + c < n;
        const size_t k = ok ? (size_t)b * n + jc + c : 0;
        vE[c] = ok ? E.v[k] : 0.0f; rE[c] = ok ? E.refrac_count[k] : 0.0f;
        xE[c] = (ok && E.traces) ? E.x[k] : 0.0f;
        vI[c] = ok ? I.v[k] : 0.0f; rI[c] = ok ? I.refrac_count[k] : 0.0f;
        if (ok && E.s[k]) sEprev |= 1u << c;
        if (ok && I.s[k]) sIprev |= 1u << c;
    }
    if (act && stdp) {
        *(float4 *)(tx + b * TJ + 4 * cg) =
            make_float4(wdep ? xE[0] : xE[0] * C.nu0, wdep ? xE[1] : xE[1] * C.nu0, wdep ? xE[2] : xE[2] * C.nu0,
                        wdep ? xE[3] : xE[3] * C.nu0);
    }
    __syncthreads();
    bool livep = false;  // my (sample, column group) pair has a non-zero Ae trace
    if (act && stdp && (xE[0] != 0.0f || xE[1] != 0.0f || xE[2] != 0.0f || xE[3] != 0.0f)) {
        livep = true;
        atomicOr(&M.nz4[cg][b >> 5], 1u << (b & 31));
        live[atomicAdd(&M.nlive, 1)] = (uint16_t)tid;
    }

    const uint32_t bytesE = (uint32_t)Q.SB, bytesT = (uint32_t)(sizeof(uint32_t) * (size_t)P * BW);
    if (tid == 0) {  // stage slot 0 (spikes of step -1) and slot 1 (spikes of step 0)
        mbar_expect_tx(&M.mbar[0], bytesE + bytesT);
        bulk_g2s(evb, Q.evS, bytesE, &M.mbar[0]);
        bulk_g2s(inT, Q.inT, bytesT, &M.mbar[0]);
        mbar_expect_tx(&M.mbar[1], bytesE + bytesT);
        bulk_g2s(evb + evblk, Q.evS + Q.SB, bytesE, &M.mbar[1]);
        bulk_g2s(inT + P * BW, Q.inT + (size_t)P * BW, bytesT, &M.mbar[1]);
    }
    // input trace of step 0 for the samples this CTA owns, published in xpub slot 0
    // Warp 0 is busy with the grid barrier while the others publish: work is spread over the threads
    // ptid = tid - 32 (all threads when the block is a single warp).  The spike bits of the first
    // item of each of the first two owned samples can be loaded ahead (publish_load) so that their
    // L2 latency hides behind the early STDP.
    const int ptid = nthr > 32 ? tid - 32 : tid, pn = nthr > 32 ? nthr - 32 : nthr;
    const int tid_pf = nthr > 32 ? 32 : 0;  // the thread that issues the bulk prefetches (not thread 0: it arrives at the barrier)
    uint32_t pubw0 = 0, pubw1 = 0;
    auto publish_load = [&](int step) {
        if (!X.traces || (Q.dbg & 8) || ptid < 0 || ptid >= (P >> 2)) return;
        const int b0_ = blockIdx.x, b1_ = blockIdx.x + (int)G;
        if (b0_ < B) pubw0 = __ldg(Q.inS + ((size_t)(step + 1) * B + b0_) * Q.SW + (ptid >> 3));
        if (own > 1 && b1_ < B) pubw1 = __ldg(Q.inS + ((size_t)(step + 1) * B + b1_) * Q.SW + (ptid >> 3));
    };
    auto publish_trace = [&](int step) {
        if (!X.traces || (Q.dbg & 8) || ptid < 0) return;
        for (int o = 0; o < own; ++o) {
            const int bo = blockIdx.x + o * (int)G;
            if (bo < B) {
                const uint32_t *srow = Q.inS + ((size_t)(step + 1) * B + bo) * Q.SW;
                float4 *dst = (float4 *)(Q.xpub + ((size_t)(step % 3) * B + bo) * P);
                float4 *xo = (float4 *)(xown + o * P);
                #pragma unroll 1
                for (int i4 = ptid; i4 < (P >> 2); i4 += pn) {
                    const uint32_t word = (i4 == ptid && o < 2) ? (o == 0 ? pubw0 : pubw1) : __ldg(srow + (i4 >> 3));
                    const uint32_t bits = word >> ((i4 & 7) * 4);
                    float4 x = xo[i4];
                    x.x = trace_step(x.x, bits & 1u, X.trace_decay, X.trace_scale, X.traces_additive);
                    x.y = trace_step(x.y, bits & 2u, X.trace_decay, X.trace_scale, X.traces_additive);
                    x.z = trace_step(x.z, bits & 4u, X.trace_decay, X.trace_scale, X.traces_additive);
                    x.w = trace_step(x.w, bits & 8u, X.trace_decay, X.trace_scale, X.traces_additive);
                    xo[i4] = x;
                    dst[i4] = x;
                }
            }
        }
    };
    publish_load(0);
    publish_trace(0);
    __syncthreads();

    // spike-gather of my 4 columns for the spikes of list block `blk` (slot `slot`):
    // p[c] = sum_{i in sX[b]} W[i][c], i ascending (topology.py:437-479)
    auto gather = [&](const unsigned char *blk, int slot) -> float4 {
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
        const int cnt = ((const uint16_t *)blk)[b];
        if (Q.dbg & 4) {
        } else if (cnt <= EV_CAP) {
            const uint2 *l4 = (const uint2 *)(blk + cntb + b * (2 * EV_CAP));
            #pragma unroll 1
            for (int k = 0; k < cnt; k += 4) {
                const uint2 q = l4[k >> 2];  // 4 pixel indices; tail padded with P (zero row)
                const float4 r0 = *(const float4 *)(Wc + (q.x & 0xffffu) * WS);
                const float4 r1 = *(const float4 *)(Wc + (q.x >> 16) * WS);
                const float4 r2 = *(const float4 *)(Wc + (q.y & 0xffffu) * WS);
                const float4 r3 = *(const float4 *)(Wc + (q.y >> 16) * WS);
                p0 = p0 + r0.x; p1 = p1 + r0.y; p2 = p2 + r0.z; p3 = p3 + r0.w;
                p0 = p0 + r1.x; p1 = p1 + r1.y; p2 = p2 + r1.z; p3 = p3 + r1.w;
                p0 = p0 + r2.x; p1 = p1 + r2.y; p2 = p2 + r2.z; p3 = p3 + r2.w;
                p0 = p0 + r3.x; p1 = p1 + r3.y; p2 = p2 + r3.z; p3 = p3 + r3.w;
            }
        } else {  // dense sample: walk the bit row in global memory (rare, slow path)
            return gather_dense(Q.inS + ((size_t)slot * B + b) * Q.SW, Q.SW, Wc, WS);
        }
        return make_float4(p0, p1, p2, p3);
    };
    int myslot = -1;           // staged input-trace row of my sample's candidates (step being finalised)

    // =====================================================================================
    PROF(0)  // prologue
    uint32_t pend = 0;         // my candidates of the step being finalised are still undecided
    int dflag = 0;
    for (int t = 0; t <= T; ++t) {
        const int buf = t & 1;                                       // slot t   = spikes of step t-1
        const unsigned char *cE = evb + buf * evblk;
        const int par = t & 1, ppar = par ^ 1;                      // parity of step t / of step t-1

        // ---- late(t-1): exchange results of step t-1 (slot (t-1) % 3; t = 0: from the pre-pass)
        const int xs = (t + 2) % 3;
        unsigned long long key = 0ull;
        if (act && t > 0 && pend) key = __ldcg(Q.win + xs * B + b);
        // Ai spike count of step t-1 (lateral inhibition): issued here, consumed by the neuron update
        const unsigned int isum = act ? __ldcg(Q.sisum + xs * B + b) : 0u;
        const uint32_t lategrp = t > 0 ? M.candgrp[ppar] : 0u;      // groups that held a candidate at t-1
        // stage the input-trace rows of this tile's candidate samples (the possible winners) in
        // shared memory: issued together with the exchange loads, so one L2 round trip covers both
        if (t > 0 && stage_on && lategrp && !(Q.dbg & 16)) {
            const int ns = min(M.ncand[ppar], XR);
            const int P4 = P >> 2;
            const float4 *xsrc = (const float4 *)(Q.xpub + (size_t)((t - 1) % 3) * B * P);
            for (int idx = tid; idx < ns * P4; idx += nthr) {
                const int r = idx / P4, i4 = idx - r * P4;
                ((float4 *)xrow)[idx] = __ldcg(xsrc + (size_t)M.candb[ppar][r] * P4 + i4);
            }
        }
        PROF(1)  // exchange loads + staging issue
        if (t > 0) {
            uint32_t sE = 0;
            if (act) {
                if (pend) {  // candidates of step t-1: winner (nodes.py:1097-1105), then the trace
                    if (E.one_spike) {
                        if (key != 0ull) {
                            const int wj = (int)(uint32_t)(key & 0xffffffffull) - jc;
                            if (wj >= 0 && wj < 4 && ((candE >> wj) & 1u)) sE = 1u << wj;
                        }
                    } else sE = candE;
                    if (E.traces) {
                        #pragma unroll
                        for (int c = 0; c < 4; ++c) xE[c] = trace_step(xE[c], (sE >> c) & 1u, E.trace_decay, E.trace_scale, E.traces_additive);
                    }
                    if (update_on && stdp) {
                        if (xE[0] != 0.0f || xE[1] != 0.0f || xE[2] != 0.0f || xE[3] != 0.0f) {
                            *(float4 *)(tx + b * TJ + 4 * cg) =
                                make_float4(wdep ? xE[0] : xE[0] * C.nu0, wdep ? xE[1] : xE[1] * C.nu0,
                                            wdep ? xE[2] : xE[2] * C.nu0, wdep ? xE[3] : xE[3] * C.nu0);
                            if (!livep) {
                                livep = true;
                                atomicOr(&M.nz4[cg][b >> 5], 1u << (b & 31));
                                live[atomicAdd(&M.nlive, 1)] = (uint16_t)tid;
                            }
                        }
                        if (sE && post_on) {
                            #pragma unroll
                            for (int c = 0; c < 4; ++c)
                                if ((sE >> c) & 1u) {
                                    atomicOr(&M.wmask[4 * cg + c][b >> 5], 1u << (b & 31));
                                    atomicOr(&M.colwin, 1u << (4 * cg + c));
                                    const int k = atomicAdd(&M.nwl, 1);
                                    if (k < 4) M.wl[k] = ((uint32_t)(4 * cg + c) << 16) | ((uint32_t)b << 8) | (myslot >= 0 ? (uint32_t)myslot : 0xffu);
                                }
                        }
                    }
                    pend = 0;
                }
                // monitors (monitors.py:94-111): spikes / voltages of step t-1
                if (E.rec_s || I.rec_s || E.rec_v || I.rec_v) {
                    const size_t k0 = ((size_t)(t - 1) * B + b) * n + jc;
                    if (((n & 3) == 0) && jc + 3 < n) {  // 4 rasters bytes in one store
                        if (E.rec_s) *(uint32_t *)(E.rec_s + k0) = (sE & 1u) | ((sE & 2u) << 7) | ((sE & 4u) << 14) | ((sE & 8u) << 21);
                        if (I.rec_s) *(uint32_t *)(I.rec_s + k0) = (sIprev & 1u) | ((sIprev & 2u) << 7) | ((sIprev & 4u) << 14) | ((sIprev & 8u) << 21);
                        if (E.rec_v) *(float4 *)(E.rec_v + k0) = make_float4(vE[0], vE[1], vE[2], vE[3]);
                        if (I.rec_v) *(float4 *)(I.rec_v + k0) = make_float4(vI[0], vI[1], vI[2], vI[3]);
                    } else {
                        #pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (jc + c < n) {
                                if (E.rec_s) E.rec_s[k0 + c] = (sE >> c) & 1u;
                                if (I.rec_s) I.rec_s[k0 + c] = (sIprev >> c) & 1u;
                                if (E.rec_v) E.rec_v[k0 + c] = vE[c];
                                if (I.rec_v) I.rec_v[k0 + c] = vI[c];
                            }
                    }
                }
            }
            sEprev = sE;
            if (E.rec_count) { cE01 += (sE & 1u) | ((sE & 2u) << 15); cE23 += ((sE >> 2) & 1u) | ((sE & 8u) << 13); }
            if (I.rec_count) { cI01 += (sIprev & 1u) | ((sIprev & 2u) << 15); cI23 += ((sIprev >> 2) & 1u) | ((sIprev & 8u) << 13); }
        }
        PROF(12)  // late finalise (winner, trace, monitors)
        if (t > 0 && update_on && lategrp) {
            // STDP of step t-1 for the column groups that held a candidate (MCC_learning.py:234-299).
            __syncthreads();
            const uint32_t colwin = post_on ? M.colwin : 0u;
            const bool full = decay_on || (C.has_clamp && t == 1);
            PROF(13)  // late sync
            // Typical case — at most 4 winners in the tile, in distinct columns, their input-trace rows
            // staged: the row pass shared with the early STDP (its code is resident in the instruction
            // cache; this path gates the whole grid).  Anything else takes the general form.
            const int nwl = post_on ? M.nwl : 0;
            bool fast = !wdep && C.reduction == SNN_REDUCE_SUM && !full && nwl <= 4 && nwl == __popc(colwin);
            if (fast && nwl) {
                const uint4 wl4 = *(const uint4 *)M.wl;  // column << 16 | sample << 8 | row slot (0xff: not staged)
                if ((wl4.x & 0xffu) == 0xffu || (nwl > 1 && (wl4.y & 0xffu) == 0xffu) || (nwl > 2 && (wl4.z & 0xffu) == 0xffu) ||
                    (nwl > 3 && (wl4.w & 0xffu) == 0xffu))
                    fast = false;
            }
            if (PROFV && tid == 0) {
                long long *lc = M.lc;
                lc[0] += 1; lc[1] += fast ? 1 : 0;
                lc[4] += colwin ? 1 : 0; lc[6] += M.ncand[ppar];
                if (fast) lc[2] += clock64() - pt;   // set-up
            }
            if (fast && !M.denseflag[buf]) {
                // pre term through the list pass the early STDP uses every step (its code is resident),
                // then the post term: one scalar update per (row, winner column); constants from shared
                // memory (a rarely used kernel-parameter line costs a constant-cache miss per use here)
                if (pre_on) {
                    stdp_list<TJ, BW>(&s_cx, buf, lategrp, colwin, tid, nthr);
                    __syncthreads();
                }
                if (nwl) {
                    const float l_dts = s_cx.dts, l_nu1 = s_cx.nu1, l_wmin = s_cx.wmin, l_wmax = s_cx.wmax, l_wd = s_cx.weight_decay;
                    const int l_clamp = s_cx.has_clamp, l_P = s_cx.P;
                    #pragma unroll 1
                    for (int k = 0; k < nwl; ++k) {
                        const uint32_t e = M.wl[k];  // column << 16 | sample << 8 | staged row slot
                        float *wcol = W + (e >> 16);
                        const float *xr = xrow + (e & 0xffu) * l_P;
                        #pragma unroll 1
                        for (int i = tid; i < l_P; i += nthr) {
                            float w = wcol[i * WS];
                            const float V = 0.0f + xr[i] * l_nu1;
                            w = w + V * l_dts;
                            if (l_wd != 0.0f) w = w * l_wd;
                            if (l_clamp) w = clampf(w, l_wmin, l_wmax);
                            wcol[i * WS] = w;
                        }
                    }
                }
            } else if (fast) stdp_rows<TJ, BW>(&s_cx, buf, lategrp, nwl, colwin, 0);
            else late_generic_fn<TJ, BW>(&s_cx, &Q0.C, buf, ppar, lategrp, colwin, full ? 1 : 0, Q.xpub + (size_t)((t - 1) % 3) * B * P);
            if (PROFV && Q.prof && tid == 0) { const long long now_ = clock64(); if (fast) M.lc[7] += now_ - pt; }
            PROF(14)  // late group pass
            __syncthreads();
            if (colwin) {
                for (int k = tid; k < TJ * 8; k += nthr) (&M.wmask[0][0])[k] = 0;
                if (tid == 0) { M.colwin = 0; M.nwl = 0; }
            }
        }
        PROF(2)  // late finalise + late STDP
        if (t == T) break;

        // ---- D(t): gather, Ae / Ai update, candidates ----------------------------------------
        // this iteration's spike lists (slot t) were prefetched one iteration ago; buffer `buf` is
        // filled for the (t >> 1)-th time, which is its mbarrier phase
        while (!mbar_try_wait(&M.mbar[buf], (uint32_t)(t >> 1) & 1u)) {}
        if (tid == 0) { M.ncand[par] = 0; M.candgrp[par] = 0; }
        __syncthreads();   // late STDP done (W final for step t-1); bookkeeping of parity `par` reset
        PROF(3)  // mbarrier wait + sync
        uint32_t cand = 0, sI = 0;
        unsigned long long mykey = 0ull;
        int nI = 0;
        if (act) {
            // columns whose step-(t-1) STDP was late (or everything, first time): gather now
            // (running this gather one step ahead, in the shadow of the barrier, for the groups whose W
            // is already final was measured: no gain — the CTAs are busy, not idle, in that shadow)
            const float4 pg = gather(cE, t);
            const float p[4] = {pg.x, pg.y, pg.z, pg.w};
            const float4 th4 = *(const float4 *)(theta_s + 4 * cg);
            float th[4] = {th4.x, th4.y, th4.z, th4.w};
            #pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (E.learning) th[c] = th[c] * E.theta_decay;  // nodes.py:1078-1079 (stored after the step)
                if (jc + c < n) {
                    // network.py:225-248: X->Ae first, then Ai->Ae; the latter is rep[#spiking Ai other than j]
                    const int mI = (int)isum - (int)((sIprev >> c) & 1u);
                    float cur = 0.0f + p[c];
                    cur = cur + rep[mI];
                    if (dc_step(E, vE[c], rE[c], cur, th[c])) cand |= 1u << c;       // nodes.py:1077-1092
                    if (E.has_lbound && vE[c] < E.lbound) vE[c] = E.lbound;           // nodes.py:1108-1109
                    const bool inI = (sEprev >> c) & 1u;
                    if (inI || !ai_rest_ok || vI[c] != I.rest || rI[c] > 0.0f) {
                        float curI = inI ? (0.0f + Q.exc) : 0.0f;                     // diagonal Ae->Ai
                        if (lif_step(I, vI[c], rI[c], curI)) { sI |= 1u << c; ++nI; } // nodes.py:500-529
                    } else {
                        rI[c] = rI[c] - I.dt;  // resting, unrefractory, no input: v stays exactly at rest
                    }
                }
            }
            if (cand) {
                #pragma unroll
                for (int c = 0; c < 4; ++c)
                    if ((cand >> c) & 1u) {
                        atomicAdd(&M.cnt[par][4 * cg + c], 1);
                        if (E.one_spike) {
                            const unsigned long long k2 = snn_one_spike_key(Q.seed, (uint32_t)t + Q.step_offset, (uint32_t)Q.liE,
                                                                            (uint32_t)b, (uint32_t)(jc + c));
                            mykey = k2 > mykey ? k2 : mykey;
                        }
                    }
                atomicOr(&M.candgrp[par], 1u << cg);
            } else if (E.traces) {
                // no candidate among my neurons: their step-t trace is already final (no spike)
                #pragma unroll
                for (int c = 0; c < 4; ++c) xE[c] = trace_step(xE[c], false, E.trace_decay, E.trace_scale, E.traces_additive);
                if (update_on && stdp && livep)
                    *(float4 *)(tx + b * TJ + 4 * cg) =
                        make_float4(wdep ? xE[0] : xE[0] * C.nu0, wdep ? xE[1] : xE[1] * C.nu0,
                                    wdep ? xE[2] : xE[2] * C.nu0, wdep ? xE[3] : xE[3] * C.nu0);
            }
        }
        PROF(4)  // gather + neuron updates
        // reductions over the CG lanes that share a sample (all lanes of the warp take part)
        uint32_t anyc = cand;
        #pragma unroll
        for (int o = CG / 2; o > 0; o >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, mykey, o);
            mykey = ok > mykey ? ok : mykey;
            nI += __shfl_xor_sync(0xffffffffu, nI, o);
            anyc |= __shfl_xor_sync(0xffffffffu, anyc, o);
        }
        int slot_ = -1;
        if (act && cg == 0) {
            const int ws = t % 3;
            if (mykey) atomicMax(Q.win + ws * B + b, mykey);
            if (nI) atomicAdd(Q.sisum + ws * B + b, (unsigned int)nI);
            if (anyc && (stage_on || (PROFV && Q.prof))) {
                const int s = atomicAdd(&M.ncand[par], 1);
                if (s < XR) { M.candb[par][s] = b; slot_ = s; }
            }
        }
        myslot = __shfl_sync(0xffffffffu, slot_, (tid & 31) & ~(CG - 1));  // from the sample's first lane
        candE = cand;
        pend = cand;
        sIprev = sI;
        // clear the exchange slot step t+1 will accumulate into (last read before the previous barrier)
        if (blockIdx.x == 0)
            for (int k = tid; k < B; k += nthr) {
                Q.win[((t + 1) % 3) * B + k] = 0ull;
                Q.sisum[((t + 1) % 3) * B + k] = 0u;
            }
        __syncthreads();
        PROF(5)  // reductions, atomics, sync
        // theta = theta * decay + theta_plus * (#candidates of the column)  (nodes.py:1078-1094)
        if (tid < TJ) {
            if (E.learning) theta_s[tid] = theta_s[tid] * E.theta_decay + E.theta_plus * (float)M.cnt[par][tid];
            M.cnt[par][tid] = 0;
        }
        // prefetch slot t+2 into the buffer the gather just finished with
        if (tid == tid_pf && t + 2 <= T) {  // (not thread 0: it is about to arrive at the grid barrier)
            mbar_expect_tx(&M.mbar[buf], bytesE + bytesT);
            bulk_g2s(evb + buf * evblk, Q.evS + (size_t)(t + 2) * Q.SB, bytesE, &M.mbar[buf]);
            bulk_g2s(inT + buf * P * BW, Q.inT + (size_t)(t + 2) * P * BW, bytesT, &M.mbar[buf]);
            dflag = __ldg(Q.dense + t + 2);  // stored below, read by the early pass one step from now
        }
        // ---- arrive(t): this CTA's contributions to step t's exchange are issued -----------------
        PROF(6)  // theta, prefetch issue
#ifdef SNN_EMU
        if (tid == 0) __atomic_fetch_add(Q.bar, 1u, __ATOMIC_RELEASE);
#else
        if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(Q.bar) : "memory");
#endif
        gen += 1;
        PROF(15)  // arrive (release)
        // ---- early(t): in the shadow of the barrier ------------------------------------------------
        if (t + 1 < T) publish_load(t + 1);
        const int nb = buf ^ 1;                                      // slot t+1 = spikes of step t
        while (!mbar_try_wait(&M.mbar[nb], (uint32_t)((t + 1) >> 1) & 1u)) {}  // slot t+1 landed? (waited again by D(t+1))
        if (update_on) {
            const uint32_t allg = (1u << CG) - 1u;
            const uint32_t earlygrp = allg & ~M.candgrp[par];
            const bool full = decay_on || (C.has_clamp && t == 0);
            // list form while every sample's event list is complete; row form otherwise
            if (full || M.denseflag[nb]) stdp_rows<TJ, BW>(&s_cx, nb, earlygrp, 0, 0u, full ? 1 : 0);
            else if (pre_on) stdp_list<TJ, BW>(&s_cx, nb, earlygrp, 0u, nthr > 32 ? tid - 32 : tid, nthr > 32 ? nthr - 32 : nthr);
        }
        PROF(7)  // early STDP
        if (t + 1 < T) publish_trace(t + 1);  // input trace of step t+1 (its winners read it after barrier t+1)
        if (tid == tid_pf && t + 2 <= T) M.denseflag[buf] = dflag;
        PROF(8)  // trace publish
        // ---- wait(t) ----------------------------------------------------------------------------
        if (tid == 0 && !(Q.dbg & 1)) {
            // three polling loads kept in flight a third of an L2 round trip apart: the completion is
            // seen ~1/6 of a round trip after it becomes visible instead of ~1/2
            const unsigned int target = G * gen;
            const long long t0 = clock64();
            unsigned int r0 = ld_relaxed_u32(Q.bar), r1 = r0, r2 = r0;
            if ((int)(r0 - target) < 0) {
                r0 = ld_relaxed_u32(Q.bar);
                { const long long s0 = clock64(); while (clock64() - s0 < 220) {} }
                r1 = ld_relaxed_u32(Q.bar);
                { const long long s0 = clock64(); while (clock64() - s0 < 220) {} }
                r2 = ld_relaxed_u32(Q.bar);
                for (;;) {
                    if ((int)(r0 - target) >= 0) break;
                    r0 = ld_relaxed_u32(Q.bar);
                    if ((int)(r1 - target) >= 0) break;
                    r1 = ld_relaxed_u32(Q.bar);
                    if ((int)(r2 - target) >= 0) break;
                    r2 = ld_relaxed_u32(Q.bar);
                    if (clock64() - t0 > 4000000000LL) { if (Q.err) atomicOr(Q.err, SNN_ERR_BARRIER); M.abort = 1; break; }
                }
            }
#ifdef SNN_EMU
            __atomic_thread_fence(__ATOMIC_SEQ_CST);
#else
            asm volatile("fence.acquire.gpu;" ::: "memory");
#endif
        }
        __syncthreads();
        if (M.abort) return;
        PROF(9)  // barrier wait
        if (PROFV && Q.prof && tid == 0 && t >= 100 && t < 132) {  // per-step trace: cumulative phase cycles
            for (int k = 0; k < NPROF; ++k) Q.prof[320 * NPROF + ((t - 100) * 160 + blockIdx.x) * NPROF + k] = pc[k];
        }
    }

    // ---- epilogue: normalize() on the tile (network.py:464-465), write everything back -----
    __syncthreads();
    if (Q.normalize && C.has_norm) {
        float *part = xrow;  // [SNN_NORM_CHUNKS + 1][TJ]
        const int chunk = (P + SNN_NORM_CHUNKS - 1) / SNN_NORM_CHUNKS;
        #pragma unroll 1
        for (int idx = tid; idx < SNN_NORM_CHUNKS * TJ; idx += nthr) {
            const int c = idx / TJ, jj = idx % TJ;
            float a = 0.0f;
            const int i1 = min((c + 1) * chunk, P);
            #pragma unroll 1
            for (int i = c * chunk; i < i1; ++i) { const float x = W[i * WS + jj]; a = a + (C.norm_abs ? fabsf(x) : x); }
            part[idx] = a;
        }
        __syncthreads();
        if (tid < TJ) {
            float tot = 0.0f;
            for (int c = 0; c < SNN_NORM_CHUNKS; ++c) tot = tot + part[c * TJ + tid];
            if (tot == 0.0f) tot = 1.0f;
            part[SNN_NORM_CHUNKS * TJ + tid] = C.norm / tot;
        }
        __syncthreads();
        #pragma unroll 1
        for (int idx = tid; idx < P * TJ; idx += nthr) { const int i = idx / TJ, jj = idx % TJ; W[i * WS + jj] = W[i * WS + jj] * part[SNN_NORM_CHUNKS * TJ + jj]; }
        __syncthreads();
    }
    #pragma unroll 1
    for (int idx = tid; idx < P * TJ; idx += nthr) {
        const int i = idx / TJ, jj = idx % TJ;
        if (j0 + jj < n) {
            const size_t k = (size_t)i * n + j0 + jj;
            if (Q.delta_w) Q.delta_w[k] = W[i * WS + jj] - C.w[k];   // multi-GPU window: the caller's all-reduce buffer
            else C.w[k] = W[i * WS + jj];
        }
    }
    for (int jj = tid; jj < TJ; jj += nthr)
        if (j0 + jj < n) {
            if (Q.delta_theta) Q.delta_theta[j0 + jj] = theta_s[jj] - E.theta[j0 + jj];
            else E.theta[j0 + jj] = theta_s[jj];
        }
    if (act) {
        #pragma unroll
        for (int c = 0; c < 4; ++c)
            if (jc + c < n) {
                const size_t k = (size_t)b * n + jc + c;
                E.v[k] = vE[c]; E.refrac_count[k] = rE[c];
                if (E.traces) E.x[k] = xE[c];
                E.s[k] = (sEprev >> c) & 1u;
                I.v[k] = vI[c]; I.refrac_count[k] = rI[c];
                I.s[k] = (sIprev >> c) & 1u;
                if (E.rec_count) E.rec_count[k] += (int)(((c < 2 ? cE01 : cE23) >> ((c & 1) * 16)) & 0xffffu);
                if (I.rec_count) I.rec_count[k] += (int)(((c < 2 ? cI01 : cI23) >> ((c & 1) * 16)) & 0xffffu);
            }
    }
    PROF(11)  // epilogue (partial)
    if (PROFV && Q.prof && tid == 0)
    {
        for (int k = 0; k < NPROF; ++k) { Q.prof[blockIdx.x * NPROF + k] = pc[k]; Q.prof[(160 + blockIdx.x) * NPROF + k] = 0; }
        for (int k = 0; k < 8; ++k) atomicAdd((unsigned long long *)(Q.prof + 320 * NPROF + NPROF * 32 * 160 + 600 + k), (unsigned long long)M.lc[k]);
    }
    for (int o = 0; o < own; ++o) {
        const int bo = blockIdx.x + o * (int)G;
        if (bo < B) {
            const uint32_t *srow = Q.inS + ((size_t)T * B + bo) * Q.SW;  // slot T = spikes of step T-1
            #pragma unroll 1
            for (int i = tid; i < P; i += nthr) {
                if (X.traces) X.x[(size_t)bo * P + i] = xown[o * P + i];
                X.s[(size_t)bo * P + i] = (__ldg(srow + (i >> 5)) >> (i & 31)) & 1u;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Pre-pass: one CTA per slot.  Slot 0 = the Input layer's incoming spike state, slot t+1 = the
// external input of step t (network.py:388-392 / Input.forward nodes.py:211-221).  Produces the
// per-sample bit rows and pixel lists, the per-pixel sample masks, the Input monitor raster,
// flags non-binary input; CTA 0 also resets the exchange slots and counts the incoming Ai spikes.
__global__ void __launch_bounds__(256) snn_dc_prepass(const __grid_constant__ FusedParams Q, int BW) {
    // grid = (slot, group of 32 samples): each CTA converts 32 samples of one timestep
#ifdef SNN_EMU
    uint32_t *sbits = (uint32_t *)emu::tls_cta->dyn_smem;
#else
    extern __shared__ uint32_t sbits[];  // [32][SW]
#endif
    const int B = Q.B, P = Q.P, SW = Q.SW, PW = (P + 31) / 32;
    const int slot = blockIdx.x, grp = blockIdx.y, b0 = grp * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const snn_layer_t &X = Q.X;
    unsigned char *blk = Q.evS + (size_t)slot * Q.SB;
    uint16_t *ecnt = (uint16_t *)blk;
    uint16_t *elist = (uint16_t *)(blk + ev_count_bytes(B));
    bool nonbin = false, dense = false;
    // fast path: byte spikes (uint8 / bool input, or the layer's own spike state for slot 0) in
    // 16-byte aligned rows — one 16-byte load per lane covers 16 pixels
    const bool bytes_in = slot == 0 || (X.ext && X.ext_dtype == SNN_EXT_U8);
    const unsigned char *src0 = slot == 0 ? (const unsigned char *)X.s : (const unsigned char *)X.ext + (size_t)(slot - 1) * B * P;
    const bool fast16 = bytes_in && (P & 15) == 0 && (((size_t)src0) & 15) == 0 && (((size_t)X.rec_s) & 15) == 0;
    for (int bl = warp; bl < 32; bl += nwarp) {
        const int b = b0 + bl;
        int total = 0;
        if (b < B && fast16) {
            uint16_t *lst = elist + b * EV_CAP;
            const uint4 *row = (const uint4 *)(src0 + (size_t)b * P);
            unsigned char *rec = (slot > 0 && X.rec_s) ? X.rec_s + ((size_t)(slot - 1) * B + b) * P : nullptr;
            const int nchunk = P >> 4;
            for (int c0 = 0; c0 < nchunk; c0 += 32) {
                const int c = c0 + lane;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (c < nchunk) v = __ldg(row + c);
                if (slot > 0) nonbin |= ((v.x | v.y | v.z | v.w) & 0xfefefefeu) != 0u;
                const uint32_t zx = __vcmpne4(v.x, 0u) & 0x01010101u, zy = __vcmpne4(v.y, 0u) & 0x01010101u,
                               zz = __vcmpne4(v.z, 0u) & 0x01010101u, zw = __vcmpne4(v.w, 0u) & 0x01010101u;
                if (rec && c < nchunk) ((uint4 *)rec)[c] = make_uint4(zx, zy, zz, zw);
                // 4 flag bytes -> 4 bits: the multiply moves byte k's bit 0 to bit 24 + k
                const uint32_t m16 = ((zx * 0x01020408u) >> 24) | (((zy * 0x01020408u) >> 24) << 4) | (((zz * 0x01020408u) >> 24) << 8) |
                                     (((zw * 0x01020408u) >> 24) << 12);
                const uint32_t other = __shfl_xor_sync(0xffffffffu, m16, 1);
                if (!(lane & 1)) {  // even lane: the 32-pixel word of chunks c, c + 1
                    const int w = c >> 1;
                    if (w < SW) {
                        const uint32_t word = m16 | (other << 16);
                        sbits[bl * SW + w] = word;
                        Q.inS[((size_t)slot * B + b) * SW + w] = word;
                    }
                }
                // ascending pixel list: position = spikes before me
                const int mine = __popc(m16);
                int incl = mine;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                int pos = total + incl - mine;
                uint32_t mm = m16;
                while (mm) {
                    const int bit = __ffs(mm) - 1;
                    mm &= mm - 1;
                    if (pos < EV_CAP) lst[pos] = (uint16_t)(c * 16 + bit);
                    ++pos;
                }
                total += __shfl_sync(0xffffffffu, incl, 31);
            }
            for (int w = (P >> 5) + lane; w < SW; w += 32) {  // words past the last pixel (odd chunk count: half word above)
                if (w * 32 >= P) { sbits[bl * SW + w] = 0u; Q.inS[((size_t)slot * B + b) * SW + w] = 0u; }
            }
            const int padded = (total + 3) & ~3;  // pad to a multiple of 4 with the zero row P
            if (total < EV_CAP && lane < padded - total) lst[total + lane] = (uint16_t)P;
            if (lane == 0) ecnt[b] = (uint16_t)(total > 65535 ? 65535 : total);
            dense |= total > EV_CAP;
        } else if (b < B) {
            uint16_t *lst = elist + b * EV_CAP;
            for (int w = 0; w < SW; ++w) {
                const int i = w * 32 + lane;
                bool s = false;
                if (w < PW && i < P) {
                    if (slot == 0) s = X.s[(size_t)b * P + i] != 0;
                    else if (X.ext) {
                        const size_t idx = ((size_t)(slot - 1) * B + b) * P + i;
                        if (X.ext_dtype == SNN_EXT_U8) { const uint8_t e = ((const uint8_t *)X.ext)[idx]; s = e != 0; nonbin |= e > 1; }
                        else { const float e = ((const float *)X.ext)[idx]; s = e != 0.0f; nonbin |= (e != 0.0f && e != 1.0f); }
                    }
                    if (slot > 0 && X.rec_s) X.rec_s[((size_t)(slot - 1) * B + b) * P + i] = s ? 1 : 0;
                }
                const uint32_t word = __ballot_sync(0xffffffffu, s);
                if (lane == 0) { sbits[bl * SW + w] = word; Q.inS[((size_t)slot * B + b) * SW + w] = word; }
                if (s) {  // ascending pixel list: position = spikes before me
                    const int pos = total + __popc(word & ((1u << lane) - 1u));
                    if (pos < EV_CAP) lst[pos] = (uint16_t)i;
                }
                total += __popc(word);
            }
            const int padded = (total + 3) & ~3;  // pad to a multiple of 4 with the zero row P
            if (total < EV_CAP && lane < padded - total) lst[total + lane] = (uint16_t)P;
            if (lane == 0) ecnt[b] = (uint16_t)(total > 65535 ? 65535 : total);
            dense |= total > EV_CAP;
        } else {
            for (int w = lane; w < SW; w += 32) sbits[bl * SW + w] = 0u;
        }
    }
    if (dense && lane == 0) atomicOr(Q.dense + slot, 1);
    if (grp == 0 && threadIdx.x < 8) ecnt[B + threadIdx.x] = 0;
    __syncthreads();
    // transpose 32x32 bit blocks: inT[pixel][grp] bit b' = inS[b0+b'][pixel/32] bit pixel%32
    // (5 butterfly steps: lanes l and l ^ j swap the bit blocks whose index differs in bit j)
    for (int w = warp; w < PW; w += nwarp) {
        uint32_t x = sbits[lane * SW + w];
        #pragma unroll
        for (int j = 16; j >= 1; j >>= 1) {
            const uint32_t m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
            const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
            x = (lane & j) ? (((y & ~m) >> j) | (x & ~m)) : ((x & m) | ((y & m) << j));
        }
        const int i = w * 32 + lane;
        if (i < P) {
            Q.inT[((size_t)slot * P + i) * BW + grp] = x;
            if (grp == 0)  // zero the padding groups of the per-pixel masks
                for (int g = (B + 31) / 32; g < BW; ++g) Q.inT[((size_t)slot * P + i) * BW + g] = 0u;
        }
    }
    if (nonbin && Q.err) atomicOr(Q.err, SNN_ERR_NONBINARY);
    if (slot == 0 && grp == 0) {
        for (int k = threadIdx.x; k < 3 * B; k += blockDim.x) Q.win[k] = 0ull;
        for (int k = threadIdx.x; k < 2 * B; k += blockDim.x) Q.sisum[k] = 0u;
        for (int b = warp; b < B; b += nwarp) {  // Ai spikes of step -1 (slot 2 = (-1) mod 3)
            int c = 0;
            for (int j = lane; j < Q.n; j += 32) c += Q.I.s[(size_t)b * Q.n + j] != 0;
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0) Q.sisum[2 * B + b] = (unsigned int)c;
        }
    }
}

struct Match {
    int lX, lE, lI, cXE, cEI, cIE, TJ, BW, threads, grid, SW, SB, own;
    size_t smem;
};

int device_sms() {
    static int sms = -1;
#ifdef SNN_EMU
    if (sms < 0) { const char *v = getenv("SNN_EMU_FUSED_SMS"); sms = v && atoi(v) > 0 ? atoi(v) : 148; }
#else
    if (sms < 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
        else { sms = 148; (void)cudaGetLastError(); }
    }
#endif
    return sms;
}

bool match(const snn_net_t *net, const snn_run_opts_t *o, Match &m) {
    if (net->n_layers != 3 || net->n_conns != 3 || o->T < 1 || o->one_step) return false;
    m.lX = m.lE = m.lI = -1;
    for (int l = 0; l < 3; ++l) {
        const snn_layer_t &L = net->layers[l];
        if (L.clamp || L.unclamp || L.inject_v || L.sum_input) return false;
        if (L.kind == SNN_NODE_INPUT) m.lX = l;
        else if (L.kind == SNN_NODE_DC) m.lE = l;
        else if (L.kind == SNN_NODE_LIF) m.lI = l;
    }
    if (m.lX < 0 || m.lE < 0 || m.lI < 0) return false;
    const snn_layer_t &X = net->layers[m.lX], &E = net->layers[m.lE], &I = net->layers[m.lI];
    if (E.ext || I.ext || I.traces || E.n != I.n) return false;
    if (X.rec_count || ((E.rec_count || I.rec_count) && o->T > 65535)) return false;
    m.cXE = m.cEI = m.cIE = -1;
    for (int c = 0; c < 3; ++c) {
        const snn_conn_t &C = net->conns[c];
        if (C.b || C.mask || C.kind == SNN_CONN_CONV2D || C.rule > SNN_RULE_MCC_POSTPRE) return false;
        if (C.src == m.lX && C.tgt == m.lE) m.cXE = c;
        else if (C.src == m.lE && C.tgt == m.lI) m.cEI = c;
        else if (C.src == m.lI && C.tgt == m.lE) m.cIE = c;
    }
    if (m.cXE < 0 || m.cEI < 0 || m.cIE < 0 || m.cXE > m.cIE) return false;  // accumulation order into Ae
    const snn_conn_t &CX = net->conns[m.cXE], &CEI = net->conns[m.cEI], &CIE = net->conns[m.cIE];
    auto is_static = [](const snn_conn_t &C) {
        return (C.rule == SNN_RULE_NONE || (C.rule == SNN_RULE_NOOP && (C.weight_decay == 1.0f || C.weight_decay == 0.0f))) && !C.has_norm;
    };
    if (!is_static(CEI) || !is_static(CIE)) return false;
    if (CEI.structure != SNN_W_DIAG || CIE.structure != SNN_W_OFFDIAG) return false;
    if (CX.rule >= SNN_RULE_POSTPRE && (!X.traces || !E.traces)) return false;
    const int n = E.n, P = X.n, B = o->B;
    if (B > 256 || (P & 3) || P >= 65535) return false;
    const int sms = device_sms();
    m.SW = ((P + 31) / 32 + 3) / 4 * 4;
    m.BW = B <= 128 ? 4 : 8;
    m.SB = ev_block_bytes(B);
    for (int TJ : {4, 8, 16, 32}) {
        const int grid = (n + TJ - 1) / TJ;
        int threads = ((B * (TJ / 4)) + 31) / 32 * 32;
        if (grid > sms || threads > 1024) continue;
        const int own = (B + grid - 1) / grid;
        const SmemLayout SL = smem_layout(P, TJ, B, m.BW, n, own);
        if (SL.total > 227 * 1024) continue;
        m.TJ = TJ; m.grid = grid; m.threads = threads; m.own = own; m.smem = SL.total;
        return true;
    }
    return false;
}

template <int TJ, int BW, int VAR>
cudaError_t launch_var(const FusedParams &Q, const Match &m, cudaStream_t stream) {
#ifdef SNN_EMU
    (void)stream;
    emu::run_grid(m.grid, m.threads, m.smem, [](void *a) { snn_dc_fused_window<TJ, BW, VAR>(*(const FusedParams *)a); }, (void *)&Q);
    return cudaSuccess;
#else
    cudaError_t e = cudaFuncSetAttribute(snn_dc_fused_window<TJ, BW, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m.smem);
    if (e != cudaSuccess) return e;
    void *args[] = {(void *)&Q};
    return cudaLaunchCooperativeKernel((void *)snn_dc_fused_window<TJ, BW, VAR>, dim3(m.grid), dim3(m.threads), args, m.smem, stream);
#endif
}

template <int TJ, int BW>
cudaError_t launch_tj(const FusedParams &Q, const Match &m, cudaStream_t stream) {
    // lean variant when none of the rare options is in use
    const bool lean = !Q.X.traces_additive && !Q.E.traces_additive && !Q.E.has_lbound && !Q.I.has_lbound && !Q.E.rec_v && !Q.I.rec_v &&
                      Q.C.reduction == SNN_REDUCE_SUM && Q.C.rule != SNN_RULE_WDEP_POSTPRE;
    if (Q.prof) return lean ? launch_var<TJ, BW, 3>(Q, m, stream) : launch_var<TJ, BW, 2>(Q, m, stream);
    return lean ? launch_var<TJ, BW, 1>(Q, m, stream) : launch_var<TJ, BW, 0>(Q, m, stream);
}

struct WsLayout { size_t bar, dense, inS, inT, evS, win, sisum, xpub, prof, total; };
WsLayout ws_layout(const Match &m, int T, int B, int P) {
    WsLayout L; size_t o = 0;
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    L.bar = o; o += al(sizeof(unsigned int) * 96);
    L.dense = o; o += al(sizeof(int) * (size_t)(T + 1));
    L.inS = o; o += al(sizeof(uint32_t) * (size_t)(T + 1) * B * m.SW);
    L.inT = o; o += al(sizeof(uint32_t) * (size_t)(T + 1) * P * m.BW);
    L.evS = o; o += al((size_t)(T + 1) * m.SB);
    L.win = o; o += al(sizeof(unsigned long long) * 3 * B);
    L.sisum = o; o += al(sizeof(unsigned int) * 3 * B);
    L.xpub = o; o += al(sizeof(float) * 3 * (size_t)B * P);
    L.prof = o; o += al(sizeof(long long) * (320 * NPROF + NPROF * 32 * 160 + 640));
    L.total = o;
    return L;
}

}  // namespace

int snn_fused_dc_supported(const snn_net_t *net, const snn_run_opts_t *opts) {
    Match m;
    return match(net, opts, m) ? 1 : 0;
}

size_t snn_fused_dc_workspace_bytes(const snn_net_t *net, const snn_run_opts_t *opts) {
    Match m;
    if (!match(net, opts, m)) return 0;
    return ws_layout(m, opts->T, opts->B, net->layers[m.lX].n).total;
}

int snn_fused_dc_launch(const snn_net_t *net, const snn_run_opts_t *opts, void *ws_, size_t ws_bytes, cudaStream_t stream,
                        int *launches) {
    Match m;
    if (!match(net, opts, m)) return SNN_ERR_UNSUPPORTED;
    const int T = opts->T, B = opts->B, P = net->layers[m.lX].n;
    const WsLayout WL = ws_layout(m, T, B, P);
    if (ws_bytes < WL.total) return SNN_ERR_WORKSPACE;
    char *ws = (char *)ws_;
    FusedParams Q;
    memset(&Q, 0, sizeof(Q));
    Q.X = net->layers[m.lX]; Q.E = net->layers[m.lE]; Q.I = net->layers[m.lI];
    Q.C = net->conns[m.cXE];
    Q.exc = net->conns[m.cEI].structure_val; Q.inh_neg = net->conns[m.cIE].structure_val;
    Q.T = T; Q.B = B; Q.P = P; Q.n = Q.E.n; Q.learning = net->learning; Q.normalize = opts->normalize;
    { const char *d = getenv("SNN_B200_DEBUG"); Q.dbg = d ? atoi(d) : 0; }
    {
        const SmemLayout SL = smem_layout(P, m.TJ, B, m.BW, Q.n, m.own);
        Q.o_W = (uint32_t)SL.W; Q.o_tx = (uint32_t)SL.tx; Q.o_ev = (uint32_t)SL.ev; Q.o_inT = (uint32_t)SL.inT; Q.o_xrow = (uint32_t)SL.xrow;
        Q.o_rep = (uint32_t)SL.rep; Q.o_xown = (uint32_t)SL.xown; Q.o_theta = (uint32_t)SL.theta; Q.o_live = (uint32_t)SL.live;
        Q.o_misc = (uint32_t)SL.misc;
    }
    Q.SW = m.SW; Q.SB = m.SB; Q.liE = m.lE; Q.seed = opts->seed; Q.step_offset = opts->step_offset;
    Q.inS = (uint32_t *)(ws + WL.inS); Q.inT = (uint32_t *)(ws + WL.inT); Q.evS = (unsigned char *)(ws + WL.evS);
    Q.win = (unsigned long long *)(ws + WL.win); Q.sisum = (unsigned int *)(ws + WL.sisum);
    Q.xpub = (float *)(ws + WL.xpub); Q.bar = (unsigned int *)(ws + WL.bar); Q.err = opts->err_flag;
    Q.delta_w = opts->delta_w; Q.delta_theta = opts->delta_theta;
    const bool prof = getenv("SNN_B200_PROF") != nullptr;
    Q.prof = prof ? (long long *)(ws + WL.prof) : nullptr;
    Q.dense = (int *)(ws + WL.dense);
    // barrier words and per-slot dense flags are adjacent: one memset node
    if (cudaMemsetAsync(Q.bar, 0, WL.inS - WL.bar, stream) != cudaSuccess) return SNN_ERR_CUDA;
    if (prof && cudaMemsetAsync(Q.prof + 320 * NPROF + NPROF * 32 * 160 + 600, 0, 8 * sizeof(long long), stream) != cudaSuccess) return SNN_ERR_CUDA;
    // the two static matrices are replaced by their constants: make sure they still have that structure
    if (snn_verify_structure(net->conns[m.cEI], Q.n, Q.err, stream) != SNN_OK || snn_verify_structure(net->conns[m.cIE], Q.n, Q.err, stream) != SNN_OK) return SNN_ERR_CUDA;
#ifdef SNN_EMU
    {
        struct PreArgs { const FusedParams *Q; int BW; } pa = {&Q, m.BW};
        emu::run_grid_independent(T + 1, (B + 31) / 32, 256, sizeof(uint32_t) * 32 * (size_t)m.SW,
                                  [](void *a) { const PreArgs *p = (const PreArgs *)a; snn_dc_prepass(*p->Q, p->BW); }, &pa);
    }
#else
    snn_dc_prepass<<<dim3(T + 1, (B + 31) / 32), 256, sizeof(uint32_t) * 32 * (size_t)m.SW, stream>>>(Q, m.BW);
#endif
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) {
        if (m.BW == 4) {
            switch (m.TJ) {
                case 4: e = launch_tj<4, 4>(Q, m, stream); break;
                case 8: e = launch_tj<8, 4>(Q, m, stream); break;
                case 16: e = launch_tj<16, 4>(Q, m, stream); break;
                default: e = launch_tj<32, 4>(Q, m, stream); break;
            }
        } else {
            switch (m.TJ) {
                case 4: e = launch_tj<4, 8>(Q, m, stream); break;
                case 8: e = launch_tj<8, 8>(Q, m, stream); break;
                default: e = launch_tj<16, 8>(Q, m, stream); break;
            }
        }
    }
    if (e != cudaSuccess) {
        fprintf(stderr, "libsnn_b200: fused DC2015 window launch failed: %s\n", cudaGetErrorString(e));
        return SNN_ERR_CUDA;
    }
    if (prof) {  // debug only: synchronise and print the per-phase cycle counts (mean / max over CTAs)
        static const char *names[NPROF] = {"prologue", "exchange loads", "late sync+reset", "mbar wait+sync", "gather+neurons", "reduce+atomics",
                                           "theta+prefetch+arr", "early STDP", "trace publish", "barrier wait", "(unused)", "epilogue",
                                           "late finalise", "late sync", "late group pass", "arrive (release)"};
        cudaStreamSynchronize(stream);
        static long long hostp[320 * NPROF];
        cudaMemcpy(hostp, Q.prof, sizeof(long long) * 320 * NPROF, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[snn_b200 prof] grid=%d T=%d (cycles per timestep, thread 0 of each CTA: min / mean / max)\n", m.grid, T);
        for (int k = 0; k < NPROF; ++k) {
            double sum = 0, mx = 0, mn = 1e300;
            for (int g = 0; g < m.grid; ++g) { const double v = (double)hostp[g * NPROF + k]; sum += v; mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
            const double div = (k == 0 || k == 11) ? 1.0 : (double)T;
            double smx = 0, smean = 0;
            for (int g = 0; g < m.grid; ++g) { const double v = (double)hostp[(160 + g) * NPROF + k]; smx = v > smx ? v : smx; smean += v / m.grid; }
            fprintf(stderr, "  %-18s %10.0f %10.0f %10.0f   worst single step: mean over CTAs %8.0f, max %8.0f\n", names[k], mn / div,
                    sum / m.grid / div, mx / div, smean, smx);
        }
        {   // per-step trace (steps 101..131): which phase makes the slowest CTA of a step slow?
            static long long tr[NPROF * 32 * 160];
            cudaMemcpy(tr, Q.prof + 320 * NPROF, sizeof(tr), cudaMemcpyDeviceToHost);
            double slow[NPROF] = {0}, mean[NPROF] = {0}, s_maxw = 0, s_meanw = 0; int cnt = 0;
            for (int st = 1; st < 32; ++st) {
                int gmax = 0; double maxw = -1, meanw = 0;
                for (int g = 0; g < m.grid; ++g) {
                    double w = 0;
                    for (int k = 1; k <= 15; ++k) if (k <= 8 || k == 10 || k >= 12) w += (double)(tr[(st * 160 + g) * NPROF + k] - tr[((st - 1) * 160 + g) * NPROF + k]);
                    meanw += w / m.grid;
                    if (w > maxw) { maxw = w; gmax = g; }
                }
                for (int k = 1; k <= 15; ++k) {
                    slow[k] += (double)(tr[(st * 160 + gmax) * NPROF + k] - tr[((st - 1) * 160 + gmax) * NPROF + k]);
                    for (int g = 0; g < m.grid; ++g) mean[k] += (double)(tr[(st * 160 + g) * NPROF + k] - tr[((st - 1) * 160 + g) * NPROF + k]) / m.grid;
                }
                s_maxw += maxw; s_meanw += meanw; ++cnt;
            }
            {
                long long lcs[8];
                double tot14 = 0;
                for (int g = 0; g < m.grid; ++g) tot14 += (double)hostp[g * NPROF + 14];
                cudaMemcpy(lcs, Q.prof + 320 * NPROF + NPROF * 32 * 160 + 600, sizeof(lcs), cudaMemcpyDeviceToHost);
                fprintf(stderr, "  late passes %lld (%.1f%% of CTA-steps): fast %lld, set-up cycles %lld, set-up + first row cycles %lld, with winner %lld, pre-part cycles %lld, "
                        "candidate samples per pass %.2f; cycles per fast pass %.0f, per other pass %.0f\n", lcs[0], 100.0 * lcs[0] / ((double)m.grid * T), lcs[1], lcs[2], lcs[3], lcs[4], lcs[5],
                        (double)lcs[6] / (lcs[0] ? lcs[0] : 1), (double)lcs[7] / (lcs[1] ? lcs[1] : 1),
                        (tot14 - (double)lcs[7]) / (lcs[0] - lcs[1] ? lcs[0] - lcs[1] : 1));
            }
            fprintf(stderr, "  per-step (t=101..131): work mean %.0f, work of the slowest CTA of each step %.0f; by phase (mean CTA / slowest CTA):\n", s_meanw / cnt, s_maxw / cnt);
            for (int k = 1; k <= 15; ++k) if (k != 11) fprintf(stderr, "      %-18s %8.0f %8.0f\n", names[k], mean[k] / cnt, slow[k] / cnt);
        }
        {
            double sum = 0, mx = 0, mn = 1e300; int amx = 0, amn = 0;
            for (int g = 0; g < m.grid; ++g) {
                double v = 0;
                for (int k = 1; k <= 15; ++k) if (k <= 8 || k == 10 || k >= 12) v += (double)hostp[g * NPROF + k];
                sum += v; if (v > mx) { mx = v; amx = g; } if (v < mn) { mn = v; amn = g; }
            }
            fprintf(stderr, "  %-18s %10.0f %10.0f %10.0f   (slowest CTA %d, fastest %d)\n", "work w/o barrier", mn / T, sum / m.grid / T, mx / T, amx, amn);
        }
    }
    if (launches) *launches = 4;  // 2 structure checks + pre-pass + persistent window kernel
    return SNN_OK;
}
