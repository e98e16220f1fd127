// snn_fused_dc2.cu — fused persistent window kernel for the DiehlAndCook2015 graph, second generation
//     X (Input) --W, learned--> Ae (DiehlAndCookNodes) --exc*I--> Ai (LIFNodes) --(-inh)(1-I)--> Ae
// (reference wiring: bindsnet/models/models.py:94-244; per-step semantics: SURVEY.md App. A).
//
// Same partition as snn_fused_dc.cu — a CTA owns TJ columns (neurons of Ae and their partners in Ai)
// for all B samples, W[:, tile] lives in shared memory for the whole window, Ae state in registers,
// one split-phase grid barrier per timestep — but the step is reorganised around what the round-1
// profile showed: the step is bound by the latency of the grid exchange (about four L2 round trips:
// scripts/exchange_bench.cu measures 1.45 us for barrier + data on 100-148 CTAs; message passing
// between all pairs of CTAs is slower, 1.8-2.9 us) and by shared-memory bank conflicts of the
// spike-gather, not by arithmetic.  Therefore:
//
//   * a dedicated EXCHANGE WARP per CTA arrives at the barrier, polls it, and copies the exchanged
//     per-sample words (one_spike arg-max key, nodes.py:1097-1105; number of Ai spikes = lateral
//     inhibition, models.py:217-220) into shared-memory tables, while the compute warps spend the same
//     time on everything of step t+1 that does not depend on the exchange: the early STDP AND the
//     spike-gather of step t+1 for the column groups whose weights are already final;
//   * the input traces the STDP post term needs (x_pre of the winner's sample, learning.py:407-417 /
//     MCC_learning.py:267-299) are a pure function of the input spikes, so a pre-pass scans them for
//     the whole window ([T,B,P] fp32); the window kernel bulk-copies (cp.async.bulk + mbarrier) the
//     rows of its own candidate samples the moment a candidate appears — before the barrier, not after;
//   * Ai (LIFNodes, nodes.py:500-529) is event driven: a neuron at rest without input stays bitwise
//     at rest (decay * (rest - rest) + rest == rest), only its refractory counter runs, and that is
//     replayed in closed form at the end.  Neurons that ever received a spike live in a compact list;
//   * spike rasters and spike counts are written sparsely (the launch code clears the rasters);
//   * thread layout is column-group major (warp = 32 samples of ONE float4 column group), so early /
//     late column groups are warp-uniform and TJ need not be a power of two: n = 1600 runs as
//     134 CTAs x 12 columns instead of 100 x 16.
//
// Loop iteration t = [exchange of step t-1 lands] [winners, traces, late STDP of t-1] [step t: gather of
// the late groups, Ae update, Ai list, candidates -> atomics] [arrive] [early STDP of t, gather of t+1].
//
// Arithmetic and summation orders are those of snn_phases.cuh / oracle/snn_oracle.c (one fp32
// rounding per reference op, ascending index sums), so results are bit-identical to the generic
// kernel, to snn_fused_dc.cu and to the oracle.  Options outside the lean set (weight-dependent
// rule, mean reduction, additive traces, voltage bounds / monitors, weight decay, B > 128) stay
// with snn_fused_dc.cu.
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "snn_common.cuh"

int snn_verify_structure(const snn_conn_t &C, int n, int32_t *err, cudaStream_t stream);

namespace {

constexpr int XR = 8;        // input-trace rows staged per CTA and step (samples with a candidate)
constexpr int EV_CAP = 32;   // staged events per sample and step; longer lists take the slow path
constexpr int NPROF = 16;
// build-time experiment switches (scripts/build_variant.sh)
#ifndef V2_UNIFY_AI
#define V2_UNIFY_AI 0     // the Ai list step calls the out-of-line copy the winners' path uses
#endif
#ifndef V2_UNIFY_LIST
#define V2_UNIFY_LIST 0   // the early STDP list pass calls the out-of-line copy the late pass uses
#endif
#ifndef V2_POLL_NS
#define V2_POLL_NS 0      // back-off between two polls of the grid barrier word
#endif
constexpr unsigned AI_NONE = 0xFFFFu;

struct F2Params {
    snn_layer_t X, E, I;      // Input, DiehlAndCookNodes (Ae), LIFNodes (Ai)
    snn_conn_t C;             // X -> Ae
    float exc, inh_neg;       // diag value of Ae->Ai, off-diag value of Ai->Ae
    int32_t T, B, Bp, P, n, learning, normalize;
    int32_t SW;               // words per sample row of inS
    int32_t SB;               // bytes of one event-list block
    int32_t liE;              // index of Ae in the user's layer list (enters the tie-break hash)
    int32_t G;                // CTAs of the window kernel
    int32_t nrep;             // entries of the inhibition table rep[0..nrep]
    uint32_t o_W, o_tx, o_ev, o_inT, o_xrow, o_rep, o_theta, o_live, o_tab, o_ai, o_misc;  // smem byte offsets
    uint32_t seed, step_offset;
    uint32_t *inS;            // [T+1][B][SW]  slot t = spikes of step t-1: bit i of sample b
    uint32_t *inT;            // [T+1][P][BW]  same spikes: bit b of pixel i
    unsigned char *evS;       // [T+1][SB]     same spikes as lists: u16 count[B] (padded), then u16 idx[B][EV_CAP]
    int *dense;               // [T+1] slot holds a sample whose event list overflowed EV_CAP
    uint8_t *xage;            // [T][B][P] input traces of every step as AGES (steps since the pixel's last spike in this
                              // window, 255 = none yet): the trace is dtab[age] (pre-pass scan), or NULL
    float *x0c;               // [B][P] the Input layer's trace at the start of the window (for pixels without a spike yet)
    int *anyx0;               // set by the scan when any of those is non-zero
    float *rep;               // [nrep+1] m-fold sequential sums of the Ai->Ae weight
    unsigned int *sisum0;     // [B] Ai spikes of step -1
    unsigned long long *win;  // [3][B] one_spike arg-max keys, slot t % 3
    unsigned int *sisum;      // [3][B] Ai spike counts, slot t % 3
    unsigned int *bar;        // grid barrier: monotonic arrival counter
    int32_t *err;
    int32_t dbg;              // profiling only (env SNN_B200_DEBUG): 1 no slot prefetch, 2 no early STDP, 4 no gather ahead,
                              // 8 no trace-row staging, 16 no barrier wait — results invalid
    long long *prof;          // profiling only (env SNN_B200_PROF): [G][NPROF] phase cycles of thread 0
};

#ifdef SNN_EMU   // tests/emu: mbarrier / bulk copy / polling loads / named barriers on the CUDA-model emulation (test infrastructure)
__device__ __forceinline__ void mbar_init(uint64_t *bar, int) { emu::Mbar *m = (emu::Mbar *)bar; m->phase = 0; m->pend = emu::MBAR_ARRIVAL; }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) { emu::Mbar *m = (emu::Mbar *)bar; m->pend += (int32_t)bytes - emu::MBAR_ARRIVAL; emu::mbar_settle(m); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { ((emu::Mbar *)bar)->pend += (int32_t)bytes; }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { emu::Mbar *m = (emu::Mbar *)bar; m->pend -= emu::MBAR_ARRIVAL; emu::mbar_settle(m); }
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
__device__ __forceinline__ void bar_group(int id, int count) { emu::named_barrier(id, count); }
#else
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {  // no arrival
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void bar_group(int id, int count) {   // named barrier among `count` threads
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
#endif

struct Misc2 {  // small per-step scratch (shared memory)
    uint64_t mbar_in[2];    // staged spike lists / pixel masks, by buffer
    uint64_t mbar_x[2];     // staged input-trace rows of a step's candidate samples, by step parity
    uint32_t wl[8][XR];     // per column group: winners of the step being finalised: column << 16 | sample << 8 | staged row slot (0xff: none)
    uint32_t nz4[8][8];     // per column group: samples with a non-zero Ae trace in that group
    uint32_t wmask[32][8];  // winners of the step being finalised: per column, bit mask over samples
    int cnt[2][32];         // candidates per column (theta update), by step parity
    int candb[2][XR];       // samples whose input-trace row is staged in xrow, by step parity
    int ncand[2];           // samples with a candidate in this tile (by step parity)
    uint32_t candgrp[2];    // column groups holding a candidate (by step parity)
    uint32_t colwin[8];     // per column group: bit c: column 4g+c has a winner in the step being finalised
    uint32_t candmask[2][8][8];  // by step parity, per column group: samples holding a candidate in that group
    int ncs[2][8];          // ... how many, listed in candlist[parity][group][]
    int nwl[8];             // per column group: number of winners of the step being finalised
    int nlive[8];           // per column group: samples with a non-zero trace, listed in live[g * Bp ...]
    int nact;               // entries of the Ai list
    int nact_snap;          // ... at the last step boundary: what the list pass of the running step covers
    int abort;              // exchange time-out: leave the time loop
    int negzero;            // the weight tile held a -0.0 when it was loaded (post_rows2 must not skip rows)
    int denseflag[2];       // staged slot (by buffer) holds a sample whose event list overflowed EV_CAP
    int defer[8];           // per column group: the early pass left a row of this (candidate-holding) group unclamped; the late path
                            // clamps the group once the post terms are in (the reference clamps once, after both terms)
    long long pc[NPROF];    // phase timers (profiling variant only)
    long long rs[8][16];    // ... fine stamps inside the late path, per column group
};
static_assert(offsetof(Misc2, wl) % 16 == 0 && offsetof(Misc2, nz4) % 16 == 0 && offsetof(Misc2, wmask) % 16 == 0 &&
              offsetof(Misc2, candmask) % 16 == 0, "Misc2: 16-byte rows");

struct SmemLayout2 { size_t W, tx, ev, inT, xrow, rep, theta, live, tab, ai, misc, total; };

__host__ __device__ inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }
// One event-list block (one slot): u16 count[B] (+ 8 spare, padded to 16 bytes), then the ascending pixel
// lists in CHUNK-MAJOR order — chunk c (list positions 4c..4c+3) of all samples back to back, 8 bytes per
// (chunk, sample) — so that the 32 samples of a warp read consecutive words (no bank conflicts).
__host__ __device__ inline int ev_count_bytes(int B) { return (int)al16(2 * (size_t)(B + 8)); }
__host__ __device__ inline int ev_block_bytes(int B) { return ev_count_bytes(B) + 2 * B * EV_CAP; }
__host__ __device__ inline int ev_pos(int B, int b, int k) { return ((k >> 2) * B + b) * 4 + (k & 3); }   // u16 index behind the counts
__host__ __device__ inline int tile_stride(int TJ) { return ((TJ / 4) & 1) ? TJ : TJ + 4; }  // odd number of 16-byte chunks per row
// tab region: keyT u64 [2][Bp] | isumT u32 [2][Bp] | aispk u32 [2][Bp] | candstamp u32 [2][Bp] | candslot i32 [2][Bp] |
//             candlist u8 [2][8][Bp]
__host__ __device__ inline size_t tab_bytes(int Bp) { return (size_t)Bp * (16 + 8 + 8 + 8 + 8 + 16); }
// ai region: v f32 [cap] | rc f32 [cap] | id u16 [cap] | claim u16 [2][cap] | map u16 [cap] | fl u8 [cap]
__host__ __device__ inline size_t ai_bytes(int cap) { return al16((size_t)cap * (4 + 4 + 2 + 4 + 2 + 1)); }
__host__ __device__ inline SmemLayout2 smem_layout2(int P, int TJ, int B, int Bp, int BW, int nrep) {
    SmemLayout2 L;
    size_t o = 0;
    L.W = o; o += al16(sizeof(float) * (size_t)(P + 1) * tile_stride(TJ));  // + one zero row (list padding)
    L.tx = o; o += al16(sizeof(float) * (size_t)Bp * TJ);
    L.ev = o; o += 2 * al16((size_t)ev_block_bytes(B));
    L.inT = o; o += al16(sizeof(uint32_t) * 2 * (size_t)P * BW);
    L.xrow = o; o += al16(2 * (size_t)XR * P + sizeof(float) * 64 * 17);   // staged age rows by step parity (reused as normalize scratch)
    L.rep = o; o += al16(sizeof(float) * (size_t)(nrep + 1));
    L.theta = o; o += al16(sizeof(float) * (64 + 256));   // theta, threshold, trace-by-age table
    L.live = o; o += al16(sizeof(uint16_t) * (size_t)Bp * (TJ / 4));
    L.tab = o; o += al16(tab_bytes(Bp));
    L.ai = o; o += ai_bytes(Bp * TJ);
    L.misc = o; o += al16(sizeof(Misc2));
    L.total = o;
    return L;
}

// Constants and pointers of the STDP passes, kept in shared memory so that the passes live out of line
// (the per-step code has to stay inside the instruction cache).
// the part the STDP passes copy into registers
struct PassK {
    float *W, *tx;
    const uint32_t *inT;
    const unsigned char *evb;
    const uint8_t *live;
    const uint8_t *candlist;
    Misc2 *M;
    int P, B, Bp, evblk, cntb, WS;
    int pre_on, has_clamp;
    float dts, wmin, wmax, nu1;
};
struct PassCtx2 {
    PassK k;
    // spike-gather
    const uint32_t *inS;
    int SW;
    // Ai list (LIFNodes constants and arrays), exchange
    float *ai_v, *ai_rc;
    uint16_t *ai_id, *ai_claim, *ai_map;
    uint8_t *ai_fl;
    uint32_t *aispk;
    unsigned int *sisum;
    unsigned long long *win;
    uint8_t *I_rec_s;
    int32_t *I_rec_count;
    int I_mon;
    float I_decay, I_rest, I_dt, I_thresh, I_refrac, I_reset;
    int n, j0, TJ, aicap;
    // candidates
    uint32_t *candstamp;
    int *candslot;
    uint8_t *xrow_w;
    const uint8_t *xage;
    const float *x0c, *dtab;
    float x_decay;
    int anyx0;
    uint32_t seed, step_offset;
    int liE, one_spike, stage_on, nostage;
};

// The context lives in shared memory and most of its pointers point into shared memory; loaded back from there
// they are generic pointers to the compiler (LD/ST through the generic path, 64-bit address arithmetic).  These
// hints restore the address space: LDS/STS with 32-bit addresses.
template <class T> __device__ __forceinline__ T *sh(T *p) { __builtin_assume(__isShared((const void *)p)); return p; }
__device__ __forceinline__ PassK load_k(const PassCtx2 *cx) {
    PassK k = sh(cx)->k;
    k.W = sh(k.W); k.tx = sh(k.tx); k.inT = sh(k.inT); k.evb = sh(k.evb); k.live = sh(k.live); k.candlist = sh(k.candlist); k.M = sh(k.M);
    return k;
}

// STDP of one step in list form on ONE column group c4 (MCC_learning.py:234-299, learning.py:390-420), run by the
// nthr0 threads that own the group.  ONE body for its two uses (the per-step code has to stay inside the 32 KB
// instruction cache, the rarely executed late pass included — a cold path costs an L2 round trip per 8 instructions):
//   mode 0, early pass in the shadow of the exchange: work items = (live sample of the group, event of that sample).
//     Several samples can spike at the same pixel: the item whose sample is the LOWEST live one at that pixel owns
//     the row (no atomics), sums the traces of all of them in ascending sample order (the oracle's order) and
//     rewrites the group's 4 weights: w - U*dt, clamp.  Rows at which a sample holding a CANDIDATE of this group
//     spiked (`dm`: those samples) are left alone: the candidate's trace is undecided until the exchange lands.
//     In a group that holds a candidate at all, a weight the clamp would change is stored UNCLAMPED and the group is
//     flagged (Misc2::defer): a winner of the group may still add its post term to that row (its input trace can be
//     non-zero from an earlier spike of the pixel), and the reference clamps once, after both terms
//     (learning.py:97-104) — clamp(clamp(w - U) + V) != clamp(w - U + V) when w - U left the range.  The late path
//     clamps a flagged group after its post terms;
//   mode 1, late pass once the winners are known: exactly those rows — work items = (candidate sample, event), the
//     lowest candidate at a pixel owns the row: pre term (traces of all live samples at the pixel), then for a winner
//     column (`gwin`) the post term of its single winner, then clamp.  Rows this pass does not touch get their post
//     term from post_rows2().
// Input trace of sample b at pixel i after step t from its age byte (Nodes.forward, nodes.py:96-103: decay every
// step, set to trace_scale on a spike): dtab[age]; a pixel without a spike in this window still carries the
// trace it entered the window with, decayed t + 1 times (rare: replayed).
__device__ __noinline__ float xval_nospike(const PassCtx2 *cx, int b, int i, int t) {
    cx = sh(cx);
    float x = cx->x0c[(size_t)b * cx->k.P + i];
    #pragma unroll 1
    for (int k = 0; k <= t; ++k) x = x * cx->x_decay;
    return x;
}
__device__ __forceinline__ float xval(const PassCtx2 *cx, uint32_t age, int b, int i, int t) {
    if (age != 255u) return sh(sh(cx)->dtab)[age];
    return sh(cx)->anyx0 ? xval_nospike(cx, b, i, t) : 0.0f;
}

constexpr int EVH = 16;  // list slots enumerated per sample and round in mode 0 (mode 1: EV_CAP, one round)
template <int CG, int BW>
__device__ __forceinline__ void stdp_list_body(const PassCtx2 *cx, int sb, int c4, int mode, uint32_t gwin, const uint32_t *dm, int par_,
                                               const uint8_t *xrow, int tstep, int tid0, int nthr0) {
    const PassK c_ = load_k(cx);
    dm = sh(dm); xrow = sh(xrow);
    const int P = c_.P, WS = c_.WS, B = c_.B;
    const Misc2 &M = *c_.M;
    const uint16_t *ec = (const uint16_t *)(c_.evb + sb * c_.evblk);
    const uint16_t *el = (const uint16_t *)(c_.evb + sb * c_.evblk + c_.cntb);
    const uint4 *cT = (const uint4 *)(c_.inT + sb * P * BW);
    const uint8_t *lst = mode ? c_.candlist + (par_ * 8 + c4) * c_.Bp : c_.live + c4 * c_.Bp;
    const int per = mode ? EV_CAP : EVH;
    const int total = (mode ? M.ncs[par_][c4] : M.nlive[c4]) * per;
    uint32_t z[BW], d[BW];
    {
        const uint4 z0 = *(const uint4 *)&M.nz4[c4][0], d0 = *(const uint4 *)dm;
        z[0] = z0.x; z[1] = z0.y; z[2] = z0.z; z[3] = z0.w; d[0] = d0.x; d[1] = d0.y; d[2] = d0.z; d[3] = d0.w;
        if (BW == 8) {
            const uint4 z1 = *(const uint4 *)&M.nz4[c4][4], d1 = *(const uint4 *)(dm + 4);
            z[BW - 4] = z1.x; z[BW - 3] = z1.y; z[BW - 2] = z1.z; z[BW - 1] = z1.w;
            d[BW - 4] = d1.x; d[BW - 3] = d1.y; d[BW - 2] = d1.z; d[BW - 1] = d1.w;
        }
    }
    uint32_t anyd = 0;   // the group holds a candidate: its post terms are still to come
    #pragma unroll
    for (int g = 0; g < BW; ++g) anyd |= d[g];
    bool deferred = false;
    #pragma unroll 1
    for (int idx = tid0; idx < total; idx += nthr0) {
        const int bb = lst[idx / per];
        const int cnt = min((int)ec[bb], EV_CAP);
        #pragma unroll 1
        for (int k = idx % per; k < cnt; k += per) {
            const int i = el[ev_pos(B, bb, k)];
            uint32_t a[BW], cm[BW], anyc = 0;
            {
                const uint4 q0 = cT[i * (BW / 4)];
                a[0] = q0.x & z[0]; a[1] = q0.y & z[1]; a[2] = q0.z & z[2]; a[3] = q0.w & z[3];
                cm[0] = q0.x & d[0]; cm[1] = q0.y & d[1]; cm[2] = q0.z & d[2]; cm[3] = q0.w & d[3];
                if (BW == 8) {
                    const uint4 q1 = cT[i * (BW / 4) + 1];
                    a[BW - 4] = q1.x & z[BW - 4]; a[BW - 3] = q1.y & z[BW - 3]; a[BW - 2] = q1.z & z[BW - 2]; a[BW - 1] = q1.w & z[BW - 1];
                    cm[BW - 4] = q1.x & d[BW - 4]; cm[BW - 3] = q1.y & d[BW - 3]; cm[BW - 2] = q1.z & d[BW - 2]; cm[BW - 1] = q1.w & d[BW - 1];
                }
                #pragma unroll
                for (int g = 0; g < BW; ++g) anyc |= cm[g];
            }
            if (!mode && anyc) continue;   // early pass: the row waits for the winners
            // owner of row i = the lowest live (early) / candidate-holding (late) sample spiking at pixel i
            uint32_t lower = 0;
            #pragma unroll
            for (int g = 0; g < BW; ++g) {
                const uint32_t below = g < (bb >> 5) ? 0xffffffffu : (g == (bb >> 5) ? ((1u << (bb & 31)) - 1u) : 0u);
                lower |= (mode ? cm[g] : a[g]) & below;
            }
            if (lower) continue;
            float U0 = 0.f, U1 = 0.f, U2 = 0.f, U3 = 0.f;
            uint32_t anya = 0;
            #pragma unroll 1
            for (int g = 0; g < BW; ++g) {
                uint32_t mm = a[g];
                anya |= mm;
                while (mm) {
                    const int b2 = g * 32 + __ffs(mm) - 1;
                    mm &= mm - 1;
                    const float4 t4 = *(const float4 *)(c_.tx + b2 * (4 * CG) + 4 * c4);
                    U0 = U0 + t4.x; U1 = U1 + t4.y; U2 = U2 + t4.z; U3 = U3 + t4.w;
                }
            }
            if (!c_.pre_on) anya = 0;
            if (!(anya | gwin)) continue;
            float *wp = c_.W + i * WS + 4 * c4;
            const float4 w4 = *(const float4 *)wp;
            float wv[4] = {w4.x, w4.y, w4.z, w4.w};
            const float Uv[4] = {U0, U1, U2, U3};
            #pragma unroll
            for (int c = 0; c < 4; ++c) {
                float w = wv[c];
                if (anya) w = w - Uv[c] * c_.dts;  // x * 1.0f is exact: the classic rule's missing dt factor is dts = 1
                if ((gwin >> c) & 1u) {  // the column's single winner (fast late pass): post term
                    uint32_t e = M.wl[c4][0];
                    #pragma unroll 1
                    for (int k2 = 1; k2 < M.nwl[c4]; ++k2) if ((M.wl[c4][k2] >> 16) == (uint32_t)(4 * c4 + c)) e = M.wl[c4][k2];
                    const float V = 0.0f + xval(cx, xrow[(e & 0xffu) * P + i], (int)((e >> 8) & 0xffu), i, tstep) * c_.nu1;
                    w = w + V * c_.dts;
                }
                if (c_.has_clamp) {
                    const float wc = clampf(w, c_.wmin, c_.wmax);
                    if (!mode && anyd && wc != w) deferred = true;   // early pass, post term possibly still to come: clamp later
                    else w = wc;
                }
                wv[c] = w;
            }
            *(float4 *)wp = make_float4(wv[0], wv[1], wv[2], wv[3]);
        }
    }
    if (deferred) sh(c_.M)->defer[c4] = 1;
}

template <int CG, int BW>
__device__ __noinline__ void stdp_list2(const PassCtx2 *cx, int sb, int c4, int mode, uint32_t gwin, const uint32_t *dm, int par_,
                                        const uint8_t *xrow, int tstep, int tid0, int nthr0) {
    stdp_list_body<CG, BW>(cx, sb, c4, mode, gwin, dm, par_, xrow, tstep, tid0, nthr0);
}

// Post term of the fast late pass of column group c4 for the rows stdp_list2 did not touch: per winner
// (column, staged row): w + x_pre[b,i]*nu1*dt, clamp (MCC_learning.py:267-299, 86-110).  A row was handled
// by stdp_late2 iff a candidate-holding sample (`dm`) spiked at its pixel.  Rows whose pre-synaptic trace is
// exactly zero are skipped when `skip0` says that is exact: w + 0*nu1*dt == w bitwise unless w is -0.0 (the
// tile holds none: checked when it is loaded), and the clamp of an in-range weight is the identity.
template <int CG, int BW>
__device__ __noinline__ void post_rows2(const PassCtx2 *cx, int sb, int c4, int nwl, const uint32_t *dm, const uint8_t *xrow, int tstep, int skip0,
                                        int tid0, int nthr0) {
    const PassK c_ = load_k(cx);
    dm = sh(dm); xrow = sh(xrow);
    const int P = c_.P, WS = c_.WS;
    const Misc2 &M = *c_.M;
    const uint4 *cT = (const uint4 *)(c_.inT + sb * P * BW);
    uint32_t z[BW];   // the samples whose rows stdp_late2 handles
    {
        const uint4 z0 = *(const uint4 *)dm;
        z[0] = z0.x; z[1] = z0.y; z[2] = z0.z; z[3] = z0.w;
        if (BW == 8) { const uint4 z1 = *(const uint4 *)(dm + 4); z[BW - 4] = z1.x; z[BW - 3] = z1.y; z[BW - 2] = z1.z; z[BW - 1] = z1.w; }
    }
    #pragma unroll 1
    for (int k = 0; k < nwl; ++k) {   // distinct columns
        const uint32_t e = M.wl[c4][k];
        const uint8_t *xr = xrow + (e & 0xffu) * P;
        const int wb = (int)((e >> 8) & 0xffu);
        float *wcol = c_.W + (e >> 16);
        // four rows per thread and round with every load issued before the first use: the pass is a chain of
        // dependent shared-memory loads otherwise (it sits on the critical path of the step)
        #pragma unroll 1
        for (int i0 = tid0; i0 < P; i0 += 4 * nthr0) {
            uint32_t age[4], any[4];
            float w[4];
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = min(i0 + u * nthr0, P - 1);
                age[u] = xr[i];
                const uint4 q0 = cT[i * (BW / 4)];
                any[u] = (q0.x & z[0]) | (q0.y & z[1]) | (q0.z & z[2]) | (q0.w & z[3]);
                if (BW == 8) { const uint4 q1 = cT[i * (BW / 4) + 1]; any[u] |= (q1.x & z[BW - 4]) | (q1.y & z[BW - 3]) | (q1.z & z[BW - 2]) | (q1.w & z[BW - 1]); }
                if (i0 + u * nthr0 >= P) any[u] = 1u;   // past the last row
            }
            #pragma unroll
            for (int u = 0; u < 4; ++u) w[u] = any[u] ? 0.0f : wcol[(i0 + u * nthr0) * WS];   // rows the list pass owns are not even read
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * nthr0;
                if (any[u]) continue;   // rows with a candidate-holding sample's spike: done by the list pass
                const float xv = xval(cx, age[u], wb, i, tstep);
                if (skip0 && xv == 0.0f) continue;
                const float V = 0.0f + xv * c_.nu1;
                float wn = w[u] + V * c_.dts;
                if (c_.has_clamp) wn = clampf(wn, c_.wmin, c_.wmax);
                wcol[i * WS] = wn;
            }
        }
    }
}

// STDP of one step of column group c4 in row form, general: pre and post term of a column applied together
// (pre, post, clamp — the reference's order).  Used for slots with an overflowed event list, more than XR
// winners, two winners in one column, a winner whose trace row is not staged.
// `cand` = staged samples (rows of xrow), `xsrc` = the step's input-trace ages in global memory.  `dm` != NULL: the
// pre term only on the rows stdp_list2 left alone (a sample of `dm` spiked there); the post term on every row.
template <int CG, int BW>
__device__ __noinline__ void stdp_rows2(const PassCtx2 *cx, int sb, int c4, uint32_t gwin, const int *cand, int ns, const uint8_t *xsrc,
                                        const uint8_t *xrow, int tstep, const uint32_t *dm, int tid0, int nthr0) {
    const PassK c_ = load_k(cx);
    if (dm) dm = sh(dm);
    xrow = sh(xrow); cand = sh(cand);
    const int P = c_.P, WS = c_.WS, TJ = 4 * CG;
    const Misc2 &M = *c_.M;
    const uint4 *cT = (const uint4 *)(c_.inT + sb * P * BW);
    const uint4 z0 = c_.pre_on ? *(const uint4 *)&M.nz4[c4][0] : make_uint4(0, 0, 0, 0);
    #pragma unroll 1
    for (int i = tid0; i < P; i += nthr0) {
        uint32_t m[BW];
        const uint4 q0 = cT[i * (BW / 4)];
        m[0] = q0.x & z0.x; m[1] = q0.y & z0.y; m[2] = q0.z & z0.z; m[3] = q0.w & z0.w;
        uint32_t anym = m[0] | m[1] | m[2] | m[3];
        if (BW == 8) {
            const uint4 q1 = cT[i * (BW / 4) + 1];
            const uint4 z1 = c_.pre_on ? *(const uint4 *)&M.nz4[c4][4] : make_uint4(0, 0, 0, 0);
            m[BW - 4] = q1.x & z1.x; m[BW - 3] = q1.y & z1.y; m[BW - 2] = q1.z & z1.z; m[BW - 1] = q1.w & z1.w;
            anym |= m[BW - 4] | m[BW - 3] | m[BW - 2] | m[BW - 1];
        }
        if (dm) {   // rows already finished by the early pass: no pre term
            uint32_t df = (q0.x & dm[0]) | (q0.y & dm[1]) | (q0.z & dm[2]) | (q0.w & dm[3]);
            if (BW == 8) { const uint4 q1 = cT[i * (BW / 4) + 1]; df |= (q1.x & dm[4]) | (q1.y & dm[5]) | (q1.z & dm[6]) | (q1.w & dm[7]); }
            if (!df) anym = 0u;
        }
        const bool pre_t = anym != 0u;
        if (!(pre_t || gwin)) continue;
        float U[4] = {0.f, 0.f, 0.f, 0.f};
        if (pre_t) {
            #pragma unroll 1
            for (int g = 0; g < BW; ++g) {
                uint32_t mm = m[g];
                while (mm) {
                    const int bb = g * 32 + __ffs(mm) - 1;
                    mm &= mm - 1;
                    const float4 t4 = *(const float4 *)(c_.tx + bb * TJ + 4 * c4);
                    U[0] = U[0] + t4.x; U[1] = U[1] + t4.y; U[2] = U[2] + t4.z; U[3] = U[3] + t4.w;
                }
            }
        }
        float *wp = c_.W + i * WS + 4 * c4;
        const float4 w4 = *(const float4 *)wp;
        float wv[4] = {w4.x, w4.y, w4.z, w4.w};
        #pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            const bool post_t = (gwin >> c) & 1u;
            float w = wv[c];
            if (pre_t) w = w - U[c] * c_.dts;
            if (post_t) {
                float V = 0.0f;
                #pragma unroll 1
                for (int g = 0; g < BW; ++g) {  // winners of this column, ascending sample order
                    uint32_t mm = M.wmask[4 * c4 + c][g];
                    while (mm) {
                        const int bb = g * 32 + __ffs(mm) - 1;
                        mm &= mm - 1;
                        int sl = -1;
                        for (int q = 0; q < ns; ++q) if (cand[q] == bb) sl = q;
                        const float xv = xval(cx, sl >= 0 ? xrow[sl * P + i] : __ldcg(xsrc + (size_t)bb * P + i), bb, i, tstep);
                        V = V + xv * c_.nu1;
                    }
                }
                w = w + V * c_.dts;
            }
            if (c_.has_clamp) w = clampf(w, c_.wmin, c_.wmax);
            wv[c] = w;
        }
        *(float4 *)wp = make_float4(wv[0], wv[1], wv[2], wv[3]);
    }
}

// Spike-gather of a sample whose event list overflowed EV_CAP: walk its bit row in global memory
// (rare; out of line to keep the hot loop small).
__device__ __noinline__ float4 gather_dense2(const uint32_t *row, int SW, const float *Wc, int WS) {
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

// T-fold sequential  rc = rc - dt  (what the refractory counter of an undisturbed neuron does over the
// window, nodes.py:514); closed form when every intermediate value is an exactly representable integer.
__device__ __forceinline__ float refrac_replay(float rc, float dt, int T) {
    if (dt == 1.0f && rc == truncf(rc) && fabsf(rc) < 4194304.0f && T < 4194304) return rc - (float)T;
    #pragma unroll 1
    for (int k = 0; k < T; ++k) rc = rc - dt;
    return rc;
}

// Spike-gather of 4 columns (Wc = the thread's column group in row 0 of the tile) for the spikes of sample b in
// list block `blk` (slot `slot`): p[c] = sum_{i in sX[b]} W[i][c], i ascending (topology.py:437-479).  One copy
// for the step itself and for the gather that runs ahead in the shadow of the exchange.
template <int WS>
__device__ __forceinline__ float4 gather2(const PassCtx2 *cx, const unsigned char *blk, int slot, int b, const float *Wc) {
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
    cx = sh(cx); blk = sh(blk); Wc = sh(Wc);
    const int cnt = ((const uint16_t *)blk)[b];
    const int B = cx->k.B;
    if (cnt > EV_CAP)   // dense sample: walk the bit row in global memory (rare, slow path)
        return gather_dense2(cx->inS + ((size_t)slot * B + b) * cx->SW, cx->SW, Wc, WS);
    const uint2 *l4 = (const uint2 *)(blk + cx->k.cntb) + b;
    #pragma unroll 1
    for (int k = 0; k < cnt; k += 4) {
        const uint2 q = l4[(k >> 2) * B];  // 4 pixel indices; tail padded with P (zero row)
        const float4 r0 = *(const float4 *)(Wc + (q.x & 0xffffu) * WS);
        const float4 r1 = *(const float4 *)(Wc + (q.x >> 16) * WS);
        const float4 r2 = *(const float4 *)(Wc + (q.y & 0xffffu) * WS);
        const float4 r3 = *(const float4 *)(Wc + (q.y >> 16) * WS);
        p0 = p0 + r0.x; p1 = p1 + r0.y; p2 = p2 + r0.z; p3 = p3 + r0.w;
        p0 = p0 + r1.x; p1 = p1 + r1.y; p2 = p2 + r1.z; p3 = p3 + r1.w;
        p0 = p0 + r2.x; p1 = p1 + r2.y; p2 = p2 + r2.z; p3 = p3 + r2.w;
        p0 = p0 + r3.x; p1 = p1 + r3.y; p2 = p2 + r3.z; p3 = p3 + r3.w;
    }
    return make_float4(p0, p1, p2, p3);
}

// Ai spike monitors (monitors.py:94-111; the launch code cleared the raster): rarely compiled-in work, out of line.
__device__ __noinline__ void ai_monitor2(const PassCtx2 *cx, int sb_, int col, int t) {
    if (cx->I_rec_s) cx->I_rec_s[((size_t)t * cx->k.B + sb_) * cx->n + cx->j0 + col] = 1;
    if (cx->I_rec_count) atomicAdd(cx->I_rec_count + (size_t)sb_ * cx->n + cx->j0 + col, 1);
}

// One LIFNodes.forward step (nodes.py:500-529) of Ai list entry e with input xin at step t; a spike goes to the
// exchange (`sis`: the step's Ai spike counts per sample), to the tile's own-spike bits (`spk`) and to the monitors.
__device__ __forceinline__ void ai_step_body(const PassCtx2 *cx, int e, float xin, int t, unsigned int *sis, uint32_t *spk) {
    cx = sh(cx); spk = sh(spk);
    float *const ai_v = sh(cx->ai_v), *const ai_rc = sh(cx->ai_rc);
    uint8_t *const ai_fl = sh(cx->ai_fl);
    float v = ai_v[e], rc = ai_rc[e];
    uint32_t fl = ai_fl[e];
    v = cx->I_decay * (v - cx->I_rest) + cx->I_rest;
    if (!(fl & 1u)) { if (rc > 0.0f) xin = 0.0f; rc = rc - cx->I_dt; }   // undisturbed counter: <= 0 by construction
    v = v + xin;
    fl &= 1u;
    if (v >= cx->I_thresh) {
        rc = cx->I_refrac; v = cx->I_reset; fl = 2u;
        const int id = sh(cx->ai_id)[e], sb_ = id >> 8, col = id & 0xff;
        atomicOr(spk + sb_, 1u << col);
        atomicAdd(sis + sb_, 1u);
        if (cx->I_mon) ai_monitor2(cx, sb_, col, t);
    }
    ai_v[e] = v; ai_rc[e] = rc; ai_fl[e] = (uint8_t)fl;
}

__device__ __noinline__ void ai_step2(const PassCtx2 *cx, int e, float xin, int t, unsigned int *sis, uint32_t *spk) {
    ai_step_body(cx, e, xin, t, sis, spk);
}

// A thread found threshold crossers among its 4 neurons at step t (`cand`, nodes.py:1088-1092): count them for
// theta (nodes.py:1093-1094), send the best one_spike key to the exchange (nodes.py:1097-1105), claim the partner
// Ai neurons for step t+1, mark the column group / sample as candidate-holding, and stage the sample's input-trace
// row of step t for a possible post term.  Rare per thread: kept out of line.
__device__ __noinline__ void on_candidate2(const PassCtx2 *cx, uint32_t cand, int b, int cg, int t) {
    cx = sh(cx);
    Misc2 &M = *sh(cx->k.M);
    uint16_t *const ai_map = sh(cx->ai_map), *const ai_claim = sh(cx->ai_claim);
    const int par = t & 1, TJ = cx->TJ, jc = cx->j0 + 4 * cg, B = cx->k.B, Bp = cx->k.Bp, P = cx->k.P;
    unsigned long long mykey = 0ull;
    #pragma unroll 1
    for (int c = 0; c < 4; ++c)
        if ((cand >> c) & 1u) {
            const int col = 4 * cg + c;
            atomicAdd(&M.cnt[par][col], 1);
            if (cx->one_spike) {
                const unsigned long long k2 = snn_one_spike_key(cx->seed, (uint32_t)t + cx->step_offset, (uint32_t)cx->liE, (uint32_t)b, (uint32_t)(jc + c));
                mykey = k2 > mykey ? k2 : mykey;
            }
            // the partner Ai neuron is this thread's at step t+1, whether the candidate wins or not
            unsigned e = ai_map[b * TJ + col];
            if (e == AI_NONE) {
                e = (unsigned)atomicAdd(&M.nact, 1);
                sh(cx->ai_v)[e] = cx->I_rest; sh(cx->ai_rc)[e] = 0.0f; sh(cx->ai_id)[e] = (uint16_t)((b << 8) | col); sh(cx->ai_fl)[e] = 1;
                ai_claim[par * cx->aicap + e] = (uint16_t)AI_NONE;
                ai_map[b * TJ + col] = (uint16_t)e;
            }
            ai_claim[(par ^ 1) * cx->aicap + e] = (uint16_t)(t + 1);   // slot of parity (t + 1) & 1
        }
    atomicOr(&M.candgrp[par], 1u << cg);
    atomicOr(&M.candmask[par][cg][b >> 5], 1u << (b & 31));
    sh((uint8_t *)cx->k.candlist)[(par * 8 + cg) * Bp + atomicAdd(&M.ncs[par][cg], 1)] = (uint8_t)b;
    if (cx->one_spike) atomicMax(cx->win + (t % 3) * B + b, mykey);
    if (cx->stage_on) {  // stage x_pre[b,:] of step t for the post term, once per sample
        const uint32_t old = atomicExch(sh(cx->candstamp) + par * Bp + b, (uint32_t)(t + 1));
        if (old != (uint32_t)(t + 1)) {
            const int s = atomicAdd(&M.ncand[par], 1);
            if (s < XR && !cx->nostage) {
                M.candb[par][s] = b;
                sh(cx->candslot)[par * Bp + b] = s;
                mbar_expect_tx(&M.mbar_x[par], (uint32_t)P);
                bulk_g2s(sh(cx->xrow_w) + (par * XR + s) * P, cx->xage + ((size_t)t * B + b) * P, (uint32_t)P, &M.mbar_x[par]);
            } else sh(cx->candslot)[par * Bp + b] = -1;
        }
    }
}

// CG: float4 column groups per CTA (TJ = 4 CG columns); BW: 32-bit words of a per-pixel sample mask
// (4 -> B <= 128).  Threads = Bp * CG compute threads, thread (cg, b), Bp = B rounded up to a multiple
// of 32, plus one exchange warp.  The Bp threads of a column group form a pipeline of their own (named
// barrier 2 + cg): the CTA-wide barriers are the two per step that frame the exchange.
// VAR bit 1: phase timers compiled in.
template <int CG, int BW, int VAR>
__global__ void __launch_bounds__((32 * BW * CG + 32 < 1024 ? 32 * BW * CG + 32 : 1024), 1)
snn_dc2_window(const __grid_constant__ F2Params Q) {
    constexpr bool PROFV = (VAR & 2) != 0;
    constexpr int TJ = 4 * CG;
    constexpr int WS = (CG & 1) ? TJ : TJ + 4;
#ifdef SNN_EMU
    unsigned char *smem = (unsigned char *)emu::tls_cta->dyn_smem;
#else
    extern __shared__ __align__(16) unsigned char smem[];
#endif
    const int B = Q.B, Bp = Q.Bp, P = Q.P, n = Q.n, T = Q.T;
    const unsigned int G = gridDim.x;
    float *W = (float *)(smem + Q.o_W);
    float *tx = (float *)(smem + Q.o_tx);
    unsigned char *evb = smem + Q.o_ev;
    uint32_t *inT = (uint32_t *)(smem + Q.o_inT);
    uint8_t *xrow = smem + Q.o_xrow;             // [2][XR][P] staged age rows
    float *rep = (float *)(smem + Q.o_rep);
    float *theta_s = (float *)(smem + Q.o_theta);   // [32] theta, [32] thresh + decayed theta
    float *thr_s = theta_s + 32;
    float *dtab = theta_s + 64;                  // [256] input trace by age: trace_scale * decay^age, multiplied up step by step
    uint8_t *live = smem + Q.o_live;             // [CG][Bp] live samples per column group
    unsigned long long *keyT = (unsigned long long *)(smem + Q.o_tab);  // [2][Bp] one_spike arg-max key of a step, by step parity
    uint32_t *isumT = (uint32_t *)(keyT + 2 * Bp);                      // [2][Bp] Ai spikes of a step
    uint32_t *aispk = isumT + 2 * Bp;                                   // [2][Bp] bit col: Ai (b, col) of this tile spiked in that step
    uint32_t *candstamp = aispk + 2 * Bp;                               // [2][Bp] by step parity: step + 1 of the sample's last staged row
    int *candslot = (int *)(candstamp + 2 * Bp);                        // [2][Bp] its slot in xrow (-1: not staged)
    uint8_t *candlist = (uint8_t *)(candslot + 2 * Bp);                 // [2][8][Bp] by step parity, per column group: candidate-holding samples
    const int aicap = Bp * TJ;
    float *ai_v = (float *)(smem + Q.o_ai);
    float *ai_rc = ai_v + aicap;
    uint16_t *ai_id = (uint16_t *)(ai_rc + aicap);      // b << 8 | col
    uint16_t *ai_claim = ai_id + aicap;                  // [2][cap] by step parity: step in which the neuron's owner thread (not the
                                                         // list pass) runs it
    uint16_t *ai_map = ai_claim + 2 * aicap;             // (b * TJ + col) -> list entry, AI_NONE
    uint8_t *ai_fl = (uint8_t *)(ai_map + aicap);        // bit 0: refractory counter still the undisturbed one; bit 1: spiked last
                                                         // step; bit 2: partner spiked at step -1 (input at step 0)
    Misc2 &M = *(Misc2 *)(smem + Q.o_misc);
#ifdef SNN_EMU
    PassCtx2 &s_cx = *(PassCtx2 *)emu::tls_cta->static_smem;
#else
    __shared__ PassCtx2 s_cx;
#endif

    const int NC = Bp * CG;                        // compute threads; the warp above them is the exchange warp
    const int tid = threadIdx.x, lane = tid & 31;
    const bool isx = tid >= NC;
    const int cg = isx ? 0 : tid / Bp, b = isx ? Bp : tid - cg * Bp;   // state ownership: sample b, neurons jc..jc+3 (warp-uniform cg)
    const int j0 = blockIdx.x * TJ;
    const int jc = j0 + 4 * cg;
    const bool act = !isx && b < B && jc < n;
    const int gbar = 2 + cg;                       // the column group's named barrier
    const snn_layer_t &E = Q.E, &I = Q.I;
    const snn_conn_t &C = Q.C;
    const bool stdp = C.rule >= SNN_RULE_POSTPRE;
    const bool pre_on = stdp && C.nu0 != 0.0f, post_on = stdp && C.nu1 != 0.0f;
    const bool update_on = Q.learning && stdp;
    const bool stage_on = update_on && post_on && Q.xage != nullptr;
    const float dts = C.rule == SNN_RULE_MCC_POSTPRE ? C.dt_scale : 1.0f;
    const int evblk = (int)al16((size_t)Q.SB);
    const int cntb = ev_count_bytes(B);
    float *Wc = W + 4 * cg;
    long long *pc = M.pc;
    long long pt = clock64();
    #define PROF(k) { if (PROFV && tid == 0) { const long long now_ = clock64(); pc[k] += now_ - pt; pt = now_; } }
    #define PROFX(k) { if (PROFV && tid == NC) { const long long now_ = clock64(); pc[k] += now_ - pt; pt = now_; } }

    // ---- prologue: W tile, theta, inhibition table, tables, Ai list, state registers -----------
    bool negz = false;
    {
        constexpr int U = 8;
        const int total = (P + 1) * TJ, nthr = blockDim.x;
        #pragma unroll 1
        for (int base = 0; base < total; base += U * nthr) {
            float v[U];
            #pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = base + u * nthr + tid, i = idx / TJ, jj = idx - i * TJ;
                v[u] = (idx < total && i < P && j0 + jj < n) ? __ldg(C.w + (size_t)i * n + j0 + jj) : 0.0f;
            }
            #pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = base + u * nthr + tid, i = idx / TJ, jj = idx - i * TJ;
                if (idx < total) { W[i * WS + jj] = v[u]; negz |= __float_as_uint(v[u]) == 0x80000000u; }
            }
        }
    }
    #pragma unroll 1
    for (int jj = tid; jj < 32; jj += blockDim.x) {
        const float th = (jj < TJ && j0 + jj < n) ? E.theta[j0 + jj] : 0.0f;
        theta_s[jj] = th;
        thr_s[jj] = E.thresh + (E.learning ? th * E.theta_decay : th);   // nodes.py:1078-1079, 1088
    }
    #pragma unroll 1
    for (int k = tid; k <= Q.nrep; k += blockDim.x) rep[k] = Q.rep[k];
    #pragma unroll 1
    for (int k = tid; k < 2 * Bp; k += blockDim.x) { keyT[k] = 0ull; isumT[k] = 0u; aispk[k] = 0u; }
    #pragma unroll 1
    for (int k = tid; k < 2 * Bp; k += blockDim.x) { candstamp[k] = 0u; candslot[k] = -1; }
    {
        uint32_t *m32 = (uint32_t *)ai_map;   // aicap is even (Bp is a multiple of 32)
        #pragma unroll 1
        for (int k = tid; k < aicap / 2; k += blockDim.x) m32[k] = 0xFFFFFFFFu;
    }
    #pragma unroll 1
    for (int k = tid; k < 64; k += blockDim.x) { (&M.nz4[0][0])[k] = 0; (&M.cnt[0][0])[k] = 0; }
    #pragma unroll 1
    for (int k = tid; k < 32 * 8; k += blockDim.x) (&M.wmask[0][0])[k] = 0;
    if (tid == 0) {
        mbar_init(&M.mbar_in[0], 1);
        mbar_init(&M.mbar_in[1], 1);
        mbar_init(&M.mbar_x[0], 1);
        mbar_init(&M.mbar_x[1], 1);
#ifndef SNN_EMU
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
        M.ncand[0] = M.ncand[1] = 0; M.candgrp[0] = M.candgrp[1] = 0; M.abort = 0; M.nact = 0; M.negzero = 0;
        for (int g = 0; g < 8; ++g) { M.colwin[g] = 0; M.nwl[g] = 0; M.nlive[g] = 0; M.defer[g] = 0; }
        for (int k = 0; k < 2 * 8 * 8; ++k) (&M.candmask[0][0][0])[k] = 0;
        for (int k = 0; k < 16; ++k) (&M.ncs[0][0])[k] = 0;
        M.denseflag[0] = Q.dense[0]; M.denseflag[1] = T >= 1 ? Q.dense[1] : 0;
        for (int k = 0; k < NPROF; ++k) M.pc[k] = 0;
        s_cx.k.W = W; s_cx.k.tx = tx; s_cx.k.inT = inT; s_cx.k.evb = evb; s_cx.k.live = live; s_cx.k.candlist = candlist; s_cx.k.M = &M;
        s_cx.k.P = P; s_cx.k.B = B; s_cx.k.Bp = Bp; s_cx.k.evblk = evblk; s_cx.k.cntb = cntb; s_cx.k.WS = WS;
        s_cx.k.pre_on = pre_on; s_cx.k.has_clamp = C.has_clamp;
        s_cx.k.dts = dts; s_cx.k.wmin = C.wmin; s_cx.k.wmax = C.wmax; s_cx.k.nu1 = C.nu1;
        s_cx.inS = Q.inS; s_cx.SW = Q.SW;
        s_cx.ai_v = ai_v; s_cx.ai_rc = ai_rc; s_cx.ai_id = ai_id; s_cx.ai_claim = ai_claim; s_cx.ai_map = ai_map; s_cx.ai_fl = ai_fl;
        s_cx.aispk = aispk; s_cx.sisum = Q.sisum; s_cx.win = Q.win; s_cx.I_rec_s = I.rec_s; s_cx.I_rec_count = I.rec_count; s_cx.I_mon = (I.rec_s || I.rec_count) ? 1 : 0;
        s_cx.I_decay = I.decay; s_cx.I_rest = I.rest; s_cx.I_dt = I.dt; s_cx.I_thresh = I.thresh; s_cx.I_refrac = I.refrac; s_cx.I_reset = I.reset;
        s_cx.n = n; s_cx.j0 = j0; s_cx.TJ = TJ; s_cx.aicap = aicap;
        s_cx.candstamp = candstamp; s_cx.candslot = candslot; s_cx.xrow_w = xrow; s_cx.xage = Q.xage;
        s_cx.x0c = Q.x0c; s_cx.dtab = dtab; s_cx.x_decay = Q.X.trace_decay; s_cx.anyx0 = Q.anyx0 ? *Q.anyx0 : 0;
        {   // the trace a pixel carries `age` steps after its spike: the scale, then one multiplication per step
            float x = Q.X.trace_scale;
            for (int k = 0; k < 256; ++k) { dtab[k] = x; x = x * Q.X.trace_decay; }
        }
        s_cx.seed = Q.seed; s_cx.step_offset = Q.step_offset; s_cx.liE = Q.liE; s_cx.one_spike = E.one_spike; s_cx.stage_on = stage_on ? 1 : 0;
        s_cx.nostage = (PROFV && (Q.dbg & 8)) ? 1 : 0;
    }
    __syncthreads();

    if (negz) M.negzero = 1;
    float vE[4], rE[4], xE[4];
    uint32_t candE = 0, pend = 0, sEfin = 0;  // 4-bit masks over my neurons
    uint32_t vm = 0;                          // my neurons that exist (column < n)
    {
        const bool vec = act && (n & 3) == 0 && jc + 3 < n;
        const size_t k0 = act ? (size_t)b * n + jc : 0;
        float v0[4], r0[4];
        uint32_t sE0 = 0, sI0 = 0;
        if (vec) {
            const float4 a = *(const float4 *)(E.v + k0), r = *(const float4 *)(E.refrac_count + k0);
            const float4 x = E.traces ? *(const float4 *)(E.x + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 iv = *(const float4 *)(I.v + k0), ir = *(const float4 *)(I.refrac_count + k0);
            const uint32_t es = *(const uint32_t *)(E.s + k0), is = *(const uint32_t *)(I.s + k0);
            vE[0] = a.x; vE[1] = a.y; vE[2] = a.z; vE[3] = a.w; rE[0] = r.x; rE[1] = r.y; rE[2] = r.z; rE[3] = r.w;
            xE[0] = x.x; xE[1] = x.y; xE[2] = x.z; xE[3] = x.w;
            v0[0] = iv.x; v0[1] = iv.y; v0[2] = iv.z; v0[3] = iv.w; r0[0] = ir.x; r0[1] = ir.y; r0[2] = ir.z; r0[3] = ir.w;
            #pragma unroll
            for (int c = 0; c < 4; ++c) { if ((es >> (8 * c)) & 0xffu) sE0 |= 1u << c; if ((is >> (8 * c)) & 0xffu) sI0 |= 1u << c; }
            vm = 0xFu;
        } else {
            #pragma unroll
            for (int c = 0; c < 4; ++c) {
                const bool ok = act && jc + c < n;
                if (ok) vm |= 1u << c;
                const size_t k = ok ? k0 + c : 0;
                vE[c] = ok ? E.v[k] : 0.0f; rE[c] = ok ? E.refrac_count[k] : 0.0f;
                xE[c] = (ok && E.traces) ? E.x[k] : 0.0f;
                v0[c] = ok ? I.v[k] : I.rest; r0[c] = ok ? I.refrac_count[k] : 0.0f;
                if (ok && E.s[k]) sE0 |= 1u << c;
                if (ok && I.s[k]) sI0 |= 1u << c;
            }
        }
        // Ai: a neuron that is not exactly at rest, is refractory, or whose partner spiked at step -1 starts
        // in the list (nodes.py:500-529); everything else is at rest until a spike arrives
        if (sI0) atomicOr(&aispk[Bp + b], sI0 << (4 * cg));   // parity 1 = step -1
        #pragma unroll
        for (int c = 0; c < 4; ++c)
            if (((vm >> c) & 1u) && (v0[c] != I.rest || r0[c] > 0.0f || ((sE0 >> c) & 1u))) {
                const int e = atomicAdd(&M.nact, 1);
                ai_v[e] = v0[c]; ai_rc[e] = r0[c]; ai_id[e] = (uint16_t)((b << 8) | (4 * cg + c));
                ai_claim[e] = (uint16_t)AI_NONE; ai_claim[aicap + e] = (uint16_t)AI_NONE; ai_fl[e] = ((sE0 >> c) & 1u) ? 4 : 0;
                ai_map[b * TJ + 4 * cg + c] = (uint16_t)e;
            }
    }
    if (!isx && cg == 0 && b < B) isumT[Bp + b] = Q.sisum0[b];
    if (!isx) *(float4 *)(tx + b * TJ + 4 * cg) = (act && stdp) ? make_float4(xE[0] * C.nu0, xE[1] * C.nu0, xE[2] * C.nu0, xE[3] * C.nu0)
                                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
    bool livep = false;  // my (sample, column group) pair has a non-zero Ae trace
    if (act && stdp && (xE[0] != 0.0f || xE[1] != 0.0f || xE[2] != 0.0f || xE[3] != 0.0f)) {
        livep = true;
        atomicOr(&M.nz4[cg][b >> 5], 1u << (b & 31));
        live[cg * Bp + atomicAdd(&M.nlive[cg], 1)] = (uint8_t)b;
    }
    const uint32_t bytesE = (uint32_t)Q.SB, bytesT = (uint32_t)(sizeof(uint32_t) * (size_t)P * BW);
    const int tid_pf = NC + 1;  // the thread that issues the slot prefetches (exchange warp)
    if (tid == tid_pf) {  // stage slot 0 (spikes of step -1) and slot 1 (spikes of step 0)
        mbar_arrive_expect_tx(&M.mbar_in[0], bytesE + bytesT);
        bulk_g2s(evb, Q.evS, bytesE, &M.mbar_in[0]);
        bulk_g2s(inT, Q.inT, bytesT, &M.mbar_in[0]);
        mbar_arrive_expect_tx(&M.mbar_in[1], bytesE + bytesT);
        bulk_g2s(evb + evblk, Q.evS + Q.SB, bytesE, &M.mbar_in[1]);
        bulk_g2s(inT + P * BW, Q.inT + (size_t)P * BW, bytesT, &M.mbar_in[1]);
    }
    __syncthreads();
    if (tid == 0) M.nact_snap = M.nact;
    PROF(0)  // prologue
    if (PROFV && tid == NC) pt = clock64();

    float4 pre = make_float4(0.f, 0.f, 0.f, 0.f);   // currents of the step about to run, gathered ahead
    bool havepre = false;

    // =====================================================================================
    for (int t = 0; t <= T; ++t) {
        const int buf = t & 1;                       // slot t = spikes of step t-1
        const unsigned char *cE = evb + buf * evblk;
        const int par = t & 1, ppar = par ^ 1;       // parity of step t / of step t-1
        unsigned int *const sis_t = Q.sisum + (t % 3) * B;   // this step's Ai spike counts (exchange slot t % 3)
        const int dense_nb = __ldg(Q.dense + (t + 1 <= T ? t + 1 : T));   // slot t+1 holds an overflowed event list (used in the shadow)

        // ---- exchange of step t-1 lands: the exchange warp waits for the grid barrier and copies the
        // per-sample words of slot (t-1) % 3 into the shared-memory tables of parity `ppar`
        if (isx && t > 0) {
            if (lane == 0) {
                const unsigned int target = G * (unsigned int)t;
                unsigned int spins = 0;
                while ((int)(ld_relaxed_u32(Q.bar) - target) < 0 && !(PROFV && (Q.dbg & 16))) {
#if V2_POLL_NS
                    __nanosleep(V2_POLL_NS);
#endif
                    if ((++spins & 0xfffffu) == 0) {   // ~ a second of polling: give up
                        if (spins > (4u << 20)) { if (Q.err) atomicOr(Q.err, SNN_ERR_BARRIER); M.abort = 1; break; }
                    }
                }
#ifdef SNN_EMU
                __atomic_thread_fence(__ATOMIC_SEQ_CST);
#else
                asm volatile("fence.acquire.gpu;" ::: "memory");
#endif
            }
            __syncwarp();
            if (PROFV && tid == NC && Q.prof && t > 100 && t <= 132) Q.prof[160 * NPROF + ((t - 101) * 160 + blockIdx.x) * 2 + 1] = clock64() - pt;
            PROFX(10)  // barrier wait (exchange warp)
            const int xs = (t - 1) % 3;
            {   // all loads in flight together: one L2 round trip
                unsigned long long kk[BW];
                unsigned int ss[BW];
                #pragma unroll
                for (int u = 0; u < BW; ++u) {
                    const int k = lane + 32 * u;
                    kk[u] = k < B ? __ldcg(Q.win + xs * B + k) : 0ull;
                    ss[u] = k < B ? __ldcg(Q.sisum + xs * B + k) : 0u;
                }
                #pragma unroll
                for (int u = 0; u < BW; ++u) {
                    const int k = lane + 32 * u;
                    if (k < Bp) { keyT[ppar * Bp + k] = kk[u]; isumT[ppar * Bp + k] = ss[u]; }
                }
            }
            PROFX(11)  // exchange read
        }
        __syncthreads();
        if (PROFV && tid == NC) pt = clock64();
        const long long g_t0 = PROFV ? clock64() : 0;
        long long g_t1 = 0, g_t2 = 0, g_t3 = 0, g_ta = 0, g_tb = 0, g_info = 0;
        if (M.abort) return;
        PROF(1)  // exchange wait (compute side)

        // ---- column groups that held a candidate at step t-1: winners (nodes.py:1097-1105), Ae trace, partner
        // Ai neurons, monitors, late STDP — all of it local to the group's own threads
        const bool lateg = !isx && t > 0 && (((M.candgrp[ppar] >> cg) & 1u) || (Q.dbg & 256));   // (diagnostic bit 256: every group walks the late path every step)
        if (lateg) {
            if (PROFV && b == 0) M.rs[cg][0] = clock64();
            if (pend) {
                uint32_t sE = 0;
                if (E.one_spike) {
                    const unsigned long long key = keyT[ppar * Bp + b];
                    const int wj = (int)(uint32_t)(key & 0xffffffffull) - jc;
                    if (key != 0ull && wj >= 0 && wj < 4 && ((candE >> wj) & 1u)) sE = 1u << wj;
                } else sE = candE;
                if (PROFV) M.rs[cg][1] = clock64();
                if (E.traces) {
                    #pragma unroll
                    for (int c = 0; c < 4; ++c) xE[c] = trace_step(xE[c], (sE >> c) & 1u, E.trace_decay, E.trace_scale, 0);
                }
                if (update_on) {
                    if (xE[0] != 0.0f || xE[1] != 0.0f || xE[2] != 0.0f || xE[3] != 0.0f) {
                        *(float4 *)(tx + b * TJ + 4 * cg) = make_float4(xE[0] * C.nu0, xE[1] * C.nu0, xE[2] * C.nu0, xE[3] * C.nu0);
                        if (!livep) {
                            livep = true;
                            atomicOr(&M.nz4[cg][b >> 5], 1u << (b & 31));
                            live[cg * Bp + atomicAdd(&M.nlive[cg], 1)] = (uint8_t)b;
                        }
                    }
                }
                if (PROFV) M.rs[cg][2] = clock64();
                const int slot = (stage_on && candstamp[ppar * Bp + b] == (uint32_t)t) ? candslot[ppar * Bp + b] : -1;   // staged at step t-1
                if (PROFV) M.rs[cg][3] = clock64();
                #pragma unroll 1
                for (int c = 0; c < 4; ++c)
                    if ((candE >> c) & 1u) {
                        const int col = 4 * cg + c;
                        const bool won = (sE >> c) & 1u;
                        if (won && update_on && post_on) {
                            atomicOr(&M.wmask[col][b >> 5], 1u << (b & 31));
                            atomicOr(&M.colwin[cg], 1u << c);
                            const int k = atomicAdd(&M.nwl[cg], 1);
                            if (k < XR) M.wl[cg][k] = ((uint32_t)col << 16) | ((uint32_t)b << 8) | (slot >= 0 ? (uint32_t)slot : 0xffu);
                        }
                        // the partner Ai neuron of every candidate runs here (claimed at step t-1): input `exc`
                        // through the diagonal Ae->Ai iff the candidate won (network.py:225-248)
                        if (t < T) ai_step2(&s_cx, (int)ai_map[b * TJ + col], won ? (0.0f + Q.exc) : 0.0f, t, sis_t, aispk + par * Bp);
                        if (won) {   // monitors (monitors.py:94-111): the launch code cleared the raster
                            if (E.rec_s) E.rec_s[((size_t)(t - 1) * B + b) * n + jc + c] = 1;
                            if (E.rec_count) atomicAdd(E.rec_count + (size_t)b * n + jc + c, 1);
                        }
                    }
                if (t == T) sEfin = sE;
                pend = 0;
                if (PROFV) M.rs[cg][4] = clock64();
            }
            PROF(2)  // winners
            if (PROFV) g_ta = clock64();
            if (update_on) {
                // STDP of step t-1 for this column group (MCC_learning.py:234-299)
                bar_group(gbar, Bp);
                if (PROFV && b == 0) M.rs[cg][5] = clock64();
                const int sb_ = buf;   // slot t = spikes of step t-1
                const uint32_t gwin = post_on ? M.colwin[cg] : 0u;
                const int nwl = post_on ? M.nwl[cg] : 0;
                const bool deferg = M.defer[cg] != 0;   // set by the early pass of step t-1 (a __syncthreads ago), uniform in the group
                bool fast = nwl <= XR && nwl == __popc(gwin) && !M.denseflag[sb_];
                if (fast) {
                    #pragma unroll 1
                    for (int k = 0; k < nwl; ++k) if ((M.wl[cg][k] & 0xffu) == 0xffu) fast = false;
                }
                if (PROFV && b == 0) M.rs[cg][6] = clock64();
                const uint8_t *xr = xrow + ppar * XR * P;   // rows staged at step t-1 (buffer of its parity, filled for the ((t-1)>>1)-th time)
                if (nwl && stage_on) { while (!mbar_try_wait(&M.mbar_x[ppar], (uint32_t)((t - 1) >> 1) & 1u)) {} }
                PROF(3)  // late set-up
                if (PROFV) { g_tb = clock64(); g_info = (fast ? 2 : 0) | (nwl << 4) | ((long long)M.ncand[ppar] << 12) | ((long long)M.nlive[cg] << 20); }
                // the shadow pass finished every row of this group except those at which a candidate-holding sample
                // spiked (unless the slot is dense: then nothing of this group was done yet)
                const uint32_t *dm = &M.candmask[ppar][cg][0];
                if (fast) {
                    if (PROFV && b == 0) M.rs[cg][7] = clock64();
                    if (pre_on || gwin) stdp_list2<CG, BW>(&s_cx, sb_, cg, 1, gwin, dm, ppar, xr, t - 1, b, Bp);
                    if (PROFV && b == 0) M.rs[cg][8] = clock64();
                    if (nwl) post_rows2<CG, BW>(&s_cx, sb_, cg, nwl, dm, xr, t - 1, M.negzero ? 0 : 1, b, Bp);
                } else {
                    stdp_rows2<CG, BW>(&s_cx, sb_, cg, gwin, M.candb[ppar], min(M.ncand[ppar], XR),
                                       Q.xage ? Q.xage + (size_t)(t - 1) * B * P : nullptr, xr, t - 1, M.denseflag[sb_] ? nullptr : dm, b, Bp);
                }
                PROF(4)  // late pass
                if (PROFV && b == 0) M.rs[cg][9] = clock64();
                if (deferg) {   // rows the early pass left unclamped (rare): both terms are in now — the reference's single clamp
                    bar_group(gbar, Bp);
                    #pragma unroll 1
                    for (int i = b; i < P; i += Bp) {
                        float4 w4 = *(float4 *)(W + i * WS + 4 * cg);
                        w4.x = clampf(w4.x, C.wmin, C.wmax); w4.y = clampf(w4.y, C.wmin, C.wmax);
                        w4.z = clampf(w4.z, C.wmin, C.wmax); w4.w = clampf(w4.w, C.wmin, C.wmax);
                        *(float4 *)(W + i * WS + 4 * cg) = w4;
                    }
                }
                bar_group(gbar, Bp);   // the group's weights are final for step t-1
                if (deferg && b == 0) M.defer[cg] = 0;
                if (PROFV && b == 0 && Q.prof && ((M.candgrp[ppar] >> cg) & 1u)) {   // fine stamps -> sums over all late paths of the window
                    M.rs[cg][10] = clock64();
                    long long *fs = Q.prof + 160 * NPROF + 32 * 160 * 2 + 32 * 160 * 8 * 5;
                    for (int k = 1; k <= 10; ++k) {
                        const long long d = M.rs[cg][k] - M.rs[cg][k - 1];
                        if (d > 0 && d < 1000000) atomicAdd((unsigned long long *)fs + k, (unsigned long long)d);
                    }
                    atomicAdd((unsigned long long *)fs, 1ull);
                }
                if (gwin) {
                    if (b < 4 * BW) M.wmask[4 * cg + (b / BW)][b % BW] = 0;
                    if (b == 0) { M.colwin[cg] = 0; M.nwl[cg] = 0; }
                }
            }
        }
        if (t == 1 && update_on && C.has_clamp) {
            // the reference clamps the whole matrix every step (MCC_learning.py:101-110, learning.py:97-104);
            // after the first step that is a no-op for untouched weights, so one sweep after step 0 covers it
            __syncthreads();
            #pragma unroll 1
            for (int idx = tid; idx < P * TJ; idx += blockDim.x) { const int i = idx / TJ, jj = idx - i * TJ; W[i * WS + jj] = clampf(W[i * WS + jj], C.wmin, C.wmax); }
            __syncthreads();
        }
        if (t == T) break;

        // ---- step t: gather (unless done ahead), Ae update, candidates, Ai list ------------------------
        if (PROFV) g_t1 = clock64();
        uint32_t cand = 0;
        if (act) {
            if (!havepre) {
                while (!mbar_try_wait(&M.mbar_in[buf], (uint32_t)(t >> 1) & 1u)) {}   // slot t landed (prefetched one step ago)
                pre = gather2<WS>(&s_cx, cE, t, b, Wc);
            }
            if (PROFV) g_t2 = clock64();
            const float p[4] = {pre.x, pre.y, pre.z, pre.w};
            const float4 th4 = *(const float4 *)(thr_s + 4 * cg);
            const float thr[4] = {th4.x, th4.y, th4.z, th4.w};
            const int isum = (int)isumT[ppar * Bp + b];                 // Ai spikes of step t-1 (lateral inhibition)
            const uint32_t own = (aispk[ppar * Bp + b] >> (4 * cg)) & 0xFu;  // ... of which my neurons' own partners
            #pragma unroll
            for (int c = 0; c < 4; ++c) {
                if ((vm >> c) & 1u) {
                    // network.py:225-248: X->Ae first, then Ai->Ae; the latter is rep[#spiking Ai other than j]
                    const int mI = isum - (int)((own >> c) & 1u);
                    float cur = 0.0f + p[c];
                    float inh = rep[min(mI, Q.nrep)];
                    if (mI > Q.nrep) {   // more simultaneous Ai spikes than the table holds (never with one_spike)
                        #pragma unroll 1
                        for (int m = Q.nrep; m < mI; ++m) inh = inh + Q.inh_neg;
                    }
                    cur = cur + inh;
                    // DiehlAndCookNodes.forward up to the threshold test (nodes.py:1077-1092)
                    vE[c] = E.decay * (vE[c] - E.rest) + E.rest;
                    const float gate = rE[c] <= 0.0f ? 1.0f : 0.0f;
                    vE[c] = vE[c] + gate * cur;
                    rE[c] = rE[c] - E.dt;
                    if (vE[c] >= thr[c]) { cand |= 1u << c; rE[c] = E.refrac; vE[c] = E.reset; }
                }
            }
            if (cand) {
                on_candidate2(&s_cx, cand, b, cg, t);
            } else if (E.traces) {
                // no candidate among my neurons: their step-t trace is already final (no spike)
                #pragma unroll
                for (int c = 0; c < 4; ++c) xE[c] = xE[c] * E.trace_decay;
                if (update_on && livep) *(float4 *)(tx + b * TJ + 4 * cg) = make_float4(xE[0] * C.nu0, xE[1] * C.nu0, xE[2] * C.nu0, xE[3] * C.nu0);
            }
        }
        if (PROFV) g_t3 = clock64();
        candE = cand;
        pend = cand;
        havepre = false;
        // Ai list (LIFNodes.forward, nodes.py:500-529): the entries nobody claimed for this step get no input
        if (!isx) {
            const int nact = M.nact_snap;
            #pragma unroll 1
            for (int k = tid; k < nact; k += NC) {
                if (ai_claim[par * aicap + k] == (uint16_t)t) continue;
#if V2_UNIFY_AI
                ai_step2(&s_cx, k, (t == 0 && (ai_fl[k] & 4u)) ? (0.0f + Q.exc) : 0.0f, t, sis_t, aispk + par * Bp);
#else
                ai_step_body(&s_cx, k, (t == 0 && (ai_fl[k] & 4u)) ? (0.0f + Q.exc) : 0.0f, t, sis_t, aispk + par * Bp);
#endif
            }
        }
        PROF(6)  // gather + neurons
        if (PROFV && Q.prof && !isx && b == 0 && t >= 100 && t < 132) {
            long long *tr = Q.prof + 160 * NPROF + 32 * 160 * 2 + (((t - 100) * 160 + blockIdx.x) * 8 + cg) * 5;
            const long long now_ = clock64();
            tr[0] = g_t1 - g_t0; tr[1] = g_t2 - g_t1; tr[2] = g_t3 - g_t2; tr[3] = now_ - g_t3;
            tr[4] = (lateg ? 1 : 0) | g_info | ((g_ta - g_t0) << 32) | ((g_tb - g_ta) << 48);
        }
        __syncthreads();
        PROF(7)  // step sync
        if (PROFV && tid == NC) {   // exchange warp: S1 -> S2 = the CTA's critical path of this step
            const long long now_ = clock64();
            pc[14] += now_ - pt;
            if (Q.prof && t >= 100 && t < 132) Q.prof[160 * NPROF + ((t - 100) * 160 + blockIdx.x) * 2] = now_ - pt;
            pt = now_;
        }
        // ---- arrive(t): this CTA's contributions to step t's exchange are issued -----------------
        const int nb = buf ^ 1;                                      // slot t+1 = spikes of step t
        if (isx) {
            // clear the exchange slot step t+1 will accumulate into (last read before the previous barrier)
            if (blockIdx.x == 0)
                for (int k = lane; k < B; k += 32) { Q.win[((t + 1) % 3) * B + k] = 0ull; Q.sisum[((t + 1) % 3) * B + k] = 0u; }
            __syncwarp();
            if (lane == 0) {
#ifdef SNN_EMU
                __atomic_fetch_add(Q.bar, 1u, __ATOMIC_RELEASE);
#else
                if (PROFV && (Q.dbg & 128)) asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(Q.bar) : "memory");
                else asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(Q.bar) : "memory");
#endif
                mbar_arrive(&M.mbar_x[par]);   // the rows staged in step t: complete once the copies issued above have landed
            }
            if (PROFV && tid == tid_pf && t + 2 <= T && (Q.dbg & 1)) mbar_arrive(&M.mbar_in[buf]);
            if (tid == tid_pf && t + 2 <= T && !(PROFV && (Q.dbg & 1))) {  // prefetch slot t+2 into the buffer the gather just finished with
                const bool masks = !(PROFV && (Q.dbg & 32));   // (diagnostic: pixel masks not refreshed)
                mbar_arrive_expect_tx(&M.mbar_in[buf], bytesE + (masks ? bytesT : 0u));
                bulk_g2s(evb + buf * evblk, Q.evS + (size_t)(t + 2) * Q.SB, bytesE, &M.mbar_in[buf]);
                if (masks) bulk_g2s(inT + buf * P * BW, Q.inT + (size_t)(t + 2) * P * BW, bytesT, &M.mbar_in[buf]);
            }
            if (PROFV && tid == NC) pt = clock64();
        } else {
            // ---- the shadow of the exchange, per column group: theta, early STDP of step t, gather of step t+1
            if (b < 4) {
                // theta = theta * decay + theta_plus * (#candidates of the column)  (nodes.py:1078-1094)
                const int col = 4 * cg + b;
                float th = theta_s[col];
                if (E.learning) th = th * E.theta_decay + E.theta_plus * (float)M.cnt[par][col];
                theta_s[col] = th;
                thr_s[col] = E.thresh + (E.learning ? th * E.theta_decay : th);
                M.cnt[par][col] = 0;
            }
            if (cg == 0) aispk[ppar * Bp + b] = 0u;   // own Ai spikes of step t-1: consumed
            const bool earlyg = !((M.candgrp[par] >> cg) & 1u);
            while (!mbar_try_wait(&M.mbar_in[nb], (uint32_t)((t + 1) >> 1) & 1u)) {}
            if (tid == 0) { M.denseflag[nb] = dense_nb; M.ncand[ppar] = 0; M.candgrp[ppar] = 0; M.nact_snap = M.nact; }   // parity ppar: next used by step t+1
            // early STDP: the pre term of step t wherever it does not depend on the exchange — every row of a column
            // group without a candidate; in a group with candidates every row except those at which a candidate-holding
            // sample spiked (stdp_late2 finishes them once the winners are known)
            if (b < BW) M.candmask[ppar][cg][b] = 0;   // step t-1's: consumed by the late pass above
            if (b == 0) M.ncs[ppar][cg] = 0;
            if (update_on && pre_on && !(PROFV && (Q.dbg & 2))) {
#if V2_UNIFY_LIST
                if (!dense_nb) stdp_list2<CG, BW>(&s_cx, nb, cg, 0, 0u, &M.candmask[par][cg][0], par, xrow, t, b, Bp);
#else
                if (!dense_nb) stdp_list_body<CG, BW>(&s_cx, nb, cg, 0, 0u, &M.candmask[par][cg][0], par, xrow, t, b, Bp);
#endif
                else if (earlyg) stdp_rows2<CG, BW>(&s_cx, nb, cg, 0u, M.candb[par], 0, nullptr, xrow, t, nullptr, b, Bp);
                if (earlyg) bar_group(gbar, Bp);
            }
            PROF(8)  // early STDP
            // gather of step t+1 for a group whose weights are final now: off the critical path
            const bool ahead = t + 1 < T && earlyg && !(t == 0 && update_on && C.has_clamp) && !(PROFV && (Q.dbg & 4));
            if (ahead && act) pre = gather2<WS>(&s_cx, evb + nb * evblk, t + 1, b, Wc);
            havepre = ahead;
            PROF(9)  // gather ahead
        }
    }

    // ---- epilogue: normalize() on the tile (network.py:464-465), write everything back -----
    __syncthreads();
    PROF(12)
    if (Q.normalize && C.has_norm) {
        float *part = (float *)xrow;  // [SNN_NORM_CHUNKS + 1][TJ]
        const int chunk = (P + SNN_NORM_CHUNKS - 1) / SNN_NORM_CHUNKS;
        #pragma unroll 1
        for (int idx = tid; idx < SNN_NORM_CHUNKS * TJ; idx += blockDim.x) {
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
        #pragma unroll 4
        for (int idx = tid; idx < P * TJ; idx += blockDim.x) {
            const int i = idx / TJ, jj = idx % TJ;
            const float w = W[i * WS + jj] * part[SNN_NORM_CHUNKS * TJ + jj];
            if (j0 + jj < n) C.w[(size_t)i * n + j0 + jj] = w;
        }
    } else {
        #pragma unroll 4
        for (int idx = tid; idx < P * TJ; idx += blockDim.x) {
            const int i = idx / TJ, jj = idx % TJ;
            if (j0 + jj < n) C.w[(size_t)i * n + j0 + jj] = W[i * WS + jj];
        }
    }
    for (int jj = tid; jj < TJ; jj += blockDim.x)
        if (j0 + jj < n) E.theta[j0 + jj] = theta_s[jj];
    if (act) {
        const size_t k0 = (size_t)b * n + jc;
        float vI[4], rI[4];
        uint32_t sI = 0;
        // Ai: listed neurons from the list, the others only ran their refractory counter down
        #pragma unroll
        for (int c = 0; c < 4; ++c) {
            vI[c] = 0.0f; rI[c] = 0.0f;
            if ((vm >> c) & 1u) {
                const unsigned e = ai_map[b * TJ + 4 * cg + c];
                const float r0 = I.refrac_count[k0 + c];
                if (e == AI_NONE) { vI[c] = I.rest; rI[c] = refrac_replay(r0, I.dt, T); }
                else {
                    vI[c] = ai_v[e];
                    rI[c] = (ai_fl[e] & 1u) ? refrac_replay(r0, I.dt, T) : ai_rc[e];
                    sI |= ((ai_fl[e] >> 1) & 1u) << c;
                }
            }
        }
        if (vm == 0xFu && (n & 3) == 0) {
            *(float4 *)(E.v + k0) = make_float4(vE[0], vE[1], vE[2], vE[3]);
            *(float4 *)(E.refrac_count + k0) = make_float4(rE[0], rE[1], rE[2], rE[3]);
            if (E.traces) *(float4 *)(E.x + k0) = make_float4(xE[0], xE[1], xE[2], xE[3]);
            *(uint32_t *)(E.s + k0) = (sEfin & 1u) | ((sEfin & 2u) << 7) | ((sEfin & 4u) << 14) | ((sEfin & 8u) << 21);
            *(float4 *)(I.v + k0) = make_float4(vI[0], vI[1], vI[2], vI[3]);
            *(float4 *)(I.refrac_count + k0) = make_float4(rI[0], rI[1], rI[2], rI[3]);
            *(uint32_t *)(I.s + k0) = (sI & 1u) | ((sI & 2u) << 7) | ((sI & 4u) << 14) | ((sI & 8u) << 21);
        } else {
            #pragma unroll
            for (int c = 0; c < 4; ++c)
                if ((vm >> c) & 1u) {
                    const size_t k = k0 + c;
                    E.v[k] = vE[c]; E.refrac_count[k] = rE[c];
                    if (E.traces) E.x[k] = xE[c];
                    E.s[k] = (sEfin >> c) & 1u;
                    I.v[k] = vI[c]; I.refrac_count[k] = rI[c];
                    I.s[k] = (sI >> c) & 1u;
                }
        }
    }
    PROF(13)  // epilogue
    if (PROFV && Q.prof && tid == 0)
        for (int k = 0; k < NPROF; ++k) Q.prof[blockIdx.x * NPROF + k] = pc[k];
}

// ---------------------------------------------------------------------------------------
// Pre-pass 1 (bits): grid = (slot, group of 32 samples).  Slot 0 = the Input layer's incoming spike
// state, slot t+1 = the external input of step t (network.py:388-392 / Input.forward nodes.py:211-221).
// Produces the per-sample bit rows and ascending pixel lists, the per-pixel sample masks, the Input
// monitor raster, flags non-binary input; one CTA also counts the incoming Ai spikes and builds the
// inhibition table.
__global__ void __launch_bounds__(256) snn_dc2_prepass(const __grid_constant__ F2Params Q, int BW) {
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
                    if (pos < EV_CAP) elist[ev_pos(B, b, pos)] = (uint16_t)(c * 16 + bit);
                    ++pos;
                }
                total += __shfl_sync(0xffffffffu, incl, 31);
            }
            for (int w = (P >> 5) + lane; w < SW; w += 32) {  // words past the last pixel (odd chunk count: half word above)
                if (w * 32 >= P) { sbits[bl * SW + w] = 0u; Q.inS[((size_t)slot * B + b) * SW + w] = 0u; }
            }
            const int padded = (total + 3) & ~3;  // pad to a multiple of 4 with the zero row P
            if (total < EV_CAP && lane < padded - total) elist[ev_pos(B, b, total + lane)] = (uint16_t)P;
            if (lane == 0) ecnt[b] = (uint16_t)(total > 65535 ? 65535 : total);
            dense |= total > EV_CAP;
        } else if (b < B) {
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
                    if (pos < EV_CAP) elist[ev_pos(B, b, pos)] = (uint16_t)i;
                }
                total += __popc(word);
            }
            const int padded = (total + 3) & ~3;  // pad to a multiple of 4 with the zero row P
            if (total < EV_CAP && lane < padded - total) elist[ev_pos(B, b, total + lane)] = (uint16_t)P;
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
        for (int b = warp; b < B; b += nwarp) {  // Ai spikes of step -1
            int c = 0;
            for (int j = lane; j < Q.n; j += 32) c += Q.I.s[(size_t)b * Q.n + j] != 0;
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0) Q.sisum0[b] = (unsigned int)c;
        }
    }
    if (slot == (gridDim.x > 1 ? 1 : 0) && grp == 0 && threadIdx.x == 0) {
        // rep[m] = m-fold sequential sum of the Ai->Ae weight: what the reference's dense sum over k of
        // sI[b,k] * w_ie[k,j] evaluates to when m inhibitory neurons (other than j) spike (topology.py:437-479)
        float a = 0.0f;
        Q.rep[0] = 0.0f;
        for (int m = 1; m <= Q.nrep; ++m) { a = a + Q.inh_neg; Q.rep[m] = a; }
    }
}

// Pre-pass 2 (trace scan): the Input layer's trace for every step of the window, Nodes.forward (nodes.py:96-103)
// applied T times per pixel.  What the window kernel needs of it — the traces of a winner's sample for the STDP
// post term — is stored as one AGE byte per (step, sample, pixel): steps since the pixel's last spike in this
// window (255: none yet), a quarter of the fp32 size so that the whole window stays in L2; the trace is a table
// lookup by age.  Thread = 4 pixels of one sample; loads run SCAN_U steps ahead.  Also leaves the layer's final
// state (x, s) behind and keeps the incoming traces (for pixels that have not spiked yet).
constexpr int SCAN_U = 10;
__global__ void __launch_bounds__(256) snn_dc2_trace_scan(const __grid_constant__ F2Params Q) {
    const int P4 = Q.P >> 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Q.B * P4) return;
    const snn_layer_t &X = Q.X;
    const size_t stride4 = (size_t)Q.B * P4;  // float4 / uchar4 units per timestep
    float4 x = X.traces ? ((const float4 *)X.x)[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (Q.x0c) {
        ((float4 *)Q.x0c)[idx] = x;
        if (x.x != 0.0f || x.y != 0.0f || x.z != 0.0f || x.w != 0.0f) *Q.anyx0 = 1;
    }
    uint32_t age = 0xffffffffu;   // 4 age bytes
    uint32_t last = 0;
    for (int t0 = 0; t0 < Q.T; t0 += SCAN_U) {
        uint32_t bits[SCAN_U];
        #pragma unroll
        for (int u = 0; u < SCAN_U; ++u) {
            bits[u] = 0;
            if (t0 + u < Q.T && X.ext) {
                if (X.ext_dtype == SNN_EXT_U8) {
                    const uint32_t w = __ldg((const uint32_t *)X.ext + (size_t)(t0 + u) * stride4 + idx);
                    bits[u] = __vcmpne4(w, 0u) & 0x01010101u;
                } else {
                    const float4 f = __ldg((const float4 *)X.ext + (size_t)(t0 + u) * stride4 + idx);
                    bits[u] = (f.x != 0.0f ? 1u : 0u) | (f.y != 0.0f ? 0x100u : 0u) | (f.z != 0.0f ? 0x10000u : 0u) | (f.w != 0.0f ? 0x1000000u : 0u);
                }
            }
        }
        #pragma unroll
        for (int u = 0; u < SCAN_U; ++u) {
            if (t0 + u < Q.T) {
                const uint32_t w = bits[u];
                if (X.traces) {
                    x.x = trace_step(x.x, w & 0x1u, X.trace_decay, X.trace_scale, X.traces_additive);
                    x.y = trace_step(x.y, w & 0x100u, X.trace_decay, X.trace_scale, X.traces_additive);
                    x.z = trace_step(x.z, w & 0x10000u, X.trace_decay, X.trace_scale, X.traces_additive);
                    x.w = trace_step(x.w, w & 0x1000000u, X.trace_decay, X.trace_scale, X.traces_additive);
                    // ages: +1 per step (saturating below 255 = "none yet" is not needed: T <= 254), 0 on a spike
                    const uint32_t never = __vcmpeq4(age, 0xffffffffu);            // bytes still 255
                    age = (__vadd4(age, 0x01010101u) & ~never) | never;            // 255 stays 255
                    age &= ~(w * 0xffu);                                            // spike: byte -> 0
                    if (Q.xage) ((uint32_t *)Q.xage)[(size_t)(t0 + u) * stride4 + idx] = age;
                }
                last = w;
            }
        }
    }
    if (Q.T > 0) {
        if (X.traces) ((float4 *)X.x)[idx] = x;
        ((uint32_t *)X.s)[idx] = last;
    }
}

struct Match2 {
    int lX, lE, lI, cXE, cEI, cIE, CG, BW, threads, grid, SW, SB, Bp, nrep;
    size_t smem;
};

int device_sms2() {
    static int sms = -1;
#ifdef SNN_EMU
    if (sms < 0) { const char *v = getenv("SNN_EMU_FUSED_SMS"); sms = v && atoi(v) > 0 ? atoi(v) : 148; }
#endif
    if (sms < 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
        else { sms = 148; (void)cudaGetLastError(); }
    }
    return sms;
}

bool match2(const snn_net_t *net, const snn_run_opts_t *o, Match2 &m) {
    if (net->n_layers != 3 || net->n_conns != 3 || o->T < 1 || o->T > 254 || o->one_step) return false;   // T: one age byte per step
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
    if (X.rec_count || X.rec_v || E.rec_v || I.rec_v) return false;
    // the lean option set (everything else: snn_fused_dc.cu)
    if (X.traces_additive || E.traces_additive || E.has_lbound || I.has_lbound) return false;
    if (!(I.rest < I.thresh)) return false;   // an Ai neuron at rest must stay silent
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
    if (CX.rule == SNN_RULE_WDEP_POSTPRE || CX.reduction != SNN_REDUCE_SUM) return false;
    if (CX.weight_decay != 0.0f && CX.weight_decay != 1.0f) return false;
    if (CX.rule == SNN_RULE_NOOP) return false;
    if (CX.rule >= SNN_RULE_POSTPRE && (!X.traces || !E.traces)) return false;
    const int n = E.n, P = X.n, B = o->B;
    if (B > 128 || (P & 15) || P >= 65535 || n >= 65535) return false;   // P: 16-byte rows of age bytes for the bulk copies
    const int sms = device_sms2();
    m.SW = ((P + 31) / 32 + 3) / 4 * 4;
    m.BW = 4;
    m.SB = ev_block_bytes(B);
    m.Bp = (B + 31) / 32 * 32;
    m.nrep = n < 1023 ? n : 1023;
    for (int CG = 1; CG <= 8; ++CG) {
        const int TJ = 4 * CG;
        const int grid = (n + TJ - 1) / TJ;
        const int threads = m.Bp * CG + 32;   // compute threads + the exchange warp
        if (grid > sms || threads > 1024) continue;
        const SmemLayout2 SL = smem_layout2(P, TJ, B, m.Bp, m.BW, m.nrep);
        if (SL.total > 227 * 1024) continue;
        m.CG = CG; m.grid = grid; m.threads = threads; m.smem = SL.total;
        return true;
    }
    return false;
}

template <int CG, int BW, int VAR>
cudaError_t launch_var2(const F2Params &Q, const Match2 &m, cudaStream_t stream) {
#ifdef SNN_EMU
    (void)stream;
    emu::run_grid(m.grid, m.threads, m.smem, [](void *a) { snn_dc2_window<CG, BW, VAR>(*(const F2Params *)a); }, (void *)&Q);
    return cudaSuccess;
#endif
    cudaError_t e = cudaFuncSetAttribute(snn_dc2_window<CG, BW, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m.smem);
    if (e != cudaSuccess) return e;
    void *args[] = {(void *)&Q};
    return cudaLaunchCooperativeKernel((void *)snn_dc2_window<CG, BW, VAR>, dim3(m.grid), dim3(m.threads), args, m.smem, stream);
}

template <int CG>
cudaError_t launch_cg2(const F2Params &Q, const Match2 &m, cudaStream_t stream) {
    return Q.prof ? launch_var2<CG, 4, 2>(Q, m, stream) : launch_var2<CG, 4, 0>(Q, m, stream);
}

struct WsLayout2 { size_t dense, bar, win, sisum, anyx0, inS, inT, evS, rep, sisum0, x0c, xage, prof, total; };
WsLayout2 ws_layout2(const Match2 &m, int T, int B, int P, bool traces) {
    WsLayout2 L; size_t o = 0;
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    // dense flags, barrier word and exchange slots are adjacent: one memset node per window
    L.dense = o; o += al(sizeof(int) * (size_t)(T + 1));
    L.bar = o; o += al(sizeof(unsigned int) * 64);
    L.win = o; o += al(sizeof(unsigned long long) * 3 * (size_t)B);
    L.sisum = o; o += al(sizeof(unsigned int) * 3 * (size_t)B);
    L.anyx0 = o; o += al(sizeof(int) * 4);
    L.inS = o; o += al(sizeof(uint32_t) * (size_t)(T + 1) * B * m.SW);
    L.inT = o; o += al(sizeof(uint32_t) * (size_t)(T + 1) * P * m.BW);
    L.evS = o; o += al((size_t)(T + 1) * m.SB);
    L.rep = o; o += al(sizeof(float) * (size_t)(m.nrep + 1));
    L.sisum0 = o; o += al(sizeof(unsigned int) * (size_t)B);
    L.x0c = o; o += traces ? al(sizeof(float) * (size_t)B * P) : 0;
    L.xage = o; o += traces ? al((size_t)T * B * P) : 0;
    L.prof = o; o += al(sizeof(long long) * (160 * NPROF + 32 * 160 * 2 + 32 * 160 * 8 * 5 + 16));
    L.total = o;
    return L;
}

}  // namespace

int snn_fused_dc2_supported(const snn_net_t *net, const snn_run_opts_t *opts) {
    Match2 m;
    return match2(net, opts, m) ? 1 : 0;
}

size_t snn_fused_dc2_workspace_bytes(const snn_net_t *net, const snn_run_opts_t *opts) {
    Match2 m;
    if (!match2(net, opts, m)) return 0;
    const snn_layer_t &X = net->layers[m.lX];
    return ws_layout2(m, opts->T, opts->B, X.n, X.traces != 0).total;
}

int snn_fused_dc2_launch(const snn_net_t *net, const snn_run_opts_t *opts, void *ws_, size_t ws_bytes, cudaStream_t stream,
                         int *launches) {
    Match2 m;
    if (!match2(net, opts, m)) return SNN_ERR_UNSUPPORTED;
    const int T = opts->T, B = opts->B, P = net->layers[m.lX].n;
    const bool traces = net->layers[m.lX].traces != 0;
    const WsLayout2 WL = ws_layout2(m, T, B, P, traces);
    if (ws_bytes < WL.total) return SNN_ERR_WORKSPACE;
    char *ws = (char *)ws_;
    F2Params Q;
    memset(&Q, 0, sizeof(Q));
    Q.X = net->layers[m.lX]; Q.E = net->layers[m.lE]; Q.I = net->layers[m.lI];
    Q.C = net->conns[m.cXE];
    Q.exc = net->conns[m.cEI].structure_val; Q.inh_neg = net->conns[m.cIE].structure_val;
    Q.T = T; Q.B = B; Q.Bp = m.Bp; Q.P = P; Q.n = Q.E.n; Q.learning = net->learning; Q.normalize = opts->normalize;
    Q.G = m.grid; Q.nrep = m.nrep;
    {
        const SmemLayout2 SL = smem_layout2(P, 4 * m.CG, B, m.Bp, m.BW, m.nrep);
        Q.o_W = (uint32_t)SL.W; Q.o_tx = (uint32_t)SL.tx; Q.o_ev = (uint32_t)SL.ev; Q.o_inT = (uint32_t)SL.inT; Q.o_xrow = (uint32_t)SL.xrow;
        Q.o_rep = (uint32_t)SL.rep; Q.o_theta = (uint32_t)SL.theta; Q.o_live = (uint32_t)SL.live; Q.o_tab = (uint32_t)SL.tab;
        Q.o_ai = (uint32_t)SL.ai; Q.o_misc = (uint32_t)SL.misc;
    }
    Q.SW = m.SW; Q.SB = m.SB; Q.liE = m.lE; Q.seed = opts->seed; Q.step_offset = opts->step_offset;
    Q.inS = (uint32_t *)(ws + WL.inS); Q.inT = (uint32_t *)(ws + WL.inT); Q.evS = (unsigned char *)(ws + WL.evS);
    Q.dense = (int *)(ws + WL.dense); Q.bar = (unsigned int *)(ws + WL.bar); Q.win = (unsigned long long *)(ws + WL.win);
    Q.sisum = (unsigned int *)(ws + WL.sisum);
    Q.rep = (float *)(ws + WL.rep); Q.sisum0 = (unsigned int *)(ws + WL.sisum0);
    const bool stdp = Q.C.rule >= SNN_RULE_POSTPRE;
    const bool need_x = traces && stdp && net->learning && Q.C.nu1 != 0.0f;
    Q.xage = need_x ? (uint8_t *)(ws + WL.xage) : nullptr;
    Q.x0c = need_x ? (float *)(ws + WL.x0c) : nullptr;
    Q.anyx0 = (int *)(ws + WL.anyx0);
    Q.err = opts->err_flag;
    { const char *d = getenv("SNN_B200_DEBUG"); Q.dbg = d ? atoi(d) : 0; }
    const bool prof = getenv("SNN_B200_PROF") != nullptr;
    Q.prof = prof ? (long long *)(ws + WL.prof) : nullptr;
    int nl = 0;
    // dense flags, barrier counter, exchange slots
    if (cudaMemsetAsync(ws + WL.dense, 0, WL.inS - WL.dense, stream) != cudaSuccess) return SNN_ERR_CUDA;
    // sparse monitors: the window kernel only writes the ones
    if (Q.prof) cudaMemsetAsync(Q.prof + 160 * NPROF + 32 * 160 * 2 + 32 * 160 * 8 * 5, 0, sizeof(long long) * 16, stream);
    if (Q.E.rec_s && cudaMemsetAsync(Q.E.rec_s, 0, (size_t)T * B * Q.n, stream) != cudaSuccess) return SNN_ERR_CUDA;
    if (Q.I.rec_s && cudaMemsetAsync(Q.I.rec_s, 0, (size_t)T * B * Q.n, stream) != cudaSuccess) return SNN_ERR_CUDA;
    // the two static matrices are replaced by their constants: make sure they still have that structure
    if (snn_verify_structure(net->conns[m.cEI], Q.n, Q.err, stream) != SNN_OK || snn_verify_structure(net->conns[m.cIE], Q.n, Q.err, stream) != SNN_OK) return SNN_ERR_CUDA;
    nl += 2;
    SNN_LAUNCH(snn_dc2_prepass, dim3(T + 1, (B + 31) / 32), 256, sizeof(uint32_t) * 32 * (size_t)m.SW, stream, Q, m.BW);
    ++nl;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) {
        const int items = B * (P >> 2);
        SNN_LAUNCH(snn_dc2_trace_scan, (items + 255) / 256, 256, 0, stream, Q);
        ++nl;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) {
        switch (m.CG) {
            case 1: e = launch_cg2<1>(Q, m, stream); break;
            case 2: e = launch_cg2<2>(Q, m, stream); break;
            case 3: e = launch_cg2<3>(Q, m, stream); break;
            case 4: e = launch_cg2<4>(Q, m, stream); break;
            case 5: e = launch_cg2<5>(Q, m, stream); break;
            case 6: e = launch_cg2<6>(Q, m, stream); break;
            case 7: e = launch_cg2<7>(Q, m, stream); break;
            default: e = launch_cg2<8>(Q, m, stream); break;
        }
        ++nl;
    }
    if (e != cudaSuccess) {
        fprintf(stderr, "libsnn_b200: fused DC2015 (v2) window launch failed: %s\n", cudaGetErrorString(e));
        return SNN_ERR_CUDA;
    }
    if (prof) {  // debug only: synchronise and print the per-phase cycle counts (min / mean / max over CTAs)
        static const char *names[NPROF] = {"prologue", "exchange wait", "winners", "late set-up", "late pass", "slot wait+sync", "gather+neurons",
                                           "step sync", "early STDP", "gather ahead", "x: barrier wait", "x: exchange read", "final sync", "epilogue", "x: S1->S2 (CTA)", ""};
        cudaStreamSynchronize(stream);
        static long long hostp[160 * NPROF];
        cudaMemcpy(hostp, Q.prof, sizeof(long long) * (size_t)m.grid * NPROF, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[snn_b200 prof v2] grid=%d threads=%d T=%d (cycles per timestep, thread 0 of each CTA: min / mean / max)\n", m.grid, m.threads, T);
        for (int k = 0; k < 15; ++k) {
            double sum = 0, mx = 0, mn = 1e300;
            for (int g = 0; g < m.grid; ++g) { const double v = (double)hostp[g * NPROF + k]; sum += v; mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
            const double div = (k == 0 || k == 12 || k == 13) ? 1.0 : (double)T;
            fprintf(stderr, "  %-18s %10.0f %10.0f %10.0f\n", names[k], mn / div, sum / m.grid / div, mx / div);
        }
        if (T >= 133) {   // per-step view, steps 100..131: which CTA makes the others wait, and for how long
            static long long tr[32 * 160 * 2];
            cudaMemcpy(tr, Q.prof + 160 * NPROF, sizeof(tr), cudaMemcpyDeviceToHost);
            double a_mean = 0, a_max = 0, w_mean = 0, w_min = 0;
            for (int st = 0; st < 32; ++st) {
                double sm = 0, mx = 0, wm = 0, wn = 1e300;
                for (int g = 0; g < m.grid; ++g) {
                    const double a = (double)tr[(st * 160 + g) * 2], w = (double)tr[(st * 160 + g) * 2 + 1];
                    sm += a / m.grid; mx = a > mx ? a : mx; wm += w / m.grid; wn = w < wn ? w : wn;
                }
                a_mean += sm / 32; a_max += mx / 32; w_mean += wm / 32; w_min += wn / 32;
            }
            fprintf(stderr, "  per step (t=100..131): S1->S2 mean over CTAs %.0f, slowest CTA %.0f; barrier wait mean %.0f, of the last arriver %.0f\n",
                    a_mean, a_max, w_mean, w_min);
            // the slowest column group of the slowest CTA of each step: where its time went
            static long long gt[32 * 160 * 8 * 5];
            cudaMemcpy(gt, Q.prof + 160 * NPROF + 32 * 160 * 2, sizeof(gt), cudaMemcpyDeviceToHost);
            double ph[4] = {0, 0, 0, 0}, late_share = 0;
            for (int st = 0; st < 32; ++st) {
                double best = -1; int bg = 0, bc = 0;
                for (int g = 0; g < m.grid; ++g)
                    for (int c = 0; c < m.CG; ++c) {
                        const long long *r = gt + ((st * 160 + g) * 8 + c) * 5;
                        const double tot = (double)(r[0] + r[1] + r[2] + r[3]);
                        if (tot > best) { best = tot; bg = g; bc = c; }
                    }
                const long long *r = gt + ((st * 160 + bg) * 8 + bc) * 5;
                for (int k = 0; k < 4; ++k) ph[k] += (double)r[k] / 32;
                late_share += (double)(r[4] & 1) / 32;
                fprintf(stderr, "    step %d: CTA %d group %d: late path %lld (winners %lld, set-up %lld), gather %lld, neurons %lld, Ai %lld; fast=%lld nwl=%lld ncand=%lld nlive=%lld\n",
                        100 + st, bg, bc, r[0], (r[4] >> 32) & 0xffff, (r[4] >> 48) & 0xffff, r[1], r[2], r[3], (r[4] >> 1) & 1, (r[4] >> 4) & 0xff,
                        (r[4] >> 12) & 0xff, (r[4] >> 20) & 0xfff);
            }
            fprintf(stderr, "  slowest group per step: winners+late STDP %.0f, gather %.0f, neurons %.0f, Ai list %.0f cycles; it was a late group in %.0f%% of the steps\n",
                    ph[0], ph[1], ph[2], ph[3], 100 * late_share);
            // does a late path cost more the longer its CTA has not run one (instruction-cache eviction)?
            double dur[4] = {0, 0, 0, 0}; int cntg[4] = {0, 0, 0, 0};
            for (int g = 0; g < m.grid; ++g) {
                int last = -1;
                for (int st = 0; st < 32; ++st) {
                    bool any = false; double worst = 0;
                    for (int c = 0; c < m.CG; ++c) {
                        const long long *r = gt + ((st * 160 + g) * 8 + c) * 5;
                        if (r[4] & 1) { any = true; worst = (double)r[0] > worst ? (double)r[0] : worst; }
                    }
                    if (any) {
                        if (last >= 0) { const int gap = st - last; const int bk = gap <= 2 ? 0 : gap <= 5 ? 1 : gap <= 10 ? 2 : 3; dur[bk] += worst; ++cntg[bk]; }
                        last = st;
                    }
                }
            }
            {
                long long fs[16];
                cudaMemcpy(fs, Q.prof + 160 * NPROF + 32 * 160 * 2 + 32 * 160 * 8 * 5, sizeof(fs), cudaMemcpyDeviceToHost);
                const double nn = fs[0] ? (double)fs[0] : 1.0;
                fprintf(stderr, "  late path, mean over %lld group events: S1->key %.0f, traces/tx %.0f, slot %.0f, winners loop+Ai %.0f, group barrier %.0f, fast check %.0f, "
                                "row wait %.0f, list pass %.0f, row pass %.0f, closing barrier %.0f\n", fs[0], fs[1] / nn, fs[2] / nn, fs[3] / nn, fs[4] / nn, fs[5] / nn,
                        fs[6] / nn, fs[7] / nn, fs[8] / nn, fs[9] / nn, fs[10] / nn);
            }
            fprintf(stderr, "  late-path cycles by steps since the CTA's previous late path: <=2: %.0f (%d)  3-5: %.0f (%d)  6-10: %.0f (%d)  >10: %.0f (%d)\n",
                    cntg[0] ? dur[0] / cntg[0] : 0.0, cntg[0], cntg[1] ? dur[1] / cntg[1] : 0.0, cntg[1], cntg[2] ? dur[2] / cntg[2] : 0.0, cntg[2],
                    cntg[3] ? dur[3] / cntg[3] : 0.0, cntg[3]);
        }
    }
    if (launches) *launches = nl;  // 2 structure checks + two pre-passes + the persistent window kernel
    return SNN_OK;
}
