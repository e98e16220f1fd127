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
// The step's input spikes arrive as two bit matrices (per sample over pixels for the gather,
// per pixel over samples for the STDP pre term), produced once per window by a pre-pass and
// staged a step ahead into shared memory with cp.async.bulk (TMA bulk copy) + mbarrier.
//
// Loop iteration t = [finalise step t-1: exchange results, winner, Ae trace, STDP on the tile]
//                    [step t: gather, Ae/Ai update, candidates -> atomics] [grid barrier].
//
// Arithmetic and summation orders are those of snn_phases.cuh / oracle/snn_oracle.c, so the
// result is bit-identical to the generic kernel and to the oracle.
#include <cstdio>
#include <cstring>

#include "snn_common.cuh"

namespace {

constexpr int XR = 8;  // input-trace rows staged per CTA per step (samples with a candidate)

struct FusedParams {
    snn_layer_t X, E, I;      // Input, DiehlAndCookNodes (Ae), LIFNodes (Ai)
    snn_conn_t C;             // X -> Ae
    float exc, inh_neg;       // diag value of Ae->Ai, off-diag value of Ai->Ae
    int32_t T, B, P, n, learning, normalize;
    int32_t SW, BW;           // words per sample row of inS / per pixel row of inT (multiples of 4)
    int32_t liE;              // index of Ae in the user's layer list (enters the tie-break hash)
    uint32_t seed, step_offset;
    uint32_t *inS;            // [T+1][B][SW]  slot t = spikes of step t-1: bit i of sample b
    uint32_t *inT;            // [T+1][P][BW]  same spikes: bit b of pixel i
    unsigned long long *win;  // [3][B] arg-max keys, slot t % 3
    unsigned int *sisum;      // [3][B] Ai spike counts, slot t % 3 (slot 2 = step -1)
    float *xpub;              // [2][B][P] published input traces, slot t & 1
    unsigned int *bar;
    int32_t *err;
};

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

struct Misc {  // small per-step scratch (lives in shared memory)
    uint64_t mbar[2];
    int cnt[32];            // candidates per column (theta update)
    uint32_t wmask[32][8];  // winners: per column, bit mask over samples
    uint32_t nz[8];         // samples that have a non-zero Ae trace in this tile
    int ncand;              // samples with a candidate in this tile this step
    int candb[XR];          // ... the first XR of them (their input-trace rows get staged)
    int winany;
    uint32_t colwin;        // bit j: column j has a winner this step
    int8_t wslot[256];      // sample -> staged row slot, -1 = not staged
};

struct SmemLayout { size_t W, tx, inS, inT, xrow, rep, xown, theta, misc, total; };

__host__ __device__ inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline SmemLayout smem_layout(int P, int TJ, int B, int SW, int BW, int n, int own) {
    SmemLayout L;
    size_t o = 0;
    L.W = o; o += al16(sizeof(float) * (size_t)P * TJ);
    L.tx = o; o += al16(sizeof(float) * (size_t)B * TJ);
    L.inS = o; o += al16(sizeof(uint32_t) * 2 * (size_t)B * SW);
    L.inT = o; o += al16(sizeof(uint32_t) * 2 * (size_t)P * BW);
    L.xrow = o; o += al16(sizeof(float) * (size_t)XR * P);
    L.rep = o; o += al16(sizeof(float) * (size_t)(n + 1));
    L.xown = o; o += al16(sizeof(float) * (size_t)own * P);
    L.theta = o; o += al16(sizeof(float) * 32);
    L.misc = o; o += al16(sizeof(Misc));
    L.total = o;
    return L;
}

template <int TJ>
__global__ void __launch_bounds__(1024, 1) snn_dc_fused_window(const __grid_constant__ FusedParams Q) {
    constexpr int CG = TJ / 4;  // float4 column groups = lanes that share one sample
    extern __shared__ __align__(16) unsigned char smem[];
    const int B = Q.B, P = Q.P, n = Q.n, SW = Q.SW, BW = Q.BW, T = Q.T;
    const unsigned int G = gridDim.x;
    const int own = (B + (int)G - 1) / (int)G;
    const SmemLayout SL = smem_layout(P, TJ, B, SW, BW, n, own);
    float *W = (float *)(smem + SL.W);
    float *tx = (float *)(smem + SL.tx);
    uint32_t *inS = (uint32_t *)(smem + SL.inS);
    uint32_t *inT = (uint32_t *)(smem + SL.inT);
    float *xrow = (float *)(smem + SL.xrow);
    float *rep = (float *)(smem + SL.rep);
    float *xown = (float *)(smem + SL.xown);
    float *theta_s = (float *)(smem + SL.theta);
    Misc &M = *(Misc *)(smem + SL.misc);

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int b = tid / CG, cg = tid % CG;   // state ownership: sample b, neurons jc..jc+3
    const bool act = b < B;
    const int j0 = blockIdx.x * TJ;
    const int jc = j0 + 4 * cg;
    const int NRS = nthr / CG, rslot = tid / CG;  // row slots of the STDP row loop
    const snn_layer_t &E = Q.E, &I = Q.I, &X = Q.X;
    const snn_conn_t &C = Q.C;
    const bool stdp = C.rule >= SNN_RULE_POSTPRE;
    const bool wdep = C.rule == SNN_RULE_WDEP_POSTPRE;
    const bool pre_on = stdp && C.nu0 != 0.0f, post_on = stdp && C.nu1 != 0.0f;
    const bool decay_on = C.weight_decay != 0.0f && C.weight_decay != 1.0f;
    const bool update_on = Q.learning && C.rule != SNN_RULE_NONE && (stdp || decay_on);
    const bool stage_on = update_on && post_on && X.traces;
    const float Bf = (float)B;

    // ---- prologue: W tile, theta, inhibition table, owned input traces, state registers ----
    for (int idx = tid; idx < P * TJ; idx += nthr) {
        const int i = idx / TJ, jj = idx % TJ;
        W[idx] = (j0 + jj < n) ? C.w[(size_t)i * n + j0 + jj] : 0.0f;
    }
    for (int jj = tid; jj < TJ; jj += nthr) theta_s[jj] = (j0 + jj < n) ? E.theta[j0 + jj] : 0.0f;
    if (tid == 0) {
        // rep[m] = m-fold sequential sum of the Ai->Ae weight: what the reference's dense sum
        // over k of sI[b,k] * w_ie[k,j] evaluates to when m inhibitory neurons (other than j) spike
        float a = 0.0f;
        rep[0] = 0.0f;
        for (int m = 1; m <= n; ++m) { a = a + Q.inh_neg; rep[m] = a; }
        mbar_init(&M.mbar[0], 1);
        mbar_init(&M.mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        M.ncand = 0; M.winany = 0; M.colwin = 0;
    }
    for (int k = tid; k < 8; k += nthr) M.nz[k] = 0;
    for (int k = tid; k < 32; k += nthr) M.cnt[k] = 0;
    for (int k = tid; k < 32 * 8; k += nthr) (&M.wmask[0][0])[k] = 0;
    for (int k = tid; k < 256; k += nthr) M.wslot[k] = -1;
    if (X.traces)
        for (int o = 0; o < own; ++o) {
            const int bo = blockIdx.x + o * (int)G;
            if (bo < B)
                for (int i = tid; i < P; i += nthr) xown[o * P + i] = X.x[(size_t)bo * P + i];
        }

    float vE[4], rE[4], xE[4], vI[4], rI[4];
    uint32_t sEprev = 0, sIprev = 0, candE = 0;  // 4-bit masks over my neurons
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
        *(float4 *)(tx + (size_t)b * TJ + 4 * cg) =
            make_float4(wdep ? xE[0] : xE[0] * C.nu0, wdep ? xE[1] : xE[1] * C.nu0, wdep ? xE[2] : xE[2] * C.nu0,
                        wdep ? xE[3] : xE[3] * C.nu0);
    }
    __syncthreads();
    if (act && stdp && (xE[0] != 0.0f || xE[1] != 0.0f || xE[2] != 0.0f || xE[3] != 0.0f)) atomicOr(&M.nz[b >> 5], 1u << (b & 31));

    const uint32_t bytesS = (uint32_t)(sizeof(uint32_t) * (size_t)B * SW), bytesT = (uint32_t)(sizeof(uint32_t) * (size_t)P * BW);
    if (tid == 0) {  // stage slot 0 (spikes of step -1 = the Input layer's incoming spike state)
        mbar_expect_tx(&M.mbar[0], bytesS + bytesT);
        bulk_g2s(inS, Q.inS, bytesS, &M.mbar[0]);
        bulk_g2s(inT, Q.inT, bytesT, &M.mbar[0]);
    }
    uint32_t ph0 = 0, ph1 = 0;
    __syncthreads();

    // =====================================================================================
    for (int t = 0; t <= T; ++t) {
        const int buf = t & 1;
        const uint32_t *cS = inS + (size_t)buf * B * SW;  // spikes of step t-1, per sample
        const uint32_t *cT = inT + (size_t)buf * P * BW;  // spikes of step t-1, per pixel

        // prefetch slot t+1 into the other buffer (its readers finished before the last barrier)
        if (tid == 0 && t + 1 <= T) {
            const int nb = buf ^ 1;
            mbar_expect_tx(&M.mbar[nb], bytesS + bytesT);
            bulk_g2s(inS + (size_t)nb * B * SW, Q.inS + (size_t)(t + 1) * B * SW, bytesS, &M.mbar[nb]);
            bulk_g2s(inT + (size_t)nb * P * BW, Q.inT + (size_t)(t + 1) * P * BW, bytesT, &M.mbar[nb]);
        }

        // ---- A. exchange results of step t-1 (slot (t-1) % 3; for t = 0 the pre-pass filled it)
        const int xs = (t + 2) % 3;
        unsigned long long key = 0ull;
        unsigned int isum = 0;
        if (act) {
            isum = __ldcg(Q.sisum + (size_t)xs * B + b);
            if (t > 0) key = __ldcg(Q.win + (size_t)xs * B + b);
        }
        if (t > 0 && stage_on) {  // speculative: input-trace rows of this tile's candidate samples
            const int ns = min(M.ncand, XR);
            for (int idx = tid; idx < ns * P; idx += nthr) {
                const int r = idx / P, i = idx - r * P;
                xrow[idx] = __ldcg(Q.xpub + ((size_t)((t - 1) & 1) * B + M.candb[r]) * P + i);
            }
        }
        // wait for this iteration's spike matrices (prefetched during the previous iteration)
        {
            uint32_t &ph = buf ? ph1 : ph0;
            while (!mbar_try_wait(&M.mbar[buf], ph)) {}
            ph ^= 1u;
        }

        if (t > 0) {
            // ---- B. finalise step t-1: winner (nodes.py:1097-1105), Ae trace (nodes.py:96-103)
            uint32_t sE = 0;
            if (act) {
                if (E.one_spike) {
                    if (candE && key != 0ull) {
                        const int wj = (int)(uint32_t)(key & 0xffffffffull) - jc;
                        if (wj >= 0 && wj < 4 && ((candE >> wj) & 1u)) sE = 1u << wj;
                    }
                } else sE = candE;
                if (E.traces) {
                    #pragma unroll
                    for (int c = 0; c < 4; ++c) xE[c] = trace_step(xE[c], (sE >> c) & 1u, E.trace_decay, E.trace_scale, E.traces_additive);
                }
                if (update_on && stdp) {
                    if (xE[0] != 0.0f || xE[1] != 0.0f || xE[2] != 0.0f || xE[3] != 0.0f)
                        *(float4 *)(tx + (size_t)b * TJ + 4 * cg) =
                            make_float4(wdep ? xE[0] : xE[0] * C.nu0, wdep ? xE[1] : xE[1] * C.nu0,
                                        wdep ? xE[2] : xE[2] * C.nu0, wdep ? xE[3] : xE[3] * C.nu0);
                    if (sE) {
                        atomicOr(&M.nz[b >> 5], 1u << (b & 31));
                        #pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if ((sE >> c) & 1u) {
                                atomicOr(&M.wmask[4 * cg + c][b >> 5], 1u << (b & 31));
                                atomicOr(&M.colwin, 1u << (4 * cg + c));
                            }
                        M.winany = 1;
                    }
                }
                // monitors (monitors.py:94-111): spikes / voltages of step t-1
                if (E.rec_s || I.rec_s || E.rec_v || I.rec_v) {
                    #pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (jc + c < n) {
                            const size_t k = ((size_t)(t - 1) * B + b) * n + jc + c;
                            if (E.rec_s) E.rec_s[k] = (sE >> c) & 1u;
                            if (I.rec_s) I.rec_s[k] = (sIprev >> c) & 1u;
                            if (E.rec_v) E.rec_v[k] = vE[c];
                            if (I.rec_v) I.rec_v[k] = vI[c];
                        }
                }
            }
            sEprev = sE;
        }
        __syncthreads();

        if (t > 0 && update_on) {
            // ---- C. learning-rule update of step t-1 on the W tile --------------------------
            //   U[i,j] = reduce_b sX[b,i] * (xE[b,j]*nu0)   pre term  (MCC_learning.py:234-263)
            //   V[i,j] = reduce_b xX[b,i] * (sE[b,j]*nu1)   post term (MCC_learning.py:267-299)
            // then decay + clamp (MCC_learning.py:86-110).  Rows without a pre spike from a
            // sample with a live Ae trace, in columns without a winner, are bitwise unchanged
            // and skipped — except on the first update of the window (entries may sit outside
            // [wmin, wmax] after normalize()) or with a weight decay.
            const bool full = decay_on || (C.has_clamp && t == 1);
            uint32_t nzm[8];
            #pragma unroll
            for (int g = 0; g < 8; ++g) nzm[g] = g < BW ? M.nz[g] : 0u;
            const uint32_t mycolwin = post_on ? ((M.colwin >> (4 * cg)) & 0xFu) : 0u;
            for (int i = rslot; i < P; i += NRS) {
                uint32_t m[8];
                uint32_t anym = 0;
                #pragma unroll
                for (int g = 0; g < 8; ++g) {
                    m[g] = (pre_on && g < BW) ? (cT[(size_t)i * BW + g] & nzm[g]) : 0u;
                    anym |= m[g];
                }
                const bool pre_t = anym != 0u;
                if (!(full || pre_t || mycolwin)) continue;
                const float4 w4 = *(const float4 *)(W + (size_t)i * TJ + 4 * cg);
                float U[4] = {0.f, 0.f, 0.f, 0.f};
                if (pre_t) {
                    #pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        uint32_t mm = m[g];
                        while (mm) {
                            const int bb = g * 32 + __ffs(mm) - 1;
                            mm &= mm - 1;
                            const float4 t4 = *(const float4 *)(tx + (size_t)bb * TJ + 4 * cg);
                            U[0] = U[0] + t4.x; U[1] = U[1] + t4.y; U[2] = U[2] + t4.z; U[3] = U[3] + t4.w;
                        }
                    }
                    if (C.reduction == SNN_REDUCE_MEAN) { U[0] = U[0] / Bf; U[1] = U[1] / Bf; U[2] = U[2] / Bf; U[3] = U[3] / Bf; }
                }
                float wv[4] = {w4.x, w4.y, w4.z, w4.w};
                #pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float V = 0.0f;
                    const bool post_t = (mycolwin >> c) & 1u;
                    if (post_t) {
                        for (int g = 0; g < BW; ++g) {
                            uint32_t mm = M.wmask[4 * cg + c][g];
                            while (mm) {
                                const int bb = g * 32 + __ffs(mm) - 1;
                                mm &= mm - 1;
                                const int s = M.wslot[bb];
                                const float xsv = (s >= 0) ? xrow[(size_t)s * P + i]
                                                           : __ldcg(Q.xpub + ((size_t)((t - 1) & 1) * B + bb) * P + i);
                                V = V + xsv * (wdep ? 1.0f : C.nu1);
                            }
                        }
                        if (C.reduction == SNN_REDUCE_MEAN) V = V / Bf;
                    }
                    wv[c] = apply_rule(C, wv[c], U[c], pre_t, V, post_t);
                }
                *(float4 *)(W + (size_t)i * TJ + 4 * cg) = make_float4(wv[0], wv[1], wv[2], wv[3]);
            }
            __syncthreads();
            if (M.winany) {
                for (int k = tid; k < TJ * 8; k += nthr) (&M.wmask[0][0])[k] = 0;
            }
        }
        if (t > 0) {
            // reset the per-step candidate / winner bookkeeping
            for (int k = tid; k < B; k += nthr) M.wslot[k] = -1;
            __syncthreads();
            if (tid == 0) { M.winany = 0; M.colwin = 0; M.ncand = 0; }
        }
        if (t == T) break;

        // ---- D. step t: theta decay, gather, Ae / Ai update, candidates ---------------------
        if (tid < TJ && E.learning) theta_s[tid] = theta_s[tid] * E.theta_decay;  // nodes.py:1078-1079
        __syncthreads();
        uint32_t cand = 0, sI = 0;
        unsigned long long mykey = 0ull;
        int nI = 0;
        if (act) {
            // spike-gather: p[c] = sum_{i in sX(t-1)[b]} W[i][c], i ascending (topology.py:437-479)
            float p[4] = {0.f, 0.f, 0.f, 0.f};
            const uint32_t *row = cS + (size_t)b * SW;
            for (int w4i = 0; w4i < SW; w4i += 4) {
                const uint4 q = *(const uint4 *)(row + w4i);
                const uint32_t ww[4] = {q.x, q.y, q.z, q.w};
                #pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint32_t word = ww[u];
                    while (word) {
                        const int i = (w4i + u) * 32 + __ffs(word) - 1;
                        word &= word - 1;
                        const float4 r4 = *(const float4 *)(W + (size_t)i * TJ + 4 * cg);
                        p[0] = p[0] + r4.x; p[1] = p[1] + r4.y; p[2] = p[2] + r4.z; p[3] = p[3] + r4.w;
                    }
                }
            }
            const float4 th4 = *(const float4 *)(theta_s + 4 * cg);
            const float th[4] = {th4.x, th4.y, th4.z, th4.w};
            #pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (jc + c < n) {
                    // network.py:225-248: X->Ae first, then Ai->Ae; the latter is rep[#spiking Ai other than j]
                    const int mI = (int)isum - (int)((sIprev >> c) & 1u);
                    float cur = 0.0f + p[c];
                    cur = cur + rep[mI];
                    if (dc_step(E, vE[c], rE[c], cur, th[c])) cand |= 1u << c;       // nodes.py:1077-1092
                    if (E.has_lbound && vE[c] < E.lbound) vE[c] = E.lbound;           // nodes.py:1108-1109
                    float curI = ((sEprev >> c) & 1u) ? (0.0f + Q.exc) : 0.0f;        // diagonal Ae->Ai
                    if (lif_step(I, vI[c], rI[c], curI)) { sI |= 1u << c; ++nI; }     // nodes.py:500-529
                }
            }
            if (cand && E.one_spike) {
                #pragma unroll
                for (int c = 0; c < 4; ++c)
                    if ((cand >> c) & 1u) {
                        const unsigned long long k2 = snn_one_spike_key(Q.seed, (uint32_t)t + Q.step_offset, (uint32_t)Q.liE,
                                                                        (uint32_t)b, (uint32_t)(jc + c));
                        mykey = k2 > mykey ? k2 : mykey;
                    }
            }
            if (cand) {
                #pragma unroll
                for (int c = 0; c < 4; ++c)
                    if ((cand >> c) & 1u) atomicAdd(&M.cnt[4 * cg + c], 1);
            }
        }
        // reductions over the CG lanes that share a sample (all lanes of the warp take part)
        uint32_t anyc = cand;
        #pragma unroll
        for (int o = CG / 2; o > 0; o >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, mykey, o);
            mykey = ok > mykey ? ok : mykey;
            nI += __shfl_xor_sync(0xffffffffu, nI, o);
            anyc |= __shfl_xor_sync(0xffffffffu, anyc, o);
        }
        if (act && cg == 0) {
            const int ws = t % 3;
            if (mykey) atomicMax(Q.win + (size_t)ws * B + b, mykey);
            if (nI) atomicAdd(Q.sisum + (size_t)ws * B + b, (unsigned int)nI);
            if (anyc && stage_on) {
                const int s = atomicAdd(&M.ncand, 1);
                if (s < XR) { M.candb[s] = b; M.wslot[b] = (int8_t)s; }
            }
        }
        candE = cand;
        sIprev = sI;

        // input trace of the samples this CTA owns: x = s ? scale : x * decay (nodes.py:96-103),
        // published for the winners' post-synaptic STDP term of THIS step
        if (X.traces) {
            for (int o = 0; o < own; ++o) {
                const int bo = blockIdx.x + o * (int)G;
                if (bo < B) {
                    const uint32_t *srow = Q.inS + ((size_t)(t + 1) * B + bo) * SW;
                    float *dst = Q.xpub + ((size_t)(t & 1) * B + bo) * P;
                    for (int i = tid; i < P; i += nthr) {
                        const bool s = (__ldg(srow + (i >> 5)) >> (i & 31)) & 1u;
                        const float x = trace_step(xown[o * P + i], s, X.trace_decay, X.trace_scale, X.traces_additive);
                        xown[o * P + i] = x;
                        dst[i] = x;
                    }
                }
            }
        }
        // clear the exchange slot step t+1 will accumulate into (last read before the previous barrier)
        if (blockIdx.x == 0)
            for (int k = tid; k < B; k += nthr) {
                Q.win[(size_t)((t + 1) % 3) * B + k] = 0ull;
                Q.sisum[(size_t)((t + 1) % 3) * B + k] = 0u;
            }
        __syncthreads();
        // theta += theta_plus * (number of candidates in the column)  (nodes.py:1093-1094)
        if (tid < TJ) {
            if (E.learning) theta_s[tid] = theta_s[tid] + E.theta_plus * (float)M.cnt[tid];
            M.cnt[tid] = 0;
        }
        if (!grid_barrier(Q.bar, G, Q.err)) return;
    }

    // ---- epilogue: normalize() on the tile (network.py:464-465), write everything back -----
    __syncthreads();
    if (Q.normalize && C.has_norm) {
        float *part = xrow;  // [SNN_NORM_CHUNKS + 1][TJ]
        const int chunk = (P + SNN_NORM_CHUNKS - 1) / SNN_NORM_CHUNKS;
        for (int idx = tid; idx < SNN_NORM_CHUNKS * TJ; idx += nthr) {
            const int c = idx / TJ, jj = idx % TJ;
            float a = 0.0f;
            const int i1 = min((c + 1) * chunk, P);
            for (int i = c * chunk; i < i1; ++i) { const float x = W[(size_t)i * TJ + jj]; a = a + (C.norm_abs ? fabsf(x) : x); }
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
        for (int idx = tid; idx < P * TJ; idx += nthr) W[idx] = W[idx] * part[SNN_NORM_CHUNKS * TJ + (idx % TJ)];
        __syncthreads();
    }
    for (int idx = tid; idx < P * TJ; idx += nthr) {
        const int i = idx / TJ, jj = idx % TJ;
        if (j0 + jj < n) C.w[(size_t)i * n + j0 + jj] = W[idx];
    }
    for (int jj = tid; jj < TJ; jj += nthr)
        if (j0 + jj < n) E.theta[j0 + jj] = theta_s[jj];
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
            }
    }
    for (int o = 0; o < own; ++o) {
        const int bo = blockIdx.x + o * (int)G;
        if (bo < B) {
            const uint32_t *srow = Q.inS + ((size_t)T * B + bo) * SW;  // slot T = spikes of step T-1
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
// per-sample and per-pixel bit matrices, the Input monitor raster, flags non-binary input;
// CTA 0 also resets the exchange slots and counts the incoming Ai spikes.
__global__ void __launch_bounds__(256) snn_dc_prepass(const __grid_constant__ FusedParams Q) {
    extern __shared__ uint32_t sbits[];  // [B][SW]
    const int B = Q.B, P = Q.P, SW = Q.SW, BW = Q.BW, PW = (P + 31) / 32;
    const int slot = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const snn_layer_t &X = Q.X;
    bool nonbin = false;
    for (int b = warp; b < B; b += nwarp) {
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
            if (lane == 0) { sbits[b * SW + w] = word; Q.inS[((size_t)slot * B + b) * SW + w] = word; }
        }
    }
    __syncthreads();
    // transpose 32x32 bit blocks: inT[pixel][g] bit b' = inS[g*32+b'][pixel/32] bit pixel%32
    const int NG = (B + 31) / 32;
    for (int blk = warp; blk < BW * PW; blk += nwarp) {
        const int g = blk / PW, w = blk % PW;
        uint32_t mine = 0;
        if (g < NG) {
            const int bb = g * 32 + lane;
            const uint32_t word = bb < B ? sbits[bb * SW + w] : 0u;
            #pragma unroll
            for (int r = 0; r < 32; ++r) {
                const uint32_t mm = __ballot_sync(0xffffffffu, (word >> r) & 1u);
                if (lane == r) mine = mm;
            }
        }
        const int i = w * 32 + lane;
        if (i < P) Q.inT[((size_t)slot * P + i) * BW + g] = mine;
    }
    if (nonbin && Q.err) atomicOr(Q.err, SNN_ERR_NONBINARY);
    if (slot == 0) {
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
    int lX, lE, lI, cXE, cEI, cIE, TJ, threads, grid, SW, BW, own;
    size_t smem;
};

int device_sms() {
    static int sms = -1;
    if (sms < 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
        else { sms = 148; (void)cudaGetLastError(); }
    }
    return sms;
}

bool match(const snn_net_t *net, const snn_run_opts_t *o, Match &m) {
    if (net->n_layers != 3 || net->n_conns != 3 || o->T < 1) return false;
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
    if (X.rec_v) return false;
    m.cXE = m.cEI = m.cIE = -1;
    for (int c = 0; c < 3; ++c) {
        const snn_conn_t &C = net->conns[c];
        if (C.b) return false;
        if (C.src == m.lX && C.tgt == m.lE) m.cXE = c;
        else if (C.src == m.lE && C.tgt == m.lI) m.cEI = c;
        else if (C.src == m.lI && C.tgt == m.lE) m.cIE = c;
    }
    if (m.cXE < 0 || m.cEI < 0 || m.cIE < 0 || m.cXE > m.cIE) return false;  // accumulation order into Ae
    const snn_conn_t &CX = net->conns[m.cXE], &CEI = net->conns[m.cEI], &CIE = net->conns[m.cIE];
    auto is_static = [](const snn_conn_t &C) { return (C.rule == SNN_RULE_NONE || (C.rule == SNN_RULE_NOOP && (C.weight_decay == 1.0f || C.weight_decay == 0.0f))) && !C.has_norm; };
    if (!is_static(CEI) || !is_static(CIE)) return false;
    if (CEI.structure != SNN_W_DIAG || CIE.structure != SNN_W_OFFDIAG) return false;
    if (CX.rule >= SNN_RULE_POSTPRE && (!X.traces || !E.traces)) return false;
    const int n = E.n, P = X.n, B = o->B;
    if (B > 256) return false;
    const int sms = device_sms();
    m.SW = ((P + 31) / 32 + 3) / 4 * 4;
    m.BW = ((B + 31) / 32 + 3) / 4 * 4;
    for (int TJ : {4, 8, 16, 32}) {
        const int grid = (n + TJ - 1) / TJ;
        const int threads = ((B * (TJ / 4)) + 31) / 32 * 32;
        if (grid > sms || threads > 1024) continue;
        const int own = (B + grid - 1) / grid;
        const SmemLayout SL = smem_layout(P, TJ, B, m.SW, m.BW, n, own);
        if (SL.total > 227 * 1024) continue;
        m.TJ = TJ; m.grid = grid; m.threads = threads < 32 ? 32 : threads; m.own = own; m.smem = SL.total;
        return true;
    }
    return false;
}

template <int TJ>
cudaError_t launch_tj(const FusedParams &Q, const Match &m, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(snn_dc_fused_window<TJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m.smem);
    if (e != cudaSuccess) return e;
    void *args[] = {(void *)&Q};
    return cudaLaunchCooperativeKernel((void *)snn_dc_fused_window<TJ>, dim3(m.grid), dim3(m.threads), args, m.smem, stream);
}

struct WsLayout { size_t bar, inS, inT, win, sisum, xpub, total; };
WsLayout ws_layout(const Match &m, int T, int B, int P) {
    WsLayout L; size_t o = 0;
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    L.bar = o; o += al(sizeof(unsigned int) * 96);
    L.inS = o; o += al(sizeof(uint32_t) * (size_t)(T + 1) * B * m.SW);
    L.inT = o; o += al(sizeof(uint32_t) * (size_t)(T + 1) * P * m.BW);
    L.win = o; o += al(sizeof(unsigned long long) * 3 * B);
    L.sisum = o; o += al(sizeof(unsigned int) * 3 * B);
    L.xpub = o; o += al(sizeof(float) * 2 * (size_t)B * P);
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
    Q.SW = m.SW; Q.BW = m.BW; Q.liE = m.lE; Q.seed = opts->seed; Q.step_offset = opts->step_offset;
    Q.inS = (uint32_t *)(ws + WL.inS); Q.inT = (uint32_t *)(ws + WL.inT);
    Q.win = (unsigned long long *)(ws + WL.win); Q.sisum = (unsigned int *)(ws + WL.sisum);
    Q.xpub = (float *)(ws + WL.xpub); Q.bar = (unsigned int *)(ws + WL.bar); Q.err = opts->err_flag;
    if (cudaMemsetAsync(Q.bar, 0, sizeof(unsigned int) * 96, stream) != cudaSuccess) return SNN_ERR_CUDA;
    snn_dc_prepass<<<T + 1, 256, sizeof(uint32_t) * (size_t)B * m.SW, stream>>>(Q);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) {
        switch (m.TJ) {
            case 4: e = launch_tj<4>(Q, m, stream); break;
            case 8: e = launch_tj<8>(Q, m, stream); break;
            case 16: e = launch_tj<16>(Q, m, stream); break;
            default: e = launch_tj<32>(Q, m, stream); break;
        }
    }
    if (e != cudaSuccess) {
        fprintf(stderr, "libsnn_b200: fused DC2015 window launch failed: %s\n", cudaGetErrorString(e));
        return SNN_ERR_CUDA;
    }
    if (launches) *launches = 2;  // pre-pass + persistent window kernel
    return SNN_OK;
}
