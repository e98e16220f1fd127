// snn_phases.cuh — the per-step building blocks of the window kernels (spike gather, neuron
// update, one_spike resolution, STDP tile update, column normalisation), shared by the generic
// window kernel (snn_generic.cu) and the single-operator entry points (snn_ops.cu).
#pragma once
#include "snn_common.cuh"

namespace {

// ---------------------------------------------------------------------------------------
// Shared memory of one CTA of the generic tier (dynamic; carved the same way by the window kernel and by
// the single-operator kernels).  Phases never overlap inside a CTA, so the regions are reused:
//   acc     phase 3: per-warp [32 rows][32 columns] pre-synaptic accumulators (32 KB)
//           phase 1: the per-warp spike lists of the gather (first 16 KB) and the staged source bit rows of a
//                    convolutional gather (second 16 KB)
//   xs      phase 3: per-warp staged pre-synaptic traces of the samples with a post-synaptic event (16 KB)
//           phase 1: staged filter taps of a convolutional gather
//           phase 3 (conv MSTDP): the sample's two bit rows, its decoded source-spike list, the channel offsets
//   xt      phase 3: the target traces of the whole tile [B][32]; phase 3 (dense MSTDP): staged rule state
//   acc     phase 3 (conv MSTDP) also: the P- rows of the unit's output channels
#define SNN_P3_MAXEV 16
#define SNN_GATHER_BLOCK 1024   // source neurons per gather block (32 words): list capacity per warp
#define SNN_CONV_STAGE_WORDS 4096   // staged source bit words of a conv gather (second half of acc)
#define SNN_CONV_STAGE_TAPS 4096    // staged filter taps (xs region)
#define SNN_XT_MAX_BYTES (96 * 1024)

struct GenSmem {
    float *acc;
    uint16_t *list;     // [WARPS][SNN_GATHER_BLOCK]  (aliases acc)
    uint32_t *cbits;    // [SNN_CONV_STAGE_WORDS]     (aliases acc + 16 KB)
    float *red;         // [SNN_NORM_CHUNKS + 1][32]
    uint32_t *colmask;  // [ceil(B/32)][32]
    float *xs;          // [WARPS][SNN_P3_MAXEV][32]
    int32_t *evb;       // [SNN_P3_MAXEV + 1]
    uint8_t *evslot;    // [B]
    float *xt;          // [B][32] or NULL (batch too large to stage)
    uint32_t *live;     // [ceil(B/32)] samples whose staged trace row is not all zero
    int32_t xt_bytes;
};

__host__ __device__ inline size_t gen_xt_bytes(int B) {
    const size_t b = sizeof(float) * 32 * (size_t)B;
    return b <= SNN_XT_MAX_BYTES ? b : 0;
}
__host__ __device__ inline size_t gen_smem_bytes(int B) {
    const size_t NG = (size_t)((B + 31) / 32);
    size_t s = sizeof(float) * SNN_GEN_WARPS * 32 * 32;            // acc
    s += sizeof(float) * (SNN_NORM_CHUNKS + 1) * 32;               // red
    s += sizeof(uint32_t) * 32 * NG;                               // colmask
    s += sizeof(float) * SNN_GEN_WARPS * SNN_P3_MAXEV * 32;        // xs
    s += sizeof(int32_t) * (SNN_P3_MAXEV + 16);                    // evb (padded)
    s += ((size_t)B + 15) / 16 * 16;                               // evslot
    s += (sizeof(uint32_t) * NG + 15) / 16 * 16;                   // live
    s += gen_xt_bytes(B);
    return s;
}
__device__ __forceinline__ GenSmem gen_carve(float *smem, int B) {
    GenSmem M;
    const int NG = (B + 31) / 32;
    M.acc = smem;
    M.list = (uint16_t *)smem;
    M.cbits = (uint32_t *)(smem + SNN_GEN_WARPS * 32 * 16);
    M.red = smem + SNN_GEN_WARPS * 32 * 32;
    M.colmask = (uint32_t *)(M.red + (SNN_NORM_CHUNKS + 1) * 32);
    M.xs = (float *)(M.colmask + 32 * NG);
    M.evb = (int32_t *)(M.xs + SNN_GEN_WARPS * SNN_P3_MAXEV * 32);
    M.evslot = (uint8_t *)(M.evb + SNN_P3_MAXEV + 16);
    M.live = (uint32_t *)(M.evslot + (B + 15) / 16 * 16);
    M.xt_bytes = (int32_t)gen_xt_bytes(B);
    M.xt = M.xt_bytes ? (float *)((uint8_t *)M.live + (sizeof(uint32_t) * NG + 15) / 16 * 16) : nullptr;
    return M;
}

__device__ __forceinline__ float ld_ext(const snn_layer_t &L, size_t idx, bool &nonbin) {
    if (L.ext_dtype == SNN_EXT_U8) {
        const uint8_t e = ((const uint8_t *)L.ext)[idx];
        nonbin |= e > 1;
        return (float)e;
    }
    const float e = ((const float *)L.ext)[idx];
    nonbin |= (e != 0.0f && e != 1.0f);
    return e;
}

// Spike-gather for one sample: p[j] = sum_{i : s_src[b,i]} w[i, j], i ascending.
// Restates Connection.compute (topology.py:332-346) and MulticompartmentConnection.compute
// with a Weight feature (topology.py:437-479, topology_features.py:633-645) without ever
// materialising the [B, n_src, n_tgt] broadcast.  Per block of 32 words (1024 source neurons) the warp
// first compacts the set bits into an ascending index list in shared memory (ballot-free prefix sum of
// the lanes' popcounts), then walks the list eight entries at a time so that eight weight rows are in
// flight from L2 at once — the sum itself stays one fp32 add per spike in ascending i, like the oracle.
// Weights are read with ld.cg: the CTA that updates a tile in the learning phase is not the CTA that
// gathers from it.
// `first` = the lane's word of block 0 (sb[lane]), loaded by the caller ahead of time.  The words of up to eight
// blocks (8192 source neurons) are fetched together, so a wide but sparse source layer (the 6400 inhibitory neurons of
// BASELINE config 3: 7 blocks, almost always empty) costs one L2 round trip instead of one per block.
__device__ __forceinline__ float gather(const snn_conn_t &C, const uint32_t *__restrict__ sb, int nw_src,
                                        int n_src, int n_tgt, int j, bool valid, int lane, uint16_t *__restrict__ lst, uint32_t first) {
    float p = 0.0f;
    const float *__restrict__ wcol = C.w + j;
    for (int s0 = 0; s0 < nw_src; s0 += 256) {
        uint32_t wd[8];
        #pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int w = s0 + 32 * k + lane;
            wd[k] = (k == 0 && s0 == 0) ? first : (w < nw_src ? __ldcg(sb + w) : 0u);
        }
        uint32_t nzb = 0;   // blocks of this group that hold a spike
        #pragma unroll
        for (int k = 0; k < 8; ++k) nzb |= __any_sync(0xffffffffu, wd[k] != 0u) ? (1u << k) : 0u;
        while (nzb) {
            const int kb = __ffs(nzb) - 1;
            nzb &= nzb - 1;
            uint32_t mine = wd[0];
            #pragma unroll
            for (int k = 1; k < 8; ++k)
                if (kb == k) mine = wd[k];
            const int cnt = __popc(mine);
            int pre = cnt;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, pre, o);
                if (lane >= o) pre += v;
            }
            const int total = __shfl_sync(0xffffffffu, pre, 31);
            int q = pre - cnt;
            while (mine) {
                const int r = __ffs(mine) - 1;
                mine &= mine - 1;
                lst[q++] = (uint16_t)((lane << 5) | r);
            }
            __syncwarp();
            const int base = (s0 + 32 * kb) * 32;
            for (int e = 0; e < total; e += 8) {
                float v[8];
                #pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = base + (int)lst[min(e + k, total - 1)];
                    v[k] = (valid && e + k < total && i < n_src) ? __ldcg(wcol + (size_t)i * n_tgt) : 0.0f;
                }
                #pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (e + k < total) p = p + v[k];
            }
            __syncwarp();
        }
    }
    return p;
}

// Conv2dConnection.compute (topology.py:799-815) for one target neuron j = (co, oy, ox) of one
// sample: the sum of the filter taps whose (zero-padded) input position spiked, in ascending
// (ci, ky, kx) order, then the bias.  STAGED: the sample's source bit row and the filter taps of the
// tile's output channels sit in shared memory (phase 1 stages them once per work unit).
// The part of the gather that depends on the target neuron only (not on the sample): computed once per work unit.
struct ConvGeo {
    int co;                // output channel of neuron j
    int y0, x0;            // source row of filter row 0 (oy*sh - ph); source column of the first VALID tap of a filter row
    int kx_lo, cnt;        // first valid tap of a filter row, number of valid taps
    int ky_lo, ky_hi;      // valid filter rows (source row inside the image)
};
__device__ __forceinline__ ConvGeo conv_geo(const snn_conn_t &C, int j, bool valid) {
    ConvGeo g;
    const int L = C.hout * C.wout;
    g.co = valid ? j / L : 0;
    const int l = valid ? j - g.co * L : 0;
    const int oy = l / C.wout, ox = l - oy * C.wout;
    const int ix0 = ox * C.sw - C.pw;
    g.kx_lo = max(0, -ix0);
    g.x0 = ix0 + g.kx_lo;
    g.cnt = min(C.kw, C.win - ix0) - g.kx_lo;
    // rows: iy = y0 + ky*dh in [0, hin)
    g.y0 = oy * C.sh - C.ph;
    g.ky_lo = g.y0 < 0 ? (-g.y0 + C.dh - 1) / C.dh : 0;
    g.ky_hi = min(C.kh, g.y0 >= C.hin ? 0 : (C.hin - 1 - g.y0) / C.dh + 1);
    return g;
}

template <bool STAGED_BITS, bool STAGED_TAPS>
__device__ __forceinline__ float gather_conv(const snn_conn_t &C, const uint32_t *sb, const float *taps, int co_base, int j, bool valid,
                                             const ConvGeo &g) {
    if (!valid) return 0.0f;
    const int co = g.co;
    const int KK = C.kh * C.kw;
    float p = 0.0f;
    if (C.dw == 1 && C.kw <= 32) {
        // the kw taps of one filter row look at kw CONSECUTIVE source bits: cut that window out of the bit row (two
        // words, one funnel shift) and visit its set bits only — ascending kx, so the order of the sum is unchanged
        if (g.cnt > 0) {
            const float *tp = STAGED_TAPS ? taps + (co - co_base) * C.cin * KK : C.w + (size_t)co * C.cin * KK;
            const uint32_t cmask = g.cnt >= 32 ? 0xffffffffu : ((1u << g.cnt) - 1u);
            for (int ci = 0; ci < C.cin; ++ci)
                for (int ky = g.ky_lo; ky < g.ky_hi; ++ky) {
                    const int iy = g.y0 + ky * C.dh;
                    const int bit0 = (ci * C.hin + iy) * C.win + g.x0, w0 = bit0 >> 5, sft = bit0 & 31;
                    const uint32_t lo = STAGED_BITS ? sb[w0] : __ldcg(sb + w0);
                    const uint32_t hi = sft + g.cnt > 32 ? (STAGED_BITS ? sb[w0 + 1] : __ldcg(sb + w0 + 1)) : 0u;
                    uint32_t bits = __funnelshift_r(lo, hi, sft) & cmask;
                    const int k0 = (ci * C.kh + ky) * C.kw + g.kx_lo;
                    while (bits) {
                        const int k = k0 + __ffs(bits) - 1;
                        bits &= bits - 1;
                        p = p + (STAGED_TAPS ? tp[k] : __ldcg(tp + k));
                    }
                }
        }
        return p + C.b[co];
    }
    const int L = C.hout * C.wout, l = j - co * L, oy = l / C.wout, ox = l - oy * C.wout;   // dilated columns: tap by tap
    for (int ci = 0; ci < C.cin; ++ci)
        for (int ky = 0; ky < C.kh; ++ky) {
            const int iy = oy * C.sh - C.ph + ky * C.dh;
            if (iy < 0 || iy >= C.hin) continue;
            for (int kx = 0; kx < C.kw; ++kx) {
                const int ix = ox * C.sw - C.pw + kx * C.dw;
                if (ix < 0 || ix >= C.win) continue;
                const int i = (ci * C.hin + iy) * C.win + ix;
                const uint32_t word = STAGED_BITS ? sb[i >> 5] : __ldcg(sb + (i >> 5));
                if ((word >> (i & 31)) & 1u) {
                    const int k = (ci * C.kh + ky) * C.kw + kx;
                    p = p + (STAGED_TAPS ? taps[(co - co_base) * C.cin * KK + k] : __ldcg(C.w + (size_t)co * C.cin * KK + k));
                }
            }
        }
    return p + C.b[co];
}

// Final spikes of one neuron of one sample: traces (nodes.py:96-103), clamp / unclamp (network.py:415-429),
// recordings.  Returns the spike that is published.
__device__ __forceinline__ bool finalize_neuron(const DevNet &N, const DevLayer &D, bool s, float xold, size_t k, int b, int j, int t, int wr) {
    const snn_layer_t &L = D.L;
    bool sf = s;
    if (L.traces) {
        const float x = trace_step(xold, s, L.trace_decay, L.trace_scale, L.traces_additive);
        L.x[k] = x;
        if (D.xpub) D.xpub[((size_t)wr * N.B + b) * L.n + j] = x;
    }
    if (L.clamp && L.clamp[(L.clamp_per_step ? (size_t)t * L.n : 0) + j]) sf = true;
    if (L.unclamp && L.unclamp[(L.unclamp_per_step ? (size_t)t * L.n : 0) + j]) sf = false;
    if (t == N.T - 1) L.s[k] = sf ? 1 : 0;
    if (L.rec_s) L.rec_s[((size_t)t * N.B + b) * L.n + j] = sf ? 1 : 0;
    if (L.rec_count && sf) L.rec_count[k] += 1;
    return sf;
}

// ---------------------------------------------------------------------------------------
// phase 1.  Work unit = (layer, 32-neuron tile, chunk of N.cs samples): one warp lane per neuron, the CTA's
// warps stride over the chunk's samples.
__device__ void phase1(const DevNet &N, int li, int tile, int chunk, int t, const GenSmem &M) {
    const DevLayer &D = N.layers[li];
    const snn_layer_t &L = D.L;
    const int B = N.B, n = L.n, nw = D.nw;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = tile * SNN_TILE + lane;
    const bool valid = j < n;
    const int rd = (t + 1) & 1, wr = t & 1;
    const int b0 = chunk * N.cs, b1 = min(B, b0 + N.cs);
    const bool dc = L.kind == SNN_NODE_DC;
    const bool deferred = dc && L.one_spike;  // final spikes known only after the arg-max
    bool nonbin = false;

    if (L.kind == SNN_NODE_INPUT) {
        // Input.forward (nodes.py:211-221): s = x.  Four samples per warp are in flight at once.
        for (int bb = b0 + warp; bb < b1; bb += 4 * SNN_GEN_WARPS) {
            float e[4], xo[4];
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int b = bb + q * SNN_GEN_WARPS;
                const bool ok = valid && b < b1;
                e[q] = (ok && L.ext) ? ld_ext(L, ((size_t)t * B + b) * n + j, nonbin) : 0.0f;
                xo[q] = (ok && L.traces) ? L.x[(size_t)b * n + j] : 0.0f;
            }
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int b = bb + q * SNN_GEN_WARPS;
                if (b >= b1) break;
                const size_t k = (size_t)b * n + j;
                const bool s = e[q] != 0.0f;
                bool sf = false;
                if (valid) {
                    if (L.sum_input) L.summed[k] = L.summed[k] + (s ? 1.0f : 0.0f);
                    sf = finalize_neuron(N, D, s, xo[q], k, b, j, t, wr);
                }
                const uint32_t fw = __ballot_sync(0xffffffffu, valid && sf);
                if (lane == 0) {
                    D.bits[((size_t)wr * B + b) * nw + tile] = fw;
                    if (D.anyf && fw) atomicOr(D.anyf + (size_t)(t % 3) * B + b, 1u);
                }
            }
        }
        if (D.anyf && tile == 0)
            for (int b = b0 + threadIdx.x; b < b1; b += SNN_GEN_THREADS) D.anyf[(size_t)((t + 1) % 3) * B + b] = 0u;
        if (nonbin && N.err) atomicOr(N.err, SNN_ERR_NONBINARY);
        return;
    }

    // adaptive threshold (nodes.py:1078-1079, 1093-1094).  The batch sum of a step's threshold crossers is an
    // integer accumulated with atomics over the sample chunks (thcnt[t % 3]); every unit rebuilds the value the
    // previous step left — thdec (the decayed threshold that step used) + theta_plus * count — so no unit waits
    // for another one; chunk 0 publishes this step's decayed value and clears the counter slot of step t + 1.
    float theta = 0.0f;
    if (dc && valid) {
        if (L.learning) {
            const float prev = t == 0 ? L.theta[j]
                                      : __ldcg(D.thdec + (size_t)rd * n + j) + L.theta_plus * (float)__ldcg(D.thcnt + (size_t)((t + 2) % 3) * n + j);
            theta = prev * L.theta_decay;
            if (chunk == 0 && warp == 0) {
                D.thdec[(size_t)wr * n + j] = theta;
                D.thcnt[(size_t)((t + 1) % 3) * n + j] = 0;
            }
        } else {
            theta = L.theta[j];
        }
    }
    int cnt = 0;  // candidates of this column over this warp's samples

    // a convolutional input: stage the chunk's source bit rows and the taps of this tile's output channels
    int conv_c = -1, conv_slot = 0, co_base = 0;
    bool st_bits = false, st_taps = false;
    ConvGeo geo = {};
    for (int c = 0; c < N.n_conns && conv_c < 0; ++c)
        if (N.conns[c].tgt == li && N.conns[c].kind == SNN_CONN_CONV2D) conv_c = c;
    if (conv_c >= 0) {
        const snn_conn_t &C = N.conns[conv_c];
        const DevLayer &S = N.layers[C.src];
        geo = conv_geo(C, j, valid);
        conv_slot = (N.one_step && C.src < li) ? wr : rd;
        const int words = (b1 - b0) * S.nw;
        st_bits = words <= SNN_CONV_STAGE_WORDS;
        const int Lhw = C.hout * C.wout, K = C.cin * C.kh * C.kw;
        co_base = (tile * SNN_TILE) / Lhw;
        const int co_hi = min(n - 1, tile * SNN_TILE + SNN_TILE - 1) / Lhw;
        const int ntaps = (co_hi - co_base + 1) * K;
        st_taps = ntaps <= SNN_CONV_STAGE_TAPS;
        if (st_bits) {
            const uint32_t *src = S.bits + ((size_t)conv_slot * B + b0) * S.nw;
            for (int k0 = threadIdx.x; k0 < words; k0 += 4 * SNN_GEN_THREADS) {
                uint32_t wq[4];
                #pragma unroll
                for (int q = 0; q < 4; ++q) wq[q] = k0 + q * SNN_GEN_THREADS < words ? __ldcg(src + k0 + q * SNN_GEN_THREADS) : 0u;
                #pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (k0 + q * SNN_GEN_THREADS < words) M.cbits[k0 + q * SNN_GEN_THREADS] = wq[q];
            }
        }
        if (st_taps)
            for (int k = threadIdx.x; k < ntaps; k += SNN_GEN_THREADS) M.xs[k] = __ldcg(C.w + (size_t)co_base * K + k);
        __syncthreads();
    }
    uint16_t *lst = M.list + warp * SNN_GATHER_BLOCK;
    // the incoming connections in insertion order (the first four get their first bit words prefetched)
    int cl[4] = {-1, -1, -1, -1}, ncl = 0, nin = 0;
    for (int c = 0; c < N.n_conns; ++c)
        if (N.conns[c].tgt == li) {
            #pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q == ncl && nin < 4) cl[q] = c;
            if (nin < 4) ++ncl;
            ++nin;
        }

    // first bit words of the dense inputs and the "sample spiked at all" flags of wide sources, fetched one sample
    // ahead: the gather of sample b starts without waiting for L2
    uint32_t fwn[4], afn[4];
    auto prefetch = [&](int b) {
        #pragma unroll
        for (int q = 0; q < 4; ++q) {
            fwn[q] = 0u; afn[q] = 1u;
            if (q < ncl && b < b1 && N.conns[cl[q]].kind != SNN_CONN_CONV2D) {
                const snn_conn_t &C = N.conns[cl[q]];
                const DevLayer &S = N.layers[C.src];
                const int slot = (N.one_step && C.src < li) ? wr : rd;
                if (lane < S.nw) fwn[q] = __ldcg(S.bits + ((size_t)slot * B + b) * S.nw + lane);
                if (S.anyf) afn[q] = __ldcg(S.anyf + (size_t)(slot == wr ? t % 3 : (t + 2) % 3) * B + b);
            }
        }
    };
    prefetch(b0 + warp);

    for (int b = b0 + warp; b < b1; b += SNN_GEN_WARPS) {
        const size_t k = (size_t)b * n + j;
        uint32_t fw[4], af[4];
        #pragma unroll
        for (int q = 0; q < 4; ++q) { fw[q] = fwn[q]; af[q] = afn[q]; }
        prefetch(b + SNN_GEN_WARPS);
        float v = 0.0f, rc = 0.0f, xold = 0.0f, ic = 0.0f;
        if (valid) {
            v = L.v[k];
            if (L.kind != SNN_NODE_MCP) rc = L.refrac_count[k];
            if (L.traces && !deferred) xold = L.x[k];
            if (L.kind == SNN_NODE_CURRENT_LIF) ic = L.i[k];
        }
        // network.py:211-250: accumulate every incoming connection in insertion order
        float cur = 0.0f;
        const bool has_in = nin > 0;
        int seen = 0;
        for (int c = 0; c < N.n_conns; ++c) {
            const snn_conn_t &C = N.conns[c];
            if (C.tgt != li) continue;
            const DevLayer &S = N.layers[C.src];
            uint32_t first = 0u, anysp = 1u;
            #pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q == seen) { first = fw[q]; anysp = af[q]; }
            const bool prefetched = seen < 4;
            ++seen;
            // one-step mode (network.py:393-396): sources earlier in the insertion order have already
            // produced this step's spikes (slot wr); everything else is still at step t-1 (slot rd)
            const int slot = (N.one_step && C.src < li) ? wr : rd;
            float p;
            if (C.kind == SNN_CONN_CONV2D) {
                const uint32_t *gsb = S.bits + ((size_t)slot * B + b) * S.nw;
                if (c == conv_c && st_bits) {
                    const uint32_t *ssb = M.cbits + (size_t)(b - b0) * S.nw;
                    p = st_taps ? gather_conv<true, true>(C, ssb, M.xs, co_base, j, valid, geo) : gather_conv<true, false>(C, ssb, nullptr, 0, j, valid, geo);
                } else if (c == conv_c && st_taps) {
                    p = gather_conv<false, true>(C, gsb, M.xs, co_base, j, valid, geo);
                } else {
                    p = gather_conv<false, false>(C, gsb, nullptr, 0, j, valid, c == conv_c ? geo : conv_geo(C, j, valid));
                }
            } else {
                const uint32_t *sbr = S.bits + ((size_t)slot * B + b) * S.nw;
                if (!prefetched) first = lane < S.nw ? __ldcg(sbr + lane) : 0u;
                p = anysp ? gather(C, sbr, S.nw, S.L.n, n, j, valid, lane, lst, first) : 0.0f;
                if (C.b && valid) p = p + C.b[j];
            }
            cur = cur + p;
        }
        bool s = false;
        if (valid) {
            // (one-step mode: the connection input REPLACES the external one, network.py:393-396)
            if (L.ext && !(N.one_step && has_in)) { bool nb = false; cur = cur + ld_ext(L, ((size_t)t * B + b) * n + j, nb); }
            if (L.inject_v) v = v + L.inject_v[(L.inject_per_step ? (size_t)t * n : 0) + j];  // network.py:398-404
            float xin = cur;
            if (dc) {
                s = dc_step(L, v, rc, xin, theta);
                if (L.has_lbound && v < L.lbound) v = L.lbound;  // nodes.py:1108-1109
            } else if (L.kind == SNN_NODE_IF) {
                s = if_step(L, v, rc, xin);
            } else if (L.kind == SNN_NODE_CURRENT_LIF) {
                s = clif_step(L, v, rc, ic, xin);
                L.i[k] = ic;
            } else if (L.kind == SNN_NODE_BOOSTED_LIF) {
                s = boosted_step(L, v, rc, xin);
            } else if (L.kind == SNN_NODE_MCP) {   // McCullochPitts.forward (nodes.py:278-288): voltages equal the inputs
                v = xin;
                s = v >= L.thresh;
            } else {
                s = lif_step(L, v, rc, xin);
            }
            L.v[k] = v;
            if (L.kind != SNN_NODE_MCP) L.refrac_count[k] = rc;
            if (L.sum_input) L.summed[k] = L.summed[k] + xin;
            if (L.rec_v) L.rec_v[((size_t)t * B + b) * n + j] = v;
            cnt += s ? 1 : 0;
        }
        const uint32_t word = __ballot_sync(0xffffffffu, valid && s);
        if (deferred) {
            if (lane == 0) D.candbits[(size_t)b * nw + tile] = word;
            if (word) {
                unsigned long long key = 0ull;
                if (valid && s) key = snn_one_spike_key(N.seed, (uint32_t)t + N.step_offset, (uint32_t)li, (uint32_t)b, (uint32_t)j);
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                    key = other > key ? other : key;
                }
                if (lane == 0) atomicMax(D.keys + (size_t)wr * B + b, key);
            }
        } else {
            bool sf = false;
            if (valid) sf = finalize_neuron(N, D, s, xold, k, b, j, t, wr);
            const uint32_t fw = __ballot_sync(0xffffffffu, valid && sf);
            if (lane == 0) {
                D.bits[((size_t)wr * B + b) * nw + tile] = fw;
                if (D.anyf && fw) atomicOr(D.anyf + (size_t)(t % 3) * B + b, 1u);
            }
        }
    }
    if (D.anyf && tile == 0)   // the flag slot of step t + 1 (last read in step t - 1)
        for (int b = b0 + threadIdx.x; b < b1; b += SNN_GEN_THREADS) D.anyf[(size_t)((t + 1) % 3) * B + b] = 0u;

    // theta += theta_plus * sum_b s  (nodes.py:1093-1094)
    if (dc && L.learning && valid && cnt > 0) atomicAdd(D.thcnt + (size_t)(t % 3) * n + j, cnt);
    if (deferred && tile == 0) {
        // clear the key slot the NEXT step will arg-max into (last read two barriers ago)
        for (int b = b0 + threadIdx.x; b < b1; b += SNN_GEN_THREADS) D.keys[(size_t)rd * B + b] = 0ull;
    }
    if (conv_c >= 0) __syncthreads();   // the staged rows / taps are overwritten by the CTA's next unit
}

// ---------------------------------------------------------------------------------------
// phase 2 (DiehlAndCookNodes with one_spike): keep the arg-max candidate of each sample
// (nodes.py:1097-1105), then traces / clamp / publish as in phase 1.  Same work units as phase 1.
__device__ void phase2(const DevNet &N, int li, int tile, int chunk, int t) {
    const DevLayer &D = N.layers[li];
    const snn_layer_t &L = D.L;
    const int B = N.B, n = L.n, nw = D.nw;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = tile * SNN_TILE + lane;
    const bool valid = j < n;
    const int wr = t & 1;
    const int b0 = chunk * N.cs, b1 = min(B, b0 + N.cs);
    for (int bb = b0 + warp; bb < b1; bb += 4 * SNN_GEN_WARPS) {
        uint32_t cand[4];
        unsigned long long key[4];
        float xo[4];
        #pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int b = bb + q * SNN_GEN_WARPS;
            const bool ok = b < b1;
            cand[q] = ok ? __ldcg(D.candbits + (size_t)b * nw + tile) : 0u;
            key[q] = ok ? __ldcg(D.keys + (size_t)wr * B + b) : 0ull;
            xo[q] = (ok && valid && L.traces) ? L.x[(size_t)b * n + j] : 0.0f;
        }
        #pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int b = bb + q * SNN_GEN_WARPS;
            if (b >= b1) break;
            const size_t k = (size_t)b * n + j;
            const bool s = valid && ((cand[q] >> lane) & 1u) && key[q] != 0ull && (uint32_t)(key[q] & 0xffffffffull) == (uint32_t)j;
            bool sf = false;
            if (valid) sf = finalize_neuron(N, D, s, xo[q], k, b, j, t, wr);
            const uint32_t fw = __ballot_sync(0xffffffffu, valid && sf);
            if (lane == 0) {
                D.bits[((size_t)wr * B + b) * nw + tile] = fw;
                if (D.anyf && fw) atomicOr(D.anyf + (size_t)(t % 3) * B + b, 1u);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// phase 3: learning-rule update of the rows [wg0*32, wg1*32) of one weight tile W[:, tile] of connection `ci`.
//   U[i,j] = reduce_b s_src[b,i] * (x_tgt[b,j] * nu0)      pre-synaptic term
//   V[i,j] = reduce_b x_src[b,i] * (s_tgt[b,j] * nu1)      post-synaptic term
// Both reduce over the batch in ascending b.  The reference materialises [B,n_src,n_tgt]
// (learning.py:399-417, MCC_learning.py:234-299); here only rows with a pre-synaptic spike
// and columns with a post-synaptic spike are touched (everything else is a bitwise no-op),
// except when a full pass is required: weight decay != 1, or the first update of the window
// (entries may sit outside [wmin, wmax] after normalize()).
// Work unit = (connection, tile, row chunk): the target traces of the tile are staged in shared memory once
// (every row group reads them), a warp owns one group of 32 source rows at a time, stages the pre-synaptic
// traces of the (few) samples with a post-synaptic event for those rows, and rewrites the rows eight at a time
// so that eight weight loads are in flight.
__device__ void phase3(const DevNet &N, int ci, int tile, int wg0, int wg1, int t, const GenSmem &M) {
    const snn_conn_t &C = N.conns[ci];
    const DevLayer &S = N.layers[C.src], &G = N.layers[C.tgt];
    const int B = N.B, ns = S.L.n, nt = G.L.n, nwS = S.nw, nwG = G.nw;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = tile * SNN_TILE + lane;
    const bool valid = j < nt;
    const int wr = t & 1;
    const int NG = (B + 31) / 32;
    const bool stdp = SNN_RULE_IS_STDP(C.rule);
    const bool wdep = C.rule == SNN_RULE_WDEP_POSTPRE || C.rule == SNN_RULE_HEBBIAN;   // plain sums: nu applied after the reduction
    const bool pre_on = stdp && C.nu0 != 0.0f, post_on = stdp && C.nu1 != 0.0f;
    const bool decay_on = C.weight_decay != 0.0f && C.weight_decay != 1.0f;
    const bool full = decay_on || (C.has_clamp && t == 0);
    const float Bf = (float)B;
    const bool stage = pre_on && M.xt != nullptr;

    // Samples whose target-trace row is all zero in this tile cannot change the pre-synaptic term (adding +-0 to a sum
    // that started at +0 is a bitwise no-op), and in networks of rarely spiking neurons that is most of them: the
    // staging pass records the others in `live`, the accumulation looks at nobody else.
    if (stage) {
        if (threadIdx.x < NG) M.live[threadIdx.x] = 0u;
        __syncthreads();
        for (int b0 = warp; b0 < B; b0 += 8 * SNN_GEN_WARPS) {   // eight rows of the trace tile in flight per warp
            float tx[8];
            #pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int b = b0 + q * SNN_GEN_WARPS;
                tx[q] = (valid && b < B) ? __ldcg(G.L.x + (size_t)b * nt + j) : 0.0f;
            }
            #pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int b = b0 + q * SNN_GEN_WARPS;
                if (b < B) {
                    const float v = wdep ? tx[q] : tx[q] * C.nu0;
                    M.xt[b * 32 + lane] = v;
                    const bool nzrow = __any_sync(0xffffffffu, v != 0.0f);
                    if (lane == 0 && nzrow) atomicOr(M.live + (b >> 5), 1u << (b & 31));
                }
            }
        }
    }

    // column events: colmask[g*32 + lane] = samples of group g whose target spike hit column j
    uint32_t colany = 0;
    if (post_on) {
        for (int b = threadIdx.x; b < B; b += SNN_GEN_THREADS) M.evslot[b] = 0xFF;
        for (int g = warp; g < NG; g += SNN_GEN_WARPS) {
            const int b = g * 32 + lane;
            const uint32_t wb = b < B ? __ldcg(G.bits + ((size_t)wr * B + b) * nwG + tile) : 0u;
            uint32_t mine = 0;
            if (__any_sync(0xffffffffu, wb != 0u)) {
                #pragma unroll
                for (int r = 0; r < 32; ++r) {
                    const uint32_t m = __ballot_sync(0xffffffffu, (wb >> r) & 1u);
                    if (lane == r) mine = m;
                }
            }
            M.colmask[g * 32 + lane] = mine;
        }
        __syncthreads();
        for (int g = 0; g < NG; ++g) colany |= M.colmask[g * 32 + lane];
        // the first SNN_P3_MAXEV samples (ascending) with an event in this tile get a staging slot
        if (warp == 0) {
            int nev = 0;
            for (int g = 0; g < NG; ++g) {
                uint32_t m = M.colmask[g * 32 + lane];
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
                if (lane == 0)
                    while (m && nev < SNN_P3_MAXEV) {
                        const int bb = g * 32 + __ffs(m) - 1;
                        m &= m - 1;
                        M.evb[nev] = bb;
                        M.evslot[bb] = (uint8_t)nev;
                        ++nev;
                    }
            }
            if (lane == 0) M.evb[SNN_P3_MAXEV] = nev;
        }
    }
    const bool any_col = __syncthreads_or(colany != 0u) != 0;   // also publishes xt / evb / evslot
    if (!full && !pre_on && !any_col) return;

    float *acc = M.acc + warp * (32 * 32);
    for (int r = 0; r < 32; ++r) acc[r * 32 + lane] = 0.0f;
    float *xsw = M.xs + warp * (SNN_P3_MAXEV * 32);
    const int nev = any_col ? M.evb[SNN_P3_MAXEV] : 0;
    __syncwarp();

    // In the dense regime (large batches: nearly every source row has a spike somewhere in the batch) the weight rows
    // of a group are fetched before it is known which of them change — their L2 round trip then hides behind the
    // accumulation; for small batches only the rows that need it are read.
    const bool eager = B >= 64;
    const bool post_t = colany != 0u;
    for (int wg = wg0 + warp; wg < wg1; wg += SNN_GEN_WARPS) {
        const int i0 = wg * 32;
        float wv[8];
        if (eager) {
            #pragma unroll
            for (int q = 0; q < 8; ++q) wv[q] = (valid && i0 + q < ns) ? __ldcg(C.w + (size_t)(i0 + q) * nt + j) : 0.0f;
        }
        if (nev > 0) {   // pre-synaptic traces of the event samples for my 32 rows (coalesced, four in flight)
            const int i = i0 + lane;
            for (int e0 = 0; e0 < nev; e0 += 4) {
                float xv[4];
                #pragma unroll
                for (int q = 0; q < 4; ++q)
                    xv[q] = (e0 + q < nev && i < ns) ? __ldcg(S.xpub + ((size_t)wr * B + M.evb[min(e0 + q, nev - 1)]) * ns + i) : 0.0f;
                #pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (e0 + q < nev) xsw[(e0 + q) * 32 + lane] = xv[q];
            }
        }
        uint32_t tmask = 0, umask = 0;   // rows with a pre-synaptic spike / with a non-zero contribution to U
        if (pre_on) {
            for (int g0 = 0; g0 < NG; g0 += 8) {
                uint32_t mine[8];
                #pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int bl = (g0 + q) * 32 + lane;
                    mine[q] = (g0 + q < NG && bl < B) ? __ldcg(S.bits + ((size_t)wr * B + bl) * nwS + wg) : 0u;
                }
                #pragma unroll
                for (int q = 0; q < 8; ++q) {
                    uint32_t nz = __ballot_sync(0xffffffffu, mine[q] != 0u);
                    {   // every spiking sample marks its rows; only the live ones are accumulated
                        uint32_t orw = mine[q];
                        #pragma unroll
                        for (int o = 16; o > 0; o >>= 1) orw |= __shfl_xor_sync(0xffffffffu, orw, o);
                        tmask |= orw;
                    }
                    if (stage && g0 + q < NG) nz &= M.live[g0 + q];
                    while (nz) {
                        const int bb = __ffs(nz) - 1;
                        nz &= nz - 1;
                        uint32_t word = __shfl_sync(0xffffffffu, mine[q], bb);
                        const int b = (g0 + q) * 32 + bb;
                        float tx = 0.0f;
                        if (stage) {
                            tx = M.xt[b * 32 + lane];
                        } else {
                            if (valid) {
                                tx = __ldcg(G.L.x + (size_t)b * nt + j);
                                if (!wdep) tx = tx * C.nu0;
                            }
                            if (!__any_sync(0xffffffffu, tx != 0.0f)) continue;
                        }
                        umask |= word;
                        while (word) {
                            const int r = __ffs(word) - 1;
                            word &= word - 1;
                            acc[r * 32 + lane] = acc[r * 32 + lane] + tx;
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (!full && !(wdep ? tmask : umask) && !any_col) continue;
        // Columns with a post-synaptic event change in every row.  They are few (one spiking neuron in the tile is
        // the rule), so each is rewritten with one lane per ROW — one pass of the rule for all 32 rows of the group —
        // instead of dragging the whole warp through 32 row iterations for the sake of one lane.
        const uint32_t evcols = __ballot_sync(0xffffffffu, post_t);
        for (uint32_t ec = evcols; ec; ec &= ec - 1) {
            const int jl = __ffs(ec) - 1, jc = tile * SNN_TILE + jl, i = i0 + lane;
            if (i < ns) {
                const bool pre_t = (tmask >> lane) & 1u;
                float U = 0.0f, V = 0.0f;
                if (pre_t) {
                    U = acc[lane * 32 + jl];
                    if (C.reduction == SNN_REDUCE_MEAN) U = U / Bf;
                }
                for (int g = 0; g < NG; ++g) {
                    uint32_t m = M.colmask[g * 32 + jl];
                    while (m) {
                        const int b = g * 32 + __ffs(m) - 1;
                        m &= m - 1;
                        const int slot = M.evslot[b];
                        const float xs = slot != 0xFF ? xsw[slot * 32 + lane] : __ldcg(S.xpub + ((size_t)wr * B + b) * ns + i);
                        V = V + xs * (wdep ? 1.0f : C.nu1);
                    }
                }
                if (C.reduction == SNN_REDUCE_MEAN) V = V / Bf;
                float *wp = C.w + (size_t)i * nt + jc;
                *wp = apply_rule(C, __ldcg(wp), U, pre_t, V, true);
            }
        }
        __syncwarp();
        for (int r0 = 0; r0 < 32; r0 += 8) {
            bool nd[8];
            bool anyneed = false;
            #pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int r = r0 + q, i = i0 + r;
                const bool pre_t = (tmask >> r) & 1u;
                // a row whose U is +0 in every column is left alone: w - 0 is w, and w has been inside [wmin, wmax] since
                // the full pass of step 0 (the weight-dependent and Hebbian forms compute w + (+-0), which may flip
                // the sign of a zero: they always rewrite)
                const bool touched = wdep ? pre_t : ((umask >> r) & 1u) != 0u;
                nd[q] = valid && i < ns && !post_t && (full || touched);   // event columns: done above
                if (!eager) wv[q] = nd[q] ? __ldcg(C.w + (size_t)i * nt + j) : 0.0f;
                anyneed |= nd[q];
            }
            float wn[8];   // next group's rows: in flight while this group is rewritten
            if (eager && r0 + 8 < 32) {
                #pragma unroll
                for (int q = 0; q < 8; ++q) wn[q] = (valid && i0 + r0 + 8 + q < ns) ? __ldcg(C.w + (size_t)(i0 + r0 + 8 + q) * nt + j) : 0.0f;
            }
            if (__any_sync(0xffffffffu, anyneed)) {
                #pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int r = r0 + q, i = i0 + r;
                    const bool pre_t = (tmask >> r) & 1u;
                    if (nd[q]) {
                        float U = 0.0f;
                        if (pre_t) {
                            U = acc[r * 32 + lane];
                            if (C.reduction == SNN_REDUCE_MEAN) U = U / Bf;
                        }
                        C.w[(size_t)i * nt + j] = apply_rule(C, wv[q], U, pre_t, 0.0f, false);
                    }
                }
            }
            #pragma unroll
            for (int q = 0; q < 8; ++q)
                if ((umask >> (r0 + q)) & 1u) acc[(r0 + q) * 32 + lane] = 0.0f;
            if (eager) {
                #pragma unroll
                for (int q = 0; q < 8; ++q) wv[q] = wn[q];
            }
        }
        __syncwarp();
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// phase 3 for reward-modulated STDP and for convolutional connections.  The rule's state is double
// buffered (DevMstdp): everything is READ from slot `in` and WRITTEN to slot `out`, so no thread
// overwrites a value another thread still needs in this step.
__device__ __forceinline__ float mst_trace(float p, float decay, float a, bool s) {
    // learning.py:1564-1567 / 1999-2003:  P *= exp(-dt/tc);  P += a * s
    const float x = p * decay;
    return x + a * (s ? 1.0f : 0.0f);
}
__device__ __forceinline__ bool bit_of(const uint32_t *row, int i) { return (__ldcg(row + (i >> 5)) >> (i & 31)) & 1u; }

// learning.MSTDP._connection_update (learning.py:1504-1574) + base class decay / clamp (:87-104).
// Work is spread over the tiles of the SOURCE layer (a dense layer's target is often tiny — 10 output
// neurons in BASELINE config 4 — while its source has thousands of rows): unit = 32 rows i (one per
// lane) x all columns j (warps stride over them); the batch sum runs in ascending b per (i, j).  When it
// fits, the rule state the batch loop reads (p_plus and the pre-synaptic spikes of the unit's rows, p_minus
// and the post-synaptic spikes of all columns) is staged in shared memory first, so that the B-long
// dependent loop never waits for L2.
__device__ void phase3_mstdp_dense(const DevNet &N, int ci_, int tile, int t, const GenSmem &GS) {
    const snn_conn_t &C = N.conns[ci_];
    const DevMstdp &M = N.mst[ci_];
    const DevLayer &S = N.layers[C.src], &G = N.layers[C.tgt];
    const int B = N.B, ns = S.L.n, nt = G.L.n;
    const int in = (t + N.T) & 1, out = in ^ 1, wr = t & 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = tile * SNN_TILE + lane;
    const float Bf = (float)B;
    const float *pp = M.pp[in], *pm = M.pm[in];
    const uint8_t *sp = M.sp[in], *st = M.st[in];
    const bool staged = GS.xt != nullptr && (size_t)B * 32 + (size_t)B * nt * 5 + 16 <= sizeof(float) * SNN_GEN_WARPS * 32 * 32;
    float *pp_s = GS.xt;                                   // [B][32]
    float *pm_s = GS.acc;                                  // [B][nt]
    uint8_t *st_s = (uint8_t *)(pm_s + (size_t)B * nt);    // [B][nt]
    uint8_t *sp_s = st_s + ((size_t)B * nt + 15) / 16 * 16;  // [B][32]
    uint32_t *sbw = GS.colmask;                            // [B] source spike word of this tile, step t
    if (staged) {
        for (int b0 = warp; b0 < B; b0 += 8 * SNN_GEN_WARPS) {   // eight samples' rows in flight per warp
            float pv[8]; uint8_t sv[8];
            #pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int b = b0 + q * SNN_GEN_WARPS;
                const bool ok = i < ns && b < B;
                pv[q] = ok ? __ldcg(pp + (size_t)b * ns + i) : 0.0f;
                sv[q] = ok ? __ldcg(sp + (size_t)b * ns + i) : (uint8_t)0;
            }
            #pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int b = b0 + q * SNN_GEN_WARPS;
                if (b < B) { pp_s[b * 32 + lane] = pv[q]; sp_s[b * 32 + lane] = sv[q]; }
            }
        }
        for (int k0 = threadIdx.x; k0 < B * nt; k0 += 4 * SNN_GEN_THREADS) {
            float pv[4]; uint8_t sv[4];
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = k0 + q * SNN_GEN_THREADS;
                pv[q] = k < B * nt ? __ldcg(pm + k) : 0.0f;
                sv[q] = k < B * nt ? __ldcg(st + k) : (uint8_t)0;
            }
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = k0 + q * SNN_GEN_THREADS;
                if (k < B * nt) { pm_s[k] = pv[q]; st_s[k] = sv[q]; }
            }
        }
        for (int b = threadIdx.x; b < B; b += SNN_GEN_THREADS) sbw[b] = __ldcg(S.bits + ((size_t)wr * B + b) * S.nw + tile);
        __syncthreads();
    }
    // weight update from the eligibility of the previous step = p_plus (x) s_post + s_pre (x) p_minus
    if (C.rule == SNN_RULE_MSTDPET) {
        // learning.MSTDPET._connection_update (learning.py:2187-2249), batch size 1: the eligibility feeds a decaying
        // trace per synapse, the reward modulates the trace
        if (i < ns)
            for (int j = warp; j < nt; j += SNN_GEN_WARPS) {
                const bool ss = staged ? sp_s[lane] != 0 : __ldcg(sp + i) != 0;
                const bool tt = staged ? st_s[j] != 0 : __ldcg(st + j) != 0;
                const float ppv = staged ? pp_s[lane] : __ldcg(pp + i), pmv = staged ? pm_s[j] : __ldcg(pm + j);
                const float e = ppv * (tt ? 1.0f : 0.0f) + (ss ? 1.0f : 0.0f) * pmv;      // :2245-2247 of the previous step
                float et = __ldcg(C.e_trace + (size_t)i * nt + j) * C.e_trace_decay;   // :2229
                et = et + e / C.tc_e_trace;                                            // :2230
                C.e_trace[(size_t)i * nt + j] = et;
                float x = __ldcg(C.w + (size_t)i * nt + j) + C.et_coef * et;           // :2232-2238
                if (C.weight_decay != 0.0f) x = x * C.weight_decay;
                if (C.has_clamp) x = clampf(x, C.wmin, C.wmax);
                C.w[(size_t)i * nt + j] = x;
            }
    } else if (i < ns)
        for (int j = warp; j < nt; j += SNN_GEN_WARPS) {
            float upd = 0.0f;
            if (staged) {
                for (int b = 0; b < B; ++b) {
                    const bool ss = sp_s[b * 32 + lane] != 0, tt = st_s[b * nt + j] != 0;
                    if (!ss && !tt) continue;
                    const float e = pp_s[b * 32 + lane] * (tt ? 1.0f : 0.0f) + (ss ? 1.0f : 0.0f) * pm_s[b * nt + j];
                    upd = upd + C.reward * e;
                }
            } else {
                for (int b = 0; b < B; ++b) {
                    const bool ss = __ldcg(sp + (size_t)b * ns + i) != 0, tt = __ldcg(st + (size_t)b * nt + j) != 0;
                    if (!ss && !tt) continue;
                    const float e = __ldcg(pp + (size_t)b * ns + i) * (tt ? 1.0f : 0.0f) + (ss ? 1.0f : 0.0f) * __ldcg(pm + (size_t)b * nt + j);
                    upd = upd + C.reward * e;
                }
            }
            if (C.reduction == SNN_REDUCE_MEAN) upd = upd / Bf;
            float x = __ldcg(C.w + (size_t)i * nt + j) + C.nu0 * upd;
            if (C.weight_decay != 0.0f) x = x * C.weight_decay;
            if (C.has_clamp) x = clampf(x, C.wmin, C.wmax);
            C.w[(size_t)i * nt + j] = x;
        }
    // P+ and the pre-synaptic spikes of this step for my rows; tile 0 also does P- and the post side
    if (i < ns)
        for (int b = warp; b < B; b += SNN_GEN_WARPS) {
            const size_t k = (size_t)b * ns + i;
            bool s;
            float p;
            if (staged) { s = (sbw[b] >> lane) & 1u; p = pp_s[b * 32 + lane]; }
            else { s = bit_of(S.bits + ((size_t)wr * B + b) * S.nw, i); p = __ldcg(pp + k); }
            M.pp[out][k] = mst_trace(p, C.p_plus_decay, C.a_plus, s);
            M.sp[out][k] = s ? 1 : 0;
        }
    if (tile == 0)
        for (size_t k = threadIdx.x; k < (size_t)B * nt; k += SNN_GEN_THREADS) {
            const int b = (int)(k / nt), jj = (int)(k - (size_t)b * nt);
            const bool s = bit_of(G.bits + ((size_t)wr * B + b) * G.nw, jj);
            M.pm[out][k] = mst_trace(__ldcg(pm + k), C.p_minus_decay, C.a_minus, s);
            M.st[out][k] = s ? 1 : 0;
        }
    if (staged) __syncthreads();   // the staging buffers belong to the CTA's next unit
}

// Ascending iterator over the set bits of row[lo, hi) (a bit row in shared memory, or in L2 when STAGED is off).
struct BitIter {
    const uint32_t *row;
    int w, wend, lo, hi;
    uint32_t cur;
    bool staged;
    __device__ __forceinline__ uint32_t load(int ww) const {
        uint32_t x = staged ? row[ww] : __ldcg(row + ww);
        const int base = ww * 32;
        if (base < lo) x &= 0xffffffffu << (lo - base);
        if (base + 32 > hi) x &= (hi - base >= 32) ? 0xffffffffu : ((1u << (hi - base)) - 1u);
        return x;
    }
    __device__ __forceinline__ void init(const uint32_t *r, int lo_, int hi_, bool st) {
        row = r; lo = lo_; hi = hi_; staged = st;
        w = lo >> 5; wend = (hi - 1) >> 5;
        cur = hi > lo ? load(w) : 0u;
        if (hi <= lo) w = wend = 0;
    }
    __device__ __forceinline__ int next() {
        while (cur == 0u) {
            if (w >= wend) return -1;
            ++w;
            cur = load(w);
        }
        const int idx = w * 32 + __ffs(cur) - 1;
        cur &= cur - 1;
        return idx;
    }
};

// learning.MSTDP._conv2d_connection_update (learning.py:1942-2015) with a per-sample eligibility
// (SURVEY.md §0.8), PostPre / WeightDependentPostPre / Hebbian on a Conv2dConnection, and the decay-only
// update of a conv connection without a rule (learning.NoOp).  Called once per CTA and step: every loop is
// spread over the whole grid (cta of ncta).
__device__ void phase3_conv(const DevNet &N, int ci_, int cta, int ncta, int t, const GenSmem &GS) {
    const snn_conn_t &C = N.conns[ci_];
    const DevMstdp &M = N.mst[ci_];
    const DevLayer &S = N.layers[C.src], &G = N.layers[C.tgt];
    const int B = N.B, ns = S.L.n, nt = G.L.n;
    const int KK = C.kh * C.kw, K = C.cin * KK, L = C.hout * C.wout, NWT = C.cout * K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t start = (size_t)cta * SNN_GEN_THREADS + threadIdx.x, stride = (size_t)ncta * SNN_GEN_THREADS;
    if (SNN_RULE_IS_STDP(C.rule)) {
        // PostPre / WeightDependentPostPre / Hebbian on the im2col views (learning.py:457-497, 920-975, 1348-1380):
        // per filter tap (co, k) the inner sum runs over the output positions (ascending) of one sample, the outer
        // one over the samples (ascending) — the oracle's order.  Traces are read from the layers' own arrays:
        // the grid barriers before and after the learning phase frame them.
        const bool hebb = C.rule == SNN_RULE_HEBBIAN;
        const bool pre_on = C.nu0 != 0.0f || hebb, post_on = C.nu1 != 0.0f || hebb;
        const int wr = t & 1;
        for (size_t e = start; e < (size_t)NWT; e += stride) {
            const int co = (int)(e / K), k = (int)(e - (size_t)co * K);
            const int ci = k / KK, kk = k - ci * KK, ky = kk / C.kw, kx = kk - ky * C.kw;
            float U = 0.0f, V = 0.0f;
            for (int b = 0; b < B; ++b) {
                const uint32_t *sb = S.bits + ((size_t)wr * B + b) * S.nw, *gb = G.bits + ((size_t)wr * B + b) * G.nw;
                float u1 = 0.0f, v1 = 0.0f;
                for (int oy = 0; oy < C.hout; ++oy) {
                    const int iy = oy * C.sh - C.ph + ky;
                    if (iy < 0 || iy >= C.hin) continue;
                    for (int ox = 0; ox < C.wout; ++ox) {
                        const int ix = ox * C.sw - C.pw + kx;
                        if (ix < 0 || ix >= C.win) continue;
                        const int src = (ci * C.hin + iy) * C.win + ix, tgt = co * L + oy * C.wout + ox;
                        if (pre_on && bit_of(sb, src)) u1 = u1 + __ldcg(G.L.x + (size_t)b * nt + tgt);
                        if (post_on && bit_of(gb, tgt)) v1 = v1 + __ldcg(S.L.x + (size_t)b * ns + src);
                    }
                }
                U = U + u1; V = V + v1;
            }
            if (C.reduction == SNN_REDUCE_MEAN) { U = U / (float)B; V = V / (float)B; }
            float x = __ldcg(C.w + e);
            if (C.rule == SNN_RULE_WDEP_POSTPRE) {
                float upd = 0.0f;
                if (pre_on) upd = upd - (C.nu0 * U) * (x - C.wmin);
                if (post_on) upd = upd + (C.nu1 * V) * (C.wmax - x);
                x = x + upd;
            } else if (hebb) {
                x = x + C.nu0 * U;
                x = x + C.nu1 * V;
            } else {
                if (pre_on) x = x - C.nu0 * U;
                if (post_on) x = x + C.nu1 * V;
            }
            if (C.weight_decay != 0.0f) x = x * C.weight_decay;
            if (C.has_clamp) x = clampf(x, C.wmin, C.wmax);
            C.w[e] = x;
        }
        return;
    }
    if (C.rule != SNN_RULE_MSTDP) {  // learning.NoOp: w *= weight_decay (learning.py:93-94), no clamp
        if (C.rule == SNN_RULE_NOOP && C.weight_decay != 0.0f)
            for (size_t k = start; k < (size_t)NWT; k += stride) C.w[k] = __ldcg(C.w + k) * C.weight_decay;
        return;
    }
    const int in = (t + N.T) & 1, out = in ^ 1, wr = t & 1;
    const float *pp = M.pp[in], *pm = M.pm[in], *el = M.el[in];
    // w += nu0 * sum_b reward * eligibility(t-1)  (:1973-1974), then decay / clamp (learning.py:87-104).
    // One warp per filter tap: the lanes fetch 32 samples' eligibilities at once, the sum itself stays serial
    // in ascending b (shuffles, no memory latency in the chain).
    for (int k = cta * SNN_GEN_WARPS + warp; k < NWT; k += ncta * SNN_GEN_WARPS) {
        float upd = 0.0f;
        for (int bq = 0; bq < B; bq += 32) {
            const float val = (bq + lane < B) ? __ldcg(el + (size_t)(bq + lane) * NWT + k) : 0.0f;
            const int m = min(32, B - bq);
            for (int q = 0; q < m; ++q) upd = upd + C.reward * __shfl_sync(0xffffffffu, val, q);
        }
        if (lane == 0) {
            float x = __ldcg(C.w + k) + C.nu0 * upd;
            if (C.weight_decay != 0.0f) x = x * C.weight_decay;
            if (C.has_clamp) x = clampf(x, C.wmin, C.wmax);
            C.w[k] = x;
        }
    }
    // P+ (trace image of the source), P- (:1999-2003); four elements per thread in flight
    {
        const size_t tot = (size_t)B * ns;
        for (size_t k0 = start; k0 < tot; k0 += 4 * stride) {
            float pv[4]; bool sv[4];
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                const size_t k = k0 + q * stride;
                pv[q] = 0.0f; sv[q] = false;
                if (k < tot) {
                    const int b = (int)(k / ns), i = (int)(k - (size_t)b * ns);
                    pv[q] = __ldcg(pp + k);
                    sv[q] = bit_of(S.bits + ((size_t)wr * B + b) * S.nw, i);
                }
            }
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                const size_t k = k0 + q * stride;
                if (k < tot) M.pp[out][k] = mst_trace(pv[q], C.p_plus_decay, C.a_plus, sv[q]);
            }
        }
    }
    {
        const size_t tot = (size_t)B * nt;
        for (size_t k0 = start; k0 < tot; k0 += 4 * stride) {
            float pv[4]; bool sv[4];
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                const size_t k = k0 + q * stride;
                pv[q] = 0.0f; sv[q] = false;
                if (k < tot) {
                    const int b = (int)(k / nt), jj = (int)(k - (size_t)b * nt);
                    pv[q] = __ldcg(pm + k);
                    sv[q] = bit_of(G.bits + ((size_t)wr * B + b) * G.nw, jj);
                }
            }
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                const size_t k = k0 + q * stride;
                if (k < tot) M.pm[out][k] = mst_trace(pv[q], C.p_minus_decay, C.a_minus, sv[q]);
            }
        }
    }
    // eligibility(t)[b,co,k] = sum_l s_post[b,co,l] * P+col[b,k,l]  +  sum_l P-[b,co,l] * s_pre_col[b,k,l]
    // (:2005-2009), l = (oy, ox) ascending, with the UPDATED traces (recomputed here from slot `in`).
    // Only spiking positions contribute, so each sum walks the set bits of the sample's target (source) bit row
    // inside channel co (ci) in ascending order — the same terms in the same order as the dense double loop.
    // Unit = (sample, group of output channels); the sample's two bit rows are staged in shared memory.
    {
        // output channels per unit: about one (channel, tap) element per thread, and few enough channels for their
        // P- rows of the sample ([channels][L] floats) to be staged in the 32 KB accumulator region
        const bool stage_pm = L <= SNN_GEN_WARPS * 32 * 32;
        const int cpc = stage_pm ? min(max(1, SNN_GEN_THREADS / K), (SNN_GEN_WARPS * 32 * 32) / L) : max(1, SNN_GEN_THREADS / K);
        const int nch = (C.cout + cpc - 1) / cpc;
        const bool staged = S.nw + G.nw <= SNN_CONV_STAGE_WORDS;
        uint32_t *sb_s = (uint32_t *)GS.xs, *gb_s = (uint32_t *)GS.xs + S.nw;
        float *pm_s = GS.acc;
        // the sample's source spikes as an ascending list of (src << 16 | iy << 8 | ix), decoded once per unit instead
        // of once per (filter tap, spike); it shares the 16 KB region with the two bit rows
        uint32_t *slist = (uint32_t *)GS.xs + S.nw + G.nw;
        const int slist_cap = staged ? SNN_CONV_STAGE_WORDS - S.nw - G.nw - (C.cin + 2) : 0;
        int32_t *seg = (int32_t *)(slist + (slist_cap > 0 ? slist_cap : 0));   // [cin + 1] first list entry of every input channel
        const bool listed = staged && ns <= 65535 && C.hin <= 256 && C.win <= 256 && slist_cap >= ns;
        const bool unit_stride = C.sh == 1 && C.sw == 1;
        for (int u = cta; u < B * nch; u += ncta) {
            const int b = u / nch, ch = u - b * nch;
            const int co0 = ch * cpc, co1 = min(C.cout, co0 + cpc);
            const uint32_t *sbg = S.bits + ((size_t)wr * B + b) * S.nw, *gbg = G.bits + ((size_t)wr * B + b) * G.nw;
            const float *ppb = pp + (size_t)b * ns, *pmb = pm + (size_t)b * nt;
            if (staged || stage_pm) __syncthreads();
            if (staged) {
                for (int k = threadIdx.x; k < S.nw; k += SNN_GEN_THREADS) sb_s[k] = __ldcg(sbg + k);
                for (int k = threadIdx.x; k < G.nw; k += SNN_GEN_THREADS) gb_s[k] = __ldcg(gbg + k);
            }
            if (stage_pm) {   // P- of the unit's channels, eight loads in flight per thread
                const int tot = (co1 - co0) * L;
                const float *src = pmb + (size_t)co0 * L;
                for (int k0 = threadIdx.x; k0 < tot; k0 += 8 * SNN_GEN_THREADS) {
                    float pv8[8];
                    #pragma unroll
                    for (int q = 0; q < 8; ++q) pv8[q] = k0 + q * SNN_GEN_THREADS < tot ? __ldcg(src + k0 + q * SNN_GEN_THREADS) : 0.0f;
                    #pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (k0 + q * SNN_GEN_THREADS < tot) pm_s[k0 + q * SNN_GEN_THREADS] = pv8[q];
                }
            }
            if (staged || stage_pm) __syncthreads();
            const uint32_t *sb = staged ? sb_s : sbg, *gb = staged ? gb_s : gbg;
            if (listed) {
                if (warp == 0) {   // ordered compaction of the set bits, 32 words at a time
                    int n_ent = 0;
                    for (int w0 = 0; w0 < S.nw; w0 += 32) {
                        uint32_t mine = w0 + lane < S.nw ? sb_s[w0 + lane] : 0u;
                        const int cnt = __popc(mine);
                        int pre = cnt;
                        #pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int v = __shfl_up_sync(0xffffffffu, pre, o);
                            if (lane >= o) pre += v;
                        }
                        int q = n_ent + pre - cnt;
                        n_ent += __shfl_sync(0xffffffffu, pre, 31);
                        while (mine) {
                            const int src = (w0 + lane) * 32 + __ffs(mine) - 1;
                            mine &= mine - 1;
                            const int r = src % (C.hin * C.win), iy = r / C.win, ix = r - iy * C.win;
                            slist[q++] = ((uint32_t)src << 16) | ((uint32_t)iy << 8) | (uint32_t)ix;
                        }
                    }
                    __syncwarp();
                    // seg[c] = number of entries whose source index lies below channel c
                    for (int c = lane; c <= C.cin; c += 32) {
                        const uint32_t lim = (uint32_t)(c * C.hin * C.win);
                        int lo = 0, hi = n_ent;
                        while (lo < hi) { const int mid = (lo + hi) >> 1; if ((slist[mid] >> 16) < lim) lo = mid + 1; else hi = mid; }
                        seg[c] = lo;
                    }
                }
                __syncthreads();
            }
            for (int e = threadIdx.x; e < (co1 - co0) * K; e += SNN_GEN_THREADS) {
                const int co = co0 + e / K, k = e - (co - co0) * K;
                const int ci = k / KK, kk = k - ci * KK, ky = kk / C.kw, kx = kk - ky * C.kw;
                float s1 = 0.0f, s2 = 0.0f;
                BitIter it;
                // post-synaptic spikes of channel co
                it.init(gb, co * L, (co + 1) * L, staged);
                for (;;) {
                    const int tgt = it.next();
                    if (tgt < 0) break;
                    const int l = tgt - co * L, oy = l / C.wout, ox = l - oy * C.wout;
                    const int iy = oy * C.sh - C.ph + ky, ix = ox * C.sw - C.pw + kx;
                    if (iy < 0 || iy >= C.hin || ix < 0 || ix >= C.win) continue;
                    const int src = (ci * C.hin + iy) * C.win + ix;
                    const bool ss = ((staged ? sb[src >> 5] : __ldcg(sb + (src >> 5))) >> (src & 31)) & 1u;
                    s1 = s1 + mst_trace(__ldcg(ppb + src), C.p_plus_decay, C.a_plus, ss);
                }
                // pre-synaptic spikes of channel ci
                if (listed) {
                    const int e1 = seg[ci + 1];
                    for (int q = seg[ci]; q < e1; ++q) {
                        const uint32_t ent = slist[q];
                        const int ty = (int)((ent >> 8) & 255u) + C.ph - ky, tx = (int)(ent & 255u) + C.pw - kx;
                        if (ty < 0 || tx < 0) continue;
                        int oy = ty, ox = tx;
                        if (!unit_stride) {
                            oy = ty / C.sh; ox = tx / C.sw;
                            if (oy * C.sh != ty || ox * C.sw != tx) continue;
                        }
                        if (oy >= C.hout || ox >= C.wout) continue;
                        const int tgt = co * L + oy * C.wout + ox;
                        const float pv = stage_pm ? pm_s[tgt - co0 * L] : __ldcg(pmb + tgt);
                        const bool ts = (gb[tgt >> 5] >> (tgt & 31)) & 1u;
                        s2 = s2 + mst_trace(pv, C.p_minus_decay, C.a_minus, ts);
                    }
                    M.el[out][(size_t)b * NWT + (size_t)co * K + k] = s1 + s2;
                    continue;
                }
                // ... the same walk straight off the bit row (shapes the list does not cover), four trace loads in flight
                const int cbase = ci * C.hin * C.win;
                it.init(sb, cbase, cbase + C.hin * C.win, staged);
                bool done = false;
                while (!done) {
                    int tg[4];
                    #pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        tg[q] = -1;
                        if (done) continue;
                        const int src = it.next();
                        if (src < 0) { done = true; continue; }
                        const int r = src - cbase, iy = r / C.win, ix = r - iy * C.win;
                        const int ty = iy + C.ph - ky, tx = ix + C.pw - kx;
                        if (ty < 0 || tx < 0) continue;
                        const int oy = ty / C.sh, ox = tx / C.sw;
                        if (oy * C.sh != ty || ox * C.sw != tx || oy >= C.hout || ox >= C.wout) continue;
                        tg[q] = co * L + oy * C.wout + ox;
                    }
                    float pv[4];
                    #pragma unroll
                    for (int q = 0; q < 4; ++q) pv[q] = tg[q] >= 0 ? (stage_pm ? pm_s[tg[q] - co0 * L] : __ldcg(pmb + tg[q])) : 0.0f;
                    #pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (tg[q] >= 0) {
                            const bool ts = ((staged ? gb[tg[q] >> 5] : __ldcg(gb + (tg[q] >> 5))) >> (tg[q] & 31)) & 1u;
                            s2 = s2 + mst_trace(pv[q], C.p_minus_decay, C.a_minus, ts);
                        }
                }
                M.el[out][(size_t)b * NWT + (size_t)co * K + k] = s1 + s2;
            }
        }
        if (staged || stage_pm) __syncthreads();
    }
}

// Conv2dConnection.normalize (topology.py:824-837): every (out, in) filter scaled to sum `norm`
// (plain sum in ascending order; no guard against a zero sum, like the reference).
__device__ void normalize_conv_item(const snn_conn_t &C, int tile, int ntiles) {
    const int F = C.cout * C.cin, KK = C.kh * C.kw;
    for (int f = tile * SNN_GEN_THREADS + threadIdx.x; f < F; f += ntiles * SNN_GEN_THREADS) {
        float tot = 0.0f;
        for (int k = 0; k < KK; ++k) tot = tot + C.w[(size_t)f * KK + k];
        const float fac = C.norm / tot;
        for (int k = 0; k < KK; ++k) C.w[(size_t)f * KK + k] = C.w[(size_t)f * KK + k] * fac;
    }
}

// Connection masks (Network.run(..., masks=...), network.py:449 -> AbstractConnection.update, topology.py:127-131): the
// masked weights of this column tile are zero after every step's update.
__device__ void mask_tile(const snn_conn_t &C, int ns, int nt, int tile) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = tile * SNN_TILE + lane;
    if (j < nt)
        for (int i = warp; i < ns; i += SNN_GEN_WARPS) {
            const size_t k = (size_t)i * nt + j;
            if (C.mask[k]) C.w[k] = 0.0f;
        }
    __syncthreads();
}

// normalize(): Connection.normalize (topology.py:383-392) / AbstractFeature.normalize
// (topology_features.py:250-266) on one tile; row chunking as documented in snn_b200.h.
__device__ void normalize_tile(const snn_conn_t &C, int ns, int nt, int tile, float *s_part) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = tile * SNN_TILE + lane;
    const bool valid = j < nt;
    const int chunk = (ns + SNN_NORM_CHUNKS - 1) / SNN_NORM_CHUNKS;
    __syncthreads();
    for (int c = warp; c < SNN_NORM_CHUNKS; c += SNN_GEN_WARPS) {
        float part = 0.0f;
        const int i1 = min((c + 1) * chunk, ns);
        if (valid)
            for (int i = c * chunk; i < i1; ++i) {
                const float x = C.w[(size_t)i * nt + j];
                part = part + (C.norm_abs ? fabsf(x) : x);
            }
        s_part[c * 32 + lane] = part;
    }
    __syncthreads();
    if (warp == 0) {
        float tot = 0.0f;
        for (int c = 0; c < SNN_NORM_CHUNKS; ++c) tot = tot + s_part[c * 32 + lane];
        if (tot == 0.0f) tot = 1.0f;
        s_part[SNN_NORM_CHUNKS * 32 + lane] = C.norm / tot;
    }
    __syncthreads();
    const float f = s_part[SNN_NORM_CHUNKS * 32 + lane];
    if (valid)
        for (int i = warp; i < ns; i += SNN_GEN_WARPS) C.w[(size_t)i * nt + j] = C.w[(size_t)i * nt + j] * f;
    __syncthreads();
}


}  // namespace
