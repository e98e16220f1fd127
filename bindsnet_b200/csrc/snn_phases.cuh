// snn_phases.cuh — the per-step building blocks of the window kernels (spike gather, neuron
// update, one_spike resolution, STDP tile update, column normalisation), shared by the generic
// window kernel (snn_generic.cu) and the single-operator entry points (snn_ops.cu).
#pragma once
#include "snn_common.cuh"

namespace {


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
// materialising the [B, n_src, n_tgt] broadcast.
__device__ __forceinline__ float gather(const snn_conn_t &C, const uint32_t *__restrict__ sb, int nw_src,
                                        int n_src, int n_tgt, int j, bool valid, int lane) {
    float p = 0.0f;
    const bool shared_w = C.rule == SNN_RULE_MSTDP;
    for (int w0 = 0; w0 < nw_src; w0 += 32) {
        const uint32_t mine = (w0 + lane < nw_src) ? __ldcg(sb + w0 + lane) : 0u;
        uint32_t nz = __ballot_sync(0xffffffffu, mine != 0u);
        while (nz) {
            const int k = __ffs(nz) - 1;
            nz &= nz - 1;
            uint32_t word = __shfl_sync(0xffffffffu, mine, k);
            const int base = (w0 + k) * 32;
            while (word) {
                const int i = base + __ffs(word) - 1;
                word &= word - 1;
                // MSTDP weights are written by other CTAs (phase3_mstdp_dense): read them from L2
                if (valid && i < n_src) p = p + (shared_w ? __ldcg(C.w + (size_t)i * n_tgt + j) : C.w[(size_t)i * n_tgt + j]);
            }
        }
    }
    return p;
}

// Conv2dConnection.compute (topology.py:799-815) for one target neuron j = (co, oy, ox) of one
// sample: the sum of the filter taps whose (zero-padded) input position spiked, in ascending
// (ci, ky, kx) order, then the bias.
__device__ __forceinline__ float gather_conv(const snn_conn_t &C, const uint32_t *__restrict__ sb, int j, bool valid) {
    if (!valid) return 0.0f;
    const int L = C.hout * C.wout;
    const int co = j / L, l = j - co * L, oy = l / C.wout, ox = l - oy * C.wout;
    float p = 0.0f;
    for (int ci = 0; ci < C.cin; ++ci)
        for (int ky = 0; ky < C.kh; ++ky) {
            const int iy = oy * C.sh - C.ph + ky * C.dh;
            if (iy < 0 || iy >= C.hin) continue;
            for (int kx = 0; kx < C.kw; ++kx) {
                const int ix = ox * C.sw - C.pw + kx * C.dw;
                if (ix < 0 || ix >= C.win) continue;
                const int i = (ci * C.hin + iy) * C.win + ix;
                if ((__ldcg(sb + (i >> 5)) >> (i & 31)) & 1u) p = p + __ldcg(C.w + ((co * C.cin + ci) * C.kh + ky) * C.kw + kx);
            }
        }
    return p + C.b[co];
}

// ---------------------------------------------------------------------------------------
// phase 1
__device__ void phase1(const DevNet &N, int li, int tile, int t, float *s_red, int32_t *s_flag) {
    const DevLayer &D = N.layers[li];
    const snn_layer_t &L = D.L;
    const int B = N.B, n = L.n, nw = D.nw;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = tile * SNN_TILE + lane;
    const bool valid = j < n;
    const int rd = (t + 1) & 1, wr = t & 1;
    const bool dc = L.kind == SNN_NODE_DC;
    const bool deferred = dc && L.one_spike;  // final spikes known only after the arg-max
    bool nonbin = false;

    float theta = 0.0f;
    if (dc && valid) {
        theta = L.theta[j];
        if (L.learning) theta = theta * L.theta_decay;  // nodes.py:1078-1079
    }
    int cnt = 0;  // candidates of this column over this warp's samples

    for (int b = warp; b < B; b += SNN_GEN_WARPS) {
        const size_t k = (size_t)b * n + j;
        bool s = false;
        if (L.kind == SNN_NODE_INPUT) {
            // Input.forward (nodes.py:211-221): s = x
            float e = 0.0f;
            if (valid && L.ext) e = ld_ext(L, ((size_t)t * B + b) * n + j, nonbin);
            s = e != 0.0f;
            if (valid && L.sum_input) L.summed[k] = L.summed[k] + (s ? 1.0f : 0.0f);
        } else {
            // network.py:211-250: accumulate every incoming connection in insertion order
            float cur = 0.0f;
            bool has_in = false;
            for (int c = 0; c < N.n_conns; ++c) {
                const snn_conn_t &C = N.conns[c];
                if (C.tgt != li) continue;
                has_in = true;
                const DevLayer &S = N.layers[C.src];
                // one-step mode (network.py:393-396): sources earlier in the insertion order have already
                // produced this step's spikes (slot wr); everything else is still at step t-1 (slot rd)
                const int slot = (N.one_step && C.src < li) ? wr : rd;
                float p;
                if (C.kind == SNN_CONN_CONV2D) {
                    p = gather_conv(C, S.bits + ((size_t)slot * B + b) * S.nw, j, valid);
                } else {
                    p = gather(C, S.bits + ((size_t)slot * B + b) * S.nw, S.nw, S.L.n, n, j, valid, lane);
                    if (C.b && valid) p = p + C.b[j];
                }
                cur = cur + p;
            }
            if (valid) {
                // (one-step mode: the connection input REPLACES the external one, network.py:393-396)
                if (L.ext && !(N.one_step && has_in)) { bool nb = false; cur = cur + ld_ext(L, ((size_t)t * B + b) * n + j, nb); }
                float v = L.v[k], rc = L.refrac_count[k];
                if (L.inject_v) v = v + L.inject_v[(L.inject_per_step ? (size_t)t * n : 0) + j];  // network.py:398-404
                float xin = cur;
                if (dc) {
                    s = dc_step(L, v, rc, xin, theta);
                    if (L.has_lbound && v < L.lbound) v = L.lbound;  // nodes.py:1108-1109
                } else if (L.kind == SNN_NODE_IF) {
                    s = if_step(L, v, rc, xin);
                } else if (L.kind == SNN_NODE_CURRENT_LIF) {
                    float ic = L.i[k];
                    s = clif_step(L, v, rc, ic, xin);
                    L.i[k] = ic;
                } else {
                    s = lif_step(L, v, rc, xin);
                }
                L.v[k] = v;
                L.refrac_count[k] = rc;
                if (L.sum_input) L.summed[k] = L.summed[k] + xin;
                if (L.rec_v) L.rec_v[((size_t)t * B + b) * n + j] = v;
                cnt += s ? 1 : 0;
            }
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
            // final spikes: traces (nodes.py:96-103), clamp/unclamp (network.py:415-429), publish
            bool sf = s;
            if (valid) {
                float x = 0.0f;
                if (L.traces) {
                    x = trace_step(L.x[k], s, L.trace_decay, L.trace_scale, L.traces_additive);
                    L.x[k] = x;
                    if (D.xpub) D.xpub[((size_t)wr * B + b) * n + j] = x;
                }
                if (L.clamp && L.clamp[(L.clamp_per_step ? (size_t)t * n : 0) + j]) sf = true;
                if (L.unclamp && L.unclamp[(L.unclamp_per_step ? (size_t)t * n : 0) + j]) sf = false;
                if (t == N.T - 1) L.s[k] = sf ? 1 : 0;
                if (L.rec_s) L.rec_s[((size_t)t * B + b) * n + j] = sf ? 1 : 0;
                if (L.rec_count && sf) L.rec_count[k] += 1;
            }
            const uint32_t fw = __ballot_sync(0xffffffffu, valid && sf);
            if (lane == 0) D.bits[((size_t)wr * B + b) * nw + tile] = fw;
        }
    }

    if (dc) {
        // theta += theta_plus * sum_b s  (nodes.py:1093-1094): batch reduction is item-local
        s_red[warp * 32 + lane] = (float)cnt;
        __syncthreads();
        if (warp == 0 && valid) {
            int tot = 0;
            #pragma unroll
            for (int w = 0; w < SNN_GEN_WARPS; ++w) tot += (int)s_red[w * 32 + lane];
            if (L.learning) L.theta[j] = theta + L.theta_plus * (float)tot;
        }
        __syncthreads();
    }
    if (deferred && tile == 0) {
        // clear the key slot the NEXT step will arg-max into (last read two barriers ago)
        for (int b = threadIdx.x; b < B; b += blockDim.x) D.keys[(size_t)rd * B + b] = 0ull;
    }
    if (nonbin && N.err) atomicOr(N.err, SNN_ERR_NONBINARY);
    (void)s_flag;
}

// ---------------------------------------------------------------------------------------
// phase 2 (DiehlAndCookNodes with one_spike): keep the arg-max candidate of each sample
// (nodes.py:1097-1105), then traces / clamp / publish as in phase 1.
__device__ void phase2(const DevNet &N, int li, int tile, int t) {
    const DevLayer &D = N.layers[li];
    const snn_layer_t &L = D.L;
    const int B = N.B, n = L.n, nw = D.nw;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = tile * SNN_TILE + lane;
    const bool valid = j < n;
    const int wr = t & 1;
    for (int b = warp; b < B; b += SNN_GEN_WARPS) {
        const size_t k = (size_t)b * n + j;
        const uint32_t cand = D.candbits[(size_t)b * nw + tile];
        const unsigned long long key = __ldcg(D.keys + (size_t)wr * B + b);
        const bool s = valid && ((cand >> lane) & 1u) && key != 0ull && (uint32_t)(key & 0xffffffffull) == (uint32_t)j;
        bool sf = s;
        if (valid) {
            if (L.traces) {
                const float x = trace_step(L.x[k], s, L.trace_decay, L.trace_scale, L.traces_additive);
                L.x[k] = x;
                if (D.xpub) D.xpub[((size_t)wr * B + b) * n + j] = x;
            }
            if (L.clamp && L.clamp[(L.clamp_per_step ? (size_t)t * n : 0) + j]) sf = true;
            if (L.unclamp && L.unclamp[(L.unclamp_per_step ? (size_t)t * n : 0) + j]) sf = false;
            if (t == N.T - 1) L.s[k] = sf ? 1 : 0;
            if (L.rec_s) L.rec_s[((size_t)t * B + b) * n + j] = sf ? 1 : 0;
            if (L.rec_count && sf) L.rec_count[k] += 1;
        }
        const uint32_t fw = __ballot_sync(0xffffffffu, valid && sf);
        if (lane == 0) D.bits[((size_t)wr * B + b) * nw + tile] = fw;
    }
}

// ---------------------------------------------------------------------------------------
// phase 3: learning-rule update of one weight tile W[:, tile] of connection `ci`.
//   U[i,j] = reduce_b s_src[b,i] * (x_tgt[b,j] * nu0)      pre-synaptic term
//   V[i,j] = reduce_b x_src[b,i] * (s_tgt[b,j] * nu1)      post-synaptic term
// Both reduce over the batch in ascending b.  The reference materialises [B,n_src,n_tgt]
// (learning.py:399-417, MCC_learning.py:234-299); here only rows with a pre-synaptic spike
// and columns with a post-synaptic spike are touched (everything else is a bitwise no-op),
// except when a full pass is required: weight decay != 1, or the first update of the window
// (entries may sit outside [wmin, wmax] after normalize()).
__device__ void phase3(const DevNet &N, int ci, int tile, int t, float *s_acc, uint32_t *s_colmask, int32_t *s_flag) {
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

    // column events: s_colmask[g*32 + lane] = samples of group g whose target spike hit column j
    uint32_t colany = 0;
    if (post_on) {
        for (int g = warp; g < NG; g += SNN_GEN_WARPS) {
            const int b = g * 32 + lane;
            const uint32_t wb = b < B ? __ldcg(G.bits + ((size_t)wr * B + b) * nwG + tile) : 0u;
            uint32_t mine = 0;
            #pragma unroll
            for (int r = 0; r < 32; ++r) {
                const uint32_t m = __ballot_sync(0xffffffffu, (wb >> r) & 1u);
                if (lane == r) mine = m;
            }
            s_colmask[g * 32 + lane] = mine;
        }
        __syncthreads();
        for (int g = 0; g < NG; ++g) colany |= s_colmask[g * 32 + lane];
    }
    const bool any_col = __syncthreads_or(colany != 0u) != 0;
    if (!full && !pre_on && !any_col) return;

    float *acc = s_acc + warp * (32 * 32);
    for (int r = 0; r < 32; ++r) acc[r * 32 + lane] = 0.0f;
    __syncwarp();

    for (int wg = warp; wg < nwS; wg += SNN_GEN_WARPS) {
        uint32_t tmask = 0;
        if (pre_on) {
            for (int g = 0; g < NG; ++g) {
                const int bl = g * 32 + lane;
                const uint32_t mine = bl < B ? __ldcg(S.bits + ((size_t)wr * B + bl) * nwS + wg) : 0u;
                uint32_t nz = __ballot_sync(0xffffffffu, mine != 0u);
                while (nz) {
                    const int bb = __ffs(nz) - 1;
                    nz &= nz - 1;
                    uint32_t word = __shfl_sync(0xffffffffu, mine, bb);
                    const int b = g * 32 + bb;
                    float tx = 0.0f;
                    if (valid) {
                        tx = G.L.x[(size_t)b * nt + j];
                        if (!wdep) tx = tx * C.nu0;
                    }
                    tmask |= word;
                    while (word) {
                        const int r = __ffs(word) - 1;
                        word &= word - 1;
                        acc[r * 32 + lane] = acc[r * 32 + lane] + tx;
                    }
                }
            }
        }
        if (!full && !tmask && !any_col) continue;
        for (int r = 0; r < 32; ++r) {
            const int i = wg * 32 + r;
            if (i >= ns) break;
            const bool pre_t = (tmask >> r) & 1u;
            const bool need = valid && (full || pre_t || colany != 0u);
            if (!__any_sync(0xffffffffu, need)) continue;
            if (need) {
                float U = 0.0f, V = 0.0f;
                if (pre_t) {
                    U = acc[r * 32 + lane];
                    if (C.reduction == SNN_REDUCE_MEAN) U = U / Bf;
                }
                const bool post_t = colany != 0u;
                if (post_t) {
                    for (int g = 0; g < NG; ++g) {
                        uint32_t m = s_colmask[g * 32 + lane];
                        while (m) {
                            const int b = g * 32 + __ffs(m) - 1;
                            m &= m - 1;
                            const float xs = __ldcg(S.xpub + ((size_t)wr * B + b) * ns + i);
                            V = V + xs * (wdep ? 1.0f : C.nu1);
                        }
                    }
                    if (C.reduction == SNN_REDUCE_MEAN) V = V / Bf;
                }
                float *wp = C.w + (size_t)i * nt + j;
                *wp = apply_rule(C, *wp, U, pre_t, V, post_t);
            }
            if (pre_t) acc[r * 32 + lane] = 0.0f;
        }
        __syncwarp();
    }
    __syncthreads();
    (void)s_flag;
}

// ---------------------------------------------------------------------------------------
// phase 3 for reward-modulated STDP and for convolutional connections.  The rule's state is double
// buffered (DevMstdp): everything is READ from slot `in` and WRITTEN to slot `out`, so no thread
// overwrites a value another thread still needs in this step; the work is spread over the items
// (32-neuron tiles) of the connection's target layer with grid-stride loops.
__device__ __forceinline__ float mst_trace(float p, float decay, float a, bool s) {
    // learning.py:1564-1567 / 1999-2003:  P *= exp(-dt/tc);  P += a * s
    const float x = p * decay;
    return x + a * (s ? 1.0f : 0.0f);
}
__device__ __forceinline__ bool bit_of(const uint32_t *row, int i) { return (__ldcg(row + (i >> 5)) >> (i & 31)) & 1u; }

// learning.MSTDP._connection_update (learning.py:1504-1574) + base class decay / clamp (:87-104).
// Work is spread over the items of the SOURCE layer (a dense layer's target is often tiny — 10 output
// neurons in BASELINE config 4 — while its source has thousands of rows): item = 32 rows i (one per
// lane) x all columns j (warps stride over them); the batch sum runs in ascending b per (i, j).
__device__ void phase3_mstdp_dense(const DevNet &N, int ci_, int tile, int t) {
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
    // weight update from the eligibility of the previous step = p_plus (x) s_post + s_pre (x) p_minus
    if (i < ns)
        for (int j = warp; j < nt; j += SNN_GEN_WARPS) {
            float upd = 0.0f;
            for (int b = 0; b < B; ++b) {
                const bool ss = __ldcg(sp + (size_t)b * ns + i) != 0, tt = __ldcg(st + (size_t)b * nt + j) != 0;
                if (!ss && !tt) continue;
                const float e = __ldcg(pp + (size_t)b * ns + i) * (tt ? 1.0f : 0.0f) + (ss ? 1.0f : 0.0f) * __ldcg(pm + (size_t)b * nt + j);
                upd = upd + C.reward * e;
            }
            if (C.reduction == SNN_REDUCE_MEAN) upd = upd / Bf;
            float x = C.w[(size_t)i * nt + j] + C.nu0 * upd;
            if (C.weight_decay != 0.0f) x = x * C.weight_decay;
            if (C.has_clamp) x = clampf(x, C.wmin, C.wmax);
            C.w[(size_t)i * nt + j] = x;
        }
    // P+ and the pre-synaptic spikes of this step for my rows; tile 0 also does P- and the post side
    if (i < ns)
        for (int b = warp; b < B; b += SNN_GEN_WARPS) {
            const size_t k = (size_t)b * ns + i;
            const bool s = bit_of(S.bits + ((size_t)wr * B + b) * S.nw, i);
            M.pp[out][k] = mst_trace(__ldcg(pp + k), C.p_plus_decay, C.a_plus, s);
            M.sp[out][k] = s ? 1 : 0;
        }
    if (tile == 0)
        for (size_t k = threadIdx.x; k < (size_t)B * nt; k += SNN_GEN_THREADS) {
            const int b = (int)(k / nt), jj = (int)(k - (size_t)b * nt);
            const bool s = bit_of(G.bits + ((size_t)wr * B + b) * G.nw, jj);
            M.pm[out][k] = mst_trace(__ldcg(pm + k), C.p_minus_decay, C.a_minus, s);
            M.st[out][k] = s ? 1 : 0;
        }
}

// learning.MSTDP._conv2d_connection_update (learning.py:1942-2015) with a per-sample eligibility
// (SURVEY.md §0.8), and the decay-only update of a conv connection without a rule (learning.NoOp).
__device__ void phase3_conv(const DevNet &N, int ci_, int tile, int t) {
    const snn_conn_t &C = N.conns[ci_];
    const DevMstdp &M = N.mst[ci_];
    const DevLayer &S = N.layers[C.src], &G = N.layers[C.tgt];
    const int B = N.B, ns = S.L.n, nt = G.L.n, ntiles = G.nw;
    const int K = C.cin * C.kh * C.kw, L = C.hout * C.wout, NWT = C.cout * K;
    const size_t start = (size_t)tile * SNN_GEN_THREADS + threadIdx.x, stride = (size_t)ntiles * SNN_GEN_THREADS;
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
            const int ci = k / (C.kh * C.kw), kk = k - ci * C.kh * C.kw, ky = kk / C.kw, kx = kk - ky * C.kw;
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
            float x = C.w[e];
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
            for (size_t k = start; k < (size_t)NWT; k += stride) C.w[k] = C.w[k] * C.weight_decay;
        return;
    }
    const int in = (t + N.T) & 1, out = in ^ 1, wr = t & 1;
    const float *pp = M.pp[in], *pm = M.pm[in], *el = M.el[in];
    // w += nu0 * sum_b reward * eligibility(t-1)  (:1973-1974), then decay / clamp (learning.py:87-104)
    for (size_t k = start; k < (size_t)NWT; k += stride) {
        float upd = 0.0f;
        for (int b = 0; b < B; ++b) upd = upd + C.reward * __ldcg(el + (size_t)b * NWT + k);
        float x = C.w[k] + C.nu0 * upd;
        if (C.weight_decay != 0.0f) x = x * C.weight_decay;
        if (C.has_clamp) x = clampf(x, C.wmin, C.wmax);
        C.w[k] = x;
    }
    // P+ (trace image of the source), P- (:1999-2003)
    for (size_t k = start; k < (size_t)B * ns; k += stride) {
        const int b = (int)(k / ns), i = (int)(k - (size_t)b * ns);
        M.pp[out][k] = mst_trace(__ldcg(pp + k), C.p_plus_decay, C.a_plus, bit_of(S.bits + ((size_t)wr * B + b) * S.nw, i));
    }
    for (size_t k = start; k < (size_t)B * nt; k += stride) {
        const int b = (int)(k / nt), jj = (int)(k - (size_t)b * nt);
        M.pm[out][k] = mst_trace(__ldcg(pm + k), C.p_minus_decay, C.a_minus, bit_of(G.bits + ((size_t)wr * B + b) * G.nw, jj));
    }
    // eligibility(t)[b,co,k] = sum_l s_post[b,co,l] * P+col[b,k,l]  +  sum_l P-[b,co,l] * s_pre_col[b,k,l]
    // (:2005-2009), l = (oy, ox) ascending, with the UPDATED traces (recomputed here from slot `in`)
    for (size_t e = start; e < (size_t)B * NWT; e += stride) {
        const int b = (int)(e / NWT), r = (int)(e - (size_t)b * NWT), co = r / K, k = r - co * K;
        const int ci = k / (C.kh * C.kw), kk = k - ci * C.kh * C.kw, ky = kk / C.kw, kx = kk - ky * C.kw;
        const uint32_t *sb = S.bits + ((size_t)wr * B + b) * S.nw, *gb = G.bits + ((size_t)wr * B + b) * G.nw;
        float s1 = 0.0f, s2 = 0.0f;
        for (int oy = 0; oy < C.hout; ++oy) {
            const int iy = oy * C.sh - C.ph + ky;
            if (iy < 0 || iy >= C.hin) continue;
            for (int ox = 0; ox < C.wout; ++ox) {
                const int ix = ox * C.sw - C.pw + kx;
                if (ix < 0 || ix >= C.win) continue;
                const int src = (ci * C.hin + iy) * C.win + ix, tgt = co * L + oy * C.wout + ox;
                const bool ss = bit_of(sb, src), ts = bit_of(gb, tgt);
                if (ts) s1 = s1 + mst_trace(__ldcg(pp + (size_t)b * ns + src), C.p_plus_decay, C.a_plus, ss);
                if (ss) s2 = s2 + mst_trace(__ldcg(pm + (size_t)b * nt + tgt), C.p_minus_decay, C.a_minus, ts);
            }
        }
        M.el[out][e] = s1 + s2;
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
