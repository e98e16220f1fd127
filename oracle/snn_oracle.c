/*
 * snn_oracle.c — CPU restatement of BindsNET's Network.run() hot path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg may load this library; the product path (bindsnet_b200) never does.
 *
 * Parity pinning: the reference's own tests hold no golden vectors for this path
 * (SURVEY.md §4, §8c), so this restatement is pinned against the LIVE reference instead:
 * oracle/gen_golden.py imports /root/reference, runs it under fixed seeds and commits the
 * inputs + resulting state under tests/golden/; tests/test_oracle_golden.py replays them here.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/bindsnet/).  Arithmetic is fp32 with one rounding per reference ATen op
 * (compile with -ffp-contract=off: no FMA contraction).  Where the reference leaves a
 * summation order to ATen (sum over pre-synaptic neurons, over the batch, over rows in
 * normalize) this file fixes it to ascending index order (normalize: SNN_NORM_CHUNKS
 * chunks), which is the order the CUDA kernels use, so kernel-vs-oracle is bit-exact and
 * oracle-vs-reference is within fp32 summation-order tolerance.
 *
 * `dense` = 1 evaluates every product of the reference's dense formulation (the zeros of
 * `s.float() @ w` and of the batch-summed outer products included) — that is the costed
 * restatement used as the CPU baseline; `dense` = 0 skips exact-zero terms (bit-identical,
 * used to make large test cases fast).
 */
#include "../include/snn_b200.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    float *cur;      /* [B,n] input current of this step (network.py:240-248)            */
    uint8_t *cand;   /* [B,n] DC: threshold crossers before one_spike                    */
    int has_in;
} layer_ws_t;

typedef struct {
    float *U, *V;    /* [n_src,n_tgt] batch-reduced pre / post outer products            */
    float *tx;       /* [B,n_tgt] scratch                                                */
    uint8_t *row_t;  /* [n_src] row touched by the pre term this step                    */
    uint8_t *col_t;  /* [n_tgt] column touched by the post term this step                */
    int first_update_done;
} conn_ws_t;

static int check_plan(const snn_net_t *net, const snn_run_opts_t *o) {
    if (!net || !o || net->abi_version != SNN_ABI_VERSION) return SNN_ERR_BAD_ARG;
    if (o->delta_w || o->delta_theta) return SNN_ERR_UNSUPPORTED;   /* a device-side shortcut of the multi-GPU combine */
    if (net->n_layers < 0 || net->n_layers > SNN_MAX_LAYERS) return SNN_ERR_BAD_ARG;
    if (net->n_conns < 0 || net->n_conns > SNN_MAX_CONNS) return SNN_ERR_BAD_ARG;
    if (o->T < 0 || o->B <= 0) return SNN_ERR_BAD_ARG;
    for (int l = 0; l < net->n_layers; ++l) {
        const snn_layer_t *L = &net->layers[l];
        if (L->n <= 0 || !L->s) return SNN_ERR_BAD_ARG;
        if (L->kind != SNN_NODE_INPUT && (!L->v || (!L->refrac_count && L->kind != SNN_NODE_MCP))) return SNN_ERR_BAD_ARG;
        if (L->kind == SNN_NODE_DC && !L->theta) return SNN_ERR_BAD_ARG;
        if (L->traces && !L->x) return SNN_ERR_BAD_ARG;
        if (L->sum_input && !L->summed) return SNN_ERR_BAD_ARG;
        if (L->kind < 0 || L->kind > SNN_NODE_MCP) return SNN_ERR_UNSUPPORTED;
        if (L->kind == SNN_NODE_CURRENT_LIF && !L->i) return SNN_ERR_BAD_ARG;
    }
    for (int c = 0; c < net->n_conns; ++c) {
        const snn_conn_t *C = &net->conns[c];
        if (C->src < 0 || C->src >= net->n_layers || C->tgt < 0 || C->tgt >= net->n_layers) return SNN_ERR_BAD_ARG;
        if (!C->w) return SNN_ERR_BAD_ARG;
        if (net->layers[C->tgt].kind == SNN_NODE_INPUT) return SNN_ERR_UNSUPPORTED;
        if (C->rule < 0 || C->rule > SNN_RULE_MSTDPET) return SNN_ERR_UNSUPPORTED;
        if (C->kind < 0 || C->kind > SNN_CONN_CONV2D) return SNN_ERR_UNSUPPORTED;
        if (C->kind == SNN_CONN_CONV2D) {
            const snn_layer_t *S = &net->layers[C->src], *G = &net->layers[C->tgt];
            if (C->cin * C->hin * C->win != S->n || C->cout * C->hout * C->wout != G->n) return SNN_ERR_BAD_ARG;
            if (C->kh < 1 || C->kw < 1 || C->sh < 1 || C->sw < 1 || C->dh < 1 || C->dw < 1 || !C->b) return SNN_ERR_BAD_ARG;
            if (C->rule == SNN_RULE_MCC_POSTPRE) return SNN_ERR_UNSUPPORTED;
            if (SNN_RULE_IS_STDP(C->rule) && (C->dh != 1 || C->dw != 1)) return SNN_ERR_UNSUPPORTED; /* im2col_indices ignores dilation */
        }
        if (C->rule == SNN_RULE_MSTDPET) {
            if (C->kind != SNN_CONN_DENSE || o->B != 1) return SNN_ERR_UNSUPPORTED;
            if (!C->p_plus || !C->p_minus || !C->mst_spre || !C->mst_spost || !C->e_trace) return SNN_ERR_BAD_ARG;
        }
        if (C->rule == SNN_RULE_MSTDP) {
            if (!C->p_plus || !C->p_minus) return SNN_ERR_BAD_ARG;
            if (C->kind == SNN_CONN_CONV2D) { if (!C->elig || C->dh != 1 || C->dw != 1) return SNN_ERR_BAD_ARG; }
            else if (C->kind == SNN_CONN_DENSE) { if (!C->mst_spre || !C->mst_spost) return SNN_ERR_BAD_ARG; }
            else return SNN_ERR_UNSUPPORTED;
        }
        if (SNN_RULE_IS_STDP(C->rule)) {
            /* learning.py:373-376,597-599; MCC_learning.py:193-196: traces required */
            if (!net->layers[C->src].traces) return SNN_ERR_BAD_ARG;
            if (!net->layers[C->tgt].traces) return SNN_ERR_BAD_ARG; /* target.x is read by every STDP rule */
        }
    }
    return SNN_OK;
}

/* Connection.compute (topology.py:332-346) / MulticompartmentConnection.compute with a
 * Weight feature (topology.py:437-479, topology_features.py:633-645):
 * p[b,j] = sum_i s[b,i] * w[i,j] (+ bias), i ascending; then network.py:248 adds p into the
 * target's accumulator. */
static void conn_compute(const snn_conn_t *C, const snn_layer_t *S, int n_tgt, int B, float *cur, int dense) {
    const int ns = S->n;
#pragma omp parallel
    {
        float *p = (float *)malloc(sizeof(float) * (size_t)n_tgt);
#pragma omp for schedule(static)
        for (int b = 0; b < B; ++b) {
            const uint8_t *s = S->s + (size_t)b * ns;
            for (int j = 0; j < n_tgt; ++j) p[j] = 0.0f;
            for (int i = 0; i < ns; ++i) {
                const float sv = s[i] ? 1.0f : 0.0f;
                if (!dense && !s[i]) continue;
                const float *wr = C->w + (size_t)i * n_tgt;
                for (int j = 0; j < n_tgt; ++j) p[j] = p[j] + sv * wr[j];
            }
            float *cb = cur + (size_t)b * n_tgt;
            if (C->b && C->kind == SNN_CONN_DENSE)
                for (int j = 0; j < n_tgt; ++j) cb[j] = cb[j] + (p[j] + C->b[j]); /* topology.py:345 */
            else
                for (int j = 0; j < n_tgt; ++j) cb[j] = cb[j] + p[j];
        }
        free(p);
    }
}

/* Conv2dConnection.compute (topology.py:799-815): F.conv2d(s.float(), w, b, stride, padding,
 * dilation) on zero-padded spikes; w is [Cout,Cin,kh,kw].  Summation order (ATen leaves it to the
 * backend): ascending (ci, ky, kx), bias added last. */
static void conv_compute(const snn_conn_t *C, const snn_layer_t *S, int B, float *cur, int dense) {
    const int L = C->hout * C->wout, nt = C->cout * L, ns = S->n;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        const uint8_t *s = S->s + (size_t)b * ns;
        float *cb = cur + (size_t)b * nt;
        for (int co = 0; co < C->cout; ++co)
            for (int oy = 0; oy < C->hout; ++oy)
                for (int ox = 0; ox < C->wout; ++ox) {
                    float p = 0.0f;
                    for (int ci = 0; ci < C->cin; ++ci)
                        for (int ky = 0; ky < C->kh; ++ky) {
                            const int iy = oy * C->sh - C->ph + ky * C->dh;
                            if (iy < 0 || iy >= C->hin) continue;
                            for (int kx = 0; kx < C->kw; ++kx) {
                                const int ix = ox * C->sw - C->pw + kx * C->dw;
                                if (ix < 0 || ix >= C->win) continue;
                                const uint8_t sv = s[((size_t)ci * C->hin + iy) * C->win + ix];
                                if (!dense && !sv) continue;
                                p = p + (sv ? 1.0f : 0.0f) * C->w[(((size_t)co * C->cin + ci) * C->kh + ky) * C->kw + kx];
                            }
                        }
                    const size_t j = ((size_t)co * C->hout + oy) * C->wout + ox;
                    cb[j] = cb[j] + (p + C->b[co]);
                }
    }
}

/* Nodes.forward: spike trace + summed input (nodes.py:96-107). */
static inline void trace_and_sum(const snn_layer_t *L, size_t k, int s, float xin) {
    if (L->traces) {
        float x = L->x[k] * L->trace_decay;                 /* nodes.py:98  */
        if (L->traces_additive) x = x + L->trace_scale * (s ? 1.0f : 0.0f); /* nodes.py:101 */
        else if (s) x = L->trace_scale;                     /* nodes.py:103 */
        L->x[k] = x;
    }
    if (L->sum_input) L->summed[k] = L->summed[k] + xin;    /* nodes.py:107 */
}

static void layer_forward(const snn_net_t *net, int l, const snn_run_opts_t *o, int t, layer_ws_t *ws, int *err) {
    const snn_layer_t *L = &net->layers[l];
    const int B = o->B, n = L->n;
    const size_t BN = (size_t)B * n;
    float *cur = ws->cur;

    /* network.py:388-392: external input of this step */
    if (L->kind == SNN_NODE_INPUT) {
        /* Input.forward (nodes.py:211-221): s = x */
        for (size_t k = 0; k < BN; ++k) {
            int s = 0;
            if (L->ext_dtype == SNN_EXT_U8) {
                uint8_t e = ((const uint8_t *)L->ext)[(size_t)t * BN + k];
                s = e != 0; if (e > 1) *err |= SNN_ERR_NONBINARY;
            } else if (L->ext_dtype == SNN_EXT_F32) {
                float e = ((const float *)L->ext)[(size_t)t * BN + k];
                s = e != 0.0f; if (e != 0.0f && e != 1.0f) *err |= SNN_ERR_NONBINARY;
            }
            L->s[k] = (uint8_t)s;
            trace_and_sum(L, k, s, s ? 1.0f : 0.0f);
        }
    } else {
        if (!ws->has_in) memset(cur, 0, sizeof(float) * BN); /* network.py:408-413 */
        /* one_step mode (network.py:393-396): `current_inputs.update(self._get_inputs(layers=[l]))` REPLACES the
         * entry that held the external input whenever the layer has an incoming connection */
        const int drop_ext = o->one_step && ws->has_in;
        if (drop_ext) {
        } else if (L->ext_dtype == SNN_EXT_U8) {
            const uint8_t *e = (const uint8_t *)L->ext + (size_t)t * BN;
            for (size_t k = 0; k < BN; ++k) cur[k] = cur[k] + (float)e[k];
        } else if (L->ext_dtype == SNN_EXT_F32) {
            const float *e = (const float *)L->ext + (size_t)t * BN;
            for (size_t k = 0; k < BN; ++k) cur[k] = cur[k] + e[k];
        }
        /* network.py:398-404: voltage injection before forward */
        if (L->inject_v) {
            const float *iv = L->inject_v + (L->inject_per_step ? (size_t)t * n : 0);
            for (int b = 0; b < B; ++b)
                for (int j = 0; j < n; ++j) L->v[(size_t)b * n + j] += iv[j];
        }
        if (L->kind == SNN_NODE_LIF) {
            /* LIFNodes.forward (nodes.py:500-529) */
            for (size_t k = 0; k < BN; ++k) {
                float v = L->decay * (L->v[k] - L->rest) + L->rest;   /* :508 */
                float xin = cur[k];
                if (L->refrac_count[k] > 0.0f) xin = 0.0f;            /* :511 (in place on x) */
                float rc = L->refrac_count[k] - L->dt;                /* :514 */
                v = v + xin;                                          /* :516 */
                int s = v >= L->thresh;                               /* :519 */
                if (s) { rc = L->refrac; v = L->reset; }              /* :522-523 */
                if (L->has_lbound && v < L->lbound) v = L->lbound;    /* :526-527 */
                L->v[k] = v; L->refrac_count[k] = rc; L->s[k] = (uint8_t)s;
                trace_and_sum(L, k, s, xin);                          /* :529 (masked x) */
            }
        } else if (L->kind == SNN_NODE_IF) {
            /* IFNodes.forward (nodes.py:377-394): no leak; the gate is taken BEFORE the decrement; x is not masked */
            for (size_t k = 0; k < BN; ++k) {
                const float gate = L->refrac_count[k] <= 0.0f ? 1.0f : 0.0f;
                float v = L->v[k] + gate * cur[k];                    /* :378 */
                float rc = L->refrac_count[k] - L->dt;                /* :381 */
                int s = v >= L->thresh;                               /* :384 */
                if (s) { rc = L->refrac; v = L->reset; }              /* :387-388 */
                if (L->has_lbound && v < L->lbound) v = L->lbound;    /* :391-392 */
                L->v[k] = v; L->refrac_count[k] = rc; L->s[k] = (uint8_t)s;
                trace_and_sum(L, k, s, cur[k]);                       /* :394 */
            }
        } else if (L->kind == SNN_NODE_CURRENT_LIF) {
            /* CurrentLIFNodes.forward (nodes.py:770-791): the gate is taken AFTER the decrement, on the current i */
            for (size_t k = 0; k < BN; ++k) {
                float v = L->decay * (L->v[k] - L->rest) + L->rest;   /* :770 */
                float ic = L->i[k] * L->i_decay;                      /* :771 */
                float rc = L->refrac_count[k] - L->dt;                /* :774 */
                ic = ic + cur[k];                                     /* :777 */
                const float gate = rc <= 0.0f ? 1.0f : 0.0f;
                v = v + gate * ic;                                    /* :778 */
                int s = v >= L->thresh;                               /* :781 */
                if (s) { rc = L->refrac; v = L->reset; }              /* :784-785 */
                if (L->has_lbound && v < L->lbound) v = L->lbound;    /* :788-789 */
                L->v[k] = v; L->i[k] = ic; L->refrac_count[k] = rc; L->s[k] = (uint8_t)s;
                trace_and_sum(L, k, s, cur[k]);                       /* :791 */
            }
        } else if (L->kind == SNN_NODE_BOOSTED_LIF) {
            /* BoostedLIFNodes.forward (nodes.py:620-647): no rest, no reset value, no lower bound */
            for (size_t k = 0; k < BN; ++k) {
                float v = L->v[k] * L->decay;                         /* :628 */
                float xin = cur[k];
                if (L->refrac_count[k] > 0.0f) xin = 0.0f;            /* :632 (in place on x) */
                float rc = L->refrac_count[k] - L->dt;                /* :635 */
                v = v + xin;                                          /* :638 */
                int s = v >= L->thresh;                               /* :641 */
                if (s) { rc = L->refrac; v = 0.0f; }                  /* :644-645 */
                L->v[k] = v; L->refrac_count[k] = rc; L->s[k] = (uint8_t)s;
                trace_and_sum(L, k, s, xin);                          /* :647 (masked x) */
            }
        } else if (L->kind == SNN_NODE_MCP) {
            /* McCullochPitts.forward (nodes.py:278-288) */
            for (size_t k = 0; k < BN; ++k) {
                const float v = cur[k];                               /* :285 */
                int s = v >= L->thresh;                               /* :286 */
                L->v[k] = v; L->s[k] = (uint8_t)s;
                trace_and_sum(L, k, s, cur[k]);                       /* :288 */
            }
        } else { /* SNN_NODE_DC: DiehlAndCookNodes.forward (nodes.py:1069-1111) */
            uint8_t *cand = ws->cand;
            if (L->learning)
                for (int j = 0; j < n; ++j) L->theta[j] = L->theta[j] * L->theta_decay; /* :1078-1079 */
            for (size_t k = 0; k < BN; ++k) {
                const int j = (int)(k % (size_t)n);
                float v = L->decay * (L->v[k] - L->rest) + L->rest;   /* :1077 */
                const float gate = L->refrac_count[k] <= 0.0f ? 1.0f : 0.0f;
                v = v + gate * cur[k];                                /* :1082 */
                float rc = L->refrac_count[k] - L->dt;                /* :1085 */
                int s = v >= (L->thresh + L->theta[j]);               /* :1088 */
                if (s) { rc = L->refrac; v = L->reset; }              /* :1091-1092 */
                L->v[k] = v; L->refrac_count[k] = rc; cand[k] = (uint8_t)s;
            }
            if (L->learning) {                                        /* :1093-1094 */
                for (int j = 0; j < n; ++j) {
                    int cnt = 0;
                    for (int b = 0; b < B; ++b) cnt += cand[(size_t)b * n + j];
                    L->theta[j] = L->theta[j] + L->theta_plus * (float)cnt;
                }
            }
            for (int b = 0; b < B; ++b) {                             /* :1097-1105 */
                uint8_t *cb = cand + (size_t)b * n;
                uint8_t *sb = L->s + (size_t)b * n;
                if (L->one_spike) {
                    uint64_t best = 0;
                    for (int j = 0; j < n; ++j)
                        if (cb[j]) {
                            uint64_t key = snn_one_spike_key(o->seed, (uint32_t)t + o->step_offset, (uint32_t)l, (uint32_t)b, (uint32_t)j);
                            if (key > best) best = key;
                        }
                    for (int j = 0; j < n; ++j) sb[j] = 0;
                    if (best) sb[(uint32_t)(best & 0xFFFFFFFFu)] = 1;
                } else {
                    for (int j = 0; j < n; ++j) sb[j] = cb[j];
                }
            }
            for (size_t k = 0; k < BN; ++k) {
                float v = L->v[k];
                if (L->has_lbound && v < L->lbound) { v = L->lbound; L->v[k] = v; } /* :1108-1109 */
                trace_and_sum(L, k, L->s[k], cur[k]);                 /* :1111 */
            }
        }
    }
    /* network.py:415-429: clamp / unclamp after forward (traces already updated) */
    if (L->clamp) {
        const uint8_t *m = L->clamp + (L->clamp_per_step ? (size_t)t * n : 0);
        for (int b = 0; b < B; ++b)
            for (int j = 0; j < n; ++j) if (m[j]) L->s[(size_t)b * n + j] = 1;
    }
    if (L->unclamp) {
        const uint8_t *m = L->unclamp + (L->unclamp_per_step ? (size_t)t * n : 0);
        for (int b = 0; b < B; ++b)
            for (int j = 0; j < n; ++j) if (m[j]) L->s[(size_t)b * n + j] = 0;
    }
}

static inline float clampf(float w, float lo, float hi) {
    /* torch.clamp_: min(max(w, lo), hi) */
    w = w < lo ? lo : w;
    w = w > hi ? hi : w;
    return w;
}

/* LearningRule.update for one connection:
 *   learning.PostPre._connection_update           learning.py:390-420
 *   learning.WeightDependentPostPre._connection_update  learning.py:626-653
 *   MCC_learning.PostPre._connection_update       MCC_learning.py:224-302
 * followed by the base-class decay + clamp (learning.py:87-104, MCC_learning.py:86-110). */
static void conn_update(const snn_net_t *net, const snn_conn_t *C, const snn_run_opts_t *o, conn_ws_t *ws, int dense) {
    if (C->rule == SNN_RULE_NONE) return;
    const snn_layer_t *S = &net->layers[C->src], *G = &net->layers[C->tgt];
    const int B = o->B, ns = S->n, nt = G->n;
    const size_t NW = (size_t)ns * nt;
    float *w = C->w;
    const int stdp = SNN_RULE_IS_STDP(C->rule);
    const int hebb = C->rule == SNN_RULE_HEBBIAN;
    const int wdep = C->rule == SNN_RULE_WDEP_POSTPRE || hebb;   /* "raw" sums: nu applied after the reduction */
    const int pre_on = stdp && C->nu0 != 0.0f, post_on = stdp && C->nu1 != 0.0f;
    const float dts = C->rule == SNN_RULE_MCC_POSTPRE ? C->dt_scale : 1.0f;
    const int use_dt = C->rule == SNN_RULE_MCC_POSTPRE;
    const float Bf = (float)B;
    int any_col = 0;

    if (pre_on) {
        /* U[i,j] = reduce_b s_src[b,i] * (x_tgt[b,j] * nu0)   (WDEP: without nu0) */
        float *tx = ws->tx;
        for (size_t k = 0; k < (size_t)B * nt; ++k) tx[k] = wdep ? G->x[k] : G->x[k] * C->nu0;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < ns; ++i) {
            float *Ui = ws->U + (size_t)i * nt;
            int touched = 0;
            if (dense) touched = 1;
            else for (int b = 0; b < B; ++b) if (S->s[(size_t)b * ns + i]) { touched = 1; break; }
            ws->row_t[i] = (uint8_t)touched;
            if (!touched) continue;
            for (int j = 0; j < nt; ++j) Ui[j] = 0.0f;
            for (int b = 0; b < B; ++b) {
                const uint8_t sb = S->s[(size_t)b * ns + i];
                if (!dense && !sb) continue;
                const float sv = sb ? 1.0f : 0.0f;
                const float *txb = tx + (size_t)b * nt;
                for (int j = 0; j < nt; ++j) Ui[j] = Ui[j] + sv * txb[j];
            }
            if (C->reduction == SNN_REDUCE_MEAN) for (int j = 0; j < nt; ++j) Ui[j] = Ui[j] / Bf;
        }
    } else memset(ws->row_t, 0, (size_t)ns);

    if (post_on) {
        /* V[i,j] = reduce_b x_src[b,i] * (s_tgt[b,j] * nu1)   (WDEP: without nu1) */
        for (int j = 0; j < nt; ++j) {
            int touched = dense;
            if (!touched) for (int b = 0; b < B; ++b) if (G->s[(size_t)b * nt + j]) { touched = 1; break; }
            ws->col_t[j] = (uint8_t)touched;
            any_col |= touched;
        }
        if (any_col) {
#pragma omp parallel for schedule(static)
            for (int i = 0; i < ns; ++i) {
                float *Vi = ws->V + (size_t)i * nt;
                for (int j = 0; j < nt; ++j) if (ws->col_t[j]) Vi[j] = 0.0f;
                for (int b = 0; b < B; ++b) {
                    const float xs = S->x[(size_t)b * ns + i];
                    const uint8_t *sg = G->s + (size_t)b * nt;
                    for (int j = 0; j < nt; ++j) {
                        if (!ws->col_t[j]) continue;
                        if (!dense && !sg[j]) continue;
                        const float ts = wdep ? (sg[j] ? 1.0f : 0.0f) : (sg[j] ? 1.0f : 0.0f) * C->nu1;
                        Vi[j] = Vi[j] + xs * ts;
                    }
                }
                if (C->reduction == SNN_REDUCE_MEAN) for (int j = 0; j < nt; ++j) if (ws->col_t[j]) Vi[j] = Vi[j] / Bf;
            }
        }
    } else memset(ws->col_t, 0, (size_t)nt);

    /* per-entry application, in the reference's order: pre, post, decay, clamp.
     * Untouched entries are bitwise unchanged by the reference's full-tensor ops unless the
     * decay factor is not 1 or the entry is outside [wmin,wmax] (possible only before the
     * first clamp of the window, e.g. right after normalize()). */
    const int decay_on = C->weight_decay != 0.0f && C->weight_decay != 1.0f;
    const int full = dense || decay_on || (C->has_clamp && !ws->first_update_done);
    ws->first_update_done = 1;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < ns; ++i) {
        const int rt = ws->row_t[i];
        if (!full && !rt && !any_col) continue;
        float *wi = w + (size_t)i * nt;
        const float *Ui = ws->U + (size_t)i * nt, *Vi = ws->V + (size_t)i * nt;
        for (int j = 0; j < nt; ++j) {
            const int ct = ws->col_t[j];
            if (!full && !rt && !ct) continue;
            float x = wi[j];
            if (hebb) {                                                        /* learning.py:1124-1134 */
                if (pre_on) x = x + C->nu0 * (rt ? Ui[j] : 0.0f);
                if (post_on) x = x + C->nu1 * (ct ? Vi[j] : 0.0f);
            } else if (wdep) {
                float upd = 0.0f;
                if (pre_on) upd = upd - (C->nu0 * (rt ? Ui[j] : 0.0f)) * (x - C->wmin);  /* learning.py:643-644 */
                if (post_on) upd = upd + (C->nu1 * (ct ? Vi[j] : 0.0f)) * (C->wmax - x); /* learning.py:648-649 */
                x = x + upd;                                                             /* learning.py:651     */
            } else {
                if (pre_on && rt) x = use_dt ? x - Ui[j] * dts : x - Ui[j];   /* learning.py:405 / MCC:260-263 */
                if (post_on && ct) x = use_dt ? x + Vi[j] * dts : x + Vi[j];  /* learning.py:417 / MCC:296-299 */
            }
            if (C->weight_decay != 0.0f) x = x * C->weight_decay;             /* learning.py:93-94   */
            if (C->has_clamp) x = clampf(x, C->wmin, C->wmax);                /* learning.py:97-104  */
            wi[j] = x;
        }
    }
    (void)NW;
}

/* learning.MSTDP._connection_update (learning.py:1504-1574) on a dense Connection, then the base
 * class decay + clamp (learning.py:87-104).  The reference keeps eligibility[B,n_src,n_tgt] from the
 * previous step; it equals p_plus (x) s_post + s_pre (x) p_minus of that step (:1568-1572), so it is
 * rebuilt here from the rule's p_plus / p_minus (not yet updated for this step) and the spikes the
 * rule saw last (mst_spre / mst_spost).  Batch reduction in ascending b. */
static void mstdp_dense_update(const snn_net_t *net, const snn_conn_t *C, const snn_run_opts_t *o, int dense) {
    const snn_layer_t *S = &net->layers[C->src], *G = &net->layers[C->tgt];
    const int B = o->B, ns = S->n, nt = G->n;
    const float Bf = (float)B;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < ns; ++i)
        for (int j = 0; j < nt; ++j) {
            float upd = 0.0f;
            for (int b = 0; b < B; ++b) {
                const uint8_t ss = C->mst_spre[(size_t)b * ns + i], sp = C->mst_spost[(size_t)b * nt + j];
                if (!dense && !ss && !sp) continue;
                const float e = C->p_plus[(size_t)b * ns + i] * (sp ? 1.0f : 0.0f) + (ss ? 1.0f : 0.0f) * C->p_minus[(size_t)b * nt + j];
                upd = upd + C->reward * e;                                   /* learning.py:1559 */
            }
            if (C->reduction == SNN_REDUCE_MEAN) upd = upd / Bf;
            float x = C->w[(size_t)i * nt + j] + C->nu0 * upd;               /* learning.py:1562 */
            if (C->weight_decay != 0.0f) x = x * C->weight_decay;            /* learning.py:93-94 */
            if (C->has_clamp) x = clampf(x, C->wmin, C->wmax);               /* learning.py:97-104 */
            C->w[(size_t)i * nt + j] = x;
        }
    /* P+ / P- (learning.py:1564-1567), then remember the spikes of this step */
    for (size_t k = 0; k < (size_t)B * ns; ++k) {
        float x = C->p_plus[k] * C->p_plus_decay;
        C->p_plus[k] = x + C->a_plus * (S->s[k] ? 1.0f : 0.0f);
        C->mst_spre[k] = S->s[k] ? 1 : 0;
    }
    for (size_t k = 0; k < (size_t)B * nt; ++k) {
        float x = C->p_minus[k] * C->p_minus_decay;
        C->p_minus[k] = x + C->a_minus * (G->s[k] ? 1.0f : 0.0f);
        C->mst_spost[k] = G->s[k] ? 1 : 0;
    }
}

/* learning.MSTDPET._connection_update (learning.py:2187-2249) on a dense Connection, batch size 1 (the reference
 * flattens the batch into its [n] traces), then the base class decay + clamp (learning.py:87-104).  The eligibility of
 * the previous step (:2245-2247) is rebuilt from p_plus / p_minus and the spikes the rule saw last, like in
 * mstdp_dense_update; eligibility_trace is the materialised state. */
static void mstdpet_dense_update(const snn_net_t *net, const snn_conn_t *C) {
    const snn_layer_t *S = &net->layers[C->src], *G = &net->layers[C->tgt];
    const int ns = S->n, nt = G->n;
    for (int i = 0; i < ns; ++i)
        for (int j = 0; j < nt; ++j) {
            const size_t k = (size_t)i * nt + j;
            const float e = C->p_plus[i] * (C->mst_spost[j] ? 1.0f : 0.0f) + (C->mst_spre[i] ? 1.0f : 0.0f) * C->p_minus[j];
            float et = C->e_trace[k] * C->e_trace_decay;                      /* :2229 */
            et = et + e / C->tc_e_trace;                                      /* :2230 */
            C->e_trace[k] = et;
            float x = C->w[k] + C->et_coef * et;                              /* :2232-2238 */
            if (C->weight_decay != 0.0f) x = x * C->weight_decay;             /* learning.py:93-94 */
            if (C->has_clamp) x = clampf(x, C->wmin, C->wmax);                /* learning.py:97-104 */
            C->w[k] = x;
        }
    for (int i = 0; i < ns; ++i) {                                            /* :2241-2244 */
        const float x = C->p_plus[i] * C->p_plus_decay;
        C->p_plus[i] = x + C->a_plus * (S->s[i] ? 1.0f : 0.0f);
        C->mst_spre[i] = S->s[i] ? 1 : 0;
    }
    for (int j = 0; j < nt; ++j) {
        const float x = C->p_minus[j] * C->p_minus_decay;
        C->p_minus[j] = x + C->a_minus * (G->s[j] ? 1.0f : 0.0f);
        C->mst_spost[j] = G->s[j] ? 1 : 0;
    }
}

/* learning.MSTDP._conv2d_connection_update (learning.py:1942-2015) with the per-sample eligibility
 * the code intends (:1958-1961 allocate [B,*w.shape]; the final .view(w.size()) at :2013 only works for
 * B = 1, SURVEY.md §0.8: for B > 1 this is the reference with that view taken per sample).  p_plus is
 * kept as the [B,cin,hin,win] image whose unfold the reference stores.  bmm sums run over the output
 * positions l = (oy, ox) in ascending order. */
static void mstdp_conv_update(const snn_net_t *net, const snn_conn_t *C, const snn_run_opts_t *o, int dense) {
    const snn_layer_t *S = &net->layers[C->src], *G = &net->layers[C->tgt];
    const int B = o->B, ns = S->n, nt = G->n, L = C->hout * C->wout, K = C->cin * C->kh * C->kw;
    /* weight update from the previous step's eligibility (:1973-1974) */
    for (int co = 0; co < C->cout; ++co)
        for (int k = 0; k < K; ++k) {
            float upd = 0.0f;
            for (int b = 0; b < B; ++b) upd = upd + C->reward * C->elig[((size_t)b * C->cout + co) * K + k];
            C->w[(size_t)co * K + k] = C->w[(size_t)co * K + k] + C->nu0 * upd;
        }
    /* P+ / P- (:1999-2003) */
    for (size_t k = 0; k < (size_t)B * ns; ++k) {
        float x = C->p_plus[k] * C->p_plus_decay;
        C->p_plus[k] = x + C->a_plus * (S->s[k] ? 1.0f : 0.0f);
    }
    for (size_t k = 0; k < (size_t)B * nt; ++k) {
        float x = C->p_minus[k] * C->p_minus_decay;
        C->p_minus[k] = x + C->a_minus * (G->s[k] ? 1.0f : 0.0f);
    }
    /* eligibility (:2005-2009): bmm(target_s, p_plus_col^T) + bmm(p_minus, source_s_col^T) */
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b)
        for (int co = 0; co < C->cout; ++co)
            for (int ci = 0; ci < C->cin; ++ci)
                for (int ky = 0; ky < C->kh; ++ky)
                    for (int kx = 0; kx < C->kw; ++kx) {
                        float s1 = 0.0f, s2 = 0.0f;
                        for (int oy = 0; oy < C->hout; ++oy) {
                            const int iy = oy * C->sh - C->ph + ky;
                            if (iy < 0 || iy >= C->hin) continue;
                            for (int ox = 0; ox < C->wout; ++ox) {
                                const int ix = ox * C->sw - C->pw + kx;
                                if (ix < 0 || ix >= C->win) continue;
                                const size_t src = (size_t)b * ns + ((size_t)ci * C->hin + iy) * C->win + ix;
                                const size_t tgt = (size_t)b * nt + (size_t)co * L + (size_t)oy * C->wout + ox;
                                const uint8_t ts = G->s[tgt], ss = S->s[src];
                                if (dense || ts) s1 = s1 + (ts ? 1.0f : 0.0f) * C->p_plus[src];
                                if (dense || ss) s2 = s2 + C->p_minus[tgt] * (ss ? 1.0f : 0.0f);
                            }
                        }
                        const int k = (ci * C->kh + ky) * C->kw + kx;
                        C->elig[((size_t)b * C->cout + co) * K + k] = s1 + s2;
                    }
    /* base class (learning.py:87-104) */
    for (size_t k = 0; k < (size_t)C->cout * K; ++k) {
        float x = C->w[k];
        if (C->weight_decay != 0.0f) x = x * C->weight_decay;
        if (C->has_clamp) x = clampf(x, C->wmin, C->wmax);
        C->w[k] = x;
    }
}

/* PostPre / WeightDependentPostPre / Hebbian on a Conv2dConnection (learning.py:457-497, 920-975, 1348-1380) and the
 * base class decay + clamp (:87-104).  The reference correlates the im2col views with two bmm's and reduces over
 * the batch; here per filter tap (co, k): inner sum over the output positions l = (oy, ox) ascending per sample,
 * outer sum over b ascending. */
static void stdp_conv_update(const snn_net_t *net, const snn_conn_t *C, const snn_run_opts_t *o, int dense) {
    const snn_layer_t *S = &net->layers[C->src], *G = &net->layers[C->tgt];
    const int B = o->B, ns = S->n, nt = G->n, L = C->hout * C->wout, K = C->cin * C->kh * C->kw;
    const int pre_on = C->nu0 != 0.0f || C->rule == SNN_RULE_HEBBIAN, post_on = C->nu1 != 0.0f || C->rule == SNN_RULE_HEBBIAN;
    const float Bf = (float)B;
#pragma omp parallel for schedule(static)
    for (int co = 0; co < C->cout; ++co)
        for (int ci = 0; ci < C->cin; ++ci)
            for (int ky = 0; ky < C->kh; ++ky)
                for (int kx = 0; kx < C->kw; ++kx) {
                    float U = 0.0f, V = 0.0f;
                    for (int b = 0; b < B; ++b) {
                        float u1 = 0.0f, v1 = 0.0f;
                        for (int oy = 0; oy < C->hout; ++oy) {
                            const int iy = oy * C->sh - C->ph + ky;
                            if (iy < 0 || iy >= C->hin) continue;
                            for (int ox = 0; ox < C->wout; ++ox) {
                                const int ix = ox * C->sw - C->pw + kx;
                                if (ix < 0 || ix >= C->win) continue;
                                const size_t src = (size_t)b * ns + ((size_t)ci * C->hin + iy) * C->win + ix;
                                const size_t tgt = (size_t)b * nt + (size_t)co * L + (size_t)oy * C->wout + ox;
                                const uint8_t ts = G->s[tgt], ss = S->s[src];
                                if (pre_on && (dense || ss)) u1 = u1 + G->x[tgt] * (ss ? 1.0f : 0.0f);
                                if (post_on && (dense || ts)) v1 = v1 + (ts ? 1.0f : 0.0f) * S->x[src];
                            }
                        }
                        U = U + u1; V = V + v1;
                    }
                    if (C->reduction == SNN_REDUCE_MEAN) { U = U / Bf; V = V / Bf; }
                    const size_t k = (size_t)co * K + ((size_t)ci * C->kh + ky) * C->kw + kx;
                    float x = C->w[k];
                    if (C->rule == SNN_RULE_WDEP_POSTPRE) {
                        float upd = 0.0f;
                        if (pre_on) upd = upd - (C->nu0 * U) * (x - C->wmin);      /* learning.py:955-961 */
                        if (post_on) upd = upd + (C->nu1 * V) * (C->wmax - x);     /* :964-972 */
                        x = x + upd;
                    } else if (C->rule == SNN_RULE_HEBBIAN) {
                        x = x + C->nu0 * U;                                        /* learning.py:1374 */
                        x = x + C->nu1 * V;                                        /* :1378 */
                    } else {
                        if (pre_on) x = x - C->nu0 * U;                            /* learning.py:487-491 */
                        if (post_on) x = x + C->nu1 * V;                           /* :494-497 */
                    }
                    if (C->weight_decay != 0.0f) x = x * C->weight_decay;
                    if (C->has_clamp) x = clampf(x, C->wmin, C->wmax);
                    C->w[k] = x;
                }
}

/* Conv2dConnection.normalize (topology.py:824-837): every (out, in) filter is scaled to sum `norm`
 * (plain sum over kh*kw in ascending order; no guard against a zero sum, like the reference). */
static void normalize_conv(const snn_conn_t *C) {
    const int F = C->cout * C->cin, KK = C->kh * C->kw;
    for (int f = 0; f < F; ++f) {
        float tot = 0.0f;
        for (int k = 0; k < KK; ++k) tot = tot + C->w[(size_t)f * KK + k];
        const float fac = C->norm / tot;
        for (int k = 0; k < KK; ++k) C->w[(size_t)f * KK + k] = C->w[(size_t)f * KK + k] * fac;
    }
}

/* Connection.normalize (topology.py:383-392) / AbstractFeature.normalize
 * (topology_features.py:250-266). */
static void normalize_cols(float *w, int ns, int nt, int norm_abs, float norm) {
    const int chunk = (ns + SNN_NORM_CHUNKS - 1) / SNN_NORM_CHUNKS;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < nt; ++j) {
        float tot = 0.0f;
        for (int c = 0; c < SNN_NORM_CHUNKS; ++c) {
            float part = 0.0f;
            const int i1 = (c + 1) * chunk < ns ? (c + 1) * chunk : ns;
            for (int i = c * chunk; i < i1; ++i) {
                const float x = w[(size_t)i * nt + j];
                part = part + (norm_abs ? fabsf(x) : x);
            }
            tot = tot + part;
        }
        if (tot == 0.0f) tot = 1.0f;
        const float f = norm / tot;
        for (int i = 0; i < ns; ++i) w[(size_t)i * nt + j] = w[(size_t)i * nt + j] * f;
    }
}

int snn_oracle_abi_version(void) { return SNN_ABI_VERSION; }

int snn_oracle_delta_apply(float *w, const float *w0, const float *dw_sum, int32_t n_src, int32_t n_tgt,
                           int32_t has_clamp, float wmin, float wmax, int32_t has_norm, int32_t norm_abs,
                           float norm) {
    const size_t N = (size_t)n_src * n_tgt;
    for (size_t k = 0; k < N; ++k) {
        float x = w0[k] + dw_sum[k];
        if (has_clamp) x = clampf(x, wmin, wmax);
        w[k] = x;
    }
    if (has_norm) normalize_cols(w, n_src, n_tgt, norm_abs, norm);
    return SNN_OK;
}

/* Network.run (network.py:252-465): the timestep loop. */
int snn_oracle_run_window(const snn_net_t *net, const snn_run_opts_t *o, int dense, int threads) {
    int rc = check_plan(net, o);
    if (rc) return rc;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
    const int B = o->B, T = o->T;
    layer_ws_t lws[SNN_MAX_LAYERS];
    conn_ws_t cws[SNN_MAX_CONNS];
    memset(lws, 0, sizeof(lws)); memset(cws, 0, sizeof(cws));
    for (int l = 0; l < net->n_layers; ++l) {
        const size_t BN = (size_t)B * net->layers[l].n;
        lws[l].cur = (float *)calloc(BN, sizeof(float));
        lws[l].cand = (uint8_t *)calloc(BN, 1);
    }
    for (int c = 0; c < net->n_conns; ++c) {
        const snn_conn_t *C = &net->conns[c];
        const int ns = net->layers[C->src].n, nt = net->layers[C->tgt].n;
        if (SNN_RULE_IS_STDP(C->rule) && C->kind != SNN_CONN_CONV2D) {
            cws[c].U = (float *)calloc((size_t)ns * nt, sizeof(float));
            cws[c].V = (float *)calloc((size_t)ns * nt, sizeof(float));
            cws[c].tx = (float *)calloc((size_t)B * nt, sizeof(float));
        }
        cws[c].row_t = (uint8_t *)calloc((size_t)ns, 1);
        cws[c].col_t = (uint8_t *)calloc((size_t)nt, 1);
    }
    int err = 0;
    for (int t = 0; t < T; ++t) {
        /* 1. _get_inputs (network.py:211-250): currents from the PREVIOUS step's spikes */
        for (int l = 0; l < net->n_layers; ++l) lws[l].has_in = 0;
        for (int c = 0; c < net->n_conns && !o->one_step; ++c) {
            const snn_conn_t *C = &net->conns[c];
            const snn_layer_t *G = &net->layers[C->tgt];
            if (!lws[C->tgt].has_in) { memset(lws[C->tgt].cur, 0, sizeof(float) * (size_t)B * G->n); lws[C->tgt].has_in = 1; }
            if (C->kind == SNN_CONN_CONV2D) conv_compute(C, &net->layers[C->src], B, lws[C->tgt].cur, dense);
            else conn_compute(C, &net->layers[C->src], G->n, B, lws[C->tgt].cur, dense);
        }
        /* 2. layers in insertion order (network.py:386-429) */
        for (int l = 0; l < net->n_layers; ++l) {
            if (o->one_step) {
                /* one-step (feed-forward) mode, network.py:393-396: this layer's input is recomputed just
                 * before its forward, from the CURRENT spikes of its sources — layers earlier in the
                 * insertion order have already been updated this step */
                for (int c = 0; c < net->n_conns; ++c) {
                    const snn_conn_t *C = &net->conns[c];
                    if (C->tgt != l) continue;
                    const snn_layer_t *G = &net->layers[l];
                    if (!lws[l].has_in) { memset(lws[l].cur, 0, sizeof(float) * (size_t)B * G->n); lws[l].has_in = 1; }
                    if (C->kind == SNN_CONN_CONV2D) conv_compute(C, &net->layers[C->src], B, lws[l].cur, dense);
                    else conn_compute(C, &net->layers[C->src], G->n, B, lws[l].cur, dense);
                }
            }
            layer_forward(net, l, o, t, &lws[l], &err);
        }
        /* 3. connection updates in insertion order (network.py:431-454) */
        if (net->learning)
            for (int c = 0; c < net->n_conns; ++c) {
                const snn_conn_t *C = &net->conns[c];
                if (C->rule == SNN_RULE_MSTDP && C->kind == SNN_CONN_CONV2D) mstdp_conv_update(net, C, o, dense);
                else if (C->rule == SNN_RULE_MSTDP) mstdp_dense_update(net, C, o, dense);
                else if (C->rule == SNN_RULE_MSTDPET) mstdpet_dense_update(net, C);
                else if (C->kind == SNN_CONN_CONV2D && SNN_RULE_IS_STDP(C->rule)) stdp_conv_update(net, C, o, dense);
                else if (C->kind == SNN_CONN_CONV2D) {  /* learning.NoOp on a conv connection: decay only */
                    if (C->rule == SNN_RULE_NOOP && C->weight_decay != 0.0f)
                        for (size_t k = 0; k < (size_t)C->cout * C->cin * C->kh * C->kw; ++k) C->w[k] = C->w[k] * C->weight_decay;
                } else conn_update(net, C, o, &cws[c], dense);
            }
        /* connection masks (network.py:449 -> AbstractConnection.update, topology.py:127-131): after the update,
         * whether or not learning is on */
        for (int c = 0; c < net->n_conns; ++c) {
            const snn_conn_t *C = &net->conns[c];
            if (!C->mask || C->kind != SNN_CONN_DENSE) continue;
            const size_t NW = (size_t)net->layers[C->src].n * net->layers[C->tgt].n;
            for (size_t k = 0; k < NW; ++k) if (C->mask[k]) C->w[k] = 0.0f;
        }
        /* 4. monitors (network.py:460-461, monitors.py:94-111) */
        for (int l = 0; l < net->n_layers; ++l) {
            const snn_layer_t *L = &net->layers[l];
            const size_t BN = (size_t)B * L->n;
            if (L->rec_s) memcpy(L->rec_s + (size_t)t * BN, L->s, BN);
            if (L->rec_v && L->v) memcpy(L->rec_v + (size_t)t * BN, L->v, BN * sizeof(float));
            if (L->rec_count) for (size_t k = 0; k < BN; ++k) L->rec_count[k] += L->s[k] ? 1 : 0;
        }
    }
    /* network.py:464-465: normalize every connection once after the loop */
    if (o->normalize)
        for (int c = 0; c < net->n_conns; ++c) {
            const snn_conn_t *C = &net->conns[c];
            if (C->has_norm && C->kind == SNN_CONN_CONV2D) normalize_conv(C);
            else if (C->has_norm) normalize_cols(C->w, net->layers[C->src].n, net->layers[C->tgt].n, C->norm_abs, C->norm);
        }
    for (int l = 0; l < net->n_layers; ++l) { free(lws[l].cur); free(lws[l].cand); }
    for (int c = 0; c < net->n_conns; ++c) { free(cws[c].U); free(cws[c].V); free(cws[c].tx); free(cws[c].row_t); free(cws[c].col_t); }
    if (o->err_flag) *o->err_flag |= err;
    return SNN_OK;
}

/* ---- single-operator entry points (same restatements, exposed per object) ---- */

/* Connection.compute (topology.py:332-346) / MulticompartmentConnection.compute
 * (topology.py:437-479): out[b,j] = sum_i s[b,i] w[i,j] (+ b[j]). */
int snn_oracle_conn_compute(const snn_conn_t *C, int32_t n_src, int32_t n_tgt, int32_t B, const uint8_t *s, float *out) {
    if (!C || !C->w || !s || !out) return SNN_ERR_BAD_ARG;
    snn_layer_t S; memset(&S, 0, sizeof(S)); S.n = n_src; S.s = (uint8_t *)s;
    memset(out, 0, sizeof(float) * (size_t)B * n_tgt);
    if (C->kind == SNN_CONN_CONV2D) conv_compute(C, &S, B, out, 0);
    else conn_compute(C, &S, n_tgt, B, out, 0);
    return SNN_OK;
}

/* connection.update(learning=True) from the layers' current s/x (topology.py:112-139). */
int snn_oracle_conn_update(const snn_net_t *net, int32_t ci, int32_t B) {
    if (!net || ci < 0 || ci >= net->n_conns) return SNN_ERR_BAD_ARG;
    const snn_conn_t *C = &net->conns[ci];
    const int ns = net->layers[C->src].n, nt = net->layers[C->tgt].n;
    conn_ws_t ws; memset(&ws, 0, sizeof(ws));
    ws.U = (float *)calloc((size_t)ns * nt, sizeof(float));
    ws.V = (float *)calloc((size_t)ns * nt, sizeof(float));
    ws.tx = (float *)calloc((size_t)B * nt, sizeof(float));
    ws.row_t = (uint8_t *)calloc((size_t)ns, 1);
    ws.col_t = (uint8_t *)calloc((size_t)nt, 1);
    snn_run_opts_t o; memset(&o, 0, sizeof(o)); o.B = B; o.T = 1;
    conn_update(net, C, &o, &ws, 0);
    free(ws.U); free(ws.V); free(ws.tx); free(ws.row_t); free(ws.col_t);
    return SNN_OK;
}

/* Connection.normalize (topology.py:383-392) / AbstractFeature.normalize
 * (topology_features.py:250-266). */
int snn_oracle_conn_normalize(const snn_conn_t *C, int32_t n_src, int32_t n_tgt) {
    if (!C || !C->w) return SNN_ERR_BAD_ARG;
    if (C->has_norm && C->kind == SNN_CONN_CONV2D) normalize_conv(C);
    else if (C->has_norm) normalize_cols(C->w, n_src, n_tgt, C->norm_abs, C->norm);
    return SNN_OK;
}
