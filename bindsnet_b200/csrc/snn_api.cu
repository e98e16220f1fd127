// snn_api.cu — the C ABI of libsnn_b200.so (include/snn_b200.h): plan validation, workspace
// layout, tier selection and launches.  No torch types, no host synchronisation.
#include <cstdio>
#include <cstring>

#include "snn_common.cuh"

int snn_generic_launch(DevNet &N, cudaStream_t stream);
int snn_fused_dc_supported(const snn_net_t *net, const snn_run_opts_t *opts);
size_t snn_fused_dc_workspace_bytes(const snn_net_t *net, const snn_run_opts_t *opts);
int snn_fused_dc_launch(const snn_net_t *net, const snn_run_opts_t *opts, void *ws, size_t ws_bytes,
                        cudaStream_t stream, int *launches);
int snn_fused_dc2_supported(const snn_net_t *net, const snn_run_opts_t *opts);
size_t snn_fused_dc2_workspace_bytes(const snn_net_t *net, const snn_run_opts_t *opts);
int snn_fused_dc2_launch(const snn_net_t *net, const snn_run_opts_t *opts, void *ws, size_t ws_bytes,
                         cudaStream_t stream, int *launches);

static thread_local int g_last_launches = 0;

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

static int validate(const snn_net_t *net, const snn_run_opts_t *o) {
    if (!net || !o || net->abi_version != SNN_ABI_VERSION) return SNN_ERR_BAD_ARG;
    if (net->n_layers < 1 || net->n_layers > SNN_MAX_LAYERS) return SNN_ERR_BAD_ARG;
    if (net->n_conns < 0 || net->n_conns > SNN_MAX_CONNS) return SNN_ERR_BAD_ARG;
    if (o->T < 0 || o->B <= 0) return SNN_ERR_BAD_ARG;
    for (int l = 0; l < net->n_layers; ++l) {
        const snn_layer_t &L = net->layers[l];
        if (L.kind < SNN_NODE_INPUT || L.kind > SNN_NODE_MCP) return SNN_ERR_UNSUPPORTED;
        if (L.kind == SNN_NODE_CURRENT_LIF && !L.i) return SNN_ERR_BAD_ARG;
        if (L.n <= 0 || !L.s) return SNN_ERR_BAD_ARG;
        if (L.kind != SNN_NODE_INPUT && (!L.v || (!L.refrac_count && L.kind != SNN_NODE_MCP))) return SNN_ERR_BAD_ARG;
        if (L.kind == SNN_NODE_DC && !L.theta) return SNN_ERR_BAD_ARG;
        if (L.traces && !L.x) return SNN_ERR_BAD_ARG;
        if (L.sum_input && !L.summed) return SNN_ERR_BAD_ARG;
        if (L.ext && L.ext_dtype != SNN_EXT_U8 && L.ext_dtype != SNN_EXT_F32) return SNN_ERR_BAD_ARG;
    }
    if ((o->delta_w || o->delta_theta) && (o->normalize || !net->learning || o->one_step)) return SNN_ERR_UNSUPPORTED;
    for (int c = 0; c < net->n_conns; ++c) {
        const snn_conn_t &C = net->conns[c];
        if (C.src < 0 || C.src >= net->n_layers || C.tgt < 0 || C.tgt >= net->n_layers || !C.w) return SNN_ERR_BAD_ARG;
        if (net->layers[C.tgt].kind == SNN_NODE_INPUT) return SNN_ERR_UNSUPPORTED;
        if (C.rule < SNN_RULE_NONE || C.rule > SNN_RULE_MSTDPET) return SNN_ERR_UNSUPPORTED;
        if (C.kind < SNN_CONN_DENSE || C.kind > SNN_CONN_CONV2D) return SNN_ERR_UNSUPPORTED;
        if (C.kind == SNN_CONN_CONV2D) {
            const snn_layer_t &S = net->layers[C.src], &G = net->layers[C.tgt];
            if (C.cin * C.hin * C.win != S.n || C.cout * C.hout * C.wout != G.n || !C.b) return SNN_ERR_BAD_ARG;
            if (C.kh < 1 || C.kw < 1 || C.sh < 1 || C.sw < 1 || C.dh < 1 || C.dw < 1) return SNN_ERR_BAD_ARG;
            if (C.rule == SNN_RULE_MCC_POSTPRE) return SNN_ERR_UNSUPPORTED;
            if (SNN_RULE_IS_STDP(C.rule) && (C.dh != 1 || C.dw != 1)) return SNN_ERR_UNSUPPORTED;   // im2col_indices ignores dilation
        }
        if (C.rule == SNN_RULE_MSTDPET) {   // dense, batch size 1 (learning.py:2187-2249)
            if (C.kind != SNN_CONN_DENSE || o->B != 1) return SNN_ERR_UNSUPPORTED;
            if (!C.p_plus || !C.p_minus || !C.mst_spre || !C.mst_spost || !C.e_trace) return SNN_ERR_BAD_ARG;
        }
        if (C.rule == SNN_RULE_MSTDP) {
            if (!C.p_plus || !C.p_minus) return SNN_ERR_BAD_ARG;
            if (C.kind == SNN_CONN_CONV2D) { if (!C.elig || C.dh != 1 || C.dw != 1) return SNN_ERR_BAD_ARG; }
            else if (C.kind == SNN_CONN_DENSE) { if (!C.mst_spre || !C.mst_spost) return SNN_ERR_BAD_ARG; }
            else return SNN_ERR_UNSUPPORTED;
        }
        if (SNN_RULE_IS_STDP(C.rule) && (!net->layers[C.src].traces || !net->layers[C.tgt].traces))
            return SNN_ERR_BAD_ARG;
        if (C.mask && (C.kind != SNN_CONN_DENSE || SNN_RULE_IS_MSTDP(C.rule))) return SNN_ERR_UNSUPPORTED;
    }
    return SNN_OK;
}

static bool layer_needs_xpub(const snn_net_t *net, int l) {
    for (int c = 0; c < net->n_conns; ++c)
        if (net->conns[c].src == l && SNN_RULE_IS_STDP(net->conns[c].rule) && net->conns[c].kind != SNN_CONN_CONV2D) return true;
    return false;
}

// Carves the generic tier's workspace; with ws == nullptr only sizes it.
static size_t layout_generic(const snn_net_t *net, const snn_run_opts_t *o, char *ws, DevNet *N) {
    size_t off = 0;
    const size_t B = (size_t)o->B;
    if (N) N->bar = (unsigned int *)(ws + off);
    off += align_up(sizeof(unsigned int) * 96);
    int items = 0;
    for (int l = 0; l < net->n_layers; ++l) {
        const snn_layer_t &L = net->layers[l];
        const int nw = (L.n + 31) / 32;
        if (N) { N->layers[l].L = L; N->layers[l].nw = nw; N->layers[l].item0 = items; }
        items += nw;
        if (N) N->layers[l].bits = (uint32_t *)(ws + off);
        off += align_up(sizeof(uint32_t) * 2 * B * nw);
        const bool os = L.kind == SNN_NODE_DC && L.one_spike;
        if (N) N->layers[l].candbits = os ? (uint32_t *)(ws + off) : nullptr;
        if (os) off += align_up(sizeof(uint32_t) * B * nw);
        if (N) N->layers[l].keys = os ? (unsigned long long *)(ws + off) : nullptr;
        if (os) off += align_up(sizeof(unsigned long long) * 2 * B);
        const bool xp = L.traces && layer_needs_xpub(net, l);
        if (N) N->layers[l].xpub = xp ? (float *)(ws + off) : nullptr;
        if (xp) off += align_up(sizeof(float) * 2 * B * L.n);
        const bool th = L.kind == SNN_NODE_DC;   // adaptive threshold: decayed value per step parity + batch counters
        if (N) N->layers[l].thdec = th ? (float *)(ws + off) : nullptr;
        if (th) off += align_up(sizeof(float) * 2 * L.n);
        if (N) N->layers[l].thcnt = th ? (int32_t *)(ws + off) : nullptr;
        if (th) off += align_up(sizeof(int32_t) * 3 * L.n);
        bool wide_src = false;   // source of a dense connection with more than one gather block
        for (int c = 0; c < net->n_conns; ++c)
            if (net->conns[c].src == l && net->conns[c].kind != SNN_CONN_CONV2D && nw > 32) wide_src = true;
        if (N) N->layers[l].anyf = wide_src ? (uint32_t *)(ws + off) : nullptr;
        if (wide_src) off += align_up(sizeof(uint32_t) * 3 * B);
        if (N && os) N->any_one_spike = 1;
    }
    if (N) N->total_items = items;
    // second slot of every MSTDP rule's state (DevMstdp)
    for (int c = 0; c < net->n_conns; ++c) {
        const snn_conn_t &C = net->conns[c];
        if (!SNN_RULE_IS_MSTDP(C.rule)) continue;
        const size_t ns = (size_t)net->layers[C.src].n, nt = (size_t)net->layers[C.tgt].n;
        auto take = [&](size_t bytes) { char *p = ws ? ws + off : nullptr; off += align_up(bytes); return p; };
        char *pp = take(sizeof(float) * B * ns), *pm = take(sizeof(float) * B * nt);
        if (N) {
            DevMstdp &M = N->mst[c];
            M.pp[0] = C.p_plus; M.pm[0] = C.p_minus; M.pp[1] = (float *)pp; M.pm[1] = (float *)pm;
        }
        if (C.kind == SNN_CONN_CONV2D) {
            char *el = take(sizeof(float) * B * (size_t)C.cout * C.cin * C.kh * C.kw);
            if (N) { N->mst[c].el[0] = C.elig; N->mst[c].el[1] = (float *)el; }
        } else {
            char *sp = take(B * ns), *st = take(B * nt);
            if (N) { N->mst[c].sp[0] = C.mst_spre; N->mst[c].st[0] = C.mst_spost; N->mst[c].sp[1] = (uint8_t *)sp; N->mst[c].st[1] = (uint8_t *)st; }
        }
    }
    return off;
}

extern "C" {

int snn_b200_abi_version(void) { return SNN_ABI_VERSION; }

const char *snn_b200_build_info(void) {
    return "libsnn_b200 sm_100a (generic window + fused DC2015 windows v1/v2), ABI " "9" ", built " __DATE__ " " __TIME__;
}

int snn_b200_last_launch_count(void) { return g_last_launches; }

int snn_b200_select_tier(const snn_net_t *net, const snn_run_opts_t *opts) {
    if (validate(net, opts) != SNN_OK) return 0;
    if (opts->delta_w || opts->delta_theta)   // delta windows exist in the barrier kernel only
        return (opts->tier == 0 || opts->tier == 2) && snn_fused_dc_supported(net, opts) ? 2 : 0;
    if (opts->tier == 1) return 1;
    if (opts->tier == 3) return snn_fused_dc2_supported(net, opts) ? 3 : 0;
    // auto: the barrier kernel (tier 2) is the faster of the two fused kernels wherever both apply (B200, metric
    // configuration: 1.59 ms against 2.20 ms per 250-step window — DESIGN.md section 4); the column-group kernel
    // (tier 3) takes the shapes only it matches
    if (snn_fused_dc_supported(net, opts)) return 2;
    if (opts->tier == 2) return 0;
    if (snn_fused_dc2_supported(net, opts)) return 3;
    return 1;
}

size_t snn_b200_workspace_bytes(const snn_net_t *net, const snn_run_opts_t *opts) {
    if (validate(net, opts) != SNN_OK) return 0;
    size_t g = layout_generic(net, opts, nullptr, nullptr);
    size_t f = snn_fused_dc_supported(net, opts) ? snn_fused_dc_workspace_bytes(net, opts) : 0;
    size_t f2 = snn_fused_dc2_supported(net, opts) ? snn_fused_dc2_workspace_bytes(net, opts) : 0;
    if (f2 > f) f = f2;
    return g > f ? g : f;
}

int snn_b200_run_window(const snn_net_t *net, const snn_run_opts_t *opts, void *workspace, size_t workspace_bytes,
                        void *stream_) {
    g_last_launches = 0;
    int rc = validate(net, opts);
    if (rc != SNN_OK) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (opts->T == 0 && !opts->normalize) return SNN_OK;
    const int tier = snn_b200_select_tier(net, opts);
    if (tier == 0) return SNN_ERR_UNSUPPORTED;
    if (!workspace) return SNN_ERR_WORKSPACE;
    if (tier == 3) {
        if (workspace_bytes < snn_fused_dc2_workspace_bytes(net, opts)) return SNN_ERR_WORKSPACE;
        return snn_fused_dc2_launch(net, opts, workspace, workspace_bytes, stream, &g_last_launches);
    }
    if (tier == 2) {
        if (workspace_bytes < snn_fused_dc_workspace_bytes(net, opts)) return SNN_ERR_WORKSPACE;
        return snn_fused_dc_launch(net, opts, workspace, workspace_bytes, stream, &g_last_launches);
    }
    DevNet N;
    memset(&N, 0, sizeof(N));
    const size_t need = layout_generic(net, opts, (char *)workspace, &N);
    if (workspace_bytes < need) return SNN_ERR_WORKSPACE;
    N.n_layers = net->n_layers; N.n_conns = net->n_conns; N.learning = net->learning;
    N.T = opts->T; N.B = opts->B; N.normalize = opts->normalize;
    N.seed = opts->seed; N.step_offset = opts->step_offset; N.err = opts->err_flag;
    N.one_step = opts->one_step ? 1 : 0;
    for (int c = 0; c < net->n_conns; ++c) {
        N.conns[c] = net->conns[c];
        const snn_conn_t &C = net->conns[c];
        if (C.mask) N.any_mask = 1;
    }
    if (cudaMemsetAsync(N.bar, 0, sizeof(unsigned int) * 96, stream) != cudaSuccess) return SNN_ERR_CUDA;
    const int e = snn_generic_launch(N, stream);
    if (e != 0) {
        fprintf(stderr, "libsnn_b200: generic window launch failed: %s\n", cudaGetErrorString((cudaError_t)e));
        return SNN_ERR_CUDA;
    }
    g_last_launches = 1;  // the persistent window kernel (the memset node is not ours)
    return SNN_OK;
}

}  // extern "C"
