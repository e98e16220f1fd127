/*
 * snn_b200.h — C ABI of the B200-native SNN simulation core.
 *
 * This is the drop-in boundary for ONE path of BindsNET: the per-timestep loop of
 * `Network.run()` (reference: bindsnet/network/network.py:252-465).  The reference has no
 * native code and therefore no FFI of its own; the entry points below are what a
 * maintainer would bind from `bindsnet/network/network.py` (see INTEGRATION.md for the
 * ctypes stub).  Everything is plain C: pointers, sizes, POD structs, no torch types.
 *
 * A *window* is one call of Network.run(inputs, time=T): T timesteps over a batch of B
 * samples.  The caller describes the network as an array of layers (reference: Nodes
 * subclasses, bindsnet/network/nodes.py) and an array of connections (reference:
 * AbstractConnection subclasses, bindsnet/network/topology.py) with their learning rule
 * (bindsnet/learning/learning.py, bindsnet/learning/MCC_learning.py), in the insertion
 * order of Network.add_layer / Network.add_connection, which fixes the accumulation order
 * (network.py:225,246-248,386).
 *
 * All state pointers are the storage of the user's own tensors (layer.v, layer.x,
 * connection.w ...) and are updated IN PLACE: after the call they hold the state the
 * reference would hold after run() (network.py:380-465), including the end-of-run
 * normalize (network.py:464-465).
 *
 * The same structs are consumed by two libraries:
 *   - libsnn_b200.so   (bindsnet_b200/csrc, CUDA sm_100a; every pointer is a DEVICE pointer)
 *   - libsnn_oracle.so (oracle/, plain C test infrastructure; every pointer is a HOST pointer)
 * (and by tests/emu/libsnn_emu.so, test infrastructure: the generic kernel's CUDA sources compiled for the host on a small
 * emulation of the CUDA execution model, HOST pointers).
 */
#ifndef SNN_B200_H
#define SNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNN_ABI_VERSION 9
#define SNN_MAX_LAYERS 8
#define SNN_MAX_CONNS 12

/* ---- node kinds (reference: bindsnet/network/nodes.py) ---- */
#define SNN_NODE_INPUT 0 /* Input            nodes.py:172-228  */
#define SNN_NODE_LIF 1   /* LIFNodes         nodes.py:418-559  */
#define SNN_NODE_DC 2    /* DiehlAndCookNodes nodes.py:981-1144; with one_spike = 0 also AdaptiveLIFNodes nodes.py:829-978
                            (the same arithmetic: decay, theta decay, gated input, threshold + theta, theta += plus * sum_b s) */
#define SNN_NODE_IF 3         /* IFNodes          nodes.py:308-415: no leak, gate taken before the refractory decrement */
#define SNN_NODE_CURRENT_LIF 4 /* CurrentLIFNodes nodes.py:681-826: decaying synaptic current i, gate taken after the decrement */
#define SNN_NODE_BOOSTED_LIF 5 /* BoostedLIFNodes nodes.py:562-678: LIF without rest / reset / lbound: v *= decay, reset to 0 */
#define SNN_NODE_MCP 6         /* McCullochPitts  nodes.py:231-305: v = x, s = v >= thresh; no refractory state (refrac_count NULL) */

/* ---- connection kinds (reference: bindsnet/network/topology.py) ---- */
#define SNN_CONN_DENSE 0 /* Connection: s.float() @ w + b                topology.py:332-346 */
#define SNN_CONN_MCC 1   /* MulticompartmentConnection[Weight]: sum_i W*s  topology.py:437-479,
                            topology_features.py:633-645 (same maths, dt-scaled STDP)          */
#define SNN_CONN_CONV2D 2 /* Conv2dConnection: F.conv2d(s.float(), w, b, stride, padding, dilation)
                             topology.py:799-815; w is [Cout, Cin, kh, kw], b is [Cout]; the source layer's
                             neurons are indexed (ci, y, x), the target's (co, oy, ox), row-major */

/* ---- learning rules ---- */
#define SNN_RULE_NONE 0        /* MCC_learning.NoOp: update() does nothing    MCC_learning.py:120-146 */
#define SNN_RULE_NOOP 1        /* learning.NoOp: weight decay only, no clamp  learning.py:107-146     */
#define SNN_RULE_POSTPRE 2     /* learning.PostPre._connection_update         learning.py:390-420     */
#define SNN_RULE_WDEP_POSTPRE 3/* learning.WeightDependentPostPre             learning.py:626-653     */
#define SNN_RULE_MCC_POSTPRE 4 /* MCC_learning.PostPre._connection_update     MCC_learning.py:224-302 */
#define SNN_RULE_MSTDP 5       /* learning.MSTDP: reward-modulated STDP; _connection_update learning.py:1504-1574
                                  on SNN_CONN_DENSE, _conv2d_connection_update :1942-2015 on SNN_CONN_CONV2D */
#define SNN_RULE_HEBBIAN 6     /* learning.Hebbian: both terms positive, nu applied after the batch reduction;
                                  _connection_update learning.py:1110-1136, _conv2d_connection_update :1348-1380 */
/* On SNN_CONN_CONV2D the rules SNN_RULE_POSTPRE (learning.py:457-497), SNN_RULE_WDEP_POSTPRE (:920-975) and
 * SNN_RULE_HEBBIAN correlate the im2col views:  pre[co,k] = reduce_b sum_l x_tgt[b,co,l] * s_src_col[b,k,l],
 * post[co,k] = reduce_b sum_l s_tgt[b,co,l] * x_src_col[b,k,l]  (dilation 1), nu applied after the reduction. */
#define SNN_RULE_MSTDPET 7     /* learning.MSTDPET on a dense Connection (learning.py:2187-2249): reward-modulated STDP with an
                                  eligibility TRACE; batch size 1 only (the reference flattens the spikes of the whole batch into
                                  its [n] traces) */
#define SNN_RULE_IS_MSTDP(r) ((r) == SNN_RULE_MSTDP || (r) == SNN_RULE_MSTDPET)
#define SNN_RULE_IS_STDP(r) (((r) >= SNN_RULE_POSTPRE && (r) <= SNN_RULE_MCC_POSTPRE) || (r) == SNN_RULE_HEBBIAN)

/* ---- weight-matrix structure hints (DiehlAndCook2015's static exc/inh matrices,
 *      models.py:204,217-220) ---- */
#define SNN_W_DENSE 0   /* arbitrary dense matrix                                              */
#define SNN_W_DIAG 1    /* square, w[i][i] = structure_val, 0 elsewhere                        */
#define SNN_W_OFFDIAG 2 /* square, w[i][i] = 0, structure_val elsewhere                        */

/* ---- batch reduction of the STDP update (learning.py:76-80) ---- */
#define SNN_REDUCE_SUM 0  /* torch.sum; also torch.squeeze when B == 1 */
#define SNN_REDUCE_MEAN 1 /* torch.mean */

/* ---- external input dtype ---- */
#define SNN_EXT_NONE 0
#define SNN_EXT_U8 1  /* uint8 / bool, one byte per element */
#define SNN_EXT_F32 2 /* float32 */

/* ---- status codes (snn_*_run_window return value and *err_flag bits) ---- */
#define SNN_OK 0
#define SNN_ERR_BAD_ARG 1        /* malformed plan (sizes, indices, NULLs)                        */
#define SNN_ERR_UNSUPPORTED 2    /* valid reference configuration this build does not implement   */
#define SNN_ERR_WORKSPACE 4      /* workspace too small                                           */
#define SNN_ERR_CUDA 8           /* a CUDA runtime call failed                                    */
#define SNN_ERR_NONBINARY 16     /* (device flag) an Input layer received a value outside {0,1}   */
#define SNN_ERR_BARRIER 32       /* (device flag) grid barrier timed out — kernel bailed out      */
#define SNN_ERR_STRUCTURE 64     /* (device flag) a weight matrix does not have the structure its SNN_W_* hint claims */

/* One population of neurons.  Reference: Nodes.__init__ nodes.py:15-86 + subclass ctor. */
typedef struct snn_layer {
    int32_t kind;            /* SNN_NODE_*                                                 */
    int32_t n;               /* neurons per sample                                         */
    int32_t traces;          /* Nodes.traces                                               */
    int32_t traces_additive; /* Nodes.traces_additive                                      */
    int32_t sum_input;       /* Nodes.sum_input                                            */
    int32_t learning;        /* layer.learning (gates theta adaptation, nodes.py:1078,1093) */
    int32_t one_spike;       /* DiehlAndCookNodes.one_spike (nodes.py:1097-1105)            */
    int32_t has_lbound;      /* lbound is not None                                         */
    float dt;                /* layer.dt (nodes.py:127)                                    */
    float trace_decay;       /* exp(-dt/tc_trace), evaluated in fp32 like nodes.py:129-131 */
    float trace_scale;
    float decay;             /* exp(-dt/tc_decay)  nodes.py:546-548,1128-1130              */
    float rest, reset, thresh, refrac, lbound;
    float theta_plus;        /* DC only */
    float theta_decay;       /* DC only: exp(-dt/tc_theta_decay) nodes.py:1131-1133        */
    int32_t ext_dtype;       /* SNN_EXT_*: dtype of `ext`                                  */
    int32_t clamp_per_step;  /* 0: clamp is [n]; 1: clamp is [T,n]   (network.py:416-421)   */
    int32_t unclamp_per_step;
    int32_t inject_per_step;
    /* --- state, updated in place; shapes as in the reference --- */
    uint8_t *s;          /* [B,n] 0/1.  in: spikes of step -1, out: spikes of step T-1      */
    float *v;            /* [B,n]  (LIF, DC)                                                */
    float *refrac_count; /* [B,n]  (LIF, DC)                                                */
    float *x;            /* [B,n]  if traces                                                */
    float *theta;        /* [n]    (DC) — shared across the batch, nodes.py:1061            */
    float *summed;       /* [B,n]  if sum_input                                             */
    /* --- per-window inputs --- */
    const void *ext;        /* [T,B,n] external input (network.py:388-392) or NULL          */
    const uint8_t *clamp;   /* bool mask, force s=1 after forward (network.py:415-421)      */
    const uint8_t *unclamp; /* bool mask, force s=0 after forward (network.py:423-429)      */
    const float *inject_v;  /* added to v before forward (network.py:398-404)               */
    /* --- per-window recordings (Monitor, monitors.py:94-111); NULL = not recorded --- */
    uint8_t *rec_s; /* [T,B,n] */
    float *rec_v;   /* [T,B,n] */
    int32_t *rec_count; /* [B,n] += number of spikes of each neuron over the window (what the
                           reference's callers compute as spikes.sum(time), e.g.
                           examples/mnist/batch_eth_mnist.py:280-284); NULL = not counted */
    /* SNN_NODE_CURRENT_LIF */
    float *i;           /* [B,n] synaptic input current, updated in place (nodes.py:771,778) */
    float i_decay;      /* exp(-dt/tc_i_decay) (nodes.py:818-820) */
} snn_layer_t;

/* One dense synapse matrix.  Reference: Connection (topology.py:265-399) or
 * MulticompartmentConnection with a single Weight feature (topology.py:402-537,
 * topology_features.py:575-671). */
typedef struct snn_conn {
    int32_t kind;      /* SNN_CONN_*                                                        */
    int32_t src, tgt;  /* indices into layers[]                                             */
    int32_t rule;      /* SNN_RULE_*                                                        */
    int32_t reduction; /* SNN_REDUCE_*                                                      */
    int32_t has_norm;  /* normalize at window end (network.py:464-465)                      */
    int32_t norm_abs;  /* 1: divide by sum_i |w| (topology.py:390-392); 0: plain sum
                          (topology_features.py:264-266)                                    */
    int32_t has_clamp; /* rule clamps w to [wmin,wmax] after each update (learning.py:97-104,
                          MCC_learning.py:101-110)                                          */
    int32_t structure; /* SNN_W_*: caller-verified structure of w (plan-time hint that lets the
                          fused kernel skip a static n x n matrix; SNN_W_DENSE is always valid) */
    float nu0, nu1;    /* pre-/post-synaptic learning rates                                 */
    float wmin, wmax;
    float weight_decay;/* multiplicative per-step factor (learning.py:85,93-94); 1.0 = off  */
    float dt_scale;    /* MCC: connection.dt factor on both STDP terms (MCC_learning.py:262,298) */
    float norm;
    float structure_val; /* the constant of SNN_W_DIAG / SNN_W_OFFDIAG                      */
    float *w;          /* [n_src, n_tgt] row-major, updated in place (CONV2D: [Cout,Cin,kh,kw]) */
    const float *b;    /* [n_tgt] bias or NULL (topology.py:345); CONV2D: [Cout]            */
    /* SNN_CONN_CONV2D geometry (topology.py:738-760): source [cin,hin,win], target [cout,hout,wout] */
    int32_t cin, hin, win, cout, hout, wout, kh, kw, sh, sw, ph, pw, dh, dw;
    /* SNN_RULE_MSTDP (learning.py:1440-1574, 1942-2015).  State of the rule, updated in place:
         DENSE : p_plus [B,n_src], p_minus [B,n_tgt]; the eligibility [B,n_src,n_tgt] of the previous
                 step is NOT materialised: it is p_plus (x) s_post + s_pre (x) p_minus of that step, so
                 the spikes the rule saw last are kept instead (mst_spre [B,n_src], mst_spost [B,n_tgt],
                 one byte per neuron);
         CONV2D: p_plus is the [B,cin,hin,win] trace image whose im2col the reference stores
                 (unfold is linear), p_minus [B,cout*hout*wout], elig [B,cout,cin*kh*kw] (fp32,
                 materialised like the reference's, applied one step later).                  */
    float reward;      /* the run's scalar reward (network.py:319-377 -> kwargs["reward"])  */
    float a_plus, a_minus;           /* defaults +1 / -1 (learning.py:1543-1556)            */
    float p_plus_decay, p_minus_decay; /* exp(-dt/tc_plus), exp(-dt/tc_minus), computed by the host in fp32 */
    float *p_plus, *p_minus, *elig;
    uint8_t *mst_spre, *mst_spost;
    /* Network.run(..., masks={(source, target): mask}) (network.py:279-280,321,449): weights whose mask byte is non-zero are
       forced to 0 after every step's update, learning or not (AbstractConnection.update, topology.py:127-131).  [n_src, n_tgt]
       bytes, SNN_CONN_DENSE only (MulticompartmentConnection.update ignores the kwarg, topology.py:509-518); NULL = none. */
    const uint8_t *mask;
    /* SNN_RULE_MSTDPET (learning.py:2187-2249), dense, B = 1.  p_plus [n_src], p_minus [n_tgt], mst_spre / mst_spost as for
       SNN_RULE_MSTDP (the eligibility of the previous step is rebuilt from them); e_trace [n_src, n_tgt] is the rule's
       eligibility_trace, updated in place:  e_trace = e_trace * e_trace_decay + eligibility / tc_e_trace  (:2229-2230),
       w += et_coef * e_trace  with et_coef = nu[0] * dt * reward evaluated by the host in fp32 (:2232-2238). */
    float *e_trace;
    float e_trace_decay, tc_e_trace, et_coef;
} snn_conn_t;

typedef struct snn_net {
    int32_t abi_version; /* SNN_ABI_VERSION */
    int32_t n_layers;
    int32_t n_conns;
    int32_t learning;    /* network.learning: gates connection.update() (network.py:448-450) */
    snn_layer_t layers[SNN_MAX_LAYERS];
    snn_conn_t conns[SNN_MAX_CONNS];
} snn_net_t;

typedef struct snn_run_opts {
    int32_t T;             /* timesteps = int(time / dt)  (network.py:356)                    */
    int32_t B;             /* batch size                                                      */
    int32_t normalize;     /* 1: run every connection's normalize() after the loop            */
    int32_t tier;          /* 0 = auto (fused v1 where it matches, else fused v2, else generic), 1 = force generic kernel,
                              2 = force fused DC2015 kernel v1 (grid barrier), 3 = force fused DC2015 kernel v2
                              (exchange warp, per-column-group pipelines) */
    uint32_t seed;         /* one_spike tie-break stream (see snn_one_spike_key)              */
    uint32_t step_offset;  /* added to t in the tie-break hash (lets callers split a window)  */
    int32_t *err_flag;     /* optional int32 (device memory for the CUDA lib); OR-ed with SNN_ERR_* */
    int32_t one_step;  /* Network.run(one_step=True), network.py:383-396: each layer's input is recomputed from
                          the CURRENT spikes of its sources just before its forward (generic tier only) */
    /* Multi-GPU windows (the reference has no counterpart).  When delta_w / delta_theta are set
     * the window leaves the learned weights / adaptive thresholds of the DiehlAndCook2015 graph as they were at the
     * start and writes what it would have added instead — delta_w[i*n+j] = w_end - w_start, delta_theta[j] =
     * theta_end - theta_start — straight into the caller's all-reduce buffer (no snapshot, no separate subtraction
     * pass).  Fused DC2015 kernel (tier 2) only, normalize must be 0; anything else: SNN_ERR_UNSUPPORTED. */
    float *delta_w;
    float *delta_theta;
} snn_run_opts_t;

/*
 * one_spike tie-break.  The reference draws the single winner per sample with
 * torch.multinomial over the 0/1 candidate mask (nodes.py:1097-1105), i.e. uniformly among
 * the threshold crossers.  We draw it as the arg-max over the candidates of an i.i.d. 31-bit
 * hash of (seed, step, layer, sample, neuron) — the same distribution, but computable with one
 * atomicMax per sample across the whole grid.  The key packs the hash above the neuron index so
 * that the arg-max also returns the winner.  The oracle and the golden generator (which
 * monkey-patches torch.multinomial with it) use this very definition.
 */
#ifdef __CUDACC__
#define SNN_HD __host__ __device__
#else
#define SNN_HD
#endif
static inline SNN_HD uint32_t snn_fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16; return h;
}
static inline SNN_HD uint32_t snn_one_spike_hash(uint32_t seed, uint32_t t, uint32_t layer, uint32_t b, uint32_t j) {
    uint32_t h = snn_fmix32(seed ^ (0x9E3779B9u * (t + 1u)));
    h = snn_fmix32(h + 0x85EBCA6Bu * (layer + 1u) + b);
    h = snn_fmix32(h ^ (0xC2B2AE35u * (j + 1u)));
    return h;
}
static inline SNN_HD uint64_t snn_one_spike_key(uint32_t seed, uint32_t t, uint32_t layer, uint32_t b, uint32_t j) {
    return ((uint64_t)(snn_one_spike_hash(seed, t, layer, b, j) | 0x80000000u) << 32) | (uint64_t)j;
}

/* Row-chunking of the end-of-window column sum (normalize): rows are split into
 * SNN_NORM_CHUNKS contiguous chunks of ceil(n_src/SNN_NORM_CHUNKS) rows, each summed in
 * ascending row order, and the partial sums are then added in ascending chunk order.  Both
 * libraries use this fixed order so that their results are bit-identical. */
#define SNN_NORM_CHUNKS 16

/* ------------------------------------------------------------------------------------------
 * CUDA library (libsnn_b200.so).  Every pointer inside `net`/`opts` is a device pointer on the
 * current device; `stream` is a cudaStream_t (NULL = legacy default stream).  Calls are
 * asynchronous with respect to the host; no host synchronisation happens inside.
 * ------------------------------------------------------------------------------------------ */

/* Bytes of scratch the window needs (spike bitmaps, tie-break keys, barrier words).
 * Replaces: the per-step temporaries the reference allocates inside Network.run
 * (network.py:240-242; topology.py:454; MCC_learning.py:234-299). */
size_t snn_b200_workspace_bytes(const snn_net_t *net, const snn_run_opts_t *opts);

/* Run one window.  Replaces: Network.run's timestep loop + end-of-run normalize
 * (network.py:380-465) together with everything it dispatches to: _get_inputs
 * (network.py:211-250), Nodes.forward (nodes.py:96-107,211-221,500-529,1069-1111),
 * Connection.compute / MulticompartmentConnection.compute (topology.py:332-346,437-479),
 * LearningRule.update (learning.py:87-104,390-420,626-653; MCC_learning.py:86-110,224-302),
 * Monitor.record (monitors.py:94-111) and normalize (topology.py:383-392;
 * topology_features.py:250-266).  Returns SNN_OK or an SNN_ERR_* code (host-detectable
 * errors only; device-detected errors are OR-ed into *opts->err_flag). */
int snn_b200_run_window(const snn_net_t *net, const snn_run_opts_t *opts, void *workspace,
                        size_t workspace_bytes, void *stream);

/* Which kernel tier `tier = 0` would select for this plan: 1 generic, 2 fused DC2015 (v1), 3 fused DC2015 (v2). */
int snn_b200_select_tier(const snn_net_t *net, const snn_run_opts_t *opts);

/* Number of kernel launches the last snn_b200_run_window on this thread issued. */
int snn_b200_last_launch_count(void);

/* Multi-GPU window combine (no reference equivalent: SURVEY.md §8e).  After an
 * all-reduce(sum) of dw = w_local - w0 over ranks:  w = clamp(w0 + dw_sum, wmin, wmax),
 * then, if has_norm, the column normalisation of normalize().  n_src x n_tgt row-major. */
int snn_b200_delta_prepare(const float *w, const float *w0, float *dw, size_t count, void *stream);
int snn_b200_delta_apply(float *w, const float *w0, const float *dw_sum, int32_t n_src, int32_t n_tgt,
                         int32_t has_clamp, float wmin, float wmax, int32_t has_norm, int32_t norm_abs,
                         float norm, void *stream);
/* The same combine in place for a window run with snn_run_opts_t.delta_w / delta_theta (w still holds the weights of
 * the window start): w = clamp(w + dw_sum), normalize(), theta = theta + dtheta_sum — one launch.  theta may be NULL. */
int snn_b200_delta_apply_fused(float *w, const float *dw_sum, int32_t n_src, int32_t n_tgt, int32_t has_clamp, float wmin, float wmax,
                               int32_t has_norm, int32_t norm_abs, float norm, float *theta, const float *dtheta_sum, int32_t n_theta,
                               void *stream);

/* Single-operator entry points — the reference's per-object methods, for callers that drive
 * the objects themselves instead of through Network.run:
 *   snn_b200_conn_compute   = Connection.compute / MulticompartmentConnection.compute
 *                             (topology.py:332-346,437-479): out[b,j] = sum_i s[b,i] w[i,j] (+ b[j])
 *   snn_b200_conn_update    = connection.update(learning=True) for conns[conn_index] from the
 *                             layers' CURRENT s / x (topology.py:112-139, learning.py:87-104,390-420,
 *                             626-653; MCC_learning.py:86-110,224-302)
 *   snn_b200_conn_normalize = Connection.normalize / AbstractFeature.normalize
 *                             (topology.py:383-392, topology_features.py:250-266) */
int snn_b200_conn_compute(const snn_conn_t *conn, int32_t n_src, int32_t n_tgt, int32_t B, const uint8_t *s,
                          float *out, void *stream);
int snn_b200_conn_update(const snn_net_t *net, int32_t conn_index, int32_t B, void *workspace,
                         size_t workspace_bytes, void *stream);
int snn_b200_conn_normalize(const snn_conn_t *conn, int32_t n_src, int32_t n_tgt, void *stream);

/* On-device spike encoders — the step before the hot path: the reference's callers encode on
 * the CPU and ship [time, batch, ...] uint8 spikes to the device; these write the same tensor on the device from the
 * rate image.  `out` is [T][n] uint8 (n = batch * pixels, the element order of the rate tensor), 0/1.
 *   snn_b200_encode_poisson   = bindsnet.encoding.poisson (encodings.py:99-156): rate_hz[n] in Hz, inter-spike intervals
 *                               ~ Poisson(1000 / (rate * dt)) steps, zero intervals bumped to one, rate 0 never spikes
 *   snn_b200_encode_bernoulli = bindsnet.encoding.bernoulli (encodings.py:50-96): prob[n] = max_prob * normalised datum,
 *                               one independent trial per step
 * Counter-based Philox-4x32-10 keyed by (seed, element): same distribution as the reference, not its random stream. */
int snn_b200_encode_poisson(const float *rate_hz, int32_t n, int32_t T, float dt, uint64_t seed, uint8_t *out, void *stream);
int snn_b200_encode_bernoulli(const float *prob, int32_t n, int32_t T, uint64_t seed, uint8_t *out, void *stream);

/* Label assignment and classification from per-sample spike counts — the step after the hot path.
 * `counts` is [n_samples, n_neurons] int32: what snn_layer_t.rec_count accumulates over a window, i.e. the reference's
 * `spikes.sum(1)` (evaluation.py:41,115,160) without the [n_samples, time, n_neurons] raster.
 *   snn_b200_assign_labels = bindsnet.evaluation.assign_labels (evaluation/evaluation.py:8-61): rates [n, L] updated in
 *                            place (alpha-decayed running per-class mean counts), proportions [n, L], assignments [n] (int64)
 *   snn_b200_predict       = all_activity (evaluation.py:99-136; proportions == NULL) / proportion_weighting (:139-180):
 *                            predictions [n_samples] int64 */
int snn_b200_assign_labels(const int32_t *counts, const int64_t *labels, int32_t n_samples, int32_t n_neurons, int32_t n_labels, float alpha,
                           float *rates, float *proportions, int64_t *assignments, void *stream);
int snn_b200_predict(const int32_t *counts, const int64_t *assignments, const float *proportions, int32_t n_samples, int32_t n_neurons,
                     int32_t n_labels, int64_t *predictions, void *stream);

/* Library/ABI identification. */
int snn_b200_abi_version(void);
const char *snn_b200_build_info(void);

/* ------------------------------------------------------------------------------------------
 * Oracle library (libsnn_oracle.so) — TEST INFRASTRUCTURE, host pointers, never shipped on the
 * product path.  `dense` = 1 evaluates every product of the reference's dense formulation
 * (zeros included, like `s.float() @ w` and the batch-summed outer products); 0 skips
 * exact-zero terms.  Both give bit-identical results.  `threads` <= 0 means all cores.
 * ------------------------------------------------------------------------------------------ */
int snn_oracle_run_window(const snn_net_t *net, const snn_run_opts_t *opts, int dense, int threads);
int snn_oracle_delta_apply(float *w, const float *w0, const float *dw_sum, int32_t n_src, int32_t n_tgt,
                           int32_t has_clamp, float wmin, float wmax, int32_t has_norm, int32_t norm_abs,
                           float norm);
int snn_oracle_conn_compute(const snn_conn_t *conn, int32_t n_src, int32_t n_tgt, int32_t B, const uint8_t *s,
                            float *out);
int snn_oracle_conn_update(const snn_net_t *net, int32_t conn_index, int32_t B);
int snn_oracle_conn_normalize(const snn_conn_t *conn, int32_t n_src, int32_t n_tgt);
int snn_oracle_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SNN_B200_H */
