// snn_generic.cu — generic persistent window kernel (any topology of Input / LIF /
// DiehlAndCook populations joined by dense connections).
//
// One cooperative grid iterates the whole T-step window of Network.run (reference:
// bindsnet/network/network.py:380-465) with at most two grid barriers per step and no host
// involvement.  Work is partitioned by TARGET-NEURON COLUMNS: a work item is (layer, tile of
// 32 neurons) for all B samples, one warp lane per neuron.  An item owns its neurons' state
// (v, refrac_count, x, theta) and the column tile W[:, tile] of every connection INTO its
// layer, so the spike-gather, the neuron update, the batch reductions of theta and of the
// STDP outer products, the clamp and the end-of-window normalisation are all item-local.
// The only cross-item traffic is bit-packed spikes (32 neurons per word), the one_spike
// arg-max keys and the published pre-synaptic traces.
//
// Per step t (rd = slot of s(t-1), wr = slot of s(t)):
//   phase 1  currents from bits[rd] (network.py:211-250), neuron update (nodes.py), theta,
//            candidates -> atomicMax keys (DC one_spike) or final spikes -> bits[wr], traces
//   barrier  (only if some DiehlAndCookNodes layer has one_spike)
//   phase 2  one_spike layers: resolve the winner per sample, final spikes, traces
//   barrier
//   phase 3  STDP + decay + clamp on the item's weight tiles (learning.py / MCC_learning.py); MSTDP and
//            conv connections (weights shared between items): + one barrier
// After the last step: normalize() of the item's tiles (network.py:464-465).
#include "snn_phases.cuh"

namespace {

__device__ __forceinline__ void item_of(const DevNet &N, int item, int &li, int &tile) {
    li = 0;
    #pragma unroll 1
    for (int l = 0; l < N.n_layers; ++l)
        if (item >= N.layers[l].item0) li = l;
    tile = item - N.layers[li].item0;
}

__global__ void __launch_bounds__(SNN_GEN_THREADS) snn_generic_window(const __grid_constant__ DevNet N) {
    extern __shared__ float smem[];
    float *s_acc = smem;                                            // [8 warps][32][32]
    float *s_red = s_acc + SNN_GEN_WARPS * 32 * 32;                 // [17][32]
    uint32_t *s_colmask = (uint32_t *)(s_red + (SNN_NORM_CHUNKS + 1) * 32);  // [ceil(B/32)][32]
    __shared__ int32_t s_flag;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int G = gridDim.x;

    // prologue: pack the incoming spike state s(-1) into slot 1, clear the arg-max keys
    for (int item = blockIdx.x; item < N.total_items; item += G) {
        int li, tile; item_of(N, item, li, tile);
        const DevLayer &D = N.layers[li];
        const int j = tile * SNN_TILE + lane;
        for (int b = warp; b < N.B; b += SNN_GEN_WARPS) {
            const bool s = j < D.L.n && D.L.s[(size_t)b * D.L.n + j] != 0;
            const uint32_t w = __ballot_sync(0xffffffffu, s);
            if (lane == 0) D.bits[((size_t)1 * N.B + b) * D.nw + tile] = w;
        }
        if (D.keys && tile == 0)
            for (int b = threadIdx.x; b < 2 * N.B; b += blockDim.x) D.keys[b] = 0ull;
        // MSTDP state: step t reads slot (t + T) & 1; for an odd T the first read is slot 1, so the
        // caller's tensors (slot 0) are copied there — the last step then writes slot 0
        if (N.learning && (N.T & 1))
            for (int c = 0; c < N.n_conns; ++c) {
                const snn_conn_t &C = N.conns[c];
                if (C.tgt != li || C.rule != SNN_RULE_MSTDP) continue;
                const DevMstdp &M = N.mst[c];
                const size_t ns = (size_t)N.layers[C.src].L.n, nt = (size_t)D.L.n, Bz = (size_t)N.B;
                const size_t start = (size_t)tile * SNN_GEN_THREADS + threadIdx.x, stride = (size_t)D.nw * SNN_GEN_THREADS;
                for (size_t k = start; k < Bz * ns; k += stride) { M.pp[1][k] = M.pp[0][k]; if (M.sp[0]) M.sp[1][k] = M.sp[0][k]; }
                for (size_t k = start; k < Bz * nt; k += stride) { M.pm[1][k] = M.pm[0][k]; if (M.st[0]) M.st[1][k] = M.st[0][k]; }
                if (M.el[0]) {
                    const size_t ne = Bz * (size_t)C.cout * C.cin * C.kh * C.kw;
                    for (size_t k = start; k < ne; k += stride) M.el[1][k] = M.el[0][k];
                }
            }
    }
    if (!grid_barrier(N.bar, G, N.err)) return;

    for (int t = 0; t < N.T; ++t) {
        if (N.one_step) {
            // feed-forward mode (network.py:383-396): layer by layer in insertion order, each one reading the
            // spikes its predecessors produced in THIS step — a grid barrier per layer
            for (int l = 0; l < N.n_layers; ++l) {
                const DevLayer &D = N.layers[l];
                for (int tile = blockIdx.x; tile < D.nw; tile += G) phase1(N, l, tile, t, s_red, &s_flag);
                if (D.L.kind == SNN_NODE_DC && D.L.one_spike) {
                    if (!grid_barrier(N.bar, G, N.err)) return;
                    for (int tile = blockIdx.x; tile < D.nw; tile += G) phase2(N, l, tile, t);
                }
                if (l + 1 < N.n_layers && !grid_barrier(N.bar, G, N.err)) return;
            }
        } else {
        for (int item = blockIdx.x; item < N.total_items; item += G) {
            int li, tile; item_of(N, item, li, tile);
            phase1(N, li, tile, t, s_red, &s_flag);
        }
        }
        if (N.any_one_spike && !N.one_step) {
            if (!grid_barrier(N.bar, G, N.err)) return;
            for (int item = blockIdx.x; item < N.total_items; item += G) {
                int li, tile; item_of(N, item, li, tile);
                const snn_layer_t &L = N.layers[li].L;
                if (L.kind == SNN_NODE_DC && L.one_spike) phase2(N, li, tile, t);
            }
        }
        if (!grid_barrier(N.bar, G, N.err)) return;
        if (N.learning) {
            for (int item = blockIdx.x; item < N.total_items; item += G) {
                int li, tile; item_of(N, item, li, tile);
                for (int c = 0; c < N.n_conns; ++c) {
                    const snn_conn_t &C = N.conns[c];
                    if (C.rule == SNN_RULE_NONE) continue;
                    if (C.rule == SNN_RULE_MSTDP && C.kind != SNN_CONN_CONV2D) {  // dense MSTDP: by source rows
                        if (C.src == li) phase3_mstdp_dense(N, c, tile, t);
                        continue;
                    }
                    if (C.tgt != li) continue;
                    if (C.kind == SNN_CONN_CONV2D) phase3_conv(N, c, tile, t);
                    else phase3(N, c, tile, t, s_acc, s_colmask, &s_flag);
                }
            }
            __syncthreads();
            // PostPre-family updates touch only the item's own column tile; MSTDP (spread over the source
            // rows) and conv filters are read by other CTAs in the next step's gather
            if (N.sync_after_learning && !grid_barrier(N.bar, G, N.err)) return;
        }
        if (N.any_mask) {   // connection masks apply after the update, learning or not (topology.py:127-131)
            for (int item = blockIdx.x; item < N.total_items; item += G) {
                int li, tile; item_of(N, item, li, tile);
                for (int c = 0; c < N.n_conns; ++c) {
                    const snn_conn_t &C = N.conns[c];
                    if (C.mask && C.tgt == li && C.kind == SNN_CONN_DENSE) mask_tile(C, N.layers[C.src].L.n, N.layers[li].L.n, tile);
                }
            }
        }
    }

    if (N.normalize) {
        for (int item = blockIdx.x; item < N.total_items; item += G) {
            int li, tile; item_of(N, item, li, tile);
            for (int c = 0; c < N.n_conns; ++c)
                if (N.conns[c].tgt == li && N.conns[c].has_norm) {
                    if (N.conns[c].kind == SNN_CONN_CONV2D) normalize_conv_item(N.conns[c], tile, N.layers[li].nw);
                    else normalize_tile(N.conns[c], N.layers[N.conns[c].src].L.n, N.layers[li].L.n, tile, s_red);
                }
        }
    }
}

}  // namespace

size_t snn_generic_smem_bytes(int B) {
    return sizeof(float) * (SNN_GEN_WARPS * 32 * 32 + (SNN_NORM_CHUNKS + 1) * 32) + sizeof(uint32_t) * 32 * (size_t)((B + 31) / 32);
}

// Launch the generic window.  Returns a cudaError_t cast to int.
int snn_generic_launch(const DevNet &N, cudaStream_t stream) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = snn_generic_smem_bytes(N.B);
    e = cudaFuncSetAttribute(snn_generic_window, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, snn_generic_window, SNN_GEN_THREADS, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) return (int)cudaErrorLaunchOutOfResources;
    if (per_sm > 2) per_sm = 2;
    int grid = N.total_items < sms * per_sm ? N.total_items : sms * per_sm;
    if (grid < 1) grid = 1;
    void *args[] = {(void *)&N};
    return (int)cudaLaunchCooperativeKernel((void *)snn_generic_window, dim3(grid), dim3(SNN_GEN_THREADS), args, smem, stream);
}
