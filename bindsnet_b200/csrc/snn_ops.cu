// snn_ops.cu — single-operator entry points of the C ABI (the reference's per-object methods:
// Connection.compute, connection.update, normalize) and the multi-GPU window-combine kernels.
#include "snn_phases.cuh"
#include "snn_combine.cuh"

namespace {

// out[b,j] = sum_{i: s[b,i]} w[i,j] (+ bias).  Connection.compute (topology.py:332-346).
__global__ void __launch_bounds__(SNN_GEN_THREADS) conn_compute_kernel(snn_conn_t C, int ns, int nt, int B,
                                                                        const uint8_t *__restrict__ s, float *__restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x, j = tile * SNN_TILE + lane;
    const bool valid = j < nt;
    for (int b = blockIdx.y * SNN_GEN_WARPS + warp; b < B; b += gridDim.y * SNN_GEN_WARPS) {
        float p = 0.0f;
        for (int i0 = 0; i0 < ns; i0 += 32) {
            const bool sp = (i0 + lane < ns) && s[(size_t)b * ns + i0 + lane] != 0;
            uint32_t word = __ballot_sync(0xffffffffu, sp);
            while (word) {
                const int i = i0 + __ffs(word) - 1;
                word &= word - 1;
                if (valid) p = p + C.w[(size_t)i * nt + j];
            }
        }
        if (valid) out[(size_t)b * nt + j] = C.b ? p + C.b[j] : p;
    }
}

// Conv2dConnection.compute (topology.py:799-815): out[b, co, oy, ox] = sum of the filter taps whose (zero-padded)
// input position spiked, in ascending (ci, ky, kx) order, then the bias — the window kernels' gather_conv on
// byte spikes.  Thread = one target neuron of one sample.
__global__ void __launch_bounds__(256) conv_compute_kernel(snn_conn_t C, int ns, int nt, int B, const uint8_t *__restrict__ s,
                                                           float *__restrict__ out) {
    const size_t total = (size_t)B * nt;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(k / nt), j = (int)(k - (size_t)b * nt);
        const int L = C.hout * C.wout;
        const int co = j / L, l = j - co * L, oy = l / C.wout, ox = l - oy * C.wout;
        const uint8_t *sb = s + (size_t)b * ns;
        float p = 0.0f;
        for (int ci = 0; ci < C.cin; ++ci)
            for (int ky = 0; ky < C.kh; ++ky) {
                const int iy = oy * C.sh - C.ph + ky * C.dh;
                if (iy < 0 || iy >= C.hin) continue;
                for (int kx = 0; kx < C.kw; ++kx) {
                    const int ix = ox * C.sw - C.pw + kx * C.dw;
                    if (ix < 0 || ix >= C.win) continue;
                    if (sb[(ci * C.hin + iy) * C.win + ix]) p = p + C.w[((co * C.cin + ci) * C.kh + ky) * C.kw + kx];
                }
            }
        out[k] = p + C.b[co];
    }
}

__global__ void __launch_bounds__(SNN_GEN_THREADS) conv_normalize_kernel(snn_conn_t C) { normalize_conv_item(C, blockIdx.x, gridDim.x); }

// bit-pack the CURRENT spikes of the two layers of a connection into slot 0
__global__ void pack_bits_kernel(const uint8_t *__restrict__ s, uint32_t *__restrict__ bits, int B, int n, int nw) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gw >= B * nw) return;
    const int b = gw / nw, w = gw % nw, j = w * 32 + lane;
    const bool sp = j < n && s[(size_t)b * n + j] != 0;
    const uint32_t word = __ballot_sync(0xffffffffu, sp);
    if (lane == 0) bits[(size_t)b * nw + w] = word;
}

__global__ void __launch_bounds__(SNN_GEN_THREADS) conn_update_kernel(const __grid_constant__ DevNet N, int ci) {
    SNN_DYN_SHARED(float, smem);
    const GenSmem M = gen_carve(smem, N.B);
    phase3(N, ci, blockIdx.x, 0, N.layers[N.conns[ci].src].nw, 0, M);
}

__global__ void __launch_bounds__(SNN_GEN_THREADS) conn_normalize_kernel(snn_conn_t C, int ns, int nt) {
    SNN_SHARED(float, s_part, (SNN_NORM_CHUNKS + 1) * 32);
    normalize_tile(C, ns, nt, blockIdx.x, s_part);
}

__global__ void delta_prepare_kernel(const float *__restrict__ w, const float *__restrict__ w0, float *__restrict__ dw, size_t n) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) dw[k] = w[k] - w0[k];
}

// w = clamp(w0 + sum_r dw_r), then normalize() — all on one column tile (snn_combine.cuh).
__global__ void __launch_bounds__(SNN_GEN_THREADS) delta_apply_kernel(snn_conn_t C, const float *w0, const float *__restrict__ dws, int ns, int nt,
                                                                       float *theta, const float *__restrict__ dtheta, int n_theta) {
    if (theta)   // theta = theta0 + sum_r dtheta_r, in place, spread over the grid
        for (int k = blockIdx.x * SNN_GEN_THREADS + threadIdx.x; k < n_theta; k += gridDim.x * SNN_GEN_THREADS) theta[k] = theta[k] + dtheta[k];
    SNN_SHARED(float, s_part, (SNN_NORM_CHUNKS + 1) * 32);
    delta_apply_tile(C, w0, dws, ns, nt, blockIdx.x, s_part);
}

// Checks on the device that a square matrix has the structure a plan claims for it (SNN_W_DIAG: val on the
// diagonal, 0 elsewhere; SNN_W_OFFDIAG: 0 on the diagonal, val elsewhere) — the fused kernels replace such a
// matrix by its constant, so a matrix modified behind the host-side cache must not go unnoticed.
__global__ void __launch_bounds__(256) verify_structure_kernel(const float *__restrict__ w, int n, int structure, float val, int32_t *err) {
    const size_t total = (size_t)n * n;
    const bool diag = structure == SNN_W_DIAG;
    bool bad = false;
    if ((n & 3) == 0 && (((size_t)w) & 15) == 0) {   // 16-byte loads: the matrix is read once per window (10 MB at n = 1600)
        const float4 *w4 = (const float4 *)w;
        for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total / 4; q += (size_t)gridDim.x * blockDim.x) {
            const float4 x = __ldcs(w4 + q);
            const size_t k = 4 * q, i = k / n, j = k - i * n;   // 4 | n: the four elements share row i
            const float xs[4] = {x.x, x.y, x.z, x.w};
            #pragma unroll
            for (int e = 0; e < 4; ++e) bad |= xs[e] != (((i == j + e) == diag) ? val : 0.0f);
        }
    } else {
        for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
            const size_t i = k / n, j = k - i * n;
            bad |= w[k] != (((i == j) == diag) ? val : 0.0f);
        }
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0 && err) atomicOr(err, SNN_ERR_STRUCTURE);
}

inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? SNN_OK : SNN_ERR_CUDA; }

}  // namespace

int snn_verify_structure(const snn_conn_t &C, int n, int32_t *err, cudaStream_t stream) {
    if (C.structure != SNN_W_DIAG && C.structure != SNN_W_OFFDIAG) return SNN_OK;
    const size_t total = (size_t)n * n;
    const int blocks = (int)((total + 2047) / 2048 < 1184 ? (total + 2047) / 2048 : 1184);   // 8 elements per thread, up to 8 CTAs per SM
    SNN_LAUNCH(verify_structure_kernel, blocks > 0 ? blocks : 1, 256, 0, stream, C.w, n, C.structure, C.structure_val, err);
    return cuda_rc(cudaGetLastError());
}

extern "C" {

int snn_b200_conn_compute(const snn_conn_t *conn, int32_t n_src, int32_t n_tgt, int32_t B, const uint8_t *s, float *out,
                          void *stream) {
    if (!conn || !conn->w || !s || !out || n_src <= 0 || n_tgt <= 0 || B <= 0) return SNN_ERR_BAD_ARG;
    if (conn->kind == SNN_CONN_CONV2D) {
        if (!conn->b || conn->cin * conn->hin * conn->win != n_src || conn->cout * conn->hout * conn->wout != n_tgt) return SNN_ERR_BAD_ARG;
        const size_t total = (size_t)B * n_tgt;
        const int blocks = (int)((total + 255) / 256 < 4736 ? (total + 255) / 256 : 4736);
        SNN_LAUNCH(conv_compute_kernel, blocks, 256, 0, (cudaStream_t)stream, *conn, n_src, n_tgt, B, s, out);
        return cuda_rc(cudaGetLastError());
    }
    dim3 grid((n_tgt + SNN_TILE - 1) / SNN_TILE, (B + SNN_GEN_WARPS - 1) / SNN_GEN_WARPS);
    if (grid.y > 64) grid.y = 64;
    SNN_LAUNCH(conn_compute_kernel, grid, SNN_GEN_THREADS, 0, (cudaStream_t)stream, *conn, n_src, n_tgt, B, s, out);
    return cuda_rc(cudaGetLastError());
}

int snn_b200_conn_update(const snn_net_t *net, int32_t ci, int32_t B, void *workspace, size_t workspace_bytes, void *stream_) {
    if (!net || ci < 0 || ci >= net->n_conns || B <= 0 || !workspace) return SNN_ERR_BAD_ARG;
    const snn_conn_t &C = net->conns[ci];
    if (C.src < 0 || C.src >= net->n_layers || C.tgt < 0 || C.tgt >= net->n_layers || !C.w) return SNN_ERR_BAD_ARG;
    // the single-operator update is the dense [n_src, n_tgt] rule application; convolutional weights and the
    // reward-modulated rules (whose state lives in the window plan) are only updated inside run_window
    if (C.kind != SNN_CONN_DENSE && C.kind != SNN_CONN_MCC) return SNN_ERR_UNSUPPORTED;
    if (SNN_RULE_IS_MSTDP(C.rule)) return SNN_ERR_UNSUPPORTED;
    if (C.rule == SNN_RULE_NONE) return SNN_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    DevNet N;
    memset(&N, 0, sizeof(N));
    N.n_layers = net->n_layers; N.n_conns = net->n_conns; N.learning = 1; N.T = 1; N.B = B;
    for (int c = 0; c < net->n_conns; ++c) N.conns[c] = net->conns[c];
    size_t off = 0;
    const int ends[2] = {C.src, C.tgt};
    for (int e = 0; e < 2; ++e) {
        const int l = ends[e];
        DevLayer &D = N.layers[l];
        if (D.bits) continue;  // recurrent connection: same layer twice
        D.L = net->layers[l];
        D.nw = (D.L.n + 31) / 32;
        D.bits = (uint32_t *)((char *)workspace + off);
        off += (sizeof(uint32_t) * (size_t)B * D.nw + 255) / 256 * 256;
        D.xpub = D.L.x;  // slot 0 = the layer's current trace
        if (off > workspace_bytes) return SNN_ERR_WORKSPACE;
        const int warps = B * D.nw;
        SNN_LAUNCH(pack_bits_kernel, (warps * 32 + 255) / 256, 256, 0, stream, D.L.s, D.bits, B, D.L.n, D.nw);
    }
    const size_t smem = gen_smem_bytes(B);
    cudaFuncSetAttribute(conn_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SNN_LAUNCH(conn_update_kernel, N.layers[C.tgt].nw, SNN_GEN_THREADS, smem, stream, N, ci);
    return cuda_rc(cudaGetLastError());
}

int snn_b200_conn_normalize(const snn_conn_t *conn, int32_t n_src, int32_t n_tgt, void *stream) {
    if (!conn || !conn->w || n_src <= 0 || n_tgt <= 0) return SNN_ERR_BAD_ARG;
    if (!conn->has_norm) return SNN_OK;
    if (conn->kind == SNN_CONN_CONV2D) {
        const int F = conn->cout * conn->cin;
        SNN_LAUNCH(conv_normalize_kernel, (F + SNN_GEN_THREADS - 1) / SNN_GEN_THREADS, SNN_GEN_THREADS, 0, (cudaStream_t)stream, *conn);
        return cuda_rc(cudaGetLastError());
    }
    SNN_LAUNCH(conn_normalize_kernel, (n_tgt + SNN_TILE - 1) / SNN_TILE, SNN_GEN_THREADS, 0, (cudaStream_t)stream, *conn, n_src, n_tgt);
    return cuda_rc(cudaGetLastError());
}

int snn_b200_delta_prepare(const float *w, const float *w0, float *dw, size_t count, void *stream) {
    if (!w || !w0 || !dw) return SNN_ERR_BAD_ARG;
    const int blocks = (int)((count + 1023) / 1024 < 1184 ? (count + 1023) / 1024 : 1184);
    SNN_LAUNCH(delta_prepare_kernel, blocks > 0 ? blocks : 1, 256, 0, (cudaStream_t)stream, w, w0, dw, count);
    return cuda_rc(cudaGetLastError());
}

int snn_b200_delta_apply(float *w, const float *w0, const float *dw_sum, int32_t n_src, int32_t n_tgt, int32_t has_clamp,
                         float wmin, float wmax, int32_t has_norm, int32_t norm_abs, float norm, void *stream) {
    if (!w || !w0 || !dw_sum || n_src <= 0 || n_tgt <= 0) return SNN_ERR_BAD_ARG;
    snn_conn_t C;
    memset(&C, 0, sizeof(C));
    C.w = w; C.has_clamp = has_clamp; C.wmin = wmin; C.wmax = wmax; C.has_norm = has_norm; C.norm_abs = norm_abs; C.norm = norm;
    SNN_LAUNCH(delta_apply_kernel, (n_tgt + SNN_TILE - 1) / SNN_TILE, SNN_GEN_THREADS, 0, (cudaStream_t)stream, C, w0, dw_sum, n_src, n_tgt, nullptr, nullptr, 0);
    return cuda_rc(cudaGetLastError());
}

int snn_b200_delta_apply_fused(float *w, const float *dw_sum, int32_t n_src, int32_t n_tgt, int32_t has_clamp, float wmin, float wmax,
                               int32_t has_norm, int32_t norm_abs, float norm, float *theta, const float *dtheta_sum, int32_t n_theta,
                               void *stream) {
    if (!w || !dw_sum || n_src <= 0 || n_tgt <= 0 || (theta && (!dtheta_sum || n_theta <= 0))) return SNN_ERR_BAD_ARG;
    snn_conn_t C;
    memset(&C, 0, sizeof(C));
    C.w = w; C.has_clamp = has_clamp; C.wmin = wmin; C.wmax = wmax; C.has_norm = has_norm; C.norm_abs = norm_abs; C.norm = norm;
    SNN_LAUNCH(delta_apply_kernel, (n_tgt + SNN_TILE - 1) / SNN_TILE, SNN_GEN_THREADS, 0, (cudaStream_t)stream, C, w, dw_sum, n_src, n_tgt, theta, dtheta_sum,
                                                                                                        n_theta);
    return cuda_rc(cudaGetLastError());
}

}  // extern "C"
