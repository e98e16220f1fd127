// snn_combine.cuh — the multi-GPU window combine on one 32-column tile:
//     w = clamp(w0 + sum_r dw_r), then normalize()      [+ theta = theta0 + sum_r dtheta_r]
// One CTA per tile, so that the column sums of normalize() stay CTA-local; the rows of the tile are walked by the CTA's
// warps, EIGHT ROWS IN FLIGHT per warp in every pass (apply, column sums, scale): the passes are pure streams whose
// cost is the L2 / HBM latency of a row, so the number of rows in flight is what sets their speed (the round-2
// measurement: 50 CTAs walking 98 rows each, one dependent load -> store at a time, cost more than the all-reduce).
// Summation order of the column sums: the SNN_NORM_CHUNKS contiguous row chunks of include/snn_b200.h, each in ascending
// row order, then ascending chunk order — the order of normalize_tile and of the oracle, so the result is bit-identical.
#pragma once
#include "snn_common.cuh"

namespace {

__device__ __forceinline__ void delta_apply_tile(const snn_conn_t &C, const float *w0, const float *__restrict__ dws, int ns, int nt, int tile,
                                                 float *s_part) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = tile * SNN_TILE + lane;
    const bool valid = j < nt;
    float *wcol = C.w + j;
    const float *w0col = w0 + j, *dcol = dws + j;
    // apply: w = clamp(w0 + dw)
    for (int i0 = warp; i0 < ns; i0 += 8 * SNN_GEN_WARPS) {
        float a[8], d[8];
        #pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = i0 + q * SNN_GEN_WARPS;
            const bool ok = valid && i < ns;
            a[q] = ok ? w0col[(size_t)i * nt] : 0.0f;
            d[q] = ok ? dcol[(size_t)i * nt] : 0.0f;
        }
        #pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = i0 + q * SNN_GEN_WARPS;
            if (valid && i < ns) {
                float x = a[q] + d[q];
                if (C.has_clamp) x = clampf(x, C.wmin, C.wmax);
                wcol[(size_t)i * nt] = x;
            }
        }
    }
    if (!C.has_norm) return;
    // normalize(): Connection.normalize (topology.py:383-392) / AbstractFeature.normalize (topology_features.py:250-266)
    const int chunk = (ns + SNN_NORM_CHUNKS - 1) / SNN_NORM_CHUNKS;
    __syncthreads();
    for (int c = warp; c < SNN_NORM_CHUNKS; c += SNN_GEN_WARPS) {
        float part = 0.0f;
        const int i1 = min((c + 1) * chunk, ns);
        for (int i0 = c * chunk; i0 < i1; i0 += 8) {
            float x[8];
            #pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = (valid && i0 + q < i1) ? wcol[(size_t)(i0 + q) * nt] : 0.0f;
            #pragma unroll
            for (int q = 0; q < 8; ++q)
                if (i0 + q < i1) part = part + (C.norm_abs ? fabsf(x[q]) : x[q]);
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
    for (int i0 = warp; i0 < ns; i0 += 8 * SNN_GEN_WARPS) {
        float x[8];
        #pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = i0 + q * SNN_GEN_WARPS;
            x[q] = (valid && i < ns) ? wcol[(size_t)i * nt] : 0.0f;
        }
        #pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = i0 + q * SNN_GEN_WARPS;
            if (valid && i < ns) wcol[(size_t)i * nt] = x[q] * f;
        }
    }
}

}  // namespace
