// snn_encode.cu — on-device spike encoders, the step BEFORE the hot path.
//
// The reference's callers encode every batch on the CPU (bindsnet/encoding/encodings.py: `bernoulli` :50-96,
// `poisson` :99-156) and ship the [time, batch, ...] uint8 spike tensor to the device — 25 MB per 250-step window
// of 128 MNIST-sized samples, which at window times well below a millisecond is the end-to-end bottleneck.  These
// kernels take the rate image that is already on the device (400 KB per window) and write the spike tensor there.
//
// Random numbers: counter-based Philox-4x32-10 (implemented here, no library state), keyed by the caller's seed
// and the element index, so a spike train depends only on (seed, element) — not on grid shape or launch order.
// Semantics are the reference's, matched in distribution (the reference draws from torch's global generator, so
// bitwise equality with it is meaningless):
//   poisson  : inter-spike intervals ~ Poisson(1000 / (rate_hz * dt)) steps, a zero interval counts as one step,
//              spike times are the running sums of the intervals (encodings.py:137-154); rate 0 never spikes;
//   bernoulli: one independent trial per step with probability max_prob * p (encodings.py:84-94).
#include "snn_common.cuh"

namespace {

struct Philox {
    uint32_t key0, key1;    // seed
    uint32_t c0, c1, c2;    // element index (64 bit) + stream id; c3 = block counter
    uint32_t ctr;           // 128-bit blocks drawn so far
    uint32_t out[4];
    int have;               // unread words in out[]

    __device__ __forceinline__ void init(uint64_t seed, uint64_t element, uint32_t stream) {
        key0 = (uint32_t)seed; key1 = (uint32_t)(seed >> 32);
        c0 = (uint32_t)element; c1 = (uint32_t)(element >> 32); c2 = stream;
        ctr = 0; have = 0;
    }
    __device__ __forceinline__ void block() {
        uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = ctr++, k0 = key0, k1 = key1;
        #pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            const uint32_t y0 = hi1 ^ x1 ^ k0, y1 = lo1, y2 = hi0 ^ x3 ^ k1, y3 = lo0;
            x0 = y0; x1 = y1; x2 = y2; x3 = y3;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
        have = 4;
    }
    __device__ __forceinline__ uint32_t next() {
        if (have == 0) block();
        return out[--have];
    }
    // uniform in (0, 1]: never 0, so that log() is finite
    __device__ __forceinline__ float uniform() { return ((float)(next() >> 8) + 1.0f) * (1.0f / 16777216.0f); }
};

// Poisson(lambda) variate.  Small lambda: multiplication of uniforms (Knuth).  Large lambda: Hörmann's transformed
// rejection with squeeze (PTRS, 1993), exact for lambda >= 10.
__device__ float poisson_draw(Philox &g, float lambda) {
    if (lambda < 10.0f) {
        const float limit = __expf(-lambda);
        float prod = g.uniform();
        int k = 0;
        while (prod > limit) { prod *= g.uniform(); ++k; }
        return (float)k;
    }
    const float slam = sqrtf(lambda), loglam = __logf(lambda);
    const float b = 0.931f + 2.53f * slam, a = -0.059f + 0.02483f * b;
    const float invalpha = 1.1239f + 1.1328f / (b - 3.4f), vr = 0.9277f - 3.6224f / (b - 2.0f);
    for (;;) {
        const float U = g.uniform() - 0.5f, V = g.uniform();
        const float us = 0.5f - fabsf(U);
        const float k = floorf((2.0f * a / us + b) * U + lambda + 0.43f);
        if (us >= 0.07f && V <= vr) return k;
        if (k < 0.0f || (us < 0.013f && V > us)) continue;
        if (__logf(V) + __logf(invalpha) - __logf(a / (us * us) + b) <= -lambda + k * loglam - lgammaf(k + 1.0f)) return k;
    }
}

// thread = one input element; the threads of a warp sweep the time axis together, so every step's store is one
// coalesced row segment of out[t][.]
__global__ void __launch_bounds__(256) encode_poisson_kernel(const float *__restrict__ rate_hz, int n, int T, float dt, uint64_t seed,
                                                             uint8_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float r = rate_hz[i];
    Philox g;
    g.init(seed, (uint64_t)i, 0x504F4953u);
    const float lambda = r > 0.0f ? (1000.0f / dt) / r : 0.0f;   // mean interval in steps (encodings.py:139-140)
    // next spike time, 1-based like the reference's cumulative sum; a zero interval counts as one step
    float next = 0.0f;
    if (r > 0.0f) { const float k = poisson_draw(g, lambda); next = k < 1.0f ? 1.0f : k; }
    #pragma unroll 1
    for (int t = 0; t < T; ++t) {
        const bool s = r > 0.0f && next == (float)(t + 1);
        out[(size_t)t * n + i] = s ? 1 : 0;
        if (s) { const float k = poisson_draw(g, lambda); next += k < 1.0f ? 1.0f : k; }
    }
}

__global__ void __launch_bounds__(256) encode_bernoulli_kernel(const float *__restrict__ prob, int n, int T, uint64_t seed,
                                                               uint8_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float p = prob[i];
    Philox g;
    g.init(seed, (uint64_t)i, 0x4245524Eu);
    #pragma unroll 1
    for (int t = 0; t < T; ++t) {
        // uniform in [0, 1): a trial succeeds iff u < p, like torch.bernoulli
        const float u = (float)(g.next() >> 8) * (1.0f / 16777216.0f);
        out[(size_t)t * n + i] = u < p ? 1 : 0;
    }
}

}  // namespace

extern "C" {

int snn_b200_encode_poisson(const float *rate_hz, int32_t n, int32_t T, float dt, uint64_t seed, uint8_t *out, void *stream) {
    if (!rate_hz || !out || n <= 0 || T < 0 || !(dt > 0.0f)) return SNN_ERR_BAD_ARG;
    if (T == 0) return SNN_OK;
    SNN_LAUNCH(encode_poisson_kernel, (n + 255) / 256, 256, 0, (cudaStream_t)stream, rate_hz, n, T, dt, seed, out);
    return cudaGetLastError() == cudaSuccess ? SNN_OK : SNN_ERR_CUDA;
}

int snn_b200_encode_bernoulli(const float *prob, int32_t n, int32_t T, uint64_t seed, uint8_t *out, void *stream) {
    if (!prob || !out || n <= 0 || T < 0) return SNN_ERR_BAD_ARG;
    if (T == 0) return SNN_OK;
    SNN_LAUNCH(encode_bernoulli_kernel, (n + 255) / 256, 256, 0, (cudaStream_t)stream, prob, n, T, seed, out);
    return cudaGetLastError() == cudaSuccess ? SNN_OK : SNN_ERR_CUDA;
}

}  // extern "C"
