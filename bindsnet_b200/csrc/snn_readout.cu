// snn_readout.cu — label assignment and classification from per-sample spike COUNTS, the step AFTER the hot path
// (SURVEY.md §8f rank 2).  The reference's evaluation functions (bindsnet/evaluation/evaluation.py: `assign_labels`
// :8-61, `all_activity` :99-136, `proportion_weighting` :139-180) take [n_samples, time, n_neurons] rasters and sum
// them over time first; the window kernels already deliver that sum (snn_layer_t.rec_count / SpikeCounter), so these
// kernels start from [n_samples, n_neurons] int32 counts.  Counts are integers: every sum of them is exact in fp32
// whatever its order; the weighted sums of `proportion_weighting` use a fixed reduction order (deterministic).
#include "snn_common.cuh"

namespace {

constexpr int MAXL = 64;   // labels (classes)

// rates[j, i] = alpha * rates[j, i] + sum_{s: labels[s] == i} counts[s, j] / #{s: labels[s] == i}   (labels seen only)
// proportions[j, :] = rates[j, :] / sum_i rates[j, i]  (0 where that is 0/0);  assignments[j] = arg max_i proportions
// thread = neuron j (coalesced over j for every sample)
__global__ void __launch_bounds__(128) assign_labels_kernel(const int32_t *__restrict__ counts, const int64_t *__restrict__ labels, int S, int n,
                                                             int L, float alpha, float *__restrict__ rates, float *__restrict__ proportions,
                                                             int64_t *__restrict__ assignments) {
    SNN_SHARED(int, n_lab, MAXL);
    SNN_DYN_SHARED(float, acc);   // [L][blockDim.x]
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = threadIdx.x; i < L; i += blockDim.x) n_lab[i] = 0;
    for (int i = 0; i < L; ++i) acc[i * blockDim.x + threadIdx.x] = 0.0f;
    __syncthreads();
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const int64_t l = labels[s];
        if (l >= 0 && l < L) atomicAdd(&n_lab[(int)l], 1);
    }
    __syncthreads();
    if (j >= n) return;
    for (int s = 0; s < S; ++s) {
        const int64_t l = labels[s];
        if (l >= 0 && l < L) acc[(int)l * blockDim.x + threadIdx.x] += (float)counts[(size_t)s * n + j];
    }
    float tot = 0.0f;
    for (int i = 0; i < L; ++i) {
        float r = rates[(size_t)j * L + i];
        if (n_lab[i] > 0) r = alpha * r + acc[i * blockDim.x + threadIdx.x] / (float)n_lab[i];   // evaluation.py:44-51
        rates[(size_t)j * L + i] = r;
        tot = tot + r;
    }
    int best = 0; float bestp = -1.0f;
    for (int i = 0; i < L; ++i) {
        float p = rates[(size_t)j * L + i] / tot;   // evaluation.py:53-54
        if (p != p) p = 0.0f;
        proportions[(size_t)j * L + i] = p;
        if (p > bestp) { bestp = p; best = i; }     // first maximum, like torch.max
    }
    assignments[j] = best;
}

// rates[s, i] = sum_{j: assignments[j] == i} w[j, i] * counts[s, j] / #{j: assignments[j] == i}, w = 1 (all_activity,
// evaluation.py:99-136) or proportions (proportion_weighting, :139-180); predictions[s] = arg max_i rates[s, i].
// block = sample s; thread t sums neurons t, t + blockDim, ... in ascending order, then a fixed-order tree.
__global__ void __launch_bounds__(256) predict_kernel(const int32_t *__restrict__ counts, const int64_t *__restrict__ assignments,
                                                       const float *__restrict__ proportions, int S, int n, int L,
                                                       int64_t *__restrict__ predictions) {
    SNN_SHARED2(float, part, MAXL, 256 / 32);
    SNN_SHARED(int, n_as, MAXL);
    SNN_SHARED(float, rate, MAXL);
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < L; i += blockDim.x) n_as[i] = 0;
    __syncthreads();
    for (int j = tid; j < n; j += blockDim.x) {
        const int64_t a = assignments[j];
        if (a >= 0 && a < L) atomicAdd(&n_as[(int)a], 1);
    }
    __syncthreads();
    for (int i = 0; i < L; ++i) {
        float acc = 0.0f;
        if (n_as[i] > 0)
            for (int j = tid; j < n; j += blockDim.x)
                if (assignments[j] == i) {
                    const float c = (float)counts[(size_t)s * n + j];
                    acc = acc + (proportions ? proportions[(size_t)j * L + i] * c : c);
                }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc = acc + __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) part[i][warp] = acc;
    }
    __syncthreads();
    if (tid < L) {
        float tot = 0.0f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot = tot + part[tid][w];
        rate[tid] = n_as[tid] > 0 ? tot / (float)n_as[tid] : 0.0f;
    }
    __syncthreads();
    if (tid == 0) {
        int best = 0;
        for (int i = 1; i < L; ++i) if (rate[i] > rate[best]) best = i;   // ties: the lowest label
        predictions[s] = best;
    }
}

}  // namespace

extern "C" {

int snn_b200_assign_labels(const int32_t *counts, const int64_t *labels, int32_t S, int32_t n, int32_t n_labels, float alpha, float *rates,
                           float *proportions, int64_t *assignments, void *stream) {
    if (!counts || !labels || !rates || !proportions || !assignments || S <= 0 || n <= 0 || n_labels <= 0) return SNN_ERR_BAD_ARG;
    if (n_labels > MAXL) return SNN_ERR_UNSUPPORTED;
    const int threads = 128;
    SNN_LAUNCH(assign_labels_kernel, (n + threads - 1) / threads, threads, sizeof(float) * n_labels * threads, (cudaStream_t)stream, counts, labels, S, n, n_labels, alpha, rates, proportions, assignments);
    return cudaGetLastError() == cudaSuccess ? SNN_OK : SNN_ERR_CUDA;
}

int snn_b200_predict(const int32_t *counts, const int64_t *assignments, const float *proportions, int32_t S, int32_t n, int32_t n_labels,
                     int64_t *predictions, void *stream) {
    if (!counts || !assignments || !predictions || S <= 0 || n <= 0 || n_labels <= 0) return SNN_ERR_BAD_ARG;
    if (n_labels > MAXL) return SNN_ERR_UNSUPPORTED;
    SNN_LAUNCH(predict_kernel, S, 256, 0, (cudaStream_t)stream, counts, assignments, proportions, S, n, n_labels, predictions);
    return cudaGetLastError() == cudaSuccess ? SNN_OK : SNN_ERR_CUDA;
}

}  // extern "C"
