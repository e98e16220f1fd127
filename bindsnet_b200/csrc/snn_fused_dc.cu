// snn_fused_dc.cu — fused persistent window kernel for the DiehlAndCook2015 graph (placeholder
// until the kernel lands: reports "not supported" so every plan takes the generic tier).
#include "snn_common.cuh"

int snn_fused_dc_supported(const snn_net_t *, const snn_run_opts_t *) { return 0; }
size_t snn_fused_dc_workspace_bytes(const snn_net_t *, const snn_run_opts_t *) { return 0; }
int snn_fused_dc_launch(const snn_net_t *, const snn_run_opts_t *, void *, size_t, cudaStream_t, int *) { return SNN_ERR_UNSUPPORTED; }
