// emu_dc2.cpp — TEST INFRASTRUCTURE: the column-group fused kernel (tier 3, csrc/snn_fused_dc2.cu) compiled for the host on
// cuda_emu.h.  A translation unit of its own: it names its PTX helpers like snn_fused_dc.cu does.
#include "cuda_emu.h"

#include "../../bindsnet_b200/csrc/snn_fused_dc2.cu"
