#!/bin/bash
# scripts/build_variant.sh <name> [-DFLAG ...]: scripts/libsnn_b200_<name>.bin = the library with snn_fused_dc2.cu compiled
# under the given macros (the other objects are compiled once into /tmp/snn_objs)
set -e
name=$1; shift
C=/root/repo/bindsnet_b200/csrc; O=/tmp/snn_objs; mkdir -p $O
F="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -Xcompiler -fPIC -ccbin /usr/bin/g++"
for s in snn_api snn_generic snn_fused_dc snn_ops snn_encode snn_readout; do
  if [ ! -f $O/$s.o ] || [ $C/$s.cu -nt $O/$s.o ] || [ $C/snn_common.cuh -nt $O/$s.o ] || [ $C/snn_phases.cuh -nt $O/$s.o ] || [ /root/repo/include/snn_b200.h -nt $O/$s.o ]; then
    nvcc $F -c $C/$s.cu -o $O/$s.o &
  fi
done
nvcc $F "$@" -Xptxas -v -c $C/snn_fused_dc2.cu -o $O/dc2_$name.o 2> $O/dc2_$name.log
wait
nvcc -shared -cudart static -ccbin /usr/bin/g++ -o /root/repo/scripts/libsnn_b200_$name.bin $O/snn_api.o $O/snn_generic.o $O/snn_fused_dc.o $O/dc2_$name.o $O/snn_ops.o $O/snn_encode.o $O/snn_readout.o
grep -A2 "snn_dc2_windowILi3ELi4ELi0" $O/dc2_$name.log | grep -E "stack|registers" | head -3
