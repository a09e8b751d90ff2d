#!/bin/bash
# DEV TOOL: A/B variants of one SPD pair kernel object.  N=<matrix size> tools/build_variant.sh <name> <extra nvcc flags...>
# -> matrix-manifolds_b200/lib/libgm_b200_<name>.so (select with GM_B200_LIB=<path>); N defaults to 4.
set -e
cd "$(dirname "$0")/../matrix-manifolds_b200"
N=${N:-4}
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  --expt-relaxed-constexpr -DGM_N=$N "$@" -Xptxas -v -c csrc/gm_pairs_spd.cu -o build/var_$name.o 2> build/var_$name.ptxas.log
objs=$(ls build/*.o | grep -v "build/var_" | grep -v "gm_pairs_spd_$N.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/libgm_b200_$name.so $objs build/var_$name.o -lcudart
grep -A3 "spd_pair_stream_kernelINS_5SpdAIIfLi${N}ELb0ELb0EEEfLi2" build/var_$name.ptxas.log | grep -E "Used|spill"
