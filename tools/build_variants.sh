#!/bin/bash
# Experiment builds of libsgdx.so with sgdx_ph.cu and sgdx_phx.cu compiled under extra macros (A/B runs on one box; some variants give
# WRONG results and are for timing only):
#   tools/build_variants.sh [name:flags ...]  ->  singular-demux_b200/variants/libsgdx_<name>.so
#   default set: nohist (-DSGDX_PH_X_NOHIST), noq (-DSGDX_PH_X_NOQ), nohist_noq
# Use with SGDX_LIB=<path> python tools/bench_compact.py --no-e2e ...   (remove variants/ afterwards: it travels with gpurun)
set -e
cd "$(dirname "$0")/../singular-demux_b200/csrc"
make -j8 all >/dev/null
mkdir -p ../variants build/var
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
OBJS=$(ls build/*.o | grep -v "sgdx_ph.o\|sgdx_phx.o")
SET=("$@")
[ ${#SET[@]} -eq 0 ] && SET=("nohist:-DSGDX_PH_X_NOHIST" "noq:-DSGDX_PH_X_NOQ" "nohist_noq:-DSGDX_PH_X_NOHIST -DSGDX_PH_X_NOQ")
for v in "${SET[@]}"; do
  name=${v%%:*}; flags=${v#*:}
  $NV $flags -c -o build/var/sgdx_ph_$name.o sgdx_ph.cu
  $NV $flags -c -o build/var/sgdx_phx_$name.o sgdx_phx.cu
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libsgdx_$name.so $OBJS build/var/sgdx_ph_$name.o build/var/sgdx_phx_$name.o -lcudart_static -lpthread -ldl -lrt
done
ls -la ../variants
