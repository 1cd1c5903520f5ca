#!/bin/bash
# Experiment builds of libsgdx.so with parts of sgdx_match_ph compiled out (results are then WRONG; timing only):
#   tools/build_variants.sh  ->  singular-demux_b200/variants/libsgdx_{nohist,noq,nohist_noq}.so
# Use with SGDX_LIB=<path> python tools/bench_compact.py --no-e2e ...
set -e
cd "$(dirname "$0")/../singular-demux_b200/csrc"
make -j8 all >/dev/null
mkdir -p ../variants build/var
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
OBJS=$(ls build/*.o | grep -v sgdx_ph.o)
for v in "nohist:-DSGDX_PH_X_NOHIST" "noq:-DSGDX_PH_X_NOQ" "nohist_noq:-DSGDX_PH_X_NOHIST -DSGDX_PH_X_NOQ"; do
  name=${v%%:*}; flags=${v#*:}
  $NV $flags -c -o build/var/sgdx_ph_$name.o sgdx_ph.cu
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libsgdx_$name.so $OBJS build/var/sgdx_ph_$name.o -lcudart_static -lpthread -ldl -lrt
done
ls -la ../variants
