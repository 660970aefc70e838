#!/bin/bash
# usage: build_variant.sh <name> "<extra nvcc -D flags>"  -> spada-sim_b200/lib/libspada_b200_<name>.so
set -e
cd "$(dirname "$0")/../spada-sim_b200/csrc"
mkdir -p /tmp/var_$1 && for f in plan esc esc_cta_bitonic fused transpose longrow engine; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr $2 -c $f.cu -o /tmp/var_$1/$f.o &
done; wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libspada_b200_$1.so /tmp/var_$1/*.o -lcudart_static -ldl -lrt -lpthread
ls -la ../lib/libspada_b200_$1.so
