#!/bin/bash
# Builds tuning variants of libmbfea.so (same ABI, different -D knobs of the symmetric SpMV) for A/B runs:
#   scripts/build_variants.sh name1 "-DMBFEA_SYM_STREAM=1" name2 "-DMBFEA_SYM_MINB=5 -DMBFEA_SYM_UNROLL=1" ...
#   ->  meshbeamfea_b200/variants/libmbfea_<name>.so   (select with MBFEA_LIB=...)
set -e
cd "$(dirname "$0")/../meshbeamfea_b200/csrc"
make -s -j8
mkdir -p ../variants build
while [ $# -ge 2 ]; do
  name="$1"; flags="$2"; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ccbin /usr/bin/g++ \
     -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr $flags -c kernels.cu -o build/kernels_$name.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libmbfea_$name.so \
     build/kernels_$name.o build/mlprec.o build/mlprec_host.o build/api.o build/host_prep.o build/hierarchy.o build/dense.o \
     build/scenario.o build/nccl_dl.o build/hostmem.o -cudart static -ldl -lpthread
  echo "built variants/libmbfea_$name.so"
done
