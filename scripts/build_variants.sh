#!/bin/bash
# Builds tuning variants of libmbfea.so (same ABI, different -D knobs of the symmetric SpMV) for A/B runs:
#   scripts/build_variants.sh "MINB UNROLL [PREFETCH [L2PF]]" ...   ->  meshbeamfea_b200/variants/libmbfea_m<MINB>u<UNROLL>.so
set -e
cd "$(dirname "$0")/../meshbeamfea_b200/csrc"
make -s -j8
mkdir -p ../variants build
for spec in "$@"; do
  set -- $spec
  name="m$1u$2p${3:-0}l${4:-1}"
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ccbin /usr/bin/g++ \
     -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -DMBFEA_SYM_MINB=$1 -DMBFEA_SYM_UNROLL=$2 -DMBFEA_SYM_PREFETCH=${3:-0} -DMBFEA_SYM_L2PF=${4:-1} \
     -c kernels.cu -o build/kernels_$name.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libmbfea_$name.so \
     build/kernels_$name.o build/multilevel.o build/api.o build/host_prep.o build/hierarchy.o build/dense.o build/scenario.o build/nccl_dl.o -cudart static -ldl -lpthread
  echo "built variants/libmbfea_$name.so"
done
