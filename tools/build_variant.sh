#!/bin/bash
# Build a library variant for same-box A/B runs: bash tools/build_variant.sh <name> [extra nvcc flags...]
# -> variants/lib_<name>.so (objects compiled in parallel). Select with AIGV_LIB=$PWD/variants/lib_<name>.so.
set -e
name=$1; shift
src=aigv_assessor_b200/csrc
obj=/tmp/aigv_variant_$name
mkdir -p $obj variants
pids=()
for f in api gemm_tcgen05 attention2_tcgen05 rowops preprocess llm_glue llm_prefill motion_conv; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DAIGV_SRC_HASH=\"variant-$name\" "$@" -c $src/$f.cu -o $obj/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/lib_$name.so $obj/*.o
echo built variants/lib_$name.so
