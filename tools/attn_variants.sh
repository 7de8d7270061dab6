#!/bin/bash
# GPU box: same-box A/B of prebuilt attention variants (variants/lib_*.so) against the default build.
mkdir -p gpurun_out
for rep in 1 2; do
  echo "== default"; timeout 120 python tools/op_bench.py attn 20 2>&1 | tail -1
  for v in variants/lib_*.so; do
    [ -f $v ] || continue
    echo "== $v"; AIGV_LIB=$PWD/$v timeout 120 python tools/op_bench.py attn 20 2>&1 | tail -1
  done
done | tee gpurun_out/attn_variants.log
