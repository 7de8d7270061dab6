#!/bin/bash
# GPU box: parity of the default attention build + of every prebuilt variant, then a same-box A/B of the kernel alone.
mkdir -p gpurun_out
echo "== default parity"; timeout 300 python -m pytest tests/test_gpu_attention.py -q -m gpu 2>&1 | tail -4
for v in variants/lib_*.so; do
  [ -f $v ] || continue
  echo "== $v parity"; AIGV_LIB=$PWD/$v timeout 300 python -m pytest tests/test_gpu_attention.py -q -m gpu 2>&1 | tail -3
done
bash tools/attn_variants.sh 2>&1 | tail -20
