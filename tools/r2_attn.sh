#!/bin/bash
# Round-2 attention bring-up on the GPU box: parity tests (default library and every variant), then same-box timing.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/gpu.txt
timeout -s KILL 900 python -m pytest tests/test_gpu_attention.py -q -m gpu 2>&1 | tail -8 | tee gpurun_out/attn_tests.log
for v in variants/lib_*.so; do
  [ -f $v ] || continue
  echo "== tests with $v"; AIGV_LIB=$PWD/$v timeout -s KILL 900 python -m pytest tests/test_gpu_attention.py -q -m gpu 2>&1 | tail -6
done | tee -a gpurun_out/attn_tests.log
for rep in 1 2; do
  echo "== default"; timeout 120 python tools/op_bench.py attn 20 2>&1 | tail -1
  for v in variants/lib_*.so; do
    [ -f $v ] || continue
    echo "== $v"; AIGV_LIB=$PWD/$v timeout 120 python tools/op_bench.py attn 20 2>&1 | tail -1
  done
done | tee gpurun_out/attn_ab.log
