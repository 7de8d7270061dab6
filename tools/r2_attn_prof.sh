#!/bin/bash
mkdir -p gpurun_out
for v in poly0 poly2; do
  lib=$PWD/variants/lib_$v.so
  [ -f $lib ] || lib=$PWD/aigv_assessor_b200/libaigv_b200.so
  AIGV_LIB=$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd2 -s 2 -c 1 -f -o gpurun_out/prof_attn2_$v python tools/op_bench.py attn 2 > gpurun_out/prof_attn2_$v.log 2>&1
  tail -2 gpurun_out/prof_attn2_$v.log
done
ls -la gpurun_out/*.ncu-rep | tail -3
