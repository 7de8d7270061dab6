#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
  for v in variants/lib_*.so; do
    echo "== $v"; AIGV_LIB=$PWD/$v timeout 120 python tools/op_bench.py attn 20 2>&1 | tail -1
  done
done | tee gpurun_out/attn_abl.log
AIGV_LIB=$PWD/variants/lib_pp1.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd2 -s 2 -c 1 -f -o gpurun_out/prof_attn2_pp1 python tools/op_bench.py attn 2 > gpurun_out/prof_attn2_pp1.log 2>&1
tail -2 gpurun_out/prof_attn2_pp1.log
