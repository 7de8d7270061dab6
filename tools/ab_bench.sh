#!/bin/bash
# GPU box: same-box A/B of library variants (variants/lib_*.so) against the default build on the bench's kernel table.
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-ref-gpu --no-e2e"
for rep in 1 2; do
  for lib in "" variants/lib_*.so; do
    [ -z "$lib" ] || lib=$PWD/$lib
    AIGV_LIB=$lib timeout 600 $B 2>/dev/null > gpurun_out/ab_tmp.json
    python - <<PY
import json
d=json.load(open("gpurun_out/ab_tmp.json"))
k=d["kernels"]
print("lib=%-40s rep $rep: %.1f frames/s  fc1 %.3f fc2 %.3f qkv %.3f proj %.3f attn %.3f ms/step clk %s" % ("${lib##*/}" or "default", d["value"], k["gemm_fc1"]["ms_per_step"], k["gemm_fc2"]["ms_per_step"], k["gemm_qkv"]["ms_per_step"], k["gemm_proj"]["ms_per_step"], k["attention"]["ms_per_step"], d["clocks"]["sm_mhz"]))
PY
  done
done | tee gpurun_out/ab_bench.log
