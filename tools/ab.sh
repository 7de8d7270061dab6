#!/bin/bash
# A/B prebuilt library variants on ONE box (per-kernel times differ by several % between boxes).
# Usage: bash tools/ab.sh name=path[,ENV=VAL] ... ; writes gpurun_out/ab_<name>_<rep>.json
mkdir -p gpurun_out
for rep in 1 2; do
  for spec in "$@"; do
    name=${spec%%=*}; rest=${spec#*=}
    lib=${rest%%,*}; envs=""
    [ "$rest" != "$lib" ] && envs=${rest#*,}
    env AIGV_LIB=$PWD/$lib ${envs//,/ } timeout 300 python bench.py --steps 6 --warmup 3 2>/dev/null > gpurun_out/ab_${name}_${rep}.json
    python - <<PY
import json
d=json.load(open("gpurun_out/ab_${name}_${rep}.json"))
k=d["kernels"]
print("${name} rep${rep}: %.1f f/s  e2e %.1f  clk %s | " % (d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"]) + " ".join("%s %.2f" % (n.replace("gemm_",""), k[n]["ms_per_step"]) for n in ("attention","gemm_fc1","gemm_fc2","gemm_qkv","gemm_proj","norm") if n in k))
PY
  done
done
