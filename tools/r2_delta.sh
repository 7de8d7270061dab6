#!/bin/bash
# GPU box: parity of the "delta" proj scheme + same-box A/B against AIGV_DELTA_PROJ=0.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_extract_feature.py tests/test_gpu_gemm.py tests/test_gpu_score_parity.py -q -m gpu -s 2>&1 | grep -E "passed|failed|rror|e_max|score1|tile 63" | tee gpurun_out/delta_tests.log
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-ref-gpu"
for rep in 1 2; do
  for mode in 1 0; do
    AIGV_DELTA_PROJ=$mode timeout 600 $B 2>/dev/null > gpurun_out/delta_${mode}_$rep.json
    python - <<PY
import json
d=json.load(open("gpurun_out/delta_${mode}_$rep.json"))
k=d["kernels"]
print("delta=$mode rep $rep: %.1f frames/s e2e %.1f  proj %.3f fc2 %.3f fc1 %.3f qkv %.3f attn %.3f ms/step clk %s" % (d["value"], d["e2e"]["value"], k["gemm_proj"]["ms_per_step"], k["gemm_fc2"]["ms_per_step"], k["gemm_fc1"]["ms_per_step"], k["gemm_qkv"]["ms_per_step"], k["attention"]["ms_per_step"], d["clocks"]["sm_mhz"]))
PY
  done
done | tee gpurun_out/delta_ab.log
