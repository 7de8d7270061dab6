#!/bin/bash
# GPU box: GEMM tile-order test + same-box A/B of the grouped tile order on the InternLM2-7B prefill.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "tile_orders" 2>&1 | tail -3
for G in ${GROUPS_AB:-0 12 16 24 16 0}; do
  echo "== AIGV_GEMM_GROUP_M=$G"
  AIGV_GEMM_GROUP_M=$G timeout 300 python tools/bench_llm.py 7b 2108 ${BS_AB:-4} noref 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(' B',d['B'],'ms/prompt',round(d['ms_per_prompt'],2),'TFLOP/s',round(d['tflops'],1))"
done
