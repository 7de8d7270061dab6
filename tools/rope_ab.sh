#!/bin/bash
# GPU box: LLM parity tests with the rotary embedding fused into the wqkv epilogue, then a same-box A/B against the standalone kernel.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_llm_prefill.py tests/test_gpu_score_parity.py -x -q -m gpu 2>&1 | tail -4
AIGV_LLM_UNFUSED_ROPE=1 timeout 400 python -m pytest tests/test_gpu_llm_prefill.py -x -q -m gpu -k "oracle or everything or reference_forward" 2>&1 | tail -2
for rep in 1 2; do
for U in 0 1; do
  echo "== AIGV_LLM_UNFUSED_ROPE=$U"
  for m in 1_8b 7b; do
  AIGV_LLM_UNFUSED_ROPE=$U timeout 200 python tools/bench_llm.py $m 2108 1 noref 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(' $m B',d['B'],'ms/prompt',round(d['ms_per_prompt'],3))"
  done
done
done
