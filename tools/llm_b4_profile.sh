#!/bin/bash
# GPU box: ncu launch lists of one InternLM2-7B prefill at B = 1 and B = 4 (per-kernel time by position in the layer).
mkdir -p gpurun_out
PATHK='regex:gemm_bf16_tn|attention_fwd2|llm_seed_residual|llm_rope_kernel|llm_silu_mul|llm_norm_rows'
for B in 1 4; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k "$PATHK" -s 195 -c 195 --csv --log-file gpurun_out/llm_launches_b${B}.csv python tools/bench_llm.py 7b 2108 $B noref > gpurun_out/llm_launches_b${B}.log 2>&1
  tail -1 gpurun_out/llm_launches_b${B}.log
done
python tools/bench_llm.py 7b 2108 1 2 4 noref 2>/dev/null | tee gpurun_out/llm_b124.jsonl
