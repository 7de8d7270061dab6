#!/bin/bash
# GPU box: ncu evidence for the InternLM2-7B prefill (next-1): launch list of one prefill + full captures of the causal
# attention kernel and of the w1|w3 SwiGLU GEMM. Usage: bash tools/llm_profile.sh <tag>
tag=${1:-r2h}
mkdir -p gpurun_out
B="python tools/bench_llm.py 7b 2108 1 noref"
PATHK='regex:gemm_bf16_tn|attention_fwd2|llm_seed_residual|llm_rope_kernel|llm_silu_mul|llm_norm_rows'
# the first prefill (196 launches incl. the rope table's neighbours) is the cold warm-up: list the second one
ncu --metrics gpu__time_duration.sum --clock-control none -k "$PATHK" -s 195 -c 195 --csv --log-file gpurun_out/llm_launches_${tag}.csv $B > gpurun_out/llm_launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_fwd2 -s 34 -c 1 -o gpurun_out/prof_llm_attn_${tag} -f $B > gpurun_out/prof_llm_attn_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 132 -c 4 -o gpurun_out/prof_llm_gemm_${tag} -f $B > gpurun_out/prof_llm_gemm_${tag}.log 2>&1
ls -la gpurun_out | grep llm
