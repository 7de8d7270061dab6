#!/bin/bash
# ncu passes for profiles/: launch list of one bench step + full captures of the GEMM and attention kernels.
# Usage (GPU box): bash tools/gpu_profile.sh <round-tag>
tag=${1:-r1}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra --no-ref-gpu"
# path kernels only (weight ingest / folding kernels of engine construction are not part of a step); skip the warm-up step
PATHK='regex:gemm_bf16_tn|attention_fwd|im2col_kernel|assemble_tokens|pixel_shuffle_ln|norm_f32_to_bf16|rmsnorm_bf16_inplace'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$PATHK" -s 126 -c 126 --csv --log-file gpurun_out/launches_${tag}.csv $B > gpurun_out/launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 9 -c 4 -o gpurun_out/prof_gemm_${tag} -f $B > gpurun_out/prof_gemm_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 2 -c 1 -o gpurun_out/prof_attn_${tag} -f $B > gpurun_out/prof_attn_${tag}.log 2>&1
ls -la gpurun_out
