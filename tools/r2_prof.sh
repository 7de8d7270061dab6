#!/bin/bash
# GPU box: new tests, attention timing, ncu evidence for profiles/, compute-sanitizer racecheck / memcheck on small cases.
tag=${1:-r2c}
mkdir -p gpurun_out
for f in tests/test_gpu_attention.py tests/test_gpu_llm_glue.py; do
  timeout -s KILL 600 python -m pytest $f -q -m gpu -s 2>&1 | grep -E "passed|failed|rror|all-CUDA" | tee -a gpurun_out/check_$tag.log
done
for rep in 1 2; do timeout 120 python tools/op_bench.py attn 20 2>&1 | tail -1; done | tee gpurun_out/attn_$tag.log
bash tools/gpu_profile.sh $tag > gpurun_out/profile_$tag.log 2>&1
tail -3 gpurun_out/profile_$tag.log
# race / memory checks on small cases (the sanitizer serialises everything: keep the shapes tiny)
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $SAN --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_attention.py -q -m gpu -k "1-128-1-64 or 3-257-2-64 or 1-128-1-128" 2>&1 | tail -15 > gpurun_out/racecheck_attention_$tag.txt
timeout 600 $SAN --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "exact_small or rejects or assemble" 2>&1 | tail -15 > gpurun_out/racecheck_gemm_$tag.txt
timeout 600 $SAN --tool memcheck python -m pytest tests/test_gpu_attention.py tests/test_gpu_llm_glue.py -q -m gpu -k "1-128-1-64 or 2-100-2-64 or 1-300-2-64 or splice or mlpscore_head" 2>&1 | tail -15 > gpurun_out/memcheck_$tag.txt
tail -4 gpurun_out/racecheck_attention_$tag.txt gpurun_out/racecheck_gemm_$tag.txt gpurun_out/memcheck_$tag.txt
