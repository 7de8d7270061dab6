#!/bin/bash
tag=${1:-r2d}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_llm_glue.py tests/test_gpu_gemm.py -q -m gpu -s 2>&1 | grep -E "passed|failed|rror|all-CUDA" | tee gpurun_out/check_$tag.log
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $SAN --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "exact_small or rejects or assemble or residual_stats" > gpurun_out/racecheck_gemm_$tag.txt 2>&1
grep -E "RACECHECK SUMMARY|passed|failed|Error: Race" gpurun_out/racecheck_gemm_$tag.txt | sort | uniq -c | head
timeout 600 $SAN --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_attention.py -q -m gpu -k "1-128-1-64 or 3-257-2-64 or 1-128-1-128 or 2-100-2-64 or 1-300-2-64" > gpurun_out/racecheck_attention_$tag.txt 2>&1
grep -E "RACECHECK SUMMARY|passed|failed|Error: Race" gpurun_out/racecheck_attention_$tag.txt | sort | uniq -c | head
timeout 600 $SAN --tool memcheck python -m pytest tests/test_gpu_attention.py tests/test_gpu_llm_glue.py tests/test_gpu_gemm.py -q -m gpu -k "1-128-1-64 or 2-100-2-64 or 1-300-2-64 or splice or mlpscore_head or exact_small or residual_stats" > gpurun_out/memcheck_$tag.txt 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_$tag.txt | head
for rep in 1 2; do timeout 120 python tools/op_bench.py all 20 2>&1 | tail -5; done | tee gpurun_out/op_$tag.log
