#!/bin/bash
# GPU box: LLM prefill tests + the attention regression tests (the kernel gained a causal / GQA variant) + prefill timing.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_llm_prefill.py -q -m gpu -s -x 2>&1 | tail -12 | tee gpurun_out/llm_tests.log
AIGV_LLM_UNFUSED_SWIGLU=1 timeout -s KILL 900 python -m pytest tests/test_gpu_llm_prefill.py -q -m gpu -x -k "prefill or score" 2>&1 | tail -3 | tee -a gpurun_out/llm_tests.log
timeout -s KILL 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_gemm.py -q -m gpu 2>&1 | tail -5 | tee -a gpurun_out/llm_tests.log
for rep in 1 2; do timeout 120 python tools/op_bench.py attn 20 2>&1 | tail -1; done | tee -a gpurun_out/llm_tests.log
timeout 600 python tools/bench_llm.py 7b 2108 1 2 4 2>&1 | tee gpurun_out/llm_bench_7b.log
AIGV_LLM_UNFUSED_SWIGLU=1 timeout 600 python tools/bench_llm.py 7b 2108 1 4 2>&1 | tee gpurun_out/llm_bench_7b_unfused.log
timeout 300 python tools/bench_llm.py 1_8b 2108 1 4 2>&1 | tee gpurun_out/llm_bench_1_8b.log
