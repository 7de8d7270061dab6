#!/bin/bash
# GPU box: LLM prefill tests + the attention regression tests (the kernel gained a causal / GQA variant).
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_llm_prefill.py -q -m gpu -s -x 2>&1 | tail -40 | tee gpurun_out/llm_tests.log
timeout -s KILL 600 python -m pytest tests/test_gpu_attention.py -q -m gpu 2>&1 | tail -5 | tee -a gpurun_out/llm_tests.log
for rep in 1 2; do timeout 120 python tools_op_bench.py attn 20 2>&1 | tail -1; done | tee -a gpurun_out/llm_tests.log
