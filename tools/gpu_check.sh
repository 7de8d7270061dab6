#!/bin/bash
# Run every GPU test file in its own process (a trapped kernel poisons only its own CUDA context), with timeouts.
# Usage (on the GPU box): bash tools/gpu_check.sh [extra pytest args]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/gpu.txt
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
rc=0
for f in tests/test_gpu_rowops.py tests/test_gpu_gemm.py tests/test_gpu_attention.py tests/test_gpu_preprocess.py tests/test_gpu_extract_feature.py tests/test_gpu_score_parity.py tests/test_gpu_llm_glue.py tests/test_gpu_llm_prefill.py tests/test_gpu_slowfast.py; do
  echo "=== $f"
  timeout -s KILL 420 python -m pytest "$f" -q -m gpu -s "$@" 2>&1 | tail -40 | tee -a gpurun_out/tests.log
  [ ${PIPESTATUS[0]} -ne 0 ] && rc=1
done
exit $rc
