#!/bin/bash
# GPU box: compute-sanitizer on the round-2 additions (small cases: the sanitizer serialises the GPU).
tag=${1:-r2j}
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
out=gpurun_out/sanitizer_$tag.txt
echo "# compute-sanitizer on B200 ($tag build): round-2 additions" > $out
echo "## memcheck: front end (720p, 6-tile, 449x447), rotary, causal / GQA attention (ragged, multi-item), InternLM2 prefill (2 x 300 tokens), ClipScorer pieces" >> $out
timeout 900 $SAN --tool memcheck python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_llm_prefill.py -q -m gpu -k "oracle_and_pillow and (720 or 300-200 or 449) or rope or (gqa and (100-4-2 or 129-4-1 or 300-8-2)) or prefill_matches or rejects" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" | head -12 >> $out
echo "## racecheck: causal / GQA attention (single and multi tile, both head dims), rotary" >> $out
timeout 900 $SAN --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_llm_prefill.py -q -m gpu -k "rope or (gqa and (100-4-2 or 129-4-1 or 128-2-2 or 1-1-2-1))" 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|Error: Race" | sort | uniq -c | head -8 >> $out
echo "## racecheck: front end" >> $out
timeout 900 $SAN --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_preprocess.py -q -m gpu -k "oracle_and_pillow and (300-200 or 449)" 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|Error: Race" | sort | uniq -c | head -8 >> $out
cat $out
