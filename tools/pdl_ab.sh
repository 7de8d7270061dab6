#!/bin/bash
# GPU box: programmatic dependent launch on / off (AIGV_PDL), same box: the ViT step (instrumentation off) and the InternLM2 prefills.
mkdir -p gpurun_out
for rep in 1 2; do
for P in 1 0; do
  echo "== AIGV_PDL=$P"
  AIGV_PDL=$P timeout 200 python tools/profile_overhead.py 2>&1 | grep "off" | tail -2
  AIGV_PDL=$P timeout 200 python tools/bench_llm.py 1_8b 2108 1 noref 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(' 1.8b B',d['B'],'ms/prompt',round(d['ms_per_prompt'],3))"
  AIGV_PDL=$P timeout 200 python tools/bench_llm.py 7b 2108 1 noref 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(' 7b   B',d['B'],'ms/prompt',round(d['ms_per_prompt'],3))"
done
done
