#!/bin/bash
# GPU box: every -m gpu test (one process per file), smoke(), the bench line. Usage: bash tools/r2_validate.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
bash tools/gpu_check.sh 2>&1 | grep -E "===|passed|failed|rror|e_max|score1|tile 63" | tee gpurun_out/check_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_$tag.txt
timeout 900 python bench.py 2> gpurun_out/bench_${tag}_err.log > gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_${tag}_err.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("value %.1f e2e %.1f ms/step %.2f clk %s launches %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"], d["gpu_launches"]))
for k,v in d["kernels"].items(): print("  %-18s %.3f ms/step %s" % (k, v["ms_per_step"], ("%.0f TF" % v["tflops"]) if "tflops" in v else ""))
r=d.get("reference_gpu_stack") or {}
print("ref gpu stack:", {k:(round(v,2) if isinstance(v,float) else v) for k,v in r.items() if k!="note"})
print("6b:", {k:v for k,v in (d["extra"].get("config5_internvit_6b") or {}).items() if k not in ("kernels_ms_per_step","note")})
PY
