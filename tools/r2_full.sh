#!/bin/bash
# Full GPU validation: every -m gpu test file, op-bench A/B of the variants, then the bench line.
mkdir -p gpurun_out
bash tools/gpu_check.sh 2>&1 | grep -E "===|passed|failed|error|Error|e_max|score1|tile 63" | tee gpurun_out/check.log
for rep in 1 2; do
  echo "== default"; timeout 120 python tools/op_bench.py all 20 2>&1 | tail -5
  for v in variants/lib_*.so; do
    [ -f $v ] || continue
    echo "== $v"; AIGV_LIB=$PWD/$v timeout 120 python tools/op_bench.py all 20 2>&1 | tail -5
  done
done | tee gpurun_out/op_ab.log
timeout 900 python bench.py 2> gpurun_out/bench_err.log > gpurun_out/bench_r2a.json; tail -3 gpurun_out/bench_err.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2a.json"))
print("value %.1f e2e %.1f ms/step %.2f clk %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]))
for k,v in d["kernels"].items(): print("  %-18s %.3f ms/step %s" % (k, v["ms_per_step"], ("%.0f TF" % v["tflops"]) if "tflops" in v else ""))
print(json.dumps(d.get("reference_gpu_stack"), indent=1))
print(json.dumps(d.get("extra"), indent=1)[:3000])
print(d.get("roofline_attention"), d.get("cpu_baseline"))
PY
