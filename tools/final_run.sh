set -x
# GPU box: what the driver runs at round end, in its order (GPU tests in ONE process, smoke, the reference arm, the bench), plus a record.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/gpu_final.txt
timeout -s KILL 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/final_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/final_smoke.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null > gpurun_out/final_bench_ref.json
timeout 600 python bench.py 2>gpurun_out/final_bench_err.log > gpurun_out/final_bench.json; tail -2 gpurun_out/final_bench_err.log
python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print(d['value'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['roofline_attention']['ms_per_step']); e=d['extra']; print({k:(v.get('clips_per_s') or v.get('ms_per_prompt') or v.get('frames_per_s')) for k,v in e.items() if isinstance(v,dict)})"
ls -la gpurun_out | tail -8
