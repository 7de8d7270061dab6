set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/gpu_final.txt
timeout -s KILL 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/final_pytest.txt
timeout 600 python bench.py 2>gpurun_out/final_bench_err.log > gpurun_out/final_bench.json; tail -2 gpurun_out/final_bench_err.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/final_smoke.txt
bash tools_gpu_profile.sh r1f > gpurun_out/final_profile.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null > gpurun_out/final_bench_ref.json
ls -la gpurun_out | tail -12
