#!/bin/bash
# GPU box: front-end parity tests, kernel-only timing, one ncu --set full capture of the 720p launch.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_preprocess.py -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pre_tests.log
timeout 300 python tools/pre_bench.py 2>&1 | tee gpurun_out/pre_bench.log
if [ "$1" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:preprocess -s 2 -c 1 -f -o gpurun_out/prof_pre python tools/pre_bench.py > gpurun_out/prof_pre.log 2>&1
tail -2 gpurun_out/prof_pre.log
fi
