#!/bin/bash
# GPU box: compute-sanitizer memcheck on the last additions (delta proj / fc2 epilogues on the real 300M shapes, motion kernels).
tag=${1:-r2n}
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
out=gpurun_out/sanitizer_$tag.txt
echo "# compute-sanitizer memcheck on B200 ($tag build)" > $out
echo "## InternViT-300M one frame (delta proj kResid 3 + fc2 delta_in), tiny tower" >> $out
timeout 1200 $SAN --tool memcheck python -m pytest tests/test_gpu_extract_feature.py -q -m gpu -k "vit300m_frame or tiny_tower" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid" | head -6 >> $out
echo "## motion branch: im2col3d + gemm_ex (both paths, clipped stores), pools, tail, trunk at 224" >> $out
timeout 1200 $SAN --tool memcheck python -m pytest tests/test_gpu_slowfast.py -q -m gpu -k "im2col3d or gemm_ex or pool_join or (trunk_matches and 224)" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid" | head -6 >> $out
cat $out
