#!/bin/bash
# GPU box: compute-sanitizer on the final build (programmatic dependent launch in the GEMM / attention kernels, grouped tile order).
tag=${1:-r2z}
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
out=gpurun_out/sanitizer_$tag.txt
echo "# compute-sanitizer on B200 ($tag build: PDL launches, grouped GEMM tile order)" > $out
echo "## memcheck: GEMM tile orders with a large weight (ragged groups), bias shapes" >> $out
timeout 1500 $SAN --tool memcheck python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "tile_orders or bias" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid" | head -6 >> $out
echo "## memcheck: attention (ViT and causal GQA shapes, several items per CTA)" >> $out
timeout 1500 $SAN --tool memcheck python -m pytest tests/test_gpu_attention.py -q -m gpu -k "matches_reference and (1025-2-64 or 300-16-64 or 2108 or 300-25-128)" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid" | head -6 >> $out
echo "## memcheck: InternViT-300M one frame + tiny towers (the whole PDL chain of a forward)" >> $out
timeout 1500 $SAN --tool memcheck python -m pytest tests/test_gpu_extract_feature.py -q -m gpu -k "vit300m_frame or tiny_tower" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid" | head -6 >> $out
echo "## racecheck: tiny tower forward, one attention case" >> $out
timeout 1500 $SAN --tool racecheck python -m pytest tests/test_gpu_extract_feature.py -q -m gpu -k "tiny_tower" 2>&1 | grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|hazard" | head -6 >> $out
timeout 1500 $SAN --tool racecheck python -m pytest tests/test_gpu_attention.py -q -m gpu -k "matches_reference and 1-256-2-64" 2>&1 | grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|hazard" | head -6 >> $out
cat $out
