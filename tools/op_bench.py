#!/usr/bin/env python
"""Micro-benchmarks of single ops through the C-ABI (GPU box): attention and the four layer GEMM shapes."""
import _path  # noqa: F401  (repository root on sys.path)
import sys
import torch
sys.path.insert(0, ".")
from aigv_assessor_b200 import _lib
from tests.gpu_util import ptr, stream

lib = _lib.load()
frames, S, H, d, D, I = 64, 1025, 16, 64, 1024, 4096
if len(sys.argv) > 3:
    S = int(sys.argv[3])   # e.g. 1024: no class-token row (the side warps idle), no lone tail key
M = frames * S
which = sys.argv[1] if len(sys.argv) > 1 else "all"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10


def timeit(fn, flops, name):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:12s} {ms * 1e3:9.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)


if which in ("all", "attn"):
    qkv = (torch.randn((M, 3 * D), device="cuda") * 1.0).to(torch.bfloat16)
    out = torch.empty((M, D), dtype=torch.bfloat16, device="cuda")
    timeit(lambda: lib.aigv_attention(ptr(qkv), ptr(out), frames, S, H, d, d ** -0.5, stream()), 4.0 * frames * S * S * D, "attention")
if which in ("all", "gemm"):
    for name, N, K, gelu, f32 in (("qkv", 3 * D, D, 0, 0), ("proj", D, D, 0, 1), ("fc1", I, D, 1, 0), ("fc2", D, I, 0, 1)):
        a = torch.randn((M, K), device="cuda").to(torch.bfloat16)
        w = (torch.randn((N, K), device="cuda") * K ** -0.5).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        o = torch.zeros((M, N), dtype=torch.float32 if f32 else torch.bfloat16, device="cuda")
        fn = (lambda a=a, w=w, bias=bias, o=o, N=N, K=K, gelu=gelu, f32=f32:
              lib.aigv_gemm_epilogue(ptr(a), ptr(w), ptr(bias), ptr(bias) if f32 else None, ptr(o) if f32 else None,
                                     ptr(o) if f32 else None, None if f32 else ptr(o), M, N, K, gelu, stream()))
        timeit(fn, 2.0 * M * N * K, name)
