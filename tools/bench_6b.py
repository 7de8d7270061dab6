#!/usr/bin/env python
"""Config 5: the InternViT-6B tower (D=3200, 25 heads x 128, I=12800, RMSNorm + q/k RMSNorm, no qkv bias)
with its projector through the same library, timed on one B200. Weights are random-init, generated ON the GPU from
synth.weight_spec (5.9 G parameters: numpy on the host would take minutes). This is a side measurement (it prints one
JSON line with the same keys as bench.py's kernels table); bench.py stays the contract benchmark of the 300M workload.

Usage (GPU box): python tools/bench_6b.py [--frames 16] [--layers 48] [--steps 5]
"""
import _path  # noqa: F401  (repository root on sys.path)
import argparse
import json
import sys

import torch

from aigv_assessor_b200 import config as cfgmod
from aigv_assessor_b200 import synth
from aigv_assessor_b200.engine import VisionEngine


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--layers", type=int, default=48)
    ap.add_argument("--llm-hidden", type=int, default=6144)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    cfg = cfgmod.internvit_6b(llm_hidden=args.llm_hidden, num_layers=args.layers)
    torch.cuda.set_device(0)
    eng = VisionEngine(cfg, max_tiles=args.frames)
    g = torch.Generator(device="cuda").manual_seed(0)
    for name, shape, dist in synth.weight_spec(cfg):
        if dist[0] == "uniform":
            b = dist[1] ** -0.5
            t = (torch.rand(shape, generator=g, device="cuda") * 2 - 1) * b
        else:
            t = dist[2] + dist[1] * torch.randn(shape, generator=g, device="cuda")
        eng.set_weight(name, t.to(torch.bfloat16))
        del t
    eng.load_state_dict({})  # finalize
    pv = torch.randn((args.frames, 3, cfg.image_size, cfg.image_size), generator=g, device="cuda").to(torch.bfloat16)
    out = None
    for _ in range(args.warmup):
        out = eng.extract_feature(pv)
    torch.cuda.synchronize()
    eng.profile(True)
    eng.profile_read(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = eng.extract_feature(pv)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    prof = eng.profile_read(reset=True)
    finite = bool(torch.isfinite(out.float()).all())
    fl = cfg.flops_per_frame() * args.frames
    line = {"workload": f"InternViT-6B ({args.layers}L, D=3200, 25x128) + mlp1 -> {args.llm_hidden}", "frames": args.frames,
            "ms_per_step": ms, "frames_per_s": args.frames / ms * 1e3, "tflops": fl / ms / 1e9,
            "gflop_per_frame": cfg.flops_per_frame() / 1e9, "finite": finite,
            "out_absmax": float(out.float().abs().max()),
            "kernels_ms_per_step": {k: v[0] / args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}}
    print(json.dumps(line))
    return 0 if finite else 1


if __name__ == "__main__":
    sys.exit(main())
