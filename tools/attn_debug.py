#!/usr/bin/env python
"""GPU box: per-(frame, head, 128-row tile) error map of aigv_attention against the fp32 softmax reference."""
import _path  # noqa: F401  (repository root on sys.path)
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from aigv_assessor_b200 import _lib
from tests.gpu_util import ptr, stream

lib = _lib.load()
cases = [(2, 1025, 2, 64), (1, 384, 1, 64), (1, 385, 1, 64), (1, 512, 1, 64), (1, 1024, 1, 64), (2, 1025, 2, 128)]
if len(sys.argv) > 1:
    cases = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for frames, S, H, d in cases:
    g = torch.Generator().manual_seed(S + H + d)
    qkv = (torch.randn((frames * S, 3 * H * d), generator=g) * 1.5).to(torch.bfloat16)
    out = torch.full((frames * S, H * d), float("nan"), dtype=torch.bfloat16, device="cuda")
    dq = qkv.cuda()
    st = lib.aigv_attention(ptr(dq), ptr(out), frames, S, H, d, d ** -0.5, stream())
    torch.cuda.synchronize()
    q, k, v = qkv.float().cuda().view(frames, S, 3, H, d).permute(2, 0, 3, 1, 4)
    a = (q * d ** -0.5) @ k.transpose(-2, -1)
    ref = (a.softmax(-1) @ v).transpose(1, 2).reshape(frames, S, H, d)
    y = out.float().view(frames, S, H, d)
    err = (y - ref).abs().amax(-1) / ref.abs().max()   # [frames, S, H]
    nan = torch.isnan(y).any(-1)
    print(f"== frames {frames} S {S} H {H} d {d}: status {st} overall e_max {float(err[~nan].max()) if (~nan).any() else -1:.3e} nan rows {int(nan.sum())}")
    for f in range(frames):
        for h in range(H):
            row = []
            for t0 in range(0, S, 128):
                e = err[f, t0:t0 + 128, h]
                n = nan[f, t0:t0 + 128, h]
                row.append(("%.0e" % float(e[~n].max()) if (~n).any() else "-") + ("!" if n.any() else ""))
            print(f"  f{f} h{h}: " + " ".join(row))
