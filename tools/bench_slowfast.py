"""GPU box: time the SlowFast-R50 motion trunk (aigv_assessor_b200/slowfast.py) per 8-frame 448x448 clip."""
import _path  # noqa: F401  (repository root on sys.path)
import json
import sys

import torch

from aigv_assessor_b200 import slowfast, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
trunk = slowfast.SlowFastTrunk(synth.make_slowfast_weights(seed=11))
pv = torch.randn((B * 8, 3, 448, 448), device="cuda").to(torch.bfloat16)
for _ in range(3):
    f = trunk(pv, B)
torch.cuda.synchronize()
l0 = trunk.launches
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    f = trunk(pv, B)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(json.dumps({"clips": B, "ms_per_call": ms, "ms_per_clip": ms / B, "launches_per_call": (trunk.launches - l0) // 10,
                  "finite": bool(torch.isfinite(f).all())}))
if "prof" in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        trunk(pv, B)
        torch.cuda.synchronize()
    print(p.key_averages().table(sort_by="cuda_time_total", row_limit=12))
