#!/usr/bin/env python
"""GPU box: does the per-launch CUDA-event instrumentation (aigv_profile_*) slow the timed step? 10 steps of the bench workload with
the instrumentation off / on, alternating."""
import _path  # noqa: F401  (repository root on sys.path)
import sys
import torch
sys.path.insert(0, ".")
from aigv_assessor_b200 import synth
from aigv_assessor_b200.config import internvit_300m
from aigv_assessor_b200.engine import VisionEngine

cfg = internvit_300m(llm_hidden=2048)
eng = VisionEngine(cfg, max_tiles=64).load_state_dict(synth.make_weights(cfg, seed=0))
xs = [(torch.randn((64, 3, 448, 448), device="cuda")).to(torch.bfloat16) for _ in range(4)]


def run(n):
    for i in range(n):
        eng.extract_feature(xs[i % 4])


run(3)
torch.cuda.synchronize()
for rep in range(3):
    for on in (False, True):
        eng.profile(on)
        if on:
            eng.profile_read(reset=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        run(10)
        b.record()
        torch.cuda.synchronize()
        if on:
            eng.profile_read(reset=True)
        print(f"profile {'on ' if on else 'off'}: {a.elapsed_time(b) / 10:.3f} ms per step", flush=True)
eng.profile(False)
