"""GPU box: time aigv_preprocess_u8 alone (device-resident uint8 frames) for a few source sizes. Usage: python tools/pre_bench.py"""
import _path  # noqa: F401  (repository root on sys.path)
import torch
from aigv_assessor_b200.config import tiny
from aigv_assessor_b200.engine import VisionEngine

eng = VisionEngine(tiny(), max_tiles=1)
for (h, w, mx, nf) in [(720, 1280, 1, 64), (1080, 1920, 1, 32), (2160, 3840, 1, 8), (512, 768, 6, 64), (448, 448, 1, 64), (360, 640, 1, 64)]:
    u8 = torch.randint(0, 256, (nf, h, w, 3), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        pv, counts = eng.preprocess_frames(u8, max_num=mx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.profile(True)
    eng.profile_read(reset=True)
    e0.record()
    for _ in range(10):
        eng.preprocess_frames(u8, max_num=mx)
    e1.record()
    torch.cuda.synchronize()
    prof = eng.profile_read(reset=True)
    eng.profile(False)
    ms_loop = e0.elapsed_time(e1) / 10
    ms = sum(v[0] for k, v in prof.items() if k.startswith("preprocess")) / 10   # kernel time (events around each launch)
    tiles = sum(counts)
    b = nf * h * w * 3 + tiles * 3 * 448 * 448 * 2
    print(f"{w}x{h} max_num={mx}: {nf} frames -> {tiles} tiles  {ms*1e3/nf:8.2f} us/frame  {b/ms/1e6:8.1f} GB/s  {nf/ms*1e3:9.0f} frames/s   (host loop {ms_loop*1e3/nf:.2f} us/frame)")
