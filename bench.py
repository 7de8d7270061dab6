#!/usr/bin/env python
"""Headline benchmark: frames/sec of InternViT-300M encode + mlp1 projector @448^2 bf16 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A step = one pass of the hot path (extract_feature through the C-ABI) over one batch of synthetic 8-frame clips per
GPU. `value` is device-timed with inputs resident in HBM; `e2e` repeats the measurement with host (pinned) buffers:
H2D of the step's pixel_values and D2H of its visual tokens inside the timed region. Multi-GPU: one process per GPU
(torchrun), clips sharded with no data-path collective, one all-gather of per-clip pooled features per step
(weak scaling). `--impl reference` times the CPU oracle port of the reference path (oracle/vit_oracle_torch.py, torch
fp32 with all host threads) on a bounded sample of the same workload; the reference itself is Python and cannot travel to the
GPU box (/root/reference is absent there).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec InternViT-300M encode+projector @448^2 bf16"
WORKLOAD = "InternVL2-2B vision path: InternViT-300M (24L, D=1024, 16 heads, 1025 tokens/tile) + pixel_shuffle + mlp1(4096->2048->2048), 8-frame 448x448 clips"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(source="measured", burst=float(d["bf16_tflops"]), sustained=float(d["bf16_tflops_sustained"]),
                    hbm=float(d["hbm_gbs"]))
    return dict(source="fallback", burst=1590.0, sustained=1400.0, hbm=6650.0)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {getattr(nv, k): k for k in dir(nv) if k.startswith("nvmlClocksEventReason") or
                 k.startswith("nvmlClocksThrottleReason")}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if isinstance(bit, int) and bit and (mask & bit) == bit and bit & (bit - 1) == 0:
                        self.reasons.add(name.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        reasons = sorted(r for r in self.reasons if r not in ("None", "GpuIdle", "All"))
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=reasons, samples=len(s))


def gemm_flops_per_step(cfg, frames):
    S, D, I, L, C = cfg.tokens, cfg.hidden_size, cfg.intermediate_size, cfg.num_layers, cfg.llm_hidden
    M, Mp, Mo = frames * S, frames * cfg.num_patches, frames * cfg.out_tokens
    return {
        "gemm_qkv": L * 2.0 * M * D * 3 * D, "gemm_proj": L * 2.0 * M * D * D, "gemm_fc1": L * 2.0 * M * D * I,
        "gemm_fc2": L * 2.0 * M * D * I, "gemm_patch_embed": 2.0 * Mp * cfg.patch_k * D,
        "gemm_mlp1_a": 2.0 * Mo * cfg.shuffled_width * C, "gemm_mlp1_b": 2.0 * Mo * C * C,
        "attention": L * 4.0 * frames * S * S * D,
    }


def cpu_oracle_fps(cfg, sd, frames, repeats=1):
    """Time the CPU oracle port of the reference path (torch fp32, all host threads, like the reference's own CPU run)
    on `frames` tiles; returns (frames/s, seconds)."""
    import torch
    from aigv_assessor_b200 import synth
    from oracle import vit_oracle_torch as vt
    torch.set_num_threads(os.cpu_count())
    pv = torch.from_numpy(synth.make_pixel_values(frames, cfg, seed=99))
    tsd = vt.to_torch(sd)
    vt.extract_feature(pv[:1], tsd, cfg)  # warm the thread pool / allocator
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        vt.extract_feature(pv, tsd, cfg)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return frames / best, best


def reference_gpu_stack(cfg, sd, frames, dev, iters=10):
    """COMPARATOR, never the product path: the kernel stack the unmodified reference runs on this same GPU -- cuBLASLt
    bf16 GEMMs behind nn.Linear (modeling_intern_vit.py:129,141,189,190), the flash-attn wheel's sm_100 cubin behind
    flash_attn_varlen_qkvpacked_func (flash_attention.py:52-55) and ATen elementwise kernels -- timed per op at this
    bench's shapes and as the whole eager bf16 extract_feature path (the torch restatement of the reference's modules,
    oracle/vit_oracle_torch.py, on CUDA with FA2). SURVEY.md section 8(d) "beat this" row."""
    import torch
    import torch.nn.functional as F
    from oracle import vit_oracle_torch as vt
    res = {}
    S, D, I, H, hd, L = cfg.tokens, cfg.hidden_size, cfg.intermediate_size, cfg.num_heads, cfg.head_dim, cfg.num_layers
    M = frames * S

    def timeit(fn, n=iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    g = torch.Generator(device="cpu").manual_seed(7)
    for name, N, K in (("qkv", 3 * D, D), ("proj", D, D), ("fc1", I, D), ("fc2", D, I)):
        a = torch.randn((M, K), generator=g).to(dev, torch.bfloat16)
        w = (torch.randn((N, K), generator=g) * K ** -0.5).to(dev, torch.bfloat16)
        bias = torch.randn(N, generator=g).to(dev, torch.bfloat16)
        ms = timeit(lambda: F.linear(a, w, bias))
        res[f"gemm_{name}_ms_per_step"] = ms * L
        res[f"gemm_{name}_tflops"] = 2.0 * M * N * K / ms / 1e9
        del a, w, bias
    try:
        from flash_attn import flash_attn_varlen_qkvpacked_func
        qkv = torch.randn((M, 3, H, hd), generator=g).to(dev, torch.bfloat16)
        cu = torch.arange(0, (frames + 1) * S, step=S, dtype=torch.int32, device=dev)
        ms = timeit(lambda: flash_attn_varlen_qkvpacked_func(qkv, cu, S, 0.0, softmax_scale=None, causal=False))
        res["attention_ms_per_step"] = ms * L
        res["attention_tflops"] = 4.0 * frames * S * S * D / ms / 1e9
        res["attention_kernel"] = "flash_attn_varlen_qkvpacked_func (flash-attn wheel, sm_100 cubin: mma.sync + cp.async)"
        del qkv
        have_fa = True
    except Exception as e:  # pragma: no cover
        res["attention_unavailable"] = repr(e)[:200]
        have_fa = False
    # whole path, eager bf16 (+FA2): the reference's own GPU configuration (stage2_eval.py:932, use_flash_attn=True)
    tsd = {k: torch.from_numpy(v).to(dev, torch.bfloat16) for k, v in sd.items()}
    if have_fa:
        tsd["__use_flash_attn__"] = torch.ones(1)
    from aigv_assessor_b200 import synth
    pv = torch.from_numpy(synth.make_pixel_values(8, cfg, seed=99)).to(dev, torch.bfloat16).repeat(frames // 8, 1, 1, 1)
    ms = timeit(lambda: vt.extract_feature(pv, tsd, cfg), n=3)
    res["eager_bf16_ms_per_step"] = ms
    res["eager_bf16_frames_per_s"] = frames / ms * 1e3
    res["eager_attention"] = "flash-attn 2" if have_fa else "naive softmax (flash-attn unavailable)"
    res["note"] = ("library kernels of the reference's own GPU path on this GPU (cuBLASLt, flash-attn wheel, ATen), timed "
                   "alone per op at M = %d rows and as the eager bf16 path over %d frames; a comparator, not used by the product" % (M, frames))
    return res


def run_reference(args, rank, world, out):
    """CPU arm: the oracle port of the reference's path, all host threads, bounded sample per step."""
    if rank != 0:
        return
    import torch
    from aigv_assessor_b200 import synth
    from aigv_assessor_b200.config import internvit_300m
    from oracle import vit_oracle_torch as vt
    torch.set_num_threads(os.cpu_count())
    cfg = internvit_300m(llm_hidden=2048)
    sd = vt.to_torch(synth.make_weights(cfg, seed=0))
    sample = args.cpu_sample_frames
    pv = torch.from_numpy(synth.make_pixel_values(sample, cfg, seed=99))
    for _ in range(min(args.warmup, 1)):
        vt.extract_feature(pv, sd, cfg)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vt.extract_feature(pv, sd, cfg)
    dt = time.perf_counter() - t0
    fps = sample * args.steps / dt
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": sample,
                   "note": "CPU oracle port (torch fp32, MKL) of the reference extract_feature path; each step is a "
                           f"{sample}-frame sample of the workload"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} frame(s) per step x {args.steps} steps, torch CPU fp32, {cores} threads"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    out.emit(json.dumps(line))


def build_6b_engine(dev, frames):
    """Config 5: InternViT-6B (D=3200, 25 heads x 128, I=12800, 48 layers, RMSNorm + q/k RMSNorm) + mlp1 -> 6144 with
    random-init weights generated ON the device from synth.weight_spec (5.9 G parameters)."""
    import torch
    from aigv_assessor_b200 import synth
    from aigv_assessor_b200.config import internvit_6b
    from aigv_assessor_b200.engine import VisionEngine
    cfg = internvit_6b()
    eng = VisionEngine(cfg, max_tiles=frames, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    for name, shape, dist_ in synth.weight_spec(cfg):
        if dist_[0] == "uniform":
            t = (torch.rand(shape, generator=g, device=dev) * 2 - 1) * dist_[1] ** -0.5
        else:
            t = dist_[2] + dist_[1] * torch.randn(shape, generator=g, device=dev)
        eng.set_weight(name, t.to(torch.bfloat16))
        del t
    eng.load_state_dict({})  # finalize
    return cfg, eng


def run_native(args, rank, local_rank, world, out):
    import numpy as np
    import torch
    import torch.distributed as dist

    from aigv_assessor_b200 import synth
    from aigv_assessor_b200.config import internvit_300m
    from aigv_assessor_b200.engine import VisionEngine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = internvit_300m(llm_hidden=2048)
    frames = args.frames_per_step
    clip = 8
    assert frames % clip == 0
    sd = synth.make_weights(cfg, seed=0)
    eng = VisionEngine(cfg, max_tiles=min(frames, args.max_tiles)).load_state_dict(sd)

    # synthetic inputs: R rotating batches so a step never finds its input in L2 (and the 2 GB activation working
    # set per step exceeds the 126 MB L2 by itself)
    R = 4
    lut = torch.from_numpy(synth.normalize_lut())
    g = torch.Generator().manual_seed(1234 + rank)
    host = []
    for _ in range(R):
        u = torch.randint(0, 256, (frames, 3, cfg.image_size, cfg.image_size), generator=g, dtype=torch.uint8)
        pv = torch.stack([lut[c][u[:, c].long()] for c in range(3)], dim=1).to(torch.bfloat16)
        host.append(pv.pin_memory())
    devs = [h.to(dev) for h in host]
    out_host = torch.empty((frames, cfg.out_tokens, cfg.llm_hidden), dtype=torch.bfloat16).pin_memory()
    n_clips = frames // clip

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    gathered = None
    if world > 1:
        gathered = torch.empty((world * frames, cfg.out_tokens, cfg.llm_hidden), dtype=torch.bfloat16, device=dev)

    def timed_steps(n, batches, e=eng):
        """n passes of the hot path; clips are sharded over the ranks with NO data-path collective, and the path's only
        exchange -- one all-gather of this rank's per-clip visual tokens -- happens once, at the end."""
        o = None
        for i in range(n):
            o = e.extract_feature(batches[i % len(batches)])
        if world > 1 and o is not None and o.shape[0] == frames:
            dist.all_gather_into_tensor(gathered, o)
        return o

    timed_steps(args.warmup, devs)
    barrier()

    # ---- timed region: device-resident inputs, per-kernel CUDA events on the launching stream -------------------
    eng.profile(True)
    eng.profile_read(reset=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    timed_steps(args.steps, devs)
    e1.record()
    barrier()
    ms_local = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = eng.launch_count - l0
    prof = eng.profile_read(reset=True)
    eng.profile(False)

    # ---- e2e: uint8 video frames in pinned HOST memory -> front end -> extract_feature -> visual tokens on the HOST ----
    # The public API runs on the caller's stream, so a caller pipelines exactly like this: a copy-in stream, the
    # compute stream and a copy-out stream, double buffered, chained with events. Every step moves its own frames
    # host->device and its visual tokens device->host inside the timed region, and runs the frame front end.
    fh, fw = 720, 1280
    ms_e2e_local = float("nan")
    h2d_bytes = frames * fh * fw * 3
    if not args.no_e2e:
        gf = torch.Generator().manual_seed(77 + rank)
        host_u8 = [torch.randint(0, 256, (frames, fh, fw, 3), generator=gf, dtype=torch.uint8).pin_memory() for _ in range(2)]
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        s_cmp = torch.cuda.current_stream()
        stage = [torch.empty((frames, fh, fw, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
        outs = [torch.empty((frames, cfg.out_tokens, cfg.llm_hidden), dtype=torch.bfloat16, device=dev)
                for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        ev_copied = [torch.cuda.Event() for _ in range(2)]

        def e2e_run(n):
            for b in range(2):
                ev_free[b].record(s_cmp)
                ev_copied[b].record(s_out)
            for i in range(n):
                b = i & 1
                with torch.cuda.stream(s_in):
                    s_in.wait_event(ev_free[b])          # stage[b] no longer read by step i-2
                    stage[b].copy_(host_u8[i & 1], non_blocking=True)
                    ev_in[b].record(s_in)
                s_cmp.wait_event(ev_in[b])
                s_cmp.wait_event(ev_copied[b])           # outs[b] of step i-2 has left the device
                pvb, _ = eng.preprocess_frames(stage[b], max_num=1)   # resize + normalise + bf16 on the device
                eng.extract_feature(pvb, out=outs[b])
                ev_free[b].record(s_cmp)
                ev_out[b].record(s_cmp)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_out[b])
                    out_host.copy_(outs[b], non_blocking=True)
                    ev_copied[b].record(s_out)
            s_cmp.wait_stream(s_out)

        e2e_run(min(args.warmup, 2))
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        e2e_run(args.steps)
        f1.record()
        barrier()
        ms_e2e_local = f0.elapsed_time(f1)
        del stage, outs, host_u8

    def timed(fn, n):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b) / n)

    extra = {}
    # ---- config 3 shape: 32 frames (4 clips) per GPU per step --------------------------------------------------------
    if not args.no_extra and frames >= 32:
        half = [d[:32] for d in devs]
        timed_steps(2, half)
        ms32 = timed(lambda: timed_steps(1, half), 8)
        extra["config3_32_frames_per_gpu"] = {"frames_per_s": world * 32 / ms32 * 1e3, "ms_per_step": ms32,
                                              "note": "BASELINE config 3 shape: 32 clips per step over 8 GPUs = 4 clips (32 tiles, M = 32800) per GPU"}
    # ---- config 4: frames -> tiles -> tokens incl. the GPU front end, 16/32/64 frames x 1/3/7 tiles per frame -------
    if not args.no_extra:
        sweep = []
        for (sh, sw, mx, tiles) in ((720, 1280, 1, 1), (720, 1280, 6, 3), (512, 768, 6, 7)):
            for nf in (16, 32, 64):
                u8 = torch.randint(0, 256, (nf, sh, sw, 3), dtype=torch.uint8, device=dev)
                eng.encode_frames(u8, max_num=mx)
                ms_c = timed(lambda: eng.encode_frames(u8, max_num=mx), 2)
                sweep.append({"frames_per_clip": nf, "tiles_per_frame": tiles, "source": f"{sw}x{sh}",
                              "tiles_per_s": world * nf * tiles / ms_c * 1e3, "clips_per_s": world / ms_c * 1e3,
                              "ms_per_clip": ms_c})
                del u8
        extra["config4_sweep"] = {"rows": sweep, "note": "uint8 frames resident on the device -> dynamic tiling front end (max_num = 1 / 6 incl. "
                                  "thumbnail) -> extract_feature, micro-batched at %d tiles; one clip per GPU per pass" % eng.max_tiles}

    # ---- HBM-bound kernels of the path: achieved GB/s against the measured copy bandwidth ---------------------------
    bw = {}
    if rank == 0:
        nf = frames
        u8 = torch.randint(0, 256, (nf, fh, fw, 3), dtype=torch.uint8, device=dev)   # 177 MB > L2
        for _ in range(2):
            eng.preprocess_frames(u8, max_num=1)
        torch.cuda.synchronize()
        eng.profile(True)   # kernel time from the library's CUDA events around each launch (the python call costs ~100 us)
        eng.profile_read(reset=True)
        for _ in range(5):
            eng.preprocess_frames(u8, max_num=1)
        torch.cuda.synchronize()
        pre_ms = sum(v[0] for k, v in eng.profile_read(reset=True).items() if k.startswith("preprocess")) / 5
        eng.profile(False)
        pre_bytes = nf * (fh * fw * 3 + 3 * cfg.image_size ** 2 * 2)
        bw["preprocess_u8"] = {"GBps": pre_bytes / pre_ms / 1e6, "bytes_per_frame": pre_bytes / nf,
                               "frames_per_s": nf / pre_ms * 1e3, "source": f"{fh}x{fw} uint8 -> 448x448 bf16 tile",
                               "bound": "instruction issue (one PRMT + one IMAD per byte-tap of Pillow's 22-bit fixed-point "
                                        "resampler), not HBM; see profiles/r2e_ncu_preprocess_summary.txt"}
        del u8
        M_rows = frames * cfg.tokens
        if "norm" in prof:
            b = M_rows * cfg.hidden_size * 6
            bw["norm"] = {"GBps": b * prof["norm"][1] / prof["norm"][0] / 1e6, "bytes_per_launch": b}
        if "pixel_shuffle_ln" in prof:
            b = frames * cfg.num_patches * cfg.hidden_size * 4 + frames * cfg.out_tokens * cfg.shuffled_width * 2
            bw["pixel_shuffle_ln"] = {"GBps": b * prof["pixel_shuffle_ln"][1] / prof["pixel_shuffle_ln"][0] / 1e6,
                                      "bytes_per_launch": b}

    # ---- config 5: InternViT-6B replicas (one per GPU) ----------------------------------------------------------------
    if not args.no_extra and not args.no_6b:
        try:
            f6 = 16
            del devs
            torch.cuda.empty_cache()
            cfg6, eng6 = build_6b_engine(dev, f6)
            pv6 = torch.randn((f6, 3, cfg6.image_size, cfg6.image_size), device=dev).to(torch.bfloat16)
            for _ in range(2):
                o6 = eng6.extract_feature(pv6)
            eng6.profile(True)
            eng6.profile_read(reset=True)
            ms6 = timed(lambda: eng6.extract_feature(pv6), 3)
            prof6 = eng6.profile_read(reset=True)
            tf6 = cfg6.flops_per_frame() * f6 / ms6 / 1e9
            extra["config5_internvit_6b"] = {
                "frames_per_s": world * f6 / ms6 * 1e3, "ms_per_step": ms6, "frames_per_step_per_gpu": f6,
                "tflops_per_gpu": tf6, "frac_of_sustained_peak": tf6 / measured_peaks()["sustained"],
                "gflop_per_frame": cfg6.flops_per_frame() / 1e9, "finite": bool(torch.isfinite(o6.float()).all()),
                "kernels_ms_per_step": {k: v[0] / 3 for k, v in sorted(prof6.items(), key=lambda kv: -kv[1][0])},
                "note": "InternViT-6B 48L (D=3200, 25x128, I=12800, RMSNorm + q/k RMSNorm) + mlp1 -> 6144, random-init weights "
                        "generated on the device, one replica per GPU (weak scaling)"}
            eng6.close()
            del eng6, pv6
        except Exception as e:  # a side measurement must never take the bench line down
            extra["config5_internvit_6b"] = {"unavailable": repr(e)[:300]}

    # ---- next-1: InternLM2-7B prefill of one 8-frame clip prompt (2108 tokens), one replica per GPU ----------
    if not args.no_extra and not args.no_llm:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_llm as tbl
            torch.cuda.empty_cache()
            n_tok = 1 + 8 * 258 + 3 + 40      # <s>, 8 x (<img> 256 x <IMG_CONTEXT> </img>), the motion slot, 40 text tokens
            c7, lm = tbl.build("7b", 4 * n_tok)   # workspace for the 4-clip step of config 3 below
            x7 = (torch.randn((1, n_tok, c7["hidden_size"]), device=dev) * 0.5).to(torch.bfloat16)
            tail = list(range(n_tok - 49, n_tok - 1))

            def clip_tail():
                lm.prefill(x7)
                lm.hidden_rows([n_tok - 4])
                return lm.logits_rows(tail)
            for _ in range(2):
                lg = clip_tail()
            ms7 = timed(clip_tail, 5)
            fl7 = tbl.flops_per_prompt(c7, n_tok)
            extra["next1_llm_prefill_7b"] = {
                "prompts_per_s": world / ms7 * 1e3, "ms_per_prompt": ms7, "tokens": n_tok, "tflops_per_gpu": fl7 / ms7 / 1e9,
                "frac_of_sustained_peak": fl7 / ms7 / 1e9 / measured_peaks()["sustained"], "tflop_per_prompt": fl7 / 1e12,
                "finite": bool(torch.isfinite(lg).all()), "launches_per_prompt": 1 + 5 * c7["num_hidden_layers"] + 3,
                "note": "InternLM2-7B decoder stack (32L, 4096, 32/8 heads of 128, I=14336) on the spliced prompt of one 8-frame "
                        "clip: RMSNorm-fused wqkv with the rotary embedding in its epilogue, causal GQA tcgen05 attention, wo + residual, w1|w3 + SwiGLU epilogue, w2 + "
                        "residual; final norm + output head on row -4 and the 48 answer rows only; random-init weights on the device"}
            if world == 1 and not args.no_ref_gpu:
                try:
                    extra["next1_llm_prefill_7b"]["reference_gpu_stack"] = tbl.reference_gpu_stack(c7, 1, n_tok)
                except Exception as e:
                    extra["next1_llm_prefill_7b"]["reference_gpu_stack"] = {"unavailable": repr(e)[:200]}
            # frames -> score for one clip: pinned uint8 frames on the host -> front end -> InternViT-300M -> splice ->
            # InternLM2-7B prefill -> mlpscore, score read back to the host (the product metric of the reference: one
            # quality score per video). The vision tower's projector is widened to the LLM's 4096 for this leg.
            try:
                import numpy as np
                from aigv_assessor_b200 import llm_glue
                from aigv_assessor_b200.config import VisionPathConfig
                cfg_s = VisionPathConfig(llm_hidden=c7["hidden_size"])
                eng_s = VisionEngine(cfg_s, max_tiles=32, device=dev).load_state_dict(synth.make_weights(cfg_s, seed=3))
                hsd = synth.make_head_weights(seed=5)
                head = llm_glue.MlpScoreHead(hsd, dev)
                from aigv_assessor_b200.scoring import MotionMlp
                from aigv_assessor_b200.slowfast import SlowFastTrunk
                trunk = SlowFastTrunk(synth.make_slowfast_weights(seed=6), device=dev)
                motion = MotionMlp(hsd, dev)
                tok = (torch.randn((c7["vocab_size"], c7["hidden_size"]), device=dev) * 0.5).to(torch.bfloat16)
                ids = torch.from_numpy(synth.make_clip_prompt(8, seed=1, vocab_size=c7["vocab_size"])).to(dev)
                clip_host = torch.randint(0, 256, (8, fh, fw, 3), dtype=torch.uint8).pin_memory()

                from aigv_assessor_b200.scoring import ClipScorer
                scorer = ClipScorer.from_parts(eng_s, lm, motion, head, tok, synth.IMG_CONTEXT_ID, slowfast=trunk)
                scorer.check_splice = False   # the prompt is known good: no extra host sync per clip

                def clip_to_score():
                    fr = clip_host.to(dev, non_blocking=True)
                    pv8, _ = eng_s.preprocess_frames(fr, max_num=1)
                    out_s = scorer(pv8, ids)                  # ViT || SlowFast trunk + motion_mlp -> splice -> prefill -> heads
                    return float(out_s["score1"][0].item())   # D2H of the score: the step's result reaches the host
                for _ in range(2):
                    sc0 = clip_to_score()
                ms_c = timed(clip_to_score, 5)
                extra["clip_to_score_7b"] = {
                    "clips_per_s": world / ms_c * 1e3, "ms_per_clip": ms_c, "frames_per_clip": 8, "finite": bool(np.isfinite(sc0)),
                    "h2d_bytes_per_clip": int(clip_host.numel()), "d2h_bytes_per_clip": 4,
                    "note": "one 8-frame 1280x720 uint8 clip on the host -> H2D -> aigv_preprocess_u8 -> InternViT-300M + projector "
                            "(4096 wide) + SlowFast-R50 trunk + motion_mlp -> aigv_splice_tokens -> InternLM2-7B prefill (2108 tokens) -> "
                            "final norm on row -4 -> aigv_mlpscore -> score on the host; synchronous per clip (no overlap between clips)"}
                # BASELINE config 3: 32 clips per step over 8 GPUs = 4 clips per GPU per step, one batched forward: 32 tiles through
                # the front end / ViT / SlowFast trunk, one [4, 2108] prefill, 4 scores back to the host
                try:
                    ids4 = ids.repeat(4, 1).contiguous()
                    clips4_host = torch.randint(0, 256, (32, fh, fw, 3), dtype=torch.uint8).pin_memory()

                    def clips4_to_scores():
                        fr = clips4_host.to(dev, non_blocking=True)
                        pv32, _ = eng_s.preprocess_frames(fr, max_num=1)
                        return scorer(pv32, ids4)["score1"].cpu()
                    for _ in range(2):
                        sc4 = clips4_to_scores()
                    ms_c4 = timed(clips4_to_scores, 3)
                    extra["config3_clips_to_score_7b"] = {
                        "clips_per_s": world * 4 / ms_c4 * 1e3, "ms_per_step": ms_c4, "clips_per_step_per_gpu": 4,
                        "ms_per_clip": ms_c4 / 4, "finite": bool(torch.isfinite(sc4).all()), "scores_per_step": int(sc4.numel()),
                        "h2d_bytes_per_step": int(clips4_host.numel()), "d2h_bytes_per_step": 4 * int(sc4.numel()),
                        "note": "BASELINE config 3 shape (32 clips per step sharded over 8 GPUs): 4 clips per GPU in ONE forward -- 32 uint8 "
                                "720p frames on the host -> H2D -> front end -> InternViT-300M (32 tiles) || SlowFast trunk (4 clips) + "
                                "motion_mlp -> splice -> InternLM2-7B prefill of [4, 2108] -> mlpscore -> 4 scores on the host"}
                except Exception as e:
                    extra["config3_clips_to_score_7b"] = {"unavailable": repr(e)[:300]}
                eng_s.close()
                del eng_s, tok, trunk
            except Exception as e:
                extra["clip_to_score_7b"] = {"unavailable": repr(e)[:300]}
            lm.close()
            del lm, x7
            torch.cuda.empty_cache()
            # BASELINE configs[1]: InternVL2-2B stage-1 eval forward (InternViT-300M + projector -> InternLM2-1.8B prefill ->
            # quality-level logits of the answer span), one 8-frame clip at a time, frames on the host -> token ids on the host
            try:
                from aigv_assessor_b200 import llm_glue
                c2, lm2 = tbl.build("1_8b", n_tok)
                eng2 = VisionEngine(cfg, max_tiles=8, device=dev).load_state_dict(sd)     # the bench's InternVL2-2B vision path
                tok2 = (torch.randn((c2["vocab_size"], c2["hidden_size"]), device=dev) * 0.5).to(torch.bfloat16)
                mot2 = (torch.randn((1, c2["hidden_size"]), device=dev) * 0.5).to(torch.bfloat16)   # the motion token (resident)
                ids2 = torch.from_numpy(synth.make_clip_prompt(8, seed=1, vocab_size=c2["vocab_size"])).to(dev)
                clip2_host = torch.randint(0, 256, (8, fh, fw, 3), dtype=torch.uint8).pin_memory()
                N2 = ids2.shape[1]
                rows2 = list(range(N2 - 49, N2 - 1))

                def clip_to_logits():
                    fr = clip2_host.to(dev, non_blocking=True)
                    pv8, _ = eng2.preprocess_frames(fr, max_num=1)
                    x = llm_glue.splice_tokens(ids2, tok2, eng2.extract_feature(pv8), mot2, synth.IMG_CONTEXT_ID, check=False)
                    lm2.prefill(x)
                    return torch.argmax(lm2.logits_rows(rows2), dim=1).cpu()
                for _ in range(2):
                    a2 = clip_to_logits()
                ms_l = timed(clip_to_logits, 5)
                extra["config1_stage1_clip_to_logits_1_8b"] = {
                    "clips_per_s": world / ms_l * 1e3, "ms_per_clip": ms_l, "frames_per_clip": 8, "tokens": int(N2),
                    "answer_ids": int(a2.numel()), "h2d_bytes_per_clip": int(clip2_host.numel()), "d2h_bytes_per_clip": 8 * int(a2.numel()),
                    "note": "BASELINE configs[1] shape: one 8-frame 1280x720 uint8 clip on the host -> H2D -> front end -> InternViT-300M + "
                            "projector (2048 wide) -> splice -> InternLM2-1.8B prefill (24L, 2048, 16/8 heads of 128) -> output head on the 48 "
                            "answer rows -> argmax -> token ids on the host; synchronous per clip"}
                lm2.close()
                eng2.close()
                del lm2, eng2, tok2
                torch.cuda.empty_cache()
            except Exception as e:
                extra["config1_stage1_clip_to_logits_1_8b"] = {"unavailable": repr(e)[:300]}
        except Exception as e:  # a side measurement must never take the bench line down
            extra["next1_llm_prefill_7b"] = {"unavailable": repr(e)[:300]}

    ms = max_over_ranks(ms_local)
    ms_e2e = max_over_ranks(ms_e2e_local)
    rank_ms = [ms_local / args.steps]
    if world > 1:
        t = torch.tensor([ms_local / args.steps], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        rank_ms = [float(x[0]) for x in allt]
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt[0])

    if rank == 0:
        peaks = measured_peaks()
        total_frames = world * frames * args.steps
        value = total_frames / (ms / 1e3)
        e2e = total_frames / (ms_e2e / 1e3)
        fl = gemm_flops_per_step(cfg, frames)
        gemm_ms = sum(prof[k][0] for k in prof if k.startswith("gemm_"))
        gemm_fl = sum(fl[k] for k in fl if k.startswith("gemm_")) * args.steps
        gemm_launches = sum(prof[k][1] for k in prof if k.startswith("gemm_"))
        achieved = gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms else None
        kernels = {}
        tot_ms = sum(v[0] for v in prof.values())
        for k, (kms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
            ent = {"ms_per_step": kms / args.steps, "launches_per_step": cnt / args.steps, "share": kms / tot_ms}
            if k in fl:
                ent["tflops"] = fl[k] * args.steps / (kms / 1e3) / 1e12
            kernels[k] = ent
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("gemm_dram_bytes_per_launch")
        att = kernels.get("attention", {})
        # attention's own floor: one ex2 per score on the 16-lane/clk/SM MUFU at the clock seen under load
        exps = cfg.num_layers * frames * cfg.num_heads * cfg.tokens * cfg.tokens
        mufu_floor_ms = exps / (148 * 16 * (clocks["sm_mhz"] or 1965) * 1e6) * 1e3 if att else None
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": frames, "clips_per_step_per_gpu": n_clips,
                       "micro_batch_tiles": eng.max_tiles, "parallelism": f"dp{world} (clips sharded, replicas of 0.6 GB weights; "
                       "one all-gather of the visual tokens at the end of the timed region)",
                       "l2": f"{R} rotating input batches; ~{eng.lib.aigv_workspace_bytes(eng._h, eng.max_tiles) / 2**30:.1f} GiB "
                             "activation working set per step >> 126 MB L2",
                       "gflop_per_frame": cfg.flops_per_frame() / 1e9,
                       "timed_region": "every launch of the timed steps is bracketed by CUDA events on the launching stream (the "
                                       "roofline's per-kernel times); the events cost ~1% of the step (profiles/r2z_pdl_ab.txt: 46.7 ms "
                                       "with them, 46.3 without) and keep programmatic dependent launch from overlapping kernel "
                                       "prologues there; the e2e leg runs without them"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": out_host.numel() * 2, "ms_per_step": ms_e2e / args.steps,
                    "path": f"pinned uint8 {fh}x{fw} RGB frames -> H2D -> aigv_preprocess_u8 (Pillow-exact resize, normalise, bf16) -> "
                            "aigv_extract_feature -> D2H of the visual tokens; copies double buffered on side streams"},
            "gpu_launches": launches,
            "rank_ms_per_step": rank_ms,
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16_tn_kernel (tcgen05 cta_group::2, all GEMM launches of the timed steps)",
                         "achieved": achieved, "peak": peaks["sustained"], "unit": "TFLOP/s",
                         "frac": (achieved / peaks["sustained"]) if achieved else None, "traffic": traffic,
                         "peak_source": f"{peaks['source']} (sustained cuBLAS bf16; burst {peaks['burst']})",
                         "launches": gemm_launches, "avg_launch_ms": gemm_ms / max(gemm_launches, 1),
                         "flop_per_launch_avg": gemm_fl / max(gemm_launches, 1),
                         "whole_step_frac": value / world * cfg.flops_per_frame() / 1e12 / peaks["sustained"]},
            "roofline_attention": ({"bound": "tensor + MUFU(ex2)", "kernel": "attention_fwd2_kernel<64> (tcgen05, persistent)",
                                    "achieved": att.get("tflops"), "peak": peaks["sustained"], "unit": "TFLOP/s",
                                    "frac": att["tflops"] / peaks["sustained"], "ms_per_step": att["ms_per_step"],
                                    "mufu_floor_ms_per_step": mufu_floor_ms,
                                    "frac_of_mufu_floor": mufu_floor_ms / att["ms_per_step"]} if att else None),
            "kernels": kernels,
            "bandwidth_kernels": {k: dict(v, frac_of_hbm_peak=v["GBps"] / peaks["hbm"]) for k, v in bw.items()},
            "extra": extra,
        }
        if world == 1 and not args.no_ref_gpu:
            try:
                line["reference_gpu_stack"] = reference_gpu_stack(cfg, sd, frames, dev)
            except Exception as e:  # the comparator must never take the bench line down
                line["reference_gpu_stack"] = {"unavailable": repr(e)[:300]}
        if world == 1 and not args.no_cpu_baseline:
            fps, secs = cpu_oracle_fps(cfg, sd, args.cpu_baseline_frames)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{args.cpu_baseline_frames} frame(s) of the same workload through the torch-CPU "
                                              f"fp32 oracle port ({secs:.1f} s, {os.cpu_count()} threads)"}
        out.emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


class OnlyJsonOnStdout:
    """Libraries (NCCL's version banner, torchrun warnings) write to fd 1; the driver wants exactly ONE JSON line there.
    Everything written to stdout while this is active goes to stderr; emit() writes to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self._real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line: str):
        os.write(self._real, (line + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._real, 1)
        os.close(self._real)
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=64, help="tiles per GPU per step (8 clips of 8 frames)")
    ap.add_argument("--max-tiles", type=int, default=64, help="micro-batch the workspace is sized for")
    ap.add_argument("--cpu-sample-frames", type=int, default=2, help="--impl reference: frames per step")
    ap.add_argument("--cpu-baseline-frames", type=int, default=16, help="native arm: frames timed on the CPU oracle")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-GPU-stack comparator (cuBLASLt / FA2)")
    ap.add_argument("--no-extra", action="store_true", help="skip the config 3 / 4 / 5 side measurements")
    ap.add_argument("--no-6b", action="store_true", help="skip the InternViT-6B (config 5) side measurement")
    ap.add_argument("--no-llm", action="store_true", help="skip the InternLM2-7B prefill (next-1) side measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    with OnlyJsonOnStdout() as out:
        if args.impl == "reference":
            run_reference(args, rank, world, out)
        else:
            run_native(args, rank, local_rank, world, out)


if __name__ == "__main__":
    main()
