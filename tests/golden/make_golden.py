"""Generate the golden fixtures in this directory by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference has no tests/golden vectors of its own, so these fixtures are what pins the oracle
(oracle/vit_oracle.py, oracle/preprocess_oracle.py) and, through it, the CUDA path. Weights and inputs are NOT
stored: they are regenerated from seeds by aigv_assessor_b200.synth (numpy PCG64, platform independent).

Import shims: timm.DropPath, pytorchvideo, peft, decord, imageio are absent in this image and are
stubbed; use_flash_attn=False selects the reference's own _naive_attn CPU path.
"""
from __future__ import annotations

import hashlib
import importlib.machinery
import json
import os
import sys
import types

import numpy as np
import torch
import transformers  # noqa: F401  (must be imported before the stubs are installed)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
sys.path.insert(0, REF)

from aigv_assessor_b200 import config as cfgmod, synth  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _DropPath(torch.nn.Module):
    def __init__(self, p=0.0):
        super().__init__()

    def forward(self, x):
        return x


def install_stubs():
    _stub("timm"); _stub("timm.models"); _stub("timm.models.layers", DropPath=_DropPath)
    _stub("pytorchvideo"); _stub("pytorchvideo.models"); _stub("pytorchvideo.models.hub", slowfast_r50=lambda **k: None)
    _stub("peft", LoraConfig=object, get_peft_model=lambda *a, **k: None)
    _stub("decord", VideoReader=object, cpu=lambda *a: None)
    _stub("imageio")


def build_reference(cfg, sd_np, dtype=torch.float32):
    from internvl.model.internvl_chat.configuration_intern_vit import InternVisionConfig
    from internvl.model.internvl_chat.modeling_intern_vit import InternVisionModel
    from internvl.model.internvl_chat_eval2 import modeling_internvl_chat as EVAL2

    vc = InternVisionConfig(hidden_size=cfg.hidden_size, num_attention_heads=cfg.num_heads,
                            intermediate_size=cfg.intermediate_size, num_hidden_layers=cfg.num_layers,
                            image_size=cfg.image_size, patch_size=cfg.patch_size, qkv_bias=cfg.qkv_bias,
                            qk_normalization=cfg.qk_normalization, norm_type=cfg.norm_type,
                            layer_norm_eps=cfg.layer_norm_eps, initializer_factor=cfg.initializer_factor,
                            use_flash_attn=False, drop_path_rate=0.0, hidden_act="gelu")
    vit = InternVisionModel(vc)
    W, C = cfg.shuffled_width, cfg.llm_hidden
    mlp1 = torch.nn.Sequential(torch.nn.LayerNorm(W), torch.nn.Linear(W, C), torch.nn.GELU(), torch.nn.Linear(C, C))
    vsd = {k[len("vision_model."):]: torch.from_numpy(v) for k, v in sd_np.items() if k.startswith("vision_model.")}
    missing = vit.load_state_dict(vsd, strict=True)
    msd = {k[len("mlp1."):]: torch.from_numpy(v) for k, v in sd_np.items() if k.startswith("mlp1.")}
    mlp1.load_state_dict(msd, strict=True)

    class StandIn:  # carries exactly what EVAL2.extract_feature reads from self (EVAL2:508-531)
        pass

    s = StandIn()
    s.select_layer = cfg.select_layer
    s.downsample_ratio = cfg.downsample_ratio
    s.ps_version = cfg.ps_version
    s.vision_model = vit.to(dtype).eval()
    s.mlp1 = mlp1.to(dtype).eval()
    s.pixel_shuffle = types.MethodType(EVAL2.InternVLChatModel.pixel_shuffle, s)
    s.extract_feature = types.MethodType(EVAL2.InternVLChatModel.extract_feature, s)
    return s


def golden_tower(name, cfg, n_frames, w_seed, x_seed, full_store, weights=synth.make_weights):
    sd = weights(cfg, seed=w_seed)
    pv = synth.make_pixel_values(n_frames, cfg, seed=x_seed)
    ref = build_reference(cfg, sd)
    with torch.no_grad():
        x = torch.from_numpy(pv)
        vit_out = ref.vision_model(pixel_values=x, output_hidden_states=True, return_dict=True)
        out = ref.extract_feature(x)
        # the reference's own bf16 path, for the record (calibrates the 2e-2 budget)
        ref16 = build_reference(cfg, sd, dtype=torch.bfloat16)
        out16 = ref16.extract_feature(x.to(torch.bfloat16)).float()
    hs = [h.numpy() for h in vit_out.hidden_states]
    out = out.numpy()
    e_max16 = float(np.abs(out16.numpy() - out).max() / np.abs(out).max())
    meta = dict(name=name, cfg=cfg.to_dict(), n_frames=n_frames, w_seed=w_seed, x_seed=x_seed,
                ref_bf16_e_max=e_max16, torch=torch.__version__)
    arrs = {}
    if full_store:
        arrs["embeddings_s"] = hs[0][:, ::5]
        arrs["layer0_s"] = hs[1][:, ::5]
        arrs["vit_out"] = hs[-1]
        arrs["out"] = out
    else:
        arrs["embeddings_s"] = hs[0][:, ::41, ::16]
        arrs["layer0_s"] = hs[1][:, ::41, ::16]
        arrs["vit_out_s"] = hs[-1][:, ::41, ::16]
        arrs["out_s"] = out[:, ::8, ::16]
        meta["out_absmax"] = float(np.abs(out).max())
        meta["out_absmean"] = float(np.abs(out).mean())
        meta["vit_absmax"] = float(np.abs(hs[-1]).max())
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **{k: v.astype(np.float32) for k, v in arrs.items()})
    print(name, {k: v.shape for k, v in arrs.items()}, "ref bf16 e_max", e_max16)
    return meta


def golden_preprocess():
    from PIL import Image
    from internvl.train.dataset import build_transform, dynamic_preprocess

    tf = build_transform(False, 448)
    cases = []
    for (h, w, kind, max_num, seed) in [(720, 1280, "smooth", 1, 11), (720, 1280, "uniform", 6, 12),
                                        (512, 768, "smooth", 6, 13), (448, 448, "uniform", 1, 14),
                                        (1280, 720, "smooth", 6, 15), (1080, 1920, "smooth", 1, 16)]:
        frame = synth.make_frames(1, h, w, seed=seed, kind=kind)[0]
        tiles = dynamic_preprocess(Image.fromarray(frame), max_num=max_num, image_size=448, use_thumbnail=True)
        if max_num == 1:
            tiles = tiles[:1]  # stage1_eval.py:453 keeps [0]
        t_u8 = np.stack([np.asarray(t) for t in tiles])
        pv = torch.stack([tf(t) for t in tiles]).numpy()
        cases.append(dict(h=h, w=w, kind=kind, max_num=max_num, seed=seed, n_tiles=len(tiles),
                          tiles_sha256=hashlib.sha256(t_u8.tobytes()).hexdigest(),
                          pv_sha256=hashlib.sha256(np.ascontiguousarray(pv).tobytes()).hexdigest(),
                          pv_bf16_sha256=hashlib.sha256(synth.to_bf16_bits(pv).tobytes()).hexdigest(),
                          tile0_crop=t_u8[0, 100:104, 200:204].tolist()))
        print("preprocess", h, w, kind, max_num, "->", len(tiles), "tiles")
    import PIL
    return dict(pillow=PIL.__version__, cases=cases)


def main():
    install_stubs()
    torch.manual_seed(0)
    metas = {}
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    if only:  # regenerate selected towers only, keeping the other entries of golden_meta.json
        with open(os.path.join(HERE, "golden_meta.json")) as f:
            metas = json.load(f)
        if "tiny_stress" in only:
            metas["tiny_stress"] = golden_tower("tiny_stress", cfgmod.tiny(), n_frames=2, w_seed=106, x_seed=107,
                                                full_store=False, weights=synth.make_stress_weights)
        if "vit6b_l1" in only:
            metas["vit6b_l1"] = golden_tower("vit6b_l1", cfgmod.internvit_6b(num_layers=1), n_frames=1, w_seed=104,
                                             x_seed=105, full_store=False)
        with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
            json.dump(metas, f, indent=1)
        return
    metas["tiny"] = golden_tower("tiny", cfgmod.tiny(), n_frames=2, w_seed=100, x_seed=101, full_store=True)
    tiny_rms = cfgmod.tiny(hidden_size=256, num_heads=2, intermediate_size=512, qkv_bias=False, qk_normalization=True,
                           norm_type="rms_norm", initializer_factor=0.1)
    metas["tiny_rms"] = golden_tower("tiny_rms", tiny_rms, n_frames=1, w_seed=102, x_seed=103, full_store=False)
    metas["vit300m"] = golden_tower("vit300m", cfgmod.internvit_300m(llm_hidden=2048), n_frames=1, w_seed=0, x_seed=1,
                                    full_store=False)
    # InternViT-6B block geometry at its REAL width (config 5): D=3200, 25 heads x 128, I=12800, RMSNorm, q/k RMSNorm over
    # all heads, no qkv bias, layer scale 0.1, projector 12800 -> 6144; one layer keeps the CPU run to seconds
    metas["vit6b_l1"] = golden_tower("vit6b_l1", cfgmod.internvit_6b(num_layers=1), n_frames=1, w_seed=104, x_seed=105,
                                     full_store=False)
    # "trained-like" statistics at the tiny width: peaked attention, +-300 outlier channels, layer scales 0.1 .. 1
    metas["tiny_stress"] = golden_tower("tiny_stress", cfgmod.tiny(), n_frames=2, w_seed=106, x_seed=107,
                                        full_store=False, weights=synth.make_stress_weights)
    metas["preprocess"] = golden_preprocess()
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(metas, f, indent=1)


if __name__ == "__main__":
    main()
