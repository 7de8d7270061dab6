"""Host-side mirror of the reference's vision path on top of the C-ABI (include/aigv_b200.h).

``VisionEngine.extract_feature`` keeps the contract of ``InternVLChatModel.extract_feature``
(/root/reference/internvl/model/internvl_chat_eval2/modeling_internvl_chat.py:508-531): pixel_values
``[N,3,448,448]`` -> visual-token embeddings ``[N,256,C]`` (or ``[N,16,16,C]`` for the stage-1 family,
internvl_chat/modeling_internvl_chat.py:390-409), same dtype/device, on the caller's current CUDA stream.
``preprocess_frames`` mirrors ``load_video``'s per-frame ``dynamic_preprocess`` + transform + ``.to(bfloat16)``
(/root/reference/internvl/train/internvl/eval/stage1_eval.py:443-457,484-485,922).

torch is used for device memory and streams only; all compute is in libaigv_b200.so. There is no CPU fallback:
without a CUDA device or without the built library the constructors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Sequence

import numpy as np
import torch

from . import _lib, tiling
from .config import VisionPathConfig

_DT = {torch.float32: 0, torch.bfloat16: 1}


def _stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def merge_lora_names(sd: Mapping[str, "torch.Tensor"], scale: float = 2.0) -> dict:
    """Normalise a (possibly peft-wrapped) state dict to plain reference names, folding LoRA into the base weight:
    W + (alpha/r) * B @ A with alpha = 2r (eval2 modeling_internvl_chat.py:265-266,275-283; tools/merge_lora.py:18-21)."""
    out, lora_a, lora_b = {}, {}, {}
    for k, v in sd.items():
        k = k.replace("base_model.model.", "")
        if ".lora_A." in k:
            lora_a[k.split(".lora_A.")[0]] = v
        elif ".lora_B." in k:
            lora_b[k.split(".lora_B.")[0]] = v
        else:
            out[k.replace(".base_layer.", ".")] = v
    for mod, a in lora_a.items():
        b = lora_b[mod]
        key = mod + ".weight"
        out[key] = (out[key].float() + scale * (b.float() @ a.float())).to(out[key].dtype)
    return out


class VisionEngine:
    def __init__(self, cfg: VisionPathConfig, max_tiles: int = 64, device=None):
        if not torch.cuda.is_available():
            raise _lib.AigvError("aigv_assessor_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if cfg.select_layer != -1 or cfg.downsample_ratio != 0.5 or cfg.ps_version != "v2":
            raise ValueError("only select_layer=-1, downsample_ratio=0.5, ps_version='v2' are covered (SURVEY 8b)")
        self.cfg = cfg
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise ValueError(f"VisionEngine needs a CUDA device, got {self.device}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.load()
        self.max_tiles = int(max_tiles)
        c = _lib.Cfg(cfg.hidden_size, cfg.num_heads, cfg.intermediate_size, cfg.num_layers, cfg.image_size,
                     cfg.patch_size, cfg.llm_hidden, int(cfg.qkv_bias), int(cfg.qk_normalization),
                     0 if cfg.norm_type == "layer_norm" else 1, cfg.layer_norm_eps, cfg.mlp1_ln_eps, self.max_tiles)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            st = self.lib.aigv_create(C.byref(c), C.byref(self._h))
        if st != 0:
            h, self._h = self._h, C.c_void_p()
            try:
                _lib.check(st, h, "aigv_create")
            finally:
                if h:
                    self.lib.aigv_destroy(h)
        self._finalized = False

    def close(self):
        if getattr(self, "_h", None):
            self.lib.aigv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def set_weight(self, name: str, value):
        if isinstance(value, np.ndarray):
            value = torch.from_numpy(np.ascontiguousarray(value))
        t = value.detach()
        if t.dtype not in _DT:
            t = t.float()
        t = t.contiguous()
        shape = (C.c_int64 * t.dim())(*t.shape)
        with torch.cuda.device(self.device):
            st = self.lib.aigv_set_weight(self._h, name.encode(), C.c_void_p(t.data_ptr()), _DT[t.dtype], shape,
                                          t.dim())
        _lib.check(st, self._h, f"aigv_set_weight({name})")

    def load_state_dict(self, sd: Mapping[str, object]):
        """sd uses the reference's names: 'vision_model.*' and 'mlp1.*' (other keys are ignored). LoRA pairs of a
        peft-wrapped tower ('...base_layer.weight' + 'lora_A/lora_B.default.weight') are merged first."""
        if any(".lora_A." in k for k in sd):
            sd = merge_lora_names({k: (torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v)
                                   for k, v in sd.items()})
        for k, v in sd.items():
            if k.startswith("vision_model.") or k.startswith("mlp1."):
                self.set_weight(k, v)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.aigv_finalize_weights(self._h, None), self._h, "aigv_finalize_weights")
        self._finalized = True
        return self

    def load_checkpoint(self, path: str) -> "VisionEngine":
        """Ingest vision_model.* / mlp1.* from a safetensors file or checkpoint directory (checkpoint.py)."""
        from . import checkpoint
        return self.load_state_dict(checkpoint.load_checkpoint_dir(path))

    @classmethod
    def from_model(cls, model, max_tiles: int = 64, llm_hidden: int | None = None) -> "VisionEngine":
        """Build from a live InternVLChatModel-like object (reads model.vision_model and model.mlp1; a peft-wrapped
        vision tower is merged)."""
        vm = model.vision_model
        vsd = merge_lora_names(vm.state_dict())
        vc = getattr(vm, "config", None) or vm.base_model.model.config
        m1 = model.mlp1.state_dict()
        cfg = VisionPathConfig(
            hidden_size=vc.hidden_size, num_heads=vc.num_attention_heads, intermediate_size=vc.intermediate_size,
            num_layers=vc.num_hidden_layers, image_size=vc.image_size, patch_size=vc.patch_size,
            llm_hidden=llm_hidden or m1["3.weight"].shape[0], qkv_bias=bool(vc.qkv_bias),
            qk_normalization=bool(vc.qk_normalization), norm_type=vc.norm_type, layer_norm_eps=vc.layer_norm_eps,
            mlp1_ln_eps=float(model.mlp1[0].eps), downsample_ratio=model.downsample_ratio,
            ps_version=model.ps_version, select_layer=model.select_layer)
        dev = next(model.mlp1.parameters()).device
        eng = cls(cfg, max_tiles=max_tiles, device=dev if dev.type == "cuda" else None)
        return eng.load_from_model(model)

    def load_from_model(self, model) -> "VisionEngine":
        """(Re-)ingest vision_model.* / mlp1.* from a live model, merging a peft-wrapped tower's LoRA."""
        vsd = merge_lora_names(model.vision_model.state_dict())
        sd = {"vision_model." + k: v for k, v in vsd.items()}
        sd.update({"mlp1." + k: v for k, v in model.mlp1.state_dict().items()})
        return self.load_state_dict(sd)

    # ------------------------------------------------------------------ the hot path
    def extract_feature(self, pixel_values: torch.Tensor, flatten: bool = True,
                        out: torch.Tensor | None = None) -> torch.Tensor:
        """out (optional): a preallocated bf16 CUDA tensor [N, 256, C] to write into (avoids an allocation per call
        when the caller pipelines copies on side streams).
        The path computes in bf16 with fp32 accumulation whatever the input dtype (the reference's GPU dtype,
        stage2_eval.py:932): fp32 / fp16 pixel_values are rounded to bf16 on entry and the result is cast back."""
        if pixel_values.dim() != 4:
            raise ValueError(f"wrong pixel_values size: {pixel_values.shape}")  # modeling_intern_vit.py:345
        cfg = self.cfg
        n = pixel_values.shape[0]
        if tuple(pixel_values.shape[1:]) != (3, cfg.image_size, cfg.image_size):
            raise ValueError(f"wrong pixel_values size: {pixel_values.shape}")
        if not pixel_values.is_cuda:
            raise ValueError("pixel_values must be a CUDA tensor (no CPU path)")
        if pixel_values.device != self.device:
            raise ValueError(f"pixel_values is on {pixel_values.device}, the engine (weights, workspace) on {self.device}")
        if not self._finalized:
            raise ValueError("weights not loaded")
        in_dtype = pixel_values.dtype
        pv = pixel_values.detach().to(torch.bfloat16).contiguous()
        g2 = cfg.grid // 2
        if out is None:
            out = torch.empty((n, g2 * g2, cfg.llm_hidden), dtype=torch.bfloat16, device=pv.device)
        elif (out.dtype != torch.bfloat16 or out.device != self.device or not out.is_contiguous()
              or out.numel() != n * g2 * g2 * cfg.llm_hidden):
            raise ValueError("out must be a contiguous bf16 tensor with N*256*C elements on the engine's device")
        else:
            out = out.view(n, g2 * g2, cfg.llm_hidden)
        if n:
            with torch.cuda.device(pv.device):
                st = self.lib.aigv_extract_feature(self._h, C.c_void_p(pv.data_ptr()), n, C.c_void_p(out.data_ptr()),
                                                   C.c_void_p(_stream_ptr(pv.device)))
            _lib.check(st, self._h, "aigv_extract_feature")
        if not flatten:
            out = out.view(n, g2, g2, cfg.llm_hidden)
        return out if in_dtype == torch.bfloat16 else out.to(in_dtype)

    def vit_hidden(self, n_tiles: int) -> torch.Tensor:
        """fp32 ViT last_hidden_state [n,1025,D] of the most recent micro-batch (tests)."""
        x = torch.empty((n_tiles, self.cfg.tokens, self.cfg.hidden_size), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            st = self.lib.aigv_copy_vit_hidden(self._h, C.c_void_p(x.data_ptr()), n_tiles,
                                               C.c_void_p(_stream_ptr(self.device)))
        _lib.check(st, self._h, "aigv_copy_vit_hidden")
        return x

    # ------------------------------------------------------------------ instrumentation
    def profile(self, on: bool):
        _lib.check(self.lib.aigv_profile_enable(self._h, int(on)), self._h, "aigv_profile_enable")

    def profile_read(self, reset: bool = True) -> dict:
        """{kernel class: (total ms, launches)} measured with CUDA events on the launching stream."""
        k = self.lib.aigv_profile_num_classes()
        ms, cnt = (C.c_double * k)(), (C.c_int64 * k)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.aigv_profile_read(self._h, ms, cnt, k, int(reset)), self._h, "aigv_profile_read")
        return {self.lib.aigv_profile_class_name(i).decode(): (ms[i], cnt[i]) for i in range(k) if cnt[i]}

    @property
    def launch_count(self) -> int:
        return int(self.lib.aigv_launch_count(self._h))

    # ------------------------------------------------------------------ front end
    def preprocess_frames(self, frames_u8: torch.Tensor, max_num: int = 1, use_thumbnail: bool = True,
                          return_u8: bool = False):
        """frames_u8: uint8 [T,H,W,3] (HWC RGB, CUDA or pinned/pageable host). Returns (pixel_values bf16
        [tiles,3,448,448], tiles_per_frame[, tiles_u8 [tiles,448,448,3]]). max_num=1 is the video path of the
        reference (one tile per frame, stage1_eval.py:453); max_num>1 follows dynamic_preprocess incl. thumbnail."""
        if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[-1] != 3:
            raise ValueError(f"frames must be uint8 [T,H,W,3], got {frames_u8.dtype} {tuple(frames_u8.shape)}")
        t, hh, ww, _ = frames_u8.shape
        if not frames_u8.is_cuda:
            frames_u8 = frames_u8.to(self.device, non_blocking=True)
        elif frames_u8.device != self.device:
            raise ValueError(f"frames are on {frames_u8.device}, the engine on {self.device}")
        frames_u8 = frames_u8.contiguous()
        gi, gj = tiling.tile_grid(ww, hh, 1, max_num, self.cfg.image_size)
        per = gi * gj + (1 if (use_thumbnail and gi * gj != 1) else 0)
        s = self.cfg.image_size
        out = torch.empty((t * per, 3, s, s), dtype=torch.bfloat16, device=frames_u8.device)
        u8 = torch.empty((t * per, s, s, 3), dtype=torch.uint8, device=frames_u8.device) if return_u8 else None
        if t:
            with torch.cuda.device(frames_u8.device):
                st = self.lib.aigv_preprocess_u8(self._h, C.c_void_p(frames_u8.data_ptr()), t, hh, ww, hh * ww * 3,
                                                 ww * 3, gi, gj, int(use_thumbnail),
                                                 C.c_void_p(u8.data_ptr()) if return_u8 else None,
                                                 C.c_void_p(out.data_ptr()), C.c_void_p(_stream_ptr(out.device)))
            _lib.check(st, self._h, "aigv_preprocess_u8")
        counts = [per] * t
        return (out, counts, u8) if return_u8 else (out, counts)

    def encode_frames(self, frames_u8: torch.Tensor, max_num: int = 1, use_thumbnail: bool = True) -> torch.Tensor:
        """frames -> visual tokens in one call (front end + extract_feature)."""
        pv, _ = self.preprocess_frames(frames_u8, max_num=max_num, use_thumbnail=use_thumbnail)
        return self.extract_feature(pv)
