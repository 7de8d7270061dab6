"""Drop-in seam: rebind ``extract_feature`` on a live InternVLChatModel (any of the reference's six variants).

    import aigv_assessor_b200 as aigv
    aigv.install(model)          # model.extract_feature(pixel_values) now runs the sm_100a path

Everything above the seam — forward()'s <IMG_CONTEXT> splice, the SlowFast motion branch, the LoRA-merged LLM prefill
and the mlpscore head (eval2 modeling_internvl_chat.py:327-488) — runs unchanged on the returned tensor.
"""
from __future__ import annotations

import inspect
import types

import torch

from .engine import VisionEngine


def _variant_flattens(model) -> bool:
    """stage-2 family returns [N,256,C] (eval2 :527); CHAT/EVAL1/stage1_lora keep [N,16,16,C] (internvl_chat :406-408)."""
    try:
        src = inspect.getsource(type(model).extract_feature)
    except (OSError, TypeError):
        return True
    return "reshape(vit_embeds.shape[0], -1, vit_embeds.shape[-1])" in src.replace("\n", "")


def install(model, max_tiles: int = 64, flatten: bool | None = None):
    """Bind the B200 path as ``model.extract_feature``. Configurations the kernels do not cover (select_layer != -1,
    downsample_ratio != 0.5, ps_version != 'v2') keep the original method."""
    original = model.extract_feature
    if model.select_layer != -1 or model.downsample_ratio != 0.5 or model.ps_version != "v2":
        return None
    engine = VisionEngine.from_model(model, max_tiles=max_tiles)
    flat = _variant_flattens(model) if flatten is None else flatten

    def extract_feature(self, pixel_values):
        with torch.no_grad():
            return engine.extract_feature(pixel_values, flatten=flat)

    model.extract_feature = types.MethodType(extract_feature, model)
    model._aigv_engine = engine
    model._aigv_original_extract_feature = original
    return engine


def uninstall(model):
    if hasattr(model, "_aigv_original_extract_feature"):
        model.extract_feature = model._aigv_original_extract_feature
        del model._aigv_original_extract_feature
        del model._aigv_engine
