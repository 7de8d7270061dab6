"""Drop-in seam: rebind ``extract_feature`` on a live InternVLChatModel (any of the reference's six variants).

    import aigv_assessor_b200 as aigv
    aigv.install(model)          # model.extract_feature(pixel_values) now runs the sm_100a path

Everything above the seam — forward()'s <IMG_CONTEXT> splice, the SlowFast motion branch, the LoRA-merged LLM prefill
and the mlpscore head (eval2 modeling_internvl_chat.py:327-488) — runs unchanged on the returned tensor.

Inference only. The engine holds a SNAPSHOT of vision_model / mlp1 (LoRA merged) and returns a tensor without autograd
history, so: while the model is in training mode with trainable vision / mlp1 parameters the call is routed to the
reference's own method (gradients keep flowing, stage-1/2 training is unaffected), and after the weights change
(load_state_dict, a LoRA update, ``merge_and_unload``) call ``refresh(model)`` to re-ingest them.
"""
from __future__ import annotations

import types
import warnings

import torch

from .engine import VisionEngine


class _ProbeOut:
    def __init__(self, t):
        self.last_hidden_state = t
        self.hidden_states = (t,)


def _variant_flattens(model) -> bool:
    """Does this variant's extract_feature return [N,256,C] (stage-2 family, eval2 :527) or keep [N,16,16,C]
    (internvl_chat / eval1 / stage1_lora, internvl_chat/modeling_internvl_chat.py:406-408)?  Decided by RUNNING the
    class's own method on a 2x2-token stand-in (identity vision tower and projector): only its reshape logic executes."""
    cls = type(model)
    fn = getattr(cls, "extract_feature", None)
    if fn is None:
        return True
    probe = types.SimpleNamespace()
    probe.select_layer = -1
    probe.downsample_ratio = 0.5
    probe.ps_version = "v2"
    probe.vision_model = lambda pixel_values=None, **kw: _ProbeOut(torch.zeros(1, 5, 8))
    probe.mlp1 = lambda x: x
    ps = getattr(cls, "pixel_shuffle", None)
    if ps is None:
        return True
    probe.pixel_shuffle = types.MethodType(ps, probe)
    try:
        with torch.no_grad():
            y = fn(probe, torch.zeros(1, 3, 28, 28))
    except Exception:  # an unknown variant: the stage-2 shape is what every caller accepts (eval2 :328,377)
        return True
    return y.dim() == 3


def _trainable(model) -> bool:
    mods = [m for m in (getattr(model, "vision_model", None), getattr(model, "mlp1", None)) if m is not None]
    return any(p.requires_grad for m in mods if hasattr(m, "parameters") for p in m.parameters())


def install(model, max_tiles: int = 64, flatten: bool | None = None):
    """Bind the B200 path as ``model.extract_feature``. Configurations the kernels do not cover (select_layer != -1,
    downsample_ratio != 0.5, ps_version != 'v2') keep the original method."""
    original = model.extract_feature
    if model.select_layer != -1 or model.downsample_ratio != 0.5 or model.ps_version != "v2":
        return None
    engine = VisionEngine.from_model(model, max_tiles=max_tiles)
    flat = _variant_flattens(model) if flatten is None else flatten
    warned = []

    def extract_feature(self, pixel_values):
        if getattr(self, "training", False) and torch.is_grad_enabled() and _trainable(self):
            if not warned:
                warned.append(1)
                warnings.warn("aigv_assessor_b200: model is training with trainable vision/mlp1 parameters; "
                              "extract_feature falls back to the reference method so gradients flow")
            return original(pixel_values)
        with torch.no_grad():
            return self._aigv_engine.extract_feature(pixel_values, flatten=flat)

    model.extract_feature = types.MethodType(extract_feature, model)
    model._aigv_engine = engine
    model._aigv_original_extract_feature = original
    return engine


def refresh(model):
    """Re-ingest vision_model / mlp1 weights into the installed engine (after load_state_dict / LoRA updates)."""
    eng = getattr(model, "_aigv_engine", None)
    if eng is None:
        raise ValueError("aigv_assessor_b200.install(model) has not been called")
    eng.load_from_model(model)
    return eng


def uninstall(model):
    if hasattr(model, "_aigv_original_extract_feature"):
        model.extract_feature = model._aigv_original_extract_feature
        del model._aigv_original_extract_feature
        del model._aigv_engine
