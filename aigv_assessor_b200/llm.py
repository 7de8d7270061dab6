"""Host side of the InternLM2 prefill above the seam.

Mirrors what ``InternVLChatModel.forward`` asks of ``self.language_model``
(/root/reference/internvl/model/internvl_chat_eval2/modeling_internvl_chat.py:441-450, 469-488 ->
/root/reference/internvl/model/internlm2/modeling_internlm2.py:1036-1124): a full prefill of the spliced ``inputs_embeds``
of which only ``hidden_states[-1][:, -4]`` (mlpscore) and the logits of the answer span (argmax) are consumed. All compute
is the CUDA library behind ``aigv_llm_*``; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Sequence

import numpy as np
import torch

from . import _lib

_NAMES = ("attention.wqkv.weight", "attention.wo.weight", "attention_norm.weight", "ffn_norm.weight",
          "feed_forward.w1.weight", "feed_forward.w2.weight", "feed_forward.w3.weight")


def _sp(device) -> C.c_void_p:
    return C.c_void_p(int(torch.cuda.current_stream(device).cuda_stream))


class LanguageEngine:
    """One InternLM2 decoder stack on one GPU. ``llm`` carries the InternLM2Config fields (hidden_size,
    num_hidden_layers, num_attention_heads, num_key_value_heads, intermediate_size, vocab_size, rms_norm_eps, rope_theta,
    max_position_embeddings); ``max_tokens`` sizes the workspace (B*N per prefill)."""

    def __init__(self, llm: Mapping[str, object], max_tokens: int = 4096, device=None):
        if not torch.cuda.is_available():
            raise _lib.AigvError("LanguageEngine needs a CUDA device (no CPU path)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.llm = dict(llm)
        self.hidden = int(llm["hidden_size"])
        self.vocab = int(llm["vocab_size"])
        self.layers = int(llm["num_hidden_layers"])
        self.max_tokens = int(max_tokens)
        self.lib = _lib.load()
        cfg = _lib.LlmCfg(self.hidden, int(llm["num_attention_heads"]), int(llm["num_key_value_heads"]),
                          int(llm["intermediate_size"]), self.layers, self.vocab, float(llm["rms_norm_eps"]),
                          float(llm.get("rope_theta", 10000.0)), self.max_tokens,
                          int(llm.get("max_position_embeddings", 0)))
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            st = self.lib.aigv_llm_create(C.byref(cfg), C.byref(h))
        self._h = h
        self._check(st, "aigv_llm_create")
        self._finalized = False
        self._shape = None

    def _check(self, st: int, what: str):
        if st == 0:
            return
        msg = self.lib.aigv_status_string(st).decode()
        if self._h:
            d = self.lib.aigv_llm_last_error(self._h)
            if d:
                msg += ": " + d.decode()
        if st in (-1, -4, -5):
            raise ValueError(f"{what}: {msg}")
        raise _lib.AigvError(f"{what}: {msg}")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.aigv_llm_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def set_weight(self, name: str, value):
        t = torch.as_tensor(value)
        if t.dtype not in (torch.float32, torch.bfloat16):
            t = t.to(torch.float32)
        t = t.contiguous()
        shape = (C.c_int64 * t.dim())(*t.shape)
        with torch.cuda.device(self.device):
            st = self.lib.aigv_llm_set_weight(self._h, name.encode(), C.c_void_p(t.data_ptr()),
                                              1 if t.dtype == torch.bfloat16 else 0, shape, t.dim())
        self._check(st, f"aigv_llm_set_weight({name})")

    def load_state_dict(self, sd: Mapping[str, object], prefix: str = "language_model.") -> "LanguageEngine":
        """Reference state-dict names (modeling_internlm2.py parameter names under ``language_model.``). LoRA pairs of a
        peft-wrapped LLM ('...base_layer.weight' + 'lora_A/lora_B.default.weight'; the stage-2 checkpoints wrap wqkv / wo /
        w1 / w2 / w3, eval2 :285-305) are folded first: W + 2 B A as tools/merge_lora.py:18-25 does."""
        if any(".lora_A." in k for k in sd):
            from .engine import merge_lora_names
            sd = merge_lora_names({k: (torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v)
                                   for k, v in sd.items() if k.startswith(prefix) or "language_model" in k})
        for i in range(self.layers):
            for n in _NAMES:
                self.set_weight(f"model.layers.{i}.{n}", sd[f"{prefix}model.layers.{i}.{n}"])
        self.set_weight("model.norm.weight", sd[f"{prefix}model.norm.weight"])
        if f"{prefix}output.weight" in sd:
            self.set_weight("output.weight", sd[f"{prefix}output.weight"])
        with torch.cuda.device(self.device):
            st = self.lib.aigv_llm_finalize_weights(self._h, _sp(self.device))
        self._check(st, "aigv_llm_finalize_weights")
        self._finalized = True
        return self

    # ------------------------------------------------------------------ compute
    def prefill(self, inputs_embeds: torch.Tensor) -> "LanguageEngine":
        """inputs_embeds bf16 [B, N, hidden] on the engine's device (the output of llm_glue.splice_tokens)."""
        if inputs_embeds.dim() != 3 or inputs_embeds.shape[2] != self.hidden:
            raise ValueError(f"inputs_embeds must be [B, N, {self.hidden}], got {tuple(inputs_embeds.shape)}")
        if not inputs_embeds.is_cuda or inputs_embeds.device != self.device:
            raise ValueError(f"inputs_embeds must be on {self.device} (no CPU path)")
        if not self._finalized:
            raise ValueError("weights not loaded")
        x = inputs_embeds.detach().to(torch.bfloat16).contiguous()
        B, N, _ = x.shape
        with torch.cuda.device(self.device):
            st = self.lib.aigv_llm_prefill(self._h, C.c_void_p(x.data_ptr()), B, N, _sp(self.device))
        self._check(st, "aigv_llm_prefill")
        self._shape = (B, N)
        self._keep = x   # the kernel reads it asynchronously
        return self

    def _rows(self, rows: Sequence[int]):
        if self._shape is None:
            raise ValueError("no prefill has run")
        r = np.ascontiguousarray(np.asarray(rows, dtype=np.int32).reshape(-1))
        return r, r.ctypes.data_as(C.POINTER(C.c_int32))

    def hidden_rows(self, rows: Sequence[int]) -> torch.Tensor:
        """hidden_states[-1] (final norm applied) of flattened rows b*N + pos -> fp32 [n, hidden]."""
        r, rp = self._rows(rows)
        out = torch.empty((len(r), self.hidden), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            st = self.lib.aigv_llm_hidden_rows(self._h, rp, len(r), C.c_void_p(out.data_ptr()), _sp(self.device))
        self._check(st, "aigv_llm_hidden_rows")
        return out

    def logits_rows(self, rows: Sequence[int]) -> torch.Tensor:
        """logits of flattened rows b*N + pos -> fp32 [n, vocab]."""
        r, rp = self._rows(rows)
        out = torch.empty((len(r), self.vocab), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            st = self.lib.aigv_llm_logits_rows(self._h, rp, len(r), C.c_void_p(out.data_ptr()), _sp(self.device))
        self._check(st, "aigv_llm_logits_rows")
        return out

    def residual(self) -> torch.Tensor:
        B, N = self._shape
        out = torch.empty((B, N, self.hidden), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            st = self.lib.aigv_llm_copy_residual(self._h, C.c_void_p(out.data_ptr()), _sp(self.device))
        self._check(st, "aigv_llm_copy_residual")
        return out

    def stage2_tail(self, inputs_embeds: torch.Tensor, n_answer: int = 48):
        """What eval2 :469-488 consumes: (hidden_states[-1][:, -4] fp32 [B, hidden], argmax of the last n_answer shifted
        logits of the flattened batch, int64 [n_answer])."""
        self.prefill(inputs_embeds)
        B, N = self._shape
        hid = self.hidden_rows([b * N + N - 4 for b in range(B)])
        # shift_logits = logits[..., :-1, :] flattened over the batch: the last n_answer of them
        flat = [b * N + p for b in range(B) for p in range(N - 1)][-n_answer:]
        logits = self.logits_rows(flat)
        return hid, torch.argmax(logits, dim=1), logits

    @property
    def launch_count(self) -> int:
        return int(self.lib.aigv_llm_launch_count(self._h))


def rope_(qkv: torch.Tensor, B: int, N: int, heads: int, kv_heads: int, head_dim: int, theta: float) -> torch.Tensor:
    lib = _lib.load()
    with torch.cuda.device(qkv.device):
        st = lib.aigv_rope(C.c_void_p(qkv.data_ptr()), B, N, heads, kv_heads, head_dim, float(theta), _sp(qkv.device))
    _lib.check(st, None, "aigv_rope")
    return qkv


def attention_gqa(qkv: torch.Tensor, B: int, N: int, heads: int, kv_heads: int, head_dim: int, causal: bool = True,
                  scale: float | None = None) -> torch.Tensor:
    """qkv bf16 [B*N, (heads + 2*kv_heads)*head_dim] in InternLM2's wqkv packing -> bf16 [B*N, heads*head_dim]."""
    lib = _lib.load()
    out = torch.empty((B * N, heads * head_dim), dtype=torch.bfloat16, device=qkv.device)
    with torch.cuda.device(qkv.device):
        st = lib.aigv_attention_gqa(C.c_void_p(qkv.data_ptr()), C.c_void_p(out.data_ptr()), B, N, heads, kv_heads, head_dim,
                                    float(scale if scale is not None else head_dim ** -0.5), int(causal), _sp(qkv.device))
    _lib.check(st, None, "aigv_attention_gqa")
    return out
