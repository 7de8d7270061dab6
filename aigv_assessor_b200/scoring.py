"""The stage-2 eval forward of ``InternVLChatModel`` (/root/reference/internvl/model/internvl_chat_eval2/
modeling_internvl_chat.py:306-488) composed from this library's CUDA pieces:

    pixel_values --extract_feature--> visual tokens            aigv_extract_feature   (the seam; eval2 :328,508-531)
    SlowFast features --motion_mlp--> motion token             aigv_layernorm + 2 x aigv_gemm_epilogue   (eval2 :244-249,344-345)
    input_ids + tokens --splice--> inputs_embeds               aigv_splice_tokens     (eval2 :324,351-378)
    inputs_embeds --InternLM2 prefill--> hidden[-1][:, -4]      aigv_llm_prefill / aigv_llm_hidden_rows (eval2 :441-450,469)
    hidden --mlpscore--> score1                                aigv_mlpscore          (eval2 :475)
    answer-span logits --argmax--> 'logit'                     aigv_llm_logits_rows   (eval2 :488; stage1_eval.py:928-929)

    pixel_values --SlowFast-R50 trunk--> 2304 motion features    slowfast.SlowFastTrunk (eval2 :136-193, 337-343), when the state
                                                                 dict carries ``slowfast_model.*``; else pass ``motion_features``

What stays with the caller: tokenisation and the loss / label bookkeeping of training. Inference only; no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping

import numpy as np
import torch

from . import _lib, llm_glue
from .config import VisionPathConfig
from .engine import VisionEngine
from .llm import LanguageEngine


def _sp(device) -> C.c_void_p:
    return C.c_void_p(int(torch.cuda.current_stream(device).cuda_stream))


class MotionMlp:
    """motion_mlp = LayerNorm(2304) -> Linear(2304, 4096) -> GELU -> Linear(4096, 4096)  (eval2 :244-249)."""

    def __init__(self, sd: Mapping[str, object], device):
        self.device = torch.device(device)
        f32 = lambda k: torch.as_tensor(sd[k]).to(self.device, torch.float32).contiguous()
        b16 = lambda k: torch.as_tensor(sd[k]).to(self.device).to(torch.bfloat16).contiguous()
        self.ln_w, self.ln_b = f32("motion_mlp.0.weight"), f32("motion_mlp.0.bias")
        self.w1, self.b1 = b16("motion_mlp.1.weight"), f32("motion_mlp.1.bias")
        self.w2, self.b2 = b16("motion_mlp.3.weight"), f32("motion_mlp.3.bias")
        self.d_in, self.d_mid, self.d_out = self.w1.shape[1], self.w1.shape[0], self.w2.shape[0]
        if self.w2.shape[1] != self.d_mid or self.ln_w.numel() != self.d_in:
            raise ValueError("motion_mlp weights have inconsistent shapes")
        self.lib = _lib.load()

    def __call__(self, feat: torch.Tensor) -> torch.Tensor:
        """feat [B, 2304] (any float dtype, CUDA) -> motion token bf16 [B, 4096]"""
        if feat.dim() != 2 or feat.shape[1] != self.d_in or feat.device != self.device:
            raise ValueError(f"motion features must be [B, {self.d_in}] on {self.device}")
        x = feat.to(torch.float32).contiguous()
        n = x.shape[0]
        xn = torch.empty((n, self.d_in), dtype=torch.bfloat16, device=self.device)
        h = torch.empty((n, self.d_mid), dtype=torch.bfloat16, device=self.device)
        out = torch.empty((n, self.d_out), dtype=torch.bfloat16, device=self.device)
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(self.device):
            st = _sp(self.device)
            _lib.check(self.lib.aigv_layernorm(p(x), p(self.ln_w), p(self.ln_b), p(xn), n, self.d_in, 1e-5, 0, st), None, "motion_mlp.0")
            _lib.check(self.lib.aigv_gemm_epilogue(p(xn), p(self.w1), p(self.b1), None, None, None, p(h), n, self.d_mid, self.d_in, 1, st),
                       None, "motion_mlp.1")
            _lib.check(self.lib.aigv_gemm_epilogue(p(h), p(self.w2), p(self.b2), None, None, None, p(out), n, self.d_out, self.d_mid, 0, st),
                       None, "motion_mlp.3")
        return out


class ClipScorer:
    """Everything ``forward`` computes at eval time for a clip, on one GPU.

    ``sd`` uses the reference's state-dict names (vision_model.*, mlp1.*, motion_mlp.*, language_model.*, mlpscore.*)."""

    def __init__(self, sd: Mapping[str, object], vis_cfg: VisionPathConfig, llm_cfg: Mapping[str, object],
                 img_context_token_id: int, max_tiles: int = 64, max_tokens: int = 4096, device=None):
        if not torch.cuda.is_available():
            raise _lib.AigvError("ClipScorer needs a CUDA device (no CPU path)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.img_context_token_id = int(img_context_token_id)
        with torch.cuda.device(self.device):
            self.vision = VisionEngine(vis_cfg, max_tiles=max_tiles, device=self.device).load_state_dict(sd)
            self.language = LanguageEngine(llm_cfg, max_tokens=max_tokens, device=self.device).load_state_dict(sd)
            self.motion = MotionMlp(sd, self.device)
            self.head = llm_glue.MlpScoreHead(sd, self.device)
            self.tok = torch.as_tensor(sd["language_model.model.tok_embeddings.weight"]).to(self.device).to(torch.bfloat16).contiguous()
            self.slowfast = None
            if any(k.startswith("slowfast_model.") for k in sd):
                from .slowfast import SlowFastTrunk
                self.slowfast = SlowFastTrunk(sd, device=self.device)
            self._side = None
            self.check_splice = True   # read the splice kernel's error flag back (one host sync per call)

    @classmethod
    def from_parts(cls, vision: VisionEngine, language: LanguageEngine, motion: MotionMlp, head, tok_embeddings: torch.Tensor,
                   img_context_token_id: int, slowfast=None) -> "ClipScorer":
        """Assemble from engines that already hold their weights (e.g. generated on the device)."""
        self = cls.__new__(cls)
        self.device = vision.device
        self.img_context_token_id = int(img_context_token_id)
        self.vision, self.language, self.motion, self.head, self.slowfast = vision, language, motion, head, slowfast
        self.tok = tok_embeddings
        self._side = None
        self.check_splice = True
        return self

    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor, input_ids: torch.Tensor, motion_features: torch.Tensor | None = None,
                n_answer: int = 48) -> dict:
        """pixel_values bf16 [B*T, 3, 448, 448]; input_ids int64 [B, N]; motion_features [B, 2304] (the SlowFast output,
        eval2 :343) or None to run this library's trunk on pixel_values. Returns what the eval scripts read from the reference's forward: 'score1' fp32 [B] and 'logit', the
        argmax of the last n_answer shifted logits (stage1_eval.py:928-929 slices exactly that tail), plus 'hidden' =
        hidden_states[-1][:, -4]."""
        pixel_values = pixel_values.to(self.device)
        if motion_features is None:
            if self.slowfast is None:
                raise ValueError("no slowfast_model.* weights were given: pass motion_features")
            # the motion branch (215 small, latency-bound launches as one CUDA graph + motion_mlp) runs on a side stream next to
            # the ViT: the two only meet at the splice
            if getattr(self, "_side", None) is None:
                self._side = torch.cuda.Stream(device=self.device)
            main = torch.cuda.current_stream(self.device)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                mot = self.motion(self.slowfast(pixel_values, input_ids.shape[0]))
            vit = self.vision.extract_feature(pixel_values)
            main.wait_stream(self._side)
            mot.record_stream(main)
        else:
            vit = self.vision.extract_feature(pixel_values)
            mot = self.motion(motion_features.to(self.device))
        x = llm_glue.splice_tokens(input_ids.to(self.device), self.tok, vit, mot, self.img_context_token_id,
                                   check=getattr(self, "check_splice", True))
        hid, arg, logits = self.language.stage2_tail(x, n_answer=n_answer)
        score, pre = self.head(hid)
        return {"score1": score, "score_pre_relu": pre, "logit": arg, "logits_tail": logits, "hidden": hid}

    __call__ = forward

    def score_frames(self, frames_u8: torch.Tensor, input_ids: torch.Tensor, motion_features: torch.Tensor | None = None) -> torch.Tensor:
        """uint8 frames [T, H, W, 3] of ONE clip (decord layout) -> score1 [1]: front end + forward."""
        pv, _ = self.vision.preprocess_frames(frames_u8, max_num=1)
        return self.forward(pv, input_ids, motion_features)["score1"]

    def score_video(self, path: str, input_ids: torch.Tensor, num_segments: int = 8, bound=None) -> torch.Tensor:
        """Video file -> score1 [1]: sampling.load_video (CPU decode of the sampled frames, as the reference does) + score_frames;
        needs the motion trunk's weights (or use score_frames with your own motion features)."""
        from . import sampling
        frames, _ = sampling.load_video(path, num_segments=num_segments, bound=bound)
        return self.score_frames(frames, input_ids)

    @classmethod
    def from_model(cls, model, **kw) -> "ClipScorer":
        """Build from a live eval2-style InternVLChatModel (reads its state dict and configs; LoRA in the vision tower is
        merged by VisionEngine.load_state_dict, the LLM is expected merged as in the released stage-2 checkpoints)."""
        vc = model.vision_model.config
        m1 = model.mlp1.state_dict()
        vis = VisionPathConfig(hidden_size=vc.hidden_size, num_heads=vc.num_attention_heads,
                               intermediate_size=vc.intermediate_size, num_layers=vc.num_hidden_layers,
                               image_size=vc.image_size, patch_size=vc.patch_size, llm_hidden=m1["3.weight"].shape[0],
                               qkv_bias=bool(vc.qkv_bias), qk_normalization=bool(vc.qk_normalization),
                               norm_type=vc.norm_type, layer_norm_eps=vc.layer_norm_eps)
        lc = model.language_model.config
        llm = dict(hidden_size=lc.hidden_size, num_hidden_layers=lc.num_hidden_layers,
                   num_attention_heads=lc.num_attention_heads, num_key_value_heads=lc.num_key_value_heads,
                   intermediate_size=lc.intermediate_size, vocab_size=lc.vocab_size, rms_norm_eps=lc.rms_norm_eps,
                   rope_theta=getattr(lc, "rope_theta", 10000.0), max_position_embeddings=lc.max_position_embeddings)
        return cls(model.state_dict(), vis, llm, model.img_context_token_id, **kw)
