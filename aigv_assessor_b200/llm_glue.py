"""Host side of the first pieces ABOVE the seam: the visual-token splice and the mlpscore head
of ``InternVLChatModel.forward`` (/root/reference/internvl/model/internvl_chat_eval2/modeling_internvl_chat.py:324,351-378
and :82-94,469-475) as CUDA kernels behind ``aigv_splice_tokens`` / ``aigv_mlpscore``. The LLM prefill between them still
runs on the caller's side (the reference's InternLM2, unchanged).
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping

import torch

from . import _lib


def _sp(device) -> C.c_void_p:
    return C.c_void_p(int(torch.cuda.current_stream(device).cuda_stream))


def splice_tokens(input_ids: torch.Tensor, tok_embeddings: torch.Tensor, vit_embeds: torch.Tensor,
                  motion_embeds: torch.Tensor, img_context_token_id: int, check: bool = True) -> torch.Tensor:
    """input_ids int64 [B,N] (CUDA); tok_embeddings bf16 [V,C]; vit_embeds bf16 [*, C] (frames x 256 rows, prompt order);
    motion_embeds bf16 [B,C] -> inputs_embeds bf16 [B,N,C]. With check=True the kernel's error flag is read back (one
    host sync) and a mismatch between visual rows and <IMG_CONTEXT> slots raises ValueError (the reference prints a
    warning and falls into its except branch there, eval2 :381-386)."""
    if input_ids.dim() != 2 or input_ids.dtype != torch.int64 or not input_ids.is_cuda:
        raise ValueError("input_ids must be an int64 CUDA tensor [B, N]")
    dev = input_ids.device
    B, N = input_ids.shape
    V, Cc = tok_embeddings.shape
    for t in (tok_embeddings, vit_embeds, motion_embeds):
        if t.dtype != torch.bfloat16 or t.device != dev:
            raise ValueError("embeddings must be bf16 tensors on the device of input_ids")
    vit = vit_embeds.reshape(-1, Cc).contiguous()
    mot = motion_embeds.reshape(-1, Cc).contiguous()
    if mot.shape[0] != B or Cc % 8:
        raise ValueError("motion_embeds must hold one row per sample; the width must be a multiple of 8")
    out = torch.empty((B, N, Cc), dtype=torch.bfloat16, device=dev)
    work = torch.empty((B * N + 2 * B + 1,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        st = lib.aigv_splice_tokens(C.c_void_p(input_ids.contiguous().data_ptr()), C.c_void_p(tok_embeddings.contiguous().data_ptr()),
                                    C.c_void_p(vit.data_ptr()), C.c_void_p(mot.data_ptr()), C.c_void_p(out.data_ptr()),
                                    C.c_void_p(work.data_ptr()), B, N, Cc, V, vit.shape[0], int(img_context_token_id), _sp(dev))
    _lib.check(st, None, "aigv_splice_tokens")
    if check:
        flag = int(work[-1].item())
        if flag & 1:
            raise ValueError("input_ids holds a token id outside the embedding table")
        if flag & 6:
            raise ValueError(f"{vit.shape[0]} visual tokens do not match the <IMG_CONTEXT> slots of input_ids")
    return out


class MlpScoreHead:
    """The mlpscore regression head (eval2 :43-94): fc1..fc5 with ReLU after every layer, on hidden_states[-1][:, -4]."""

    def __init__(self, sd: Mapping[str, object], device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("MlpScoreHead needs a CUDA device (no CPU path)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.w, self.b = [], []
        for l in range(1, 6):
            self.w.append(torch.as_tensor(sd[f"mlpscore.fc{l}.weight"]).to(self.device, torch.float32).contiguous())
            self.b.append(torch.as_tensor(sd[f"mlpscore.fc{l}.bias"]).to(self.device, torch.float32).contiguous())
        dims = [4096, 1024, 256, 64, 16, 1]
        for l in range(5):
            if tuple(self.w[l].shape) != (dims[l + 1], dims[l]):
                raise ValueError(f"mlpscore.fc{l + 1}.weight has shape {tuple(self.w[l].shape)}")
        self._wp = (C.c_void_p * 5)(*[t.data_ptr() for t in self.w])
        self._bp = (C.c_void_p * 5)(*[t.data_ptr() for t in self.b])

    def __call__(self, hidden: torch.Tensor):
        """hidden fp32 [B, 4096] (CUDA) -> (score1 [B], pre-ReLU value of the last layer [B])"""
        if hidden.dim() != 2 or hidden.shape[1] != 4096 or hidden.device != self.device:
            raise ValueError("hidden must be [B, 4096] on the head's device")
        x = hidden.to(torch.float32).contiguous()
        n = x.shape[0]
        score = torch.empty((n,), dtype=torch.float32, device=self.device)
        pre = torch.empty((n,), dtype=torch.float32, device=self.device)
        lib = _lib.load()
        with torch.cuda.device(self.device):
            st = lib.aigv_mlpscore(C.c_void_p(x.data_ptr()), 4096, self._wp, self._bp, C.c_void_p(score.data_ptr()),
                                   C.c_void_p(pre.data_ptr()), n, _sp(self.device))
        _lib.check(st, None, "aigv_mlpscore")
        return score, pre
