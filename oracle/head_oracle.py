"""CPU restatement (torch fp32) of what the reference runs ON TOP of extract_feature in a stage-2 forward.

TEST INFRASTRUCTURE ONLY -- imported by tests/ (and tests/golden/make_golden_full.py); never by the product.

It exists to check the north-star score tolerance through the seam: visual-token embeddings from the CUDA path are
spliced into the prompt exactly like the reference does and pushed through the (unchanged) caller's arithmetic --
token splice, motion token, InternLM2 prefill, mlpscore -- and the resulting score1 / logits are compared with the
reference's own. Pinned by tests/golden/full_*.npz, which tests/golden/make_golden_full.py generates by running the
UNMODIFIED /root/reference internvl_chat_eval2.InternVLChatModel on CPU (tests/test_oracle_golden.py).

Follows (all under /root/reference/internvl/model/internvl_chat_eval2/):
  modeling_internvl_chat.py:324-378  input_embeds, <IMG_CONTEXT> splice (all but the last slot <- vit_embeds, last <- motion)
  modeling_internvl_chat.py:337-345  frames -> slowfast -> motion_mlp (the slowfast trunk is a deterministic stand-in here
                                     and in the golden generator: it needs a network download)
  modeling_internvl_chat.py:441-488  language_model(inputs_embeds, output_hidden_states) -> hidden_states[-1][:, -4] ->
                                     mlpscore -> score1; argmax of the shifted logits
  modeling_internvl_chat.py:82-94    MLP.forward (ReLU after every layer, including the last)
  modeling_internlm2.py:129-143 RMSNorm, :161-260 rotary (dynamic NTK: unchanged base below max_position_embeddings),
  :264-278 SwiGLU, :360-440 attention (fused wqkv, GQA, causal mask, fp32 softmax), :624-695 layer, :868-998 model
  (final norm included in hidden_states[-1]), :1075-1096 output head.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def fake_slowfast(pixel_values: torch.Tensor, n_clips: int) -> torch.Tensor:
    """Deterministic stand-in for the SlowFast trunk (eval2 :136-174): [B*T,3,H,W] -> [B,2304]. The same function is
    bound as `slowfast` in the golden generator, so reference and oracle see identical motion features."""
    m = pixel_values.float().reshape(n_clips, -1).mean(dim=1)
    return torch.sin(torch.arange(2304, dtype=torch.float32)[None] * 0.01 + m[:, None]) * 0.5


def motion_embeds(feat: torch.Tensor, sd) -> torch.Tensor:
    """motion_mlp = LayerNorm(2304) -> Linear(2304,4096) -> GELU -> Linear(4096,4096)  (eval2 :244-249,344-345)"""
    x = F.layer_norm(feat, (2304,), sd["motion_mlp.0.weight"], sd["motion_mlp.0.bias"], 1e-5)
    x = F.gelu(F.linear(x, sd["motion_mlp.1.weight"], sd["motion_mlp.1.bias"]))
    return F.linear(x, sd["motion_mlp.3.weight"], sd["motion_mlp.3.bias"])


def splice(input_ids: torch.Tensor, tok_emb: torch.Tensor, vit_embeds: torch.Tensor, motion: torch.Tensor,
           img_context_token_id: int) -> torch.Tensor:
    """eval2 :351-378: every <IMG_CONTEXT> slot but the last of each sample takes a visual token, the last one the motion
    token."""
    B, N = input_ids.shape
    C = tok_emb.shape[1]
    emb = tok_emb[input_ids].clone().reshape(B * N, C)
    selected = input_ids == img_context_token_id
    cs = torch.cumsum(selected, dim=1)
    last = (cs == cs.max(dim=1, keepdim=True)[0]) & selected
    sel1 = (selected & ~last).reshape(B * N)
    sel2 = last.reshape(B * N)
    emb[sel1] = emb[sel1] * 0.0 + vit_embeds.reshape(-1, C)
    emb[sel2] = emb[sel2] * 0.0 + motion.reshape(-1, C)
    return emb.reshape(B, N, C)


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    v = x.float().pow(2).mean(-1, keepdim=True)
    return w * (x.float() * torch.rsqrt(v + eps)).to(x.dtype)


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def internlm2(inputs_embeds: torch.Tensor, sd, llm: dict):
    """-> (hidden_states[-1] = final-norm output [B,N,H], logits [B,N,V])"""
    H, nh, nkv = llm["hidden_size"], llm["num_attention_heads"], llm["num_key_value_heads"]
    hd, groups, eps = H // nh, nh // nkv, llm["rms_norm_eps"]
    B, N, _ = inputs_embeds.shape
    inv_freq = 1.0 / (llm["rope_theta"] ** (torch.arange(0, hd, 2).float() / hd))
    freqs = torch.einsum("i,j->ij", torch.arange(N).float(), inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos()[None, None], emb.sin()[None, None]
    mask = torch.full((N, N), torch.finfo(torch.float32).min).triu(1)[None, None]
    x = inputs_embeds
    lm = "language_model.model."
    for i in range(llm["num_hidden_layers"]):
        lp = f"{lm}layers.{i}."
        h = rms_norm(x, sd[lp + "attention_norm.weight"], eps)
        qkv = F.linear(h, sd[lp + "attention.wqkv.weight"]).view(B, N, nkv, groups + 2, hd)
        q = qkv[..., :groups, :].reshape(B, N, nh, hd).transpose(1, 2)
        k = qkv[..., -2, :].transpose(1, 2)
        v = qkv[..., -1, :].transpose(1, 2)
        q = q * cos + _rotate_half(q) * sin
        k = k * cos + _rotate_half(k) * sin
        k = k[:, :, None].expand(B, nkv, groups, N, hd).reshape(B, nh, N, hd)
        v = v[:, :, None].expand(B, nkv, groups, N, hd).reshape(B, nh, N, hd)
        a = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(hd) + mask
        a = F.softmax(a, dim=-1, dtype=torch.float32)
        o = torch.matmul(a, v).transpose(1, 2).reshape(B, N, H)
        x = x + F.linear(o, sd[lp + "attention.wo.weight"])
        h = rms_norm(x, sd[lp + "ffn_norm.weight"], eps)
        ff = F.linear(F.silu(F.linear(h, sd[lp + "feed_forward.w1.weight"])) * F.linear(h, sd[lp + "feed_forward.w3.weight"]),
                      sd[lp + "feed_forward.w2.weight"])
        x = x + ff
    x = rms_norm(x, sd[lm + "norm.weight"], eps)
    return x, F.linear(x, sd["language_model.output.weight"]).float()


def mlpscore(x: torch.Tensor, sd):
    """-> (score1 [B], the last layer's pre-ReLU value [B])"""
    for l in range(1, 5):
        x = F.relu(F.linear(x, sd[f"mlpscore.fc{l}.weight"], sd[f"mlpscore.fc{l}.bias"]))
    pre = F.linear(x, sd["mlpscore.fc5.weight"], sd["mlpscore.fc5.bias"])
    return F.relu(pre).squeeze(1), pre.squeeze(1)


def to_torch(sd_np) -> dict:
    return {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)) for k, v in sd_np.items()}


@torch.no_grad()
def stage2_forward(vit_embeds, pixel_values, input_ids, sd, llm: dict, img_context_token_id: int) -> dict:
    """The stage-2 forward above the seam. vit_embeds: [n_frames, 256, C] (reference's or the CUDA path's); pixel_values
    feed only the motion stand-in. Returns score1, its pre-ReLU value, hidden_states[-1][:, -4] and argmax of the shifted
    logits (eval2 :469-488)."""
    vit_embeds = torch.as_tensor(vit_embeds, dtype=torch.float32)
    pixel_values = torch.as_tensor(pixel_values, dtype=torch.float32)
    input_ids = torch.as_tensor(input_ids, dtype=torch.long)
    B = input_ids.shape[0]
    motion = motion_embeds(fake_slowfast(pixel_values, B), sd)
    emb = splice(input_ids, sd["language_model.model.tok_embeddings.weight"], vit_embeds, motion, img_context_token_id)
    hidden, logits = internlm2(emb, sd, llm)
    score1, pre = mlpscore(hidden[:, -4, :], sd)
    shift_logits = logits[..., :-1, :].reshape(-1, logits.shape[-1])
    return dict(score1=score1.numpy(), score_pre=pre.numpy(), hidden_m4=hidden[:, -4, :].numpy(),
                logit=torch.argmax(shift_logits, dim=1).numpy(), logits_tail=shift_logits[-48:].numpy())
