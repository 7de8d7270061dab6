"""ORACLE (test infrastructure, NOT product code): the same restatement as oracle/vit_oracle.py, written on torch CPU
fp32 tensors so it runs multi-threaded (MKL GEMMs + parallel elementwise) like the reference's own CPU path does.

It exists to give bench.py a CPU baseline that is as fast as the reference would be on the same cores (the numpy
restatement is single-threaded outside BLAS); tests pin it to the same reference goldens. Only tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.

VIT = /root/reference/internvl/model/internvl_chat/modeling_intern_vit.py,
EVAL2 = /root/reference/internvl/model/internvl_chat_eval2/modeling_internvl_chat.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def to_torch(sd):
    return {k: torch.from_numpy(v) for k, v in sd.items()}


def _rms(x, w, eps):  # VIT:32-43
    return w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))


def _norm(x, sd, name, cfg):  # VIT:208-209
    if cfg.norm_type == "layer_norm":
        return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], cfg.layer_norm_eps)
    return _rms(x, sd[name + ".weight"], cfg.layer_norm_eps)


def embeddings(pv, sd, cfg, pre="vision_model."):  # VIT:95-107
    n = pv.shape[0]
    x = F.conv2d(pv, sd[pre + "embeddings.patch_embedding.weight"], sd[pre + "embeddings.patch_embedding.bias"],
                 stride=cfg.patch_size).flatten(2).transpose(1, 2)
    cls = sd[pre + "embeddings.class_embedding"].expand(n, 1, -1)
    return torch.cat([cls, x], dim=1) + sd[pre + "embeddings.position_embedding"]


def _flash_attention(x, sd, lp, cfg):
    """The reference's GPU attention path: InternAttention._flash_attn (VIT:162-177) -> FlashAttention.forward
    (flash_attention.py:44-56): qkv packed as [(B*S), 3, H, D], cu_seqlens = arange(0, (B+1)*S, S), non-causal,
    softmax_scale None (= D^-0.5), dropout 0. Needs the flash-attn wheel, CUDA and fp16/bf16 (flash_attention.py:41-42).
    Used by bench.py's reference_gpu_stack comparator only."""
    from flash_attn import flash_attn_varlen_qkvpacked_func
    n, s, d = x.shape
    h, hd = cfg.num_heads, cfg.head_dim
    qkv = F.linear(x, sd[lp + "attn.qkv.weight"], sd.get(lp + "attn.qkv.bias")).view(n, s, 3, h, hd)
    if cfg.qk_normalization:  # VIT:166-170
        q, k, v = qkv.unbind(2)
        q = _rms(q.flatten(-2, -1), sd[lp + "attn.q_norm.weight"], cfg.layer_norm_eps).view(n, s, h, hd)
        k = _rms(k.flatten(-2, -1), sd[lp + "attn.k_norm.weight"], cfg.layer_norm_eps).view(n, s, h, hd)
        qkv = torch.stack([q, k, v], dim=2)
    cu = torch.arange(0, (n + 1) * s, step=s, dtype=torch.int32, device=x.device)
    ctx = flash_attn_varlen_qkvpacked_func(qkv.reshape(n * s, 3, h, hd), cu, s, 0.0, softmax_scale=None, causal=False)
    return F.linear(ctx.reshape(n, s, d), sd[lp + "attn.proj.weight"], sd[lp + "attn.proj.bias"])


def attention(x, sd, lp, cfg):  # VIT:143-160 (_naive_attn)
    if getattr(cfg, "_use_flash_attn", False) or sd.get("__use_flash_attn__") is not None:
        return _flash_attention(x, sd, lp, cfg)
    n, s, d = x.shape
    h, hd = cfg.num_heads, cfg.head_dim
    qkv = F.linear(x, sd[lp + "attn.qkv.weight"], sd.get(lp + "attn.qkv.bias"))
    q, k, v = qkv.reshape(n, s, 3, h, hd).permute(2, 0, 3, 1, 4).unbind(0)
    if cfg.qk_normalization:
        q = _rms(q.transpose(1, 2).flatten(-2, -1), sd[lp + "attn.q_norm.weight"], cfg.layer_norm_eps)
        q = q.view(n, s, h, hd).transpose(1, 2)
        k = _rms(k.transpose(1, 2).flatten(-2, -1), sd[lp + "attn.k_norm.weight"], cfg.layer_norm_eps)
        k = k.view(n, s, h, hd).transpose(1, 2)
    a = ((q * hd ** -0.5) @ k.transpose(-2, -1)).softmax(dim=-1)
    ctx = (a @ v).transpose(1, 2).reshape(n, s, d)
    return F.linear(ctx, sd[lp + "attn.proj.weight"], sd[lp + "attn.proj.bias"])


def encoder_layer(x, sd, lp, cfg):  # VIT:224-226
    x = x + attention(_norm(x, sd, lp + "norm1", cfg), sd, lp, cfg) * sd[lp + "ls1"]
    hmid = F.gelu(F.linear(_norm(x, sd, lp + "norm2", cfg), sd[lp + "mlp.fc1.weight"], sd[lp + "mlp.fc1.bias"]))
    return x + F.linear(hmid, sd[lp + "mlp.fc2.weight"], sd[lp + "mlp.fc2.bias"]) * sd[lp + "ls2"]


def pixel_shuffle(x, scale_factor=0.5, ps_version="v2"):  # EVAL2:492-506
    n, w, h, c = x.size()
    x = x.view(n, w, int(h * scale_factor), int(c / scale_factor)).permute(0, 2, 1, 3).contiguous()
    x = x.view(n, int(h * scale_factor), int(w * scale_factor), int(c / (scale_factor * scale_factor)))
    if ps_version != "v1":
        x = x.permute(0, 2, 1, 3).contiguous()
    return x


@torch.no_grad()
def extract_feature(pv, sd, cfg, return_vit=False):  # EVAL2:508-531 with select_layer == -1
    if pv.dim() != 4:
        raise ValueError(f"wrong pixel_values size: {pv.shape}")
    x = embeddings(pv, sd, cfg)
    for i in range(cfg.num_layers):
        x = encoder_layer(x, sd, f"vision_model.encoder.layers.{i}.", cfg)
    vit = x
    x = x[:, 1:, :]
    hw = int(x.shape[1] ** 0.5)
    x = pixel_shuffle(x.reshape(x.shape[0], hw, hw, -1), cfg.downsample_ratio, cfg.ps_version)
    x = x.reshape(x.shape[0], -1, x.shape[-1])
    x = F.layer_norm(x, (x.shape[-1],), sd["mlp1.0.weight"], sd["mlp1.0.bias"], cfg.mlp1_ln_eps)
    x = F.linear(F.gelu(F.linear(x, sd["mlp1.1.weight"], sd["mlp1.1.bias"])), sd["mlp1.3.weight"], sd["mlp1.3.bias"])
    return (x, vit) if return_vit else x
