"""ORACLE (test infrastructure, NOT product code): numpy fp32 restatement of the reference's
``extract_feature`` path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this module. The product path (aigv_assessor_b200/) never does.

Parity pinning: the reference ships no tests or golden vectors for this path, so this
restatement is pinned against the reference's own Python modules executed in the build container on seeded
weights: ``tests/golden/make_golden.py`` (committed) imports /root/reference, runs it, and stores the fixtures
under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against them.

Each function cites the reference lines it follows. VIT = /root/reference/internvl/model/internvl_chat/
modeling_intern_vit.py, EVAL2 = /root/reference/internvl/model/internvl_chat_eval2/modeling_internvl_chat.py.
Weights are a dict keyed by the reference's state-dict names (fp32 numpy arrays).
"""
from __future__ import annotations

import numpy as np
from scipy.special import erf as _erf

F32 = np.float32


def layer_norm(x, weight, bias, eps):
    """torch.nn.LayerNorm over the last dim (VIT:208-209 norm1/norm2; EVAL2:239 mlp1[0]); biased variance."""
    x = x.astype(F32, copy=False)
    mu = x.mean(axis=-1, keepdims=True, dtype=F32)
    xc = x - mu
    var = (xc * xc).mean(axis=-1, keepdims=True, dtype=F32)
    y = xc / np.sqrt(var + F32(eps))
    return (y * weight + bias).astype(F32)


def rms_norm(x, weight, eps):
    """InternRMSNorm (VIT:32-43): x * rsqrt(mean(x^2) + eps) * weight, stats in fp32."""
    x = x.astype(F32, copy=False)
    var = (x * x).mean(axis=-1, keepdims=True, dtype=F32)
    return (weight * (x / np.sqrt(var + F32(eps)))).astype(F32)


def gelu(x):
    """ACT2FN['gelu'] / nn.GELU(): exact erf GELU (VIT:188; EVAL2:241)."""
    x = x.astype(F32, copy=False)
    return (x * F32(0.5) * (F32(1.0) + _erf(x * F32(0.7071067811865476)).astype(F32))).astype(F32)


def linear(x, w, b=None):
    """nn.Linear: x @ w.T + b with w stored [out,in] (VIT:129,141,189,190; EVAL2:240,242)."""
    y = x.astype(F32, copy=False) @ w.T
    if b is not None:
        y = y + b
    return y.astype(F32, copy=False)


def embeddings(pixel_values, sd, cfg, pre="vision_model."):
    """InternVisionEmbeddings.forward (VIT:95-107). Conv2d k=s=P is a GEMM over (c,ky,kx)-flattened patches;
    tokens are row-major over the patch grid (flatten(2).transpose(1,2)); cls prepended; position embedding added.
    `_get_pos_embed` (VIT:87-93) is the identity when the grid equals image_size//patch_size (probed in SURVEY 6)."""
    n, c, hh, ww = pixel_values.shape
    P, g = cfg.patch_size, cfg.grid
    assert hh == ww == cfg.image_size, "oracle restates the native-resolution path only"
    w = sd[pre + "embeddings.patch_embedding.weight"].reshape(cfg.hidden_size, -1)
    b = sd[pre + "embeddings.patch_embedding.bias"]
    x = pixel_values.astype(F32).reshape(n, c, g, P, g, P).transpose(0, 2, 4, 1, 3, 5).reshape(n, g * g, c * P * P)
    pe = x @ w.T + b
    cls = np.broadcast_to(sd[pre + "embeddings.class_embedding"].reshape(1, 1, -1), (n, 1, cfg.hidden_size))
    out = np.concatenate([cls, pe], axis=1) + sd[pre + "embeddings.position_embedding"]
    return out.astype(F32)


def _norm(x, sd, name, cfg):
    if cfg.norm_type == "layer_norm":
        return layer_norm(x, sd[name + ".weight"], sd[name + ".bias"], cfg.layer_norm_eps)
    return rms_norm(x, sd[name + ".weight"], cfg.layer_norm_eps)


def attention(x, sd, lp, cfg):
    """InternAttention._naive_attn (VIT:143-160): qkv channel order (three, head, d); q scaled BEFORE q@k^T
    (VIT:153); softmax over keys; no mask; dropout p=0; then proj (VIT:158)."""
    n, s, d = x.shape
    h, hd = cfg.num_heads, cfg.head_dim
    qkv = linear(x, sd[lp + "attn.qkv.weight"], sd.get(lp + "attn.qkv.bias"))
    qkv = qkv.reshape(n, s, 3, h, hd).transpose(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]  # [n,h,s,hd]
    if cfg.qk_normalization:  # VIT:148-151: RMSNorm over all heads flattened
        q = rms_norm(q.transpose(0, 2, 1, 3).reshape(n, s, d), sd[lp + "attn.q_norm.weight"], cfg.layer_norm_eps)
        q = q.reshape(n, s, h, hd).transpose(0, 2, 1, 3)
        k = rms_norm(k.transpose(0, 2, 1, 3).reshape(n, s, d), sd[lp + "attn.k_norm.weight"], cfg.layer_norm_eps)
        k = k.reshape(n, s, h, hd).transpose(0, 2, 1, 3)
    scale = F32(hd ** -0.5)
    ctx = np.empty((n, h, s, hd), dtype=F32)
    for i in range(n):
        for j in range(h):
            a = (q[i, j] * scale) @ k[i, j].T
            a = a - a.max(axis=-1, keepdims=True)
            e = np.exp(a, dtype=F32)
            p = e / e.sum(axis=-1, keepdims=True, dtype=F32)
            ctx[i, j] = p @ v[i, j]
    ctx = ctx.transpose(0, 2, 1, 3).reshape(n, s, d)
    return linear(ctx, sd[lp + "attn.proj.weight"], sd[lp + "attn.proj.bias"])


def mlp(x, sd, lp):
    """InternMLP.forward (VIT:192-196): fc1 -> gelu -> fc2."""
    hmid = gelu(linear(x, sd[lp + "mlp.fc1.weight"], sd[lp + "mlp.fc1.bias"]))
    return linear(hmid, sd[lp + "mlp.fc2.weight"], sd[lp + "mlp.fc2.bias"])


def encoder_layer(x, sd, lp, cfg):
    """InternVisionEncoderLayer.forward (VIT:224-226): x + attn(norm1(x))*ls1 ; x + mlp(norm2(x))*ls2
    (drop_path is the identity in eval)."""
    x = x + attention(_norm(x, sd, lp + "norm1", cfg), sd, lp, cfg) * sd[lp + "ls1"]
    x = x + mlp(_norm(x, sd, lp + "norm2", cfg), sd, lp) * sd[lp + "ls2"]
    return x.astype(F32)


def vision_model(pixel_values, sd, cfg, pre="vision_model.", return_stages=False):
    """InternVisionModel.forward (VIT:324-362) with select_layer == -1: embeddings, L layers, no final norm."""
    if pixel_values.ndim != 4:
        raise ValueError(f"wrong pixel_values size: {pixel_values.shape}")  # VIT:345
    x = embeddings(pixel_values, sd, cfg, pre)
    stages = {"embeddings": x} if return_stages else None
    for i in range(cfg.num_layers):
        x = encoder_layer(x, sd, f"{pre}encoder.layers.{i}.", cfg)
        if return_stages:
            stages[f"layer{i}"] = x
    return (x, stages) if return_stages else x


def pixel_shuffle(x, scale_factor=0.5, ps_version="v2"):
    """InternVLChatModel.pixel_shuffle (EVAL2:492-506), restated view-for-view."""
    n, w, h, c = x.shape
    x = x.reshape(n, w, int(h * scale_factor), int(c / scale_factor))
    x = np.ascontiguousarray(x.transpose(0, 2, 1, 3))
    x = x.reshape(n, int(h * scale_factor), int(w * scale_factor), int(c / (scale_factor * scale_factor)))
    if ps_version != "v1":
        x = np.ascontiguousarray(x.transpose(0, 2, 1, 3))
    return x


def mlp1(x, sd, cfg):
    """mlp1 = LayerNorm(4D) -> Linear(4D,C) -> GELU -> Linear(C,C) (EVAL2:238-243)."""
    x = layer_norm(x, sd["mlp1.0.weight"], sd["mlp1.0.bias"], cfg.mlp1_ln_eps)
    x = gelu(linear(x, sd["mlp1.1.weight"], sd["mlp1.1.bias"]))
    return linear(x, sd["mlp1.3.weight"], sd["mlp1.3.bias"])


def extract_feature(pixel_values, sd, cfg, flatten=True, return_stages=False):
    """InternVLChatModel.extract_feature (EVAL2:508-531): ViT last hidden state, drop cls, reshape to the patch
    grid, pixel_shuffle, flatten to [N,256,4D] (stage-2 family; CHAT:406-408 keeps [N,16,16,C]), mlp1."""
    assert cfg.select_layer == -1
    vit = vision_model(pixel_values, sd, cfg, return_stages=return_stages)
    stages = None
    if return_stages:
        vit, stages = vit
    x = vit[:, 1:, :]
    hw = int(x.shape[1] ** 0.5)
    x = x.reshape(x.shape[0], hw, hw, -1)
    x = pixel_shuffle(x, scale_factor=cfg.downsample_ratio, ps_version=cfg.ps_version)
    if return_stages:
        stages["pixel_shuffle"] = x
    if flatten:
        x = x.reshape(x.shape[0], -1, x.shape[-1])
    y = mlp1(x, sd, cfg)
    return (y, stages) if return_stages else y


def error_metrics(y, ref):
    """Parity metric of SURVEY.md section 8(d): e_max = max|y-ref|/max|ref|, e_l2 = ||y-ref||/||ref||."""
    y = np.asarray(y, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    e_max = float(np.abs(y - ref).max() / np.abs(ref).max())
    e_l2 = float(np.linalg.norm((y - ref).ravel()) / np.linalg.norm(ref.ravel()))
    return e_max, e_l2
