"""Seeded synthetic weights and frames for tests, smoke and bench (there are no checkpoints offline).

Weights are produced under the reference's own state-dict names (SURVEY.md section 8b "Weight ABI";
/root/reference/internvl/model/internvl_chat/modeling_intern_vit.py:66-85,110-141,184-214 and
/root/reference/internvl/model/internvl_chat_eval2/modeling_internvl_chat.py:238-243) so the same dict loads into
the reference modules (tests/golden/make_golden.py), the numpy oracle and the CUDA engine.

All values are rounded to bf16-representable fp32 so that casting to bf16 is exact on every side.
numpy's PCG64 stream is platform independent, so the GPU box regenerates identical tensors from the seed.
"""
from __future__ import annotations

import numpy as np

from .config import VisionPathConfig


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 -> nearest-even bf16, returned as fp32 (bf16-representable)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32)
    rounding = ((u >> 16) & 1) + np.uint32(0x7FFF)
    r = ((u + rounding) & np.uint32(0xFFFF0000)).astype(np.uint32)
    return r.view(np.float32).reshape(x.shape)


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 bit pattern (uint16), round-to-nearest-even."""
    return (bf16_round(x).view(np.uint32) >> 16).astype(np.uint16)


def from_bf16_bits(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


def weight_spec(cfg: VisionPathConfig):
    """Yield (name, shape, dist) for every tensor of the path under the reference's state-dict names.

    dist is ("uniform", fan_in) -- U(-1/sqrt(fan_in), 1/sqrt(fan_in)), torch's Linear/Conv default -- or
    ("normal", std, mean). make_weights() draws them in this order; GPU-side generators (tools/bench_6b.py) reuse it.
    """
    D, I, C, W = cfg.hidden_size, cfg.intermediate_size, cfg.llm_hidden, cfg.shuffled_width
    P = cfg.patch_size
    pre = "vision_model."
    yield pre + "embeddings.class_embedding", (1, 1, D), ("normal", 1.0, 0.0)
    yield pre + "embeddings.position_embedding", (1, cfg.tokens, D), ("normal", 1.0, 0.0)
    yield pre + "embeddings.patch_embedding.weight", (D, 3, P, P), ("uniform", 3 * P * P)
    yield pre + "embeddings.patch_embedding.bias", (D,), ("uniform", 3 * P * P)
    for i in range(cfg.num_layers):
        lp = f"{pre}encoder.layers.{i}."
        yield lp + "ls1", (D,), ("normal", 0.1 * cfg.initializer_factor, cfg.initializer_factor)
        yield lp + "ls2", (D,), ("normal", 0.1 * cfg.initializer_factor, cfg.initializer_factor)
        yield lp + "attn.qkv.weight", (3 * D, D), ("uniform", D)
        if cfg.qkv_bias:
            yield lp + "attn.qkv.bias", (3 * D,), ("uniform", D)
        if cfg.qk_normalization:
            yield lp + "attn.q_norm.weight", (D,), ("normal", 0.1, 1.0)
            yield lp + "attn.k_norm.weight", (D,), ("normal", 0.1, 1.0)
        yield lp + "attn.proj.weight", (D, D), ("uniform", D)
        yield lp + "attn.proj.bias", (D,), ("uniform", D)
        yield lp + "mlp.fc1.weight", (I, D), ("uniform", D)
        yield lp + "mlp.fc1.bias", (I,), ("uniform", D)
        yield lp + "mlp.fc2.weight", (D, I), ("uniform", I)
        yield lp + "mlp.fc2.bias", (D,), ("uniform", I)
        yield lp + "norm1.weight", (D,), ("normal", 0.1, 1.0)
        yield lp + "norm2.weight", (D,), ("normal", 0.1, 1.0)
        if cfg.norm_type == "layer_norm":
            yield lp + "norm1.bias", (D,), ("normal", 0.1, 0.0)
            yield lp + "norm2.bias", (D,), ("normal", 0.1, 0.0)
    yield "mlp1.0.weight", (W,), ("normal", 0.1, 1.0)
    yield "mlp1.0.bias", (W,), ("normal", 0.1, 0.0)
    yield "mlp1.1.weight", (C, W), ("uniform", W)
    yield "mlp1.1.bias", (C,), ("uniform", W)
    yield "mlp1.3.weight", (C, C), ("uniform", C)
    yield "mlp1.3.bias", (C,), ("uniform", C)


def make_weights(cfg: VisionPathConfig, seed: int = 0) -> dict[str, np.ndarray]:
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: dict[str, np.ndarray] = {}
    for name, shape, dist in weight_spec(cfg):
        if dist[0] == "uniform":
            b = 1.0 / np.sqrt(dist[1])
            sd[name] = bf16_round(rng.uniform(-b, b, size=shape).astype(np.float32))
        else:
            sd[name] = bf16_round((dist[2] + dist[1] * rng.standard_normal(size=shape)).astype(np.float32))
    return sd


def make_stress_weights(cfg: VisionPathConfig, seed: int = 0) -> dict[str, np.ndarray]:
    """make_weights pushed towards what TRAINED towers look like, for the pieces whose numerics differ there: qkv weights
    x4 (peaked attention rows: the stale-max softmax and its rescale path), a few outlier channels of +-300 in the
    position embedding (massive activations: the E[x^2] - mean^2 LayerNorm statistics of the fused epilogues), layer
    scales spread over 0.1 .. 1."""
    sd = make_weights(cfg, seed)
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    D = cfg.hidden_size
    pos = sd["vision_model.embeddings.position_embedding"].copy()
    for ch, val in ((5 % D, 300.0), (33 % D, -300.0), (77 % D, 200.0)):
        pos[..., ch] += np.float32(val)
    sd["vision_model.embeddings.position_embedding"] = bf16_round(pos)
    for i in range(cfg.num_layers):
        lp = f"vision_model.encoder.layers.{i}."
        sd[lp + "attn.qkv.weight"] = bf16_round(sd[lp + "attn.qkv.weight"] * np.float32(4.0))
        for ls in ("ls1", "ls2"):
            sd[lp + ls] = bf16_round(rng.uniform(0.1, 1.0, size=(D,)).astype(np.float32))
    return sd


def make_frames(n: int, height: int, width: int, seed: int = 1, kind: str = "smooth") -> np.ndarray:
    """uint8 HWC frames [n,H,W,3] (decord layout, /root/reference/internvl/train/internvl/eval/stage1_eval.py:452).

    kind="uniform": iid uniform bytes (worst case for resampling); "smooth": low-frequency sinusoids + noise.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    if kind == "uniform":
        return rng.integers(0, 256, size=(n, height, width, 3), dtype=np.uint8)
    yy = np.arange(height, dtype=np.float32)[:, None] / height
    xx = np.arange(width, dtype=np.float32)[None, :] / width
    out = np.empty((n, height, width, 3), dtype=np.uint8)
    for f in range(n):
        for c in range(3):
            img = np.full((height, width), 128.0, dtype=np.float32)
            for _ in range(4):
                fx, fy = rng.uniform(0.5, 6.0, size=2)
                ph = rng.uniform(0, 2 * np.pi)
                amp = rng.uniform(10, 40)
                img += amp * np.sin(2 * np.pi * (fx * xx + fy * yy) + ph).astype(np.float32)
            img += rng.normal(0, 8.0, size=(height, width)).astype(np.float32)
            out[f, :, :, c] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return out


def make_pixel_values(n: int, cfg: VisionPathConfig, seed: int = 1) -> np.ndarray:
    """Already-normalised pixel_values [n,3,S,S] (bf16-representable fp32): ImageNet-normalised random bytes."""
    rng = np.random.Generator(np.random.PCG64(seed))
    u = rng.integers(0, 256, size=(n, 3, cfg.image_size, cfg.image_size), dtype=np.uint8)
    return bf16_round(normalize_lut()[np.arange(3)[None, :, None, None], u])


IMAGENET_MEAN = (0.485, 0.456, 0.406)  # /root/reference/internvl/train/constants.py:10
IMAGENET_STD = (0.229, 0.224, 0.225)   # /root/reference/internvl/train/constants.py:11


def normalize_lut() -> np.ndarray:
    """[3,256] fp32 table of Normalize(ToTensor(u8)) = ((u/255) - mean)/std computed in fp32 like torchvision
    (/root/reference/internvl/train/dataset.py:272-273)."""
    u = np.arange(256, dtype=np.float32) / np.float32(255.0)
    lut = np.empty((3, 256), dtype=np.float32)
    for c in range(3):
        lut[c] = (u - np.float32(IMAGENET_MEAN[c])) / np.float32(IMAGENET_STD[c])
    return lut


# ----------------------------------------------------------------------------------------------------------------------
# Everything ABOVE the seam (test fixtures only): a small InternLM2 + motion_mlp + mlpscore under the reference's
# state-dict names, so that score parity can be checked through the unchanged caller of extract_feature
# (/root/reference/internvl/model/internvl_chat_eval2/modeling_internvl_chat.py:238-250,327-488). The reference hard-codes
# 4096 for motion_mlp / mlpscore (eval2 :244-249,44), so the LLM width is 4096; depth / FFN width / vocabulary are small.
HEAD_LLM = dict(hidden_size=4096, num_hidden_layers=2, num_attention_heads=32, num_key_value_heads=8,
                intermediate_size=1024, vocab_size=512, rms_norm_eps=1e-5, rope_theta=1000000.0,
                max_position_embeddings=32768)


def head_weight_spec(llm: dict = HEAD_LLM):
    H, I, V = llm["hidden_size"], llm["intermediate_size"], llm["vocab_size"]
    hd = H // llm["num_attention_heads"]
    qkv = (llm["num_attention_heads"] + 2 * llm["num_key_value_heads"]) * hd
    lm = "language_model.model."
    yield lm + "tok_embeddings.weight", (V, H), ("normal", 0.5, 0.0)
    for i in range(llm["num_hidden_layers"]):
        lp = f"{lm}layers.{i}."
        yield lp + "attention.wqkv.weight", (qkv, H), ("uniform", H / 8.0)   # gain 2.8: attention that is not flat, so score1 responds to the visual tokens
        yield lp + "attention.wo.weight", (H, H), ("uniform", H)
        yield lp + "feed_forward.w1.weight", (I, H), ("uniform", H)
        yield lp + "feed_forward.w3.weight", (I, H), ("uniform", H)
        yield lp + "feed_forward.w2.weight", (H, I), ("uniform", I)
        yield lp + "attention_norm.weight", (H,), ("normal", 0.1, 1.0)
        yield lp + "ffn_norm.weight", (H,), ("normal", 0.1, 1.0)
    yield lm + "norm.weight", (H,), ("normal", 0.1, 1.0)
    yield "language_model.output.weight", (V, H), ("uniform", H)
    yield "motion_mlp.0.weight", (2304,), ("normal", 0.1, 1.0)
    yield "motion_mlp.0.bias", (2304,), ("normal", 0.1, 0.0)
    yield "motion_mlp.1.weight", (4096, 2304), ("uniform", 2304)
    yield "motion_mlp.1.bias", (4096,), ("uniform", 2304)
    yield "motion_mlp.3.weight", (4096, 4096), ("uniform", 4096)
    yield "motion_mlp.3.bias", (4096,), ("uniform", 4096)
    # mlpscore (eval2 :43-94): ReLU after EVERY layer incl. the last, so zero-mean weights give score1 == 0. Weights are
    # U(-b, b) as usual; the biases are positive so units stay alive, and fc5's weights are positive (score1 > 0).
    dims = [H, 1024, 256, 64, 16, 1]
    for l in range(5):
        yield f"mlpscore.fc{l + 1}.weight", (dims[l + 1], dims[l]), (("uniform", dims[l] / 4.0) if l < 4 else ("posuniform", dims[l]))
        yield f"mlpscore.fc{l + 1}.bias", (dims[l + 1],), ("normal", 0.02, 0.1 if l < 4 else 0.0)
    yield "mlpscore.ln1.weight", (1024,), ("normal", 0.0, 1.0)   # constructed by the reference, never used in forward
    yield "mlpscore.ln1.bias", (1024,), ("normal", 0.0, 0.0)


def make_head_weights(seed: int = 0, llm: dict = HEAD_LLM) -> dict[str, np.ndarray]:
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: dict[str, np.ndarray] = {}
    for name, shape, dist in head_weight_spec(llm):
        if dist[0] == "uniform":
            b = 1.0 / np.sqrt(dist[1])
            sd[name] = bf16_round(rng.uniform(-b, b, size=shape).astype(np.float32))
        elif dist[0] == "posuniform":
            sd[name] = bf16_round(rng.uniform(0.0, 2.0 / np.sqrt(dist[1]), size=shape).astype(np.float32))
        else:
            sd[name] = bf16_round((dist[2] + dist[1] * rng.standard_normal(size=shape)).astype(np.float32))
    return sd


IMG_CONTEXT_ID, IMG_START_ID, IMG_END_ID = 7, 10, 11


def make_clip_prompt(n_frames: int, tokens_per_frame: int = 256, n_text: int = 40, seed: int = 0,
                     vocab_size: int = 512) -> np.ndarray:
    """Token ids with the call shape of the video prompt (stage1_eval.py:477-494): <s>, per frame <img> + 256 x
    <IMG_CONTEXT> + </img>, one more <img><IMG_CONTEXT></img> for the motion token, then text."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ids = [1]
    for _ in range(n_frames):
        ids += [IMG_START_ID] + [IMG_CONTEXT_ID] * tokens_per_frame + [IMG_END_ID]
    ids += [IMG_START_ID, IMG_CONTEXT_ID, IMG_END_ID]
    ids += [int(t) for t in rng.integers(20, vocab_size, size=n_text)]
    return np.asarray([ids], dtype=np.int64)


def make_slowfast_weights(seed: int = 0) -> dict[str, np.ndarray]:
    """Seeded stand-in for pytorchvideo's slowfast_r50 checkpoint (absent here): He-scaled conv weights, BatchNorm
    statistics around identity (gamma ~ 1, beta ~ 0, running_mean ~ 0, running_var ~ 1) so activations stay O(1) through the
    50 layers. Names / shapes: aigv_assessor_b200.slowfast.layer_table()."""
    from .slowfast import layer_table
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: dict[str, np.ndarray] = {}
    for wname, nname, shape in layer_table():
        fan_in = shape[1] * shape[2] * shape[3] * shape[4]
        gain = 1.0 if wname.endswith("conv_c.weight") or "branch1" in wname else np.sqrt(2.0)
        sd[wname] = bf16_round((gain / np.sqrt(fan_in) * rng.standard_normal(size=shape)).astype(np.float32))
        c = shape[0]
        sd[nname + ".weight"] = (1.0 + 0.1 * rng.standard_normal(size=(c,))).astype(np.float32)
        sd[nname + ".bias"] = (0.05 * rng.standard_normal(size=(c,))).astype(np.float32)
        sd[nname + ".running_mean"] = (0.05 * rng.standard_normal(size=(c,))).astype(np.float32)
        sd[nname + ".running_var"] = (1.0 + 0.1 * rng.random(size=(c,))).astype(np.float32)
    return sd
