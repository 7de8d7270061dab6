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


def make_weights(cfg: VisionPathConfig, seed: int = 0) -> dict[str, np.ndarray]:
    rng = np.random.Generator(np.random.PCG64(seed))
    D, I, C, W = cfg.hidden_size, cfg.intermediate_size, cfg.llm_hidden, cfg.shuffled_width
    P = cfg.patch_size
    sd: dict[str, np.ndarray] = {}

    def uniform(shape, fan_in):
        b = 1.0 / np.sqrt(fan_in)
        return bf16_round(rng.uniform(-b, b, size=shape).astype(np.float32))

    def normal(shape, std=1.0, mean=0.0):
        return bf16_round((mean + std * rng.standard_normal(size=shape)).astype(np.float32))

    pre = "vision_model."
    sd[pre + "embeddings.class_embedding"] = normal((1, 1, D))
    sd[pre + "embeddings.position_embedding"] = normal((1, cfg.tokens, D))
    sd[pre + "embeddings.patch_embedding.weight"] = uniform((D, 3, P, P), 3 * P * P)
    sd[pre + "embeddings.patch_embedding.bias"] = uniform((D,), 3 * P * P)
    for i in range(cfg.num_layers):
        lp = f"{pre}encoder.layers.{i}."
        sd[lp + "ls1"] = normal((D,), 0.1 * cfg.initializer_factor, cfg.initializer_factor)
        sd[lp + "ls2"] = normal((D,), 0.1 * cfg.initializer_factor, cfg.initializer_factor)
        sd[lp + "attn.qkv.weight"] = uniform((3 * D, D), D)
        if cfg.qkv_bias:
            sd[lp + "attn.qkv.bias"] = uniform((3 * D,), D)
        if cfg.qk_normalization:
            sd[lp + "attn.q_norm.weight"] = normal((D,), 0.1, 1.0)
            sd[lp + "attn.k_norm.weight"] = normal((D,), 0.1, 1.0)
        sd[lp + "attn.proj.weight"] = uniform((D, D), D)
        sd[lp + "attn.proj.bias"] = uniform((D,), D)
        sd[lp + "mlp.fc1.weight"] = uniform((I, D), D)
        sd[lp + "mlp.fc1.bias"] = uniform((I,), D)
        sd[lp + "mlp.fc2.weight"] = uniform((D, I), I)
        sd[lp + "mlp.fc2.bias"] = uniform((D,), I)
        sd[lp + "norm1.weight"] = normal((D,), 0.1, 1.0)
        sd[lp + "norm2.weight"] = normal((D,), 0.1, 1.0)
        if cfg.norm_type == "layer_norm":
            sd[lp + "norm1.bias"] = normal((D,), 0.1)
            sd[lp + "norm2.bias"] = normal((D,), 0.1)
    sd["mlp1.0.weight"] = normal((W,), 0.1, 1.0)
    sd["mlp1.0.bias"] = normal((W,), 0.1)
    sd["mlp1.1.weight"] = uniform((C, W), W)
    sd["mlp1.1.bias"] = uniform((C,), W)
    sd["mlp1.3.weight"] = uniform((C, C), C)
    sd["mlp1.3.bias"] = uniform((C,), C)
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
