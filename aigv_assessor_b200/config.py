"""Shape/config source of truth for the vision hot path.

Mirrors the fields of the reference's ``InternVisionConfig``
(/root/reference/internvl/model/internvl_chat/configuration_intern_vit.py:62-107) that the
``extract_feature`` path reads, plus the projector width (``llm_hidden``) that
``InternVLChatModel.__init__`` derives from ``llm_config.hidden_size``
(/root/reference/internvl/model/internvl_chat_eval2/modeling_internvl_chat.py:235-243).
"""
from __future__ import annotations

from dataclasses import dataclass, asdict

NORM_LAYER = 0  # nn.LayerNorm (InternViT-300M, config.json:163)
NORM_RMS = 1    # InternRMSNorm (InternViT-6B default, configuration_intern_vit.py:75)


@dataclass(frozen=True)
class VisionPathConfig:
    hidden_size: int = 1024
    num_heads: int = 16
    intermediate_size: int = 4096
    num_layers: int = 24
    image_size: int = 448
    patch_size: int = 14
    llm_hidden: int = 2048          # mlp1 output width C
    qkv_bias: bool = True
    qk_normalization: bool = False
    norm_type: str = "layer_norm"   # or "rms_norm"
    layer_norm_eps: float = 1e-6
    mlp1_ln_eps: float = 1e-5       # nn.LayerNorm default in mlp1[0]
    initializer_factor: float = 1.0  # ls1/ls2 init
    downsample_ratio: float = 0.5
    ps_version: str = "v2"
    select_layer: int = -1

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def num_patches(self) -> int:
        return self.grid * self.grid

    @property
    def tokens(self) -> int:
        return self.num_patches + 1

    @property
    def patch_k(self) -> int:
        return 3 * self.patch_size * self.patch_size

    @property
    def out_tokens(self) -> int:
        return int(self.grid * self.downsample_ratio) ** 2

    @property
    def shuffled_width(self) -> int:
        return self.hidden_size * int(1 / self.downsample_ratio) ** 2

    def flops_per_frame(self) -> float:
        """Algorithmic FLOPs (2*M*N*K) per 448x448 tile; matches SURVEY.md section 8(d)."""
        S, D, I, L, C = self.tokens, self.hidden_size, self.intermediate_size, self.num_layers, self.llm_hidden
        per_layer = 2 * S * D * 3 * D + 2 * S * D * D + 2 * 2 * S * D * I + 4 * S * S * D
        embed = 2 * self.num_patches * self.patch_k * D
        T, W = self.out_tokens, self.shuffled_width
        mlp1 = 2 * T * W * C + 2 * T * C * C
        return float(L * per_layer + embed + mlp1)

    def to_dict(self):
        return asdict(self)


def internvit_300m(llm_hidden: int = 2048, **kw) -> VisionPathConfig:
    """InternViT-300M as in the reference's config.json:117-200 (vision_config)."""
    return VisionPathConfig(llm_hidden=llm_hidden, **kw)


def internvit_6b(llm_hidden: int = 6144, num_layers: int = 48, **kw) -> VisionPathConfig:
    """InternViT-6B: the defaults of InternVisionConfig (configuration_intern_vit.py:63-83)."""
    return VisionPathConfig(hidden_size=3200, num_heads=25, intermediate_size=12800, num_layers=num_layers,
                            llm_hidden=llm_hidden, qkv_bias=False, qk_normalization=True, norm_type="rms_norm",
                            initializer_factor=0.1, **kw)


def tiny(llm_hidden: int = 256, **kw) -> VisionPathConfig:
    """A 2-layer, 2-head (d=64) tower with the real 448/14 geometry: small enough to ship golden tensors."""
    base = dict(hidden_size=128, num_heads=2, intermediate_size=512, num_layers=2, llm_hidden=llm_hidden)
    base.update(kw)
    return VisionPathConfig(**base)
