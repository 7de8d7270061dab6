"""Fused tcgen05 attention vs the oracle's naive softmax(q*scale @ k^T) @ v, through aigv_attention (C-ABI)."""
import numpy as np
import pytest
import torch

from aigv_assessor_b200 import _lib
from tests.gpu_util import dev, emax, ptr, stream

pytestmark = pytest.mark.gpu


def _attn_ref(qkv, frames, S, H, d):
    q, k, v = qkv.float().cuda().view(frames, S, 3, H, d).permute(2, 0, 3, 1, 4)
    a = (q * d ** -0.5) @ k.transpose(-2, -1)
    return (a.softmax(-1) @ v).transpose(1, 2).reshape(frames * S, H * d)


@pytest.mark.parametrize("frames,S,H,d", [(1, 128, 1, 64), (1, 256, 2, 64), (2, 1025, 2, 64), (3, 130, 4, 64),
                                          (1, 1025, 16, 64), (2, 95, 1, 64), (3, 257, 2, 64), (2, 513, 3, 64), (2, 257, 2, 128),
                                          (1, 128, 1, 128), (2, 1025, 2, 128), (1, 1025, 25, 128), (2, 95, 3, 128)])
def test_attention_matches_reference(frames, S, H, d):
    g = torch.Generator().manual_seed(S + H + d)
    qkv = (torch.randn((frames * S, 3 * H * d), generator=g) * (1.5 if d == 64 else 1.2)).to(torch.bfloat16)
    out = torch.full((frames * S, H * d), float("nan"), dtype=torch.bfloat16, device="cuda")
    lib = _lib.load()
    dq = dev(qkv)
    st = lib.aigv_attention(ptr(dq), ptr(out), frames, S, H, d, d ** -0.5, stream())
    assert st == 0
    torch.cuda.synchronize()
    ref = _attn_ref(qkv, frames, S, H, d)
    y = out.float().cpu().numpy()
    assert np.isfinite(y).all()
    assert emax(y, ref.cpu().numpy()) < 1e-2


def test_attention_peaked_rows_are_stable():
    # one dominant key per row: exercises the running-max rescale across KV blocks
    frames, S, H, d = 1, 1025, 1, 64
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn((S, 3 * d), generator=g)
    qkv[:, :2 * d] *= 6.0
    qkv = qkv.to(torch.bfloat16)
    out = torch.empty((S, d), dtype=torch.bfloat16, device="cuda")
    lib = _lib.load()
    dq = dev(qkv)
    assert lib.aigv_attention(ptr(dq), ptr(out), frames, S, H, d, d ** -0.5, stream()) == 0
    torch.cuda.synchronize()
    ref = _attn_ref(qkv, frames, S, H, d)
    assert emax(out.float().cpu().numpy(), ref.cpu().numpy()) < 1e-2


def test_attention_monotone_scores_force_a_rescale_every_block():
    # scores grow with the key index, so the running maximum overtakes the stale reference again and again
    frames, S, H, d = 1, 1025, 2, 64
    g = torch.Generator().manual_seed(11)
    q = torch.randn((S, H, d), generator=g) * 0.1 + 1.0
    k = torch.randn((S, H, d), generator=g) * 0.1 + (torch.arange(S).float() / S * 6.0).view(S, 1, 1)
    v = torch.randn((S, H, d), generator=g)
    qkv = torch.stack([q, k, v], dim=1).reshape(S, 3 * H * d).to(torch.bfloat16)
    out = torch.empty((S, H * d), dtype=torch.bfloat16, device="cuda")
    lib = _lib.load()
    dq = dev(qkv)
    assert lib.aigv_attention(ptr(dq), ptr(out), frames, S, H, d, d ** -0.5, stream()) == 0
    torch.cuda.synchronize()
    ref = _attn_ref(qkv, frames, S, H, d)
    assert emax(out.float().cpu().numpy(), ref.cpu().numpy()) < 1e-2


def test_attention_rejects_what_it_does_not_cover():
    lib = _lib.load()
    x = torch.zeros((128, 3 * 96), dtype=torch.bfloat16, device="cuda")
    o = torch.zeros((128, 96), dtype=torch.bfloat16, device="cuda")
    assert lib.aigv_attention(ptr(x), ptr(o), 1, 128, 1, 96, 0.1, stream()) == -5   # head_dim 96
    x = torch.zeros((100, 3 * 64), dtype=torch.bfloat16, device="cuda")
    o = torch.zeros((100, 64), dtype=torch.bfloat16, device="cuda")
    assert lib.aigv_attention(ptr(x), ptr(o), 1, 100, 1, 64, 0.1, stream()) == -5   # 100 % 64 = 36 > 32
