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
                                          (1, 128, 1, 128), (2, 1025, 2, 128), (1, 1025, 25, 128), (2, 95, 3, 128),
                                          # round-2 kernel: any tail, dead second tiles, several items per persistent CTA
                                          (2, 100, 2, 64), (1, 300, 2, 64), (1, 640, 1, 64), (1, 1, 2, 64), (2, 129, 1, 64),
                                          (3, 384, 2, 128), (4, 1025, 16, 64), (5, 513, 16, 64), (2, 1025, 25, 128),
                                          (37, 200, 5, 64),
                                          # several items per CTA where a single-live-tile item is followed by another one
                                          # (128-wide heads have ONE Q buffer: the producer's early K/V load deadlocked there)
                                          (4, 300, 25, 128), (3, 2108, 32, 128), (8, 300, 16, 64)])
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
    assert lib.aigv_attention(ptr(x), ptr(o), 0, 128, 1, 64, 0.1, stream()) == -1   # no frames


def test_attention_batch_is_bit_identical_to_single_frames():
    # the persistent kernel walks many (frame, head, chunk) items per CTA: a frame's result must not depend on which
    # CTA / pass / S-buffer rotation it lands on
    frames, S, H, d = 24, 1025, 16, 64
    g = torch.Generator().manual_seed(3)
    qkv = (torch.randn((frames * S, 3 * H * d), generator=g) * 1.5).to(torch.bfloat16).cuda()
    out = torch.empty((frames * S, H * d), dtype=torch.bfloat16, device="cuda")
    lib = _lib.load()
    assert lib.aigv_attention(ptr(qkv), ptr(out), frames, S, H, d, d ** -0.5, stream()) == 0
    for f in (0, 11, 23):
        one = torch.empty((S, H * d), dtype=torch.bfloat16, device="cuda")
        sub = qkv[f * S:(f + 1) * S].contiguous()
        assert lib.aigv_attention(ptr(sub), ptr(one), 1, S, H, d, d ** -0.5, stream()) == 0
        torch.cuda.synchronize()
        assert torch.equal(one, out[f * S:(f + 1) * S])
    ref = _attn_ref(qkv[:2 * S].cpu(), 2, S, H, d)
    assert emax(out[:2 * S].float().cpu().numpy(), ref.cpu().numpy()) < 1e-2


def test_exp2_on_the_fma_pipe_matches_exp2f():
    # the share of the softmax exponentials evaluated as a polynomial: |rel err| <= 1e-4 (bf16 P rounds at 2e-3)
    lib = _lib.load()
    x = torch.cat([torch.linspace(-120.0, 8.0, 1 << 20), torch.tensor([0.0, -0.5, 0.5, 8.0, -119.99, -1e-8]),
                   -torch.rand(4096) * 30]).float().cuda()
    y = torch.empty_like(x)
    assert lib.aigv_exp2_poly(ptr(x), ptr(y), x.numel(), stream()) == 0
    torch.cuda.synchronize()
    ref = torch.exp2(x.double())
    rel = ((y.double() - ref) / ref).abs().max().item()
    assert rel < 1e-4, rel
    # far below the clamp the result is ~2^-126 (never negative, never large)
    z = torch.tensor([-200.0, -1e4, -3e38, float("-inf")], device="cuda")
    w = torch.empty_like(z)
    assert lib.aigv_exp2_poly(ptr(z), ptr(w), 4, stream()) == 0
    torch.cuda.synchronize()
    assert (w >= 0).all() and (w < 1e-37).all()
