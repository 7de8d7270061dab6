"""CPU restatement (torch fp32) of the motion branch's SlowFast-R50 trunk. TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED. The reference does not contain this network: it calls ``pytorchvideo.models.hub.slowfast_r50(pretrained=True)``
(/root/reference/internvl/model/internvl_chat_eval2/modeling_internvl_chat.py:23,136-193), an un-vendored, unpinned
third-party dependency (pytorchvideo; not in requirements.txt, not in this image, weights from a network download), keeps
its first five blocks + the two average pools + the adaptive pool, and feeds them ``pack_pathway_output`` (:95-133). What is
restated here is pytorchvideo's PUBLISHED architecture for ``create_slowfast`` with the slowfast_r50 defaults
(Feichtenhofer et al., "SlowFast Networks for Video Recognition", ICCV 2019, R50 8x8 instantiation):

  stem      slow Conv3d(3, 64, (1,7,7), stride (1,2,2), pad (0,3,3)), fast Conv3d(3, 8, (5,7,7), (1,2,2), (2,3,3)); BN; ReLU;
            MaxPool3d((1,3,3), (1,2,2), (0,1,1))
  fusion    after the stem and after res2 / res3 / res4: Conv3d(Cf, 2 Cf, (7,1,1), stride (4,1,1), pad (3,0,0)) + BN + ReLU on
            the fast pathway, concatenated to the slow pathway's channels
  res2-5    bottleneck blocks (3, 4, 6, 3): conv_a (kt,1,1) [slow kt = 1,1,3,3; fast kt = 3] + BN + ReLU, conv_b (1,3,3) with the
            stage's spatial stride (1,2,2,2) in its first block + BN + ReLU, conv_c 1x1x1 + BN, shortcut conv 1x1x1 (strided) +
            BN in the first block of a stage, ReLU after the sum; widths slow 64..512 -> 256..2048, fast 8..64 -> 32..256
  tail      eval2 :183-189: repeat_interleave(4) along t, AvgPool3d((8,7,7)) / AvgPool3d((32,7,7)) (stride 1),
            AdaptiveAvgPool3d(1), concat -> [B, 2304]

State-dict names follow pytorchvideo's module tree under the reference's wrapper (``slowfast_model.feature_extraction.<i>.``).
Anchors that ARE the reference's: the call shape (pack_pathway_output's linspace pick of T//4 slow frames, the tail, the
2304-wide output that motion_mlp's LayerNorm(2304) consumes, eval2 :244-249). The CUDA path (aigv_assessor_b200/slowfast.py)
is tested against THIS restatement, so a mistake in the restated architecture would not be caught: hence "unpinned". One
Two weak anchors exist (tests/test_host_cpu.py): with the 400-class head the reference drops, the restated network has
34 566 488 learnable parameters = the 34.57 M pytorchvideo's model zoo lists for slowfast_r50, and 65.71 G multiply-accumulates
for the Kinetics 8x8 setting (8 / 32 frames, 256 x 256) = the published 65.71 GFLOPs: widths, kernel sizes, strides and the
fusion are right; the order of BN / ReLU and the paddings are not covered by them.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

PREFIX = "slowfast_model.feature_extraction."
DEPTHS = (3, 4, 6, 3)
SLOW_INNER, FAST_INNER = (64, 128, 256, 512), (8, 16, 32, 64)
SLOW_OUT, FAST_OUT = (256, 512, 1024, 2048), (32, 64, 128, 256)
SLOW_KT, FAST_KT = (1, 1, 3, 3), (3, 3, 3, 3)
SPATIAL_STRIDE = (1, 2, 2, 2)
BN_EPS = 1e-5


def weight_spec():
    """(name, shape, kind) for every tensor of the trunk; kind in {'conv', 'bn_w', 'bn_b', 'bn_mean', 'bn_var'}."""
    def bn(prefix, c):
        yield prefix + ".weight", (c,), "bn_w"
        yield prefix + ".bias", (c,), "bn_b"
        yield prefix + ".running_mean", (c,), "bn_mean"
        yield prefix + ".running_var", (c,), "bn_var"

    b0 = PREFIX + "0."
    yield b0 + "multipathway_blocks.0.conv.weight", (64, 3, 1, 7, 7), "conv"
    yield from bn(b0 + "multipathway_blocks.0.norm", 64)
    yield b0 + "multipathway_blocks.1.conv.weight", (8, 3, 5, 7, 7), "conv"
    yield from bn(b0 + "multipathway_blocks.1.norm", 8)
    fast_dims = (8,) + FAST_OUT[:3]
    for i in range(4):   # fusion after the stem and after res2..res4
        bi = f"{PREFIX}{i}.multipathway_fusion."
        yield bi + "conv_fast_to_slow.weight", (2 * fast_dims[i], fast_dims[i], 7, 1, 1), "conv"
        yield from bn(bi + "norm", 2 * fast_dims[i])
    slow_in = (64 + 16, 256 + 64, 512 + 128, 1024 + 256)
    fast_in = (8,) + FAST_OUT[:3]
    for s in range(4):
        for p, (din, inner, dout, kt) in enumerate(((slow_in[s], SLOW_INNER[s], SLOW_OUT[s], SLOW_KT[s]),
                                                    (fast_in[s], FAST_INNER[s], FAST_OUT[s], FAST_KT[s]))):
            for b in range(DEPTHS[s]):
                rb = f"{PREFIX}{s + 1}.multipathway_blocks.{p}.res_blocks.{b}."
                cin = din if b == 0 else dout
                if b == 0:
                    yield rb + "branch1_conv.weight", (dout, cin, 1, 1, 1), "conv"
                    yield from bn(rb + "branch1_norm", dout)
                yield rb + "branch2.conv_a.weight", (inner, cin, kt, 1, 1), "conv"
                yield from bn(rb + "branch2.norm_a", inner)
                yield rb + "branch2.conv_b.weight", (inner, inner, 1, 3, 3), "conv"
                yield from bn(rb + "branch2.norm_b", inner)
                yield rb + "branch2.conv_c.weight", (dout, inner, 1, 1, 1), "conv"
                yield from bn(rb + "branch2.norm_c", dout)


def pack_pathways(pixel_values: torch.Tensor, n_clips: int):
    """eval2 :337-343 + :95-133: [B*T, 3, H, W] -> (slow [B, 3, T//4, H, W], fast [B, 3, T, H, W])."""
    bt, c, h, w = pixel_values.shape
    frames = pixel_values.view(n_clips, bt // n_clips, c, h, w).permute(0, 2, 1, 3, 4)
    idx = torch.linspace(0, frames.shape[2] - 1, frames.shape[2] // 4).long()
    return frames.index_select(2, idx), frames


def _bn(x, sd, name):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"], sd[name + ".bias"],
                        training=False, eps=BN_EPS)


def _res_block(x, sd, rb, first, stride, kt):
    y = F.relu(_bn(F.conv3d(x, sd[rb + "branch2.conv_a.weight"], padding=(kt // 2, 0, 0)), sd, rb + "branch2.norm_a"))
    y = F.relu(_bn(F.conv3d(y, sd[rb + "branch2.conv_b.weight"], stride=(1, stride, stride), padding=(0, 1, 1)), sd, rb + "branch2.norm_b"))
    y = _bn(F.conv3d(y, sd[rb + "branch2.conv_c.weight"]), sd, rb + "branch2.norm_c")
    if first:
        x = _bn(F.conv3d(x, sd[rb + "branch1_conv.weight"], stride=(1, stride, stride)), sd, rb + "branch1_norm")
    return F.relu(x + y)


@torch.no_grad()
def trunk(slow: torch.Tensor, fast: torch.Tensor, sd) -> tuple[torch.Tensor, torch.Tensor]:
    """feature_extraction (blocks 0-4): -> (slow [B,2048,T/4,H/32,W/32], fast [B,256,T,H/32,W/32])"""
    b0 = PREFIX + "0."
    xs = F.relu(_bn(F.conv3d(slow, sd[b0 + "multipathway_blocks.0.conv.weight"], stride=(1, 2, 2), padding=(0, 3, 3)), sd,
                    b0 + "multipathway_blocks.0.norm"))
    xf = F.relu(_bn(F.conv3d(fast, sd[b0 + "multipathway_blocks.1.conv.weight"], stride=(1, 2, 2), padding=(2, 3, 3)), sd,
                    b0 + "multipathway_blocks.1.norm"))
    xs = F.max_pool3d(xs, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    xf = F.max_pool3d(xf, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    for i in range(5):
        if i > 0:
            s = i - 1
            for b in range(DEPTHS[s]):
                st = SPATIAL_STRIDE[s] if b == 0 else 1
                xs = _res_block(xs, sd, f"{PREFIX}{i}.multipathway_blocks.0.res_blocks.{b}.", b == 0, st, SLOW_KT[s])
                xf = _res_block(xf, sd, f"{PREFIX}{i}.multipathway_blocks.1.res_blocks.{b}.", b == 0, st, FAST_KT[s])
        if i < 4:
            fu = f"{PREFIX}{i}.multipathway_fusion."
            f2s = F.relu(_bn(F.conv3d(xf, sd[fu + "conv_fast_to_slow.weight"], stride=(4, 1, 1), padding=(3, 0, 0)), sd, fu + "norm"))
            xs = torch.cat([xs, f2s], dim=1)
    return xs, xf


@torch.no_grad()
def motion_features(pixel_values: torch.Tensor, n_clips: int, sd) -> torch.Tensor:
    """What eval2 :337-344 hands to motion_mlp: [B*T, 3, H, W] -> [B, 2304]."""
    slow, fast = pack_pathways(pixel_values.float(), n_clips)
    xs, xf = trunk(slow, fast, sd)
    xs = xs.repeat_interleave(4, dim=2)
    xf = xf.repeat_interleave(4, dim=2)
    xs = F.avg_pool3d(xs, (8, 7, 7), stride=(1, 1, 1))
    xf = F.avg_pool3d(xf, (32, 7, 7), stride=(1, 1, 1))
    xs = F.adaptive_avg_pool3d(xs, 1)
    xf = F.adaptive_avg_pool3d(xf, 1)
    return torch.cat([xs, xf], dim=1).flatten(1)
