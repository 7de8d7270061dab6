"""Motion branch: the SlowFast-R50 trunk behind ``InternVLChatModel.slowfast_model``
(/root/reference/internvl/model/internvl_chat_eval2/modeling_internvl_chat.py:136-193, fed by pack_pathway_output :95-133 from
the same pixel_values the ViT sees, :337-343) as a schedule of this library's kernels: every convolution is a GEMM on
channels-last activations (1x1x1: directly; others: after an ``aigv_im2col3d`` gather) with eval-mode BatchNorm folded into
the weights and ReLU in the epilogue; the fast-to-slow fusion's concat is a column offset of the GEMM output; the tail
(repeat_interleave + AvgPool3d + AdaptiveAvgPool3d) is one weighted reduction.

The network itself is pytorchvideo's ``slowfast_r50`` -- a third-party dependency the reference downloads at construction and
that is absent here -- so the layer graph below restates its published architecture (see oracle/slowfast_oracle.py for the
statement and the "parity unpinned" caveat); parameter names follow pytorchvideo's module tree under the reference's
wrapper (``slowfast_model.feature_extraction.<block>.``), which is what ``model.state_dict()`` would hold.
No CPU path; inference only.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping

import numpy as np
import torch

from . import _lib

PREFIX = "slowfast_model.feature_extraction."
DEPTHS = (3, 4, 6, 3)
SLOW = dict(inner=(64, 128, 256, 512), out=(256, 512, 1024, 2048), kt=(1, 1, 3, 3), stem=(64, 1))
FAST = dict(inner=(8, 16, 32, 64), out=(32, 64, 128, 256), kt=(3, 3, 3, 3), stem=(8, 5))
SPATIAL_STRIDE = (1, 2, 2, 2)
BN_EPS = 1e-5
FEATURES = SLOW["out"][3] + FAST["out"][3]   # 2304: what motion_mlp's LayerNorm expects (eval2 :244)


def layer_table():
    """[(conv weight name, norm name, (cout, cin, kt, kh, kw))] of the whole trunk, in execution order per pathway."""
    t = []
    b0 = PREFIX + "0."
    t.append((b0 + "multipathway_blocks.0.conv.weight", b0 + "multipathway_blocks.0.norm", (64, 3, 1, 7, 7)))
    t.append((b0 + "multipathway_blocks.1.conv.weight", b0 + "multipathway_blocks.1.norm", (8, 3, 5, 7, 7)))
    fast_dims = (FAST["stem"][0],) + FAST["out"][:3]
    for i in range(4):
        fu = f"{PREFIX}{i}.multipathway_fusion."
        t.append((fu + "conv_fast_to_slow.weight", fu + "norm", (2 * fast_dims[i], fast_dims[i], 7, 1, 1)))
    slow_in = tuple(c + 2 * f for c, f in zip((SLOW["stem"][0],) + SLOW["out"][:3], fast_dims))
    for s in range(4):
        for p, (P, din) in enumerate(((SLOW, slow_in[s]), (FAST, fast_dims[s]))):
            for b in range(DEPTHS[s]):
                rb = f"{PREFIX}{s + 1}.multipathway_blocks.{p}.res_blocks.{b}."
                cin = din if b == 0 else P["out"][s]
                if b == 0:
                    t.append((rb + "branch1_conv.weight", rb + "branch1_norm", (P["out"][s], cin, 1, 1, 1)))
                t.append((rb + "branch2.conv_a.weight", rb + "branch2.norm_a", (P["inner"][s], cin, P["kt"][s], 1, 1)))
                t.append((rb + "branch2.conv_b.weight", rb + "branch2.norm_b", (P["inner"][s], P["inner"][s], 1, 3, 3)))
                t.append((rb + "branch2.conv_c.weight", rb + "branch2.norm_c", (P["out"][s], P["inner"][s], 1, 1, 1)))
    return t


def _sp(device) -> C.c_void_p:
    return C.c_void_p(int(torch.cuda.current_stream(device).cuda_stream))


def _p(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


class _Conv:
    """One Conv3d + folded BatchNorm: weight bf16 [cout, kpad] with K ordered (dt, dh, dw, cin), bias fp32 [cout]."""

    def __init__(self, sd, wname, nname, shape, device):
        w = torch.as_tensor(np.asarray(sd[wname], dtype=np.float32) if not torch.is_tensor(sd[wname]) else sd[wname]).float()
        if tuple(w.shape) != tuple(shape):
            raise ValueError(f"{wname} has shape {tuple(w.shape)}, expected {tuple(shape)}")
        g = [torch.as_tensor(sd[nname + s]).float() for s in (".weight", ".bias", ".running_mean", ".running_var")]
        scale = g[0] / torch.sqrt(g[3] + BN_EPS)
        w = w * scale.view(-1, 1, 1, 1, 1)
        self.cout, self.cin, self.kt, self.kh, self.kw = shape
        self.K = self.cin * self.kt * self.kh * self.kw
        self.kpad = (self.K + 7) // 8 * 8
        self.npad = (self.cout + 31) // 32 * 32     # zero rows: narrow layers still run as a 32-wide tensor-core GEMM
        wk = torch.zeros((self.npad, self.kpad), dtype=torch.float32)
        wk[:self.cout, :self.K] = w.permute(0, 2, 3, 4, 1).reshape(self.cout, self.K)
        self.w = wk.to(device).to(torch.bfloat16).contiguous()
        bias = torch.zeros((self.npad,), dtype=torch.float32)
        bias[:self.cout] = g[1] - g[2] * scale
        self.b = bias.to(device).contiguous()


class _Act:
    """Channels-last activation [n, T, H, W, C] living in columns [col, col + C) of a row-major matrix with row stride ld."""

    def __init__(self, buf, n, T, H, W, C, ld=None, col=0):
        self.buf, self.n, self.T, self.H, self.W, self.C = buf, n, T, H, W, C
        self.ld = C if ld is None else ld
        self.col = col

    @property
    def rows(self):
        return self.n * self.T * self.H * self.W

    def ptr(self):
        return C.c_void_p(self.buf.data_ptr() + 2 * self.col)


class SlowFastTrunk:
    """use_graph: the ~215 launches of a call are captured once per input shape into a CUDA graph and replayed (the Python
    loop that issues them costs more host time than the GPU needs to run them)."""

    def __init__(self, sd: Mapping[str, object], device=None, use_graph: bool = True):
        if not torch.cuda.is_available():
            raise _lib.AigvError("SlowFastTrunk needs a CUDA device (no CPU path)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.load()
        self.conv = {w: _Conv(sd, w, nm, shape, self.device) for (w, nm, shape) in layer_table()}
        self.launches = 0
        self.use_graph = bool(use_graph)
        self._graphs = {}

    # ------------------------------------------------------------------ ops
    def _check(self, st, what):
        _lib.check(st, None, what)
        self.launches += 1

    def _new(self, rows, cols):
        return torch.empty((rows, cols), dtype=torch.bfloat16, device=self.device)

    def _conv(self, x: _Act, name: str, stride=(1, 1, 1), pad=(0, 0, 0), relu=True, out: _Act | None = None,
              strides_in=None) -> _Act:
        """y = act(conv(x) folded-BN); x channels-last (or, with strides_in = (s_n, s_t, s_h, s_w, s_c), any layout)."""
        cv = self.conv[name]
        if cv.cin != x.C:
            raise ValueError(f"{name}: {x.C} input channels, expected {cv.cin}")
        To = (x.T + 2 * pad[0] - cv.kt) // stride[0] + 1
        Ho = (x.H + 2 * pad[1] - cv.kh) // stride[1] + 1
        Wo = (x.W + 2 * pad[2] - cv.kw) // stride[2] + 1
        rows = x.n * To * Ho * Wo
        st = _sp(self.device)
        pointwise = (cv.kt, cv.kh, cv.kw) == (1, 1, 1) and stride == (1, 1, 1) and strides_in is None
        if pointwise:
            a_ptr, lda = x.ptr(), x.ld
        else:
            patches = self._new(rows, cv.kpad)
            if strides_in is None:
                s_w = x.ld
                strides_in = (x.T * x.H * x.W * s_w, x.H * x.W * s_w, x.W * s_w, s_w, 1)
            self._check(self.lib.aigv_im2col3d(x.ptr(), _p(patches), x.n, x.T, x.H, x.W, x.C, *strides_in, cv.kt, cv.kh, cv.kw,
                                               *stride, *pad, cv.kpad, st), f"im2col3d({name})")
            a_ptr, lda = _p(patches), cv.kpad
        if out is None:
            out = _Act(self._new(rows, cv.cout), x.n, To, Ho, Wo, cv.cout)
        self._check(self.lib.aigv_gemm_ex(a_ptr, lda, _p(cv.w), _p(cv.b), out.ptr(), out.ld, rows, cv.cout, cv.npad, cv.kpad, int(relu), st),
                    f"gemm({name})")
        return out

    def _res_block(self, x: _Act, rb: str, first: bool, stride: int, kt: int, out: _Act | None) -> _Act:
        y = self._conv(x, rb + "branch2.conv_a.weight", pad=(kt // 2, 0, 0))
        y = self._conv(y, rb + "branch2.conv_b.weight", stride=(1, stride, stride), pad=(0, 1, 1))
        y = self._conv(y, rb + "branch2.conv_c.weight", relu=False)
        sc = self._conv(x, rb + "branch1_conv.weight", stride=(1, stride, stride), relu=False) if first else x
        if sc.ld != sc.C or sc.col != 0:
            raise ValueError("internal: shortcut must be contiguous")
        if out is None:
            out = _Act(self._new(y.rows, y.C), y.n, y.T, y.H, y.W, y.C)
        self._check(self.lib.aigv_add_relu(sc.ptr(), y.ptr(), out.ptr(), y.rows, y.C, out.ld, _sp(self.device)), "add_relu")
        return out

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor, n_clips: int) -> torch.Tensor:
        """pixel_values bf16 [B*T, 3, H, W] (the ViT's input, eval2 :337) -> motion features fp32 [B, 2304]."""
        if pixel_values.dim() != 4 or pixel_values.shape[1] != 3 or pixel_values.device != self.device:
            raise ValueError(f"pixel_values must be [B*T, 3, H, W] on {self.device}")
        bt, _, H, W = pixel_values.shape
        if n_clips <= 0 or bt % n_clips:
            raise ValueError("pixel_values do not split into n_clips clips")
        T = bt // n_clips
        if T < 4 or H % 32 or W % 32 or H // 32 < 7 or W // 32 < 7:
            raise ValueError("the trunk needs >= 4 frames per clip and frames of at least 224 x 224 (multiples of 32)")
        pv = pixel_values.detach().to(torch.bfloat16).contiguous()
        with torch.cuda.device(self.device):
            if not self.use_graph:
                return self._forward(pv, n_clips, T, H, W)
            key = (n_clips, T, H, W)
            if key not in self._graphs:
                static_in = pv.clone()
                self._forward(static_in, n_clips, T, H, W)      # warm-up outside the capture (one-time kernel attributes)
                torch.cuda.current_stream(self.device).synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    static_out = self._forward(static_in, n_clips, T, H, W)
                self._graphs[key] = (graph, static_in, static_out)
            graph, static_in, static_out = self._graphs[key]
            static_in.copy_(pv)
            graph.replay()
            return static_out.clone()

    __call__ = forward

    def _forward(self, pv, B, T, H, W):
        st = _sp(self.device)
        # pack_pathway_output (eval2 :95-133): fast = every frame, slow = the linspace pick of T // 4 frames
        idx = torch.linspace(0, T - 1, T // 4, device=self.device).long()
        slow_frames = pv.view(B, T, 3, H, W).index_select(1, idx).contiguous()           # [B, T/4, 3, H, W]
        Ts = T // 4
        nchw = lambda t_: (t_ * 3 * H * W, 3 * H * W, W, 1, H * W)                       # (s_n, s_t, s_h, s_w, s_c) of [B, t_, 3, H, W]
        b0 = PREFIX + "0."
        fast_dims = (FAST["stem"][0],) + FAST["out"][:3]
        slow_dims = (SLOW["stem"][0],) + SLOW["out"][:3]
        # stems: conv + BN + ReLU, then MaxPool3d((1,3,3), (1,2,2), (0,1,1)); the slow result lands in the left columns of the
        # concat buffer the fusion fills on the right
        xs = self._conv(_Act(slow_frames, B, Ts, H, W, 3), b0 + "multipathway_blocks.0.conv.weight", stride=(1, 2, 2), pad=(0, 3, 3),
                        strides_in=nchw(Ts))
        xf = self._conv(_Act(pv, B, T, H, W, 3), b0 + "multipathway_blocks.1.conv.weight", stride=(1, 2, 2), pad=(2, 3, 3),
                        strides_in=nchw(T))
        Hp, Wp = (xs.H + 2 - 3) // 2 + 1, (xs.W + 2 - 3) // 2 + 1
        cat_w = slow_dims[0] + 2 * fast_dims[0]
        cat = _Act(self._new(B * Ts * Hp * Wp, cat_w), B, Ts, Hp, Wp, slow_dims[0], ld=cat_w, col=0)
        self._check(self.lib.aigv_maxpool3d(xs.ptr(), cat.ptr(), B * Ts, xs.H, xs.W, xs.C, 3, 3, 2, 2, 1, 1, cat_w, st), "maxpool(slow)")
        fpool = _Act(self._new(B * T * Hp * Wp, xf.C), B, T, Hp, Wp, xf.C)
        self._check(self.lib.aigv_maxpool3d(xf.ptr(), fpool.ptr(), B * T, xf.H, xf.W, xf.C, 3, 3, 2, 2, 1, 1, xf.C, st), "maxpool(fast)")
        xs, xf = cat, fpool
        for i in range(5):
            if i > 0:
                s = i - 1
                for b in range(DEPTHS[s]):
                    stride = SPATIAL_STRIDE[s] if b == 0 else 1
                    last = b == DEPTHS[s] - 1
                    out_s = None
                    if last and i < 4:   # the stage's output goes straight into the next concat buffer
                        Ho, Wo = (xs.H if b else (xs.H + 2 - 3) // stride + 1), (xs.W if b else (xs.W + 2 - 3) // stride + 1)
                        cw = SLOW["out"][s] + 2 * FAST["out"][s]
                        nxt = self._new(B * Ts * Ho * Wo, cw)
                        out_s = _Act(nxt, B, Ts, Ho, Wo, SLOW["out"][s], ld=cw, col=0)
                    xs = self._res_block(xs, f"{PREFIX}{i}.multipathway_blocks.0.res_blocks.{b}.", b == 0, stride, SLOW["kt"][s], out_s)
                    xf = self._res_block(xf, f"{PREFIX}{i}.multipathway_blocks.1.res_blocks.{b}.", b == 0, stride, FAST["kt"][s], None)
            if i < 4:
                # FuseFastToSlow: conv (7,1,1) stride (4,1,1) pad (3,0,0) + BN + ReLU on the fast pathway, written next to the
                # slow pathway's channels (xs occupies the left xs.C columns of a buffer that is xs.ld wide)
                cs, total = xs.C, xs.ld
                if total != cs + 2 * xf.C:
                    raise ValueError("internal: concat buffer width")
                dst = _Act(xs.buf, B, Ts, xs.H, xs.W, 2 * xf.C, ld=total, col=cs)
                self._conv(xf, f"{PREFIX}{i}.multipathway_fusion.conv_fast_to_slow.weight", stride=(4, 1, 1), pad=(3, 0, 0), out=dst)
                xs = _Act(xs.buf, B, Ts, xs.H, xs.W, total)
        # tail (eval2 :183-191)
        feat = torch.empty((B, FEATURES), dtype=torch.float32, device=self.device)
        self._check(self.lib.aigv_slowfast_tail(xs.ptr(), _p(feat), B, xs.T, xs.H, xs.W, xs.C, xs.ld, 4, 8, 7, 7, FEATURES, 0, st), "tail(slow)")
        self._check(self.lib.aigv_slowfast_tail(xf.ptr(), _p(feat), B, xf.T, xf.H, xf.W, xf.C, xf.ld, 4, 32, 7, 7, FEATURES, SLOW["out"][3], st),
                    "tail(fast)")
        return feat
