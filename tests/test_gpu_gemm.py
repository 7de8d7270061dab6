"""tcgen05 GEMM + fused epilogue vs fp32 references, through aigv_gemm_epilogue (C-ABI)."""
import numpy as np
import pytest
import torch

from tests.gpu_util import dev, emax, gemm

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16)


def _ref(a, w, bias=None, colscale=None, residual=None, gelu=False):
    y = a.float().cuda() @ w.float().cuda().T
    if bias is not None:
        y = y + bias.cuda()
    if gelu:
        y = torch.nn.functional.gelu(y)
    if colscale is not None:
        y = y * colscale.cuda()
    if residual is not None:
        y = y + residual.cuda()
    return y


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 256, 128), (1025, 384, 128), (8200, 1024, 1024),
                                   (2050, 3072, 1024), (1024, 128, 640), (512, 2048, 4096), (77, 96, 72)])
def test_gemm_bias_bf16(M, N, K):
    torch.backends.cuda.matmul.allow_tf32 = False
    a, w = _rand((M, K), 1), _rand((N, K), 2, K ** -0.5)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(3))
    da, dw, db = dev(a), dev(w), dev(bias)
    y = gemm(da, dw, bias=db)
    ref = _ref(a, w, bias)
    # bf16 output rounding: 2^-9 relative of the value; gate on normalised max error
    assert emax(y.float().cpu().numpy(), ref.cpu().numpy()) < 6e-3


@pytest.mark.parametrize("M,N,K", [(2108, 6400, 4096), (5400, 6656, 4096), (8432, 6400, 4096)])
def test_gemm_tile_orders_with_a_large_weight(M, N, K):
    """W beyond 48 MB and A smaller than W selects the M-fastest tile order (<= 20 row tiles) or M fastest inside groups of
    16 row tiles (more; the last group is ragged: 22 = 16 + 6, 33 = 16 + 16 + 1): every output tile must still be written once."""
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(M + N)
    a = torch.randn((M, K), device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn((N, K), device="cuda", generator=g) * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    y = gemm(a, w, bias=bias)
    ref = a.float() @ w.float().t() + bias
    err = float((y.float() - ref).abs().max() / ref.abs().max())
    assert err < 6e-3


def test_gemm_gelu_bf16():
    M, N, K = 2050, 4096, 1024
    a, w = _rand((M, K), 4), _rand((N, K), 5, K ** -0.5)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(6))
    da, dw, db = dev(a), dev(w), dev(bias)
    y = gemm(da, dw, bias=db, gelu=True)
    ref = _ref(a, w, bias, gelu=True)
    assert emax(y.float().cpu().numpy(), ref.cpu().numpy()) < 6e-3


def test_gemm_scale_residual_f32_inplace():
    M, N, K = 2050, 1024, 4096
    a, w = _rand((M, K), 7), _rand((N, K), 8, K ** -0.5)
    g = torch.Generator().manual_seed(9)
    bias, ls, res = torch.randn(N, generator=g), 0.1 * torch.randn(N, generator=g), torch.randn((M, N), generator=g)
    x = dev(res.clone())
    da, dw, db, dl = dev(a), dev(w), dev(bias), dev(ls)
    gemm(da, dw, bias=db, colscale=dl, residual=x, out=x)  # in place, like the residual stream
    ref = _ref(a, w, bias, ls, res)
    assert emax(x.cpu().numpy(), ref.cpu().numpy()) < 1e-4


def test_gemm_f32_exact_small_integers():
    # integer-valued bf16 inputs with small K: fp32 accumulation is exact -> bit-exact against the fp32 reference
    M, N, K = 256, 512, 256
    g = torch.Generator().manual_seed(10)
    a = torch.randint(-4, 5, (M, K), generator=g).to(torch.bfloat16)
    w = torch.randint(-4, 5, (N, K), generator=g).to(torch.bfloat16)
    da, dw = dev(a), dev(w)
    y = gemm(da, dw, out_dtype=torch.float32)
    ref = a.double() @ w.double().T
    assert np.array_equal(y.cpu().numpy().astype(np.float64), ref.numpy())


def test_gemm_rejects_bad_shapes():
    from aigv_assessor_b200 import _lib
    lib = _lib.load()
    a, w = dev(_rand((64, 64), 1)), dev(_rand((40, 64), 2))
    out = torch.empty((64, 40), dtype=torch.bfloat16, device="cuda")
    from tests.gpu_util import ptr, stream
    assert lib.aigv_gemm_epilogue(ptr(a), ptr(w), None, None, None, None, ptr(out), 64, 40, 64, 0, stream()) == -1
    assert lib.aigv_gemm_epilogue(None, ptr(w), None, None, None, None, ptr(out), 64, 64, 64, 0, stream()) == -1


# ---- LayerNorm-fused schedule: producer (residual + bf16 copy + statistics) and consumer (norm applied in the epilogue)
def _lib_loaded():
    from aigv_assessor_b200 import _lib
    return _lib.load()


@pytest.mark.parametrize("M,N,K", [(2050, 1024, 1024), (1025, 1024, 4096), (300, 256, 128), (77, 64, 72), (1500, 3200, 256)])
def test_gemm_residual_stats(M, N, K):
    from tests.gpu_util import ptr, stream
    lib = _lib_loaded()
    a, w = _rand((M, K), 11), _rand((N, K), 12, K ** -0.5)
    g = torch.Generator().manual_seed(13)
    bias, ls, res = torch.randn(N, generator=g), 0.3 * torch.randn(N, generator=g), 2.0 * torch.randn((M, N), generator=g)
    x = dev(res.clone())
    da, dw, db, dl = dev(a), dev(w), dev(bias), dev(ls)
    parts = lib.aigv_stats_slices(N)
    assert parts == (N + 63) // 64
    xb = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
    stats = torch.full((parts, M, 2), float("nan"), device="cuda")
    st = lib.aigv_gemm_residual_stats(ptr(da), ptr(dw), ptr(db), ptr(dl), ptr(x), ptr(xb), ptr(stats), M, N, K, stream())
    assert st == 0
    torch.cuda.synchronize()
    ref = _ref(a, w, bias, ls, res)
    assert emax(x.cpu().numpy(), ref.cpu().numpy()) < 1e-4
    # the bf16 copy is the rounding of the fp32 stream this launch wrote, bit for bit
    assert torch.equal(xb, x.to(torch.bfloat16))
    s = stats.sum(0)
    assert torch.isfinite(s).all()
    x64 = x.double()
    assert torch.allclose(s[:, 0].double(), x64.sum(1), rtol=1e-5, atol=1e-2)
    assert torch.allclose(s[:, 1].double(), (x64 * x64).sum(1), rtol=1e-5, atol=1e-2)
    # deterministic: a second run on the same inputs reproduces every bit (no atomics in the statistics)
    x2 = dev(res.clone())
    stats2 = torch.zeros_like(stats)
    xb2 = torch.zeros_like(xb)
    assert lib.aigv_gemm_residual_stats(ptr(da), ptr(dw), ptr(db), ptr(dl), ptr(x2), ptr(xb2), ptr(stats2), M, N, K, stream()) == 0
    torch.cuda.synchronize()
    assert torch.equal(x, x2) and torch.equal(stats, stats2) and torch.equal(xb, xb2)


@pytest.mark.parametrize("rms", [False, True])
@pytest.mark.parametrize("M,N,D,gelu", [(2050, 3072, 1024, False), (1025, 4096, 1024, True), (333, 384, 128, False)])
def test_gemm_norm_fused_matches_layernorm_then_linear(M, N, D, gelu, rms):
    from tests.gpu_util import ptr, stream
    lib = _lib_loaded()
    g = torch.Generator().manual_seed(21)
    x = torch.randn((M, D), generator=g) * 1.5 + 0.7 * torch.randn((1, D), generator=g)  # per-channel offsets
    x[:, 5] += 40.0  # a massive-activation channel, as ViT residual streams have
    gamma = 1.0 + 0.2 * torch.randn(D, generator=g)
    beta = None if rms else 0.3 * torch.randn(D, generator=g)
    w = _rand((N, D), 22, D ** -0.5)
    bias = torch.randn(N, generator=g)
    eps = 1e-6
    dx = dev(x)
    xb = dx.to(torch.bfloat16)
    parts = 3
    stats = torch.zeros((parts, M, 2), device="cuda")
    # slices in arbitrary split: the consumer only needs their sum
    cuts = [0, D // 4, D // 2, D]
    for p in range(parts):
        seg = dx[:, cuts[p]:cuts[p + 1]]
        stats[p, :, 0] = seg.sum(1)
        stats[p, :, 1] = (seg * seg).sum(1)
    dw, dg, dbias = dev(w), dev(gamma), dev(bias)
    dbeta = dev(beta) if beta is not None else None
    wf = torch.empty_like(dw)
    wsum = torch.empty(N, device="cuda")
    bf = torch.empty(N, device="cuda")
    assert lib.aigv_fold_norm(ptr(dw), ptr(dg), ptr(dbeta), ptr(dbias), ptr(wf), ptr(wsum), ptr(bf), N, D, stream()) == 0
    out = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    st = lib.aigv_gemm_norm_fused(ptr(xb), ptr(wf), None if rms else ptr(wsum), ptr(bf), ptr(stats), parts, D, eps,
                                  ptr(out), M, N, D, int(gelu), stream())
    assert st == 0
    torch.cuda.synchronize()
    xd = dx.double()
    if rms:
        nrm = xd * torch.rsqrt((xd * xd).mean(1, keepdim=True) + eps) * dg.double()
    else:
        nrm = torch.nn.functional.layer_norm(xd, (D,), dg.double(), dbeta.double(), eps)
    ref = nrm @ dw.double().T + dbias.double()
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    # same error budget as "bf16 LayerNorm output, then bf16 GEMM" (the reference's own rounding points)
    assert emax(out.float().cpu().numpy(), ref.cpu().numpy()) < 8e-3


def test_assemble_tokens_stats():
    from tests.gpu_util import ptr, stream
    lib = _lib_loaded()
    n, P, D = 3, 16, 256
    g = torch.Generator().manual_seed(31)
    pe, cls, pos = torch.randn((n * P, D), generator=g), torch.randn(D, generator=g), torch.randn((P + 1, D), generator=g)
    dpe, dcls, dpos = dev(pe), dev(cls), dev(pos)
    M = n * (P + 1)
    x = torch.empty((M, D), device="cuda")
    xb = torch.empty((M, D), dtype=torch.bfloat16, device="cuda")
    stats = torch.full((2, M, 2), float("nan"), device="cuda")
    assert lib.aigv_assemble_tokens_stats(ptr(dpe), ptr(dcls), ptr(dpos), ptr(x), ptr(xb), ptr(stats), 2, n, P, D, stream()) == 0
    torch.cuda.synchronize()
    ref = torch.cat([cls.expand(n, 1, D), pe.view(n, P, D)], 1) + pos
    assert torch.equal(x.cpu(), ref.view(M, D))
    assert torch.equal(xb, x.to(torch.bfloat16))
    assert torch.all(stats[1] == 0)
    assert torch.allclose(stats[0, :, 0].cpu(), ref.view(M, D).sum(1), rtol=1e-5, atol=1e-4)
    assert torch.allclose(stats[0, :, 1].cpu(), (ref.view(M, D) ** 2).sum(1), rtol=1e-5, atol=1e-4)
