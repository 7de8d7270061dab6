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
