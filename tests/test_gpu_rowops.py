"""Row-wise kernels (norms, im2col, token assembly, pixel-shuffle+LN) vs the numpy oracle, through the C-ABI."""
import numpy as np
import pytest
import torch

from aigv_assessor_b200 import _lib, synth
from oracle import vit_oracle as vo
from tests.gpu_util import dev, emax, ptr, stream

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,D,rms", [(1025, 1024, 0), (77, 128, 0), (300, 3200, 1), (64, 256, 1)])
def test_norm(rows, D, rms):
    rng = np.random.default_rng(rows + D)
    x = (rng.standard_normal((rows, D)) * 2 + 0.5).astype(np.float32)
    w = (1 + 0.1 * rng.standard_normal(D)).astype(np.float32)
    b = (0.1 * rng.standard_normal(D)).astype(np.float32)
    y = torch.empty((rows, D), dtype=torch.bfloat16, device="cuda")
    lib = _lib.load()
    dx, dw, db = dev(x), dev(w), dev(b)  # keep the device buffers alive until the kernel has run
    st = lib.aigv_layernorm(ptr(dx), ptr(dw), None if rms else ptr(db), ptr(y), rows, D, 1e-6, rms, stream())
    assert st == 0
    torch.cuda.synchronize()
    ref = vo.rms_norm(x, w, 1e-6) if rms else vo.layer_norm(x, w, b, 1e-6)
    got = y.float().cpu().numpy()
    assert emax(got, ref) < 4e-3


@pytest.mark.parametrize("rows,D", [(200, 256), (77, 3200), (33, 4096)])  # register-resident (<=1024, <=3328) and two-pass
def test_rmsnorm_inplace_on_column_block(rows, D):
    rng = np.random.default_rng(3)
    x = synth.bf16_round(rng.standard_normal((rows, 3 * D)).astype(np.float32))
    w = (1 + 0.1 * rng.standard_normal(D)).astype(np.float32)
    t, dw = dev(x, torch.bfloat16), dev(w)
    lib = _lib.load()
    assert lib.aigv_rmsnorm_inplace(ptr(t), ptr(dw), rows, D, 3 * D, D, 1e-6, stream()) == 0
    torch.cuda.synchronize()
    got = t.float().cpu().numpy()
    ref = x.copy()
    ref[:, D:2 * D] = vo.rms_norm(x[:, D:2 * D], w, 1e-6)
    assert np.array_equal(got[:, :D], x[:, :D]) and np.array_equal(got[:, 2 * D:], x[:, 2 * D:])
    assert emax(got, ref) < 4e-3


def test_im2col_and_assemble_equal_oracle_embeddings():
    from aigv_assessor_b200.config import tiny
    from tests.gpu_util import gemm
    cfg = tiny()
    sd = synth.make_weights(cfg, seed=5)
    pv = synth.make_pixel_values(2, cfg, seed=6)
    n, D, P, kp = 2, cfg.hidden_size, cfg.num_patches, 640
    lib = _lib.load()
    patches = torch.empty((n * P, kp), dtype=torch.bfloat16, device="cuda")
    dpv = dev(pv, torch.bfloat16)
    assert lib.aigv_im2col(ptr(dpv), ptr(patches), n, cfg.image_size, cfg.patch_size, kp, stream()) == 0
    torch.cuda.synchronize()
    # bit-exact gather check against the oracle's patch flattening
    g, ps = cfg.grid, cfg.patch_size
    ref_p = pv.reshape(n, 3, g, ps, g, ps).transpose(0, 2, 4, 1, 3, 5).reshape(n * P, 588)
    got_p = patches.float().cpu().numpy()
    assert np.array_equal(got_p[:, :588], ref_p) and not got_p[:, 588:].any()
    w = np.zeros((D, kp), np.float32)
    w[:, :588] = sd["vision_model.embeddings.patch_embedding.weight"].reshape(D, 588)
    pe = gemm(patches, dev(w, torch.bfloat16), bias=dev(sd["vision_model.embeddings.patch_embedding.bias"]),
              out_dtype=torch.float32)
    x = torch.empty((n, P + 1, D), dtype=torch.float32, device="cuda")
    dcls = dev(sd["vision_model.embeddings.class_embedding"].reshape(-1))
    dpos = dev(sd["vision_model.embeddings.position_embedding"].reshape(-1, D))
    assert lib.aigv_assemble_tokens(ptr(pe), ptr(dcls), ptr(dpos), ptr(x), n, P, D, stream()) == 0
    torch.cuda.synchronize()
    ref = vo.embeddings(pv, sd, cfg)
    assert emax(x.cpu().numpy(), ref) < 1e-5


@pytest.mark.parametrize("apply_ln,D", [(0, 128), (1, 128), (1, 256), (1, 1024)])
def test_pixel_shuffle_ln(apply_ln, D):
    n, g = 2, 32
    rng = np.random.default_rng(8)
    x = synth.bf16_round(rng.standard_normal((n, g * g + 1, D)).astype(np.float32))
    w = (1 + 0.1 * rng.standard_normal(4 * D)).astype(np.float32)
    b = (0.1 * rng.standard_normal(4 * D)).astype(np.float32)
    y = torch.empty((n * 256, 4 * D), dtype=torch.bfloat16, device="cuda")
    lib = _lib.load()
    dx, dw, db = dev(x), dev(w), dev(b)
    assert lib.aigv_pixel_shuffle(ptr(dx), ptr(dw), ptr(db), ptr(y), n, g, D, 1e-5, apply_ln, stream()) == 0
    torch.cuda.synchronize()
    ref = vo.pixel_shuffle(x[:, 1:].reshape(n, g, g, D), 0.5, "v2").reshape(n * 256, 4 * D)
    got = y.float().cpu().numpy()
    if apply_ln:
        assert emax(got, vo.layer_norm(ref, w, b, 1e-5)) < 4e-3
    else:
        assert np.array_equal(got, ref)  # pure data movement: bit exact
