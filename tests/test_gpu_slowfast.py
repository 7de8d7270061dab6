"""the SlowFast-R50 motion trunk as GEMMs on the library (aigv_assessor_b200/slowfast.py) against the torch
fp32 restatement of pytorchvideo's published architecture (oracle/slowfast_oracle.py -- PARITY UNPINNED: the reference does
not contain this network, see the oracle's header), plus the per-op entry points against torch."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from aigv_assessor_b200 import _lib, slowfast, synth
from oracle import slowfast_oracle as so
from tests.gpu_util import ptr, stream

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16)


@pytest.mark.parametrize("shape,kernel,stride,pad", [((2, 4, 9, 10, 16), (3, 1, 1), (1, 1, 1), (1, 0, 0)),
                                                     ((1, 2, 12, 12, 8), (1, 3, 3), (1, 2, 2), (0, 1, 1)),
                                                     ((2, 8, 6, 5, 8), (7, 1, 1), (4, 1, 1), (3, 0, 0)),
                                                     ((1, 3, 8, 8, 24), (1, 1, 1), (1, 2, 2), (0, 0, 0))])
def test_im2col3d_plus_gemm_is_conv3d(shape, kernel, stride, pad):
    n, T, H, W, Cc = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = _bf(torch.randn((n, T, H, W, Cc), generator=g))
    cout = 40
    w = _bf(torch.randn((cout, Cc) + kernel, generator=g) * 0.2)
    b = torch.randn((cout,), generator=g)
    ref = F.relu(F.conv3d(x.float().permute(0, 4, 1, 2, 3), w.float(), b, stride=stride, padding=pad)).permute(0, 2, 3, 4, 1)
    lib = _lib.load()
    K = Cc * kernel[0] * kernel[1] * kernel[2]
    kpad = (K + 7) // 8 * 8
    rows = ref.shape[0] * ref.shape[1] * ref.shape[2] * ref.shape[3]
    xd, patches = x.cuda(), torch.empty((rows, kpad), dtype=torch.bfloat16, device="cuda")
    wk = torch.zeros((cout, kpad))
    wk[:, :K] = w.float().permute(0, 2, 3, 4, 1).reshape(cout, K)
    wd, bd = _bf(wk).cuda(), b.cuda()
    out = torch.empty((rows, cout), dtype=torch.bfloat16, device="cuda")
    assert lib.aigv_im2col3d(ptr(xd), ptr(patches), n, T, H, W, Cc, T * H * W * Cc, H * W * Cc, W * Cc, Cc, 1, *kernel, *stride, *pad,
                             kpad, stream()) == 0
    assert lib.aigv_gemm_ex(ptr(patches), kpad, ptr(wd), ptr(bd), ptr(out), cout, rows, cout, cout, kpad, 1, stream()) == 0
    torch.cuda.synchronize()
    err = float((out.float().cpu().view(ref.shape) - ref).abs().max() / ref.abs().max())
    assert err < 1e-2, err


@pytest.mark.parametrize("M,N,K,ldc,pad", [(1000, 64, 80, 96, False), (300, 256, 576, 256, False), (77, 16, 56, 80, False),
                                            (513, 8, 736, 8, False), (77, 16, 56, 80, True), (5000, 8, 736, 8, True), (400, 8, 8, 40, True),
                                            (130, 40, 24, 48, True)])
def test_gemm_ex_both_paths_with_leading_dimensions(M, N, K, ldc, pad):
    """pad: W / bias zero padded to a multiple of 32 rows -> the tensor-core kernel with a clipped store; else N as is."""
    g = torch.Generator().manual_seed(M + N)
    a, w, b = _bf(torch.randn((M, K + 8), generator=g)), _bf(torch.randn((N, K), generator=g) * 0.1), torch.randn((N,), generator=g)
    out = torch.full((M, ldc), 7.0, dtype=torch.bfloat16, device="cuda")
    n_pad = (N + 31) // 32 * 32 if pad else N
    wp, bp = torch.zeros((n_pad, K), dtype=torch.bfloat16), torch.zeros((n_pad,))
    wp[:N], bp[:N] = w, b
    ad, wd, bd = a.cuda(), wp.cuda(), bp.cuda()
    assert _lib.load().aigv_gemm_ex(ptr(ad), K + 8, ptr(wd), ptr(bd), ptr(out), ldc, M, N, n_pad, K, 0, stream()) == 0
    torch.cuda.synchronize()
    ref = a[:, :K].float() @ w.float().t() + b
    o = out.float().cpu()
    assert float((o[:, :N] - ref).abs().max() / ref.abs().max()) < 1e-2
    assert (o[:, N:] == 7.0).all()          # columns beyond N untouched


def test_pool_join_and_tail_ops():
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    x = _bf(torch.randn((3, 11, 13, 16), generator=g))                      # [n*t, H, W, C]
    ref = F.max_pool3d(x.float().permute(3, 0, 1, 2)[None], (1, 3, 3), (1, 2, 2), (0, 1, 1))[0].permute(1, 2, 3, 0)
    out = torch.empty(ref.shape, dtype=torch.bfloat16, device="cuda")
    assert lib.aigv_maxpool3d(ptr(x.cuda()), ptr(out), 3, 11, 13, 16, 3, 3, 2, 2, 1, 1, 16, stream()) == 0
    assert torch.equal(out.float().cpu(), ref)
    a, b = _bf(torch.randn((50, 24), generator=g)), _bf(torch.randn((50, 24), generator=g))
    o = torch.zeros((50, 40), dtype=torch.bfloat16, device="cuda")
    assert lib.aigv_add_relu(ptr(a.cuda()), ptr(b.cuda()), ptr(o), 50, 24, 40, stream()) == 0
    assert torch.equal(o[:, :24].float().cpu(), _bf(F.relu(a.float() + b.float())).float()) and float(o[:, 24:].abs().max()) == 0
    # tail: repeat_interleave(4) + AvgPool3d((8,7,7), stride 1) + AdaptiveAvgPool3d(1)  (eval2 :183-189)
    y = _bf(torch.randn((2, 2, 9, 8, 40), generator=g))                      # [n, T, H, W, C]
    r = y.float().permute(0, 4, 1, 2, 3).repeat_interleave(4, dim=2)
    r = F.adaptive_avg_pool3d(F.avg_pool3d(r, (8, 7, 7), stride=(1, 1, 1)), 1).flatten(1)
    f = torch.zeros((2, 50), dtype=torch.float32, device="cuda")
    assert lib.aigv_slowfast_tail(ptr(y.cuda()), ptr(f), 2, 2, 9, 8, 40, 40, 4, 8, 7, 7, 50, 10, stream()) == 0
    torch.cuda.synchronize()
    assert float((f[:, 10:].cpu() - r).abs().max()) < 1e-5 * float(r.abs().max()) + 1e-6
    assert float(f[:, :10].abs().max()) == 0


@pytest.mark.parametrize("B,T,size", [(2, 8, 224), (1, 8, 448)])
def test_trunk_matches_the_restated_architecture(B, T, size):
    sd = synth.make_slowfast_weights(seed=11)
    g = torch.Generator().manual_seed(size)
    pv = _bf(torch.randn((B * T, 3, size, size), generator=g))
    trunk = slowfast.SlowFastTrunk(sd)
    feat = trunk(pv.cuda(), B).cpu()
    torch.cuda.synchronize()
    again = trunk((pv * 0.5).cuda(), B)          # graph replay with new data, then the first input again
    assert not torch.equal(again.cpu(), feat)
    assert torch.equal(trunk(pv.cuda(), B).cpu(), feat)
    eager = slowfast.SlowFastTrunk(sd, use_graph=False)(pv.cuda(), B).cpu()
    assert torch.equal(eager, feat)
    ref = so.motion_features(pv.float(), B, {k: torch.from_numpy(v) for k, v in sd.items()})
    assert feat.shape == (B, 2304) and torch.isfinite(feat).all()
    e_max = float((feat - ref).abs().max() / ref.abs().max())
    e_l2 = float((feat - ref).norm() / ref.norm())
    print(f"slowfast trunk {B}x{T}x{size}: e_max {e_max:.2e} e_l2 {e_l2:.2e} ({trunk.launches} launches)")
    assert e_max < 2e-2 and e_l2 < 1e-2


def test_trunk_rejects_bad_input():
    trunk = slowfast.SlowFastTrunk(synth.make_slowfast_weights(seed=1))
    with pytest.raises(ValueError):
        trunk(torch.zeros((8, 3, 100, 100), dtype=torch.bfloat16, device="cuda"), 1)     # too small for the 7x7 pool
    with pytest.raises(ValueError):
        trunk(torch.zeros((7, 3, 224, 224), dtype=torch.bfloat16, device="cuda"), 2)     # does not split into clips
    bad = dict(synth.make_slowfast_weights(seed=1))
    name = slowfast.layer_table()[0][0]
    bad[name] = bad[name][:, :, :, :5]
    with pytest.raises(ValueError):
        slowfast.SlowFastTrunk(bad)


def test_clip_scorer_with_its_own_motion_trunk(golden_dir):
    """ClipScorer(motion_features=None): frames -> SlowFast trunk -> motion_mlp -> splice -> InternLM2 -> mlpscore entirely on the
    library, against the same chain in the oracles (slowfast_oracle -> head_oracle) on the CUDA visual tokens."""
    import json, os
    from aigv_assessor_b200.config import VisionPathConfig
    from aigv_assessor_b200.scoring import ClipScorer
    from oracle import head_oracle as ho
    with open(os.path.join(golden_dir, "golden_meta_full.json")) as f:
        meta = json.load(f)["tiny"]
    cfg = VisionPathConfig(**meta["vis_cfg"])
    sd = dict(synth.make_weights(cfg, seed=meta["seeds"]["vis"]))
    sd_head = synth.make_head_weights(seed=meta["seeds"]["head"])
    sd_head["mlpscore.fc5.weight"] = synth.bf16_round(sd_head["mlpscore.fc5.weight"] * np.float32(meta["fc5_scale"]))
    sd_sf = synth.make_slowfast_weights(seed=21)
    sd.update(sd_head)
    sd.update(sd_sf)
    pv = synth.make_pixel_values(meta["n_frames"], cfg, seed=meta["seeds"]["x"])
    ids = synth.make_clip_prompt(meta["n_frames"], seed=meta["seeds"]["ids"], vocab_size=meta["llm"]["vocab_size"])
    scorer = ClipScorer(sd, cfg, meta["llm"], synth.IMG_CONTEXT_ID, max_tiles=8, max_tokens=ids.shape[1])
    assert scorer.slowfast is not None
    pvd = torch.from_numpy(pv).to(torch.bfloat16).cuda()
    out = scorer(pvd, torch.from_numpy(ids).cuda())
    vit = scorer.vision.extract_feature(pvd).float().cpu()
    torch.cuda.synchronize()
    sdt = ho.to_torch(sd_head)
    feat = so.motion_features(torch.from_numpy(pv), 1, {k: torch.from_numpy(v) for k, v in sd_sf.items()})
    emb = ho.splice(torch.from_numpy(ids), sdt["language_model.model.tok_embeddings.weight"], vit, ho.motion_embeds(feat, sdt),
                    synth.IMG_CONTEXT_ID)
    hidden, _ = ho.internlm2(emb, sdt, meta["llm"])
    ref, _ = ho.mlpscore(hidden[:, -4, :], sdt)
    d = abs(float(out["score1"][0]) - float(ref[0]))
    print(f"ClipScorer with its own SlowFast trunk: score1 {float(out['score1'][0]):.4f} vs oracle chain {float(ref[0]):.4f} |d| {d:.2e}")
    assert d <= 0.02
