"""SURVEY 8f next-1, second piece: the InternLM2 prefill as CUDA kernels (rotary embedding, causal grouped-query attention
on tcgen05, RMSNorm-fused wqkv / w1|w3 GEMMs, SwiGLU, fused-residual wo / w2, final norm + output head on selected rows)
against oracle/head_oracle.py, which tests/test_oracle_golden.py pins to the unmodified reference forward."""
import json
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from aigv_assessor_b200 import llm as L
from aigv_assessor_b200 import llm_glue, synth
from oracle import head_oracle as ho
from oracle import vit_oracle as vo

pytestmark = pytest.mark.gpu


def _qkv(B, N, nh, nkv, hd, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn((B * N, (nh + 2 * nkv) * hd), generator=g) * scale).to(torch.bfloat16)


def _split(qkv, B, N, nh, nkv, hd):
    """InternLM2's rearrange (modeling_internlm2.py:375-389): -> q [B,nh,N,hd], k/v [B,nkv,N,hd] (fp32)."""
    G = nh // nkv
    t = qkv.float().view(B, N, nkv, G + 2, hd)
    q = t[..., :G, :].reshape(B, N, nh, hd).transpose(1, 2)
    return q, t[..., -2, :].transpose(1, 2), t[..., -1, :].transpose(1, 2)


def _ref_attention(qkv, B, N, nh, nkv, hd, causal):
    q, k, v = _split(qkv, B, N, nh, nkv, hd)
    G = nh // nkv
    k = k[:, :, None].expand(B, nkv, G, N, hd).reshape(B, nh, N, hd)
    v = v[:, :, None].expand(B, nkv, G, N, hd).reshape(B, nh, N, hd)
    a = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(hd)
    if causal:
        a = a + torch.full((N, N), float("-inf")).triu(1)
    return torch.matmul(F.softmax(a, dim=-1), v).transpose(1, 2).reshape(B * N, nh * hd)


@pytest.mark.parametrize("B,N,nh,nkv,hd", [(1, 1, 2, 1, 64), (2, 100, 4, 2, 64), (1, 128, 2, 2, 128), (1, 129, 4, 1, 128),
                                           (2, 300, 8, 2, 128), (1, 513, 4, 2, 64), (1, 2108, 8, 2, 128),
                                           # more work items than persistent CTAs (the shape of the 7B prefill) / two prompts
                                           (1, 2108, 32, 8, 128), (2, 1100, 16, 8, 128), (3, 700, 16, 8, 64)])
@pytest.mark.parametrize("causal", [True, False])
def test_gqa_attention_matches_fp32_softmax(B, N, nh, nkv, hd, causal):
    qkv = _qkv(B, N, nh, nkv, hd, seed=N + nh)
    out = L.attention_gqa(qkv.cuda(), B, N, nh, nkv, hd, causal=causal)
    torch.cuda.synchronize()
    ref = _ref_attention(qkv, B, N, nh, nkv, hd, causal)
    err = float((out.float().cpu() - ref).abs().max() / ref.abs().max())
    assert torch.isfinite(out.float()).all()
    assert err < 1.2e-2, err   # bf16 P and bf16 output against an fp32 softmax


def test_causal_attention_peaked_scores_and_first_rows():
    """Large logits (forced rescales of the stale maximum) and the rows that see one / two keys."""
    B, N, nh, nkv, hd = 1, 400, 4, 2, 128
    qkv = _qkv(B, N, nh, nkv, hd, seed=9, scale=3.0)
    out = L.attention_gqa(qkv.cuda(), B, N, nh, nkv, hd, causal=True).float().cpu()
    ref = _ref_attention(qkv, B, N, nh, nkv, hd, True)
    assert float((out - ref).abs().max() / ref.abs().max()) < 1.5e-2
    _, _, v = _split(qkv, B, N, nh, nkv, hd)
    # row 0 attends to key 0 only: the output is exactly v[0] of its kv group
    v0 = v[0, :, 0, :].repeat_interleave(nh // nkv, dim=0).reshape(-1)
    assert torch.equal(out[0], v0)


@pytest.mark.parametrize("B,N,nh,nkv,hd", [(2, 77, 4, 2, 128), (1, 300, 4, 4, 64)])
def test_rope_matches_the_rotate_half_form(B, N, nh, nkv, hd):
    theta = 1e6
    qkv = _qkv(B, N, nh, nkv, hd, seed=3)
    out = L.rope_(qkv.clone().cuda(), B, N, nh, nkv, hd, theta).float().cpu()
    q, k, v = _split(qkv, B, N, nh, nkv, hd)
    inv = 1.0 / (theta ** (torch.arange(0, hd, 2).float() / hd))
    fr = torch.einsum("i,j->ij", torch.arange(N).float(), inv)
    emb = torch.cat((fr, fr), dim=-1)
    cos, sin = emb.cos()[None, None], emb.sin()[None, None]
    rq = q * cos + ho._rotate_half(q) * sin
    rk = k * cos + ho._rotate_half(k) * sin
    oq, ok, ov = _split(out, B, N, nh, nkv, hd)
    assert torch.equal(ov, v)                                      # v untouched
    assert float((oq - rq).abs().max()) < 2e-2 * float(rq.abs().max())   # one bf16 rounding of an O(4) value
    assert float((ok - rk).abs().max()) < 2e-2 * float(rk.abs().max())
    assert float((oq - rq).abs().mean()) < 2e-3 * float(rq.abs().max())


def _head_case(golden_dir, name="tiny"):
    with open(os.path.join(golden_dir, "golden_meta_full.json")) as f:
        meta = json.load(f)[name]
    sd_head = synth.make_head_weights(seed=meta["seeds"]["head"])
    sd_head["mlpscore.fc5.weight"] = synth.bf16_round(sd_head["mlpscore.fc5.weight"] * np.float32(meta["fc5_scale"]))
    return meta, sd_head, np.load(os.path.join(golden_dir, f"full_{name}.npz"))


def test_prefill_matches_the_oracle_on_random_embeddings(golden_dir):
    """Two prompts of 300 tokens through the 2-layer 4096-wide InternLM2 of the goldens: residual stream, final-norm rows and
    logits against the fp32 oracle on the same bf16 inputs."""
    meta, sd_head, _ = _head_case(golden_dir)
    llm = meta["llm"]
    sdt = ho.to_torch(sd_head)
    g = torch.Generator().manual_seed(17)
    B, N = 2, 300
    emb = (torch.randn((B, N, llm["hidden_size"]), generator=g) * 0.5).to(torch.bfloat16)
    eng = L.LanguageEngine(llm, max_tokens=B * N).load_state_dict(sd_head)
    eng.prefill(emb.cuda())
    rows = [N - 4, N - 1, N + 5, 2 * N - 4]
    hid = eng.hidden_rows(rows).cpu()
    logits = eng.logits_rows(rows).cpu()
    torch.cuda.synchronize()
    ref_h, ref_l = ho.internlm2(emb.float(), sdt, llm)
    ref_h, ref_l = ref_h.reshape(B * N, -1)[rows], ref_l.reshape(B * N, -1)[rows]
    e_h = vo.error_metrics(hid.numpy(), ref_h.numpy())
    e_l = vo.error_metrics(logits.numpy(), ref_l.numpy())
    print(f"prefill vs oracle: hidden rows e_max/e_l2 {e_h}  logits {e_l}")
    assert e_h[0] < 2e-2 and e_l[0] < 2e-2
    assert eng.launch_count >= 1 + 5 * llm["num_hidden_layers"]   # seed + (wqkv+rope, attention, wo, w1|w3+SwiGLU, w2) per layer


def test_score_with_everything_on_the_gpu(golden_dir):
    """The whole stage-2 forward on the GPU -- CUDA visual tokens -> CUDA splice -> CUDA InternLM2 prefill -> CUDA mlpscore and
    output head -- against the reference's own score1 / answer-span argmax (north_star: 0.02 absolute). Only the SlowFast
    stand-in and motion_mlp (one 2304-vector per clip) stay on the host."""
    from aigv_assessor_b200.config import VisionPathConfig
    from aigv_assessor_b200.engine import VisionEngine
    meta, sd_head, g = _head_case(golden_dir)
    llm = meta["llm"]
    cfg = VisionPathConfig(**meta["vis_cfg"])
    sd_vis = synth.make_weights(cfg, seed=meta["seeds"]["vis"])
    pv = synth.make_pixel_values(meta["n_frames"], cfg, seed=meta["seeds"]["x"])
    ids = synth.make_clip_prompt(meta["n_frames"], seed=meta["seeds"]["ids"], vocab_size=llm["vocab_size"])
    sdt = ho.to_torch(sd_head)
    eng = VisionEngine(cfg, max_tiles=8).load_state_dict(sd_vis)
    emb = eng.extract_feature(torch.from_numpy(pv).to(torch.bfloat16).cuda())
    motion = ho.motion_embeds(ho.fake_slowfast(torch.from_numpy(pv), 1), sdt).to(torch.bfloat16).cuda()
    tok = sdt["language_model.model.tok_embeddings.weight"].to(torch.bfloat16).cuda()
    x = llm_glue.splice_tokens(torch.from_numpy(ids).cuda(), tok, emb, motion, synth.IMG_CONTEXT_ID)
    lm = L.LanguageEngine(llm, max_tokens=ids.shape[1]).load_state_dict(sd_head)
    hid, arg, logits = lm.stage2_tail(x, n_answer=48)
    score, pre = llm_glue.MlpScoreHead(sd_head, "cuda")(hid)
    torch.cuda.synchronize()
    d = abs(float(score[0]) - float(g["score1"][0]))
    e_hid = vo.error_metrics(hid.cpu().numpy(), g["hidden_m4"])
    ref_logits = g["logits_tail"]
    top2 = np.sort(ref_logits, axis=1)[:, -2:]
    decided = (top2[:, 1] - top2[:, 0]) > 1e-2 * np.abs(ref_logits).max()
    same = arg.cpu().numpy() == g["logit_tail"]
    e_log = vo.error_metrics(logits.cpu().numpy(), ref_logits)
    print(f"all on the GPU: score1 {float(score[0]):.4f} vs reference {float(g['score1'][0]):.4f} |d| {d:.2e} (tol 0.02); "
          f"hidden[-4] e_max/e_l2 {e_hid}; logits {e_log}; argmax equal {int(same.sum())}/48 ({int(decided.sum())} decided)")
    assert d <= 0.02
    assert same[decided].all()


def test_prefill_rejects_bad_calls(golden_dir):
    meta, sd_head, _ = _head_case(golden_dir)
    llm = meta["llm"]
    eng = L.LanguageEngine(llm, max_tokens=64)
    x = torch.zeros((1, 8, llm["hidden_size"]), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(ValueError):
        eng.prefill(x)                       # weights not loaded
    eng.load_state_dict(sd_head)
    with pytest.raises(ValueError):
        eng.prefill(torch.zeros((1, 8, 128), dtype=torch.bfloat16, device="cuda"))   # wrong width
    with pytest.raises(ValueError):
        eng.prefill(torch.zeros((1, 65, llm["hidden_size"]), dtype=torch.bfloat16, device="cuda"))   # beyond max_tokens
    with pytest.raises(ValueError):
        eng.hidden_rows([0])                 # no prefill yet
    eng.prefill(x)
    with pytest.raises(ValueError):
        eng.hidden_rows([8])                 # outside the last prefill
    with pytest.raises(ValueError):
        eng.set_weight("model.layers.0.attention.wq.weight", torch.zeros(4))


def test_clip_scorer_reproduces_the_reference_forward(golden_dir):
    """scoring.ClipScorer = the eval-time forward (eval2 :306-488) with every module of the reference replaced by this
    library (motion_mlp included; only the SlowFast stand-in's 2304-vector comes from the host): score1 and the answer-span
    argmax against the golden the unmodified reference produced, for both vision towers of the goldens."""
    from aigv_assessor_b200.config import VisionPathConfig
    from aigv_assessor_b200.scoring import ClipScorer
    for name in ("tiny", "vit300m"):
        meta, sd_head, g = _head_case(golden_dir, name)
        cfg = VisionPathConfig(**meta["vis_cfg"])
        sd = dict(synth.make_weights(cfg, seed=meta["seeds"]["vis"]))
        sd.update(sd_head)
        pv = synth.make_pixel_values(meta["n_frames"], cfg, seed=meta["seeds"]["x"])
        ids = synth.make_clip_prompt(meta["n_frames"], seed=meta["seeds"]["ids"], vocab_size=meta["llm"]["vocab_size"])
        scorer = ClipScorer(sd, cfg, meta["llm"], synth.IMG_CONTEXT_ID, max_tiles=8, max_tokens=ids.shape[1])
        feat = ho.fake_slowfast(torch.from_numpy(pv), 1)           # what eval2 :343 hands to motion_mlp
        out = scorer(torch.from_numpy(pv).to(torch.bfloat16).cuda(), torch.from_numpy(ids).cuda(), feat.cuda())
        torch.cuda.synchronize()
        d = abs(float(out["score1"][0]) - float(g["score1"][0]))
        ref_logits = g["logits_tail"]
        top2 = np.sort(ref_logits, axis=1)[:, -2:]
        decided = (top2[:, 1] - top2[:, 0]) > 1e-2 * np.abs(ref_logits).max()
        same = out["logit"].cpu().numpy() == g["logit_tail"]
        ref_mot = ho.motion_embeds(feat, ho.to_torch(sd_head))
        e_mot = float((scorer.motion(feat.cuda()).float().cpu() - ref_mot).abs().max() / ref_mot.abs().max())
        print(f"ClipScorer[{name}]: score1 {float(out['score1'][0]):.4f} vs reference {float(g['score1'][0]):.4f} |d| {d:.2e}; "
              f"motion token e_max {e_mot:.2e}; argmax equal {int(same.sum())}/48 ({int(decided.sum())} decided)")
        assert e_mot < 2e-2
        assert d <= 0.02
        assert same[decided].all()
        del scorer


def test_clip_scorer_from_a_model_shaped_object(golden_dir):
    """ClipScorer.from_model on an object shaped like the eval2 InternVLChatModel (vision_model.config, mlp1, language_model.config,
    state_dict(), img_context_token_id): the configs are read off the model, the result equals the explicitly built scorer."""
    import types
    from aigv_assessor_b200.config import VisionPathConfig
    from aigv_assessor_b200.scoring import ClipScorer
    meta, sd_head, g = _head_case(golden_dir, "tiny")
    cfg = VisionPathConfig(**meta["vis_cfg"])
    sd = dict(synth.make_weights(cfg, seed=meta["seeds"]["vis"]))
    sd.update(sd_head)
    llm = meta["llm"]
    model = types.SimpleNamespace()
    model.vision_model = types.SimpleNamespace(config=types.SimpleNamespace(
        hidden_size=cfg.hidden_size, num_attention_heads=cfg.num_heads, intermediate_size=cfg.intermediate_size,
        num_hidden_layers=cfg.num_layers, image_size=cfg.image_size, patch_size=cfg.patch_size, qkv_bias=cfg.qkv_bias,
        qk_normalization=cfg.qk_normalization, norm_type=cfg.norm_type, layer_norm_eps=cfg.layer_norm_eps))
    model.mlp1 = types.SimpleNamespace(state_dict=lambda: {"3.weight": torch.from_numpy(sd["mlp1.3.weight"])})
    model.language_model = types.SimpleNamespace(config=types.SimpleNamespace(
        hidden_size=llm["hidden_size"], num_hidden_layers=llm["num_hidden_layers"], num_attention_heads=llm["num_attention_heads"],
        num_key_value_heads=llm["num_key_value_heads"], intermediate_size=llm["intermediate_size"], vocab_size=llm["vocab_size"],
        rms_norm_eps=llm["rms_norm_eps"], rope_theta=llm["rope_theta"], max_position_embeddings=llm["max_position_embeddings"]))
    model.state_dict = lambda: {k: torch.from_numpy(v) for k, v in sd.items()}
    model.img_context_token_id = synth.IMG_CONTEXT_ID
    pv = synth.make_pixel_values(meta["n_frames"], cfg, seed=meta["seeds"]["x"])
    ids = synth.make_clip_prompt(meta["n_frames"], seed=meta["seeds"]["ids"], vocab_size=llm["vocab_size"])
    scorer = ClipScorer.from_model(model, max_tiles=8, max_tokens=ids.shape[1])
    feat = ho.fake_slowfast(torch.from_numpy(pv), 1)
    out = scorer(torch.from_numpy(pv).to(torch.bfloat16).cuda(), torch.from_numpy(ids).cuda(), feat.cuda())
    torch.cuda.synchronize()
    assert abs(float(out["score1"][0]) - float(g["score1"][0])) <= 0.02
    with pytest.raises(ValueError):
        scorer(torch.from_numpy(pv).to(torch.bfloat16).cuda(), torch.from_numpy(ids).cuda())   # no slowfast_model.* weights, no features


@pytest.mark.parametrize("B,N,hidden,nh,nkv,inter", [(3, 77, 2048, 16, 8, 8192), (1, 1, 2048, 16, 8, 2048), (2, 130, 512, 8, 8, 1024),
                                                     (1, 260, 256, 4, 1, 512)])
def test_prefill_other_widths_and_tiny_prompts(B, N, hidden, nh, nkv, inter):
    """InternLM2-1.8B's width (2048, 16/8 heads of 128), multi-head without grouping, a single kv head of 64, one-token prompts."""
    llm = dict(hidden_size=hidden, num_hidden_layers=2, num_attention_heads=nh, num_key_value_heads=nkv, intermediate_size=inter,
               vocab_size=96, rms_norm_eps=1e-5, rope_theta=1e6, max_position_embeddings=4096)
    sd = synth.make_head_weights(seed=hidden + N, llm=llm)
    g = torch.Generator().manual_seed(N)
    emb = (torch.randn((B, N, hidden), generator=g) * 0.5).to(torch.bfloat16)
    eng = L.LanguageEngine(llm, max_tokens=B * N).load_state_dict(sd)
    eng.prefill(emb.cuda())
    rows = sorted({0, N - 1, (B - 1) * N, B * N - 1})
    hid, logits = eng.hidden_rows(rows).cpu(), eng.logits_rows(rows).cpu()
    torch.cuda.synchronize()
    ref_h, ref_l = ho.internlm2(emb.float(), ho.to_torch(sd), llm)
    e_h = vo.error_metrics(hid.numpy(), ref_h.reshape(B * N, -1)[rows].numpy())
    e_l = vo.error_metrics(logits.numpy(), ref_l.reshape(B * N, -1)[rows].numpy())
    assert e_h[0] < 2e-2 and e_l[0] < 2e-2, (e_h, e_l)


def test_prefill_trained_like_statistics(golden_dir):
    """What trained LLMs look like where the numerics care: massive-activation channels in the residual stream (+-400), peaked
    attention (wqkv x3), a large-norm first token -- RMSNorm statistics from sum-of-squares slices, the stale-max softmax with
    causal masking, the SwiGLU epilogue's exponent range."""
    meta, sd_head, _ = _head_case(golden_dir)
    llm = dict(meta["llm"])
    sd = dict(sd_head)
    for i in range(llm["num_hidden_layers"]):
        k = f"language_model.model.layers.{i}.attention.wqkv.weight"
        sd[k] = synth.bf16_round(sd[k] * np.float32(3.0))
    g = torch.Generator().manual_seed(23)
    B, N, H = 1, 700, llm["hidden_size"]
    emb = torch.randn((B, N, H), generator=g) * 0.5
    emb[:, :, 7] += 400.0
    emb[:, :, 1033] -= 250.0
    emb[:, 0, :] *= 20.0
    emb = emb.to(torch.bfloat16)
    eng = L.LanguageEngine(llm, max_tokens=B * N).load_state_dict(sd)
    eng.prefill(emb.cuda())
    rows = [0, 1, 350, N - 4, N - 1]
    hid, logits = eng.hidden_rows(rows).cpu(), eng.logits_rows(rows).cpu()
    torch.cuda.synchronize()
    ref_h, ref_l = ho.internlm2(emb.float(), ho.to_torch(sd), llm)
    e_h = vo.error_metrics(hid.numpy(), ref_h[0, rows].numpy())
    e_l = vo.error_metrics(logits.numpy(), ref_l[0, rows].numpy())
    print(f"trained-like prefill: hidden rows e_max/e_l2 {e_h} logits {e_l}; |hidden| max {float(ref_h.abs().max()):.1f}")
    assert torch.isfinite(hid).all() and torch.isfinite(logits).all()
    assert e_h[0] < 2e-2 and e_l[0] < 2e-2


def test_sharded_eval_loop_scores_and_metrics(golden_dir):
    """sharding.score_clips_sharded (world 1): four synthetic clips -> score1 each -> SRCC / PLCC / KRCC as the eval script prints
    them; clip 0 is the golden clip, so its score is the reference's."""
    from aigv_assessor_b200 import sharding
    from aigv_assessor_b200.config import VisionPathConfig
    from aigv_assessor_b200.scoring import ClipScorer
    meta, sd_head, g = _head_case(golden_dir, "tiny")
    cfg = VisionPathConfig(**meta["vis_cfg"])
    sd = dict(synth.make_weights(cfg, seed=meta["seeds"]["vis"]))
    sd.update(sd_head)
    ids = torch.from_numpy(synth.make_clip_prompt(meta["n_frames"], seed=meta["seeds"]["ids"], vocab_size=meta["llm"]["vocab_size"])).cuda()
    scorer = ClipScorer(sd, cfg, meta["llm"], synth.IMG_CONTEXT_ID, max_tiles=8, max_tokens=3 * ids.shape[1])
    clips, feats = [], []
    for k in range(4):
        pv = synth.make_pixel_values(meta["n_frames"], cfg, seed=meta["seeds"]["x"] + 100 * k)
        clips.append(torch.from_numpy(pv).to(torch.bfloat16).cuda())
        feats.append(ho.fake_slowfast(torch.from_numpy(pv), 1).cuda())
    res = sharding.score_clips_sharded(scorer, clips, ids, mos=[3.0, 2.0, 4.0, 1.0], motion_features=feats)
    torch.cuda.synchronize()
    # grouped forwards (3 clips: the workspace holds 3 prompts, then 1) against one clip per forward
    one = sharding.score_clips_sharded(scorer, clips, ids, motion_features=feats, clips_per_forward=1)
    torch.cuda.synchronize()
    assert float((res["scores"] - one["scores"]).abs().max()) <= 5e-3
    assert abs(float(one["scores"][0]) - float(g["score1"][0])) <= 0.02
    assert res["scores"].shape == (4,) and torch.isfinite(res["scores"]).all()
    assert abs(float(res["scores"][0]) - float(g["score1"][0])) <= 0.02
    assert all(-1.0 <= res[k] <= 1.0 for k in ("srcc", "plcc", "krcc"))


def test_prefill_full_width_batch_equals_single_prompt():
    """InternLM2-7B's widths (4096, 32/8 heads of 128, I = 14336; two layers) on four copies of one 2108-token clip prompt: 33 row tiles
    select the grouped GEMM tile order for w1|w3 (M fastest inside groups of 16 row tiles) where a single prompt runs plain M fastest.
    Tile order, work-list dealing and batch position must not change a single bit of a row: the residual stream of every copy equals
    the single-prompt prefill exactly (a size-independent property, checked at the real size)."""
    llm = dict(hidden_size=4096, num_hidden_layers=2, num_attention_heads=32, num_key_value_heads=8, intermediate_size=14336,
               vocab_size=128, rms_norm_eps=1e-5, rope_theta=1e6, max_position_embeddings=32768)
    B, N, H = 4, 2108, 4096
    eng = L.LanguageEngine(llm, max_tokens=B * N)
    g = torch.Generator(device="cuda").manual_seed(7)

    def rnd(shape, fan_in):
        return ((torch.rand(shape, device="cuda", generator=g) * 2 - 1) * fan_in ** -0.5).to(torch.bfloat16)
    for i in range(2):
        p = f"model.layers.{i}."
        eng.set_weight(p + "attention.wqkv.weight", rnd((6144, H), H / 4.0))
        eng.set_weight(p + "attention.wo.weight", rnd((H, H), H))
        eng.set_weight(p + "feed_forward.w1.weight", rnd((14336, H), H))
        eng.set_weight(p + "feed_forward.w3.weight", rnd((14336, H), H))
        eng.set_weight(p + "feed_forward.w2.weight", rnd((H, 14336), 14336))
        eng.set_weight(p + "attention_norm.weight", torch.ones(H, device="cuda"))
        eng.set_weight(p + "ffn_norm.weight", torch.ones(H, device="cuda"))
    eng.set_weight("model.norm.weight", torch.ones(H, device="cuda"))
    eng.set_weight("output.weight", rnd((128, H), H))
    eng._check(eng.lib.aigv_llm_finalize_weights(eng._h, None), "finalize")
    eng._finalized = True
    one = (torch.randn((1, N, H), device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    eng.prefill(one)
    x1 = eng.residual().clone()
    eng.prefill(one.repeat(B, 1, 1).contiguous())
    xb = eng.residual()
    torch.cuda.synchronize()
    assert torch.isfinite(x1).all() and float(x1.abs().max()) > 0.1
    for b in range(B):
        assert torch.equal(xb[b], x1[0]), f"copy {b} of the prompt differs from the single-prompt prefill"
    eng.close()
