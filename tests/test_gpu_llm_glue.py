"""SURVEY 8f next-1, first pieces: the visual-token splice and the mlpscore head as CUDA kernels, against the head oracle
(oracle/head_oracle.py: pinned to the unmodified reference forward by tests/test_oracle_golden.py)."""
import numpy as np
import pytest
import torch

from aigv_assessor_b200 import llm_glue, synth
from oracle import head_oracle as ho

pytestmark = pytest.mark.gpu


def _rand_bf16(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16)


@pytest.mark.parametrize("B,frames", [(1, 8), (3, 2)])
def test_splice_is_bit_identical_to_the_reference_assignment(B, frames):
    V, Cc = 512, 4096
    ids = np.concatenate([synth.make_clip_prompt(frames, seed=40 + b, n_text=37 + (0 if B == 1 else 0)) for b in range(B)], axis=0)
    tok = _rand_bf16((V, Cc), 1)
    vit = _rand_bf16((B * frames, 256, Cc), 2)
    mot = _rand_bf16((B, Cc), 3)
    ref = ho.splice(torch.from_numpy(ids), tok.float(), vit.float(), mot.float(), synth.IMG_CONTEXT_ID)
    out = llm_glue.splice_tokens(torch.from_numpy(ids).cuda(), tok.cuda(), vit.cuda(), mot.cuda(), synth.IMG_CONTEXT_ID)
    torch.cuda.synchronize()
    assert out.shape == (B, ids.shape[1], Cc) and out.dtype == torch.bfloat16
    assert torch.equal(out.float().cpu(), ref)   # pure data movement: bit exact


def test_splice_rejects_mismatched_visual_rows_and_bad_ids():
    ids = torch.from_numpy(synth.make_clip_prompt(2, seed=1)).cuda()
    tok, mot = _rand_bf16((512, 256), 1).cuda(), _rand_bf16((1, 256), 3).cuda()
    with pytest.raises(ValueError):
        llm_glue.splice_tokens(ids, tok, _rand_bf16((3, 256, 256), 2).cuda(), mot, synth.IMG_CONTEXT_ID)  # 3 frames for 2
    bad = ids.clone()
    bad[0, -1] = 9999
    with pytest.raises(ValueError):
        llm_glue.splice_tokens(bad, tok, _rand_bf16((2, 256, 256), 2).cuda(), mot, synth.IMG_CONTEXT_ID)


def test_mlpscore_head_matches_the_oracle():
    sd = synth.make_head_weights(seed=201)
    g = torch.Generator().manual_seed(5)
    hidden = torch.randn((5, 4096), generator=g)
    head = llm_glue.MlpScoreHead(sd, "cuda")
    score, pre = head(hidden.cuda())
    torch.cuda.synchronize()
    ref_s, ref_p = ho.mlpscore(hidden, ho.to_torch(sd))
    assert float((pre.cpu() - ref_p).abs().max()) < 1e-4 * float(ref_p.abs().max() + 1)
    assert float((score.cpu() - ref_s).abs().max()) < 1e-4 * float(ref_s.abs().max() + 1)
    assert (score >= 0).all()


def test_score_through_cuda_embeddings_cuda_splice_and_cuda_head(golden_dir):
    """The stage-2 forward with every piece the repo owns on the GPU: CUDA visual tokens -> CUDA splice -> (the reference's
    LLM arithmetic, oracle) -> CUDA mlpscore; score1 against the reference's own value (tolerance 0.02)."""
    import json, os
    from aigv_assessor_b200.config import VisionPathConfig
    from aigv_assessor_b200.engine import VisionEngine
    with open(os.path.join(golden_dir, "golden_meta_full.json")) as f:
        meta = json.load(f)["tiny"]
    g = np.load(os.path.join(golden_dir, "full_tiny.npz"))
    cfg = VisionPathConfig(**meta["vis_cfg"])
    sd_vis = synth.make_weights(cfg, seed=meta["seeds"]["vis"])
    sd_head = synth.make_head_weights(seed=meta["seeds"]["head"])
    sd_head["mlpscore.fc5.weight"] = synth.bf16_round(sd_head["mlpscore.fc5.weight"] * np.float32(meta["fc5_scale"]))
    pv = synth.make_pixel_values(meta["n_frames"], cfg, seed=meta["seeds"]["x"])
    ids = synth.make_clip_prompt(meta["n_frames"], seed=meta["seeds"]["ids"], vocab_size=meta["llm"]["vocab_size"])
    sdt = ho.to_torch(sd_head)
    eng = VisionEngine(cfg, max_tiles=8).load_state_dict(sd_vis)
    emb = eng.extract_feature(torch.from_numpy(pv).to(torch.bfloat16).cuda())
    motion = ho.motion_embeds(ho.fake_slowfast(torch.from_numpy(pv), 1), sdt).to(torch.bfloat16).cuda()
    tok = sdt["language_model.model.tok_embeddings.weight"].to(torch.bfloat16).cuda()   # weights are bf16-representable
    x = llm_glue.splice_tokens(torch.from_numpy(ids).cuda(), tok, emb, motion, synth.IMG_CONTEXT_ID)
    hidden, _ = ho.internlm2(x.float().cpu(), sdt, meta["llm"])
    score, pre = llm_glue.MlpScoreHead(sd_head, "cuda")(hidden[:, -4, :].cuda())
    d = abs(float(score[0]) - float(g["score1"][0]))
    print(f"all-CUDA pieces: score1 {float(score[0]):.4f} vs reference {float(g['score1'][0]):.4f} |d| {d:.2e}")
    assert d <= 0.02
