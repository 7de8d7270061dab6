"""Pin the numpy oracle against fixtures produced by the UNMODIFIED reference (tests/golden/make_golden.py)."""
import hashlib
import os

import numpy as np
import pytest

from aigv_assessor_b200 import synth
from aigv_assessor_b200.config import VisionPathConfig
from oracle import preprocess_oracle as po
from oracle import vit_oracle as vo


def _case(golden_dir, golden_meta, name):
    meta = golden_meta[name]
    cfg = VisionPathConfig(**meta["cfg"])
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    sd = synth.make_weights(cfg, seed=meta["w_seed"])
    pv = synth.make_pixel_values(meta["n_frames"], cfg, seed=meta["x_seed"])
    return cfg, g, sd, pv


def test_tiny_tower_matches_reference(golden_dir, golden_meta):
    cfg, g, sd, pv = _case(golden_dir, golden_meta, "tiny")
    out, st = vo.extract_feature(pv, sd, cfg, return_stages=True)
    # fp32 vs fp32: only summation-order noise is allowed
    assert vo.error_metrics(st["embeddings"][:, ::5], g["embeddings_s"])[0] < 1e-5
    assert vo.error_metrics(st["layer0"][:, ::5], g["layer0_s"])[0] < 1e-5
    assert vo.error_metrics(st["layer1"], g["vit_out"])[0] < 1e-5
    assert out.shape == (2, 256, cfg.llm_hidden)
    assert vo.error_metrics(out, g["out"])[0] < 1e-5


def test_tiny_rms_qknorm_tower_matches_reference(golden_dir, golden_meta):
    cfg, g, sd, pv = _case(golden_dir, golden_meta, "tiny_rms")
    assert cfg.norm_type == "rms_norm" and cfg.qk_normalization and not cfg.qkv_bias
    out, st = vo.extract_feature(pv, sd, cfg, return_stages=True)
    assert vo.error_metrics(st["embeddings"][:, ::41, ::16], g["embeddings_s"])[0] < 1e-5
    assert vo.error_metrics(st["layer1"][:, ::41, ::16], g["vit_out_s"])[0] < 1e-5
    assert vo.error_metrics(out[:, ::8, ::16], g["out_s"])[0] < 1e-5


@pytest.mark.slow
def test_vit300m_one_frame_matches_reference(golden_dir, golden_meta):
    cfg, g, sd, pv = _case(golden_dir, golden_meta, "vit300m")
    out, st = vo.extract_feature(pv, sd, cfg, return_stages=True)
    assert vo.error_metrics(st["layer0"][:, ::41, ::16], g["layer0_s"])[0] < 1e-5
    assert vo.error_metrics(st["layer23"][:, ::41, ::16], g["vit_out_s"])[0] < 2e-5
    assert vo.error_metrics(out[:, ::8, ::16], g["out_s"])[0] < 2e-5
    assert abs(np.abs(out).max() - golden_meta["vit300m"]["out_absmax"]) < 1e-4


def test_pixel_shuffle_index_law():
    # out[n,I,J,(2di+dj)C+c] = in[n,2I+di,2J+dj,c] for ps_version v2
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 8, 8, 5)).astype(np.float32)
    y = vo.pixel_shuffle(x, 0.5, "v2")
    assert y.shape == (2, 4, 4, 20)
    for (I, J, di, dj, c) in [(0, 0, 0, 0, 0), (1, 2, 1, 0, 3), (3, 3, 1, 1, 4), (2, 0, 0, 1, 2)]:
        assert np.array_equal(y[:, I, J, (2 * di + dj) * 5 + c], x[:, 2 * I + di, 2 * J + dj, c])


def test_preprocess_matches_reference(golden_meta):
    for case in golden_meta["preprocess"]["cases"]:
        frame = synth.make_frames(1, case["h"], case["w"], seed=case["seed"], kind=case["kind"])
        tiles, pv, counts = po.load_video_frames(frame, max_num=case["max_num"], use_thumbnail=True)
        assert counts == [case["n_tiles"]]
        assert hashlib.sha256(tiles.tobytes()).hexdigest() == case["tiles_sha256"], case
        assert tiles[0, 100:104, 200:204].tolist() == case["tile0_crop"]
        assert hashlib.sha256(np.ascontiguousarray(pv).tobytes()).hexdigest() == case["pv_sha256"], case
        assert hashlib.sha256(synth.to_bf16_bits(pv).tobytes()).hexdigest() == case["pv_bf16_sha256"]


def test_resize_matches_installed_pillow():
    from PIL import Image
    for (h, w, ow, oh, kind) in [(720, 1280, 448, 448, "uniform"), (300, 200, 896, 448, "smooth"),
                                 (512, 768, 1344, 896, "uniform"), (449, 447, 448, 448, "uniform")]:
        f = synth.make_frames(1, h, w, seed=7, kind=kind)[0]
        a = po.resize_bicubic_u8(f, ow, oh)
        b = np.asarray(Image.fromarray(f).resize((ow, oh)))
        assert np.array_equal(a, b), (h, w, ow, oh)


def test_normalize_lut_is_the_transform():
    lut = synth.normalize_lut()
    tile = np.arange(256, dtype=np.uint8).reshape(1, 256, 1).repeat(3, axis=2)
    pv = po.transform(tile)
    for c in range(3):
        assert np.array_equal(pv[c, 0], lut[c])


@pytest.mark.parametrize("name", ["tiny", "tiny_rms"])
def test_torch_oracle_matches_reference_goldens(golden_dir, golden_meta, name):
    import torch
    from oracle import vit_oracle_torch as vt
    cfg, g, sd, pv = _case(golden_dir, golden_meta, name)
    out, vit = vt.extract_feature(torch.from_numpy(pv), vt.to_torch(sd), cfg, return_vit=True)
    if name == "tiny":
        assert vo.error_metrics(vit.numpy(), g["vit_out"])[0] < 1e-5
        assert vo.error_metrics(out.numpy(), g["out"])[0] < 1e-5
    else:
        assert vo.error_metrics(vit.numpy()[:, ::41, ::16], g["vit_out_s"])[0] < 1e-5
        assert vo.error_metrics(out.numpy()[:, ::8, ::16], g["out_s"])[0] < 1e-5


def _full_case(golden_dir, name):
    import json
    with open(os.path.join(golden_dir, "golden_meta_full.json")) as f:
        meta = json.load(f)[name]
    g = np.load(os.path.join(golden_dir, f"full_{name}.npz"))
    cfg = VisionPathConfig(**meta["vis_cfg"])
    sd_vis = synth.make_weights(cfg, seed=meta["seeds"]["vis"])
    sd_head = synth.make_head_weights(seed=meta["seeds"]["head"])
    sd_head["mlpscore.fc5.weight"] = synth.bf16_round(sd_head["mlpscore.fc5.weight"] * np.float32(meta["fc5_scale"]))
    pv = synth.make_pixel_values(meta["n_frames"], cfg, seed=meta["seeds"]["x"])
    ids = synth.make_clip_prompt(meta["n_frames"], seed=meta["seeds"]["ids"], vocab_size=meta["llm"]["vocab_size"])
    return meta, g, cfg, sd_vis, sd_head, pv, ids


def test_full_model_score_oracle_matches_reference(golden_dir):
    """vit oracle -> head oracle reproduces the UNMODIFIED reference stage-2 forward (tests/golden/make_golden_full.py:
    the real internvl_chat_eval2.InternVLChatModel on CPU): score1, its pre-ReLU value, hidden_states[-1][:, -4] and the
    argmax of the shifted logits for one 8-frame clip."""
    from oracle import head_oracle as ho
    meta, g, cfg, sd_vis, sd_head, pv, ids = _full_case(golden_dir, "tiny")
    emb = vo.extract_feature(pv, sd_vis, cfg)
    assert vo.error_metrics(emb[:, ::8, ::16], g["vit_embeds_s"])[0] < 1e-5
    o = ho.stage2_forward(emb, pv, ids, ho.to_torch(sd_head), meta["llm"], synth.IMG_CONTEXT_ID)
    assert ids.shape[1] == meta["n_tokens"] and (ids == synth.IMG_CONTEXT_ID).sum() == 8 * 256 + 1
    assert abs(float(o["score1"][0]) - float(g["score1"][0])) < 1e-3 and float(g["score1"][0]) > 1.0
    assert abs(float(o["score_pre"][0]) - float(g["score_pre"][0])) < 1e-3
    assert vo.error_metrics(o["hidden_m4"], g["hidden_m4"])[0] < 1e-4
    assert np.array_equal(o["logit"][-48:], g["logit_tail"])
    # the +-0.02 score tolerance is not vacuous: on the InternViT-300M case a 1e-2 (of max) gaussian perturbation of the
    # tokens moves score1 by more than the tolerance, while the reference's own bf16 vision path stays inside it
    import json
    with open(os.path.join(golden_dir, "golden_meta_full.json")) as f:
        m300 = json.load(f)["vit300m"]
    assert m300["score_sensitivity_1e2"] > 0.02 > m300["ref_bf16_vision_dscore"]


@pytest.mark.slow
def test_vit6b_width_one_layer_matches_reference(golden_dir, golden_meta):
    """InternViT-6B geometry at its real width (D=3200, 25 heads x 128, I=12800, RMSNorm + q/k RMSNorm, ls 0.1, 12800->6144)."""
    cfg, g, sd, pv = _case(golden_dir, golden_meta, "vit6b_l1")
    assert (cfg.hidden_size, cfg.num_heads, cfg.intermediate_size, cfg.llm_hidden) == (3200, 25, 12800, 6144)
    out, st = vo.extract_feature(pv, sd, cfg, return_stages=True)
    assert vo.error_metrics(st["layer0"][:, ::41, ::16], g["vit_out_s"])[0] < 2e-5
    assert vo.error_metrics(out[:, ::8, ::16], g["out_s"])[0] < 2e-5


def test_tiny_stress_tower_matches_reference(golden_dir, golden_meta):
    """Trained-like statistics (peaked attention, +-300 outlier channels, layer scales 0.1..1) through the oracle."""
    meta = golden_meta["tiny_stress"]
    cfg = VisionPathConfig(**meta["cfg"])
    g = np.load(os.path.join(golden_dir, "tiny_stress.npz"))
    sd = synth.make_stress_weights(cfg, seed=meta["w_seed"])
    pv = synth.make_pixel_values(meta["n_frames"], cfg, seed=meta["x_seed"])
    out, st = vo.extract_feature(pv, sd, cfg, return_stages=True)
    assert meta["vit_absmax"] > 250.0
    assert vo.error_metrics(st["layer1"][:, ::41, ::16], g["vit_out_s"])[0] < 1e-4
    assert np.abs(out[:, ::8, ::16] - g["out_s"]).max() / meta["out_absmax"] < 1e-4


@pytest.mark.slow
def test_torch_oracle_the_timed_cpu_baseline_matches_reference_on_vit300m(golden_dir, golden_meta):
    """oracle/vit_oracle_torch.py is what bench.py times as the CPU baseline: pin it on the full-size golden too."""
    import torch
    from oracle import vit_oracle_torch as vt
    cfg, g, sd, pv = _case(golden_dir, golden_meta, "vit300m")
    out, vit = vt.extract_feature(torch.from_numpy(pv), vt.to_torch(sd), cfg, return_vit=True)
    assert vo.error_metrics(vit.numpy()[:, ::41, ::16], g["vit_out_s"])[0] < 2e-5
    assert vo.error_metrics(out.numpy()[:, ::8, ::16], g["out_s"])[0] < 2e-5
