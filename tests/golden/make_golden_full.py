"""Full-model goldens for SCORE parity through the seam: run the UNMODIFIED reference stage-2 model on CPU.

    python tests/golden/make_golden_full.py          (build container only: needs /root/reference)

Builds the real `internvl_chat_eval2.InternVLChatModel`: the hard-coded
open('/DATA/.../config.json') of InternVLChatConfig (configuration_internvl_chat.py:42-46) is answered with the config
we want, `slowfast` (network download, eval2 :145) is replaced by oracle.head_oracle.fake_slowfast, a 1-rank gloo group
satisfies torch.distributed.get_rank() (eval2 :347). All weights are overwritten from seeds (aigv_assessor_b200.synth:
vision tower + mlp1 = make_weights, LLM + motion_mlp + mlpscore = make_head_weights), so nothing large is stored.
The forward is the reference's own (eval2 :306-488): extract_feature -> <IMG_CONTEXT> splice -> InternLM2 -> mlpscore.

Stored per case in tests/golden/full_<name>.npz: score1, its pre-ReLU value, hidden_states[-1][:, -4], argmax of the
shifted logits (tail), the last 48 rows of the shifted logits, vit_embeds (sub-sampled), plus in golden_meta_full.json
the sensitivity of score1 to a 1e-2 relative perturbation of vit_embeds (so the +-0.02 test is known not to be vacuous).
"""
from __future__ import annotations

import builtins
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (stubs, sys.path for the repo and /root/reference)

from aigv_assessor_b200 import config as cfgmod, synth  # noqa: E402
from oracle import head_oracle as ho  # noqa: E402

CFG_PATH = "/DATA/DATA1/wjr/intern/InternVL/internvl_chat/internvl/model/internvl_chat_stage2/config.json"


class FakeSlowFast(torch.nn.Module):
    """Bound in place of eval2.slowfast: the deterministic stand-in shared with the oracle."""

    def forward(self, x):  # x = [slow_pathway, fast_pathway], fast = all frames [B,3,T,H,W] (eval2 :103-120,337-343)
        fast = x[1]
        B = fast.shape[0]
        pv = fast.permute(0, 2, 1, 3, 4).reshape(-1, *fast.shape[1:2], *fast.shape[3:])
        return ho.fake_slowfast(pv, B).view(B, 2304, 1, 1, 1)


def build_full_model(vis, llm):
    vision_config = dict(hidden_size=vis.hidden_size, num_attention_heads=vis.num_heads,
                         intermediate_size=vis.intermediate_size, num_hidden_layers=vis.num_layers, image_size=448,
                         patch_size=14, qkv_bias=vis.qkv_bias, qk_normalization=vis.qk_normalization,
                         norm_type=vis.norm_type, layer_norm_eps=vis.layer_norm_eps,
                         initializer_factor=vis.initializer_factor, use_flash_attn=False, drop_path_rate=0.0,
                         hidden_act="gelu", architectures=["InternVisionModel"])
    llm_config = dict(architectures=["InternLM2ForCausalLM"], hidden_act="silu", bias=False, attn_implementation="eager",
                      pad_token_id=2, bos_token_id=1, eos_token_id=2, tie_word_embeddings=False,
                      rope_scaling={"factor": 2.0, "type": "dynamic"}, **llm)
    real_open = builtins.open

    def patched(path, *a, **k):
        if path == CFG_PATH:
            return io.StringIO(json.dumps(dict(llm_config=llm_config, vision_config=vision_config)))
        return real_open(path, *a, **k)

    builtins.open = patched
    try:
        from internvl.model.internvl_chat_eval2 import modeling_internvl_chat as E2
        from internvl.model.internvl_chat_eval2.configuration_internvl_chat import InternVLChatConfig
        E2.slowfast = FakeSlowFast
        with contextlib.redirect_stdout(io.StringIO()):  # the reference prints whole weight tensors at construction
            cfg = InternVLChatConfig(select_layer=-1, downsample_ratio=0.5, ps_version="v2", template="internlm2-chat",
                                     force_image_size=448)
            model = E2.InternVLChatModel(cfg, use_flash_attn=False).eval()
    finally:
        builtins.open = real_open
    return model


def load_weights(model, sd_vis, sd_head):
    sd = {k: torch.from_numpy(v) for k, v in {**sd_vis, **sd_head}.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("slowfast_model") for k in missing), missing


def golden_case(name, vis, n_frames, seeds):
    llm = synth.HEAD_LLM
    sd_vis = synth.make_weights(vis, seed=seeds["vis"])
    sd_head = synth.make_head_weights(seed=seeds["head"])
    pv = synth.make_pixel_values(n_frames, vis, seed=seeds["x"])
    ids = synth.make_clip_prompt(n_frames, seed=seeds["ids"], vocab_size=llm["vocab_size"])
    model = build_full_model(vis, llm)
    model.img_context_token_id = synth.IMG_CONTEXT_ID

    def run(sd_h, hook_embeds=None):
        load_weights(model, sd_vis, sd_h)
        cap = {}
        h1 = model.mlpscore.fc5.register_forward_hook(lambda m, i, o: cap.__setitem__("pre", o.detach().clone()))
        h2 = model.language_model.model.norm.register_forward_hook(lambda m, i, o: cap.__setitem__("hid", o.detach().clone()))
        orig = model.extract_feature

        def wrapped(pixel_values):
            e = orig(pixel_values)
            cap["vit"] = e.detach().clone()
            return e if hook_embeds is None else hook_embeds(e)

        model.extract_feature = wrapped
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            out = model(mos=torch.tensor([3.0]), pixel_values=torch.from_numpy(pv), input_ids=torch.from_numpy(ids),
                        attention_mask=torch.ones(ids.shape, dtype=torch.long),
                        image_flags=torch.ones(n_frames, 1, dtype=torch.long), labels=torch.from_numpy(ids))
        model.extract_feature = orig
        h1.remove(); h2.remove()
        return out, cap

    # fc5 is rescaled once so that score1 lands near 3 (MOS-like); the factor is recorded and re-applied by the tests
    out0, cap0 = run(sd_head)
    pre0 = float(cap0["pre"].reshape(-1)[0])
    assert pre0 > 0, "pre-ReLU score must be positive (positive fc5 weights)"
    fc5_scale = float(synth.bf16_round(np.float32(3.0 / pre0)))
    sd_head = dict(sd_head)
    sd_head["mlpscore.fc5.weight"] = synth.bf16_round(sd_head["mlpscore.fc5.weight"] * np.float32(fc5_scale))
    out, cap = run(sd_head)
    score = out["score1"].numpy()
    # sensitivity: relative 1e-2 gaussian perturbation of the visual tokens
    g = torch.Generator().manual_seed(1)
    outp, _ = run(sd_head, hook_embeds=lambda e: e + 1e-2 * e.abs().max() * torch.randn(e.shape, generator=g))
    sens = float(np.abs(outp["score1"].numpy() - score).max())
    # the oracle restatement fed with the reference's own embeddings (pins oracle/head_oracle.py)
    o = ho.stage2_forward(cap["vit"].numpy(), pv, ids, ho.to_torch(sd_head), llm, synth.IMG_CONTEXT_ID)
    # for calibration: the score the reference's OWN bf16 vision path (its production dtype, stage2_eval.py:932) gives
    ref16 = mg.build_reference(vis, sd_vis, dtype=torch.bfloat16)
    with torch.no_grad():
        e16 = ref16.extract_feature(torch.from_numpy(pv).to(torch.bfloat16)).float().numpy()
    o16 = ho.stage2_forward(e16, pv, ids, ho.to_torch(sd_head), llm, synth.IMG_CONTEXT_ID)
    d16 = abs(float(o16["score1"][0]) - float(score[0]))
    hid = cap["hid"][:, -4, :].numpy()
    print(name, "score1", score, "pre", cap["pre"].reshape(-1).numpy(), "| oracle", o["score1"], "| dscore @1e-2 noise", sens,
          "| reference bf16 vision path dscore", d16)
    logits_arg = out["logit"].numpy()
    np.savez_compressed(os.path.join(HERE, f"full_{name}.npz"), score1=score.astype(np.float32),
                        score_pre=cap["pre"].reshape(-1).numpy().astype(np.float32), hidden_m4=hid.astype(np.float32),
                        logit_tail=logits_arg[-48:].astype(np.int64), logits_tail=o["logits_tail"].astype(np.float32),
                        vit_embeds_s=cap["vit"].numpy()[:, ::8, ::16].astype(np.float32))
    assert abs(float(o["score1"][0]) - float(score[0])) < 1e-3 and np.array_equal(o["logit"], logits_arg)
    return dict(name=name, vis_cfg=vis.to_dict(), llm=llm, n_frames=n_frames, seeds=seeds, fc5_scale=fc5_scale,
                n_tokens=int(ids.shape[1]), score1=float(score[0]), score_sensitivity_1e2=sens,
                ref_bf16_vision_dscore=d16,
                oracle_score1=float(o["score1"][0]), vit_absmax=float(cap["vit"].abs().max()), torch=torch.__version__)


def main():
    mg.install_stubs()
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29541")
    dist.init_process_group("gloo", rank=0, world_size=1)
    metas = {}
    metas["tiny"] = golden_case("tiny", cfgmod.tiny(llm_hidden=4096), 8, dict(vis=200, head=201, x=202, ids=203))
    if "--skip-300m" not in sys.argv:
        metas["vit300m"] = golden_case("vit300m", cfgmod.internvit_300m(llm_hidden=4096), 8,
                                       dict(vis=0, head=211, x=212, ids=213))
    with open(os.path.join(HERE, "golden_meta_full.json"), "w") as f:
        json.dump(metas, f, indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
