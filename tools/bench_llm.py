"""GPU box: time the InternLM2 prefill (aigv_llm_prefill) on random-init weights generated on the device.
Usage: python tools/bench_llm.py [7b|1_8b] [N] [B ...]   -> one JSON line per batch size."""
import _path  # noqa: F401  (repository root on sys.path)
import json
import sys
import time

import torch

from aigv_assessor_b200 import llm as L

CFGS = {
    # InternLM2.5-7B (InternVL2-8B, the AIGV-Assessor checkpoint: mlpscore / motion_mlp are 4096 wide)
    "7b": dict(hidden_size=4096, num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=8, intermediate_size=14336,
               vocab_size=92553, rms_norm_eps=1e-5, rope_theta=1e6, max_position_embeddings=32768),
    # InternLM2-1.8B (InternVL2-2B, BASELINE configs[1])
    "1_8b": dict(hidden_size=2048, num_hidden_layers=24, num_attention_heads=16, num_key_value_heads=8, intermediate_size=8192,
                 vocab_size=92553, rms_norm_eps=1e-5, rope_theta=1e6, max_position_embeddings=32768),
}


def flops_per_prompt(c, N):
    H, I, L_ = c["hidden_size"], c["intermediate_size"], c["num_hidden_layers"]
    hd = H // c["num_attention_heads"]
    qw = (c["num_attention_heads"] + 2 * c["num_key_value_heads"]) * hd
    gemm = 2.0 * N * H * (qw + H + 2 * I + I) * L_
    attn = 2.0 * 2.0 * (N * (N + 1) / 2) * H * L_      # causal: QK^T and PV over the lower triangle
    return gemm + attn


def build(name, max_tokens):
    c = CFGS[name]
    eng = L.LanguageEngine(c, max_tokens=max_tokens)
    H, I = c["hidden_size"], c["intermediate_size"]
    hd = H // c["num_attention_heads"]
    qw = (c["num_attention_heads"] + 2 * c["num_key_value_heads"]) * hd
    g = torch.Generator(device="cuda").manual_seed(0)

    def rnd(shape, fan_in):
        b = fan_in ** -0.5
        return ((torch.rand(shape, device="cuda", generator=g) * 2 - 1) * b).to(torch.bfloat16)

    for i in range(c["num_hidden_layers"]):
        p = f"model.layers.{i}."
        eng.set_weight(p + "attention.wqkv.weight", rnd((qw, H), H))
        eng.set_weight(p + "attention.wo.weight", rnd((H, H), H))
        eng.set_weight(p + "feed_forward.w1.weight", rnd((I, H), H))
        eng.set_weight(p + "feed_forward.w3.weight", rnd((I, H), H))
        eng.set_weight(p + "feed_forward.w2.weight", rnd((H, I), I))
        eng.set_weight(p + "attention_norm.weight", torch.ones(H, device="cuda"))
        eng.set_weight(p + "ffn_norm.weight", torch.ones(H, device="cuda"))
    eng.set_weight("model.norm.weight", torch.ones(H, device="cuda"))
    eng.set_weight("output.weight", rnd((c["vocab_size"], H), H))
    st = eng.lib.aigv_llm_finalize_weights(eng._h, None)
    eng._check(st, "finalize")
    eng._finalized = True
    return c, eng


def reference_gpu_stack(c, B, N, reps=3):
    """The reference's own GPU path for the same prefill -- InternLM2FlashAttention2 (modeling_internlm2.py:444-620) + eager
    bf16 nn.Linear / RMSNorm / rotary / SwiGLU ops -- restated with library kernels (cuBLASLt via torch.matmul, the flash-attn
    wheel's causal kernel, ATen elementwise) on random weights of the same shapes. One layer's weights are reused for every
    layer (the timing does not depend on the values; 32 distinct layers would double the memory). A comparator only."""
    import torch.nn.functional as F
    from flash_attn import flash_attn_func
    H, I, nl = c["hidden_size"], c["intermediate_size"], c["num_hidden_layers"]
    nh, nkv = c["num_attention_heads"], c["num_key_value_heads"]
    hd, G = H // nh, nh // nkv
    dev = "cuda"
    bf = torch.bfloat16
    w = {k: (torch.randn(shape, device=dev) * 0.02).to(bf) for k, shape in
         dict(wqkv=((nh + 2 * nkv) * hd, H), wo=(H, H), w1=(I, H), w3=(I, H), w2=(H, I)).items()}
    g1, g2 = torch.ones(H, device=dev, dtype=bf), torch.ones(H, device=dev, dtype=bf)
    inv = 1.0 / (c["rope_theta"] ** (torch.arange(0, hd, 2, device=dev).float() / hd))
    fr = torch.outer(torch.arange(N, device=dev).float(), inv)
    emb = torch.cat((fr, fr), -1)
    cos, sin = emb.cos().to(bf)[None, :, None, :], emb.sin().to(bf)[None, :, None, :]

    def rms(x, g):
        v = x.float().pow(2).mean(-1, keepdim=True)
        return g * (x.float() * torch.rsqrt(v + c["rms_norm_eps"])).to(bf)

    def rot(x):
        return torch.cat((-x[..., hd // 2:], x[..., :hd // 2]), -1)

    def layer(x):
        h = rms(x, g1)
        qkv = F.linear(h, w["wqkv"]).view(B, N, nkv, G + 2, hd)
        q = qkv[..., :G, :].reshape(B, N, nh, hd)
        k, v = qkv[..., -2, :], qkv[..., -1, :]
        q = q * cos + rot(q) * sin
        k = k * cos + rot(k) * sin
        a = flash_attn_func(q, k, v, causal=True).reshape(B, N, H)
        x = x + F.linear(a, w["wo"])
        h = rms(x, g2)
        return x + F.linear(F.silu(F.linear(h, w["w1"])) * F.linear(h, w["w3"]), w["w2"])

    x0 = (torch.randn((B, N, H), device=dev) * 0.5).to(bf)
    with torch.no_grad():
        x = x0
        for _ in range(2):
            x = layer(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            x = x0
            for _ in range(nl):
                x = layer(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"ms_per_prefill": ms, "tflops": flops_per_prompt(c, N) * B / ms / 1e9,
            "stack": "torch.matmul (cuBLASLt) + flash-attn 2 causal + ATen elementwise, eager bf16"}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "7b"
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 2108
    noref = "noref" in sys.argv
    Bs = [int(v) for v in sys.argv[3:] if v != "noref"] or [1, 2, 4]
    t0 = time.time()
    c, eng = build(name, max(Bs) * N)
    torch.cuda.synchronize()
    print(f"# {name}: weights on the device in {time.time() - t0:.1f} s", file=sys.stderr)
    for B in Bs:
        x = (torch.randn((B, N, c["hidden_size"]), device="cuda") * 0.5).to(torch.bfloat16)
        rows = [b * N + N - 4 for b in range(B)]
        for _ in range(2):
            eng.prefill(x)
            hid = eng.hidden_rows(rows)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            eng.prefill(x)
            hid = eng.hidden_rows(rows)
            lg = eng.logits_rows([b * N + p for b in range(B) for p in range(N - 49, N - 1)][-48:])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = flops_per_prompt(c, N) * B
        ref = None
        if B == Bs[0] and not noref:
            try:
                ref = reference_gpu_stack(c, B, N)
            except Exception as e:  # comparator only
                ref = {"unavailable": repr(e)[:200]}
        print(json.dumps({"model": name, "B": B, "N": N, "reference_gpu_stack": ref, "ms_per_prefill": ms, "ms_per_prompt": ms / B, "tokens_per_s": B * N / ms * 1e3,
                          "tflops": fl / ms / 1e9, "tflop_per_prompt": fl / B / 1e12, "finite": bool(torch.isfinite(hid).all() and torch.isfinite(lg).all())}),
              flush=True)


if __name__ == "__main__":
    main()
