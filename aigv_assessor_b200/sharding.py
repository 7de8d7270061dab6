"""Clip-granular data parallelism: the path shards with no data-path collective (every tile is independent,
SURVEY.md section 8e); the only exchange is one all-gather of per-clip results at the end.

The reference's eval loop is unsharded (stage1_eval.py:898-901: plain DataLoader, every rank would redo the whole
set); this module is the new functionality that replaces it for multi-GPU runs. Works with NCCL (GPU) and gloo (CPU
tests).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership: rank r takes clips r, r+W, ... (equal frames per clip => balanced)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_items, world))


def gather_clip_results(local: torch.Tensor, n_items: int, rank: int | None = None, world: int | None = None,
                        group=None) -> torch.Tensor:
    """local: [len(shard_indices(n_items, rank, world)), ...] results for this rank's clips (scores or features).
    Returns [n_items, ...] in original clip order on every rank. Uneven tails are padded for the collective."""
    if not (dist.is_available() and dist.is_initialized()):
        if local.shape[0] != n_items:
            raise ValueError("single-process gather expects all items")
        return local
    world = dist.get_world_size(group) if world is None else world
    rank = dist.get_rank(group) if rank is None else rank
    per = (n_items + world - 1) // world
    mine = len(shard_indices(n_items, rank, world))
    if local.shape[0] != mine:
        raise ValueError(f"rank {rank} holds {local.shape[0]} items, expected {mine}")
    padded = local.new_zeros((per,) + tuple(local.shape[1:]))
    padded[:mine] = local
    out = local.new_empty((world * per,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    out = out.view((world, per) + tuple(local.shape[1:]))
    # rank r slot k  ->  clip r + k*world
    full = out.transpose(0, 1).reshape((per * world,) + tuple(local.shape[1:]))
    return full[:n_items].contiguous()


def encode_clips_sharded(engine, clips: Sequence[torch.Tensor], pool: bool = False) -> torch.Tensor:
    """clips: list of pixel_values [frames,3,448,448] (same frame count). Each rank encodes its clips and the
    visual tokens (or their per-clip mean when pool=True) are all-gathered. Returns [n_clips, frames*256, C] (or
    [n_clips, C])."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    mine = shard_indices(len(clips), rank, world)
    outs = []
    for i in mine:
        y = engine.extract_feature(clips[i])
        y = y.reshape(-1, y.shape[-1])
        outs.append(y.float().mean(0) if pool else y)
    if outs:
        local = torch.stack(outs)
    else:
        c = engine.cfg.llm_hidden
        f = clips[0].shape[0] * engine.cfg.out_tokens if clips else 0
        local = torch.empty((0, c) if pool else (0, f, c), device=engine.device,
                            dtype=torch.float32 if pool else torch.bfloat16)
    return gather_clip_results(local, len(clips), rank, world)


def score_clips_sharded(scorer, clips, input_ids, mos=None, motion_features=None, clips_per_forward: int = 4) -> dict:
    """The sharded counterpart of the stage-2 eval loop (stage2_eval.py:915-975: one forward per video, score1 collected,
    SRCC / PLCC / KRCC printed at the end): every rank scores its round-robin share of ``clips`` with ``scorer``
    (scoring.ClipScorer), the scores are all-gathered once, and -- when ``mos`` is given -- the reference's metrics are computed.

    clips: list of pixel_values [frames, 3, 448, 448] (or uint8 frames [frames, H, W, 3]); input_ids: one [1, N] prompt for all
    clips or a list of them; motion_features: optional list of [1, 2304] (else the scorer's own trunk runs).
    clips_per_forward: consecutive clips of a rank with the same frame count and prompt length go through ONE forward (the
    prefill pads every call to whole 256-row tiles: 2108 tokens alone waste 8.5 % of it, four prompts together 0.2 %; measured
    29.4 against 26.7 clips/s with InternLM2-7B). Per-clip results do not depend on the grouping beyond bf16 rounding order."""
    from . import metrics
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    mine = shard_indices(len(clips), rank, world)

    def prep(i):
        ids = input_ids[i] if isinstance(input_ids, (list, tuple)) else input_ids
        clip = clips[i]
        if clip.dtype == torch.uint8:
            clip, _ = scorer.vision.preprocess_frames(clip, max_num=1)
        mf = None if motion_features is None else motion_features[i]
        return clip, ids.reshape(1, -1), mf

    local, group = [], []

    def flush():
        if not group:
            return
        pv = torch.cat([g[0] for g in group]) if len(group) > 1 else group[0][0]
        ids = torch.cat([g[1] for g in group]) if len(group) > 1 else group[0][1]
        mf = None if group[0][2] is None else torch.cat([g[2].reshape(1, -1) for g in group])
        local.append(scorer(pv, ids, mf)["score1"].reshape(-1).float())
        group.clear()

    cap = max(1, int(clips_per_forward))
    for i in mine:
        item = prep(i)
        if group and (item[0].shape != group[0][0].shape or item[1].shape != group[0][1].shape
                      or (len(group) + 1) * item[0].shape[0] > scorer.vision.max_tiles
                      or (len(group) + 1) * item[1].shape[1] > scorer.language.max_tokens):
            flush()
        group.append(item)
        if len(group) >= cap:
            flush()
    flush()
    local_t = torch.cat(local) if local else torch.empty((0,), dtype=torch.float32, device=scorer.device)
    scores = gather_clip_results(local_t, len(clips), rank, world)
    out = {"scores": scores}
    if mos is not None and len(clips) >= 2:
        out.update(metrics.correlations(list(mos), scores.cpu().tolist()))
    return out
