"""Result metrics of the sharded eval loop: host logic only, no GPU work.

Mirrors the reference's own post-processing so a multi-GPU run reports the numbers the single-GPU script prints:
  * level_from_text  -- the quality word -> level mapping of the stage-1 eval loop
                        (internvl/train/internvl/eval/stage1_eval.py:944-955: first match of bad/poor/fair/good/excellent
                        in that order, 0 when none);
  * correlations     -- SRCC / PLCC / KRCC exactly as save_and_evaluate computes them (stage1_eval.py:652-678:
                        scipy.stats spearmanr, pearsonr, kendalltau on (mos, prediction)) plus its substring accuracy.
Typical use with the clip sharding of sharding.py:
    scores = sharding.gather_clip_results(local_scores, n_clips)      # one NCCL all_gather
    if rank == 0: print(metrics.correlations(mos, scores.tolist()))
"""
from __future__ import annotations

from typing import Iterable, Sequence

_LEVELS = (("bad", 1), ("poor", 2), ("fair", 3), ("good", 4), ("excellent", 5))


def level_from_text(output: str) -> int:
    """stage1_eval.py:944-955 (an if/elif chain: the FIRST word of this order that occurs anywhere in the text wins)."""
    for word, level in _LEVELS:
        if word in output:
            return level
    return 0


def accuracy(answers: Sequence[str], outputs: Sequence[str]) -> float:
    """stage1_eval.py:661-665: a sample counts as right when the decoded output is a substring of the answer."""
    if len(answers) != len(outputs):
        raise ValueError("answers and outputs differ in length")
    if not answers:
        raise ValueError("empty result list")
    return sum(1 for a, o in zip(answers, outputs) if o in a) / len(answers)


def correlations(mos: Iterable[float], predicted: Iterable[float]) -> dict:
    """SRCC, PLCC, KRCC of predictions against mean opinion scores (stage1_eval.py:666-670)."""
    from scipy.stats import kendalltau, pearsonr, spearmanr  # the reference's own dependency for this step

    m, p = list(mos), list(predicted)
    if len(m) != len(p):
        raise ValueError("mos and predictions differ in length")
    if len(m) < 2:
        raise ValueError("need at least two samples")
    return {"srcc": float(spearmanr(m, p)[0]), "plcc": float(pearsonr(m, p)[0]), "krcc": float(kendalltau(m, p)[0])}
