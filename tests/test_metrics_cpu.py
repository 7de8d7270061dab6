"""Host-side result metrics (aigv_assessor_b200/metrics.py) against the reference's post-processing semantics."""
import numpy as np
import pytest

from aigv_assessor_b200 import metrics


def test_level_mapping_follows_the_reference_elif_order():
    # stage1_eval.py:944-955: 'bad' is tested first, then poor, fair, good, excellent; no match -> 0
    assert metrics.level_from_text("The quality of this video is excellent.") == 5
    assert metrics.level_from_text("good") == 4
    assert metrics.level_from_text("fair") == 3
    assert metrics.level_from_text("poor") == 2
    assert metrics.level_from_text("bad") == 1
    assert metrics.level_from_text("not bad, actually good") == 1   # first branch of the chain wins
    assert metrics.level_from_text("") == 0 and metrics.level_from_text("Excellent") == 0  # case sensitive, like `in`


def test_correlations_match_scipy_and_known_values():
    rng = np.random.default_rng(0)
    mos = rng.uniform(1, 5, 200)
    pred = mos + rng.normal(0, 0.7, 200)
    c = metrics.correlations(mos, pred)
    from scipy.stats import kendalltau, pearsonr, spearmanr
    assert c["srcc"] == pytest.approx(spearmanr(mos, pred)[0])
    assert c["plcc"] == pytest.approx(pearsonr(mos, pred)[0])
    assert c["krcc"] == pytest.approx(kendalltau(mos, pred)[0])
    perfect = metrics.correlations([1, 2, 3, 4], [10, 20, 30, 40])
    assert perfect == {"srcc": pytest.approx(1.0), "plcc": pytest.approx(1.0), "krcc": pytest.approx(1.0)}
    anti = metrics.correlations([1, 2, 3, 4], [4, 3, 2, 1])
    assert anti["srcc"] == pytest.approx(-1.0) and anti["krcc"] == pytest.approx(-1.0)


def test_accuracy_and_errors():
    assert metrics.accuracy(["the answer is good", "poor"], ["good", "fair"]) == 0.5
    with pytest.raises(ValueError):
        metrics.accuracy(["a"], [])
    with pytest.raises(ValueError):
        metrics.accuracy([], [])
    with pytest.raises(ValueError):
        metrics.correlations([1.0], [2.0])
    with pytest.raises(ValueError):
        metrics.correlations([1.0, 2.0], [2.0])
