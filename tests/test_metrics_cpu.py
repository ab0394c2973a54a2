"""Evaluation metrics (cm3p_b200/metrics.py) against the loop restatement of the reference's compute_metrics
(oracle/metrics_oracle.py) and, when the reference checkout is present, against the reference's own function."""
import os
import re

import pytest
import torch

from cm3p_b200 import metrics as M
from oracle import metrics_oracle as O


def _random_eval(seed, B=9, V=23, vocab=50, L=40):
    g = torch.Generator().manual_seed(seed)
    lpb = torch.randn(B, B, V, generator=g)
    classes = torch.randint(1, 5, (B, V), generator=g)
    classes[:, 0] = 0
    classes[:, V - 3:] = -1                       # padding variations never take part
    classes[1, 1:] = 2                            # an example with a single variation class
    classes[2, 1:V - 3] = classes[2, 1]           # another
    classes[3, :] = -1                            # only padding + original
    classes[3, 0] = 0
    lpb[torch.arange(B), torch.arange(B), 0] += 1.5  # the original usually, not always, wins
    logits = torch.randn(B, L, vocab, generator=g)
    labels = torch.randint(0, vocab, (B, L), generator=g)
    labels[torch.rand(B, L, generator=g) < 0.8] = -100
    return lpb, classes, logits, labels


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_vectorised_metrics_match_reference_loop(seed):
    lpb, classes, logits, labels = _random_eval(seed)
    for c in M.CLASSES_RANGE:
        want = O.variation_counts(lpb, classes, c, True)
        got = M.variation_accuracy(lpb[torch.arange(len(lpb)), torch.arange(len(lpb))], classes, c)
        assert got == want, (c, got, want)
    # accumulation over two batches + reset, names and None handling
    M.accumulated_metrics.clear()
    preds = (lpb, None, None, None, logits)
    assert M.compute_metrics(M.EvalPrediction(preds, labels, {"metadata_variation_classes": classes}), False) is None
    res = M.compute_metrics(M.EvalPrediction(preds, labels, {"metadata_variation_classes": classes}), True)
    c3 = O.variation_counts(lpb, classes, 3, True)
    assert res["accuracy_tags"] == pytest.approx(c3[0] / c3[1]) and res["top5_accuracy_tags"] == pytest.approx(c3[2] / c3[1])
    mlm = O.masked_lm_counts(logits, labels)
    assert res["accuracy_masked_lm"] == pytest.approx(mlm[0] / mlm[1])
    assert res["top5_accuracy_masked_lm"] == pytest.approx(mlm[2] / mlm[1])
    assert "top5_accuracy_year" not in res and "accuracy_year" in res
    assert M.accumulated_metrics == {}


def _reference_compute_metrics():
    """The reference's own function, exec'd from its source (train.py imports hydra, which is not installed)."""
    path = "/root/reference/train.py"
    if not os.path.isfile(path):
        return None
    src = open(path).read()
    m = re.search(r"^def compute_metrics\(.*?(?=^# noinspection PyArgumentList|^@hydra\.main)", src, flags=re.S | re.M)
    if not m:
        return None
    ns = {"torch": torch, "accumulated_metrics": {}, "EvalPrediction": M.EvalPrediction}
    exec(m.group(0), ns)
    return ns


@pytest.mark.parametrize("seed", [3, 4])
def test_metrics_match_the_reference_function(seed):
    ns = _reference_compute_metrics()
    if ns is None:
        pytest.skip("reference checkout not available (GPU box)")
    lpb, classes, logits, labels = _random_eval(seed, B=6, V=40)
    preds = (lpb, None, None, None, logits)
    ep = M.EvalPrediction(preds, labels, {"metadata_variation_classes": classes})
    want = ns["compute_metrics"](ep, True)
    M.accumulated_metrics.clear()
    got = M.compute_metrics(ep, True)
    assert set(got) == set(want)
    for k in want:
        assert got[k] == pytest.approx(want[k]), k
    # classification branch
    g = torch.Generator().manual_seed(seed)
    cl, lab = torch.randn(31, 2, generator=g), torch.randint(0, 2, (31,), generator=g)
    want = ns["compute_metrics"](M.EvalPrediction(cl, lab, {}), True)
    got = M.compute_metrics(M.EvalPrediction(cl, lab, {}), True)
    assert got == pytest.approx(want)
