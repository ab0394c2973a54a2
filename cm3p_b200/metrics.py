"""Evaluation metrics of the training entry (reference: /root/reference/train.py:38-160 `compute_metrics`).

Same interface (`compute_metrics(eval_pred, compute_result)` with `batch_eval_metrics` semantics: called once per
evaluation batch, accumulates, returns the dict and resets when `compute_result` is true), same metric names
(`accuracy_<class>`, `top5_accuracy_<class>` for masked_lm / tags / mapper / classification) and the same
definitions:

* zero-shot variation accuracy: for every example i and variation class c in {1 year, 2 status, 3 tags, 4 mapper},
  among the metadata variations of example i whose class is c or 0 (the original), is the highest
  `logits_per_beatmap[i, i, :]` the original?  Examples without a variation of that class are skipped; the class
  -1 padding variations never take part (train.py:92-139).
* masked-LM accuracy / top-5 over the positions whose label is not -100 (train.py:78-91), classification accuracy /
  top-5 for 1-D labels (train.py:61-77).

What differs from the reference is only the execution: no Python loop over examples and classes — the groups are
boolean masks over the (B, V) diagonal of the logits, evaluated with a handful of tensor ops on whatever device the
logits live on (at the reference's 1000 test variations per beatmap, `configs/train/default.yaml:147`, the loop is
4 * B host round trips per batch).  Integer / comparison work only.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Optional

import torch

VARIATION_CLASSES = {-200: "classification", -100: "masked_lm", -1: "padding", 0: "original", 1: "year", 2: "status",
                     3: "tags", 4: "mapper"}
CLASSES_RANGE = range(1, 5)
CLASSES_WITH_TOP5 = (-100, 3, 4)


@dataclass
class EvalPrediction:
    """Duck-type of transformers.EvalPrediction (predictions, label_ids, inputs)."""
    predictions: Any
    label_ids: Optional[torch.Tensor] = None
    inputs: dict = field(default_factory=dict)


accumulated_metrics: dict = {}


def _add(var_class: int, correct: int, total: int, top5: int) -> None:
    m = accumulated_metrics.setdefault(var_class, {"correct": 0, "total": 0, "top5_correct": 0})
    m["correct"] += int(correct)
    m["total"] += int(total)
    m["top5_correct"] += int(top5)


def variation_accuracy(diag_logits: torch.Tensor, classes: torch.Tensor, var_class: int):
    """diag_logits [B, V] = logits_per_beatmap[i, i, :], classes [B, V] -> (correct, total, top5_correct)."""
    mask = (classes == var_class) | (classes == 0)
    count = mask.sum(dim=1)
    valid = count > 1
    neg = torch.finfo(diag_logits.dtype).min
    masked = torch.where(mask, diag_logits, torch.full_like(diag_logits, neg))
    pred = masked.argmax(dim=1, keepdim=True)
    correct = ((classes.gather(1, pred).squeeze(1) == 0) & valid).sum()
    k = min(5, diag_logits.shape[1])
    top = masked.topk(k, dim=1)
    in_group = mask.gather(1, top.indices)  # masked-out entries can only enter the top-k of groups smaller than k
    top5 = (((classes.gather(1, top.indices) == 0) & in_group).any(dim=1) & valid).sum()
    return int(correct), int(valid.sum()), int(top5)


def compute_metrics(eval_pred, compute_result: bool) -> Optional[dict]:
    global accumulated_metrics
    labels = getattr(eval_pred, "label_ids", None)
    preds = eval_pred.predictions
    if labels is not None and len(labels) > 0:
        labels = torch.as_tensor(labels)
        if labels.ndim == 1:  # classification (train.py:61-77)
            logits = torch.as_tensor(preds[0] if isinstance(preds, tuple) else preds)
            correct = (logits.argmax(-1) == labels).sum()
            top5 = (logits.topk(min(5, logits.size(-1)), dim=-1).indices == labels.unsqueeze(-1)).any(dim=-1).sum()
            _add(-200, correct, labels.size(0), top5)
        else:  # masked LM (train.py:78-91): predictions[4] = CM3POutput.logits
            logits = torch.as_tensor(preds[4] if isinstance(preds, tuple) else preds)
            mask = labels != -100
            sel, tgt = logits[mask], labels[mask]
            correct = (sel.argmax(-1) == tgt).sum()
            top5 = (sel.topk(min(5, sel.size(-1)), dim=-1).indices == tgt.unsqueeze(-1)).any(dim=-1).sum()
            _add(-100, correct, mask.sum(), top5)
    inputs = getattr(eval_pred, "inputs", None) or {}
    if "metadata_variation_classes" in inputs:
        lpb = torch.as_tensor(preds[0])  # logits_per_beatmap (B, B, V): the diagonal pairs every beatmap with its own metadata
        classes = torch.as_tensor(inputs["metadata_variation_classes"]).to(lpb.device)
        B = lpb.shape[0]
        idx = torch.arange(B, device=lpb.device)
        diag = lpb[idx, idx].float()
        for var_class in CLASSES_RANGE:
            c, t, t5 = variation_accuracy(diag, classes, var_class)
            _add(var_class, c, t, t5 if var_class in CLASSES_WITH_TOP5 else 0)
    if not compute_result:
        return None
    result = {}
    for var_class, m in accumulated_metrics.items():
        name = VARIATION_CLASSES.get(var_class, f"class_{var_class}")
        result[f"accuracy_{name}"] = m["correct"] / m["total"] if m["total"] > 0 else None
        if var_class in CLASSES_WITH_TOP5:
            result[f"top5_accuracy_{name}"] = m["top5_correct"] / m["total"] if m["total"] > 0 else None
    accumulated_metrics = {}
    return result
