"""Loop restatement of the reference's evaluation metrics.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/train.py:38-160 (`compute_metrics`) statement by statement — per example, per variation
class, Python scalars — so that `cm3p_b200/metrics.py` (vectorised) can be checked against it on random inputs
(tests/test_metrics_cpu.py).  Parity status: the reference's function needs `transformers.EvalPrediction` and a
module-global accumulator only; `tests/test_metrics_cpu.py` also runs the reference's own function (imported from
/root/reference/train.py when hydra is importable; it is not in this image, so the function body is exec'd from
the source file without the hydra decorator) against both.
"""
from __future__ import annotations

import torch


def variation_counts(logits_per_beatmap, metadata_variation_classes, var_class, with_top5):
    """train.py:105-133 for one variation class -> (correct, total, top5_correct)."""
    correct = total = top5_correct = 0
    batch_size = logits_per_beatmap.shape[0]
    for i in range(batch_size):
        class_mask = (metadata_variation_classes[i] == var_class) | (metadata_variation_classes[i] == 0)  # :109-110
        if class_mask.sum() <= 1:  # :112
            continue
        group_logits = logits_per_beatmap[i, i][class_mask]  # :116
        group_classes = metadata_variation_classes[i][class_mask]  # :117
        total += 1  # :120
        predicted_index = torch.argmax(group_logits).item()  # :122
        if group_classes[predicted_index] == 0:  # :123
            correct += 1
        if with_top5:  # :126-129
            top5_indices = torch.topk(group_logits, k=min(5, group_logits.size(0))).indices
            if (group_classes[top5_indices] == 0).any():
                top5_correct += 1
    return correct, total, top5_correct


def masked_lm_counts(logits, labels):
    """train.py:78-91 -> (correct, total, top5_correct)."""
    mask = labels != -100
    correct = (logits.argmax(-1)[mask] == labels[mask]).sum().item()
    total = mask.sum().item()
    top5_indices = torch.topk(logits, k=min(5, logits.size(-1)), dim=-1).indices
    top5_correct = (top5_indices[mask] == labels[mask].unsqueeze(-1)).any(dim=-1).sum().item()
    return correct, total, top5_correct
