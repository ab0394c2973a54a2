"""Extraction pipeline around the model (extract_beatmap_embeddings.py:217-266) and zero-shot scoring with many
metadata variations per beatmap (train.py:92-139; configs/train/default.yaml eval: 1000 variations)."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from cm3p_b200.configuration_cm3p import CM3PConfig, small_config_dict
from cm3p_b200.synthetic import synthetic_batch, synthetic_state_dict


def _model():
    from cm3p_b200.modeling_cm3p import CM3PModel
    cfg = CM3PConfig(**copy.deepcopy(small_config_dict()))
    model = CM3PModel(cfg)
    sd = synthetic_state_dict(cfg, seed=0)
    model.load_state_dict(sd, strict=True)
    return cfg, sd, model.cuda().eval()


def test_per_beatmap_mean_embeddings_match_reference_recipe():
    from cm3p_b200.extraction import extract_beatmap_embeddings
    from oracle import cm3p_oracle as O
    cfg, sd, model = _model()
    ids_per_batch = [[11, 11, 7, None], [7, 11, 5, 5]]
    batches, want_acc = [], {}
    for n, ids in enumerate(ids_per_batch):
        b = synthetic_batch(cfg, batch=4, seq_len=320, seed=30 + n)
        b = {k: b[k] for k in ("input_ids", "attention_mask", "input_features")}
        with torch.no_grad():
            e = O.model_forward(sd, cfg, **b, return_loss=False)["beatmap_embeds"].numpy()
        for i, bid in enumerate(ids):          # the reference's host-side accumulation (:243-253)
            if bid is None:
                continue
            s = want_acc.setdefault(bid, {"sum": np.zeros_like(e[i]), "count": 0})
            s["sum"] += e[i]
            s["count"] += 1
        b["beatmap_id"] = ids
        batches.append(b)
    got_ids, got = extract_beatmap_embeddings(model, batches)
    assert got_ids == [11, 7, 5]
    for row, bid in enumerate(got_ids):
        mean = want_acc[bid]["sum"] / want_acc[bid]["count"]
        mean = mean / np.sqrt((mean ** 2).sum())
        cos = float(np.dot(got[row].cpu().numpy(), mean))
        assert cos >= 0.999, (bid, cos)
        assert abs(float(got[row].norm()) - 1.0) < 1e-4


def test_zero_shot_scoring_with_1000_variations():
    """Eval-time shape: few windows, 1000 metadata variations each -> logits_per_beatmap (B, B, V)."""
    from oracle import cm3p_oracle as O
    cfg, sd, model = _model()
    B, V = 4, 1000
    batch = synthetic_batch(cfg, batch=B, seq_len=300, variations=V, seed=8, pad_variations=100)
    with torch.no_grad():
        out = model(**{k: v.cuda() for k, v in batch.items()})
        want = O.model_forward(sd, cfg, **batch)
    assert out.logits_per_beatmap.shape == (B, B, V) and out.logits_per_metadata.shape == (B, V, B)
    got = out.logits_per_beatmap.float().cpu()
    scale = float(np.exp(sd["logit_scale"]))
    assert float((got.double() - want["logits_per_beatmap"].double()).abs().max()) <= 0.02 * scale
    assert abs(float(out.loss) - float(want["loss"])) <= 1e-2 * abs(float(want["loss"]))
    # zero-shot decision per window (train.py:118-131): argmax over its own variations agrees with the oracle
    own_got = torch.stack([got[i, i] for i in range(B)])
    own_want = torch.stack([want["logits_per_beatmap"][i, i] for i in range(B)])
    top_got = own_got.topk(5, dim=-1).indices
    for i in range(B):
        assert int(own_want[i].argmax()) in top_got[i].tolist()
