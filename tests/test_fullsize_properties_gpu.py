"""Size-independent properties at BASELINE.json's full sizes (where the CPU oracle would take minutes):
batch-permutation equivariance and batch-split consistency of the embedding path at configs[1] size
(base model, 64 windows x L=2000), padding invariance, and dense == sparse masked-LM loss at the
configs[4] sequence length (8192 tokens)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

from cm3p_b200.configuration_cm3p import CM3PConfig, base_config_dict
from cm3p_b200.synthetic import synthetic_batch, synthetic_state_dict


def _cos(a, b):
    return torch.nn.functional.cosine_similarity(a.double(), b.double(), dim=-1)


@pytest.fixture(scope="module")
def base_model():
    from cm3p_b200.modeling_cm3p import CM3PModel
    cfg = CM3PConfig(attn_implementation="flash_attention_2", **copy.deepcopy(base_config_dict()))
    model = CM3PModel(cfg)
    model.load_state_dict(synthetic_state_dict(cfg, seed=0), strict=True)
    return cfg, model.cuda().to(torch.bfloat16).eval()


def _embed(model, batch, idx=None):
    feed = {k: (v if idx is None else v[idx]).cuda() for k, v in batch.items()
            if k in ("input_ids", "attention_mask", "input_features")}
    with torch.no_grad():
        return model(**feed, return_loss=False).beatmap_embeds.float().cpu()


def test_full_size_batch_permutation_and_split(base_model):
    cfg, model = base_model
    batch = synthetic_batch(cfg, batch=64, seq_len=2000, seed=1, min_len=600)
    full = _embed(model, batch)
    assert full.shape == (64, 512) and torch.isfinite(full).all()
    assert float((full.norm(dim=-1) - 1).abs().max()) < 2e-2          # L2-normalised (bf16 output)
    perm = torch.randperm(64, generator=torch.Generator().manual_seed(3))
    assert float(_cos(_embed(model, batch, perm), full[perm]).min()) >= 0.99999
    halves = torch.cat([_embed(model, batch, torch.arange(0, 32)), _embed(model, batch, torch.arange(32, 64))])
    assert float(_cos(halves, full).min()) >= 0.99999
    # idempotence
    assert torch.equal(_embed(model, batch), full)


def test_padding_invariance(base_model):
    cfg, model = base_model
    batch = synthetic_batch(cfg, batch=4, seq_len=1500, seed=5, min_len=700)
    ref = _embed(model, batch)
    pad = 2000 - 1500
    wider = dict(batch)
    wider["input_ids"] = torch.nn.functional.pad(batch["input_ids"], (0, pad), value=cfg.beatmap_config.pad_token_id)
    wider["attention_mask"] = torch.nn.functional.pad(batch["attention_mask"], (0, pad), value=0)
    assert float(_cos(_embed(model, wider), ref).min()) >= 0.99999


def test_masked_lm_8k_tokens_dense_equals_sparse():
    """configs[4] sequence length: 8192-token windows through global + window layers, fwd + bwd."""
    from cm3p_b200.modeling_cm3p import CM3PForMaskedLM
    cfg = CM3PConfig(**copy.deepcopy(base_config_dict(has_decoder_head=True)))
    bc = cfg.beatmap_config
    sd = {k: v for k, v in synthetic_state_dict(cfg, seed=0).items()
          if k.startswith(("beatmap_model.", "head.", "decoder."))}
    batch = synthetic_batch(cfg, batch=2, seq_len=8192, seed=7, min_len=6000, with_labels=True)
    feed = {k: batch[k].cuda() for k in ("input_ids", "attention_mask", "input_features", "labels")}
    losses, gnorms = [], []
    for sparse in (False, True):
        c = copy.deepcopy(bc)
        c.sparse_prediction = sparse
        model = CM3PForMaskedLM(c)
        model.load_state_dict(sd, strict=True)
        model = model.cuda().train()
        out = model(**feed)
        out.loss.backward()
        torch.cuda.synchronize()
        losses.append(float(out.loss.detach()))
        gnorms.append(float(torch.sqrt(sum(p.grad.double().pow(2).sum() for p in model.parameters()
                                            if p.grad is not None))))
        if not sparse:
            assert out.logits.shape == (2, 8192, bc.vocab_size)
        del model, out
        torch.cuda.empty_cache()
    assert all(l == l and l > 0 for l in losses)
    assert abs(losses[0] - losses[1]) <= 1e-3 * abs(losses[0]), losses
    assert abs(gnorms[0] - gnorms[1]) <= 1e-2 * gnorms[0], gnorms
    # random-init model: loss close to ln(vocab)
    import math
    assert abs(losses[0] - math.log(bc.vocab_size)) < 1.0
