"""End-to-end parity of the CUDA path (through the public CM3PModel API -> C ABI) against
(a) the goldens produced by the unmodified reference and (b) the CPU oracle, on identical seeded
weights and inputs.  Tolerances are the north-star's: embedding cosine >= 0.999, loss within 1e-2
relative (bf16 activations)."""
import copy
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from cm3p_b200.configuration_cm3p import CM3PConfig, base_config_dict
from cm3p_b200.synthetic import synthetic_batch, synthetic_state_dict
from oracle.make_golden import CASES


def _cos(a, b):
    a, b = torch.as_tensor(a).double().flatten(0, -2), torch.as_tensor(b).double().flatten(0, -2)
    return torch.nn.functional.cosine_similarity(a, b, dim=-1)


def _build(cfg_dict, wseed, gain, attn_impl=None):
    from cm3p_b200.modeling_cm3p import CM3PModel
    cfg = CM3PConfig(attn_implementation=attn_impl, **copy.deepcopy(cfg_dict))
    model = CM3PModel(cfg)
    sd = synthetic_state_dict(cfg, seed=wseed, gain=gain)
    model.load_state_dict(sd, strict=True)
    return cfg, sd, model.cuda().eval()


@pytest.mark.parametrize("name", list(CASES))
def test_cuda_path_matches_reference_golden(golden_dir, name):
    case = CASES[name]
    gold = np.load(os.path.join(golden_dir, f"{name}.npz"))
    cfg, sd, model = _build(case["cfg"], case["wseed"], case["gain"])
    batch = synthetic_batch(cfg, **case["batch"])
    feed = {k: v.cuda() for k, v in batch.items()}
    with torch.no_grad():
        out = model(**feed)
    torch.cuda.synchronize()
    cb = _cos(out.beatmap_embeds.cpu(), gold["beatmap_embeds"])
    cm = _cos(out.metadata_embeds.cpu(), gold["metadata_embeds"])
    assert float(cb.min()) >= 0.999, f"beatmap embeds cosine {cb.tolist()}"
    assert float(cm.min()) >= 0.999, f"metadata embeds cosine min {float(cm.min())}"
    # contrastive part of the loss (the goldens' loss includes 0.5*MLM when the decoder head is on)
    lpm = out.logits_per_metadata.float().cpu()
    assert lpm.shape == gold["logits_per_metadata"].shape
    assert out.logits_per_beatmap.shape == gold["logits_per_beatmap"].shape
    scale = float(np.exp(sd["logit_scale"]))
    assert float((lpm.double() - torch.from_numpy(gold["logits_per_metadata"])).abs().max()) <= 0.02 * scale
    if not cfg.has_decoder_head:
        assert abs(float(out.loss) - float(gold["loss"])) <= 1e-2 * abs(float(gold["loss"]))
    # last hidden state on real rows (padded output, like the reference's sdpa path)
    mask = batch["attention_mask"].bool()
    last = out.beatmap_model_output.last_hidden_state.float().cpu()
    assert last.shape[:2] == mask.shape
    probe = torch.stack([torch.cat([last[b][mask[b]][:6], last[b][mask[b]][-2:]]) for b in range(mask.shape[0])])
    ch = _cos(probe, gold["hidden_probe"])
    assert float(ch.min()) >= 0.995, f"hidden-state cosine min {float(ch.min())}"
    if "mlm_logits_probe" in gold.files:
        got = out.logits[:, 205:213, :16].float().cpu()
        want = torch.from_numpy(gold["mlm_logits_probe"])
        assert float(_cos(got, want).min()) >= 0.99


def test_cuda_path_matches_oracle_base_config():
    """Base (production) architecture, random reference-like init, B=3, ragged lengths."""
    from oracle import cm3p_oracle as O
    cfg, sd, model = _build(base_config_dict(), wseed=11, gain=None, attn_impl="flash_attention_2")
    batch = synthetic_batch(cfg, batch=3, seq_len=700, variations=4, seed=12, min_len=300, pad_variations=1)
    with torch.no_grad():
        want = O.model_forward(sd, cfg, **batch)
        out = model(**{k: v.cuda() for k, v in batch.items()})
    torch.cuda.synchronize()
    assert float(_cos(out.beatmap_embeds.cpu(), want["beatmap_embeds"]).min()) >= 0.999
    assert float(_cos(out.metadata_embeds.cpu(), want["metadata_embeds"]).min()) >= 0.999
    assert abs(float(out.loss) - float(want["loss"])) <= 1e-2 * abs(float(want["loss"]))
    # flash_attention_2 setting -> unpadded hidden states (quirk Q6)
    T = int(batch["attention_mask"].sum())
    assert out.beatmap_model_output.last_hidden_state.shape == (T, cfg.beatmap_config.hidden_size)
    real = want["beatmap_last_hidden"][batch["attention_mask"].bool()]
    assert float(_cos(out.beatmap_model_output.last_hidden_state.float().cpu(), real).min()) >= 0.99


def test_inference_pure_bf16_weights():
    """extract_beatmap_embeddings.py path: model.to(bfloat16), return_loss=False, beatmap tower only."""
    from oracle import cm3p_oracle as O
    cfg, sd, model = _build(CASES["small_b8_l512_v1"]["cfg"], 0, None)
    model = model.to(torch.bfloat16)
    batch = synthetic_batch(cfg, batch=5, seq_len=420, seed=21)
    with torch.no_grad():
        out = model(input_ids=batch["input_ids"].cuda(), attention_mask=batch["attention_mask"].cuda(),
                    input_features=batch["input_features"].cuda(), return_loss=False)
        want = O.model_forward(sd, cfg, input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                               input_features=batch["input_features"], return_loss=False)
    assert out.loss is None and out.metadata_embeds is None
    assert out.beatmap_embeds.dtype == torch.bfloat16
    assert float(_cos(out.beatmap_embeds.float().cpu(), want["beatmap_embeds"]).min()) >= 0.999


def test_non_cuda_input_fails_loudly():
    cfg, sd, model = _build(CASES["small_b8_l512_v1"]["cfg"], 0, None)
    batch = synthetic_batch(cfg, batch=2, seq_len=300, seed=3)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"], return_loss=False)
