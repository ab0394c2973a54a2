"""Generate tests/golden/*.npz by running the UNMODIFIED reference model.  TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

For every case below: seeded weights (`cm3p_b200.synthetic.synthetic_state_dict`) are loaded into
the reference `CM3PModel` by state-dict, the seeded batch (`synthetic_batch`) is run forward (and
backward for the `grads` cases) on CPU with `sdpa`, and the outputs are stored.  Weights and inputs
are *not* stored: both are regenerated from the seeds (numpy MT19937), so fixtures stay a few KB.
`tests/test_oracle_golden.py` checks `oracle/cm3p_oracle.py` against these vectors; the GPU parity
tests then check the CUDA path against the oracle and against these vectors directly.
"""
from __future__ import annotations

import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cm3p_b200.configuration_cm3p import CM3PConfig, small_config_dict  # noqa: E402
from cm3p_b200.synthetic import synthetic_batch, synthetic_state_dict  # noqa: E402
from oracle.ref_shim import build_reference_model  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def _cfg(cls_embed=True, has_decoder_head=False):
    d = small_config_dict()
    d["metadata_config"]["cls_embed"] = cls_embed
    d["beatmap_config"]["cls_embed"] = cls_embed
    if has_decoder_head:
        d["has_decoder_head"] = True
        d["loss_type"] = "ForMaskedLM"
    return d


# name -> dict(config, batch kwargs, weight seed/gain, dtype, grads?)
CASES = {
    # BASELINE.json config 1: small config, B=8, L=512, fp32, sdpa on CPU, reference-like init
    "small_b8_l512_v1": dict(cfg=_cfg(), batch=dict(batch=8, seq_len=512, variations=1, seed=1),
                             wseed=0, gain=None, dtype="float32", grads=False),
    # stress init, variations incl. a class -1 padding variation, fp64, with gradients
    "small_b4_l400_v3_grads": dict(cfg=_cfg(), batch=dict(batch=4, seq_len=400, variations=3, seed=2,
                                                         pad_variations=1),
                                   wseed=3, gain=1.0, dtype="float64", grads=True),
    # masked-mean pooling on both towers
    "small_b3_l300_mean": dict(cfg=_cfg(cls_embed=False), batch=dict(batch=3, seq_len=300, variations=2, seed=4),
                               wseed=5, gain=1.0, dtype="float64", grads=True),
    # decoder head + 0.5 * MLM loss
    "small_b3_l320_mlm": dict(cfg=_cfg(has_decoder_head=True),
                              batch=dict(batch=3, seq_len=320, variations=2, seed=6, with_labels=True),
                              wseed=7, gain=1.0, dtype="float64", grads=True),
}


def run_case(name, case):
    dtype = getattr(torch, case["dtype"])
    model, ref_cfg = build_reference_model(copy.deepcopy(case["cfg"]), "sdpa")
    ours_cfg = CM3PConfig(**copy.deepcopy(case["cfg"]))
    sd = synthetic_state_dict(ours_cfg, seed=case["wseed"], gain=case["gain"], dtype=dtype)
    model = model.to(dtype)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    # the only tolerated difference: nothing.  Key schema must match exactly.
    assert not missing and not unexpected, (missing, unexpected)
    batch = synthetic_batch(ours_cfg, **case["batch"])
    feed = {k: (v.to(dtype) if v.is_floating_point() else v.clone()) for k, v in batch.items()}
    if case["grads"]:
        out = model(**feed)
        out.loss.backward()
    else:
        with torch.no_grad():
            out = model(**feed)
    mask = batch["attention_mask"].bool()
    rec = {
        "loss": np.float64(out.loss.item()),
        "beatmap_embeds": out.beatmap_embeds.detach().double().numpy(),
        "metadata_embeds": out.metadata_embeds.detach().double().numpy(),
        "logits_per_metadata": out.logits_per_metadata.detach().double().numpy(),
        "logits_per_beatmap": out.logits_per_beatmap.detach().double().numpy(),
        # real rows of the beatmap tower's last hidden state: first 6 and last 2 of every window
        "hidden_probe": np.stack([
            torch.cat([out.beatmap_model_output.last_hidden_state[b][mask[b]][:6],
                       out.beatmap_model_output.last_hidden_state[b][mask[b]][-2:]]).detach().double().numpy()
            for b in range(mask.shape[0])]),
        "input_ids_sum": np.int64(batch["input_ids"].sum().item()),
        "features_sum": np.float64(batch["input_features"].double().sum().item()),
    }
    if out.logits is not None:
        rec["mlm_logits_probe"] = out.logits[:, 205:213, :16].detach().double().numpy()
    if case["grads"]:
        names, norms = [], []
        for k, p in model.named_parameters():
            if p.grad is not None:
                names.append(k)
                norms.append(float(p.grad.double().norm()))
        rec["grad_names"] = np.array(names)
        rec["grad_norms"] = np.array(norms, dtype=np.float64)
        rec["grad_global_norm"] = np.float64(np.sqrt((np.array(norms) ** 2).sum()))
        g = dict(model.named_parameters())["beatmap_model.encoder.layers.1.attn.Wqkv.weight"].grad
        rec["grad_probe_wqkv1"] = g[:8, :8].detach().double().numpy()
        g = dict(model.named_parameters())["beatmap_model.audio_encoder.conv1.weight"].grad
        rec["grad_probe_conv1"] = g[:4, :4, :].detach().double().numpy()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: loss={rec['loss']:.10f} -> {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    only = sys.argv[1:]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        run_case(name, case)
