"""Pin the CPU oracle against vectors produced by the UNMODIFIED reference (oracle/make_golden.py)."""
import copy
import os

import numpy as np
import pytest
import torch

from cm3p_b200.configuration_cm3p import CM3PConfig
from cm3p_b200.synthetic import synthetic_batch, synthetic_state_dict
from oracle import cm3p_oracle as O
from oracle.make_golden import CASES


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, f"{name}.npz"), allow_pickle=False)


def _run(case, grads):
    cfg = CM3PConfig(**copy.deepcopy(case["cfg"]))
    dtype = getattr(torch, case["dtype"])
    sd = synthetic_state_dict(cfg, seed=case["wseed"], gain=case["gain"], dtype=dtype)
    batch = synthetic_batch(cfg, **case["batch"])
    feed = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in batch.items()}
    if grads:
        out, g = O.forward_backward(sd, cfg, feed)
    else:
        with torch.no_grad():
            out, g = O.model_forward(sd, cfg, **feed), None
    return cfg, batch, out, g


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(golden_dir, name):
    case = CASES[name]
    gold = _load(golden_dir, name)
    cfg, batch, out, grads = _run(case, case["grads"])
    # the seeded inputs regenerate identically
    assert int(batch["input_ids"].sum()) == int(gold["input_ids_sum"])
    assert abs(float(batch["input_features"].double().sum()) - float(gold["features_sum"])) < 1e-6
    tol = 2e-5 if case["dtype"] == "float32" else 1e-9
    assert abs(float(out["loss"].detach()) - float(gold["loss"])) <= tol * max(1.0, abs(float(gold["loss"])))
    for key in ("beatmap_embeds", "metadata_embeds", "logits_per_metadata", "logits_per_beatmap"):
        got = out[key].detach().double().numpy()
        assert got.shape == gold[key].shape, key
        np.testing.assert_allclose(got, gold[key], rtol=0, atol=tol * 20, err_msg=key)
    mask = batch["attention_mask"].bool()
    last = out["beatmap_last_hidden"].detach().double()
    probe = np.stack([torch.cat([last[b][mask[b]][:6], last[b][mask[b]][-2:]]).numpy()
                      for b in range(mask.shape[0])])
    np.testing.assert_allclose(probe, gold["hidden_probe"], rtol=0,
                               atol=(2e-3 if case["dtype"] == "float32" else 1e-8))
    if "mlm_logits_probe" in gold.files:
        np.testing.assert_allclose(out["logits"][:, 205:213, :16].detach().double().numpy(),
                                   gold["mlm_logits_probe"], rtol=0, atol=1e-8)
    if case["grads"]:
        names = [str(n) for n in gold["grad_names"]]
        assert set(names) == set(grads), set(names) ^ set(grads)
        got = np.array([float(grads[n].double().norm()) for n in names])
        np.testing.assert_allclose(got, gold["grad_norms"], rtol=1e-7, atol=1e-12)
        assert abs(O.global_grad_norm(grads) - float(gold["grad_global_norm"])) < 1e-8 * float(gold["grad_global_norm"])
        np.testing.assert_allclose(grads["beatmap_model.encoder.layers.1.attn.Wqkv.weight"][:8, :8].numpy(),
                                   gold["grad_probe_wqkv1"], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(grads["beatmap_model.audio_encoder.conv1.weight"][:4, :4].numpy(),
                                   gold["grad_probe_conv1"], rtol=1e-5, atol=1e-9)


def test_loss_quirk_padding_variations_are_negatives():
    """Q2: class -1 variations stay in the softmax denominators (modeling_cm3p.py:33-51)."""
    torch.manual_seed(0)
    sim = torch.randn(3, 4, 3, dtype=torch.float64)
    classes = torch.tensor([[0, 1, 2, -1], [3, 0, -1, -1], [0, 4, 1, 2]])
    full = O.cm3p_loss(sim, classes)
    t = torch.tensor([0, 1, 0])
    rows = sim[torch.arange(3), t]
    ml = torch.nn.functional.cross_entropy(rows, torch.arange(3))
    bl = torch.nn.functional.cross_entropy(sim.permute(2, 0, 1).reshape(3, 12), torch.arange(3) * 4 + t)
    assert abs(float(full) - float((ml + bl) / 2)) < 1e-12


def test_muon_oracle_matches_reference(golden_dir):
    """oracle/muon_oracle.py vs three steps of the unmodified utils/muon_utils.Muon (bit-exact on CPU)."""
    from oracle import muon_oracle as M
    from oracle.make_golden_muon import HYPER, SPECS, STEPS, seeded
    torch.set_num_threads(1)
    gold = np.load(os.path.join(golden_dir, "muon_steps.npz"))
    params, grads = seeded()
    state, use = {}, {n: m for n, _, m in SPECS}
    for t in range(STEPS):
        M.muon_step(params, grads[t], state, use, lr=HYPER["lr"], adamw_lr=HYPER["adamw_lr"],
                    adamw_betas=HYPER["adamw_betas"], adamw_wd=HYPER["adamw_wd"], adamw_eps=HYPER["adamw_eps"])
        for n in params:
            np.testing.assert_allclose(params[n].numpy(), gold[f"step{t}/{n}"], rtol=0, atol=1e-7, err_msg=f"{t}/{n}")
