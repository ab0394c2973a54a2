"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/cm3p_b200.h
declares; the host-side mirror keeps the reference's contracts (config fields, state-dict schema,
output field order, unpadding bookkeeping).  No compute entry point is called here."""
import copy
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "cm3p_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cm3p_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from cm3p_b200 import _lib, build
    build.build()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/cm3p_b200.h but not exported"
    assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)
    assert lib.cm3p_version() == 200


def test_compute_entry_fails_loudly_without_gpu():
    """No CPU fallback: on a box without an sm_100 device the entry returns an error, not a result."""
    if torch.cuda.is_available():
        pytest.skip("needs a CPU-only box")
    from cm3p_b200 import _lib
    lib = _lib.load()
    rc = lib.cm3p_layernorm_fwd(None, None, None, None, 8, 64, 1e-5, None)
    assert rc != 0
    assert "sm_100" in _lib.last_error()


def test_config_contract_matches_reference_yaml_fields():
    from cm3p_b200 import CM3PConfig
    from cm3p_b200.configuration_cm3p import base_config_dict
    cfg = CM3PConfig(**copy.deepcopy(base_config_dict()))
    bc, mc, ac = cfg.beatmap_config, cfg.metadata_config, cfg.beatmap_config.audio_config
    assert (bc.hidden_size, bc.num_hidden_layers, bc.num_attention_heads, bc.intermediate_size) == (768, 22, 12, 1152)
    assert (ac.hidden_size, ac.num_hidden_layers, ac.num_attention_heads, ac.intermediate_size) == (512, 6, 8, 1024)
    assert (mc.hidden_size, mc.num_hidden_layers, mc.num_attention_heads, mc.intermediate_size) == (256, 6, 4, 512)
    assert (bc.global_attn_every_n_layers, bc.local_attention, bc.global_rope_theta, bc.local_rope_theta) == \
        (3, 128, 160000.0, 10000.0)
    assert mc.global_attn_every_n_layers == 1 and mc.global_rope_theta == 10000.0
    assert ac.projector_intermediate_size == 2048 and ac.projector_dim == 768 and ac.n_mels == 80
    assert cfg.projection_dim == 512 and abs(cfg.logit_scale_init_value - 2.6592) < 1e-9
    # round trip through the HF serialisation
    again = CM3PConfig.from_dict(cfg.to_dict())
    assert again.beatmap_config.audio_config.hidden_size == 512
    assert "reference_compile" not in cfg.beatmap_config.to_dict()
    # attn_implementation is accepted and ignored (one backend)
    assert CM3PConfig(attn_implementation="sdpa").beatmap_config.hidden_size == 768


def test_state_dict_schema_and_output_order():
    from cm3p_b200 import CM3PConfig
    from cm3p_b200.configuration_cm3p import small_config_dict
    from cm3p_b200.modeling_cm3p import CM3PModel, CM3POutput
    from cm3p_b200.synthetic import model_param_shapes
    d = small_config_dict()
    d["has_decoder_head"] = True
    cfg = CM3PConfig(**d)
    model = CM3PModel(cfg)
    sd = model.state_dict()
    want = model_param_shapes(cfg)
    assert set(sd) == set(want)
    for k, shape in want.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    assert "beatmap_model.encoder.layers.0.attn_norm.weight" not in sd  # Identity in layer 0
    assert [f for f in CM3POutput.__dataclass_fields__] == [
        "loss", "logits_per_beatmap", "logits_per_metadata", "metadata_embeds", "beatmap_embeds", "logits",
        "metadata_model_output", "beatmap_model_output"]
    # names that train.py:331-340 uses to route parameters to AdamW vs Muon
    assert any("embed" in k for k in sd) and hasattr(model, "beatmap_model") and hasattr(model, "metadata_model")


def test_unpad_bookkeeping_matches_reference_helper():
    from cm3p_b200.modeling_cm3p import _repad, _unpad
    mask = torch.tensor([[1, 1, 1, 0, 0], [1, 1, 1, 1, 1], [1, 0, 0, 0, 0]])
    up = _unpad(mask, 3, 5, torch.device("cpu"))
    # reference: indices = nonzero(mask.flatten()); cu = pad(cumsum(lens)); max_seqlen = max(lens)
    assert up.src_index.tolist() == torch.nonzero(mask.flatten()).flatten().tolist()
    assert up.cu_seqlens.tolist() == [0, 3, 8, 9] and up.max_len == 5 and up.total == 9
    assert up.positions.tolist() == [0, 1, 2, 0, 1, 2, 3, 4, 0]
    x = torch.arange(9.0)[:, None].repeat(1, 2)
    padded = _repad(x, up)
    assert padded.shape == (3, 5, 2) and float(padded[0, 3, 0]) == 0.0 and float(padded[1, 4, 1]) == 7.0
    full = _unpad(None, 2, 4, torch.device("cpu"))
    assert full.cu_seqlens.tolist() == [0, 4, 8] and full.total == 8


def test_wi_interleave_roundtrip():
    from cm3p_b200 import ops
    w = torch.arange(2 * 32 * 3, dtype=torch.float32).reshape(64, 3)
    il = ops.interleave_wi(w)
    assert torch.equal(ops.deinterleave_wi(il), w)
    assert torch.equal(il[:16], w[:16]) and torch.equal(il[16:32], w[32:48]) and torch.equal(il[32:48], w[16:32])
