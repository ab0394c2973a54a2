"""Deterministic synthetic weights and batches.

Everything is drawn from `numpy.random.RandomState` (MT19937 + legacy `standard_normal`, whose
streams are frozen by numpy's compatibility policy) so that the oracle goldens committed under
`tests/golden/` can be regenerated bit-identically on the GPU box, where `/root/reference` does
not exist.  Used by tests, `bench.py`, `__graft_entry__.smoke()` and `oracle/make_golden.py`.

The batch layout follows the processor's output contract
(`/root/reference/cm3p/processing_cm3p.py:634-643`, `tokenization_cm3p.py:166-222`): per window
`[AUDIO_BOS] [AUDIO]xA [AUDIO_EOS] [CLS] [BOS] tok... [EOS] [PAD]...`.
"""
from __future__ import annotations

import numpy as np
import torch

from .configuration_cm3p import CM3PConfig


def encoder_param_shapes(prefix: str, cfg, with_tok_embeddings: bool = True) -> dict:
    """State-dict keys/shapes of one ModernBERT tower in the reference's schema."""
    H, I = cfg.hidden_size, cfg.intermediate_size
    shapes = {}
    if with_tok_embeddings:
        shapes[f"{prefix}.embeddings.tok_embeddings.weight"] = (cfg.vocab_size, H)
    shapes[f"{prefix}.embeddings.norm.weight"] = (H,)
    for n in range(cfg.num_hidden_layers):
        if n > 0:
            shapes[f"{prefix}.layers.{n}.attn_norm.weight"] = (H,)
        shapes[f"{prefix}.layers.{n}.attn.Wqkv.weight"] = (3 * H, H)
        shapes[f"{prefix}.layers.{n}.attn.Wo.weight"] = (H, H)
        shapes[f"{prefix}.layers.{n}.mlp_norm.weight"] = (H,)
        shapes[f"{prefix}.layers.{n}.mlp.Wi.weight"] = (2 * I, H)
        shapes[f"{prefix}.layers.{n}.mlp.Wo.weight"] = (H, I)
    shapes[f"{prefix}.final_norm.weight"] = (H,)
    return shapes


def model_param_shapes(config: CM3PConfig) -> dict:
    """All state-dict keys of `CM3PModel` with their shapes (order = reference module order)."""
    mc, bc = config.metadata_config, config.beatmap_config
    ac = bc.audio_config
    shapes = {"logit_scale": ()}
    shapes.update(encoder_param_shapes("metadata_model.encoder", mc))
    ap = "beatmap_model.audio_encoder"
    shapes[f"{ap}.conv1.weight"] = (ac.hidden_size, ac.n_mels, 3)
    shapes[f"{ap}.conv1.bias"] = (ac.hidden_size,)
    shapes[f"{ap}.conv2.weight"] = (ac.hidden_size, ac.hidden_size, 3)
    shapes[f"{ap}.conv2.bias"] = (ac.hidden_size,)
    shapes.update(encoder_param_shapes(f"{ap}.encoder", ac))
    shapes[f"{ap}.multi_modal_projector.linear_1.weight"] = (ac.projector_dim, ac.projector_intermediate_size)
    shapes[f"{ap}.multi_modal_projector.linear_2.weight"] = (ac.projector_dim, ac.projector_dim)
    shapes.update(encoder_param_shapes("beatmap_model.encoder", bc))
    shapes["beatmap_projection.weight"] = (config.projection_dim, bc.hidden_size)
    shapes["metadata_projection.weight"] = (config.projection_dim, mc.hidden_size)
    if config.has_decoder_head:
        shapes["head.dense.weight"] = (bc.hidden_size, bc.hidden_size)
        shapes["head.norm.weight"] = (bc.hidden_size,)
        shapes["decoder.weight"] = (bc.vocab_size, bc.hidden_size)
        if bc.decoder_bias:
            shapes["decoder.bias"] = (bc.vocab_size,)
    return shapes


def synthetic_state_dict(config: CM3PConfig, seed: int = 0, gain: float | None = None,
                         dtype=torch.float32) -> dict:
    """Seeded weights in the reference's key schema.

    gain=None  -> "reference-like": N(0, initializer_range) matrices, unit LayerNorm weights
                  (what `_init_weights`, modeling_cm3p.py:262-297, produces up to the RNG stream).
    gain=g     -> stress init: matrices N(0, g/sqrt(fan_in)) so every branch contributes O(g) to the
                  residual stream, LayerNorm weights 1 + 0.1 N(0,1), conv biases 0.1 N(0,1).
    """
    rs = np.random.RandomState(seed)
    out = {}
    for name, shape in model_param_shapes(config).items():
        if name == "logit_scale":
            arr = np.array(config.logit_scale_init_value, dtype=np.float64)
        elif name.endswith("norm.weight"):
            arr = np.ones(shape) if gain is None else 1.0 + 0.1 * rs.standard_normal(shape)
        elif name.endswith(".bias"):
            arr = np.zeros(shape) if gain is None else 0.1 * rs.standard_normal(shape)
        else:
            fan_in = int(np.prod(shape[1:]))
            if gain is None:
                std = config.initializer_range
                if name.endswith("projection.weight"):
                    std = shape[1] ** -0.5 * config.initializer_factor
            else:
                std = gain / np.sqrt(fan_in)
                if "tok_embeddings" in name:
                    std = 1.0
            arr = std * rs.standard_normal(shape)
        out[name] = torch.from_numpy(np.asarray(arr, dtype=np.float64)).to(dtype)
    return out


def synthetic_batch(config: CM3PConfig, batch: int, seq_len: int, variations: int = 1, seed: int = 1,
                    min_len: int | None = None, n_audio_tokens: int | None = None,
                    n_frames: int | None = None, with_labels: bool = False, meta_len: int = 128,
                    pad_variations: int = 0) -> dict:
    """One model-input dict with the schema of SURVEY.md §8a row 0 (CPU tensors).

    lengths: len_0 = seq_len, others U{min_len..seq_len}.  `variations` == 1 gives 2-D metadata
    (B, S); otherwise (B, V, S) with classes[:, 0] = 0, others U{1..4} and the last
    `pad_variations` columns set to the class -1 "padding variation" (all-UNK rows, quirk Q2).
    """
    bc, mc = config.beatmap_config, config.metadata_config
    ac = bc.audio_config
    rs = np.random.RandomState(seed)
    if n_frames is None:
        n_frames = 1600
    frames_per_token = 2 * (ac.projector_intermediate_size // ac.hidden_size)  # conv2 stride 2, then 4->1
    if n_audio_tokens is None:
        n_audio_tokens = n_frames // frames_per_token
    assert n_audio_tokens * frames_per_token == n_frames
    n_special = n_audio_tokens + 5  # BOS_A, A.., EOS_A, CLS, BOS ... EOS
    if min_len is None:
        min_len = max(n_special + 1, (seq_len * 3) // 10)
    min_len = max(min_len, n_special + 1)
    assert seq_len >= min_len

    lens = rs.randint(min_len, seq_len + 1, size=batch)
    lens[0] = seq_len
    n_plain = min(bc.vocab_size, bc.audio_sos_token_id, bc.audio_token_id, bc.audio_eos_token_id) - 70
    n_plain = max(n_plain, 8)
    ids = rs.randint(0, n_plain, size=(batch, seq_len)).astype(np.int64)
    cls_id, mask_id = n_plain + 5, n_plain + 6
    bos_id, eos_id = bc.bos_token_id, bc.eos_token_id
    pad_id = bc.pad_token_id if bc.pad_token_id is not None else 0
    if bos_id is None or bos_id >= bc.vocab_size:
        bos_id = 1
    if eos_id is None or eos_id >= bc.vocab_size:
        eos_id = 2
    mask = np.zeros((batch, seq_len), dtype=np.int64)
    for b in range(batch):
        n = int(lens[b])
        ids[b, 0] = bc.audio_sos_token_id
        ids[b, 1:1 + n_audio_tokens] = bc.audio_token_id
        ids[b, 1 + n_audio_tokens] = bc.audio_eos_token_id
        ids[b, 2 + n_audio_tokens] = cls_id
        ids[b, 3 + n_audio_tokens] = bos_id
        ids[b, n - 1] = eos_id
        ids[b, n:] = pad_id
        mask[b, :n] = 1
    feats = rs.standard_normal((batch, ac.n_mels, n_frames)).astype(np.float32)

    V = variations
    mlen = rs.randint(17, 26, size=(batch, V))
    mlen = np.minimum(mlen, meta_len)
    n_meta_plain = max(mc.vocab_size - 200, 8) if mc.vocab_size > 400 else mc.vocab_size
    mids = rs.randint(3, n_meta_plain, size=(batch, V, meta_len)).astype(np.int64)
    mmask = (np.arange(meta_len)[None, None, :] < mlen[:, :, None]).astype(np.int64)
    mids = np.where(mmask == 1, mids, mc.pad_token_id if mc.pad_token_id is not None else 0)
    classes = rs.randint(1, 5, size=(batch, V)).astype(np.int64)
    classes[:, 0] = 0
    if pad_variations:
        unk = 3
        classes[:, V - pad_variations:] = -1
        mids[:, V - pad_variations:, :] = np.where(mmask[:, V - pad_variations:, :] == 1, unk,
                                                   mids[:, V - pad_variations:, :])
    out = {
        "input_ids": torch.from_numpy(ids),
        "attention_mask": torch.from_numpy(mask),
        "input_features": torch.from_numpy(feats),
    }
    if V == 1:
        out["metadata_ids"] = torch.from_numpy(mids[:, 0])
        out["metadata_attention_mask"] = torch.from_numpy(mmask[:, 0])
    else:
        out["metadata_ids"] = torch.from_numpy(mids)
        out["metadata_attention_mask"] = torch.from_numpy(mmask)
        out["metadata_variation_classes"] = torch.from_numpy(classes)
    if with_labels:
        # MLM recipe of utils/mmrs_dataset.py:195-217: 15 % of non-special positions, 80/10/10.
        labels = np.full((batch, seq_len), -100, dtype=np.int64)
        masked_ids = ids.copy()
        for b in range(batch):
            n = int(lens[b])
            cand = np.arange(n_special - 1, n - 1)
            pick = cand[rs.rand(cand.size) < 0.15]
            labels[b, pick] = ids[b, pick]
            r = rs.rand(pick.size)
            masked_ids[b, pick[r < 0.8]] = mask_id
            rnd = pick[(r >= 0.8) & (r < 0.9)]
            masked_ids[b, rnd] = rs.randint(0, n_plain, size=rnd.size)
        out["input_ids"] = torch.from_numpy(masked_ids)
        out["labels"] = torch.from_numpy(labels)
    return out
