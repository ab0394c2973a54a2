"""CPU oracle for the CM3P contrastive / embedding hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32 or fp64) restatement of what the reference computes on
the path named by BASELINE.json:north_star.  It is the *checker*: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.
Nothing under `cm3p_b200/` imports it, and the product path has no CPU fallback.

Parity status: PINNED.  `oracle/make_golden.py` ran the unmodified reference model
(`/root/reference/cm3p/modeling_cm3p.py` on top of the installed `transformers` ModernBERT, through
`oracle/ref_shim.py`) in this container and committed its outputs under `tests/golden/`;
`tests/test_oracle_golden.py` checks this restatement against those vectors.  The reference's own
tests hold no numerical fixtures for this path.

What each function follows (reference file:line, relative to /root/reference; "MB:" =
transformers/models/modernbert/modeling_modernbert.py of the installed 5.5.0, the third-party
dependency that holds the encoder arithmetic — reference pin `transformers==4.55.0`, Dockerfile:4):

  layer_norm          MB:63 (nn.LayerNorm, eps=norm_eps, no bias)
  rope_cos_sin        MB:94-172  (inv_freq = theta^(-2k/d); cos/sin cast to activation dtype)
  apply_rope          MB:190-228 (rotate-half, fp32 math, cast back)
  attention_masks     transformers/masking_utils.py:121-131 (|i-j| <= local_attention//2) + key padding
  encoder_forward     MB:313-342 (pre-norm block, layer 0 attn_norm = Identity), MB:446-490
  audio_encoder       cm3p/modeling_cm3p.py:484-528 (+ projector :470-481)
  beatmap_tower       cm3p/modeling_cm3p.py:547-650 (audio scatter :603-605, pooling :624-642)
  metadata_tower      cm3p/modeling_cm3p.py:315-403
  cm3p_loss           cm3p/modeling_cm3p.py:27-51
  model_forward       cm3p/modeling_cm3p.py:849-1012 (norm :54-62, logits :976-982, MLM :987-996)

The arithmetic is written on padded (B, L) batches with SDPA exactly like the reference's CPU
`sdpa` path, so timing it on host cores is a fair stand-in for that path (`cpu_baseline.kind =
"port"`).  Padded *query* rows are meaningless in the reference too (SURVEY.md §8c gotcha 5); here
they are allowed to see every key so they stay finite, and they are never read by real rows.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# config helpers (duck-typed: works with cm3p_b200.CM3PConfig, the reference's CM3PConfig or a dict)

def _ns(cfg):
    if isinstance(cfg, dict):
        cfg = dict(cfg)
        for k in ("metadata_config", "beatmap_config", "audio_config"):
            if k in cfg and isinstance(cfg[k], dict):
                cfg[k] = _ns(cfg[k])
        return SimpleNamespace(**cfg)
    return cfg


def layer_norm(x, weight, eps):
    return F.layer_norm(x, (x.shape[-1],), weight, None, eps)


def rope_cos_sin(positions, head_dim, theta, dtype):
    """positions (B, L) -> cos, sin (B, L, head_dim) in `dtype` (rounded like the reference)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    freqs = positions[:, :, None].float() * inv_freq[None, None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def _rotate_half(x):
    half = x.shape[-1] // 2
    return torch.cat((-x[..., half:], x[..., :half]), dim=-1)


def apply_rope(q, k, cos, sin):
    """q, k (B, h, L, d); cos/sin (B, L, d).  Math in fp32 and cast back — also for fp64 inputs,
    exactly as MB:222-228 (`q.float()`), so an fp64 run of the reference rounds q/k to fp32 here."""
    cos, sin = cos[:, None], sin[:, None]
    qf, kf = q.float(), k.float()
    return (qf * cos + _rotate_half(qf) * sin).to(q.dtype), (kf * cos + _rotate_half(kf) * sin).to(k.dtype)


def attention_masks(key_mask, L, window_half, device):
    """bool masks (B,1,L,L): global = key is real; sliding = also |i-j| <= window_half.
    Padded query rows are opened up completely so softmax stays finite (their output is unused)."""
    B = key_mask.shape[0]
    km = key_mask.bool()
    glob = km[:, None, None, :].expand(B, 1, L, L)
    idx = torch.arange(L, device=device)
    band = (idx[:, None] - idx[None, :]).abs() <= window_half
    slid = glob & band[None, None]
    pad_q = ~km[:, None, :, None]
    return glob | pad_q, slid | pad_q


def encoder_forward(sd, prefix, cfg, embeds, key_mask, positions=None, return_all=False):
    """ModernBERT trunk on already-embedded inputs.  embeds (B, L, H); key_mask (B, L) 0/1."""
    cfg = _ns(cfg)
    B, L, H = embeds.shape
    h = cfg.num_attention_heads
    d = H // h
    dt = embeds.dtype
    if positions is None:
        positions = torch.arange(L)[None].expand(B, L)
    gmask, smask = attention_masks(key_mask, L, cfg.local_attention // 2, embeds.device)
    cs_g = rope_cos_sin(positions, d, cfg.global_rope_theta, dt)
    cs_l = rope_cos_sin(positions, d, cfg.local_rope_theta, dt)

    x = layer_norm(embeds, sd[f"{prefix}.embeddings.norm.weight"], cfg.norm_eps)
    hidden = [x]
    for n in range(cfg.num_hidden_layers):
        p = f"{prefix}.layers.{n}"
        is_global = n % cfg.global_attn_every_n_layers == 0
        a = x if n == 0 else layer_norm(x, sd[f"{p}.attn_norm.weight"], cfg.norm_eps)
        qkv = F.linear(a, sd[f"{p}.attn.Wqkv.weight"]).view(B, L, 3, h, d)
        q, k, v = (t.transpose(1, 2) for t in qkv.unbind(dim=2))
        cos, sin = cs_g if is_global else cs_l
        q, k = apply_rope(q, k, cos, sin)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=gmask if is_global else smask,
                                           scale=d ** -0.5)
        o = o.transpose(1, 2).reshape(B, L, H)
        x = x + F.linear(o, sd[f"{p}.attn.Wo.weight"])
        m = layer_norm(x, sd[f"{p}.mlp_norm.weight"], cfg.norm_eps)
        u, g = F.linear(m, sd[f"{p}.mlp.Wi.weight"]).chunk(2, dim=-1)
        x = x + F.linear(F.gelu(u) * g, sd[f"{p}.mlp.Wo.weight"])
        hidden.append(x)
    out = layer_norm(x, sd[f"{prefix}.final_norm.weight"], cfg.norm_eps)
    return (out, hidden) if return_all else out


def audio_encoder(sd, prefix, acfg, input_features):
    """(B, n_mels, F) log-mel -> audio_embeds (B * F/8, projector_dim), last_hidden (B, F/2, H_a)."""
    acfg = _ns(acfg)
    x = F.gelu(F.conv1d(input_features, sd[f"{prefix}.conv1.weight"], sd[f"{prefix}.conv1.bias"], padding=1))
    x = F.gelu(F.conv1d(x, sd[f"{prefix}.conv2.weight"], sd[f"{prefix}.conv2.bias"], stride=2, padding=1))
    x = x.permute(0, 2, 1).contiguous()
    B, T, _ = x.shape
    ones = torch.ones(B, T, dtype=torch.long)
    last = encoder_forward(sd, f"{prefix}.encoder", acfg, x, ones)
    y = last.reshape(-1, acfg.projector_intermediate_size)
    y = F.gelu(F.linear(y, sd[f"{prefix}.multi_modal_projector.linear_1.weight"]))
    y = F.linear(y, sd[f"{prefix}.multi_modal_projector.linear_2.weight"])
    return y, last


def _pool(last, mask, cls_embed):
    if cls_embed:
        return last[..., 0, :]
    m = mask.unsqueeze(-1).float()
    s = (last * m).sum(dim=-2) / torch.clamp(m.sum(dim=-2), min=1e-9)
    return s.to(last.dtype)


def beatmap_tower(sd, bcfg, input_ids, attention_mask, input_features=None, prefix="beatmap_model"):
    bcfg = _ns(bcfg)
    emb = sd[f"{prefix}.encoder.embeddings.tok_embeddings.weight"][input_ids]
    audio_last = None
    if input_features is not None:
        audio_embeds, audio_last = audio_encoder(sd, f"{prefix}.audio_encoder", bcfg.audio_config, input_features)
        emb = emb.clone()
        emb[input_ids == bcfg.audio_token_id] = audio_embeds.to(emb.dtype)
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    last = encoder_forward(sd, f"{prefix}.encoder", bcfg, emb, attention_mask)
    return last, _pool(last, attention_mask, bcfg.cls_embed), audio_last


def metadata_tower(sd, mcfg, metadata_ids, metadata_attention_mask, prefix="metadata_model"):
    mcfg = _ns(mcfg)
    shape = metadata_ids.shape
    ids = metadata_ids.reshape(-1, shape[-1])
    mask = (metadata_attention_mask.reshape(-1, shape[-1]) if metadata_attention_mask is not None
            else torch.ones_like(ids))
    emb = sd[f"{prefix}.encoder.embeddings.tok_embeddings.weight"][ids]
    last = encoder_forward(sd, f"{prefix}.encoder", mcfg, emb, mask)
    last = last.view(*shape, -1)
    return last, _pool(last, mask.view(*shape), mcfg.cls_embed)


def cm3p_loss(similarity, metadata_variation_classes=None):
    """similarity (B, V, B) [or (B, B)] -> scalar, all B*V metadata rows are negatives (quirk Q2)."""
    if similarity.dim() == 3:
        Bm, V, Bb = similarity.shape
        true_idx = (metadata_variation_classes == 0).int().argmax(dim=1)
        rows = similarity[torch.arange(Bm), true_idx]
        metadata_loss = F.cross_entropy(rows, torch.arange(Bm))
        per_beatmap = similarity.permute(2, 0, 1).reshape(Bb, Bm * V)
        beatmap_loss = F.cross_entropy(per_beatmap, torch.arange(Bm) * V + true_idx)
    else:
        n = similarity.shape[0]
        metadata_loss = F.cross_entropy(similarity, torch.arange(n))
        beatmap_loss = F.cross_entropy(similarity.t(), torch.arange(n))
    return (metadata_loss + beatmap_loss) / 2.0


def mlm_head(sd, bcfg, last_hidden):
    bcfg = _ns(bcfg)
    y = F.gelu(F.linear(last_hidden, sd["head.dense.weight"]))
    y = layer_norm(y, sd["head.norm.weight"], bcfg.norm_eps)
    return F.linear(y, sd["decoder.weight"], sd.get("decoder.bias"))


def model_forward(sd, cfg, input_ids=None, attention_mask=None, input_features=None, metadata_ids=None,
                  metadata_attention_mask=None, metadata_variation_classes=None, labels=None,
                  return_loss=True, num_items_in_batch=None):
    """Restatement of CM3PModel.forward; returns a dict with the CM3POutput field names."""
    cfg = _ns(cfg)
    out = dict(loss=None, logits_per_beatmap=None, logits_per_metadata=None, metadata_embeds=None,
               beatmap_embeds=None, logits=None, beatmap_last_hidden=None, metadata_last_hidden=None,
               audio_last_hidden=None)
    loss = 0 if return_loss else None
    if input_ids is not None:
        last, pooled, audio_last = beatmap_tower(sd, cfg.beatmap_config, input_ids, attention_mask, input_features)
        e = F.linear(pooled, sd["beatmap_projection.weight"])
        out["beatmap_embeds"] = e / e.pow(2).sum(dim=-1, keepdim=True).pow(0.5)
        out["beatmap_last_hidden"] = last
        out["audio_last_hidden"] = audio_last
    if metadata_ids is not None:
        mlast, mpooled = metadata_tower(sd, cfg.metadata_config, metadata_ids, metadata_attention_mask)
        e = F.linear(mpooled, sd["metadata_projection.weight"])
        out["metadata_embeds"] = e / e.pow(2).sum(dim=-1, keepdim=True).pow(0.5)
        out["metadata_last_hidden"] = mlast
    if out["beatmap_embeds"] is not None and out["metadata_embeds"] is not None:
        lpm = torch.matmul(out["metadata_embeds"], out["beatmap_embeds"].t()) * sd["logit_scale"].exp()
        out["logits_per_metadata"] = lpm
        out["logits_per_beatmap"] = lpm.permute(2, 0, 1) if lpm.dim() == 3 else lpm.t()
        if return_loss:
            loss = cm3p_loss(lpm, metadata_variation_classes)
    if getattr(cfg, "has_decoder_head", False) and input_ids is not None:
        logits = mlm_head(sd, cfg.beatmap_config, out["beatmap_last_hidden"])
        out["logits"] = logits
        if labels is not None and return_loss:
            V = logits.shape[-1]
            flat, tgt = logits.reshape(-1, V).float(), labels.reshape(-1)
            if num_items_in_batch is None:
                mlm = F.cross_entropy(flat, tgt, ignore_index=-100)
            else:
                mlm = F.cross_entropy(flat, tgt, ignore_index=-100, reduction="sum") / num_items_in_batch
            loss = loss + 0.5 * mlm
    out["loss"] = loss
    return out


def forward_backward(sd, cfg, batch, wrt=None):
    """Loss + gradients w.r.t. every floating weight (autograd through the restatement)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    out = model_forward(leaves, cfg, **batch)
    out["loss"].backward()
    grads = {k: v.grad for k, v in leaves.items() if v.grad is not None}
    return out, grads


def global_grad_norm(grads):
    return math.sqrt(sum(float(g.double().pow(2).sum()) for g in grads.values()))
