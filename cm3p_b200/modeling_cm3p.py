"""CM3P model classes backed by the sm_100a kernels (host-side mirror of the reference interface).

Mirrors `/root/reference/cm3p/modeling_cm3p.py`: same class names, `forward` keyword sets,
`CM3POutput` field order (:237-244, relied on positionally by train.py:77,101), attribute names
(`model.beatmap_model`, `model.metadata_model`, used for freezing at train.py:34,317-321) and
state-dict keys, so checkpoints and callers written for the reference work
unchanged.  What is different is everything underneath: no `transformers.ModernBertModel`, no
SDPA / flash-attn dispatch — every tower runs *unpadded* through `cm3p_b200.ops` (C ABI ->
hand-written CUDA).  The nn.Embedding / nn.Linear / nn.LayerNorm / nn.Conv1d members below are
parameter containers only (they give the reference's key names and init); their own `forward`
is never called, and there is no CPU or PyTorch fallback: calling a model on a non-CUDA tensor
raises.

Numerical regime: activations bf16, LayerNorm / softmax / loss statistics fp32, fp32 accumulation
on the tensor cores — i.e. the reference's `bf16` inference (`model.to(bfloat16)`,
extract_beatmap_embeddings.py:161-168) and its autocast-bf16 training (configs/train/*.yaml
`bf16: true`).  fp32 parameters are kept as master weights; bf16 working copies are re-packed
whenever a parameter changes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional

import torch
from torch import nn
from transformers import AutoModel, AutoModelForMaskedLM, AutoModelForSequenceClassification
from transformers.modeling_outputs import BaseModelOutput, BaseModelOutputWithPooling, MaskedLMOutput
from transformers.modeling_utils import PreTrainedModel
from transformers.utils import ModelOutput

from . import ops
from .configuration_cm3p import CM3PAudioConfig, CM3PBeatmapConfig, CM3PConfig, CM3PMetadataConfig


# ------------------------------------------------------------------------------------------------
# outputs (field names / order = reference modeling_cm3p.py:137-250)

@dataclass
class BeatmapClassifierOutput(ModelOutput):
    loss: Optional[torch.FloatTensor] = None
    logits: Optional[torch.FloatTensor] = None
    hidden_states: Optional[tuple] = None
    attentions: Optional[tuple] = None


@dataclass
class CM3PAudioModelOutput(BaseModelOutput):
    audio_embeds: Optional[torch.FloatTensor] = None


@dataclass
class CM3PBeatmapModelOutput(BaseModelOutputWithPooling):
    beatmap_embeds: Optional[torch.FloatTensor] = None
    audio_model_output: Optional[CM3PAudioModelOutput] = None


@dataclass
class CM3PMetadataModelOutput(BaseModelOutput):
    metadata_embeds: Optional[torch.FloatTensor] = None


@dataclass
class CM3POutput(ModelOutput):
    loss: Optional[torch.FloatTensor] = None
    logits_per_beatmap: Optional[torch.Tensor] = None
    logits_per_metadata: Optional[torch.Tensor] = None
    metadata_embeds: Optional[torch.FloatTensor] = None
    beatmap_embeds: Optional[torch.FloatTensor] = None
    logits: Optional[torch.FloatTensor] = None
    metadata_model_output: BaseModelOutputWithPooling = None
    beatmap_model_output: BaseModelOutputWithPooling = None

    def to_tuple(self) -> tuple[Any]:
        return tuple(
            self[k] if k not in ["metadata_model_output", "beatmap_model_output"] else getattr(self, k).to_tuple()
            for k in self.keys()
        )


# ------------------------------------------------------------------------------------------------
# parameter containers with the reference's (ModernBERT's) key names

class _Embeddings(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.tok_embeddings = nn.Embedding(cfg.vocab_size, cfg.hidden_size)
        self.norm = nn.LayerNorm(cfg.hidden_size, eps=cfg.norm_eps, bias=cfg.norm_bias)


class _Attention(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.Wqkv = nn.Linear(cfg.hidden_size, 3 * cfg.hidden_size, bias=cfg.attention_bias)
        self.Wo = nn.Linear(cfg.hidden_size, cfg.hidden_size, bias=cfg.attention_bias)


class _MLP(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.Wi = nn.Linear(cfg.hidden_size, 2 * int(cfg.intermediate_size), bias=cfg.mlp_bias)
        self.Wo = nn.Linear(int(cfg.intermediate_size), cfg.hidden_size, bias=cfg.mlp_bias)


class _Layer(nn.Module):
    def __init__(self, cfg, layer_idx: int):
        super().__init__()
        # layer 0 has no attn_norm (ModernBERT uses nn.Identity there, MB:317-320)
        self.attn_norm = nn.Identity() if layer_idx == 0 else nn.LayerNorm(cfg.hidden_size, eps=cfg.norm_eps,
                                                                           bias=cfg.norm_bias)
        self.attn = _Attention(cfg)
        self.mlp_norm = nn.LayerNorm(cfg.hidden_size, eps=cfg.norm_eps, bias=cfg.norm_bias)
        self.mlp = _MLP(cfg)


def _check_supported(cfg) -> None:
    problems = []
    if cfg.hidden_size % cfg.num_attention_heads or cfg.hidden_size // cfg.num_attention_heads != 64:
        problems.append("head_dim must be 64")
    if cfg.norm_bias or cfg.attention_bias or cfg.mlp_bias:
        problems.append("norm/attention/mlp biases are not supported (reference configs have none)")
    if cfg.hidden_activation != "gelu":
        problems.append("hidden_activation must be 'gelu' (exact erf GELU)")
    if int(cfg.intermediate_size) % 16 or cfg.hidden_size % 64:
        problems.append("intermediate_size % 16 and hidden_size % 64 required")
    if any(getattr(cfg, k, 0.0) for k in ("attention_dropout", "embedding_dropout", "mlp_dropout")):
        problems.append("dropout must be 0 (reference configs)")
    if problems:
        raise ValueError(f"{type(cfg).__name__} is outside what the sm_100a kernels implement: " + "; ".join(problems))


class ModernBertEncoder(nn.Module):
    """Stands where `transformers.ModernBertModel` stands in the reference (attribute `encoder`).

    Holds the parameters under ModernBERT's names and runs the trunk on an *unpadded* token matrix:
        x0 = LN(E); per layer  x += Wo.Attn(RoPE(Wqkv.LN_a(x)));  x += Wo2.(gelu(u)*g);  out = LN_f(x)
    (SURVEY.md §8a "Encoder math spec"; MB:313-342, MB:446-490).
    """

    def __init__(self, cfg):
        super().__init__()
        _check_supported(cfg)
        self.config = cfg
        self.embeddings = _Embeddings(cfg)
        self.layers = nn.ModuleList([_Layer(cfg, i) for i in range(cfg.num_hidden_layers)])
        self.final_norm = nn.LayerNorm(cfg.hidden_size, eps=cfg.norm_eps, bias=cfg.norm_bias)
        self._packed = None
        self._packed_key = None

    def get_input_embeddings(self):
        return self.embeddings.tok_embeddings

    def set_input_embeddings(self, value):
        self.embeddings.tok_embeddings = value

    # -- bf16 working copies of the weights in the layouts the kernels want --------------------
    def packed(self):
        key = tuple((p.data_ptr(), p._version, p.dtype) for p in self.parameters())
        if self._packed is None or key != self._packed_key:
            with torch.no_grad():
                bf, f32 = torch.bfloat16, torch.float32
                pk = {
                    "tok_emb": self.embeddings.tok_embeddings.weight.detach().to(bf).contiguous(),
                    "emb_norm": self.embeddings.norm.weight.detach().to(f32).contiguous(),
                    "final_norm": self.final_norm.weight.detach().to(f32).contiguous(),
                    "layers": [],
                }
                for i, layer in enumerate(self.layers):
                    e = {
                        "attn_norm": None if i == 0 else layer.attn_norm.weight.detach().to(f32).contiguous(),
                        "wqkv": layer.attn.Wqkv.weight.detach().to(bf).contiguous(),
                        "wo": layer.attn.Wo.weight.detach().to(bf).contiguous(),
                        "mlp_norm": layer.mlp_norm.weight.detach().to(f32).contiguous(),
                        "wi": ops.interleave_wi(layer.mlp.Wi.weight.detach().to(bf)).contiguous(),
                        "wo2": layer.mlp.Wo.weight.detach().to(bf).contiguous(),
                    }
                    if ops.FUSE_LAYERNORM:
                        # LayerNorm folded into the GEMM that consumes it (cm3p_gemm_bf16_ln): W' = W.diag(gamma)
                        # in bf16 and its row sums (of the rounded W', which is what the tensor core multiplies)
                        if i > 0:
                            wq = (layer.attn.Wqkv.weight.detach().float() * e["attn_norm"][None, :]).to(bf).contiguous()
                            e["wqkv_ln"], e["cqkv"] = wq, wq.float().sum(dim=1).contiguous()
                        wi = (layer.mlp.Wi.weight.detach().float() * e["mlp_norm"][None, :]).to(bf)
                        wi = ops.interleave_wi(wi).contiguous()
                        e["wi_ln"], e["ci"] = wi, wi.float().sum(dim=1).contiguous()
                    pk["layers"].append(e)
            self._packed, self._packed_key = pk, key
        return self._packed

    def run_layers(self, x: torch.Tensor, cu_seqlens: torch.Tensor, max_seqlen: int,
                   positions: torch.Tensor, groups=None) -> torch.Tensor:
        """x [T,H] bf16 = already-normalised embeddings (updated in place) -> final-normed [T,H]."""
        cfg, pk = self.config, self.packed()
        T, H = x.shape
        heads = cfg.num_attention_heads
        dev = x.device
        tab_g = ops.rope_table(cfg.global_rope_theta, cfg.max_position_embeddings, dev)
        tab_l = ops.rope_table(cfg.local_rope_theta, cfg.max_position_embeddings, dev)
        a = torch.empty_like(x)
        qkv = torch.empty((T, 3 * H), device=dev, dtype=torch.bfloat16)
        h = torch.empty((T, int(cfg.intermediate_size)), device=dev, dtype=torch.bfloat16)
        eps = cfg.norm_eps
        n_layers = len(pk["layers"])
        if ops.FUSE_LAYERNORM:
            # The two pre-norm LayerNorms of every block are folded into the GEMMs around them: the residual
            # GEMM that writes x also accumulates its row statistics, the consuming GEMM applies them.
            stats = torch.empty((2 * n_layers, (H + 255) // 256, T, 2), device=dev, dtype=torch.float32)
        for i, w in enumerate(pk["layers"]):
            is_global = cfg.layer_is_global(i)
            tab = tab_g if is_global else tab_l
            if i == 0:
                ops.gemm(x, w["wqkv"], epilogue=ops.EPI_ROPE, out=qkv, positions=positions, rope_table=tab,
                         rope_cols=2 * H)
            elif ops.FUSE_LAYERNORM:
                ops.gemm(x, w["wqkv_ln"], epilogue=ops.EPI_ROPE, out=qkv, positions=positions, rope_table=tab,
                         rope_cols=2 * H, row_stats=stats[2 * i - 1], col_corr=w["cqkv"], ln_eps=eps)
            else:
                ops.layernorm(x, w["attn_norm"], eps, out=a)
                ops.gemm(a, w["wqkv"], epilogue=ops.EPI_ROPE, out=qkv, positions=positions, rope_table=tab,
                         rope_cols=2 * H)
            ops.attn_varlen_fwd(qkv, cu_seqlens, max_seqlen, heads, -1 if is_global else cfg.window_half, out=a,
                                groups=groups)
            if ops.FUSE_LAYERNORM:
                ops.gemm(a, w["wo"], epilogue=ops.EPI_RESIDUAL, out=x, aux=x, stats_out=stats[2 * i])
                ops.gemm(x, w["wi_ln"], epilogue=ops.EPI_GEGLU, out=h, row_stats=stats[2 * i], col_corr=w["ci"],
                         ln_eps=eps)
                ops.gemm(h, w["wo2"], epilogue=ops.EPI_RESIDUAL, out=x, aux=x,
                         stats_out=stats[2 * i + 1] if i + 1 < n_layers else None)
            else:
                ops.gemm(a, w["wo"], epilogue=ops.EPI_RESIDUAL, out=x, aux=x)
                ops.layernorm(x, w["mlp_norm"], eps, out=a)
                ops.gemm(a, w["wi"], epilogue=ops.EPI_GEGLU, out=h)
                ops.gemm(h, w["wo2"], epilogue=ops.EPI_RESIDUAL, out=x, aux=x)
        return ops.layernorm(x, pk["final_norm"], eps, out=a)


# ------------------------------------------------------------------------------------------------
# batch preparation: padded (B, L) ids + mask -> unpadded token list (host plumbing, integer only)

@dataclass
class _Unpadded:
    src_index: torch.Tensor   # [T] int32 flat index into the padded (B*L) layout
    cu_seqlens: torch.Tensor  # [B+1] int32
    positions: torch.Tensor   # [T] int32 column of the token in its padded row (== reference position_ids)
    lens_cpu: list
    total: int
    max_len: int
    batch: int
    seq_len: int
    groups: Any = None        # ops.PackedGroups when the batch is made of many short sequences (metadata tower)


def _unpad(attention_mask: Optional[torch.Tensor], batch: int, seq_len: int, device, pack: bool = False) -> _Unpadded:
    """Same information as `_unpad_cm3p_input` (modeling_cm3p.py:65-103): one small D2H copy of the
    B sequence lengths is the only host sync of a forward pass."""
    if attention_mask is None:
        lens_cpu = [seq_len] * batch
        total = batch * seq_len
        src = torch.arange(total, device=device, dtype=torch.int32)
        cu = torch.arange(0, total + 1, seq_len, device=device, dtype=torch.int32)
    else:
        m = attention_mask.reshape(batch, seq_len)
        if m.dtype != torch.bool:
            m = m != 0
        lens = m.sum(dim=-1, dtype=torch.int32)
        lens_cpu = lens.tolist()
        total = int(sum(lens_cpu))
        cu = torch.zeros(batch + 1, device=device, dtype=torch.int32)
        cu[1:] = torch.cumsum(lens, dim=0, dtype=torch.int32)
        src = torch.nonzero_static(m.reshape(-1), size=total).reshape(-1).to(torch.int32)
    pos = torch.remainder(src, seq_len).to(torch.int32)
    up = _Unpadded(src, cu, pos, lens_cpu, total, max(lens_cpu) if lens_cpu else 0, batch, seq_len)
    if pack and seq_len <= 128 and batch >= 4 and 0 < total <= 64 * batch:
        # many short sequences (metadata: ~21 real tokens of 128): several of them share one 128-row attention tile
        up.groups = ops.attn_pack_groups(cu, total)
    return up


def _repad(x: torch.Tensor, up: _Unpadded) -> torch.Tensor:
    """(T, ...) -> zero-padded (B, L, ...) like `_pad_cm3p_output` (modeling_cm3p.py:106-134)."""
    out = x.new_zeros((up.batch * up.seq_len,) + tuple(x.shape[1:]))
    out.index_copy_(0, up.src_index.long(), x)
    return out.view(up.batch, up.seq_len, *x.shape[1:])


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"cm3p_b200: {what} is on {t.device}; this implementation only runs on a CUDA "
                           "sm_100a device (there is no CPU fallback)")


# ------------------------------------------------------------------------------------------------
# towers

class CM3PPreTrainedModel(PreTrainedModel):
    config_class = CM3PConfig
    base_model_prefix = "cm3p"
    supports_gradient_checkpointing = False
    _supports_flash_attn = True
    _supports_flash_attn_2 = True
    _supports_sdpa = True
    _supports_flex_attn = False

    def _init_weights(self, module):
        """Same distributions as the reference (modeling_cm3p.py:262-297; ModernBERT's own
        trunc-normal init for the encoder Linear/Embedding weights)."""
        cfg = self.config
        std = getattr(cfg, "initializer_range", 0.02)
        if isinstance(module, (nn.Linear, nn.Conv1d)):
            nn.init.normal_(module.weight, std=std)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            nn.init.trunc_normal_(module.weight, std=std, a=-2 * std, b=2 * std)
        elif isinstance(module, nn.LayerNorm):
            module.weight.data.fill_(1.0)
            if module.bias is not None:
                module.bias.data.zero_()
        if isinstance(module, CM3PModel):
            nn.init.normal_(module.metadata_projection.weight,
                            std=module.metadata_embed_dim ** -0.5 * cfg.initializer_factor)
            nn.init.normal_(module.beatmap_projection.weight,
                            std=module.beatmap_embed_dim ** -0.5 * cfg.initializer_factor)
            module.logit_scale.data.fill_(cfg.logit_scale_init_value)
        elif isinstance(module, (CM3PBeatmapModelWithProjection, CM3PMetadataModelWithProjection)):
            proj = getattr(module, "beatmap_projection", None) or getattr(module, "metadata_projection")
            nn.init.normal_(proj.weight, std=cfg.hidden_size ** -0.5 * cfg.initializer_factor)


def _out_dtype(module: nn.Module) -> torch.dtype:
    p = next(module.parameters())
    return p.dtype if p.dtype in (torch.bfloat16, torch.float16, torch.float32) else torch.float32


class CM3PMetadataTransformer(nn.Module):
    """reference: modeling_cm3p.py:300-403."""

    def __init__(self, config: CM3PMetadataConfig):
        super().__init__()
        self.config = config
        self.encoder = ModernBertEncoder(config)

    def get_input_embeddings(self):
        return self.encoder.get_input_embeddings()

    def set_input_embeddings(self, value):
        self.encoder.set_input_embeddings(value)

    def encode(self, input_ids, attention_mask):
        """-> (last_hidden unpadded [T,H] bf16, _Unpadded over the flattened (B*V, S) batch)."""
        _require_cuda(input_ids, "metadata input_ids")
        S = input_ids.shape[-1]
        ids = input_ids.reshape(-1, S).contiguous()
        up = _unpad(attention_mask, ids.shape[0], S, ids.device, pack=True)
        pk = self.encoder.packed()
        x = ops.embed_gather_ln(ids.reshape(-1), up.src_index, None, pk["tok_emb"], None, pk["emb_norm"],
                                self.config.norm_eps, rows=up.total)
        last = self.encoder.run_layers(x, up.cu_seqlens, up.max_len, up.positions, groups=up.groups)
        return last, up

    def forward(self, input_ids=None, attention_mask=None, indices=None, cu_seqlens=None, max_seqlen=None,
                batch_size=None, seq_len=None, output_attentions=None, output_hidden_states=None,
                output_pooler: bool = True) -> BaseModelOutputWithPooling:
        if input_ids is None:
            raise ValueError("You have to specify input_ids")
        last, up = self.encode(input_ids, attention_mask)
        pooled = None
        if output_pooler:
            pooled = ops.pool_project_normalize(last, up.cu_seqlens, not self.config.cls_embed, None)[0]
            pooled = pooled.view(*input_ids.shape[:-1], -1).to(_out_dtype(self))
        hidden = _repad(last, up).view(*input_ids.shape, -1).to(_out_dtype(self))
        return BaseModelOutputWithPooling(last_hidden_state=hidden, pooler_output=pooled)


class CM3PMultiModalProjector(nn.Module):
    def __init__(self, config: CM3PAudioConfig):
        super().__init__()
        if config.projector_hidden_act != "gelu":
            raise ValueError("projector_hidden_act must be 'gelu'")
        self.linear_1 = nn.Linear(config.projector_intermediate_size, config.projector_dim, bias=False)
        self.linear_2 = nn.Linear(config.projector_dim, config.projector_dim, bias=False)


class CM3PAudioEncoder(nn.Module):
    """reference: modeling_cm3p.py:484-528 — conv1d x2 (+GELU, implicit GEMM) -> ModernBERT(audio) -> 4:1 projector."""

    def __init__(self, config: CM3PAudioConfig):
        super().__init__()
        self.config = config
        self.conv1 = nn.Conv1d(config.n_mels, config.hidden_size, kernel_size=3, padding=1)
        self.conv2 = nn.Conv1d(config.hidden_size, config.hidden_size, kernel_size=3, stride=2, padding=1)
        self.encoder = ModernBertEncoder(config)
        self.multi_modal_projector = CM3PMultiModalProjector(config)
        self._packed = None
        self._packed_key = None

    def packed(self):
        ps = [self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias,
              self.multi_modal_projector.linear_1.weight, self.multi_modal_projector.linear_2.weight]
        key = tuple((p.data_ptr(), p._version, p.dtype) for p in ps)
        if self._packed is None or key != self._packed_key:
            with torch.no_grad():
                bf, f32 = torch.bfloat16, torch.float32
                c = self.config
                self._packed = {
                    # implicit-GEMM conv weights: K index = tap * c_pad + c_in (c_pad = c_in rounded up to 64)
                    "w1": ops.pack_conv_weight(self.conv1.weight),
                    "b1": self.conv1.bias.detach().to(f32).contiguous(),
                    "w2": ops.pack_conv_weight(self.conv2.weight),
                    "b2": self.conv2.bias.detach().to(f32).contiguous(),
                    "p1": self.multi_modal_projector.linear_1.weight.detach().to(bf).contiguous(),
                    "p2": self.multi_modal_projector.linear_2.weight.detach().to(bf).contiguous(),
                }
            self._packed_key = key
        return self._packed

    def forward(self, input_features: torch.Tensor, output_attentions=None,
                output_hidden_states=None) -> CM3PAudioModelOutput:
        _require_cuda(input_features, "input_features")
        cfg, pk = self.config, self.packed()
        B, _, Fr = input_features.shape
        feats = ops.transpose_cast(input_features.float().contiguous())   # [B, F, n_mels] bf16 channels-last
        y1 = ops.conv1d_k3(feats, pk["w1"], pk["b1"], stride=1, gelu=True)  # [B, F, H_a] channels-last
        y2 = ops.conv1d_k3(y1, pk["w2"], pk["b2"], stride=2, gelu=True)     # [B, F/2, H_a]
        T2 = Fr // 2
        dev = feats.device
        x = ops.layernorm(y2.view(B * T2, cfg.hidden_size), self.encoder.packed()["emb_norm"], cfg.norm_eps)
        cu = torch.arange(0, B * T2 + 1, T2, device=dev, dtype=torch.int32)
        pos = torch.remainder(torch.arange(B * T2, device=dev, dtype=torch.int32), T2).to(torch.int32)
        last = self.encoder.run_layers(x, cu, T2, pos)                     # [B*T2, H_a]
        grouped = last.view(-1, cfg.projector_intermediate_size)           # 4 consecutive frames per token
        hmid = ops.gemm(grouped, pk["p1"], epilogue=ops.EPI_GELU)
        audio_embeds = ops.gemm(hmid, pk["p2"])
        return CM3PAudioModelOutput(audio_embeds=audio_embeds, last_hidden_state=last.view(B, T2, -1))


class CM3PBeatmapTransformer(nn.Module):
    """reference: modeling_cm3p.py:531-650."""

    def __init__(self, config: CM3PBeatmapConfig):
        super().__init__()
        self.config = config
        self.audio_encoder = CM3PAudioEncoder(config.audio_config)
        self.encoder = ModernBertEncoder(config)

    def get_input_embeddings(self):
        return self.encoder.get_input_embeddings()

    def set_input_embeddings(self, value):
        self.encoder.set_input_embeddings(value)

    def encode(self, input_ids, input_features, attention_mask, up=None):
        """-> (last_hidden unpadded [T,H] bf16, _Unpadded, audio output | None)."""
        _require_cuda(input_ids, "input_ids")
        B, L = input_ids.shape
        dev = input_ids.device
        if up is None:
            up = _unpad(attention_mask, B, L, dev)
        ids_flat = input_ids.reshape(-1).contiguous()
        audio_out, audio_embeds, slot = None, None, None
        if input_features is not None:
            audio_out = self.audio_encoder(input_features)
            audio_embeds = audio_out.audio_embeds
            # running count of [AUDIO] tokens in row-major order == the reference's boolean-mask
            # scatter `inputs_embeds[input_ids == audio_token_id] = audio_embeds` (:603-605)
            is_audio = ids_flat == self.config.audio_token_id
            running = torch.cumsum(is_audio, dim=0, dtype=torch.int32) - 1
            slot = torch.where(is_audio, running, torch.full_like(running, -1))
            slot = slot.index_select(0, up.src_index.long()).contiguous()
        pk = self.encoder.packed()
        x = ops.embed_gather_ln(ids_flat, up.src_index, slot, pk["tok_emb"], audio_embeds, pk["emb_norm"],
                                self.config.norm_eps, rows=up.total)
        last = self.encoder.run_layers(x, up.cu_seqlens, up.max_len, up.positions)
        return last, up, audio_out

    def forward(self, input_ids=None, input_features=None, attention_mask=None, sliding_window_mask=None,
                position_ids=None, inputs_embeds=None, indices=None, cu_seqlens=None, max_seqlen=None,
                batch_size=None, seq_len=None, output_attentions=None, output_hidden_states=None,
                output_pooler: bool = True, pad_output: bool = True) -> CM3PBeatmapModelOutput:
        if input_ids is None:
            raise ValueError("cm3p_b200 needs input_ids (inputs_embeds-only calls are not supported)")
        if inputs_embeds is not None or position_ids is not None:
            raise NotImplementedError("custom inputs_embeds / position_ids are not supported by the fused kernels")
        last, up, audio_out = self.encode(input_ids, input_features, attention_mask)
        pooled = None
        if output_pooler:
            pooled = ops.pool_project_normalize(last, up.cu_seqlens, not self.config.cls_embed, None)[0]
            pooled = pooled.to(_out_dtype(self))
        hidden = (_repad(last, up) if pad_output else last).to(_out_dtype(self))
        return CM3PBeatmapModelOutput(last_hidden_state=hidden, pooler_output=pooled, audio_model_output=audio_out)


class CM3PMetadataModel(CM3PPreTrainedModel):
    config_class = CM3PMetadataConfig
    main_input_name = "input_ids"

    def __init__(self, config: CM3PMetadataConfig):
        super().__init__(config)
        self.metadata_model = CM3PMetadataTransformer(config)
        self.post_init()

    def get_input_embeddings(self):
        return self.metadata_model.get_input_embeddings()

    def set_input_embeddings(self, value):
        self.metadata_model.set_input_embeddings(value)

    def forward(self, input_ids=None, attention_mask=None, output_attentions=None, output_hidden_states=None):
        return self.metadata_model(input_ids=input_ids, attention_mask=attention_mask)


class CM3PBeatmapModel(CM3PPreTrainedModel):
    config_class = CM3PBeatmapConfig
    main_input_name = "input_ids"

    def __init__(self, config: CM3PBeatmapConfig):
        super().__init__(config)
        self.beatmap_model = CM3PBeatmapTransformer(config)
        self.post_init()

    def get_input_embeddings(self):
        return self.beatmap_model.get_input_embeddings()

    def set_input_embeddings(self, value):
        self.beatmap_model.set_input_embeddings(value)

    def forward(self, input_ids=None, input_features=None, attention_mask=None, position_ids=None,
                inputs_embeds=None, output_attentions=None, output_hidden_states=None):
        return self.beatmap_model(input_ids=input_ids, input_features=input_features, attention_mask=attention_mask,
                                  position_ids=position_ids, inputs_embeds=inputs_embeds)


# ------------------------------------------------------------------------------------------------
# the dual-tower model

def _pack_linear(module: nn.Linear, cache: dict, name: str) -> torch.Tensor:
    w = module.weight
    key = (w.data_ptr(), w._version, w.dtype)
    hit = cache.get(name)
    if hit is None or hit[0] != key:
        cache[name] = (key, w.detach().to(torch.bfloat16).contiguous())
    return cache[name][1]


class CM3PModel(CM3PPreTrainedModel):
    """reference: modeling_cm3p.py:727-1012 (same constructor members, same forward keywords)."""

    config_class = CM3PConfig

    def __init__(self, config: CM3PConfig):
        super().__init__(config)
        if not isinstance(config.metadata_config, CM3PMetadataConfig):
            raise TypeError("config.metadata_config is expected to be of type CM3PMetadataConfig but is of type"
                            f" {type(config.metadata_config)}.")
        if not isinstance(config.beatmap_config, CM3PBeatmapConfig):
            raise TypeError("config.beatmap_config is expected to be of type CM3PBeatmapConfig but is of type"
                            f" {type(config.beatmap_config)}.")
        mc, bc = config.metadata_config, config.beatmap_config
        self.projection_dim = config.projection_dim
        self.metadata_embed_dim = mc.hidden_size
        self.beatmap_embed_dim = bc.hidden_size
        self.loss_type = config.loss_type
        self.metadata_model = CM3PMetadataTransformer(mc)
        self.beatmap_model = CM3PBeatmapTransformer(bc)
        self.beatmap_projection = nn.Linear(self.beatmap_embed_dim, self.projection_dim, bias=False)
        self.metadata_projection = nn.Linear(self.metadata_embed_dim, self.projection_dim, bias=False)
        self.logit_scale = nn.Parameter(torch.tensor(float(config.logit_scale_init_value)))
        if config.has_decoder_head:
            self.head = CM3PPredictionHead(bc)
            self.decoder = nn.Linear(bc.hidden_size, bc.vocab_size, bias=bc.decoder_bias)
        self._wcache: dict = {}
        self.post_init()

    # -- feature helpers (reference :773-841): projection without L2 normalisation --------------
    def get_metadata_features(self, input_ids=None, output_attentions=None, output_hidden_states=None):
        last, up = self.metadata_model.encode(input_ids, None)
        w = _pack_linear(self.metadata_projection, self._wcache, "mp")
        _, proj, _, _, _ = ops.pool_project_normalize(last, up.cu_seqlens, not self.config.metadata_config.cls_embed, w)
        return proj.view(*input_ids.shape[:-1], -1).to(_out_dtype(self))

    def get_beatmap_features(self, input_ids=None, input_features=None, attention_mask=None, position_ids=None,
                             inputs_embeds=None, output_attentions=None, output_hidden_states=None):
        last, up, _ = self.beatmap_model.encode(input_ids, input_features, attention_mask)
        w = _pack_linear(self.beatmap_projection, self._wcache, "bp")
        _, proj, _, _, _ = ops.pool_project_normalize(last, up.cu_seqlens, not self.config.beatmap_config.cls_embed, w)
        return proj.to(_out_dtype(self))

    def forward(self, input_ids=None, input_features=None, metadata_ids=None, attention_mask=None,
                metadata_attention_mask=None, position_ids=None, inputs_embeds=None,
                metadata_variation_classes=None, labels=None, indices=None, cu_seqlens=None, max_seqlen=None,
                batch_size=None, seq_len=None, return_loss: Optional[bool] = True, output_attentions=None,
                output_hidden_states=None, output_logits=None, **kwargs) -> CM3POutput:
        cfg = self.config
        output_logits = output_logits if output_logits is not None else cfg.has_decoder_head
        if (metadata_ids is not None and metadata_ids.dim() == 3 and return_loss
                and metadata_variation_classes is None):
            raise ValueError("When providing multiple metadata variations, metadata_variation_classes must be "
                             "provided in order to compute loss correctly.")
        if output_logits and not cfg.has_decoder_head:
            raise ValueError("Cannot return logits when the model is not configured with a decoder head.")
        if inputs_embeds is not None or position_ids is not None or indices is not None or cu_seqlens is not None:
            raise NotImplementedError("cm3p_b200 unpads internally; pass padded input_ids + attention_mask")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .training import forward_with_grad  # explicit-backward training path
            return forward_with_grad(self, input_ids=input_ids, input_features=input_features,
                                     metadata_ids=metadata_ids, attention_mask=attention_mask,
                                     metadata_attention_mask=metadata_attention_mask,
                                     metadata_variation_classes=metadata_variation_classes, labels=labels,
                                     return_loss=return_loss, output_logits=output_logits, **kwargs)

        odt = _out_dtype(self)
        # the reference returns unpadded hidden states under flash_attention_2 (quirk Q6) and padded
        # ones otherwise; the arithmetic is identical, only the output formatting differs
        pad_outputs = getattr(cfg, "_attn_implementation", None) != "flash_attention_2"
        beatmap_embeds = metadata_embeds = logits_per_beatmap = logits_per_metadata = logits = None
        beatmap_outputs = metadata_outputs = None
        loss = 0 if return_loss else None
        be16 = me16 = mlm_loss = None

        if input_ids is not None:
            last, up, audio_out = self.beatmap_model.encode(input_ids, input_features, attention_mask)
            w = _pack_linear(self.beatmap_projection, self._wcache, "bp")
            pooled, _, _, be32, be16 = ops.pool_project_normalize(last, up.cu_seqlens,
                                                                  not cfg.beatmap_config.cls_embed, w)
            beatmap_embeds = be32.to(odt)
            hidden = (_repad(last, up) if pad_outputs else last).to(odt)
            beatmap_outputs = CM3PBeatmapModelOutput(last_hidden_state=hidden, pooler_output=pooled.to(odt),
                                                     audio_model_output=audio_out)
            if output_logits:
                logits_u = self._mlm_logits(last)
                if labels is not None and return_loss:
                    # loss += 0.5 * MLM cross-entropy (:994-996), also on the evaluation path
                    _, loss_sum, count = ops.vocab_ce_fwd(logits_u, logits_u.shape[1], labels.reshape(-1).contiguous(),
                                                          up.src_index)
                    n_items = kwargs.get("num_items_in_batch")
                    denom = count if n_items is None else torch.as_tensor(n_items, device=count.device,
                                                                          dtype=torch.float32).reshape(1)
                    mlm_loss = (loss_sum / denom).reshape(())
                logits = _repad(logits_u, up).to(odt)

        if metadata_ids is not None:
            mlast, mup = self.metadata_model.encode(metadata_ids, metadata_attention_mask)
            w = _pack_linear(self.metadata_projection, self._wcache, "mp")
            mpooled, _, _, me32, me16 = ops.pool_project_normalize(mlast, mup.cu_seqlens,
                                                                   not cfg.metadata_config.cls_embed, w)
            lead = tuple(metadata_ids.shape[:-1])
            metadata_embeds = me32.view(*lead, -1).to(odt)
            mhidden = _repad(mlast, mup).view(*metadata_ids.shape, -1).to(odt) if pad_outputs else mlast.to(odt)
            metadata_outputs = BaseModelOutputWithPooling(last_hidden_state=mhidden,
                                                          pooler_output=mpooled.view(*lead, -1).to(odt))

        if be16 is not None and me16 is not None:
            # S = M . B^T * exp(logit_scale), the factor read on the device in the GEMM epilogue (:976-977)
            S = ops.gemm(me16, be16, epilogue=ops.EPI_SCALE_F32,
                         aux=self.logit_scale.detach().float().reshape(1))  # [Bm*V, Bb] fp32
            Bb = be16.shape[0]
            if metadata_ids.dim() == 3:
                Bm, V = metadata_ids.shape[:2]
                logits_per_metadata = S.view(Bm, V, Bb)
                logits_per_beatmap = logits_per_metadata.permute(2, 0, 1)
            else:
                Bm, V = metadata_ids.shape[0], 1
                logits_per_metadata = S
                logits_per_beatmap = S.t()
            if return_loss:
                if Bm != Bb:
                    raise ValueError(f"metadata batch {Bm} != beatmap batch {Bb}")
                if V > 1:
                    true_idx = (metadata_variation_classes == 0).int().argmax(dim=1).to(torch.int32)
                else:
                    true_idx = torch.zeros(Bm, device=S.device, dtype=torch.int32)
                loss = ops.clip_loss_fwd(S, true_idx.contiguous(), V)[0].reshape(())

        if mlm_loss is not None:
            loss = loss + 0.5 * mlm_loss

        return CM3POutput(loss=loss, logits_per_beatmap=logits_per_beatmap, logits_per_metadata=logits_per_metadata,
                          metadata_embeds=metadata_embeds, beatmap_embeds=beatmap_embeds, logits=logits,
                          metadata_model_output=metadata_outputs, beatmap_model_output=beatmap_outputs)

    def _mlm_logits(self, last: torch.Tensor) -> torch.Tensor:
        """decoder(norm(gelu(dense(h)))) on unpadded rows (reference :987-993, :1229-1238)."""
        bc = self.config.beatmap_config
        wd = _pack_linear(self.head.dense, self._wcache, "hd")
        wv = _pack_linear(self.decoder, self._wcache, "dec")
        y = ops.gemm(last, wd, epilogue=ops.EPI_GELU)
        y = ops.layernorm(y, self.head.norm.weight.detach().float().contiguous(), bc.norm_eps)
        if self.decoder.bias is not None:
            return ops.gemm(y, wv, epilogue=ops.EPI_BIAS, aux=self.decoder.bias.detach().float().contiguous())
        return ops.gemm(y, wv)


class CM3PPredictionHead(nn.Module):
    def __init__(self, config: CM3PBeatmapConfig):
        super().__init__()
        if config.classifier_bias or config.classifier_activation != "gelu":
            raise ValueError("prediction head: classifier_bias=False and classifier_activation='gelu' required")
        self.config = config
        self.dense = nn.Linear(config.hidden_size, config.hidden_size, config.classifier_bias)
        self.norm = nn.LayerNorm(config.hidden_size, eps=config.norm_eps, bias=config.norm_bias)


class CM3PMetadataModelWithProjection(CM3PPreTrainedModel):
    """reference :1016-1066 — projection WITHOUT L2 normalisation."""
    config_class = CM3PMetadataConfig

    def __init__(self, config: CM3PMetadataConfig):
        super().__init__(config)
        self.metadata_model = CM3PMetadataTransformer(config)
        self.metadata_projection = nn.Linear(config.hidden_size, config.projection_dim, bias=False)
        self._wcache: dict = {}
        self.post_init()

    def forward(self, input_ids=None, attention_mask=None, output_attentions=None, output_hidden_states=None):
        last, up = self.metadata_model.encode(input_ids, attention_mask)
        w = _pack_linear(self.metadata_projection, self._wcache, "mp")
        _, proj, _, _, _ = ops.pool_project_normalize(last, up.cu_seqlens, not self.config.cls_embed, w)
        odt = _out_dtype(self)
        return CM3PMetadataModelOutput(metadata_embeds=proj.view(*input_ids.shape[:-1], -1).to(odt),
                                       last_hidden_state=_repad(last, up).view(*input_ids.shape, -1).to(odt))


class CM3PBeatmapModelWithProjection(CM3PPreTrainedModel):
    """reference :1069-1128 — projection WITHOUT L2 normalisation."""
    config_class = CM3PBeatmapConfig

    def __init__(self, config: CM3PBeatmapConfig):
        super().__init__(config)
        self.beatmap_model = CM3PBeatmapTransformer(config)
        self.beatmap_projection = nn.Linear(config.hidden_size, config.projection_dim, bias=False)
        self._wcache: dict = {}
        self.post_init()

    def forward(self, input_ids=None, input_features=None, attention_mask=None, position_ids=None,
                inputs_embeds=None, output_attentions=None, output_hidden_states=None):
        last, up, _ = self.beatmap_model.encode(input_ids, input_features, attention_mask)
        w = _pack_linear(self.beatmap_projection, self._wcache, "bp")
        pooled, proj, _, _, _ = ops.pool_project_normalize(last, up.cu_seqlens, not self.config.cls_embed, w)
        odt = _out_dtype(self)
        return CM3PBeatmapModelOutput(beatmap_embeds=proj.to(odt), pooler_output=pooled.to(odt),
                                      last_hidden_state=_repad(last, up).to(odt))


def _trains(module: nn.Module) -> bool:
    return torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters())


class CM3PForMaskedLM(CM3PPreTrainedModel):
    """reference :1241-1379 — beatmap tower + prediction head + decoder, MLM cross-entropy (ignore -100),
    optional `sparse_prediction` (only labelled positions go through the head)."""
    config_class = CM3PBeatmapConfig
    base_model_prefix = "beatmap_model"
    # The reference declares `_tied_weights_keys = ["decoder.weight"]` (:1244).  With the installed
    # transformers the tie only happens when `config.tie_word_embeddings` is true (default: untied, which is
    # what the oracle does); checkpoints with or without a separate decoder.weight load.
    _tied_weights_keys = {"decoder.weight": "beatmap_model.encoder.embeddings.tok_embeddings.weight"}

    def __init__(self, config: CM3PBeatmapConfig):
        super().__init__(config)
        self.config = config
        self.beatmap_model = CM3PBeatmapTransformer(config)
        self.head = CM3PPredictionHead(config)
        self.decoder = nn.Linear(config.hidden_size, config.vocab_size, bias=config.decoder_bias)
        self.sparse_prediction = config.sparse_prediction
        self.sparse_pred_ignore_index = config.sparse_pred_ignore_index
        self._wcache: dict = {}
        self.post_init()

    def get_output_embeddings(self):
        return self.decoder

    def set_output_embeddings(self, new_embeddings: nn.Linear):
        self.decoder = new_embeddings

    def get_input_embeddings(self):
        return self.beatmap_model.get_input_embeddings()

    def forward(self, input_ids=None, input_features=None, attention_mask=None, sliding_window_mask=None,
                position_ids=None, inputs_embeds=None, labels=None, indices=None, cu_seqlens=None, max_seqlen=None,
                batch_size=None, seq_len=None, output_attentions=None, output_hidden_states=None,
                **kwargs) -> MaskedLMOutput:
        if inputs_embeds is not None or position_ids is not None or indices is not None or cu_seqlens is not None:
            raise NotImplementedError("cm3p_b200 unpads internally; pass padded input_ids + attention_mask")
        from .training import masked_lm_forward
        loss, logits = masked_lm_forward(self, input_ids=input_ids, input_features=input_features,
                                         attention_mask=attention_mask, labels=labels, train=_trains(self), **kwargs)
        return MaskedLMOutput(loss=loss, logits=logits)


class CM3PForBeatmapClassification(CM3PPreTrainedModel):
    """reference :1137-1226 — classifier on the pooled beatmap representation (e.g. the ranked classifier)."""
    config_class = CM3PBeatmapConfig
    base_model_prefix = "beatmap_model"

    def __init__(self, config: CM3PBeatmapConfig):
        super().__init__(config)
        self.num_labels = config.num_labels
        self.beatmap_model = CM3PBeatmapTransformer(config)
        self.classifier = nn.Linear(config.hidden_size, config.num_labels) if config.num_labels > 0 else nn.Identity()
        self._wcache: dict = {}
        self.post_init()

    def forward(self, input_ids=None, input_features=None, attention_mask=None, position_ids=None,
                inputs_embeds=None, labels=None, output_attentions=None,
                output_hidden_states=None) -> BeatmapClassifierOutput:
        if inputs_embeds is not None or position_ids is not None:
            raise NotImplementedError("custom inputs_embeds / position_ids are not supported by the fused kernels")
        from .training import classification_forward
        loss, logits = classification_forward(self, input_ids=input_ids, input_features=input_features,
                                              attention_mask=attention_mask, labels=labels, train=_trains(self))
        return BeatmapClassifierOutput(loss=loss, logits=logits)


def _register():
    for cfg_cls, model_cls in ((CM3PMetadataConfig, CM3PMetadataModel), (CM3PBeatmapConfig, CM3PBeatmapModel),
                               (CM3PConfig, CM3PModel)):
        try:
            AutoModel.register(cfg_cls, model_cls)
        except ValueError:
            pass
    for auto_cls, model_cls in ((AutoModelForSequenceClassification, CM3PForBeatmapClassification),
                                (AutoModelForMaskedLM, CM3PForMaskedLM)):
        try:
            auto_cls.register(CM3PBeatmapConfig, model_cls)
        except ValueError:
            pass


_register()

__all__ = ["CM3PModel", "CM3PPreTrainedModel", "CM3PMetadataModel", "CM3PMetadataModelWithProjection",
           "CM3PBeatmapModel", "CM3PBeatmapModelWithProjection", "CM3PForBeatmapClassification", "CM3PForMaskedLM",
           "CM3POutput"]
