"""Explicit forward + backward of the contrastive train step on the sm_100a kernels.

The reference trains through torch autograd (`loss.backward()` inside
`transformers.Trainer.training_step`, reached from /root/reference/train.py:360-375).  Here the
forward pass keeps exactly the activations the backward kernels need, and the backward pass is a
hand-scheduled sequence of C-ABI calls (cm3p_b200.ops): dgrad / wgrad tcgen05 GEMMs, the varlen
attention backward, LayerNorm / GeGLU / GELU / pooling / loss backward kernels.  It is attached to
autograd through ONE `torch.autograd.Function` whose inputs are the trainable parameters and whose
output is the loss, so `loss.backward()`, gradient accumulation, `torch.nn.parallel.
DistributedDataParallel` hooks (gradient all-reduce over NCCL) and optimizers work unchanged.

Per encoder layer the forward keeps x_in, qkv (rotated), the attention output, the log-sum-exp,
x_mid and the GeGLU pre-activation (13.8 KB / token / layer for the beatmap tower); the GeGLU
product is recomputed in the backward pass, and so are the LayerNorm outputs unless there is
room to keep them (`CM3P_SAVE_LN`: "auto" keeps them when they fit in a quarter of the free
device memory, +3 KB / token / layer; "1" / "0" force it).

Numerics: bf16 activations and gradients of activations, fp32 accumulation, fp32 parameter
gradients (what autocast-bf16 training in the reference produces up to rounding; north-star
tolerance: loss / gradient norms within 1e-2 relative).
"""
from __future__ import annotations

import torch

from . import distributed as dp_utils
from . import ops

BF16, F32 = torch.bfloat16, torch.float32


class GradStore:
    """fp32 gradient buffers, one per parameter, carved out of a single zero-filled allocation.

    Data parallel: the buffer is laid out in the order in which the backward pass first touches the parameters
    (recorded on the first step: every parameter gradient is written by exactly one launch, so a prefix of that
    order is final as soon as a later parameter is touched).  It is cut into buckets of ~`BUCKET_BYTES`; a bucket's
    NCCL all-reduce is enqueued (async, on NCCL's own stream, ordered after the launches already on the compute
    stream) the moment the backward pass moves past it, so the reduction of the metadata tower / MLM head / upper
    beatmap layers travels over NVLink while the lower layers are still being differentiated.
    """

    BUCKET_BYTES = 64 << 20
    OVERLAP = True  # False: one reduction of the whole buffer after the backward pass (A/B measurements)
    TAIL_EVENTS = None  # a list: finish() appends (start, end) CUDA events around its wait for the collectives (bench.py)

    def __init__(self, params, order=None, dp=None, sum_reduce=False):
        self.params = list(params)
        self.dp = dp if (dp is not None and dp.world_size > 1) else None
        self.sum_reduce = sum_reduce
        laid = self.params if order is None else [self.params[i] for i in order]
        offs, total = [], 0
        for p in laid:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4  # keep every view 16-byte aligned
        dev = self.params[0].device
        self.flat = torch.zeros(total, device=dev, dtype=F32)
        self.views = {id(p): self.flat[o:o + p.numel()].view(p.shape) for p, o in zip(laid, offs)}
        self.touched: list[int] = []       # first-touch order (indices into self.params), recorded when order is None
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._seen: set[int] = set()
        self._works = []
        self._hold = 0  # > 0: a section of the backward pass revisits parameters (chunked recompute): nothing is final
        # buckets over the laid-out order: (first position, end position, flat start, flat end)
        self._pos = ({id(p): k for k, p in enumerate(laid)}
                     if (order is not None and self.dp is not None and self.OVERLAP) else None)
        self._buckets, self._next = [], 0
        if self._pos is not None:
            start_k, start_o = 0, 0
            for k, (p, o) in enumerate(zip(laid, offs)):
                end_o = o + (p.numel() + 3) // 4 * 4
                if (end_o - start_o) * 4 >= self.BUCKET_BYTES or k == len(laid) - 1:
                    self._buckets.append((start_k, k + 1, start_o, end_o))
                    start_k, start_o = k + 1, end_o

    def __call__(self, p: torch.Tensor) -> torch.Tensor:
        pid = id(p)
        if pid not in self._seen:
            self._seen.add(pid)
            self.touched.append(self._index[pid])
            if self._pos is not None and self._hold == 0:
                # every bucket that lies entirely before this parameter is final: send it off
                k = self._pos[pid]
                while self._next < len(self._buckets) and self._buckets[self._next][1] <= k:
                    self._launch(self._buckets[self._next])
                    self._next += 1
        return self.views[pid]

    def hold(self) -> None:
        self._hold += 1

    def release(self) -> None:
        self._hold -= 1

    def _launch(self, bucket) -> None:
        _, _, o0, o1 = bucket
        self._works.append(dp_utils.reduce_gradients_async(self.flat[o0:o1], self.dp, self.sum_reduce))

    def finish(self) -> None:
        """Reduce whatever has not been sent yet and make the compute stream wait for all reductions."""
        if self.dp is None:
            return
        tail = GradStore.TAIL_EVENTS
        if tail is not None:  # how long the compute stream stalls for collectives the backward pass could not hide
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
        if self._pos is None:
            self._works.append(dp_utils.reduce_gradients_async(self.flat, self.dp, self.sum_reduce))
        else:
            while self._next < len(self._buckets):
                self._launch(self._buckets[self._next])
                self._next += 1
        for w in self._works:
            w.wait()
        self._works = []
        if tail is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            tail.append((e0, e1))

    def order(self) -> list[int]:
        """Parameter indices in first-touch order, untouched (frozen / unused) parameters last."""
        rest = [i for i in range(len(self.params)) if i not in set(self.touched)]
        return self.touched + rest


def _wgrad(dy: torch.Tensor, x: torch.Tensor, out: torch.Tensor) -> None:
    """out[N_out, K_in] (fp32) += dy[T, N_out]^T . x[T, K_in]  (K of the GEMM = tokens, split across CTAs)."""
    ops.gemm(dy, x, trans_a=True, trans_b=True, epilogue=ops.EPI_SCALE_F32, accumulate=True, out=out)


def _dgrad(dy: torch.Tensor, w: torch.Tensor, out: torch.Tensor | None = None, **kw) -> torch.Tensor:
    """dx[T, K_in] = dy[T, N_out] . w[N_out, K_in]."""
    return ops.gemm(dy, w, trans_b=True, out=out, **kw)


# ------------------------------------------------------------------------------------------------
# ModernBERT trunk

def _keep_layernorm_outputs(nbytes: int, dev) -> bool:
    mode = ops.SAVE_LAYERNORM
    if mode in ("0", "1"):
        return mode == "1"
    free, _ = torch.cuda.mem_get_info(dev)
    free += torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)  # cached blocks are reusable
    return nbytes < free // 4


def encoder_forward(enc, x0: torch.Tensor, cu_seqlens, max_seqlen: int, positions, groups=None):
    """x0 [T,H] = normalised embeddings.  -> (final-normed hidden [T,H], saved activations)."""
    cfg, pk = enc.config, enc.packed()
    T, H = x0.shape
    heads, I = cfg.num_attention_heads, int(cfg.intermediate_size)
    dev = x0.device
    tab_g = ops.rope_table(cfg.global_rope_theta, cfg.max_position_embeddings, dev)
    tab_l = ops.rope_table(cfg.local_rope_theta, cfg.max_position_embeddings, dev)
    a = torch.empty_like(x0)
    layers = []
    x = x0
    eps, n_layers = cfg.norm_eps, len(pk["layers"])
    fuse = ops.FUSE_LAYERNORM  # LayerNorms folded into the neighbouring GEMMs (the backward recomputes them from x)
    if fuse:
        stats = torch.empty((2 * n_layers, (H + 255) // 256, T, 2), device=dev, dtype=F32)
    # keep the LayerNorm outputs for the weight-gradient GEMMs instead of recomputing them in the backward
    keep_ln = (not fuse) and _keep_layernorm_outputs(2 * n_layers * T * H * 2, dev)
    for i, w in enumerate(pk["layers"]):
        is_global = cfg.layer_is_global(i)
        tab = tab_g if is_global else tab_l
        norm_a = norm_m = None
        if i == 0:
            qkv = ops.gemm(x, w["wqkv"], epilogue=ops.EPI_ROPE, positions=positions, rope_table=tab, rope_cols=2 * H)
        elif fuse:
            qkv = ops.gemm(x, w["wqkv_ln"], epilogue=ops.EPI_ROPE, positions=positions, rope_table=tab,
                           rope_cols=2 * H, row_stats=stats[2 * i - 1], col_corr=w["cqkv"], ln_eps=eps)
        else:
            if keep_ln:
                a = torch.empty_like(x0)
            ops.layernorm(x, w["attn_norm"], eps, out=a)
            qkv = ops.gemm(a, w["wqkv"], epilogue=ops.EPI_ROPE, positions=positions, rope_table=tab, rope_cols=2 * H)
            norm_a = a if keep_ln else None
        lse = torch.empty((heads, T), device=dev, dtype=F32)
        o = ops.attn_varlen_fwd(qkv, cu_seqlens, max_seqlen, heads, -1 if is_global else cfg.window_half, lse=lse,
                                groups=groups)
        ug = torch.empty((T, 2 * I), device=dev, dtype=BF16)
        if fuse:
            x1 = ops.gemm(o, w["wo"], epilogue=ops.EPI_RESIDUAL, aux=x, stats_out=stats[2 * i])
            h = ops.gemm(x1, w["wi_ln"], epilogue=ops.EPI_GEGLU_SAVE, c2=ug, row_stats=stats[2 * i], col_corr=w["ci"],
                         ln_eps=eps)
            x2 = ops.gemm(h, w["wo2"], epilogue=ops.EPI_RESIDUAL, aux=x1,
                          stats_out=stats[2 * i + 1] if i + 1 < n_layers else None)
        else:
            x1 = ops.gemm(o, w["wo"], epilogue=ops.EPI_RESIDUAL, aux=x)
            if keep_ln:
                a = torch.empty_like(x0)
            ops.layernorm(x1, w["mlp_norm"], eps, out=a)
            h = ops.gemm(a, w["wi"], epilogue=ops.EPI_GEGLU_SAVE, c2=ug)
            x2 = ops.gemm(h, w["wo2"], epilogue=ops.EPI_RESIDUAL, aux=x1)
            norm_m = a if keep_ln else None
        layers.append(dict(x_in=x, qkv=qkv, o=o, lse=lse, x1=x1, ug=ug, window=-1 if is_global else cfg.window_half,
                           tab=tab, norm_a=norm_a, norm_m=norm_m))
        x = x2
    last = ops.layernorm(x, pk["final_norm"], cfg.norm_eps)
    saved = dict(layers=layers, x_final=x, cu=cu_seqlens, max_seqlen=max_seqlen, positions=positions, groups=groups)
    return last, saved


def encoder_backward(enc, saved, dlast: torch.Tensor, g: GradStore) -> torch.Tensor:
    """dlast = gradient of the final-normed hidden states.  Accumulates every weight gradient of the
    trunk into `g` and returns the gradient w.r.t. x0 (the normalised embeddings)."""
    cfg, pk = enc.config, enc.packed()
    T, H = dlast.shape
    heads, I = cfg.num_attention_heads, int(cfg.intermediate_size)
    dev = dlast.device
    eps = cfg.norm_eps
    cu, max_seqlen, positions = saved["cu"], saved["max_seqlen"], saved["positions"]

    dx = ops.layernorm_bwd(saved["x_final"], dlast, pk["final_norm"], eps, dgamma=g(enc.final_norm.weight))
    saved["x_final"] = None
    # scratch reused by every layer
    dh = torch.empty((T, I), device=dev, dtype=BF16)
    dug = torch.empty((T, 2 * I), device=dev, dtype=BF16)
    h = torch.empty((T, I), device=dev, dtype=BF16)
    norm = torch.empty((T, H), device=dev, dtype=BF16)
    dtmp = torch.empty((T, H), device=dev, dtype=BF16)
    dqkv = torch.empty((T, 3 * H), device=dev, dtype=BF16)
    delta = torch.empty((heads, T), device=dev, dtype=F32)
    g_wi_il = torch.empty((2 * I, H), device=dev, dtype=F32)

    for i in reversed(range(len(pk["layers"]))):
        w, layer, s = pk["layers"][i], enc.layers[i], saved["layers"][i]
        # ---- MLP:  x2 = x1 + (gelu(u) * gate) Wo2^T,  [u | gate] = LN_m(x1) Wi^T
        _dgrad(dx, w["wo2"], out=dh)
        ops.geglu_bwd(s["ug"], dh, dug=dug, h=h)
        _wgrad(dx, h, g(layer.mlp.Wo.weight))
        norm_m = s["norm_m"]
        if norm_m is None:
            norm_m = ops.layernorm(s["x1"], w["mlp_norm"], eps, out=norm)
        g_wi_il.zero_()
        _wgrad(dug, norm_m, g_wi_il)
        g(layer.mlp.Wi.weight).add_(ops.deinterleave_wi(g_wi_il))
        _dgrad(dug, w["wi"], out=dtmp)
        ops.layernorm_bwd(s["x1"], dtmp, w["mlp_norm"], eps, dres=dx, dx=dx, dgamma=g(layer.mlp_norm.weight))
        # ---- attention:  x1 = x_in + Attn(RoPE(LN_a(x_in) Wqkv^T)) Wo^T
        _dgrad(dx, w["wo"], out=dtmp)  # d(attention output)
        _wgrad(dx, s["o"], g(layer.attn.Wo.weight))
        ops.attn_varlen_bwd(s["qkv"], s["o"], dtmp, s["lse"], cu, max_seqlen, heads, s["window"],
                            positions=positions, rope_table=s["tab"], dqkv=dqkv, delta=delta,
                            groups=saved["groups"])
        if i == 0:
            _wgrad(dqkv, s["x_in"], g(layer.attn.Wqkv.weight))
            _dgrad(dqkv, w["wqkv"], out=dx, epilogue=ops.EPI_RESIDUAL, aux=dx)
        else:
            norm_a = s["norm_a"]
            if norm_a is None:
                norm_a = ops.layernorm(s["x_in"], w["attn_norm"], eps, out=norm)
            _wgrad(dqkv, norm_a, g(layer.attn.Wqkv.weight))
            _dgrad(dqkv, w["wqkv"], out=dtmp)
            ops.layernorm_bwd(s["x_in"], dtmp, w["attn_norm"], eps, dres=dx, dx=dx,
                              dgamma=g(layer.attn_norm.weight))
        saved["layers"][i] = None  # release this layer's activations
    return dx


# ------------------------------------------------------------------------------------------------
# audio encoder (conv front-end + trunk + 4:1 projector)

def audio_forward(audio, input_features: torch.Tensor):
    cfg, pk = audio.config, audio.packed()
    B, _, Fr = input_features.shape
    C = cfg.hidden_size
    xt = ops.transpose_cast(input_features.float().contiguous())      # [B, F, n_mels] bf16 channels-last
    z1 = ops.conv1d_k3(xt, pk["w1"], pk["b1"], stride=1, gelu=False)  # pre-activations are kept for the backward
    y1 = ops.gelu_fwd(z1)
    z2 = ops.conv1d_k3(y1, pk["w2"], pk["b2"], stride=2, gelu=False)
    y2 = ops.gelu_fwd(z2)
    T2 = Fr // 2
    dev = xt.device
    x0 = ops.layernorm(y2.view(B * T2, C), audio.encoder.packed()["emb_norm"], cfg.norm_eps)
    cu = torch.arange(0, B * T2 + 1, T2, device=dev, dtype=torch.int32)
    pos = torch.remainder(torch.arange(B * T2, device=dev, dtype=torch.int32), T2).to(torch.int32)
    last, enc_saved = encoder_forward(audio.encoder, x0, cu, T2, pos)
    grouped = last.view(-1, cfg.projector_intermediate_size)
    zp = ops.gemm(grouped, pk["p1"])
    hp = ops.gelu_fwd(zp)
    audio_embeds = ops.gemm(hp, pk["p2"])
    saved = dict(xt=xt, z1=z1, y1=y1, z2=z2, y2=y2, enc=enc_saved, last=last, zp=zp, hp=hp, B=B, Fr=Fr)
    return audio_embeds, last.view(B, T2, -1), saved


def audio_backward(audio, saved, d_audio_embeds: torch.Tensor, g: GradStore) -> None:
    cfg, pk = audio.config, audio.packed()
    B, Fr, C = saved["B"], saved["Fr"], cfg.hidden_size
    proj = audio.multi_modal_projector
    _wgrad(d_audio_embeds, saved["hp"], g(proj.linear_2.weight))
    dhp = _dgrad(d_audio_embeds, pk["p2"])
    dzp = ops.gelu_bwd(saved["zp"], dhp)
    grouped = saved["last"].view(-1, cfg.projector_intermediate_size)
    _wgrad(dzp, grouped, g(proj.linear_1.weight))
    dlast = _dgrad(dzp, pk["p1"]).view(-1, C)
    dx0 = encoder_backward(audio.encoder, saved["enc"], dlast, g)
    T2 = Fr // 2
    dy2 = ops.layernorm_bwd(saved["y2"].view(B * T2, C), dx0, audio.encoder.packed()["emb_norm"], cfg.norm_eps,
                            dgamma=g(audio.encoder.embeddings.norm.weight))
    dz2 = ops.gelu_bwd(saved["z2"].view(B * T2, C), dy2)
    ops.colsum_f32(dz2, g(audio.conv2.bias))
    # conv2 weight gradient: implicit GEMM over (window, frame) against the shifted conv1 output
    g_w2p = torch.zeros_like(pk["w2"], dtype=F32)
    ops.conv1d_k3_wgrad(dz2.view(B, T2, C), saved["y1"], g_w2p, stride=2)
    g(audio.conv2.weight).add_(ops.unpack_conv_weight_grad(g_w2p, C))
    # conv2 input gradient: dA2 = dz2 . W2 (all three taps), scattered back onto the frames + conv1's GELU backward
    da2 = _dgrad(dz2, pk["w2"])
    dz1 = ops.conv2_col2im_gelu_bwd(da2, saved["z1"]).view(B * Fr, C)
    ops.colsum_f32(dz1, g(audio.conv1.bias))
    g_w1p = torch.zeros_like(pk["w1"], dtype=F32)
    ops.conv1d_k3_wgrad(dz1.view(B, Fr, C), saved["xt"], g_w1p, stride=1)
    g(audio.conv1.weight).add_(ops.unpack_conv_weight_grad(g_w1p, cfg.n_mels))


# ------------------------------------------------------------------------------------------------
# towers

def beatmap_forward(tower, input_ids, input_features, attention_mask, up=None):
    from .modeling_cm3p import _unpad
    B, L = input_ids.shape
    dev = input_ids.device
    if up is None:
        up = _unpad(attention_mask, B, L, dev)
    ids_flat = input_ids.reshape(-1).contiguous()
    audio_embeds = slot = audio_last = audio_saved = None
    if input_features is not None:
        audio_embeds, audio_last, audio_saved = audio_forward(tower.audio_encoder, input_features)
        is_audio = ids_flat == tower.config.audio_token_id
        running = torch.cumsum(is_audio, dim=0, dtype=torch.int32) - 1
        slot = torch.where(is_audio, running, torch.full_like(running, -1))
        slot = slot.index_select(0, up.src_index.long()).contiguous()
    pk = tower.encoder.packed()
    x0 = ops.embed_gather_ln(ids_flat, up.src_index, slot, pk["tok_emb"], audio_embeds, pk["emb_norm"],
                             tower.config.norm_eps, rows=up.total)
    last, enc_saved = encoder_forward(tower.encoder, x0, up.cu_seqlens, up.max_len, up.positions)
    saved = dict(ids=ids_flat, up=up, slot=slot, audio_embeds=audio_embeds, audio=audio_saved, enc=enc_saved)
    return last, up, audio_last, saved


# Saved activations of the beatmap tower (+ audio encoder) above which the train step switches to recompute: the
# forward pass runs without keeping anything (inference kernels), the loss and the gradient of the embeddings are
# formed on the whole batch, and the backward pass re-runs the forward with saving, window chunk by window chunk
# (GradCache-style; exact gradients, one extra forward).  BASELINE.json configs[3] (512 windows per GPU with global
# negatives) needs ~290 GB of saved activations otherwise.  None = 84 % of the device memory.
BEATMAP_SAVE_BUDGET = None


def _beatmap_save_budget(dev) -> int:
    if BEATMAP_SAVE_BUDGET is not None:
        return int(BEATMAP_SAVE_BUDGET)
    return int(0.84 * torch.cuda.get_device_properties(dev).total_memory)


def _beatmap_saved_bytes(tower, tokens: int, windows: int, frames: int) -> int:
    est = tokens * _encoder_saved_bytes_per_token(tower.config)
    ac = tower.audio_encoder.config
    return est + windows * (frames // 2) * (_encoder_saved_bytes_per_token(ac) + 10 * ac.hidden_size)


def beatmap_forward_auto(tower, input_ids, input_features, attention_mask):
    """beatmap_forward, or — when the saved activations would not fit — a forward pass that keeps nothing and leaves
    the saving to the chunked backward.  -> (last, up, audio_last, saved)."""
    from .modeling_cm3p import _unpad
    B, L = input_ids.shape
    frames = input_features.shape[-1] if input_features is not None else 0
    up = _unpad(attention_mask, B, L, input_ids.device)
    if _beatmap_saved_bytes(tower, up.total, B if input_features is not None else 0, frames) <= _beatmap_save_budget(
            input_ids.device):
        return beatmap_forward(tower, input_ids, input_features, attention_mask, up=up)
    last, up, audio_out = tower.encode(input_ids, input_features, attention_mask, up=up)
    audio_last = None if audio_out is None else audio_out.last_hidden_state
    saved = dict(chunked=True, up=up, input_ids=input_ids, input_features=input_features,
                 attention_mask=attention_mask, frames=frames)
    return last, up, audio_last, saved


def beatmap_backward(tower, saved, dlast, g: GradStore) -> None:
    if saved.get("chunked"):
        # recompute: forward with saving + backward for as many windows at a time as fit the budget; the weight
        # gradients of all chunks accumulate into the same buffers, so no gradient bucket may leave before the end
        import numpy as np
        up = saved["up"]
        ids, feats, mask = saved["input_ids"], saved["input_features"], saved["attention_mask"]
        B = ids.shape[0]
        cum = np.concatenate(([0], np.cumsum(np.asarray(up.lens_cpu, dtype=np.int64))))
        budget = _beatmap_save_budget(ids.device)
        g.hold()
        try:
            b0 = 0
            while b0 < B:
                b1 = b0 + 1
                while b1 < B and _beatmap_saved_bytes(tower, int(cum[b1 + 1] - cum[b0]), b1 + 1 - b0,
                                                      saved["frames"]) <= budget:
                    b1 += 1
                _, _, _, sv_c = beatmap_forward(tower, ids[b0:b1], None if feats is None else feats[b0:b1],
                                                None if mask is None else mask[b0:b1])
                beatmap_backward(tower, sv_c, dlast[int(cum[b0]):int(cum[b1])], g)
                del sv_c
                b0 = b1
        finally:
            g.release()
        return
    enc, up = tower.encoder, saved["up"]
    pk = enc.packed()
    dx0 = encoder_backward(enc, saved["enc"], dlast, g)
    d_audio = None
    if saved["audio_embeds"] is not None:
        d_audio = torch.zeros_like(saved["audio_embeds"])
    ops.embed_gather_ln_bwd(saved["ids"], up.src_index, saved["slot"], pk["tok_emb"], saved["audio_embeds"],
                            pk["emb_norm"], dx0, tower.config.norm_eps, d_tok_emb=g(enc.embeddings.tok_embeddings.weight),
                            d_audio_embeds=d_audio, dgamma=g(enc.embeddings.norm.weight))
    if d_audio is not None:
        audio_backward(tower.audio_encoder, saved["audio"], d_audio, g)


# Saved activations of the metadata tower above which its backward pass re-runs the forward in chunks of sequences
# instead of keeping every layer's activations alive next to the beatmap tower's.  At the reference's 256 metadata
# variations per beatmap (configs/train/v7.yaml:40) the tower sees 65 536 sequences = 1.4 M tokens per GPU: 42 GB
# of saved activations, on top of the ~145 GB of the beatmap tower.
METADATA_SAVE_BUDGET = 6 << 30


def _encoder_saved_bytes_per_token(cfg) -> int:
    H, I = cfg.hidden_size, int(cfg.intermediate_size)
    return cfg.num_hidden_layers * 2 * (8 * H + 2 * I)  # x_in, qkv, o, x1, ug + the two LayerNorm outputs


def metadata_forward(tower, metadata_ids, attention_mask):
    from .modeling_cm3p import _unpad
    S = metadata_ids.shape[-1]
    ids = metadata_ids.reshape(-1, S).contiguous()
    up = _unpad(attention_mask, ids.shape[0], S, ids.device, pack=True)
    pk = tower.encoder.packed()
    ids_flat = ids.reshape(-1)
    x0 = ops.embed_gather_ln(ids_flat, up.src_index, None, pk["tok_emb"], None, pk["emb_norm"], tower.config.norm_eps,
                             rows=up.total)
    if up.total * _encoder_saved_bytes_per_token(tower.config) > METADATA_SAVE_BUDGET:
        # too large to keep: forward without saving anything; metadata_backward re-runs it chunk by chunk
        last = tower.encoder.run_layers(x0, up.cu_seqlens, up.max_len, up.positions, groups=up.groups)
        return last, up, dict(ids=ids_flat, up=up, enc=None)
    last, enc_saved = encoder_forward(tower.encoder, x0, up.cu_seqlens, up.max_len, up.positions, groups=up.groups)
    return last, up, dict(ids=ids_flat, up=up, enc=enc_saved)


def metadata_backward(tower, saved, dlast, g: GradStore) -> None:
    enc, up = tower.encoder, saved["up"]
    pk = enc.packed()
    cfg = tower.config
    if saved["enc"] is not None:
        dx0 = encoder_backward(enc, saved["enc"], dlast, g)
        ops.embed_gather_ln_bwd(saved["ids"], up.src_index, None, pk["tok_emb"], None, pk["emb_norm"], dx0,
                                cfg.norm_eps, d_tok_emb=g(enc.embeddings.tok_embeddings.weight),
                                d_audio_embeds=None, dgamma=g(enc.embeddings.norm.weight))
        return
    # chunked recompute (the sequences are independent): forward + backward of `n_seq` sequences at a time; the
    # weight gradients of all chunks accumulate into the same buffers, so no gradient bucket may leave before the end
    per_tok = _encoder_saved_bytes_per_token(cfg)
    tok_budget = max(1, METADATA_SAVE_BUDGET // per_tok)
    import numpy as np
    cum = np.concatenate(([0], np.cumsum(np.asarray(up.lens_cpu, dtype=np.int64))))
    n_seq = len(up.lens_cpu)
    g.hold()
    try:
        s0 = 0
        while s0 < n_seq:
            # as many whole sequences as fit the token budget (at least one)
            s1 = int(np.searchsorted(cum, cum[s0] + tok_budget, side="right")) - 1
            s1 = min(max(s1, s0 + 1), n_seq)
            t0, t1 = int(cum[s0]), int(cum[s1])
            if t1 > t0:
                cu = (up.cu_seqlens[s0:s1 + 1] - t0).contiguous()
                src = up.src_index[t0:t1].contiguous()
                pos = up.positions[t0:t1].contiguous()
                groups = ops.attn_pack_groups(cu, t1 - t0) if up.groups is not None else None
                x0 = ops.embed_gather_ln(saved["ids"], src, None, pk["tok_emb"], None, pk["emb_norm"], cfg.norm_eps,
                                         rows=t1 - t0)
                max_len = int(np.max(cum[s0 + 1:s1 + 1] - cum[s0:s1]))
                _, enc_saved = encoder_forward(enc, x0, cu, max_len, pos, groups=groups)
                dx0 = encoder_backward(enc, enc_saved, dlast[t0:t1], g)
                ops.embed_gather_ln_bwd(saved["ids"], src, None, pk["tok_emb"], None, pk["emb_norm"], dx0, cfg.norm_eps,
                                        d_tok_emb=g(enc.embeddings.tok_embeddings.weight), d_audio_embeds=None,
                                        dgamma=g(enc.embeddings.norm.weight))
                del enc_saved, dx0, x0
            s0 = s1
    finally:
        g.release()


# ------------------------------------------------------------------------------------------------
# projection head + loss

def head_forward(last, cu_seqlens, mean_pool: bool, proj_w):
    pooled, proj, inv, emb32, emb16 = ops.pool_project_normalize(last, cu_seqlens, mean_pool, proj_w)
    return emb32, emb16, dict(pooled=pooled, proj=proj, inv=inv, cu=cu_seqlens, mean_pool=mean_pool, T=last.shape[0],
                              H=last.shape[1])


def head_backward(saved, demb32: torch.Tensor, proj_w: torch.Tensor, g_proj: torch.Tensor,
                  dlast: torch.Tensor | None = None) -> torch.Tensor:
    """demb32 [B,P] fp32 (gradient of the normalised embeddings) -> dlast [T,H] bf16 (added to `dlast` if given)."""
    dproj = ops.l2norm_bwd(saved["proj"], saved["inv"], demb32)
    _wgrad(dproj, saved["pooled"], g_proj)
    dpooled = _dgrad(dproj, proj_w)
    return pooled_backward(saved, dpooled, dlast)


def pooled_backward(saved, dpooled: torch.Tensor, dlast: torch.Tensor | None = None) -> torch.Tensor:
    accumulate = dlast is not None
    if dlast is None:
        dlast = torch.empty((saved["T"], saved["H"]), device=dpooled.device, dtype=BF16)
    ops.pool_bwd(dpooled, saved["cu"], saved["mean_pool"], dlast, accumulate=accumulate)
    return dlast


# ------------------------------------------------------------------------------------------------
# MLM prediction head + vocabulary cross-entropy (reference: CM3PPredictionHead :1229-1238, decoder,
# ForMaskedLMLoss via self.loss_function :994-996 / :1365-1367)

def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(F32).contiguous()


def mlm_head_forward(head, decoder, norm_eps: float, hidden: torch.Tensor, wcache: dict):
    """hidden [M,H] bf16 -> (logits view [M,V] of a [M, ld] buffer, saved)."""
    from .modeling_cm3p import _pack_linear
    wd = _pack_linear(head.dense, wcache, "hd")
    wv = _pack_linear(decoder, wcache, "dec")
    gamma = _f32(head.norm.weight)
    zd = ops.gemm(hidden, wd)
    yd = ops.gelu_fwd(zd)
    n = ops.layernorm(yd, gamma, norm_eps)
    M, V = hidden.shape[0], wv.shape[0]
    ld = (V + 7) // 8 * 8
    buf = torch.empty((M, ld), device=hidden.device, dtype=BF16)
    logits = buf[:, :V]
    if decoder.bias is not None:
        ops.gemm(n, wv, epilogue=ops.EPI_BIAS, aux=_f32(decoder.bias), out=logits)
    else:
        ops.gemm(n, wv, out=logits)
    return logits, dict(zd=zd, hidden=hidden, buf=buf, V=V, wd=wd, wv=wv, gamma=gamma, eps=norm_eps)


def mlm_head_backward(head, decoder, saved, g: GradStore) -> torch.Tensor:
    """saved["buf"] holds d(loss)/d(logits) (written in place by vocab_ce_bwd) -> d(hidden) [M,H]."""
    buf, V = saved["buf"], saved["V"]
    dl = buf[:, :V]
    yd = ops.gelu_fwd(saved["zd"])
    n = ops.layernorm(yd, saved["gamma"], saved["eps"])
    _wgrad(dl, n, g(decoder.weight))
    if decoder.bias is not None:
        tmp = torch.zeros((buf.shape[1],), device=buf.device, dtype=F32)
        ops.colsum_f32(buf, tmp)
        g(decoder.bias).add_(tmp[:V])
    dn = _dgrad(dl, saved["wv"])
    dyd = ops.layernorm_bwd(yd, dn, saved["gamma"], saved["eps"], dgamma=g(head.norm.weight))
    dzd = ops.gelu_bwd(saved["zd"], dyd)
    _wgrad(dzd, saved["hidden"], g(head.dense.weight))
    return _dgrad(dzd, saved["wd"])


def mlm_loss_forward(logits, vocab, labels_flat, src_index, num_items_in_batch):
    """-> (mean CE over labelled rows as a device scalar, saved for the backward)."""
    row_lse, loss_sum, count = ops.vocab_ce_fwd(logits, vocab, labels_flat, src_index)
    if num_items_in_batch is None:
        denom = count
    else:
        denom = torch.as_tensor(num_items_in_batch, device=logits.device, dtype=F32).reshape(1)
    return loss_sum / denom, dict(row_lse=row_lse, denom=denom, labels=labels_flat, src_index=src_index, vocab=vocab)


class _TrainStep(torch.autograd.Function):
    """loss = f(parameters): forward already ran (state in `st`), backward runs the explicit pass."""

    @staticmethod
    def forward(ctx, st, loss_value, *params):
        ctx.st = st
        ctx.n = len(params)
        return loss_value.clone()

    @staticmethod
    def backward(ctx, grad_out):
        st = ctx.st
        ctx.st = None
        grads = st.backward(grad_out)
        return (None, None, *grads)


class _StepState:
    """Everything the backward pass of one training step needs; `backward_fn(saved, grad_out, g)` is the
    model-specific schedule (contrastive / masked-LM / classification)."""

    def __init__(self, model, params, backward_fn):
        self.model = model
        self.params = params
        self.backward_fn = backward_fn
        self.saved = {}

    def backward(self, grad_out: torch.Tensor):
        model, sv = self.model, self.saved
        self.saved = None
        if sv is None:
            raise RuntimeError("cm3p_b200: backward called twice on the same training step")
        dp = getattr(model, "_dp", None)
        # global negatives: every rank differentiates the SAME global contrastive loss, so parameter gradients are
        # summed; every other objective (local negatives, MLM, classification) is a per-rank mean -> averaged
        sum_reduce = bool(dp is not None and dp.global_negatives and self.backward_fn is _contrastive_backward)
        key = (tuple(id(p) for p in self.params), self.backward_fn)
        cached = getattr(model, "_grad_order", None)
        order = cached[1] if (cached is not None and cached[0] == key) else None
        g = GradStore(self.params, order=order, dp=dp, sum_reduce=sum_reduce)
        gout = grad_out.detach().to(device=self.params[0].device, dtype=F32).reshape(1).contiguous()
        self.backward_fn(model, sv, gout, g)
        g.finish()
        if order is None:
            model._grad_order = (key, g.order())
        return [g(p).to(p.dtype) if p.requires_grad else None for p in self.params]


def _contrastive_backward(model, sv, gout, g: GradStore) -> None:
    dp = sv["dp"]
    world = dp.world_size if (dp is not None and dp.global_negatives) else 1
    dls = g(model.logit_scale).view(1)
    dS = ops.clip_loss_bwd(sv["S"], sv["true_idx"], sv["row_lse"], sv["col_lse"], sv["V"], gout, dls)
    if world > 1:
        # every rank evaluated the full loss, so d(logit_scale) is already complete on each rank (gradients are summed)
        dls.div_(world)
    dme = ops.gemm(dS, sv["be16"], trans_b=True, epilogue=ops.EPI_SCALE_F32, aux=sv["ls"])
    dbe = ops.gemm(dS, sv["me16"], trans_a=True, trans_b=True, epilogue=ops.EPI_SCALE_F32, aux=sv["ls"])
    del dS
    if dp is not None and dp.global_negatives:
        # every rank holds d(global loss)/d(all embeddings); its own rows are its row block
        dme = dp_utils.local_rows(dme, dp).contiguous()
        dbe = dp_utils.local_rows(dbe, dp).contiguous()
    # metadata tower first (small), then the beatmap tower
    dlast_m = head_backward(sv["mhead"], dme, sv["w_mp"], g(model.metadata_projection.weight))
    metadata_backward(model.metadata_model, sv["meta"], dlast_m, g)
    del dlast_m
    dlast_b = None
    if sv.get("mlm") is not None:
        # auxiliary masked-LM loss: loss += 0.5 * CE  (modeling_cm3p.py:994-996)
        ce = sv["mlm_ce"]
        # (a per-rank mean: pre-divided by the world size where the gradients of this step are summed over ranks)
        ops.vocab_ce_bwd(sv["mlm"]["buf"], ce["vocab"], ce["labels"], ce["src_index"], ce["row_lse"],
                         (gout * (0.5 / world) / ce["denom"]).contiguous())
        dlast_b = mlm_head_backward(model.head, model.decoder, sv["mlm"], g)
        sv["mlm"] = None
    dlast_b = head_backward(sv["bhead"], dbe, sv["w_bp"], g(model.beatmap_projection.weight), dlast=dlast_b)
    beatmap_backward(model.beatmap_model, sv["beat"], dlast_b, g)


def forward_with_grad(model, *, input_ids=None, input_features=None, metadata_ids=None, attention_mask=None,
                      metadata_attention_mask=None, metadata_variation_classes=None, labels=None,
                      return_loss=True, output_logits=False, **kwargs):
    """Training-mode `CM3PModel.forward` (reference: cm3p/modeling_cm3p.py:849-1012 under autograd)."""
    from .modeling_cm3p import (BaseModelOutputWithPooling, CM3PBeatmapModelOutput, CM3POutput, _out_dtype,
                                _pack_linear, _repad, _require_cuda)
    cfg = model.config
    if input_ids is None or metadata_ids is None or not return_loss:
        raise NotImplementedError(
            "cm3p_b200 training path: the contrastive step needs input_ids, metadata_ids and return_loss=True; "
            "call single-tower / no-loss forwards under torch.no_grad()")
    _require_cuda(input_ids, "input_ids")
    odt = _out_dtype(model)
    pad_outputs = getattr(cfg, "_attn_implementation", None) != "flash_attention_2"
    params = [p for p in model.parameters()]
    st = _StepState(model, params, _contrastive_backward)
    sv = st.saved

    last_b, up_b, audio_last, sv["beat"] = beatmap_forward_auto(model.beatmap_model, input_ids, input_features,
                                                                attention_mask)
    sv["w_bp"] = _pack_linear(model.beatmap_projection, model._wcache, "bp")
    be32, be16, sv["bhead"] = head_forward(last_b, up_b.cu_seqlens, not cfg.beatmap_config.cls_embed, sv["w_bp"])
    last_m, up_m, sv["meta"] = metadata_forward(model.metadata_model, metadata_ids, metadata_attention_mask)
    sv["w_mp"] = _pack_linear(model.metadata_projection, model._wcache, "mp")
    me32, me16, sv["mhead"] = head_forward(last_m, up_m.cu_seqlens, not cfg.metadata_config.cls_embed, sv["w_mp"])

    ls = model.logit_scale.detach().float().reshape(1)  # exp() is taken on the device, in the GEMM epilogues
    if metadata_ids.dim() == 3:
        Bm, V = metadata_ids.shape[:2]
        if metadata_variation_classes is None:
            raise ValueError("When providing multiple metadata variations, metadata_variation_classes must be "
                             "provided in order to compute loss correctly.")
        true_idx = (metadata_variation_classes == 0).int().argmax(dim=1).to(torch.int32).contiguous()
    else:
        Bm, V = metadata_ids.shape[0], 1
        true_idx = torch.zeros(Bm, device=be16.device, dtype=torch.int32)
    if Bm != be16.shape[0]:
        raise ValueError(f"metadata batch {Bm} != beatmap batch {be16.shape[0]}")
    dp = getattr(model, "_dp", None)
    sv["dp"] = dp
    if dp is not None and dp.global_negatives and dp.world_size > 1:
        # global negatives: all-gather the normalised embeddings (rank-major == concatenated batch order)
        be16, me16 = dp_utils.all_gather_rows(be16, dp), dp_utils.all_gather_rows(me16, dp)
        true_idx = dp_utils.all_gather_rows(true_idx, dp)
        Bm = Bm * dp.world_size
    S = ops.gemm(me16, be16, epilogue=ops.EPI_SCALE_F32, aux=ls)
    Bb = be16.shape[0]
    if metadata_ids.dim() == 3:
        logits_per_metadata = S.view(Bm, V, Bb)
        logits_per_beatmap = logits_per_metadata.permute(2, 0, 1)
    else:
        logits_per_metadata, logits_per_beatmap = S, S.t()
    loss_val, row_lse, col_lse = ops.clip_loss_fwd(S, true_idx, V)
    sv.update(S=S, true_idx=true_idx, row_lse=row_lse, col_lse=col_lse, V=V, ls=ls, be16=be16, me16=me16)

    logits_out = None
    if cfg.has_decoder_head and output_logits:
        # decoder(head(last_hidden)) on every real token (:987-993), only when logits are requested (the reference
        # gates the MLM term on output_logits, :987, :994); bf16 like the reference under autocast
        bc = cfg.beatmap_config
        logits_u, mlm_saved = mlm_head_forward(model.head, model.decoder, bc.norm_eps, last_b, model._wcache)
        if labels is not None:
            mlm, sv["mlm_ce"] = mlm_loss_forward(logits_u, mlm_saved["V"], labels.reshape(-1).contiguous(),
                                                 up_b.src_index, kwargs.get("num_items_in_batch"))
            sv["mlm"] = mlm_saved
            loss_val = loss_val + 0.5 * mlm
        if output_logits:
            logits_out = _repad(logits_u, up_b)  # a copy: the unpadded buffer is overwritten in the backward

    loss = _TrainStep.apply(st, loss_val.reshape(()), *params)

    lead = tuple(metadata_ids.shape[:-1])
    hidden_b = (_repad(last_b, up_b) if pad_outputs else last_b).to(odt)
    hidden_m = (_repad(last_m, up_m).view(*metadata_ids.shape, -1) if pad_outputs else last_m).to(odt)
    from .modeling_cm3p import CM3PAudioModelOutput
    audio_out = None if audio_last is None else CM3PAudioModelOutput(last_hidden_state=audio_last)
    return CM3POutput(
        loss=loss, logits_per_beatmap=logits_per_beatmap, logits_per_metadata=logits_per_metadata,
        metadata_embeds=me32.view(*lead, -1).to(odt), beatmap_embeds=be32.to(odt), logits=logits_out,
        metadata_model_output=BaseModelOutputWithPooling(
            last_hidden_state=hidden_m, pooler_output=sv["mhead"]["pooled"].view(*lead, -1).to(odt)),
        beatmap_model_output=CM3PBeatmapModelOutput(
            last_hidden_state=hidden_b, pooler_output=sv["bhead"]["pooled"].to(odt), audio_model_output=audio_out))


# ------------------------------------------------------------------------------------------------
# CM3PForMaskedLM / CM3PForBeatmapClassification (reference :1241-1379, :1137-1226)

def _mlm_backward(model, sv, gout, g: GradStore) -> None:
    ce = sv["mlm_ce"]
    ops.vocab_ce_bwd(sv["mlm"]["buf"], ce["vocab"], ce["labels"], ce["src_index"], ce["row_lse"],
                     (gout / ce["denom"]).contiguous())
    dh = mlm_head_backward(model.head, model.decoder, sv["mlm"], g)
    if sv["rows"] is not None:  # sparse prediction: only the labelled rows went through the head
        dlast = torch.zeros((sv["T"], dh.shape[1]), device=dh.device, dtype=BF16)
        ops.scatter_add_rows(dh, sv["rows"], dlast)
    else:
        dlast = dh
    beatmap_backward(model.beatmap_model, sv["beat"], dlast, g)


def masked_lm_forward(model, *, input_ids, input_features, attention_mask, labels, train: bool, **kwargs):
    """CM3PForMaskedLM.forward.  -> (loss | None, padded logits) ; with `train` the loss carries the backward."""
    from .modeling_cm3p import _repad, _require_cuda
    cfg = model.config
    _require_cuda(input_ids, "input_ids")
    params = list(model.parameters())
    st = _StepState(model, params, _mlm_backward)
    sv = st.saved
    if train:
        last, up, _, sv["beat"] = beatmap_forward(model.beatmap_model, input_ids, input_features, attention_mask)
    else:
        last, up, _ = model.beatmap_model.encode(input_ids, input_features, attention_mask)
    rows, hidden, src = None, last, up.src_index
    labels_flat = labels.reshape(-1).contiguous() if labels is not None else None
    if cfg.sparse_prediction and labels is not None:
        # only tokens whose label is not the ignore index go through the head (:1349-1357)
        lab_u = labels_flat.index_select(0, up.src_index.long())
        rows = torch.nonzero(lab_u != cfg.sparse_pred_ignore_index).reshape(-1).to(torch.int32)
        hidden = ops.gather_rows(last, rows)
        src = up.src_index.index_select(0, rows.long()).contiguous()
    logits_u, mlm_saved = mlm_head_forward(model.head, model.decoder, cfg.norm_eps, hidden, model._wcache)
    loss = None
    if labels is not None:
        mlm, sv["mlm_ce"] = mlm_loss_forward(logits_u, mlm_saved["V"], labels_flat, src,
                                             kwargs.get("num_items_in_batch"))
        loss = mlm.reshape(())
    if rows is not None:
        logits_out = logits_u.clone()  # the reference returns the (n_masked, vocab) logits in this mode
    else:
        logits_out = _repad(logits_u, up)
    if train and loss is not None:
        sv.update(mlm=mlm_saved, rows=rows, T=last.shape[0])
        loss = _TrainStep.apply(st, loss, *params)
    return loss, logits_out


def _cls_backward(model, sv, gout, g: GradStore) -> None:
    ce = sv["ce"]
    buf = sv["buf"]
    ops.vocab_ce_bwd(buf, ce["vocab"], ce["labels"], None, ce["row_lse"], (gout / ce["denom"]).contiguous())
    dl = buf[:, :ce["vocab"]]
    _wgrad(dl, sv["pooled"], g(model.classifier.weight))
    tmp = torch.zeros((buf.shape[1],), device=buf.device, dtype=F32)
    ops.colsum_f32(buf, tmp)
    g(model.classifier.bias).add_(tmp[:ce["vocab"]])
    dpooled = _dgrad(dl, sv["wc"])
    dlast = pooled_backward(sv["head"], dpooled)
    beatmap_backward(model.beatmap_model, sv["beat"], dlast, g)


def classification_forward(model, *, input_ids, input_features, attention_mask, labels, train: bool):
    """CM3PForBeatmapClassification.forward: logits = classifier(pooled); single-label CE when labels given."""
    from .modeling_cm3p import _pack_linear, _require_cuda
    cfg = model.config
    _require_cuda(input_ids, "input_ids")
    params = list(model.parameters())
    st = _StepState(model, params, _cls_backward)
    sv = st.saved
    if train:
        last, up, _, sv["beat"] = beatmap_forward(model.beatmap_model, input_ids, input_features, attention_mask)
    else:
        last, up, _ = model.beatmap_model.encode(input_ids, input_features, attention_mask)
    mean_pool = not cfg.cls_embed
    pooled = ops.pool_project_normalize(last, up.cu_seqlens, mean_pool, None)[0]
    if not isinstance(model.classifier, torch.nn.Linear):
        return None, pooled.float()
    wc = _pack_linear(model.classifier, model._wcache, "cls")
    C = wc.shape[0]
    ld = (C + 7) // 8 * 8
    buf = torch.zeros((pooled.shape[0], ld), device=pooled.device, dtype=BF16)
    logits = buf[:, :C]
    ops.gemm(pooled, wc, epilogue=ops.EPI_BIAS, aux=_f32(model.classifier.bias), out=logits)
    logits_out = logits.float()
    loss = None
    if labels is not None:
        problem = cfg.problem_type
        if problem is None:
            problem = "regression" if C == 1 else (
                "single_label_classification" if labels.dtype in (torch.long, torch.int) else
                "multi_label_classification")
            cfg.problem_type = problem
        if problem != "single_label_classification":
            raise NotImplementedError(f"cm3p_b200: problem_type {problem!r} has no CUDA loss kernel (only "
                                      "single_label_classification, the reference's ranked classifier)")
        lab = labels.reshape(-1).to(torch.int64).contiguous()
        row_lse, loss_sum, count = ops.vocab_ce_fwd(logits, C, lab, None)
        loss = (loss_sum / count).reshape(())
        if train:
            sv.update(ce=dict(row_lse=row_lse, denom=count, labels=lab, vocab=C), buf=buf, pooled=pooled, wc=wc,
                      head=dict(cu=up.cu_seqlens, mean_pool=mean_pool, T=last.shape[0], H=last.shape[1]))
            loss = _TrainStep.apply(st, loss, *params)
    return loss, logits_out
