"""Thin torch-tensor wrappers over the C ABI (include/cm3p_b200.h).

PyTorch is used here only for device memory and the current CUDA stream; every function forwards
raw pointers to libcm3p_b200.so and raises on a non-zero status.  No function in this module has a
PyTorch implementation to fall back to.
"""
from __future__ import annotations

import torch

from . import _lib

EPI_STORE, EPI_RESIDUAL, EPI_GELU, EPI_BIAS_GELU, EPI_BIAS = 0, 1, 2, 3, 4
EPI_GEGLU, EPI_GEGLU_SAVE, EPI_ROPE, EPI_SCALE_F32 = 5, 6, 7, 8

# Fold the pre-norm LayerNorms of the encoder blocks into the neighbouring GEMMs (cm3p_gemm_bf16_ln): the residual
# GEMM's epilogue leaves per-tile (sum, sum^2) partials of the rows it writes (summed in tile order by the consumer:
# deterministic and batch-invariant), the consumer GEMM multiplies by W . diag(gamma) and applies
# rstd * (acc - mean * colsum) in its epilogue.  On by default since the epilogues stopped being the critical path
# (pair-MMA GEMM, round 2): -2.2 % inference step time, -0.5 .. -1.4 % train step, all parity tests unchanged.
# CM3P_FUSE_LN=0 keeps the separate LayerNorm kernels.
import os as _os

FUSE_LAYERNORM = _os.environ.get("CM3P_FUSE_LN", "1") == "1"
# training: keep LayerNorm outputs for the backward ("1"), recompute them ("0"), or decide by free memory ("auto")
SAVE_LAYERNORM = _os.environ.get("CM3P_SAVE_LN", "auto")

# Muon: orthogonalise all same-shape weight matrices with grouped GEMM launches (off = one matrix at a time)
GROUPED_MUON = True

# kernel launches issued through this module (bench.py reports it as `gpu_launches`)
LAUNCH_COUNT = 0
_LAUNCHES_PER_CALL = {
    "gemm": 1, "attn": 1, "layernorm": 1, "embed": 1, "pool_project": 3, "pool": 1, "clip_loss": 2,
    "attn_bwd": 2, "rowwise": 1,
}


def _count(kind: str) -> None:
    global LAUNCH_COUNT
    LAUNCH_COUNT += _LAUNCHES_PER_CALL[kind]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# run-time knobs of the library (include/cm3p_b200.h CM3P_OPT_*)
OPT_FWD_BLOCKS_PER_CTA, OPT_BWD_OUTER_PER_CTA, OPT_GEMM_CLUSTER = 0, 1, 2
OPT_ATTN_FORCE_TILE_KERNELS, OPT_WGRAD_DETERMINISTIC, OPT_TMAP_CACHE, OPT_ATTN_WINDOW_WALK = 3, 4, 5, 6


def set_option(option: int, value: int) -> None:
    _lib.check(_lib.load().cm3p_set_option(int(option), int(value)), "cm3p_set_option")


def get_option(option: int) -> int:
    return int(_lib.load().cm3p_get_option(int(option)))


# zeroed int32 counters the split-K weight-gradient GEMMs use as per-tile turnstiles (ordered, bit-reproducible
# accumulation); one buffer per (device, stream): launches on one stream are serialised, and the kernel leaves the
# counters zeroed
_SEM_COUNT = 16384
_SEM_CACHE: dict = {}


def _tile_sem(device) -> torch.Tensor:
    key = (device.index, _stream())
    t = _SEM_CACHE.get(key)
    if t is None:
        t = torch.zeros(_SEM_COUNT, device=device, dtype=torch.int32)
        _SEM_CACHE[key] = t
    return t


def _ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (cm3p_b200 has no CPU path)")
    if t.device.index != torch.cuda.current_device():
        # launches go to the CURRENT device's stream: run under `torch.cuda.device(t.device)` for other GPUs
        raise RuntimeError(f"{name} lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")


def gemm(a: torch.Tensor, b: torch.Tensor, *, epilogue: int = EPI_STORE, out: torch.Tensor | None = None,
         aux: torch.Tensor | None = None, c2: torch.Tensor | None = None, scale: float = 1.0,
         accumulate: bool = False, positions: torch.Tensor | None = None, rope_table: torch.Tensor | None = None,
         rope_cols: int = 0, trans_a: bool = False, trans_b: bool = False, stats_out: torch.Tensor | None = None,
         row_stats: torch.Tensor | None = None, col_corr: torch.Tensor | None = None,
         ln_eps: float = 1e-5, groups: int = 1) -> torch.Tensor:
    """out[M,N] = epilogue(A @ B^T).  A: [M,K] (or [K,M] if trans_a); B: [N,K] (or [K,N] if trans_b).

    groups > 1: `groups` independent problems in one launch, operands stacked along their outer dimension
    (A [groups*m, K] or transposed [groups*K, m]; B [groups*N, K] or [groups*K, N]; out / aux [groups*m, N]).

    LayerNorm folding (cm3p_gemm_bf16_ln): `stats_out` [ceil(N/256),M,2] fp32 (per-tile row sum / sum of squares of the rows an
    EPI_RESIDUAL GEMM writes); `row_stats` + `col_corr` on an EPI_ROPE / EPI_GEGLU(_SAVE) GEMM whose B is
    W.diag(gamma) apply the normalisation of the A rows in the epilogue."""
    _req(a, torch.bfloat16, "a")
    _req(b, torch.bfloat16, "b")
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if trans_a else a.shape
    N, Kb = (b.shape[1], b.shape[0]) if trans_b else b.shape
    group_m = 0
    if groups > 1:
        if trans_a:
            group_m, M, K = M, M * groups, K // groups
        else:
            group_m = M // groups
        if trans_b:
            Kb //= groups
        else:
            N //= groups
    if K != Kb:
        raise ValueError(f"gemm: inner dimensions differ ({K} vs {Kb})")
    n_out = N // 2 if epilogue in (EPI_GEGLU, EPI_GEGLU_SAVE) else N
    out_dtype = torch.float32 if epilogue == EPI_SCALE_F32 else torch.bfloat16
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=out_dtype)
    else:
        _req(out, out_dtype, "out")
        assert out.shape == (M, n_out) and out.stride(1) == 1
    if epilogue == EPI_GEGLU_SAVE and c2 is None:
        c2 = torch.empty((M, N), device=a.device, dtype=torch.bfloat16)
    lib = _lib.load()
    if stats_out is not None or row_stats is not None:
        assert not trans_a and not trans_b and not accumulate
        rc = lib.cm3p_gemm_bf16_ln(
            a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(), out.stride(0), M, N, K, epilogue,
            _ptr(aux), (aux.stride(0) if aux is not None and aux.dim() == 2 else 0), _ptr(c2),
            (c2.stride(0) if c2 is not None else 0), _ptr(positions), _ptr(rope_table), rope_cols, _ptr(stats_out),
            _ptr(row_stats), _ptr(col_corr), float(ln_eps), _stream())
        _lib.check(rc, "cm3p_gemm_bf16_ln")
        _count("gemm")
        return out
    sem = _tile_sem(a.device) if (accumulate and epilogue == EPI_SCALE_F32) else None
    rc = lib.cm3p_gemm_bf16(
        a.data_ptr(), a.stride(0), int(trans_a), b.data_ptr(), b.stride(0), int(trans_b), out.data_ptr(),
        out.stride(0), M, N, K, epilogue, _ptr(aux), (aux.stride(0) if aux is not None and aux.dim() == 2 else 0),
        _ptr(c2), (c2.stride(0) if c2 is not None else 0), float(scale), int(accumulate), _ptr(positions),
        _ptr(rope_table), rope_cols, _ptr(sem), (_SEM_COUNT if sem is not None else 0), group_m, _stream())
    _lib.check(rc, "cm3p_gemm_bf16")
    _count("gemm")
    return out


class PackedGroups:
    """Group table of the packed short-sequence attention kernels (cm3p_attn_pack_groups)."""

    def __init__(self, table: torch.Tensor, count: torch.Tensor, max_groups: int):
        self.table, self.count, self.max_groups = table, count, max_groups


def attn_pack_groups(cu_seqlens: torch.Tensor, total_tokens: int) -> PackedGroups:
    """Packs consecutive sequences (each <= 128 tokens) into groups of <= 128 tokens; device-side, no host sync."""
    _req(cu_seqlens, torch.int32, "cu_seqlens")
    batch = cu_seqlens.numel() - 1
    max_groups = max(1, min(batch, 2 * (int(total_tokens) // 128) + (batch + 63) // 64 + 1))
    table = torch.empty((max_groups, 2), device=cu_seqlens.device, dtype=torch.int32)
    count = torch.empty((1,), device=cu_seqlens.device, dtype=torch.int32)
    rc = _lib.load().cm3p_attn_pack_groups(cu_seqlens.data_ptr(), batch, table.data_ptr(), count.data_ptr(), max_groups,
                                           _stream())
    _lib.check(rc, "cm3p_attn_pack_groups")
    _count("rowwise")
    return PackedGroups(table, count, max_groups)


def attn_varlen_fwd(qkv: torch.Tensor, cu_seqlens: torch.Tensor, max_seqlen: int, heads: int, window: int = -1,
                    out: torch.Tensor | None = None, lse: torch.Tensor | None = None,
                    groups: PackedGroups | None = None) -> torch.Tensor:
    """qkv [T, 3*heads*64] bf16 (rotated) -> out [T, heads*64] bf16.  window < 0 = global.
    groups: packed short sequences (attn_pack_groups; every sequence <= 128 tokens)."""
    _req(qkv, torch.bfloat16, "qkv")
    _req(cu_seqlens, torch.int32, "cu_seqlens")
    T = qkv.shape[0]
    H = heads * 64
    assert qkv.shape[1] == 3 * H and qkv.is_contiguous()
    if out is None:
        out = torch.empty((T, H), device=qkv.device, dtype=torch.bfloat16)
    g = groups if (groups is not None and window < 0) else None
    rc = _lib.load().cm3p_attn_varlen_fwd(qkv.data_ptr(), out.data_ptr(), _ptr(lse), cu_seqlens.data_ptr(), T,
                                          cu_seqlens.numel() - 1, heads, 64, int(max_seqlen), int(window),
                                          None if g is None else g.table.data_ptr(),
                                          None if g is None else g.count.data_ptr(),
                                          0 if g is None else g.max_groups, _stream())
    _lib.check(rc, "cm3p_attn_varlen_fwd")
    _count("attn")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, eps: float, out: torch.Tensor | None = None,
              stats: torch.Tensor | None = None) -> torch.Tensor:
    _req(x, torch.bfloat16, "x")
    _req(gamma, torch.float32, "gamma")
    assert x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    rows = x.numel() // x.shape[-1]
    rc = _lib.load().cm3p_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), out.data_ptr(), _ptr(stats), rows,
                                        x.shape[-1], float(eps), _stream())
    _lib.check(rc, "cm3p_layernorm_fwd")
    _count("layernorm")
    return out


def embed_gather_ln(ids: torch.Tensor, src_index: torch.Tensor | None, audio_slot: torch.Tensor | None,
                    tok_emb: torch.Tensor, audio_embeds: torch.Tensor | None, gamma: torch.Tensor, eps: float,
                    rows: int, out: torch.Tensor | None = None, stats: torch.Tensor | None = None) -> torch.Tensor:
    _req(ids, torch.int64, "ids")
    _req(tok_emb, torch.bfloat16, "tok_emb")
    _req(gamma, torch.float32, "gamma")
    H = tok_emb.shape[1]
    if out is None:
        out = torch.empty((rows, H), device=tok_emb.device, dtype=torch.bfloat16)
    rc = _lib.load().cm3p_embed_gather_ln(ids.data_ptr(), _ptr(src_index), _ptr(audio_slot), tok_emb.data_ptr(),
                                          _ptr(audio_embeds), gamma.data_ptr(), out.data_ptr(), _ptr(stats), rows, H,
                                          tok_emb.shape[0], float(eps), _stream())
    _lib.check(rc, "cm3p_embed_gather_ln")
    _count("embed")
    return out


def transpose_cast(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """fp32 [B, C, F] (the processor's log-mel layout) -> bf16 [B, F, C] channels-last."""
    _req(x, torch.float32, "x")
    assert x.dim() == 3 and x.is_contiguous()
    B, C, F = x.shape
    if out is None:
        out = torch.empty((B, F, C), device=x.device, dtype=torch.bfloat16)
    _lib.check(_lib.load().cm3p_transpose_cast_bf16(x.data_ptr(), out.data_ptr(), B, C, F, _stream()),
               "cm3p_transpose_cast_bf16")
    _count("rowwise")
    return out


def pack_conv_weight(weight: torch.Tensor) -> torch.Tensor:
    """conv.weight [C_out, C_in, 3] -> bf16 [C_out, 3 * c_pad], column tap * c_pad + c, zero for c >= C_in
    (c_pad = C_in rounded up to 64: the K blocks of the implicit GEMM never straddle a tap)."""
    c_out, c_in, k = weight.shape
    assert k == 3
    c_pad = (c_in + 63) // 64 * 64
    w = torch.zeros((c_out, 3, c_pad), device=weight.device, dtype=torch.bfloat16)
    w[:, :, :c_in] = weight.detach().permute(0, 2, 1).to(torch.bfloat16)
    return w.reshape(c_out, 3 * c_pad).contiguous()


def unpack_conv_weight_grad(dw: torch.Tensor, c_in: int) -> torch.Tensor:
    """[C_out, 3 * c_pad] (tap, c) -> view shaped like conv.weight [C_out, C_in, 3]."""
    c_out = dw.shape[0]
    return dw.view(c_out, 3, -1)[:, :, :c_in].permute(0, 2, 1)


def conv1d_k3(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, stride: int, gelu: bool,
              out: torch.Tensor | None = None) -> torch.Tensor:
    """Implicit-GEMM conv1d (kernel 3, padding 1): x bf16 [B, F, C_in] channels-last, weight from pack_conv_weight
    -> bf16 [B, F / stride, C_out] (= gelu(conv + bias) when `gelu`, else the pre-activation)."""
    _req(x, torch.bfloat16, "x")
    _req(weight, torch.bfloat16, "weight")
    _req(bias, torch.float32, "bias")
    B, F, C = x.shape
    assert x.is_contiguous() and weight.is_contiguous() and weight.shape[1] % 3 == 0
    c_out, c_pad = weight.shape[0], weight.shape[1] // 3
    if out is None:
        out = torch.empty((B, F // stride, c_out), device=x.device, dtype=torch.bfloat16)
    rc = _lib.load().cm3p_conv1d_k3_fwd(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), out.data_ptr(), B, C, c_pad, F,
                                        c_out, stride, int(gelu), _stream())
    _lib.check(rc, "cm3p_conv1d_k3_fwd")
    _count("gemm")
    return out


def conv1d_k3_wgrad(dz: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, stride: int) -> torch.Tensor:
    """dw [C_out, 3 * c_pad] fp32 += conv weight gradient; dz bf16 [B, F / stride, C_out], x bf16 [B, F, C_in]."""
    _req(dz, torch.bfloat16, "dz")
    _req(x, torch.bfloat16, "x")
    _req(dw, torch.float32, "dw")
    B, F, C = x.shape
    c_out = dz.shape[-1]
    assert dz.is_contiguous() and x.is_contiguous() and dw.is_contiguous() and dw.shape[0] == c_out
    assert dz.numel() == B * (F // stride) * c_out and dw.shape[1] % 3 == 0
    sem = _tile_sem(x.device)
    rc = _lib.load().cm3p_conv1d_k3_wgrad(dz.data_ptr(), x.data_ptr(), dw.data_ptr(), B, C, dw.shape[1] // 3, F, c_out,
                                          stride, sem.data_ptr(), _SEM_COUNT, _stream())
    _lib.check(rc, "cm3p_conv1d_k3_wgrad")
    _count("gemm")
    return dw


def pool_project_normalize(hidden: torch.Tensor, cu_seqlens: torch.Tensor, mean_pool: bool,
                           proj_w: torch.Tensor | None, want_bf16: bool = True):
    """-> (pooled bf16 [B,H], proj fp32 [B,P] | None, inv_norm [B] | None, embeds fp32 | None, embeds bf16 | None)."""
    _req(hidden, torch.bfloat16, "hidden")
    _req(cu_seqlens, torch.int32, "cu_seqlens")
    B = cu_seqlens.numel() - 1
    H = hidden.shape[1]
    dev = hidden.device
    pooled = torch.empty((B, H), device=dev, dtype=torch.bfloat16)
    proj = inv = emb = emb16 = None
    P = 0
    if proj_w is not None:
        _req(proj_w, torch.bfloat16, "proj_w")
        P = proj_w.shape[0]
        proj = torch.empty((B, P), device=dev, dtype=torch.float32)
        inv = torch.empty((B,), device=dev, dtype=torch.float32)
        emb = torch.empty((B, P), device=dev, dtype=torch.float32)
        emb16 = torch.empty((B, P), device=dev, dtype=torch.bfloat16) if want_bf16 else None
    rc = _lib.load().cm3p_pool_project_normalize(hidden.data_ptr(), cu_seqlens.data_ptr(), int(mean_pool),
                                                 _ptr(proj_w), pooled.data_ptr(), _ptr(proj), _ptr(inv), _ptr(emb),
                                                 _ptr(emb16), B, H, P, _stream())
    _lib.check(rc, "cm3p_pool_project_normalize")
    _count("pool_project" if proj_w is not None else "pool")
    return pooled, proj, inv, emb, emb16


def clip_loss_fwd(S: torch.Tensor, true_idx: torch.Tensor, V: int):
    """S fp32 [Bm*V, Bb] -> (loss [1] fp32, row_lse [Bm], col_lse [Bb])."""
    _req(S, torch.float32, "S")
    _req(true_idx, torch.int32, "true_idx")
    assert S.is_contiguous()
    Bb = S.shape[1]
    Bm = S.shape[0] // V
    row_lse = torch.empty((Bm,), device=S.device, dtype=torch.float32)
    col_lse = torch.empty((Bb,), device=S.device, dtype=torch.float32)
    loss = torch.empty((1,), device=S.device, dtype=torch.float32)
    rc = _lib.load().cm3p_clip_loss_fwd(S.data_ptr(), true_idx.data_ptr(), row_lse.data_ptr(), col_lse.data_ptr(),
                                        loss.data_ptr(), Bm, V, Bb, _stream())
    _lib.check(rc, "cm3p_clip_loss_fwd")
    _count("clip_loss")
    return loss, row_lse, col_lse


# ------------------------------------------------------------------------------------------------
# backward entry points (include/cm3p_b200.h, "Backward entry points")

def attn_varlen_bwd(qkv: torch.Tensor, out: torch.Tensor, dout: torch.Tensor, lse: torch.Tensor,
                    cu_seqlens: torch.Tensor, max_seqlen: int, heads: int, window: int = -1,
                    positions: torch.Tensor | None = None, rope_table: torch.Tensor | None = None,
                    dqkv: torch.Tensor | None = None, delta: torch.Tensor | None = None,
                    groups: PackedGroups | None = None) -> torch.Tensor:
    """-> dqkv [T, 3*heads*64] bf16, gradient w.r.t. the un-rotated Wqkv output when positions/rope_table
    are given (otherwise w.r.t. the qkv passed in)."""
    for t, n in ((qkv, "qkv"), (out, "out"), (dout, "dout")):
        _req(t, torch.bfloat16, n)
    _req(lse, torch.float32, "lse")
    _req(cu_seqlens, torch.int32, "cu_seqlens")
    T, H = out.shape
    assert qkv.shape == (T, 3 * H) and dout.shape == (T, H) and lse.shape == (heads, T)
    assert qkv.is_contiguous() and out.is_contiguous() and dout.is_contiguous() and lse.is_contiguous()
    if dqkv is None:
        dqkv = torch.empty_like(qkv)
    if delta is None:
        delta = torch.empty_like(lse)
    g = groups if (groups is not None and window < 0) else None
    rc = _lib.load().cm3p_attn_varlen_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(),
                                          delta.data_ptr(), dqkv.data_ptr(), cu_seqlens.data_ptr(), _ptr(positions),
                                          _ptr(rope_table), T, cu_seqlens.numel() - 1, heads, 64, int(max_seqlen),
                                          int(window), None if g is None else g.table.data_ptr(),
                                          None if g is None else g.count.data_ptr(),
                                          0 if g is None else g.max_groups, _stream())
    _lib.check(rc, "cm3p_attn_varlen_bwd")
    _count("attn_bwd" if g is None else "attn")
    return dqkv


def layernorm_bwd(x: torch.Tensor, dy: torch.Tensor, gamma: torch.Tensor, eps: float,
                  dres: torch.Tensor | None = None, dx: torch.Tensor | None = None,
                  dgamma: torch.Tensor | None = None) -> torch.Tensor:
    """dx = LN'(x).dy (+ dres); dgamma (fp32 [H]) is accumulated into."""
    _req(x, torch.bfloat16, "x")
    _req(dy, torch.bfloat16, "dy")
    _req(gamma, torch.float32, "gamma")
    assert x.is_contiguous() and dy.is_contiguous() and x.shape == dy.shape
    if dx is None:
        dx = torch.empty_like(x)
    rows = x.numel() // x.shape[-1]
    rc = _lib.load().cm3p_layernorm_bwd(x.data_ptr(), dy.data_ptr(), gamma.data_ptr(), _ptr(dres), dx.data_ptr(),
                                        _ptr(dgamma), rows, x.shape[-1], float(eps), _stream())
    _lib.check(rc, "cm3p_layernorm_bwd")
    _count("rowwise")
    return dx


def embed_gather_ln_bwd(ids, src_index, audio_slot, tok_emb, audio_embeds, gamma, dy, eps, d_tok_emb=None,
                        d_audio_embeds=None, dgamma=None) -> None:
    _req(dy, torch.bfloat16, "dy")
    rows, H = dy.shape
    rc = _lib.load().cm3p_embed_gather_ln_bwd(ids.data_ptr(), _ptr(src_index), _ptr(audio_slot), tok_emb.data_ptr(),
                                              _ptr(audio_embeds), gamma.data_ptr(), dy.data_ptr(), _ptr(d_tok_emb),
                                              _ptr(d_audio_embeds), _ptr(dgamma), rows, H, tok_emb.shape[0], float(eps),
                                              _stream())
    _lib.check(rc, "cm3p_embed_gather_ln_bwd")
    _count("rowwise")


def geglu_bwd(ug: torch.Tensor, dh: torch.Tensor, dug: torch.Tensor | None = None, h: torch.Tensor | None = None,
              want_h: bool = True):
    """ug [T,2I] interleaved pre-activation, dh [T,I] -> (dug [T,2I], h [T,I] = gelu(u)*g recomputed)."""
    _req(ug, torch.bfloat16, "ug")
    _req(dh, torch.bfloat16, "dh")
    rows, I = dh.shape
    assert ug.shape == (rows, 2 * I) and ug.is_contiguous() and dh.is_contiguous()
    if dug is None:
        dug = torch.empty_like(ug)
    if h is None and want_h:
        h = torch.empty_like(dh)
    rc = _lib.load().cm3p_geglu_bwd(ug.data_ptr(), dh.data_ptr(), dug.data_ptr(), _ptr(h), rows, I, _stream())
    _lib.check(rc, "cm3p_geglu_bwd")
    _count("rowwise")
    return dug, h


def gelu_fwd(z: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    _req(z, torch.bfloat16, "z")
    assert z.is_contiguous()
    if out is None:
        out = torch.empty_like(z)
    _lib.check(_lib.load().cm3p_gelu_fwd(z.data_ptr(), out.data_ptr(), z.numel(), _stream()), "cm3p_gelu_fwd")
    _count("rowwise")
    return out


def gelu_bwd(z: torch.Tensor, dy: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    _req(z, torch.bfloat16, "z")
    _req(dy, torch.bfloat16, "dy")
    assert z.is_contiguous() and dy.is_contiguous() and z.numel() == dy.numel()
    if out is None:
        out = torch.empty_like(z)
    _lib.check(_lib.load().cm3p_gelu_bwd(z.data_ptr(), dy.data_ptr(), out.data_ptr(), z.numel(), _stream()),
               "cm3p_gelu_bwd")
    _count("rowwise")
    return out


def colsum_f32(dy: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out [N] fp32 += column sums of dy [rows, N] bf16."""
    _req(dy, torch.bfloat16, "dy")
    _req(out, torch.float32, "out")
    assert dy.dim() == 2 and dy.is_contiguous() and out.numel() == dy.shape[1]
    _lib.check(_lib.load().cm3p_colsum_f32(dy.data_ptr(), out.data_ptr(), dy.shape[0], dy.shape[1], _stream()),
               "cm3p_colsum_f32")
    _count("rowwise")
    return out


def pool_bwd(dpooled: torch.Tensor, cu_seqlens: torch.Tensor, mean_pool: bool, dhidden: torch.Tensor,
             accumulate: bool) -> torch.Tensor:
    _req(dpooled, torch.bfloat16, "dpooled")
    _req(dhidden, torch.bfloat16, "dhidden")
    B, H = dpooled.shape
    assert dpooled.is_contiguous() and dhidden.is_contiguous() and dhidden.shape[1] == H
    rc = _lib.load().cm3p_pool_bwd(dpooled.data_ptr(), cu_seqlens.data_ptr(), dhidden.data_ptr(), int(mean_pool),
                                   int(accumulate), B, H, _stream())
    _lib.check(rc, "cm3p_pool_bwd")
    _count("rowwise")
    return dhidden


def l2norm_bwd(proj: torch.Tensor, inv_norm: torch.Tensor, dembeds: torch.Tensor) -> torch.Tensor:
    """-> dproj bf16 [B,P]."""
    _req(proj, torch.float32, "proj")
    _req(dembeds, torch.float32, "dembeds")
    B, P = proj.shape
    assert proj.is_contiguous() and dembeds.is_contiguous() and dembeds.shape == proj.shape
    out = torch.empty((B, P), device=proj.device, dtype=torch.bfloat16)
    rc = _lib.load().cm3p_l2norm_bwd(proj.data_ptr(), inv_norm.data_ptr(), dembeds.data_ptr(), out.data_ptr(), B, P,
                                     _stream())
    _lib.check(rc, "cm3p_l2norm_bwd")
    _count("rowwise")
    return out


def clip_loss_bwd(S: torch.Tensor, true_idx: torch.Tensor, row_lse: torch.Tensor, col_lse: torch.Tensor, V: int,
                  grad_out: torch.Tensor | None, dlogit_scale: torch.Tensor):
    """-> dS bf16 [Bm*V, ld] (ld = Bb rounded up to 8; columns >= Bb are padding), dlogit_scale += sum dS*S."""
    _req(S, torch.float32, "S")
    R, Bb = S.shape
    ld = (Bb + 7) // 8 * 8
    dS = torch.empty((R, ld), device=S.device, dtype=torch.bfloat16)
    rc = _lib.load().cm3p_clip_loss_bwd(S.data_ptr(), true_idx.data_ptr(), row_lse.data_ptr(), col_lse.data_ptr(),
                                        _ptr(grad_out), dS.data_ptr(), ld, dlogit_scale.data_ptr(), R // V, V, Bb,
                                        _stream())
    _lib.check(rc, "cm3p_clip_loss_bwd")
    _count("rowwise")
    return dS[:, :Bb]


def conv2_col2im_gelu_bwd(da2: torch.Tensor, z1: torch.Tensor) -> torch.Tensor:
    """da2 [B*F/2, 3C] bf16, z1 [B,F,C] bf16 (conv1 pre-activation) -> dz1 [B,F,C]."""
    _req(da2, torch.bfloat16, "da2")
    _req(z1, torch.bfloat16, "z1")
    B, F, C = z1.shape
    assert da2.is_contiguous() and z1.is_contiguous() and da2.shape == (B * (F // 2), 3 * C)
    out = torch.empty_like(z1)
    rc = _lib.load().cm3p_conv2_col2im_gelu_bwd(da2.data_ptr(), z1.data_ptr(), out.data_ptr(), B, F, C, _stream())
    _lib.check(rc, "cm3p_conv2_col2im_gelu_bwd")
    _count("rowwise")
    return out


def vocab_ce_fwd(logits: torch.Tensor, vocab: int, labels: torch.Tensor, src_index: torch.Tensor | None,
                 ignore_index: int = -100):
    """logits bf16 [rows, ld>=vocab]; labels int64 (padded, flat).  -> (row_lse [rows], loss_sum [1], count [1])."""
    _req(logits, torch.bfloat16, "logits")
    _req(labels, torch.int64, "labels")
    assert logits.dim() == 2 and logits.stride(1) == 1
    rows = logits.shape[0]
    row_lse = torch.empty((rows,), device=logits.device, dtype=torch.float32)
    acc = torch.zeros((2,), device=logits.device, dtype=torch.float32)
    rc = _lib.load().cm3p_vocab_ce_fwd(logits.data_ptr(), logits.stride(0), labels.data_ptr(), _ptr(src_index),
                                       int(ignore_index), row_lse.data_ptr(), acc[0:1].data_ptr(), acc[1:2].data_ptr(),
                                       rows, int(vocab), _stream())
    _lib.check(rc, "cm3p_vocab_ce_fwd")
    _count("rowwise")
    return row_lse, acc[0:1], acc[1:2]


def vocab_ce_bwd(logits: torch.Tensor, vocab: int, labels: torch.Tensor, src_index: torch.Tensor | None,
                 row_lse: torch.Tensor, scale: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    """Overwrites logits [rows, ld] with d(loss)/d(logits) * scale (device scalar) and returns it."""
    _req(logits, torch.bfloat16, "logits")
    _req(scale, torch.float32, "scale")
    rc = _lib.load().cm3p_vocab_ce_bwd(logits.data_ptr(), logits.stride(0), labels.data_ptr(), _ptr(src_index),
                                       int(ignore_index), row_lse.data_ptr(), scale.data_ptr(), logits.shape[0],
                                       int(vocab), _stream())
    _lib.check(rc, "cm3p_vocab_ce_bwd")
    _count("rowwise")
    return logits


def gather_rows(x: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    _req(x, torch.bfloat16, "x")
    _req(index, torch.int32, "index")
    assert x.is_contiguous()
    out = torch.empty((index.numel(), x.shape[1]), device=x.device, dtype=torch.bfloat16)
    rc = _lib.load().cm3p_gather_rows(x.data_ptr(), index.data_ptr(), out.data_ptr(), index.numel(), x.shape[1],
                                      _stream())
    _lib.check(rc, "cm3p_gather_rows")
    _count("rowwise")
    return out


def scatter_add_rows(dx_rows: torch.Tensor, index: torch.Tensor, dx: torch.Tensor) -> torch.Tensor:
    _req(dx_rows, torch.bfloat16, "dx_rows")
    _req(dx, torch.bfloat16, "dx")
    assert dx_rows.is_contiguous() and dx.is_contiguous() and dx_rows.shape[1] == dx.shape[1]
    rc = _lib.load().cm3p_scatter_add_rows(dx_rows.data_ptr(), index.data_ptr(), dx.data_ptr(), index.numel(),
                                           dx.shape[1], _stream())
    _lib.check(rc, "cm3p_scatter_add_rows")
    _count("rowwise")
    return dx


# ------------------------------------------------------------------------------------------------
# host-side preparation helpers (layout only, no arithmetic on the hot path)

_ROPE_CACHE: dict = {}


def rope_table(theta: float, max_pos: int, device) -> torch.Tensor:
    """[max_pos, 32, 2] fp32 (cos, sin) with the reference's fp32 recipe (MB:94-172): inv_freq =
    theta^(-2k/64), angle = float(pos) * inv_freq."""
    key = (float(theta), int(max_pos), str(device))
    tab = _ROPE_CACHE.get(key)
    if tab is None:
        inv_freq = 1.0 / (theta ** (torch.arange(0, 64, 2, dtype=torch.int64).float() / 64))
        ang = torch.arange(max_pos, dtype=torch.float32)[:, None] * inv_freq[None, :]
        tab = torch.stack((ang.cos(), ang.sin()), dim=-1).contiguous().to(device)
        _ROPE_CACHE[key] = tab
    return tab


def interleave_wi(w: torch.Tensor) -> torch.Tensor:
    """Wi [2I, H] (rows: I 'input' then I 'gate', MB:90) -> rows interleaved in groups of 16 so one
    32-column accumulator chunk of the GEMM epilogue holds 16 inputs and their 16 gates."""
    two_i, H = w.shape
    I = two_i // 2
    assert I % 16 == 0, "intermediate_size must be a multiple of 16"
    return torch.stack((w[:I].reshape(I // 16, 16, H), w[I:].reshape(I // 16, 16, H)), dim=1).reshape(two_i, H)


def deinterleave_wi(w: torch.Tensor) -> torch.Tensor:
    two_i, H = w.shape
    I = two_i // 2
    v = w.reshape(I // 16, 2, 16, H)
    return torch.cat((v[:, 0].reshape(I, H), v[:, 1].reshape(I, H)), dim=0)
