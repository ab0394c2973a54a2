"""Backward kernels on a real B200 (through the C ABI) against torch autograd of the same op in fp32
on the same bf16-rounded inputs."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from cm3p_b200 import ops
    return ops


def _rand(shape, scale=1.0, seed=0, dtype=torch.bfloat16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).to(DEV)


def _close(name, got, want, atol, rtol):
    got, want = got.float(), want.float()
    err = (got - want).abs()
    bad = err > atol + rtol * want.abs()
    if bad.any():
        idx = bad.nonzero()[:6].tolist()
        lines = [f"{name}: {int(bad.sum())}/{bad.numel()} mismatches, max abs err {float(err.max()):.4g}, "
                 f"ref max {float(want.abs().max()):.4g}"]
        for i in idx:
            lines.append(f"  at {i}: got {float(got[tuple(i)]):.5g} want {float(want[tuple(i)]):.5g}")
        pytest.fail("\n".join(lines))


def _relerr(name, got, want, tol):
    got, want = got.double(), want.double()
    rel = float((got - want).norm() / want.norm().clamp_min(1e-30))
    assert rel <= tol, f"{name}: relative Frobenius error {rel:.4g} > {tol}"


# ----------------------------------------------------------------------------------- split-K wgrad
@pytest.mark.parametrize("T,Nout,Kin", [(40000, 768, 320), (5000, 96, 64), (129, 2304, 768)])
def test_gemm_wgrad_split_k_atomic(T, Nout, Kin):
    ops = _ops()
    dy, x = _rand((T, Nout), 0.5, seed=1), _rand((T, Kin), 0.5, seed=2)
    want = dy.float().t() @ x.float()
    out = torch.ones((Nout, Kin), device=DEV, dtype=torch.float32)
    ops.gemm(dy, x, trans_a=True, trans_b=True, epilogue=ops.EPI_SCALE_F32, accumulate=True, out=out)
    torch.cuda.synchronize()
    _relerr("wgrad split-K", out - 1.0, want, 2e-3)
    # second accumulation lands on top of the first
    ops.gemm(dy, x, trans_a=True, trans_b=True, epilogue=ops.EPI_SCALE_F32, accumulate=True, out=out, scale=0.5)
    _relerr("wgrad split-K x1.5", out - 1.0, want * 1.5, 2e-3)


def test_gemm_wgrad_bit_reproducible():
    """Ordered split-K accumulation (per-tile turnstile): the weight gradient is bit-identical from run to run and
    across different amounts of concurrent work; the atomic path agrees with it to rounding."""
    ops = _ops()
    T, Nout, Kin = 60000, 768, 768
    dy, x = _rand((T, Nout), 0.5, seed=1), _rand((T, Kin), 0.5, seed=2)
    outs = []
    ops.set_option(ops.OPT_WGRAD_DETERMINISTIC, 1)
    try:
        for _ in range(4):
            out = torch.zeros((Nout, Kin), device=DEV, dtype=torch.float32)
            ops.gemm(dy, x, trans_a=True, trans_b=True, epilogue=ops.EPI_SCALE_F32, accumulate=True, out=out)
            outs.append(out)
        torch.cuda.synchronize()
    finally:
        ops.set_option(ops.OPT_WGRAD_DETERMINISTIC, 0)
    for o in outs[1:]:
        assert torch.equal(o, outs[0]), "ordered split-K accumulation is not bit-reproducible"
    # default path (fp32 atomics): run-to-run spread bounded at rounding level, and equal to the ordered sum to rounding
    atomics = []
    for _ in range(3):
        a = torch.zeros((Nout, Kin), device=DEV, dtype=torch.float32)
        ops.gemm(dy, x, trans_a=True, trans_b=True, epilogue=ops.EPI_SCALE_F32, accumulate=True, out=a)
        atomics.append(a)
    torch.cuda.synchronize()
    atomic = atomics[0]
    for a in atomics[1:]:
        _relerr("atomic split-K run-to-run spread", a, atomic, 1e-6)
    _relerr("atomic vs ordered split-K", atomic, outs[0], 1e-5)
    _relerr("ordered split-K vs fp32", outs[0], dy.float().t() @ x.float(), 2e-3)


# ------------------------------------------------------------------------------------- attention
def _attn_ref_autograd(qkv, dout, cu, heads, window, positions=None, table=None):
    """fp32 autograd reference; if a rope table is given, `qkv` is the un-rotated projection."""
    T = qkv.shape[0]
    x = qkv.float().clone().requires_grad_(True)
    q3 = x.view(T, 3, heads, 64)
    if table is not None:
        cos, sin = table[positions.long(), :, 0], table[positions.long(), :, 1]
        cos, sin = torch.cat((cos, cos), -1)[:, None, None], torch.cat((sin, sin), -1)[:, None, None]
        rot = torch.cat((-q3[..., 32:], q3[..., :32]), dim=-1)
        qk = (q3 * cos + rot * sin)[:, :2]
        q3 = torch.cat((qk, q3[:, 2:]), dim=1)
    rotated = q3.reshape(T, -1)
    outs = []
    for b in range(len(cu) - 1):
        s, e = cu[b], cu[b + 1]
        q, k, v = (q3[s:e, i].transpose(0, 1) for i in range(3))
        sc = q @ k.transpose(1, 2) / 8.0
        if window >= 0:
            idx = torch.arange(e - s, device=qkv.device)
            sc = sc.masked_fill((idx[:, None] - idx[None, :]).abs() > window, float("-inf"))
        outs.append((sc.softmax(-1) @ v).transpose(0, 1).reshape(e - s, heads * 64))
    out = torch.cat(outs)
    out.backward(dout.float())
    return rotated.detach(), out.detach(), x.grad


@pytest.mark.parametrize("lens,heads,window,rope", [
    ([128], 1, -1, False), ([64], 1, -1, False), ([300, 77, 129, 512, 1], 2, -1, False),
    ([300, 77, 129, 512, 1], 2, 64, False), ([1000, 613], 3, -1, True), ([1000, 613, 190], 3, 64, True),
    ([25, 17, 21, 19], 4, -1, True), ([800] * 2, 8, 64, False)])
def test_attention_bwd(lens, heads, window, rope):
    ops = _ops()
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    T, H = cu[-1], heads * 64
    raw = _rand((T, 3 * H), 1.0, seed=5)
    dout = _rand((T, H), 1.0, seed=6)
    cu_t = torch.tensor(cu, dtype=torch.int32, device=DEV)
    pos = tab = None
    if rope:
        pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).to(DEV)
        tab = ops.rope_table(10000.0, 1024, DEV)
    rotated, want_out, want_dqkv = _attn_ref_autograd(raw, dout, cu, heads, window, pos, tab)
    qkv = rotated.to(torch.bfloat16).contiguous() if rope else raw
    if rope:  # reference gradient must see the same bf16-rounded rotated q/k the kernel sees
        _, want_out, _ = _attn_ref_autograd(qkv, dout, cu, heads, window)
    lse = torch.empty((heads, T), device=DEV, dtype=torch.float32)
    out = ops.attn_varlen_fwd(qkv, cu_t, max(lens), heads, window, lse=lse)
    dqkv = ops.attn_varlen_bwd(qkv, out, dout, lse, cu_t, max(lens), heads, window, positions=pos, rope_table=tab)
    torch.cuda.synchronize()
    assert torch.isfinite(dqkv.float()).all()
    tag = f"attn_bwd lens={lens} h={heads} w={window} rope={rope}"
    want = want_dqkv.view(T, 3, H)
    got = dqkv.float().view(T, 3, H)
    for i, n in enumerate("qkv"):
        _relerr(f"{tag} d{n}", got[:, i], want[:, i], 2e-2)
        _close(f"{tag} d{n}", got[:, i], want[:, i], 0.03 * float(want[:, i].abs().max()) + 1e-3, 5e-2)


@pytest.mark.parametrize("name", ["mixed", "metadata", "ones", "full"])
@pytest.mark.parametrize("rope", [False, True])
def test_attention_bwd_packed(name, rope):
    """Packed short-sequence backward (one kernel, 5 GEMMs per group of sequences) vs fp32 autograd and vs the
    one-tile-per-sequence kernels."""
    from test_kernels_gpu import _packed_lens
    ops = _ops()
    lens, heads = _packed_lens(name), 4
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    T, H = cu[-1], heads * 64
    raw = _rand((T, 3 * H), 1.0, seed=5)
    dout = _rand((T, H), 1.0, seed=6)
    cu_t = torch.tensor(cu, dtype=torch.int32, device=DEV)
    pos = tab = None
    if rope:
        pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).to(DEV)
        tab = ops.rope_table(10000.0, 128, DEV)
    rotated, _, want_dqkv = _attn_ref_autograd(raw, dout, cu, heads, -1, pos, tab)
    qkv = rotated.to(torch.bfloat16).contiguous() if rope else raw
    groups = ops.attn_pack_groups(cu_t, T)
    lse = torch.empty((heads, T), device=DEV, dtype=torch.float32)
    out = ops.attn_varlen_fwd(qkv, cu_t, max(lens), heads, -1, lse=lse, groups=groups)
    dqkv = ops.attn_varlen_bwd(qkv, out, dout, lse, cu_t, max(lens), heads, -1, positions=pos, rope_table=tab,
                               groups=groups)
    dqkv_tile = ops.attn_varlen_bwd(qkv, out, dout, lse, cu_t, max(lens), heads, -1, positions=pos, rope_table=tab)
    torch.cuda.synchronize()
    assert torch.isfinite(dqkv.float()).all()
    want = want_dqkv.view(T, 3, H)
    got = dqkv.float().view(T, 3, H)
    ref2 = dqkv_tile.float().view(T, 3, H)
    for i, n in enumerate("qkv"):
        tag = f"attn_bwd packed {name} rope={rope} d{n}"
        if name == "ones" and n != "v":
            # one key per query: dq = dk = 0 exactly; P (dP - delta) only cancels up to fp32 summation order
            _close(tag, got[:, i], want[:, i], 1e-3 * float(want.abs().max()), 0.0)
            continue
        _relerr(tag, got[:, i], want[:, i], 2e-2)
        _close(tag, got[:, i], want[:, i], 0.03 * float(want[:, i].abs().max()) + 1e-3, 5e-2)
        _relerr(tag + " vs tile kernels", got[:, i], ref2[:, i], 1e-2)


@pytest.mark.parametrize("outer_per_cta", [1, 2, 3, 16])
@pytest.mark.parametrize("window", [-1, 64, 0, 200])
def test_attention_bwd_streaming(outer_per_cta, window):
    """Several outer tiles per CTA (double-buffered 128-row operands and dQ accumulator, write-out one tile
    late, single-tile outer tiles at sequence ends): same gradients whatever the split."""
    ops = _ops()
    ops.set_option(ops.OPT_BWD_OUTER_PER_CTA, outer_per_cta)
    try:
        _attention_bwd_streaming_case(ops, outer_per_cta, window)
    finally:
        ops.set_option(ops.OPT_BWD_OUTER_PER_CTA, 0)


def _attention_bwd_streaming_case(ops, outer_per_cta, window):
    lens, heads = [1100, 257, 1, 640, 129, 385], 2
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    T, H = cu[-1], heads * 64
    raw = _rand((T, 3 * H), 1.0, seed=7)
    dout = _rand((T, H), 1.0, seed=8)
    cu_t = torch.tensor(cu, dtype=torch.int32, device=DEV)
    pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).to(DEV)
    tab = ops.rope_table(10000.0, 2048, DEV)
    rotated, _, want_dqkv = _attn_ref_autograd(raw, dout, cu, heads, window, pos, tab)
    qkv = rotated.to(torch.bfloat16).contiguous()
    lse = torch.empty((heads, T), device=DEV, dtype=torch.float32)
    out = ops.attn_varlen_fwd(qkv, cu_t, max(lens), heads, window, lse=lse)
    dqkv = ops.attn_varlen_bwd(qkv, out, dout, lse, cu_t, max(lens), heads, window, positions=pos, rope_table=tab)
    torch.cuda.synchronize()
    assert torch.isfinite(dqkv.float()).all()
    want = want_dqkv.view(T, 3, H)
    got = dqkv.float().view(T, 3, H)
    scale = float(want.abs().max())
    for i, n in enumerate("qkv"):
        tag = f"attn_bwd streaming opc={outer_per_cta} w={window} d{n}"
        if window == 0 and n != "v":
            # a single key per query: dq = dk = 0 exactly; the kernel's P (dP - delta) only cancels up to the
            # bf16 rounding of the forward output, so compare on the scale of dv (window 0 is here for the
            # one-inner-tile-per-outer-tile path of the streaming kernels)
            _close(tag, got[:, i], want[:, i], 0.03 * scale, 5e-2)
            continue
        _relerr(tag, got[:, i], want[:, i], 2e-2)
        _close(tag, got[:, i], want[:, i], 0.03 * float(want[:, i].abs().max()) + 1e-3, 5e-2)


# --------------------------------------------------------------------------------- row-wise kernels
@pytest.mark.parametrize("H,with_res", [(768, True), (512, False), (256, True), (64, True), (1024, False)])
def test_layernorm_bwd(H, with_res):
    ops = _ops()
    rows = 3001
    x = _rand((rows, H), 2.0, seed=1) + 0.5
    dy = _rand((rows, H), 1.0, seed=2)
    g = _rand((H,), 0.2, seed=3, dtype=torch.float32) + 1.0
    res = _rand((rows, H), 1.0, seed=4) if with_res else None
    xr = x.float().clone().requires_grad_(True)
    gr = g.clone().requires_grad_(True)
    F.layer_norm(xr, (H,), gr, None, 1e-5).backward(dy.float())
    want_dx = xr.grad + (res.float() if with_res else 0)
    dgamma = torch.full((H,), 2.0, device=DEV)
    dx = ops.layernorm_bwd(x, dy, g, 1e-5, dres=res, dgamma=dgamma)
    torch.cuda.synchronize()
    _close("ln_bwd dx", dx, want_dx, 3e-2, 2e-2)
    _relerr("ln_bwd dgamma", dgamma - 2.0, gr.grad, 2e-3)


def test_embed_gather_ln_bwd():
    ops = _ops()
    B, L, H, vocab, A = 3, 50, 128, 300, 5
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, vocab - 1, (B, L), generator=g)
    ids[:, 1:1 + A] = vocab - 1
    lens = torch.tensor([50, 20, 33])
    mask = torch.arange(L)[None] < lens[:, None]
    src = mask.flatten().nonzero().flatten().to(torch.int32)
    is_audio = (ids == vocab - 1).flatten()
    slot_all = torch.where(is_audio, torch.cumsum(is_audio.int(), 0) - 1,
                           torch.full_like(is_audio, -1, dtype=torch.int32))
    slot = slot_all[src.long()].to(torch.int32)
    tok, aud = _rand((vocab, H), seed=1), _rand((B * A, H), seed=2)
    gam = _rand((H,), 0.1, seed=3, dtype=torch.float32) + 1.0
    T = src.numel()
    dy = _rand((T, H), seed=4)
    # autograd reference
    tokr, audr, gamr = tok.float().requires_grad_(True), aud.float().requires_grad_(True), gam.clone().requires_grad_(True)
    emb = tokr[ids.flatten().to(DEV)]
    emb = torch.where(is_audio.to(DEV)[:, None], torch.zeros_like(emb), emb)
    scat = torch.zeros_like(emb)
    scat[is_audio.to(DEV)] = audr
    F.layer_norm((emb + scat)[src.long().to(DEV)], (H,), gamr, None, 1e-5).backward(dy.float())
    d_tok = torch.zeros((vocab, H), device=DEV)
    d_aud = torch.zeros((B * A, H), device=DEV, dtype=torch.bfloat16)
    dgam = torch.zeros((H,), device=DEV)
    ops.embed_gather_ln_bwd(ids.to(DEV).flatten(), src.to(DEV), slot.to(DEV), tok, aud, gam, dy, 1e-5, d_tok, d_aud, dgam)
    torch.cuda.synchronize()
    _close("embed_bwd d_tok", d_tok, tokr.grad, 2e-3, 2e-3)
    _close("embed_bwd d_audio", d_aud, audr.grad, 3e-2, 2e-2)
    _relerr("embed_bwd dgamma", dgam, gamr.grad, 2e-3)


@pytest.mark.parametrize("I", [1152, 96])
def test_geglu_bwd(I):
    ops = _ops()
    rows = 777
    ug_plain = _rand((rows, 2 * I), 1.5, seed=1)  # [u | g]
    dh = _rand((rows, I), 1.0, seed=2)
    ug_il = ops.interleave_wi(ug_plain.t().contiguous()).t().contiguous()
    r = ug_plain.float().clone().requires_grad_(True)
    hh = F.gelu(r[:, :I]) * r[:, I:]
    hh.backward(dh.float())
    dug, h = ops.geglu_bwd(ug_il, dh)
    torch.cuda.synchronize()
    got = ops.deinterleave_wi(dug.t().contiguous()).t()
    _close("geglu_bwd dug", got, r.grad, 3e-2, 2e-2)
    _close("geglu_bwd h", h, hh.detach(), 3e-2, 1e-2)


def test_gelu_fwd_bwd_colsum():
    ops = _ops()
    z, dy = _rand((1000, 512), 2.0, seed=1), _rand((1000, 512), 1.0, seed=2)
    zr = z.float().clone().requires_grad_(True)
    y = F.gelu(zr)
    y.backward(dy.float())
    _close("gelu_fwd", ops.gelu_fwd(z), y.detach(), 2e-2, 1e-2)
    _close("gelu_bwd", ops.gelu_bwd(z, dy), zr.grad, 2e-2, 1e-2)
    out = torch.full((512,), -1.0, device=DEV)
    ops.colsum_f32(dy, out)
    _relerr("colsum", out + 1.0, dy.float().sum(0), 1e-4)
    big = _rand((70001, 64), 1.0, seed=3)
    out2 = torch.zeros((64,), device=DEV)
    ops.colsum_f32(big, out2)
    _relerr("colsum tall", out2, big.float().sum(0), 1e-4)


@pytest.mark.parametrize("mean_pool", [False, True])
def test_pool_l2norm_bwd(mean_pool):
    ops = _ops()
    lens = [300, 1, 77, 129]
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    H, P = 128, 64
    cu_t = torch.tensor(cu, dtype=torch.int32, device=DEV)
    dp = _rand((len(lens), H), seed=1)
    base = _rand((cu[-1], H), seed=2)
    want = torch.zeros((cu[-1], H), device=DEV)
    for i, n in enumerate(lens):
        if mean_pool:
            want[cu[i]:cu[i + 1]] = dp[i].float() / n
        else:
            want[cu[i]] = dp[i].float()
    dh = torch.full((cu[-1], H), 7.0, device=DEV, dtype=torch.bfloat16)
    ops.pool_bwd(dp, cu_t, mean_pool, dh, accumulate=False)
    _close("pool_bwd", dh, want, 1e-3, 1e-2)
    dh2 = base.clone()
    ops.pool_bwd(dp, cu_t, mean_pool, dh2, accumulate=True)
    _close("pool_bwd acc", dh2, want + base.float(), 3e-2, 1e-2)
    # l2norm
    e = _rand((9, P), 1.0, seed=3, dtype=torch.float32)
    de = _rand((9, P), 1.0, seed=4, dtype=torch.float32)
    er = e.clone().requires_grad_(True)
    (er / er.pow(2).sum(-1, keepdim=True).sqrt()).backward(de)
    inv = 1.0 / e.norm(dim=-1)
    _close("l2norm_bwd", ops.l2norm_bwd(e, inv.contiguous(), de), er.grad, 1e-2, 1e-2)


@pytest.mark.parametrize("B,V", [(8, 1), (5, 3), (64, 8), (33, 17)])
def test_clip_loss_bwd(B, V):
    ops = _ops()
    S = _rand((B, V, B), 3.0, seed=1, dtype=torch.float32)
    g = torch.Generator().manual_seed(2)
    t = torch.randint(0, V, (B,), generator=g).to(DEV)
    Sr = S.clone().requires_grad_(True)
    rows = Sr[torch.arange(B, device=DEV), t]
    ml = F.cross_entropy(rows, torch.arange(B, device=DEV))
    bl = F.cross_entropy(Sr.permute(2, 0, 1).reshape(B, B * V), torch.arange(B, device=DEV) * V + t)
    ((ml + bl) / 2 * 1.7).backward()
    S2 = S.view(B * V, B).contiguous()
    loss, row_lse, col_lse = ops.clip_loss_fwd(S2, t.to(torch.int32), V)
    dls = torch.zeros((1,), device=DEV)
    gout = torch.tensor([1.7], device=DEV)
    dS = ops.clip_loss_bwd(S2, t.to(torch.int32), row_lse, col_lse, V, gout, dls)
    torch.cuda.synchronize()
    want = Sr.grad.view(B * V, B)
    _close("clip_loss_bwd dS", dS, want, 2e-3 * float(want.abs().max()) + 1e-6, 1e-2)
    assert abs(float(dls) - float((want * S2).sum())) <= 2e-2 * float((want * S2).abs().sum())


@pytest.mark.parametrize("B,Fr,Co", [(2, 320, 64), (3, 1600, 512)])
def test_conv_training_path(B, Fr, Co):
    """Implicit-GEMM conv forward (bias, pre-activation kept) + gelu == conv1d + gelu, and the implicit weight
    gradient / col2im input gradient against autograd."""
    ops = _ops()
    C = 80
    x = _rand((B, C, Fr), seed=1, dtype=torch.float32)
    w1, b1 = _rand((Co, C, 3), 0.1, seed=2), _rand((Co,), 0.1, seed=3, dtype=torch.float32)
    w2, b2 = _rand((Co, Co, 3), 0.05, seed=4), _rand((Co,), 0.1, seed=5, dtype=torch.float32)
    w1p, w2p = ops.pack_conv_weight(w1), ops.pack_conv_weight(w2)
    xt = ops.transpose_cast(x)
    z1 = ops.conv1d_k3(xt, w1p, b1, stride=1, gelu=False)
    y1 = ops.gelu_fwd(z1)
    z2 = ops.conv1d_k3(y1, w2p, b2, stride=2, gelu=False)
    y2 = ops.gelu_fwd(z2).view(B * Fr // 2, Co)
    # autograd reference on the same bf16-rounded intermediates
    xr = x.bfloat16().float()
    w1r, b1r = w1.float().requires_grad_(True), b1.clone().requires_grad_(True)
    w2r, b2r = w2.float().requires_grad_(True), b2.clone().requires_grad_(True)
    r1 = F.gelu(F.conv1d(xr, w1r, b1r, padding=1))
    r2 = F.gelu(F.conv1d(r1, w2r, b2r, stride=2, padding=1)).permute(0, 2, 1).reshape(B * Fr // 2, Co)
    _close("conv train y2", y2, r2.detach(), 3e-2, 2e-2)
    dy2 = _rand((B * Fr // 2, Co), 1.0, seed=6)
    r2.backward(dy2.float())
    # ours
    dz2 = ops.gelu_bwd(z2.view(B * Fr // 2, Co), dy2)
    db2 = torch.zeros((Co,), device=DEV)
    ops.colsum_f32(dz2, db2)
    dw2p = torch.zeros((Co, w2p.shape[1]), device=DEV)
    ops.conv1d_k3_wgrad(dz2.view(B, Fr // 2, Co), y1, dw2p, stride=2)
    da2 = ops.gemm(dz2, w2p, trans_b=True)
    dz1 = ops.conv2_col2im_gelu_bwd(da2, z1.view(B, Fr, Co))
    db1 = torch.zeros((Co,), device=DEV)
    ops.colsum_f32(dz1.view(B * Fr, Co), db1)
    dw1p = torch.zeros((Co, w1p.shape[1]), device=DEV)
    ops.conv1d_k3_wgrad(dz1.view(B, Fr, Co), xt, dw1p, stride=1)
    torch.cuda.synchronize()
    _relerr("conv db2", db2, b2r.grad, 2e-2)
    _relerr("conv dw2", ops.unpack_conv_weight_grad(dw2p, Co), w2r.grad, 2e-2)
    _relerr("conv db1", db1, b1r.grad, 3e-2)
    _relerr("conv dw1", ops.unpack_conv_weight_grad(dw1p, C), w1r.grad, 3e-2)
    # the padded channel columns of the conv1 weight gradient (c >= C_in) stay exactly zero
    assert float(dw1p.view(Co, 3, -1)[:, :, C:].abs().max()) == 0.0
    # second call accumulates
    ops.conv1d_k3_wgrad(dz1.view(B, Fr, Co), xt, dw1p, stride=1)
    _relerr("conv dw1 x2", ops.unpack_conv_weight_grad(dw1p, C), 2 * w1r.grad, 3e-2)


@pytest.mark.parametrize("V", [3968, 500, 2])
def test_vocab_cross_entropy_fwd_bwd(V):
    ops = _ops()
    rows, L = 900, 1000
    ld = (V + 7) // 8 * 8
    buf = torch.full((rows, ld), 3.0, device=DEV, dtype=torch.bfloat16)
    logits = buf[:, :V]
    logits.copy_(_rand((rows, V), 2.0, seed=1))
    g = torch.Generator().manual_seed(3)
    labels = torch.randint(0, V, (L,), generator=g)
    labels[torch.rand(L, generator=g) < 0.7] = -100
    src = torch.randperm(L, generator=g)[:rows].sort().values.to(torch.int32)
    tgt = labels[src.long()].to(DEV)
    ref = logits.float().clone().requires_grad_(True)
    want = F.cross_entropy(ref, tgt, ignore_index=-100, reduction="sum")
    (want * 0.37).backward()
    row_lse, loss_sum, count = ops.vocab_ce_fwd(logits, V, labels.to(DEV), src.to(DEV))
    assert abs(float(loss_sum) - float(want)) <= 2e-3 * abs(float(want)) + 1e-3
    assert int(count) == int((tgt != -100).sum())
    ops.vocab_ce_bwd(logits, V, labels.to(DEV), src.to(DEV), row_lse, torch.tensor([0.37], device=DEV))
    torch.cuda.synchronize()
    _close("vocab_ce_bwd", logits, ref.grad, 2e-3, 2e-2)
    if ld > V:
        assert float(buf[:, V:].abs().max()) == 0.0


def test_gather_scatter_rows():
    ops = _ops()
    x = _rand((500, 128), seed=1)
    idx = torch.tensor([3, 7, 8, 250, 499], dtype=torch.int32, device=DEV)
    rows = ops.gather_rows(x, idx)
    assert torch.equal(rows, x[idx.long()])
    dx = _rand((500, 128), seed=2)
    want = dx.float().clone()
    want[idx.long()] += rows.float()
    ops.scatter_add_rows(rows, idx, dx)
    _close("scatter_add_rows", dx, want, 2e-2, 1e-2)
