"""Per-kernel numerics on a real B200: every CUDA kernel (called through the C ABI) against a plain
PyTorch fp32 reference of the same op on the same bf16-rounded inputs."""
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


def _report(name, got, want, atol, rtol):
    got, want = got.float(), want.float()
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    bad = err > tol
    if bad.any():
        idx = bad.nonzero()[:8].tolist()
        msg = [f"{name}: {int(bad.sum())}/{bad.numel()} mismatches, max abs err {float(err.max()):.4g}"]
        for i in idx:
            msg.append(f"  at {i}: got {float(got[tuple(i)]):.5g} want {float(want[tuple(i)]):.5g}")
        if got.dim() == 2:
            # localise: error per 128-row block and per 32-column chunk
            R, C = got.shape
            rb = [float(err[r:r + 128].max()) for r in range(0, R, 128)][:16]
            cb = [float(err[:, c:c + 32].max()) for c in range(0, C, 32)][:24]
            msg.append(f"  max err per 128-row block: {['%.3g' % v for v in rb]}")
            msg.append(f"  max err per 32-col chunk : {['%.3g' % v for v in cb]}")
        pytest.fail("\n".join(msg))


# ------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 384, 128), (1000, 768, 768), (4096, 2304, 768),
                                   (77, 64, 240), (515, 1536, 512), (100, 7, 64), (9, 3, 512),
                                   (700, 300, 200), (257, 136, 72)])  # pair-MMA form with M / N / K tails
def test_gemm_store(M, N, K):
    ops = _ops()
    a, b = _rand((M, K), seed=1), _rand((N, K), 0.05, seed=2)
    out = ops.gemm(a, b)
    torch.cuda.synchronize()
    _report(f"gemm_store {M}x{N}x{K}", out, a.float() @ b.float().t(), 2e-2, 1e-2)


def test_gemm_residual_inplace():
    ops = _ops()
    M, N, K = 1000, 768, 1152
    a, b, r = _rand((M, K), seed=1), _rand((N, K), 0.03, seed=2), _rand((M, N), seed=3)
    want = a.float() @ b.float().t() + r.float()
    x = r.clone()
    ops.gemm(a, b, epilogue=ops.EPI_RESIDUAL, out=x, aux=x)
    _report("gemm_residual", x, want, 3e-2, 1e-2)


def test_gemm_gelu_and_bias():
    ops = _ops()
    M, N, K = 600, 512, 240
    a, b = _rand((M, K), seed=1), _rand((N, K), 0.1, seed=2)
    bias = _rand((N,), 0.5, seed=3, dtype=torch.float32)
    acc = a.float() @ b.float().t()
    _report("gemm_gelu", ops.gemm(a, b, epilogue=ops.EPI_GELU), F.gelu(acc), 2e-2, 1e-2)
    _report("gemm_bias_gelu", ops.gemm(a, b, epilogue=ops.EPI_BIAS_GELU, aux=bias), F.gelu(acc + bias), 2e-2, 1e-2)
    _report("gemm_bias", ops.gemm(a, b, epilogue=ops.EPI_BIAS, aux=bias), acc + bias, 2e-2, 1e-2)


@pytest.mark.parametrize("I,H", [(1152, 768), (96, 128)])
def test_gemm_geglu(I, H):
    ops = _ops()
    M = 700
    a, wi = _rand((M, H), seed=1), _rand((2 * I, H), 0.05, seed=2)
    acc = a.float() @ wi.float().t()
    want = F.gelu(acc[:, :I]) * acc[:, I:]
    wi_il = ops.interleave_wi(wi).contiguous()
    assert torch.equal(ops.deinterleave_wi(wi_il), wi)
    got = ops.gemm(a, wi_il, epilogue=ops.EPI_GEGLU)
    _report("gemm_geglu", got, want, 2e-2, 2e-2)
    raw = torch.empty((M, 2 * I), device=DEV, dtype=torch.bfloat16)
    got2 = ops.gemm(a, wi_il, epilogue=ops.EPI_GEGLU_SAVE, c2=raw)
    _report("gemm_geglu_save.out", got2, want, 2e-2, 2e-2)
    _report("gemm_geglu_save.raw", ops.deinterleave_wi(raw.t().contiguous()).t(), acc, 2e-2, 1e-2)


def test_gemm_cluster_forms_agree():
    """The three GEMM forms (single CTAs, CTA pairs with B multicast, CTA pairs with one cta_group::2 MMA = default)
    accumulate in the same order: bitwise-equal outputs, staged and direct epilogues, odd numbers of M tiles."""
    ops = _ops()
    M, H, I = 128 * 5 + 40, 768, 1152
    a, w = _rand((M, H), seed=11), _rand((2 * I, H), 0.05, seed=12)
    res = _rand((M, 2 * I), seed=13)
    wi_il = ops.interleave_wi(w).contiguous()
    dy = _rand((M, 256), 0.5, seed=14)
    outs = {}
    try:
        for mode in (2, 3, 1):
            ops.set_option(ops.OPT_GEMM_CLUSTER, mode)
            raw = torch.empty((M, 2 * I), device=DEV, dtype=torch.bfloat16)
            dw = torch.zeros((256, H), device=DEV)
            ops.set_option(ops.OPT_WGRAD_DETERMINISTIC, 1)
            ops.gemm(dy, a, trans_a=True, trans_b=True, epilogue=ops.EPI_SCALE_F32, accumulate=True, out=dw)
            ops.set_option(ops.OPT_WGRAD_DETERMINISTIC, 0)
            outs[mode] = (ops.gemm(a, w), ops.gemm(a, w, epilogue=ops.EPI_RESIDUAL, aux=res),
                          ops.gemm(a, wi_il, epilogue=ops.EPI_GEGLU_SAVE, c2=raw), raw, dw,
                          ops.gemm(res, w, trans_b=True))
    finally:
        ops.set_option(ops.OPT_GEMM_CLUSTER, 2)
        ops.set_option(ops.OPT_WGRAD_DETERMINISTIC, 0)
    torch.cuda.synchronize()
    for mode in (3, 1):
        for k, (x, y) in enumerate(zip(outs[2], outs[mode])):
            assert torch.equal(x, y), f"form {mode} differs from the pair-MMA form in output {k}"
    _report("gemm_pair_form", outs[2][0], a.float() @ w.float().t(), 2e-2, 1e-2)


def test_gemm_rope():
    ops = _ops()
    heads, T = 3, 500
    H = heads * 64
    a, w = _rand((T, H), seed=1), _rand((3 * H, H), 0.08, seed=2)
    pos = (torch.arange(T, dtype=torch.int32) % 211).to(DEV)
    tab = ops.rope_table(160000.0, 256, DEV)
    got = ops.gemm(a, w, epilogue=ops.EPI_ROPE, positions=pos, rope_table=tab, rope_cols=2 * H)
    acc = (a.float() @ w.float().t()).view(T, 3, heads, 64)
    cos, sin = tab[pos.long(), :, 0], tab[pos.long(), :, 1]  # [T, 32]
    cos, sin = torch.cat((cos, cos), -1)[:, None, None], torch.cat((sin, sin), -1)[:, None, None]
    rot = torch.cat((-acc[..., 32:], acc[..., :32]), dim=-1)
    want = acc.clone()
    want[:, :2] = (acc * cos + rot * sin)[:, :2]
    _report("gemm_rope", got, want.view(T, 3 * H), 3e-2, 1e-2)


def test_gemm_scale_f32_accumulate():
    ops = _ops()
    M, N, K = 260, 200, 512
    a, b = _rand((M, K), 0.1, seed=1), _rand((N, K), 0.1, seed=2)
    acc = a.float() @ b.float().t()
    out = ops.gemm(a, b, epilogue=ops.EPI_SCALE_F32, scale=14.25)
    _report("gemm_scale_f32", out, acc * 14.25, 1e-3, 1e-3)
    ops.gemm(a, b, epilogue=ops.EPI_SCALE_F32, scale=0.5, accumulate=True, out=out)
    _report("gemm_scale_f32_acc", out, acc * 14.75, 1e-3, 1e-3)
    # row pitch that is not 16-byte aligned (logits of an odd batch): scalar store path
    out3 = ops.gemm(a, b[:3].contiguous(), epilogue=ops.EPI_SCALE_F32, scale=2.0)
    _report("gemm_scale_f32_n3", out3, acc[:, :3] * 2.0, 1e-3, 1e-3)


def test_gemm_transposed_operands():
    """dgrad form (B stored [K,N]) and wgrad form (A stored [K,M], B stored [K,N])."""
    ops = _ops()
    T, Nout, Kin = 900, 768, 320
    dy, w, x = _rand((T, Nout), seed=1), _rand((Nout, Kin), 0.05, seed=2), _rand((T, Kin), seed=3)
    dx = ops.gemm(dy, w, trans_b=True)  # dX = dY @ W
    _report("gemm_dgrad", dx, dy.float() @ w.float(), 5e-2, 1e-2)
    dw = ops.gemm(dy, x, trans_a=True, trans_b=True, epilogue=ops.EPI_SCALE_F32)  # dW = dY^T @ X
    _report("gemm_wgrad", dw, dy.float().t() @ x.float(), 5e-2, 1e-2)


# ------------------------------------------------------------------------------------- attention
@pytest.fixture
def blocks_per_cta(request):
    """Force how many 256-query blocks one forward CTA streams (cm3p_set_option, restored afterwards)."""
    ops = _ops()
    ops.set_option(ops.OPT_FWD_BLOCKS_PER_CTA, request.param)
    yield request.param
    ops.set_option(ops.OPT_FWD_BLOCKS_PER_CTA, 0)


def _attn_ref(qkv, cu, heads, window):
    T = qkv.shape[0]
    out = torch.empty((T, heads * 64), device=qkv.device, dtype=torch.float32)
    q3 = qkv.float().view(T, 3, heads, 64)
    for b in range(len(cu) - 1):
        s, e = cu[b], cu[b + 1]
        q, k, v = (q3[s:e, i].transpose(0, 1) for i in range(3))  # [h, L, 64]
        sc = q @ k.transpose(1, 2) / 8.0
        if window >= 0:
            idx = torch.arange(e - s, device=qkv.device)
            sc = sc.masked_fill((idx[:, None] - idx[None, :]).abs() > window, float("-inf"))
        out[s:e] = (sc.softmax(-1) @ v).transpose(0, 1).reshape(e - s, heads * 64)
    return out


@pytest.mark.parametrize("lens,heads,window", [
    ([128], 1, -1), ([300, 77, 129, 512, 1], 2, -1), ([300, 77, 129, 512, 1], 2, 64),
    ([2000, 613, 1500], 12, -1), ([2000, 613, 1500], 12, 64), ([800] * 4, 8, 64), ([25, 17, 21, 19], 4, -1)])
def test_attention_fwd(lens, heads, window):
    ops = _ops()
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    T = cu[-1]
    qkv = _rand((T, 3 * heads * 64), 1.0, seed=5)
    cu_t = torch.tensor(cu, dtype=torch.int32, device=DEV)
    lse = torch.empty((heads, T), device=DEV, dtype=torch.float32)
    out = ops.attn_varlen_fwd(qkv, cu_t, max(lens), heads, window, lse=lse)
    torch.cuda.synchronize()
    want = _attn_ref(qkv, cu, heads, window)
    _report(f"attn lens={lens} h={heads} w={window}", out, want, 2e-2, 2e-2)
    # lse (log2 domain) against the reference for the first head of the first sequence
    L = lens[0]
    q3 = qkv.float().view(T, 3, heads, 64)
    sc = (q3[:L, 0, 0] @ q3[:L, 1, 0].t()) / 8.0
    if window >= 0:
        idx = torch.arange(L, device=DEV)
        sc = sc.masked_fill((idx[:, None] - idx[None, :]).abs() > window, float("-inf"))
    _report("attn lse", lse[0, :L], torch.logsumexp(sc, -1) / math.log(2.0), 2e-2, 1e-3)


@pytest.mark.parametrize("blocks_per_cta", [1, 2, 3, 16], indirect=True)
@pytest.mark.parametrize("window", [-1, 64, 0, 200])
def test_attention_fwd_streaming(blocks_per_cta, window):
    """Several 256-query blocks per CTA (double-buffered Q, O reuse across blocks, inactive second Q tile in the
    last block, block counts that do not divide the sequence): same result whatever the split."""
    ops = _ops()
    lens, heads = [2000, 257, 1, 640, 1153, 129, 512], 2
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    T = cu[-1]
    qkv = _rand((T, 3 * heads * 64), 1.0, seed=11)
    cu_t = torch.tensor(cu, dtype=torch.int32, device=DEV)
    lse = torch.empty((heads, T), device=DEV, dtype=torch.float32)
    out = ops.attn_varlen_fwd(qkv, cu_t, max(lens), heads, window, lse=lse)
    torch.cuda.synchronize()
    _report(f"attn streaming bpc={blocks_per_cta} w={window}", out, _attn_ref(qkv, cu, heads, window), 2e-2, 2e-2)
    assert bool(torch.isfinite(lse).all())


def _packed_lens(name):
    g = torch.Generator().manual_seed(3)
    if name == "mixed":
        return [1, 17, 25, 128, 25, 17, 1, 1, 64, 64, 21, 128, 127, 2]
    if name == "metadata":  # the metadata tower: many sequences of 17..25 tokens (more than one 64-sequence chunk)
        return torch.randint(17, 26, (300,), generator=g).tolist()
    if name == "ones":
        return [1] * 200
    if name == "full":
        return [128] * 5
    raise KeyError(name)


@pytest.mark.parametrize("name", ["mixed", "metadata", "ones", "full"])
@pytest.mark.parametrize("heads", [1, 4])
def test_attention_fwd_packed(name, heads):
    """Packed short sequences (several sequences per 128-row tile, block-diagonal mask) == one tile per sequence
    == the fp32 reference; the group table covers every sequence exactly once."""
    ops = _ops()
    lens = _packed_lens(name)
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    T = cu[-1]
    qkv = _rand((T, 3 * heads * 64), 1.0, seed=5)
    cu_t = torch.tensor(cu, dtype=torch.int32, device=DEV)
    groups = ops.attn_pack_groups(cu_t, T)
    n_groups = int(groups.count.item())
    assert 0 < n_groups <= groups.max_groups
    tab = groups.table[:n_groups].cpu()
    tab = tab[tab[:, 0].argsort()]
    assert int(tab[0, 0]) == 0 and int(tab[-1, 1]) == len(lens)
    assert bool((tab[1:, 0] == tab[:-1, 1]).all())
    tok = torch.tensor(cu)[tab[:, 1].long()] - torch.tensor(cu)[tab[:, 0].long()]
    assert int(tok.max()) <= 128 and int(tok.min()) >= 1
    lse = torch.empty((heads, T), device=DEV, dtype=torch.float32)
    lse_p = torch.empty_like(lse)
    out = ops.attn_varlen_fwd(qkv, cu_t, max(lens), heads, -1, lse=lse)
    out_p = ops.attn_varlen_fwd(qkv, cu_t, max(lens), heads, -1, lse=lse_p, groups=groups)
    torch.cuda.synchronize()
    want = _attn_ref(qkv, cu, heads, -1)
    _report(f"attn packed {name} h={heads}", out_p, want, 2e-2, 2e-2)
    _report(f"attn packed vs tile kernel {name} h={heads}", out_p, out, 1e-2, 1e-2)
    _report(f"attn packed lse {name} h={heads}", lse_p, lse, 1e-3, 1e-4)


def test_attention_more_than_65535_sequences():
    """B*V = 256*256 + 1 metadata sequences: the sequence index lives in grid.x (grid.z stops at 65535)."""
    ops = _ops()
    n, heads = 65537, 1
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(1, 4, (n,), generator=g)
    cu = torch.zeros(n + 1, dtype=torch.int32)
    cu[1:] = lens.cumsum(0)
    T = int(cu[-1])
    qkv = _rand((T, 3 * heads * 64), 1.0, seed=9)
    cu_t = cu.to(DEV)
    out = ops.attn_varlen_fwd(qkv, cu_t, 3, heads, -1)                      # one tile per sequence
    out_p = ops.attn_varlen_fwd(qkv, cu_t, 3, heads, -1, groups=ops.attn_pack_groups(cu_t, T))
    torch.cuda.synchronize()
    # last sequence against the reference, and the two kernels against each other everywhere
    s, e = int(cu[-2]), int(cu[-1])
    q3 = qkv.float().view(T, 3, 64)
    want = ((q3[s:e, 0] @ q3[s:e, 1].t()) / 8.0).softmax(-1) @ q3[s:e, 2]
    _report("attn >65535 sequences (tail)", out[s:e], want, 2e-2, 2e-2)
    _report("attn >65535 sequences packed vs tile", out_p, out, 1e-2, 1e-2)


@pytest.mark.parametrize("trans", [False, True])
def test_gemm_grouped(trans):
    """Grouped GEMM (operands stacked along their outer dimension) == a loop over the groups; the shapes are the
    Newton-Schulz products of Muon over all layers at once."""
    ops = _ops()
    G, m, K = 5, 256, 320
    if not trans:
        x = _rand((G * m, K), 0.3, seed=1)                                # A_g = X_g X_g^T
        got = ops.gemm(x, x, groups=G)
        want = torch.cat([x[g * m:(g + 1) * m].float() @ x[g * m:(g + 1) * m].float().t() for g in range(G)])
        _report("gemm grouped X X^T", got, want, 5e-2, 2e-2)
        bmat = _rand((G * m, m), 0.1, seed=2)                             # X_g <- B_g X_g + aux  (B operand [K, N])
        aux = _rand((G * m, K), 1.0, seed=3)
        got = ops.gemm(bmat, x, trans_b=True, epilogue=ops.EPI_RESIDUAL, aux=aux, groups=G)
        want = torch.cat([bmat[g * m:(g + 1) * m].float() @ x[g * m:(g + 1) * m].float() for g in range(G)]) + aux.float()
        _report("gemm grouped B X + aux", got, want, 5e-2, 2e-2)
    else:
        R = 640                                                           # tall matrices: A_g = G_g^T G_g
        x = _rand((G * R, m), 0.3, seed=4)
        got = ops.gemm(x, x, trans_a=True, trans_b=True, groups=G)
        want = torch.cat([x[g * R:(g + 1) * R].float().t() @ x[g * R:(g + 1) * R].float() for g in range(G)])
        _report("gemm grouped G^T G", got, want, 5e-2, 2e-2)


# --------------------------------------------------------------------------------- row-wise kernels
@pytest.mark.parametrize("H", [768, 512, 256, 128, 64])
def test_layernorm(H):
    ops = _ops()
    x = _rand((1001, H), 2.0, seed=1) + 0.5
    g = _rand((H,), 0.2, seed=2, dtype=torch.float32) + 1.0
    stats = torch.empty((1001, 2), device=DEV, dtype=torch.float32)
    y = ops.layernorm(x, g, 1e-5, stats=stats)
    want = F.layer_norm(x.float(), (H,), g, None, 1e-5)
    _report("layernorm", y, want, 2e-2, 1e-2)
    _report("layernorm mean", stats[:, 0], x.float().mean(-1), 1e-4, 1e-4)
    _report("layernorm rstd", stats[:, 1], (x.float().var(-1, unbiased=False) + 1e-5).rsqrt(), 1e-4, 1e-4)


def test_embed_gather_ln():
    ops = _ops()
    B, L, H, vocab, A = 3, 50, 128, 300, 5
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, vocab - 1, (B, L), generator=g)
    ids[:, 1:1 + A] = vocab - 1  # audio token
    lens = torch.tensor([50, 20, 33])
    mask = torch.arange(L)[None] < lens[:, None]
    src = mask.flatten().nonzero().flatten().to(torch.int32)
    is_audio = (ids == vocab - 1).flatten()
    slot_all = torch.where(is_audio, torch.cumsum(is_audio.int(), 0) - 1, torch.full_like(is_audio, -1, dtype=torch.int32))
    slot = slot_all[src.long()].to(torch.int32)
    tok = _rand((vocab, H), seed=1)
    aud = _rand((B * A, H), seed=2)
    gam = _rand((H,), 0.1, seed=3, dtype=torch.float32) + 1.0
    y = ops.embed_gather_ln(ids.to(DEV), src.to(DEV), slot.to(DEV), tok, aud, gam, 1e-5, rows=src.numel())
    emb = tok.float()[ids.flatten().to(DEV)]
    emb[is_audio.to(DEV)] = aud.float()
    want = F.layer_norm(emb[src.long().to(DEV)], (H,), gam, None, 1e-5)
    _report("embed_gather_ln", y, want, 2e-2, 1e-2)


@pytest.mark.parametrize("B,Fr,Co", [(3, 320, 64), (2, 1600, 512), (5, 200, 128)])
def test_conv1d_gelu_both_layers(B, Fr, Co):
    """Implicit-GEMM conv1d (no im2col matrix): shifted tensor-map boxes per tap, zero padding = out-of-bounds fill,
    stride 2 through the frame-parity dimension, windows whose length is not a multiple of the 128-row tile."""
    ops = _ops()
    C = 80
    x = _rand((B, C, Fr), seed=1, dtype=torch.float32)
    w1, b1 = _rand((Co, C, 3), 0.1, seed=2), _rand((Co,), 0.1, seed=3, dtype=torch.float32)
    w2, b2 = _rand((Co, Co, 3), 0.05, seed=4), _rand((Co,), 0.1, seed=5, dtype=torch.float32)
    xt = ops.transpose_cast(x)
    _report("transpose_cast", xt.reshape(B * Fr, C), x.bfloat16().permute(0, 2, 1).reshape(B * Fr, C), 0.0, 0.0)
    y1 = ops.conv1d_k3(xt, ops.pack_conv_weight(w1), b1, stride=1, gelu=True)
    want1 = F.gelu(F.conv1d(x.bfloat16().float(), w1.float(), b1, padding=1)).permute(0, 2, 1)
    _report("conv1", y1.reshape(B * Fr, Co), want1.reshape(B * Fr, Co), 2e-2, 1e-2)
    y2 = ops.conv1d_k3(y1, ops.pack_conv_weight(w2), b2, stride=2, gelu=True)
    want2 = F.gelu(F.conv1d(y1.float().permute(0, 2, 1), w2.float(), b2, stride=2, padding=1)).permute(0, 2, 1)
    _report("conv2", y2.reshape(B * Fr // 2, Co), want2.reshape(B * Fr // 2, Co), 2e-2, 1e-2)
    z2 = ops.conv1d_k3(y1, ops.pack_conv_weight(w2), b2, stride=2, gelu=False)   # pre-activation (training path)
    want_z2 = F.conv1d(y1.float().permute(0, 2, 1), w2.float(), b2, stride=2, padding=1).permute(0, 2, 1)
    _report("conv2 pre-activation", z2.reshape(B * Fr // 2, Co), want_z2.reshape(B * Fr // 2, Co), 3e-2, 1e-2)


@pytest.mark.parametrize("mean_pool", [False, True])
def test_pool_project_normalize(mean_pool):
    ops = _ops()
    lens = [300, 1, 77, 129]
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    H, P = 128, 64
    x = _rand((cu[-1], H), seed=1)
    w = _rand((P, H), 0.1, seed=2)
    cu_t = torch.tensor(cu, dtype=torch.int32, device=DEV)
    pooled, proj, inv, emb, emb16 = ops.pool_project_normalize(x, cu_t, mean_pool, w)
    if mean_pool:
        want_pool = torch.stack([x.float()[cu[i]:cu[i + 1]].mean(0) for i in range(len(lens))])
    else:
        want_pool = x.float()[cu_t[:-1].long()]
    _report("pooled", pooled, want_pool, 1e-2, 1e-2)
    e = pooled.float() @ w.float().t()
    _report("proj", proj, e, 1e-3, 1e-3)
    _report("embeds", emb, e / e.norm(dim=-1, keepdim=True), 1e-4, 1e-3)
    _report("embeds16", emb16, e / e.norm(dim=-1, keepdim=True), 1e-2, 1e-2)


@pytest.mark.parametrize("B,V", [(8, 1), (5, 3), (64, 8), (33, 17)])
def test_clip_loss(B, V):
    ops = _ops()
    S = _rand((B, V, B), 3.0, seed=1, dtype=torch.float32)
    g = torch.Generator().manual_seed(2)
    t = torch.randint(0, V, (B,), generator=g).to(DEV)
    loss, row_lse, col_lse = ops.clip_loss_fwd(S.view(B * V, B), t.to(torch.int32), V)
    rows = S[torch.arange(B, device=DEV), t]
    ml = F.cross_entropy(rows, torch.arange(B, device=DEV))
    bl = F.cross_entropy(S.permute(2, 0, 1).reshape(B, B * V), torch.arange(B, device=DEV) * V + t)
    _report("clip_loss", loss, ((ml + bl) / 2).reshape(1), 1e-4, 1e-4)
    _report("col_lse", col_lse, torch.logsumexp(S.view(B * V, B), 0), 1e-4, 1e-4)


# ------------------------------------------------------------------- LayerNorm folded into GEMMs
def test_gemm_layernorm_folding():
    """residual GEMM emits row statistics; the next GEMM (RoPE / GeGLU epilogue) applies the LayerNorm:
    rstd * (x . W'^T - mean * colsum(W')) == LN(x; gamma) . W^T."""
    ops = _ops()
    T, H, I, heads = 1000, 768, 1152, 12
    a, wo = _rand((T, H), seed=1), _rand((H, H), 0.03, seed=2)
    x0 = _rand((T, H), 2.0, seed=3) + 0.7  # non-zero mean rows
    gamma = _rand((H,), 0.2, seed=4, dtype=torch.float32) + 1.0
    stats = torch.full(((H + 255) // 256, T, 2), float("nan"), device=DEV)  # one partial per 256-column tile
    x = ops.gemm(a, wo, epilogue=ops.EPI_RESIDUAL, aux=x0, stats_out=stats)
    torch.cuda.synchronize()
    xf = x.float()
    _report("stats sum", stats[..., 0].sum(0), xf.sum(-1), 5e-2, 1e-3)
    _report("stats sumsq", stats[..., 1].sum(0), (xf * xf).sum(-1), 5e-1, 1e-3)
    ln = F.layer_norm(xf, (H,), gamma, None, 1e-5)
    # --- GeGLU consumer
    wi = _rand((2 * I, H), 0.05, seed=5)
    wi_ln = ops.interleave_wi((wi.float() * gamma[None]).to(torch.bfloat16)).contiguous()
    ci = wi_ln.float().sum(1).contiguous()
    got = ops.gemm(x, wi_ln, epilogue=ops.EPI_GEGLU, row_stats=stats, col_corr=ci, ln_eps=1e-5)
    acc = ln @ wi.float().t()
    _report("geglu(LN folded)", got, F.gelu(acc[:, :I]) * acc[:, I:], 4e-2, 3e-2)
    raw = torch.empty((T, 2 * I), device=DEV, dtype=torch.bfloat16)
    got2 = ops.gemm(x, wi_ln, epilogue=ops.EPI_GEGLU_SAVE, c2=raw, row_stats=stats, col_corr=ci, ln_eps=1e-5)
    _report("geglu_save(LN folded).raw", ops.deinterleave_wi(raw.t().contiguous()).t(), acc, 4e-2, 2e-2)
    _report("geglu_save(LN folded).out", got2, F.gelu(acc[:, :I]) * acc[:, I:], 4e-2, 3e-2)
    # --- RoPE consumer
    wq = _rand((3 * H, H), 0.05, seed=6)
    wq_ln = (wq.float() * gamma[None]).to(torch.bfloat16).contiguous()
    cq = wq_ln.float().sum(1).contiguous()
    pos = (torch.arange(T, dtype=torch.int32) % 333).to(DEV)
    tab = ops.rope_table(160000.0, 512, DEV)
    got3 = ops.gemm(x, wq_ln, epilogue=ops.EPI_ROPE, positions=pos, rope_table=tab, rope_cols=2 * H, row_stats=stats,
                    col_corr=cq, ln_eps=1e-5)
    accq = (ln @ wq.float().t()).view(T, 3, heads, 64)
    cos, sin = tab[pos.long(), :, 0], tab[pos.long(), :, 1]
    cos, sin = torch.cat((cos, cos), -1)[:, None, None], torch.cat((sin, sin), -1)[:, None, None]
    rot = torch.cat((-accq[..., 32:], accq[..., :32]), dim=-1)
    want = accq.clone()
    want[:, :2] = (accq * cos + rot * sin)[:, :2]
    _report("rope(LN folded)", got3, want.view(T, 3 * H), 5e-2, 2e-2)
