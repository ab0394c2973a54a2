"""Small invocations of every tcgen05 / TMA kernel family, for `compute-sanitizer --tool racecheck|synccheck|memcheck`
(one tool per gpurun call: tools/jobs/sanitize_*.sh).  Sizes are tiny on purpose: the tools slow kernels down 10-100x.
Every case also checks its result against a plain fp32 PyTorch reference, so a tool-induced timing change that
exposes a race shows up as a wrong answer even when the tool itself stays silent."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cm3p_b200 import ops  # noqa: E402

DEV = "cuda"


def rnd(shape, scale=1.0, seed=0, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).to(DEV)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def attn_ref(qkv, cu, heads, window, dout=None):
    T = qkv.shape[0]
    x = qkv.float().clone().requires_grad_(dout is not None)
    q3 = x.view(T, 3, heads, 64)
    outs = []
    for b in range(len(cu) - 1):
        s, e = cu[b], cu[b + 1]
        q, k, v = (q3[s:e, i].transpose(0, 1) for i in range(3))
        sc = q @ k.transpose(1, 2) / 8.0
        if window >= 0:
            idx = torch.arange(e - s, device=DEV)
            sc = sc.masked_fill((idx[:, None] - idx[None, :]).abs() > window, float("-inf"))
        outs.append((sc.softmax(-1) @ v).transpose(0, 1).reshape(e - s, heads * 64))
    out = torch.cat(outs)
    if dout is None:
        return out.detach(), None
    out.backward(dout.float())
    return out.detach(), x.grad


def case_gemm():
    a, b = rnd((300, 192), seed=1), rnd((320, 192), 0.05, seed=2)
    want = a.float() @ b.float().t()
    assert rel(ops.gemm(a, b), want) < 2e-2
    r = rnd((300, 320), seed=3)
    assert rel(ops.gemm(a, b, epilogue=ops.EPI_RESIDUAL, aux=r), want + r.float()) < 2e-2
    dy, x = rnd((2000, 128), 0.5, seed=4), rnd((2000, 192), 0.5, seed=5)
    out = torch.zeros((128, 192), device=DEV)
    ops.gemm(dy, x, trans_a=True, trans_b=True, epilogue=ops.EPI_SCALE_F32, accumulate=True, out=out)
    assert rel(out, dy.float().t() @ x.float()) < 5e-3
    ops.set_option(ops.OPT_WGRAD_DETERMINISTIC, 1)
    out2 = torch.zeros((128, 192), device=DEV)
    ops.gemm(dy, x, trans_a=True, trans_b=True, epilogue=ops.EPI_SCALE_F32, accumulate=True, out=out2)
    ops.set_option(ops.OPT_WGRAD_DETERMINISTIC, 0)
    assert rel(out2, dy.float().t() @ x.float()) < 5e-3
    xg = rnd((3 * 256, 128), 0.3, seed=6)
    got = ops.gemm(xg, xg, groups=3)
    want = torch.cat([xg[g * 256:(g + 1) * 256].float() @ xg[g * 256:(g + 1) * 256].float().t() for g in range(3)])
    assert rel(got, want) < 2e-2


def case_attention(window):
    lens, heads = [300, 129, 1], 2
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    T = cu[-1]
    qkv, dout = rnd((T, 3 * heads * 64), seed=7), rnd((T, heads * 64), seed=8)
    cu_t = torch.tensor(cu, dtype=torch.int32, device=DEV)
    lse = torch.empty((heads, T), device=DEV)
    out = ops.attn_varlen_fwd(qkv, cu_t, max(lens), heads, window, lse=lse)
    dqkv = ops.attn_varlen_bwd(qkv, out, dout, lse, cu_t, max(lens), heads, window)
    torch.cuda.synchronize()
    want_out, want_d = attn_ref(qkv, cu, heads, window, dout)
    assert rel(out, want_out) < 2e-2, rel(out, want_out)
    assert rel(dqkv, want_d) < 3e-2, rel(dqkv, want_d)


def case_packed():
    lens, heads = [17, 25, 1, 128, 21, 19, 64], 2
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    T = cu[-1]
    qkv, dout = rnd((T, 3 * heads * 64), seed=9), rnd((T, heads * 64), seed=10)
    cu_t = torch.tensor(cu, dtype=torch.int32, device=DEV)
    groups = ops.attn_pack_groups(cu_t, T)
    lse = torch.empty((heads, T), device=DEV)
    out = ops.attn_varlen_fwd(qkv, cu_t, max(lens), heads, -1, lse=lse, groups=groups)
    dqkv = ops.attn_varlen_bwd(qkv, out, dout, lse, cu_t, max(lens), heads, -1, groups=groups)
    torch.cuda.synchronize()
    want_out, want_d = attn_ref(qkv, cu, heads, -1, dout)
    assert rel(out, want_out) < 2e-2 and rel(dqkv, want_d) < 3e-2


def case_conv():
    import torch.nn.functional as F
    B, C, Fr, Co = 2, 80, 160, 64
    x = rnd((B, C, Fr), seed=11, dtype=torch.float32)
    w, b = rnd((Co, C, 3), 0.1, seed=12), rnd((Co,), 0.1, seed=13, dtype=torch.float32)
    xt = ops.transpose_cast(x)
    y = ops.conv1d_k3(xt, ops.pack_conv_weight(w), b, stride=1, gelu=True)
    want = F.gelu(F.conv1d(x.bfloat16().float(), w.float(), b, padding=1)).permute(0, 2, 1)
    assert rel(y, want) < 2e-2
    dz = rnd((B, Fr, Co), seed=14)
    dw = torch.zeros((Co, 3 * 128), device=DEV)
    ops.conv1d_k3_wgrad(dz, xt, dw, stride=1)
    wr = w.float().requires_grad_(True)
    F.conv1d(x.bfloat16().float(), wr, None, padding=1).backward(dz.float().permute(0, 2, 1))
    assert rel(ops.unpack_conv_weight_grad(dw, C), wr.grad) < 2e-2


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "attn", "packed", "conv"]
    if "gemm" in which:
        case_gemm()
    if "attn" in which:
        case_attention(-1)
        case_attention(64)
    if "packed" in which:
        case_packed()
    if "conv" in which:
        case_conv()
    torch.cuda.synchronize()
    print("sanitize_cases ok:", " ".join(which))
