"""One launch of each hot kernel at its train-step shape (B windows of L=2000, base config), for
`ncu --set full -k regex:cm3p`.  No warm-up on purpose: every launch in this script is a profiling target.

    python tools/ncu_targets.py [B]          (default 64 windows = 83 k tokens; 256 = the train step)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cm3p_b200 import ops  # noqa: E402

DEV = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L, HEADS, H, I = 2000, 12, 768, 1152


def lengths(batch, lo, hi, seed=0):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(lo, hi + 1, (batch,), generator=g).tolist()
    lens[0] = hi
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    return lens, cu


def gemms(T):
    x = torch.randn(T, H, device=DEV).bfloat16()
    res = torch.randn(T, H, device=DEV).bfloat16()
    wqkv = (torch.randn(3 * H, H, device=DEV) * 0.03).bfloat16()
    wo = (torch.randn(H, H, device=DEV) * 0.03).bfloat16()
    wi = (torch.randn(2 * I, H, device=DEV) * 0.03).bfloat16()
    wo2 = (torch.randn(H, I, device=DEV) * 0.03).bfloat16()
    pos = (torch.arange(T, device=DEV, dtype=torch.int32) % L)
    tab = ops.rope_table(160000.0, 2048, DEV)
    qkv = ops.gemm(x, wqkv, epilogue=ops.EPI_ROPE, positions=pos, rope_table=tab, rope_cols=2 * H)  # Wqkv + RoPE
    ops.gemm(x, wo, epilogue=ops.EPI_RESIDUAL, aux=res)                                            # Wo + residual
    hact = ops.gemm(x, wi, epilogue=ops.EPI_GEGLU)                                                  # Wi + GeGLU
    ops.gemm(hact, wo2, epilogue=ops.EPI_RESIDUAL, aux=res)                                         # Wo2 + residual
    # backward of Wqkv: dgrad (K = 2304) and split-K weight gradient (K = tokens)
    dx = torch.empty(T, H, device=DEV, dtype=torch.bfloat16)
    ops.gemm(qkv, wqkv, trans_b=True, out=dx)
    dw = torch.zeros(3 * H, H, device=DEV)
    ops.gemm(qkv, x, trans_a=True, trans_b=True, epilogue=ops.EPI_SCALE_F32, accumulate=True, out=dw)
    dwo = torch.zeros(H, H, device=DEV)
    ops.gemm(res, x, trans_a=True, trans_b=True, epilogue=ops.EPI_SCALE_F32, accumulate=True, out=dwo)
    return qkv


def attention(batch, seq, heads, window, lo, packed=False):
    lens, cu = lengths(batch, lo, seq)
    T = cu[-1]
    qkv = torch.randn(T, 3 * heads * 64, device=DEV).bfloat16()
    dout = torch.randn(T, heads * 64, device=DEV).bfloat16()
    cu_t = torch.tensor(cu, dtype=torch.int32, device=DEV)
    lse = torch.empty(heads, T, device=DEV)
    groups = ops.attn_pack_groups(cu_t, T) if packed else None
    out = ops.attn_varlen_fwd(qkv, cu_t, seq, heads, window, lse=lse, groups=groups)
    pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).to(DEV)
    tab = ops.rope_table(160000.0, 2048, DEV)
    dqkv, delta = torch.empty_like(qkv), torch.empty_like(lse)
    ops.attn_varlen_bwd(qkv, out, dout, lse, cu_t, seq, heads, window, positions=pos, rope_table=tab, dqkv=dqkv,
                        delta=delta, groups=groups)


def rowwise(T):
    x = torch.randn(T, H, device=DEV).bfloat16()
    dy = torch.randn(T, H, device=DEV).bfloat16()
    dres = torch.randn(T, H, device=DEV).bfloat16()
    g = torch.ones(H, device=DEV)
    ops.layernorm(x, g, 1e-5)
    ops.layernorm_bwd(x, dy, g, 1e-5, dres=dres, dx=torch.empty_like(x), dgamma=torch.zeros(H, device=DEV))
    ug = torch.randn(T, 2 * I, device=DEV).bfloat16()
    dh = torch.randn(T, I, device=DEV).bfloat16()
    ops.geglu_bwd(ug, dh, dug=torch.empty_like(ug), h=torch.empty_like(dh))


if __name__ == "__main__":
    _, cu = lengths(B, 600, L)
    T = cu[-1]
    gemms(T)
    attention(B, L, HEADS, -1, 600)      # global layers of the beatmap tower
    attention(B, L, HEADS, 64, 600)      # sliding-window layers (band-walk backward)
    attention(B * 8, 25, 4, -1, 17, packed=True)      # metadata tower at V = 8: packed short-sequence kernels
    rowwise(T)
    torch.cuda.synchronize()
    print(f"ncu_targets ok: B={B} T={T}")
