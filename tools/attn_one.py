"""Single attention forward (+ backward with "bwd") launch at the bench shape; used with profiling builds
(CM3P_LIB_PATH=variants/libprof.so, built with CM3P_NVCC_EXTRA=-DCM3P_ATTN_PROF)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cm3p_b200 import ops  # noqa: E402

args = [a for a in sys.argv[1:] if a != "bwd"]
window = int(args[0]) if args else -1
B, L, heads = 64, 2000, 12
g = torch.Generator().manual_seed(0)
lens = torch.randint(600, L + 1, (B,), generator=g).tolist()
lens[0] = L
cu = [0]
for n in lens:
    cu.append(cu[-1] + n)
T = cu[-1]
qkv = torch.randn(T, 3 * heads * 64, device="cuda").bfloat16()
cu_t = torch.tensor(cu, dtype=torch.int32, device="cuda")
out = torch.empty(T, heads * 64, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(heads, T, device="cuda")
for _ in range(2):
    ops.attn_varlen_fwd(qkv, cu_t, L, heads, window, out=out, lse=lse)
    torch.cuda.synchronize()
if "bwd" in sys.argv[1:]:
    dout = torch.randn(T, heads * 64, device="cuda").bfloat16()
    dqkv, delta = torch.empty_like(qkv), torch.empty_like(lse)
    pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).cuda()
    tab = ops.rope_table(160000.0, 2048, "cuda")
    ops.attn_varlen_bwd(qkv, out, dout, lse, cu_t, L, heads, window, positions=pos, rope_table=tab, dqkv=dqkv,
                        delta=delta)
    torch.cuda.synchronize()
