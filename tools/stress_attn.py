"""Race hunt for the streaming attention kernels: random ragged batches, forward + backward, and the result
must be bit-identical whatever the number of blocks / outer tiles one CTA streams (each 128- / 256-row tile is
computed independently of the split) and from run to run.

    python tools/stress_attn.py [iterations]
"""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cm3p_b200 import ops  # noqa: E402

DEV = "cuda"


def run(qkv, dout, cu_t, L, heads, window, pos, tab, split):
    if split is None:
        os.environ.pop("CM3P_FWD_BLOCKS_PER_CTA", None)
        os.environ.pop("CM3P_BWD_OUTER_PER_CTA", None)
    else:
        os.environ["CM3P_FWD_BLOCKS_PER_CTA"] = str(split)
        os.environ["CM3P_BWD_OUTER_PER_CTA"] = str(split)
    T = qkv.shape[0]
    lse = torch.empty((heads, T), device=DEV, dtype=torch.float32)
    out = ops.attn_varlen_fwd(qkv, cu_t, L, heads, window, lse=lse)
    dqkv = ops.attn_varlen_bwd(qkv, out, dout, lse, cu_t, L, heads, window, positions=pos, rope_table=tab)
    torch.cuda.synchronize()
    return out, lse, dqkv


def main(iters):
    rng = random.Random(0)
    tab = ops.rope_table(160000.0, 2048, DEV)
    bad = 0
    for it in range(iters):
        B = rng.choice([1, 2, 3, 7, 16, 40])
        top = rng.choice([130, 300, 700, 1300, 2000])
        lens = [rng.randint(1, top) for _ in range(B)]
        lens[rng.randrange(B)] = top
        heads = rng.choice([1, 2, 4, 8, 12])
        window = rng.choice([-1, -1, 64, 64, 8])
        cu = [0]
        for n in lens:
            cu.append(cu[-1] + n)
        T = cu[-1]
        g = torch.Generator(device=DEV).manual_seed(it)
        qkv = torch.randn((T, 3 * heads * 64), device=DEV, generator=g).bfloat16()
        dout = torch.randn((T, heads * 64), device=DEV, generator=g).bfloat16()
        cu_t = torch.tensor(cu, dtype=torch.int32, device=DEV)
        pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).to(DEV)
        ref = run(qkv, dout, cu_t, max(lens), heads, window, pos, tab, 1)
        for split in (None, rng.choice([2, 3, 5, 16]), None):
            got = run(qkv, dout, cu_t, max(lens), heads, window, pos, tab, split)
            for name, a, b in zip(("out", "lse", "dqkv"), ref, got):
                if not torch.isfinite(b.float()).all() or not torch.equal(a, b):
                    bad += 1
                    diff = (a.float() - b.float()).abs().max().item()
                    print(f"MISMATCH it={it} {name} split={split} B={B} top={top} heads={heads} window={window} "
                          f"max abs diff {diff:.4g}", flush=True)
    print(f"stress_attn: {iters} random problems x 3 splits, {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 100))
