"""Micro-benchmarks of the individual kernels at base-config shapes (CUDA events, L2-cold inputs by
rotating through buffers larger than L2).  Prints one line per kernel; used to decide what to tune."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cm3p_b200 import ops  # noqa: E402

DEV = "cuda"


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


def bench_gemm(M, N, K, epi=0, name=""):
    a = torch.randn(M, K, device=DEV).bfloat16()
    b = (torch.randn(N, K, device=DEV) * 0.05).bfloat16()
    kw = {}
    n_out = N
    if epi == ops.EPI_RESIDUAL:
        kw["aux"] = torch.randn(M, N, device=DEV).bfloat16()
    if epi in (ops.EPI_GEGLU, ops.EPI_GEGLU_SAVE):
        n_out = N // 2
    if epi == ops.EPI_GEGLU_SAVE:
        kw["c2"] = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    if epi == ops.EPI_ROPE:
        kw["positions"] = (torch.arange(M, device=DEV, dtype=torch.int32) % 2000)
        kw["rope_table"] = ops.rope_table(160000.0, 2048, DEV)
        kw["rope_cols"] = 2 * N // 3
    out = torch.empty(M, n_out, device=DEV, dtype=torch.bfloat16)
    t = timeit(lambda: ops.gemm(a, b, epilogue=epi, out=out, **kw))
    t_ref = timeit(lambda: torch.matmul(a, b.t()))
    flops = 2.0 * M * N * K
    print(json.dumps(dict(kernel=f"gemm{name}", M=M, N=N, K=K, epi=epi, ms=round(t * 1e3, 4),
                          tflops=round(flops / t / 1e12, 1), cublas_ms=round(t_ref * 1e3, 4),
                          cublas_tflops=round(flops / t_ref / 1e12, 1))), flush=True)


def bench_attn(B, L, heads, window):
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(min(600, max(1, L - 8)), L + 1, (B,), generator=g).tolist()
    lens[0] = L
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    T = cu[-1]
    qkv = torch.randn(T, 3 * heads * 64, device=DEV).bfloat16()
    cu_t = torch.tensor(cu, dtype=torch.int32, device=DEV)
    out = torch.empty(T, heads * 64, device=DEV, dtype=torch.bfloat16)
    t = timeit(lambda: ops.attn_varlen_fwd(qkv, cu_t, L, heads, window, out=out))
    if window < 0:
        flops = sum(4.0 * n * n * 64 * heads for n in lens)
    else:
        flops = sum(4.0 * n * min(n, 2 * window + 1) * 64 * heads for n in lens)
    print(json.dumps(dict(kernel="attn_fwd", B=B, L=L, T=T, heads=heads, window=window, ms=round(t * 1e3, 4),
                          tflops=round(flops / t / 1e12, 1))), flush=True)
    try:
        from flash_attn import flash_attn_varlen_qkvpacked_func
        q4 = qkv.view(T, 3, heads, 64)
        ws = (window, window) if window >= 0 else (-1, -1)
        t2 = timeit(lambda: flash_attn_varlen_qkvpacked_func(q4, cu_t, L, window_size=ws))
        print(json.dumps(dict(kernel="flash_attn2_ref", window=window, ms=round(t2 * 1e3, 4),
                              tflops=round(flops / t2 / 1e12, 1))), flush=True)
    except Exception as ex:  # library comparison only
        print(json.dumps(dict(kernel="flash_attn2_ref", error=str(ex)[:200])), flush=True)


def bench_attn_bwd(B, L, heads, window):
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(600, L + 1, (B,), generator=g).tolist()
    lens[0] = L
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    T = cu[-1]
    qkv = torch.randn(T, 3 * heads * 64, device=DEV).bfloat16()
    dout = torch.randn(T, heads * 64, device=DEV).bfloat16()
    cu_t = torch.tensor(cu, dtype=torch.int32, device=DEV)
    lse = torch.empty(heads, T, device=DEV)
    out = ops.attn_varlen_fwd(qkv, cu_t, L, heads, window, lse=lse)
    dqkv, delta = torch.empty_like(qkv), torch.empty_like(lse)
    # positions + RoPE table as in the train step (the dQ / dK epilogues undo the forward's rotation)
    pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).to(DEV)
    tab = ops.rope_table(160000.0, 2048, DEV)
    t = timeit(lambda: ops.attn_varlen_bwd(qkv, out, dout, lse, cu_t, L, heads, window, positions=pos, rope_table=tab,
                                           dqkv=dqkv, delta=delta))
    if window < 0:
        flops = sum(10.0 * n * n * 64 * heads for n in lens)
    else:
        flops = sum(10.0 * n * min(n, 2 * window + 1) * 64 * heads for n in lens)
    print(json.dumps(dict(kernel="attn_bwd", B=B, L=L, T=T, heads=heads, window=window, ms=round(t * 1e3, 4),
                          tflops_5gemm=round(flops / t / 1e12, 1))), flush=True)
    try:
        from flash_attn import flash_attn_varlen_qkvpacked_func
        q4 = qkv.view(T, 3, heads, 64).clone().requires_grad_(True)
        ws = (window, window) if window >= 0 else (-1, -1)
        o = flash_attn_varlen_qkvpacked_func(q4, cu_t, L, window_size=ws)
        do4 = dout.view(T, heads, 64)
        t2 = timeit(lambda: torch.autograd.grad(o, q4, do4, retain_graph=True))
        print(json.dumps(dict(kernel="flash_attn2_bwd_ref", window=window, ms=round(t2 * 1e3, 4),
                              tflops_5gemm=round(flops / t2 / 1e12, 1))), flush=True)
    except Exception as ex:  # library comparison only
        print(json.dumps(dict(kernel="flash_attn2_bwd_ref", error=str(ex)[:200])), flush=True)


def bench_ln(T, H):
    x = torch.randn(T, H, device=DEV).bfloat16()
    g = torch.ones(H, device=DEV)
    y = torch.empty_like(x)
    t = timeit(lambda: ops.layernorm(x, g, 1e-5, out=y))
    print(json.dumps(dict(kernel="layernorm", T=T, H=H, ms=round(t * 1e3, 4),
                          gbs=round(2.0 * T * H * 2 / t / 1e9, 1))), flush=True)


def bench_gemm_bwd(T, N, K, name=""):
    """The two backward GEMMs of a Linear [N, K] over T tokens: dgrad dx = dy.W and wgrad dW += dy^T.x (split-K)."""
    dy = torch.randn(T, N, device=DEV).bfloat16()
    x = torch.randn(T, K, device=DEV).bfloat16()
    w = (torch.randn(N, K, device=DEV) * 0.05).bfloat16()
    dx = torch.empty(T, K, device=DEV, dtype=torch.bfloat16)
    dw = torch.zeros(N, K, device=DEV)
    flops = 2.0 * T * N * K
    t = timeit(lambda: ops.gemm(dy, w, trans_b=True, out=dx))
    t_ref = timeit(lambda: torch.matmul(dy, w))
    print(json.dumps(dict(kernel=f"dgrad{name}", T=T, N=N, K=K, ms=round(t * 1e3, 4), tflops=round(flops / t / 1e12, 1),
                          cublas_tflops=round(flops / t_ref / 1e12, 1))), flush=True)
    t = timeit(lambda: ops.gemm(dy, x, trans_a=True, trans_b=True, epilogue=ops.EPI_SCALE_F32, accumulate=True, out=dw))
    t_ref = timeit(lambda: torch.matmul(dy.t(), x))
    print(json.dumps(dict(kernel=f"wgrad{name}", T=T, N=N, K=K, ms=round(t * 1e3, 4), tflops=round(flops / t / 1e12, 1),
                          cublas_tflops=round(flops / t_ref / 1e12, 1))), flush=True)


def bench_ln_bwd(T, H):
    x = torch.randn(T, H, device=DEV).bfloat16()
    dy = torch.randn(T, H, device=DEV).bfloat16()
    dres = torch.randn(T, H, device=DEV).bfloat16()
    g = torch.ones(H, device=DEV)
    dx = torch.empty_like(x)
    dgamma = torch.zeros(H, device=DEV)
    t = timeit(lambda: ops.layernorm_bwd(x, dy, g, 1e-5, dres=dres, dx=dx, dgamma=dgamma))
    print(json.dumps(dict(kernel="layernorm_bwd", T=T, H=H, ms=round(t * 1e3, 4),
                          gbs=round(4.0 * T * H * 2 / t / 1e9, 1))), flush=True)


def bench_geglu_bwd(T, I):
    ug = torch.randn(T, 2 * I, device=DEV).bfloat16()
    dh = torch.randn(T, I, device=DEV).bfloat16()
    dug, h = torch.empty_like(ug), torch.empty_like(dh)
    t = timeit(lambda: ops.geglu_bwd(ug, dh, dug=dug, h=h))
    print(json.dumps(dict(kernel="geglu_bwd", T=T, I=I, ms=round(t * 1e3, 4),
                          gbs=round(6.0 * T * I * 2 / t / 1e9, 1))), flush=True)


if __name__ == "__main__":
    for kv in [a[6:] for a in sys.argv[1:] if a.startswith("--opt=")]:   # library option KEY=VALUE (A/B runs)
        ops.set_option(int(kv.split("=")[0]), int(kv.split("=")[1]))
    if "winbwd" in sys.argv[1:]:
        B = int(os.environ.get("CM3P_BENCH_B", "256"))
        bench_attn_bwd(B, 2000, 12, 64)
        bench_attn_bwd(B, 800, 8, 64)
        sys.exit(0)
    if "gemmbwd" in sys.argv[1:]:
        T = 343608
        bench_gemm_bwd(T, 2304, 768, "_wqkv")
        bench_gemm_bwd(T, 768, 768, "_wo")
        bench_gemm_bwd(T, 2304, 768, "_wi")
        bench_gemm_bwd(T, 768, 1152, "_wo2")
        sys.exit(0)
    if "rowwise" in sys.argv[1:]:
        bench_geglu_bwd(343608, 1152)
        bench_ln(343608, 768)
        bench_ln_bwd(343608, 768)
        sys.exit(0)
    if "attn" in sys.argv[1:]:
        B = int(os.environ.get("CM3P_BENCH_B", "64"))  # 256 = the train step's windows per GPU
        if B != 64:
            bench_attn(B, 2000, 12, -1)
            bench_attn(B, 2000, 12, 64)
            bench_attn(B, 800, 8, -1)
            bench_attn(B, 800, 8, 64)
            if "bwd" in sys.argv[1:]:
                bench_attn_bwd(B, 2000, 12, -1)
                bench_attn_bwd(B, 2000, 12, 64)
                bench_attn_bwd(B, 800, 8, -1)
                bench_attn_bwd(B, 800, 8, 64)
            sys.exit(0)
        bench_attn(64, 2000, 12, -1)
        bench_attn(64, 2000, 12, 64)
        bench_attn(64, 800, 8, -1)
        bench_attn(512, 25, 4, -1)
        if "bwd" in sys.argv[1:]:
            bench_attn_bwd(64, 2000, 12, -1)
            bench_attn_bwd(64, 2000, 12, 64)
        sys.exit(0)
    T = 64 * 1300
    for (N, K, epi, name) in [(2304, 768, ops.EPI_ROPE, "_wqkv_rope"), (2304, 768, 0, "_wqkv_plain"),
                              (768, 768, ops.EPI_RESIDUAL, "_wo_res"), (2304, 768, ops.EPI_GEGLU, "_wi_geglu"),
                              (2304, 768, ops.EPI_GEGLU_SAVE, "_wi_geglu_save"),
                              (768, 1152, ops.EPI_RESIDUAL, "_wo2_res")]:
        bench_gemm(T, N, K, epi, name)
    bench_gemm(8192, 8192, 8192, 0, "_square")
    bench_gemm(64 * 800, 1536, 512, ops.EPI_ROPE, "_audio_wqkv")
    bench_attn(64, 2000, 12, -1)
    bench_attn(64, 2000, 12, 64)
    bench_attn(64, 800, 8, -1)
    bench_ln(T, 768)
