"""Muon optimizer on the sm_100a kernels (reference: /root/reference/utils/muon_utils.py:60-204).

Same constructor arguments, parameter routing (`use_muon` = ndim >= 2 and size(0) < 10000, everything in
`adamw_params` goes to the internal AdamW) and state keys (`use_muon`, `momentum_buffer`, `step`,
`moment1`, `moment2`) as the reference, so optimizer checkpoints are interchangeable and
`train.py:325-352` can construct it unchanged.  What differs is the execution: the Newton-Schulz
orthogonalisation (:35-57, 3 GEMMs per iteration in bf16) runs on the tcgen05 GEMM through the C ABI
— with the tall/wide transposes folded into the TMA operand majors instead of materialised — and the
momentum / AdamW / update arithmetic runs in fused element-wise kernels (cm3p_muon_momentum,
cm3p_bf16_normalize, cm3p_bf16_axpy, cm3p_muon_apply, cm3p_adamw_step).  No DTensor handling: the
framework replicates parameters (pure data parallelism).
"""
from __future__ import annotations

from typing import Generator

import torch

from . import _lib, ops

NS_COEFFS = (3.4445, -4.7750, 2.0315)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def _axpy(out, a, x, y=None):
    rows, cols = x.shape
    rc = _lib.load().cm3p_bf16_axpy(out.data_ptr(), out.stride(0), float(a), x.data_ptr(), x.stride(0),
                                    None if y is None else y.data_ptr(), 0 if y is None else y.stride(0), rows, cols,
                                    _stream())
    _lib.check(rc, "cm3p_bf16_axpy")
    return out


def newton_schulz5(x: torch.Tensor, steps: int) -> torch.Tensor:
    """x: bf16 [R, C] (already normalised) -> orthogonalised bf16 [R, C] (a fresh or the same buffer).

    Reference iteration on X = G (wide) or G^T (tall): A = X X^T; B = b A + (c A) A; X = a X + B X.
    Here G keeps its layout: wide  (R <= C): A = G G^T, G <- a G + B G;
                             tall  (R >  C): A = G^T G, G <- a G + G B   (B symmetric).
    """
    a, b, c = NS_COEFFS
    R, C = x.shape
    tall = R > C
    n = C if tall else R
    dev = x.device
    ld = _pad8(n)
    A = torch.empty((n, ld), device=dev, dtype=torch.bfloat16)[:, :n]
    cA = torch.empty((n, ld), device=dev, dtype=torch.bfloat16)[:, :n]
    B = torch.empty((n, ld), device=dev, dtype=torch.bfloat16)[:, :n]
    ax = torch.empty((R, _pad8(C)), device=dev, dtype=torch.bfloat16)[:, :C]
    nxt = torch.empty((R, _pad8(C)), device=dev, dtype=torch.bfloat16)[:, :C]  # ping-pong: X is also a GEMM operand
    for _ in range(steps):
        if tall:
            ops.gemm(x, x, trans_a=True, trans_b=True, out=A)     # G^T G
        else:
            ops.gemm(x, x, out=A)                                 # G G^T
        _axpy(cA, c, A)                                           # bf16(c * A)
        ops.gemm(cA, A, out=B)                                    # (c A) A   (A symmetric: A . A^T == A . A)
        _axpy(B, b, A, B)                                         # bf16(bf16(b A) + (cA)A)
        _axpy(ax, a, x)                                           # bf16(a X)
        if tall:
            ops.gemm(x, B, epilogue=ops.EPI_RESIDUAL, aux=ax, out=nxt)               # G B + a G
        else:
            ops.gemm(B, x, trans_b=True, epilogue=ops.EPI_RESIDUAL, aux=ax, out=nxt)  # B G + a G
        x, nxt = nxt, x
    return x


class Muon(torch.optim.Optimizer):
    def __init__(self, muon_params, lr=0.004, momentum=0.95, nesterov=True, ns_steps=6, adamw_params=None,
                 adamw_lr=0.002, adamw_betas=(0.95, 0.95), adamw_eps=1e-8, adamw_wd=0):
        defaults = dict(lr=lr, momentum=momentum, nesterov=nesterov, ns_steps=ns_steps, adamw_lr_ratio=adamw_lr / lr,
                        adamw_betas=adamw_betas, adamw_eps=adamw_eps, adamw_wd=adamw_wd)
        if isinstance(muon_params, Generator):
            muon_params = list(muon_params)
        if isinstance(adamw_params, Generator):
            adamw_params = list(adamw_params)
        elif adamw_params is None:
            adamw_params = []
        super().__init__([*muon_params, *adamw_params], defaults)

        def each(params):
            if len(params) and isinstance(params[0], dict):
                for group in params:
                    yield from group["params"]
            else:
                yield from params

        for p in each(muon_params):
            self.state[p]["use_muon"] = 1 if (p.ndim >= 2 and p.size(0) < 10000) else 0
        for p in each(adamw_params):
            self.state[p]["use_muon"] = 0

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            lr, momentum = group["lr"], group["momentum"]
            b1, b2 = group["adamw_betas"]
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("cm3p_b200.Muon: parameters must be fp32 CUDA tensors (no CPU fallback)")
                g = g.detach().float().contiguous()
                state = self.state[p]
                if state["use_muon"] == 1:
                    R = p.shape[0]
                    C = p.numel() // R
                    if "momentum_buffer" not in state:
                        state["momentum_buffer"] = torch.zeros((R, C), device=p.device, dtype=torch.float32)
                    buf = state["momentum_buffer"]
                    xbuf = torch.empty((R, _pad8(C)), device=p.device, dtype=torch.bfloat16)
                    contiguous_x = xbuf.shape[1] == C
                    x = xbuf[:, :C]
                    sumsq = torch.zeros((1,), device=p.device, dtype=torch.float32)
                    if contiguous_x:
                        _lib.check(lib.cm3p_muon_momentum(g.data_ptr(), buf.data_ptr(), x.data_ptr(), R * C,
                                                          float(momentum), int(group["nesterov"]), sumsq.data_ptr(),
                                                          _stream()), "cm3p_muon_momentum")
                        _lib.check(lib.cm3p_bf16_normalize(x.data_ptr(), R * C, sumsq.data_ptr(), 1e-7, _stream()),
                                   "cm3p_bf16_normalize")
                    else:  # row pitch padded to 16 bytes for TMA: go through a dense staging copy
                        dense = torch.empty((R, C), device=p.device, dtype=torch.bfloat16)
                        _lib.check(lib.cm3p_muon_momentum(g.data_ptr(), buf.data_ptr(), dense.data_ptr(), R * C,
                                                          float(momentum), int(group["nesterov"]), sumsq.data_ptr(),
                                                          _stream()), "cm3p_muon_momentum")
                        _lib.check(lib.cm3p_bf16_normalize(dense.data_ptr(), R * C, sumsq.data_ptr(), 1e-7,
                                                           _stream()), "cm3p_bf16_normalize")
                        x.copy_(dense)
                    x = newton_schulz5(x, group["ns_steps"])
                    upd = x if x.is_contiguous() else x.contiguous()
                    post = max(1.0, R / C) ** 0.5
                    _lib.check(lib.cm3p_muon_apply(p.data_ptr(), upd.data_ptr(), R * C, float(post), float(-lr),
                                                   _stream()), "cm3p_muon_apply")
                    # the kernels write through raw pointers: tell autograd / the bf16 weight-pack caches
                    torch.autograd.graph.increment_version(p)
                else:
                    if "step" not in state:
                        state["step"] = 0
                        state["moment1"] = torch.zeros_like(g)
                        state["moment2"] = torch.zeros_like(g)
                    state["step"] += 1
                    step = state["step"]
                    scale = (1 - b1 ** step) / (1 - b2 ** step) ** 0.5
                    adamw_lr = lr * group["adamw_lr_ratio"]
                    _lib.check(lib.cm3p_adamw_step(p.data_ptr(), g.data_ptr(), state["moment1"].data_ptr(),
                                                   state["moment2"].data_ptr(), p.numel(), float(b1), float(b2),
                                                   float(group["adamw_eps"]), float(1 - adamw_lr * group["adamw_wd"]),
                                                   float(lr / scale), _stream()), "cm3p_adamw_step")
                    torch.autograd.graph.increment_version(p)
        return loss


def split_muon_adamw(model):
    """The reference's routing rule (train.py:331-340): names containing 'embed' / 'proj_out' or ndim <= 1 -> AdamW."""
    adamw = [p for n, p in model.named_parameters()
             if (any(kw in n.lower() for kw in ("embed", "proj_out")) or p.ndim <= 1)]
    ids = {id(p) for p in adamw}
    muon = [p for _, p in model.named_parameters() if id(p) not in ids]
    return muon, adamw
