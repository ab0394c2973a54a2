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


def newton_schulz5(x: torch.Tensor, steps: int, groups: int = 1) -> torch.Tensor:
    """x: bf16 [groups*R, C] (already normalised, `groups` same-shape matrices stacked along the rows) ->
    orthogonalised bf16 of the same shape (a fresh or the same buffer).

    Reference iteration on X = G (wide) or G^T (tall): A = X X^T; B = b A + (c A) A; X = a X + B X.
    Here G keeps its layout: wide  (R <= C): A = G G^T, G <- a G + B G;
                             tall  (R >  C): A = G^T G, G <- a G + G B   (B symmetric).
    With groups > 1 every product is ONE grouped tcgen05 GEMM launch over all matrices (the 22 layers' Wqkv / Wi /
    Wo / Wo2 have four shapes between them): a single 768x768 Gram matrix is 18 output tiles on 148 SMs.
    """
    a, b, c = NS_COEFFS
    RG, C = x.shape
    R = RG // groups
    tall = R > C
    n = C if tall else R
    dev = x.device
    ld = _pad8(n)
    A = torch.empty((groups * n, ld), device=dev, dtype=torch.bfloat16)[:, :n]
    cA = torch.empty((groups * n, ld), device=dev, dtype=torch.bfloat16)[:, :n]
    B = torch.empty((groups * n, ld), device=dev, dtype=torch.bfloat16)[:, :n]
    ax = torch.empty((RG, _pad8(C)), device=dev, dtype=torch.bfloat16)[:, :C]
    nxt = torch.empty((RG, _pad8(C)), device=dev, dtype=torch.bfloat16)[:, :C]  # ping-pong: X is also a GEMM operand
    for _ in range(steps):
        if tall:
            ops.gemm(x, x, trans_a=True, trans_b=True, out=A, groups=groups)     # G^T G
        else:
            ops.gemm(x, x, out=A, groups=groups)                                 # G G^T
        _axpy(cA, c, A)                                           # bf16(c * A)
        ops.gemm(cA, A, out=B, groups=groups)                     # (c A) A   (A symmetric: A . A^T == A . A)
        _axpy(B, b, A, B)                                         # bf16(bf16(b A) + (cA)A)
        _axpy(ax, a, x)                                           # bf16(a X)
        if tall:
            ops.gemm(x, B, epilogue=ops.EPI_RESIDUAL, aux=ax, out=nxt, groups=groups)               # G B + a G
        else:
            ops.gemm(B, x, trans_b=True, epilogue=ops.EPI_RESIDUAL, aux=ax, out=nxt, groups=groups)  # B G + a G
        x, nxt = nxt, x
    return x


def _groupable(R: int, C: int) -> bool:
    """Shapes the grouped GEMM takes (group rows % 256, K % 64) for all three products of an iteration."""
    n = min(R, C)
    return n % 256 == 0 and R % 64 == 0 and C % 64 == 0 and (R <= C or R % 256 == 0)


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
            muon_by_shape: dict = {}
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
                    muon_by_shape.setdefault((R, p.numel() // R, p.device), []).append((p, g))
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
            # Muon parameters, all matrices of one shape at a time: momentum + normalisation per matrix into one
            # stacked bf16 buffer, Newton-Schulz on the stack (grouped GEMMs), update per matrix
            for (R, C, dev), items in muon_by_shape.items():
                chunk = len(items) if (_groupable(R, C) and ops.GROUPED_MUON) else 1
                for i0 in range(0, len(items), chunk):
                    part = items[i0:i0 + chunk]
                    G = len(part)
                    dense = _pad8(C) == C
                    xs = torch.empty((G * R, C), device=dev, dtype=torch.bfloat16)
                    sumsq = torch.zeros((G,), device=dev, dtype=torch.float32)
                    for i, (p, g) in enumerate(part):
                        state = self.state[p]
                        if "momentum_buffer" not in state:
                            state["momentum_buffer"] = torch.zeros((R, C), device=dev, dtype=torch.float32)
                        xi = xs[i * R:(i + 1) * R]
                        _lib.check(lib.cm3p_muon_momentum(g.data_ptr(), state["momentum_buffer"].data_ptr(),
                                                          xi.data_ptr(), R * C, float(momentum),
                                                          int(group["nesterov"]), sumsq[i:i + 1].data_ptr(), _stream()),
                                   "cm3p_muon_momentum")
                        _lib.check(lib.cm3p_bf16_normalize(xi.data_ptr(), R * C, sumsq[i:i + 1].data_ptr(), 1e-7,
                                                           _stream()), "cm3p_bf16_normalize")
                    if dense:
                        x = xs
                    else:  # row pitch padded to 16 bytes for TMA (only ungrouped shapes get here)
                        x = torch.empty((G * R, _pad8(C)), device=dev, dtype=torch.bfloat16)[:, :C]
                        x.copy_(xs)
                    x = newton_schulz5(x, group["ns_steps"], groups=G)
                    upd = x if x.is_contiguous() else x.contiguous()
                    post = max(1.0, R / C) ** 0.5
                    for i, (p, _) in enumerate(part):
                        _lib.check(lib.cm3p_muon_apply(p.data_ptr(), upd[i * R:(i + 1) * R].data_ptr(), R * C,
                                                       float(post), float(-lr), _stream()), "cm3p_muon_apply")
                        # the kernels write through raw pointers: tell autograd / the bf16 weight-pack caches
                        torch.autograd.graph.increment_version(p)
        return loss


def split_muon_adamw(model):
    """The reference's routing rule (train.py:331-340): names containing 'embed' / 'proj_out' or ndim <= 1 -> AdamW."""
    adamw = [p for n, p in model.named_parameters()
             if (any(kw in n.lower() for kw in ("embed", "proj_out")) or p.ndim <= 1)]
    ids = {id(p) for p in adamw}
    muon = [p for _, p in model.named_parameters() if id(p) not in ids]
    return muon, adamw
