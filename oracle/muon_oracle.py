"""CPU restatement of the reference's Muon optimizer step.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/utils/muon_utils.py: `zeropower_via_newtonschulz5` (:35-57) and `Muon.step`
(:138-203), without the DTensor branches (:19-32, :158-163) that only matter under FSDP.
Parity status: PINNED — `oracle/make_golden_muon.py` ran the unmodified reference class in this
container (it imports only torch) and committed `tests/golden/muon_steps.npz`;
`tests/test_oracle_golden.py::test_muon_oracle_matches_reference` holds this file to it bit-for-bit.
"""
from __future__ import annotations

import torch


def newton_schulz5(G: torch.Tensor, steps: int, eps: float = 1e-7) -> torch.Tensor:
    a, b, c = (3.4445, -4.7750, 2.0315)
    X = G.bfloat16()
    X = X / (X.norm() + eps)
    if G.size(0) > G.size(1):
        X = X.T
    for _ in range(steps):
        A = X @ X.T
        B = b * A + c * A @ A
        X = a * X + B @ X
    if G.size(0) > G.size(1):
        X = X.T
    return X


def muon_step(params: dict, grads: dict, state: dict, use_muon: dict, lr: float, momentum: float = 0.95,
              nesterov: bool = True, ns_steps: int = 6, adamw_lr: float | None = None, adamw_betas=(0.95, 0.95),
              adamw_eps: float = 1e-8, adamw_wd: float = 0.0) -> None:
    """In-place update of `params` (name -> fp32 tensor); `state` persists across calls."""
    adamw_lr = lr / 2 if adamw_lr is None else adamw_lr
    for name, p in params.items():
        g = grads.get(name)
        if g is None:
            continue
        st = state.setdefault(name, {})
        if use_muon[name]:
            g2 = g.view(g.size(0), -1)
            if "momentum_buffer" not in st:
                st["momentum_buffer"] = torch.zeros_like(g2)
            buf = st["momentum_buffer"]
            buf.mul_(momentum).add_(g2)
            u = g2.add(buf, alpha=momentum) if nesterov else buf
            u = newton_schulz5(u, ns_steps)
            u = u * max(1, u.size(0) / u.size(1)) ** 0.5
            p.add_(u.view_as(p).type_as(p), alpha=-lr)
        else:
            if "step" not in st:
                st["step"] = 0
                st["moment1"] = torch.zeros_like(g)
                st["moment2"] = torch.zeros_like(g)
            st["step"] += 1
            step = st["step"]
            st["moment1"].lerp_(g, 1 - adamw_betas[0])
            st["moment2"].lerp_(g.square(), 1 - adamw_betas[1])
            upd = st["moment1"] / (adamw_eps + st["moment2"].sqrt())
            scale = (1 - adamw_betas[0] ** step) / (1 - adamw_betas[1] ** step) ** 0.5
            p.mul_(1 - adamw_lr * adamw_wd)
            p.add_(upd, alpha=-lr / scale)
