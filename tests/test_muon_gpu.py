"""cm3p_b200.Muon on a B200 against the reference optimizer's goldens (utils/muon_utils.Muon, 3 steps)
and against the CPU oracle at production-sized matrices.  Newton-Schulz runs in bf16 in both, so the
comparison is on the applied update: cosine >= 0.99 per parameter and matching magnitude."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.make_golden_muon import HYPER, SPECS, STEPS, seeded


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-300))


def test_muon_matches_reference_goldens(golden_dir):
    from cm3p_b200.muon import Muon
    gold = np.load(os.path.join(golden_dir, "muon_steps.npz"))
    params, grads = seeded()
    ps = {n: torch.nn.Parameter(v.clone().cuda()) for n, v in params.items()}
    opt = Muon(muon_params=[ps[n] for n, _, m in SPECS if m], adamw_params=[ps[n] for n, _, m in SPECS if not m],
               **HYPER)
    prev = {n: v.clone() for n, v in params.items()}
    for t in range(STEPS):
        for n, p in ps.items():
            p.grad = grads[t][n].clone().cuda()
        opt.step()
        torch.cuda.synchronize()
        for n, _, is_muon in SPECS:
            want = torch.from_numpy(gold[f"step{t}/{n}"])
            got = ps[n].detach().cpu()
            d_want, d_got = want - prev[n], got - prev[n]
            if is_muon:
                assert _cos(d_got, d_want) >= 0.99, (t, n, _cos(d_got, d_want))
                assert abs(float(d_got.norm()) - float(d_want.norm())) <= 0.05 * float(d_want.norm()), (t, n)
            else:
                torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-7)
            prev[n] = want.clone()
            ps[n].data.copy_(want.cuda())  # re-synchronise so bf16 noise does not compound across steps
    # state keys are the reference's (checkpoint interchange)
    st = opt.state[ps["layers.0.attn.Wqkv.weight"]]
    assert set(st) == {"use_muon", "momentum_buffer"}
    assert set(opt.state[ps["final_norm.weight"]]) == {"use_muon", "step", "moment1", "moment2"}


@pytest.mark.parametrize("shape", [(2304, 768), (768, 1152), (768, 768), (512, 80, 3)])
def test_newton_schulz_production_shapes(shape):
    from cm3p_b200.muon import Muon
    from oracle import muon_oracle as M
    g = torch.Generator().manual_seed(1)
    p0 = torch.randn(shape, generator=g) * 0.02
    grad = torch.randn(shape, generator=g) * 0.01
    p = torch.nn.Parameter(p0.clone().cuda())
    p.grad = grad.clone().cuda()
    Muon(muon_params=[p], lr=1e-3).step()
    torch.cuda.synchronize()
    ref = {"w": p0.clone()}
    M.muon_step(ref, {"w": grad}, {}, {"w": True}, lr=1e-3)
    d_got, d_want = p.detach().cpu() - p0, ref["w"] - p0
    assert _cos(d_got, d_want) >= 0.99
    assert abs(float(d_got.norm()) - float(d_want.norm())) <= 0.05 * float(d_want.norm())
