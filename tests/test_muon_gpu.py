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


def test_optimizer_step_invalidates_weight_packs():
    """The kernels update parameters through raw pointers; the model's bf16 weight packs must notice
    (regression test: a stale pack would make training a silent no-op)."""
    import copy
    from cm3p_b200.configuration_cm3p import CM3PConfig, small_config_dict
    from cm3p_b200.modeling_cm3p import CM3PModel
    from cm3p_b200.muon import Muon, split_muon_adamw
    from cm3p_b200.synthetic import synthetic_batch, synthetic_state_dict
    cfg = CM3PConfig(**copy.deepcopy(small_config_dict()))
    model = CM3PModel(cfg)
    model.load_state_dict(synthetic_state_dict(cfg, seed=0), strict=True)
    model = model.cuda().train()
    muon, adamw = split_muon_adamw(model)
    opt = Muon(muon_params=muon, adamw_params=adamw, lr=5e-3, adamw_lr=1e-3)
    batch = {k: v.cuda() for k, v in synthetic_batch(cfg, batch=4, seq_len=320, variations=2, seed=1).items()}
    losses = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        out = model(**batch)
        out.loss.backward()
        opt.step()
        losses.append(float(out.loss.detach()))
    assert losses[-1] < losses[0] - 0.05, losses  # far beyond atomics noise: the updated weights are being used
    v0 = [p._version for p in model.parameters()]
    opt.zero_grad(set_to_none=True)
    model(**batch).loss.backward()
    opt.step()
    assert all(b > a for a, b in zip(v0, [p._version for p in model.parameters()]))


def test_grouped_newton_schulz_equals_one_matrix_at_a_time():
    """All same-shape matrices are orthogonalised with grouped GEMM launches; every group member must come out
    bit-identical to the matrix-at-a-time path (same tiles, same K order)."""
    from cm3p_b200 import ops
    from cm3p_b200.muon import Muon
    g = torch.Generator().manual_seed(5)
    shapes = [(512, 256)] * 4 + [(256, 768)] * 3 + [(256, 256)] * 2 + [(512, 80, 3)]
    p0 = [torch.randn(s, generator=g) * 0.02 for s in shapes]
    grads = [torch.randn(s, generator=g) * 0.01 for s in shapes]

    def run(grouped):
        ops.GROUPED_MUON = grouped
        try:
            ps = [torch.nn.Parameter(v.clone().cuda()) for v in p0]
            opt = Muon(muon_params=ps, lr=1e-3)
            for _ in range(2):
                for p, gr in zip(ps, grads):
                    p.grad = gr.clone().cuda()
                opt.step()
            torch.cuda.synchronize()
            return [p.detach().clone() for p in ps]
        finally:
            ops.GROUPED_MUON = True

    a, b = run(True), run(False)
    for i, (x, y) in enumerate(zip(a, b)):
        assert torch.equal(x, y), (i, shapes[i], float((x - y).abs().max()))
        assert float((x.cpu() - p0[i]).abs().max()) > 0
