"""2-rank NCCL test of the data-parallel train step (needs >= 2 B200s; skipped on a 1-GPU box).

global negatives: loss on every rank == 1-GPU loss on the concatenated batch, and the summed
gradients == the 1-GPU gradients (SURVEY.md §8e parity definition).  local negatives: reduced
gradient == mean of the per-shard 1-GPU gradients (the reference's DDP semantics)."""
import copy
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

from cm3p_b200.configuration_cm3p import CM3PConfig, small_config_dict
from cm3p_b200.synthetic import synthetic_batch, synthetic_state_dict

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model_and_batch(dev):
    from cm3p_b200.modeling_cm3p import CM3PModel
    cfg = CM3PConfig(**copy.deepcopy(small_config_dict()))
    model = CM3PModel(cfg)
    model.load_state_dict(synthetic_state_dict(cfg, seed=3, gain=1.0), strict=True)
    batch = synthetic_batch(cfg, batch=8, seq_len=400, variations=3, seed=2, pad_variations=1)
    return model.to(dev).train(), batch


def _grads(model):
    return {k: p.grad.detach().float().cpu() for k, p in model.named_parameters() if p.grad is not None}


def _worker(rank, world, port, mode, out_dir):
    import torch.distributed as dist
    from cm3p_b200 import distributed as D
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    try:
        model, batch = _model_and_batch(dev)
        D.enable_data_parallel(model, global_negatives=(mode == "global"))
        n = batch["input_ids"].shape[0] // world
        shard = {k: v[rank * n:(rank + 1) * n].to(dev) for k, v in batch.items()}
        out = model(**shard)
        out.loss.backward()
        torch.cuda.synchronize()
        torch.save({"loss": float(out.loss.detach()), "grads": _grads(model),
                    "lpm_shape": tuple(out.logits_per_metadata.shape)}, os.path.join(out_dir, f"{mode}_{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < WORLD, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["global", "local"])
def test_two_rank_train_step_matches_single_gpu(tmp_path, mode):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(WORLD, _free_port(), mode, str(tmp_path)), nprocs=WORLD, join=True)
    res = [torch.load(os.path.join(tmp_path, f"{mode}_{r}.pt")) for r in range(WORLD)]
    dev = torch.device("cuda", 0)
    if mode == "global":
        model, batch = _model_and_batch(dev)
        out = model(**{k: v.to(dev) for k, v in batch.items()})
        out.loss.backward()
        want, want_loss = _grads(model), float(out.loss.detach())
        assert res[0]["lpm_shape"] == (8, 3, 8)
        for r in range(WORLD):
            assert abs(res[r]["loss"] - want_loss) <= 2e-3 * abs(want_loss)
    else:
        parts = []
        for r in range(WORLD):
            model, batch = _model_and_batch(dev)
            shard = {k: v[r * 4:(r + 1) * 4].to(dev) for k, v in batch.items()}
            out = model(**shard)
            out.loss.backward()
            parts.append(_grads(model))
            assert abs(res[r]["loss"] - float(out.loss.detach())) <= 2e-3 * abs(float(out.loss.detach()))
        want = {k: (parts[0][k] + parts[1][k]) / 2 for k in parts[0]}
    for k, w in want.items():
        for r in range(WORLD):
            g = res[r]["grads"][k]
            num = float((g.double() - w.double()).norm())
            den = float(w.double().norm())
            if den > 1e-6:
                assert num <= 2e-2 * den, (mode, k, num / den)
    # both ranks hold the same reduced gradient
    for k in res[0]["grads"]:
        assert torch.equal(res[0]["grads"][k], res[1]["grads"][k]), k
