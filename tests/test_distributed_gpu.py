"""Multi-rank NCCL test of the data-parallel train step (needs >= 2 B200s; skipped on a 1-GPU box; uses every
visible GPU up to 8, so the same file is the 2-rank and the 8-rank parity test).

global negatives: loss on every rank == 1-GPU loss on the concatenated batch, and the summed
gradients == the 1-GPU gradients (SURVEY.md §8e parity definition).  local negatives: reduced
gradient == mean of the per-shard 1-GPU gradients (the reference's DDP semantics).  global_mlm: global negatives
plus the per-rank 0.5 * MLM auxiliary loss, whose gradient must come out as the MEAN over ranks (it is pre-divided
by the world size because the gradients of that mode are summed).  The bucketed, overlapped reduction is what runs:
the second step of every worker uses the recorded backward order."""
import copy
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

from cm3p_b200.configuration_cm3p import CM3PConfig, small_config_dict
from cm3p_b200.synthetic import synthetic_batch, synthetic_state_dict

WORLD = max(2, min(8, torch.cuda.device_count())) if torch.cuda.is_available() else 2
PER_RANK = 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model_and_batch(dev, mlm=False):
    from cm3p_b200.modeling_cm3p import CM3PModel
    d = small_config_dict()
    if mlm:
        d["has_decoder_head"] = True
    cfg = CM3PConfig(**copy.deepcopy(d))
    model = CM3PModel(cfg)
    model.load_state_dict(synthetic_state_dict(cfg, seed=3, gain=1.0), strict=True)
    batch = synthetic_batch(cfg, batch=PER_RANK * WORLD, seq_len=400, variations=3, seed=2, pad_variations=1,
                            with_labels=mlm)
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
        from cm3p_b200 import training
        training.GradStore.BUCKET_BYTES = 1 << 18  # several buckets even for the small test model
        model, batch = _model_and_batch(dev, mlm=(mode == "global_mlm"))
        D.enable_data_parallel(model, global_negatives=(mode != "local"))
        n = batch["input_ids"].shape[0] // world
        shard = {k: v[rank * n:(rank + 1) * n].to(dev) for k, v in batch.items()}
        for _ in range(2):  # step 1 records the backward order, step 2 reduces bucket by bucket during the backward
            model.zero_grad(set_to_none=True)
            out = model(**shard)
            out.loss.backward()
        torch.cuda.synchronize()
        torch.save({"loss": float(out.loss.detach()), "grads": _grads(model),
                    "lpm_shape": tuple(out.logits_per_metadata.shape)}, os.path.join(out_dir, f"{mode}_{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _single_gpu_grads(dev, batch_slice, mlm, with_labels=True):
    model, batch = _model_and_batch(dev, mlm=mlm)
    feed = {k: v[batch_slice].to(dev) for k, v in batch.items() if with_labels or k != "labels"}
    out = model(**feed)
    out.loss.backward()
    return _grads(model), float(out.loss.detach())


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_global_negatives_with_mlm_auxiliary_loss(tmp_path):
    """global negatives + has_decoder_head + labels: d/dtheta [global contrastive + mean_r 0.5 MLM_r]."""
    import torch.multiprocessing as mp
    mode = "global_mlm"
    mp.spawn(_worker, args=(WORLD, _free_port(), mode, str(tmp_path)), nprocs=WORLD, join=True)
    res = [torch.load(os.path.join(tmp_path, f"{mode}_{r}.pt")) for r in range(WORLD)]
    dev = torch.device("cuda", 0)
    full, _ = _single_gpu_grads(dev, slice(None), True, with_labels=False)       # global contrastive term
    want = {k: v.clone() for k, v in full.items()}
    for r in range(WORLD):
        sl = slice(r * PER_RANK, (r + 1) * PER_RANK)
        with_l, _ = _single_gpu_grads(dev, sl, True, with_labels=True)
        without, _ = _single_gpu_grads(dev, sl, True, with_labels=False)
        for k in with_l:                                                             # 0.5 * MLM_r, averaged over ranks
            base = without.get(k, torch.zeros_like(with_l[k]))
            want[k] = want.get(k, torch.zeros_like(with_l[k])) + (with_l[k] - base) / WORLD
    for k, w in want.items():
        den = float(w.double().norm())
        if den <= 1e-6:
            continue
        for r in range(WORLD):
            num = float((res[r]["grads"][k].double() - w.double()).norm())
            assert num <= 3e-2 * den, (k, num / den)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
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
        assert res[0]["lpm_shape"] == (PER_RANK * WORLD, 3, PER_RANK * WORLD)
        for r in range(WORLD):
            assert abs(res[r]["loss"] - want_loss) <= 2e-3 * abs(want_loss)
    else:
        parts = []
        for r in range(WORLD):
            model, batch = _model_and_batch(dev)
            shard = {k: v[r * PER_RANK:(r + 1) * PER_RANK].to(dev) for k, v in batch.items()}
            out = model(**shard)
            out.loss.backward()
            parts.append(_grads(model))
            assert abs(res[r]["loss"] - float(out.loss.detach())) <= 2e-3 * abs(float(out.loss.detach()))
        want = {k: sum(p[k] for p in parts) / WORLD for k in parts[0]}
    for k, w in want.items():
        for r in range(WORLD):
            g = res[r]["grads"][k]
            num = float((g.double() - w.double()).norm())
            den = float(w.double().norm())
            if den > 1e-6:
                assert num <= 2e-2 * den, (mode, k, num / den)
    # every rank holds the same reduced gradient
    for k in res[0]["grads"]:
        for r in range(1, WORLD):
            assert torch.equal(res[0]["grads"][k], res[r]["grads"][k]), k
