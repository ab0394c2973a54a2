"""world_size-2 `gloo` test of the data-parallel host logic (cm3p_b200/distributed.py) on CPU.

The CUDA kernels cannot run here, so the per-rank compute is the CPU oracle; what is under test is
the sharding algebra the CUDA train step uses (cm3p_b200/training.py): all-gather of the normalised
embeddings, full loss on every rank, own row block of the embedding gradient, one SUM all-reduce of
the flat parameter-gradient buffer (logit_scale pre-divided by the world size) — against the
single-process oracle on the concatenated batch.  Local-negative mode is checked
against the mean of the per-shard gradients (DDP semantics of the reference).
"""
import copy
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cm3p_b200 import distributed as D
from cm3p_b200.configuration_cm3p import CM3PConfig, small_config_dict
from cm3p_b200.synthetic import synthetic_batch, synthetic_state_dict
from oracle import cm3p_oracle as O

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup():
    cfg = CM3PConfig(**copy.deepcopy(small_config_dict()))
    sd = synthetic_state_dict(cfg, seed=3, gain=1.0, dtype=torch.float64)
    batch = synthetic_batch(cfg, batch=4, seq_len=300, variations=3, seed=2, pad_variations=1)
    batch = {k: (v.double() if v.is_floating_point() else v) for k, v in batch.items()}
    return cfg, sd, batch


def _shard(batch, rank, world):
    n = batch["input_ids"].shape[0] // world
    return {k: v[rank * n:(rank + 1) * n] for k, v in batch.items()}


def _flat(grads, names):
    return torch.cat([grads[n].reshape(-1) for n in names])


def _worker(rank, world, port, mode, out_dir):
    torch.set_num_threads(2)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        cfg, sd, batch = _setup()
        dp = D.DataParallel(group=None, world_size=world, rank=rank, global_negatives=(mode == "global"))
        shard = _shard(batch, rank, world)
        leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
        names = sorted(leaves)
        if mode == "global":
            out = O.model_forward(leaves, cfg, **shard, return_loss=False)
            be_loc = out["beatmap_embeds"]
            me_loc = out["metadata_embeds"].reshape(-1, be_loc.shape[-1])
            be_all = D.all_gather_rows(be_loc.detach(), dp).requires_grad_(True)
            me_all = D.all_gather_rows(me_loc.detach(), dp).requires_grad_(True)
            classes_all = D.all_gather_rows(shard["metadata_variation_classes"], dp)
            Bg, V = classes_all.shape
            scale_leaf = leaves["logit_scale"]
            lpm = (me_all @ be_all.t() * scale_leaf.exp()).view(Bg, V, Bg)
            loss = O.cm3p_loss(lpm, classes_all)
            loss.backward()
            scale_leaf.grad.div_(world)  # complete on every rank already
            torch.autograd.backward([be_loc, me_loc], [D.local_rows(be_all.grad, dp), D.local_rows(me_all.grad, dp)])
        else:
            out = O.model_forward(leaves, cfg, **shard)
            loss = out["loss"]
            loss.backward()
        grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
        flat = _flat(grads, names)
        D.reduce_gradients(flat, dp)
        torch.save({"loss": float(loss.detach()), "flat": flat}, os.path.join(out_dir, f"{mode}_{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["global", "local"])
def test_data_parallel_algebra_gloo(tmp_path, mode):
    port = _free_port()
    mp.spawn(_worker, args=(WORLD, port, mode, str(tmp_path)), nprocs=WORLD, join=True)
    res = [torch.load(os.path.join(tmp_path, f"{mode}_{r}.pt")) for r in range(WORLD)]
    cfg, sd, batch = _setup()
    names = sorted(sd)
    # every rank ends with the same reduced gradient
    assert torch.equal(res[0]["flat"], res[1]["flat"])
    if mode == "global":
        wout, want = O.forward_backward(sd, cfg, batch)
        want = {k: want.get(k, torch.zeros_like(v)) for k, v in sd.items()}
        assert abs(res[0]["loss"] - float(wout["loss"])) < 1e-10  # full loss on every rank == concatenated batch
        assert abs(res[1]["loss"] - float(wout["loss"])) < 1e-10
        ref = _flat(want, names)
    else:
        parts = []
        for r in range(WORLD):
            _, g = O.forward_backward(sd, cfg, _shard(batch, r, WORLD))
            parts.append(_flat({k: g.get(k, torch.zeros_like(v)) for k, v in sd.items()}, names))
        ref = sum(parts) / WORLD
    err = float((res[0]["flat"] - ref).abs().max())
    assert err <= 1e-9 * max(1.0, float(ref.abs().max())), err


def test_gather_and_slice_roundtrip_single_process():
    dp = D.DataParallel(group=None, world_size=1, rank=0, global_negatives=True)
    x = torch.arange(12.0).view(4, 3)
    assert torch.equal(D.all_gather_rows(x, dp), x)
    assert torch.equal(D.local_rows(x, dp), x)
    dp4 = D.DataParallel(group=None, world_size=4, rank=2)
    assert torch.equal(D.local_rows(x, dp4), x[2:3])


# ------------------------------------------------------------------------------------------------
# bucketed, overlapped gradient reduction (training.GradStore): layout in backward order, a bucket is sent the
# moment the backward pass touches a parameter behind it, mean for per-rank objectives / sum for the global loss

def _gradstore_worker(rank, world, port, sum_reduce, out_dir):
    from cm3p_b200 import training as T
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        params = [torch.zeros(n) for n in (1000, 7, 4096, 300, 1, 2048, 513)]
        touch = [4, 0, 2, 6, 1, 5]  # backward order; parameter 3 is never touched (frozen)
        dp = D.DataParallel(group=None, world_size=world, rank=rank, global_negatives=sum_reduce)
        T.GradStore.BUCKET_BYTES = 8192  # several buckets for these sizes

        def backward(order):
            g = T.GradStore(params, order=order, dp=dp, sum_reduce=sum_reduce)
            sent = []
            for i in touch:
                g(params[i]).add_(float(rank + 1) * (i + 1))
                sent.append(len(g._works))
            g.finish()
            return g, sent

        g1, sent1 = backward(None)           # first step: records the order, one reduction at the end
        order = g1.order()
        g2, sent2 = backward(order)          # steady state: buckets leave during the backward pass
        torch.save({"order": order, "sent1": sent1, "sent2": sent2, "buckets": len(g2._buckets),
                    "vals1": [float(g1(p).flatten()[0]) for p in params],
                    "vals2": [float(g2(p).flatten()[0]) for p in params]},
                   os.path.join(out_dir, f"gs_{int(sum_reduce)}_{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sum_reduce", [False, True])
def test_gradstore_bucketed_reduction_gloo(tmp_path, sum_reduce):
    mp.spawn(_gradstore_worker, args=(WORLD, _free_port(), sum_reduce, str(tmp_path)), nprocs=WORLD, join=True)
    res = [torch.load(os.path.join(tmp_path, f"gs_{int(sum_reduce)}_{r}.pt")) for r in range(WORLD)]
    touch = [4, 0, 2, 6, 1, 5]
    assert res[0]["order"] == touch + [3]
    assert res[0]["sent1"] == [0] * len(touch)          # nothing leaves early on the recording step
    assert res[0]["buckets"] >= 3
    assert res[0]["sent2"][-1] >= 1 and res[0]["sent2"] == sorted(res[0]["sent2"])  # buckets left during backward
    ranks_total = sum(r + 1 for r in range(WORLD))
    for i in range(7):
        want = 0.0 if i == 3 else (i + 1) * ranks_total / (1 if sum_reduce else WORLD)
        for r in range(WORLD):
            assert abs(res[r]["vals1"][i] - want) < 1e-6 and abs(res[r]["vals2"][i] - want) < 1e-6, (i, res[r])
