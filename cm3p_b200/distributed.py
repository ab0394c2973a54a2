"""Data-parallel plumbing of the train step: one process per GPU, NCCL over NVLink / NVSwitch.

The reference never touches `torch.distributed`; under `torchrun` it trains with whatever
`transformers.Trainer` + `accelerate` set up, i.e. DDP with *local* negatives: every rank computes the
contrastive loss over its own batch and gradients are mean-all-reduced (SURVEY.md §2a, §8e;
/root/reference/train.py:360-375).  Two modes are provided here:

* local negatives (reference semantics): no data-path collective; the fp32 gradients of all
  parameters live in ONE flat buffer (training.GradStore) laid out in backward order and all-reduced
  (mean) in ~64 MB buckets, each enqueued on NCCL's stream as soon as the backward pass has moved past
  it, so the transfer overlaps the rest of the backward pass.
* global negatives (BASELINE.json configs[3]): the L2-normalised embeddings (bf16, B x 512 per
  tower and rank — a few MB in total) are all-gathered, every rank evaluates the full
  (B_global*V) x B_global logits and loss redundantly on the tensor cores (17 GFLOP at 4096 x 4096 —
  microseconds), so the loss equals the reference's single-process loss on the concatenated batch
  and the gradient w.r.t. a rank's own embeddings is simply its row block of the full embedding
  gradient: the reduce-scatter of SURVEY.md §8e degenerates to a slice, with no second collective.
  Parameter gradients are then SUMMED over ranks.

Everything here works on any backend (`nccl` on the GPUs, `gloo` in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist


@dataclass
class DataParallel:
    group: Optional[object]
    world_size: int
    rank: int
    global_negatives: bool = False


def enable_data_parallel(model, group=None, global_negatives: bool = False) -> DataParallel:
    """Attach data-parallel gradient synchronisation to a CM3PModel (call once, after init_process_group)."""
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("enable_data_parallel: torch.distributed is not initialised")
    dp = DataParallel(group=group, world_size=dist.get_world_size(group), rank=dist.get_rank(group),
                      global_negatives=global_negatives)
    model._dp = dp
    return dp


def disable_data_parallel(model) -> None:
    model._dp = None


def all_gather_rows(x: torch.Tensor, dp: DataParallel) -> torch.Tensor:
    """[rows, ...] per rank -> [world*rows, ...] in rank order (every rank must pass the same shape)."""
    x = x.contiguous()
    if dp.world_size == 1:
        return x
    parts = [torch.empty_like(x) for _ in range(dp.world_size)]
    dist.all_gather(parts, x, group=dp.group)
    return torch.cat(parts, dim=0)


def local_rows(full: torch.Tensor, dp: DataParallel) -> torch.Tensor:
    """The row block of a gathered tensor that belongs to this rank."""
    n = full.shape[0] // dp.world_size
    return full[dp.rank * n:(dp.rank + 1) * n]


def reduce_gradients(flat: torch.Tensor, dp: DataParallel) -> torch.Tensor:
    """One collective for all parameter gradients: SUM over ranks; mean for local negatives (DDP
    semantics), plain sum for global negatives (every rank already holds d(global loss))."""
    if dp.world_size > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=dp.group)
        if not dp.global_negatives:
            flat.div_(dp.world_size)
    return flat


class _Done:
    def wait(self):
        return None


def reduce_gradients_async(flat: torch.Tensor, dp: DataParallel, sum_reduce: bool):
    """Enqueue the all-reduce of one gradient bucket and return a handle with `.wait()`.  NCCL: the collective
    runs on the process group's own stream, ordered after everything already enqueued on the current stream, and
    overlaps with what the caller enqueues next; the mean is taken by the collective itself (ReduceOp.AVG)."""
    if dp.world_size == 1:
        return _Done()
    backend = dist.get_backend(dp.group)
    if sum_reduce:
        return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=dp.group, async_op=True)
    if backend == "nccl":
        return dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=dp.group, async_op=True)
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=dp.group, async_op=True)  # gloo has no AVG

    class _Mean:
        def wait(self_inner):
            work.wait()
            flat.div_(dp.world_size)

    return _Mean()


def broadcast_parameters(model, dp: DataParallel, src: int = 0) -> None:
    """Make every rank start from rank `src`'s weights (what DDP does at construction)."""
    if dp.world_size == 1:
        return
    for p in model.parameters():
        dist.broadcast(p.data, src=src, group=dp.group)
