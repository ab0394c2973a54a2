"""Embedding extraction around the model (reference: /root/reference/extract_beatmap_embeddings.py:217-266).

The reference copies every batch of window embeddings to the host and keeps a Python dict of numpy sums;
here the per-beatmap sums live on the GPU (fp32 atomics), and the mean + re-normalisation is one kernel
at the end, so the only D2H traffic is the final (n_beatmaps, 512) table.  The parquet / metadata-merge
part of the script (:266-320) is host bookkeeping and out of scope.
"""
from __future__ import annotations

import torch

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class BeatmapEmbeddingAccumulator:
    """sum of window embeddings per beatmap id -> unit-length mean embedding per beatmap."""

    def __init__(self, proj_dim: int, device, capacity: int = 1024):
        self.P, self.device = proj_dim, torch.device(device)
        self.slots: dict[int, int] = {}
        self.sums = torch.zeros((capacity, proj_dim), device=self.device, dtype=torch.float32)
        self.counts = torch.zeros((capacity,), device=self.device, dtype=torch.float32)

    def _grow(self, need: int) -> None:
        if need <= self.sums.shape[0]:
            return
        cap = max(need, 2 * self.sums.shape[0])
        sums = torch.zeros((cap, self.P), device=self.device, dtype=torch.float32)
        counts = torch.zeros((cap,), device=self.device, dtype=torch.float32)
        sums[: self.sums.shape[0]] = self.sums
        counts[: self.counts.shape[0]] = self.counts
        self.sums, self.counts = sums, counts

    def add(self, embeds: torch.Tensor, beatmap_ids) -> None:
        """embeds: (B, P) window embeddings on the GPU; beatmap_ids: B ints (None = skip, as the reference does)."""
        if torch.is_tensor(beatmap_ids):
            beatmap_ids = beatmap_ids.tolist()
        idx = []
        for bid in beatmap_ids:
            if bid is None:
                idx.append(-1)
                continue
            s = self.slots.get(int(bid))
            if s is None:
                s = self.slots[int(bid)] = len(self.slots)
            idx.append(s)
        self._grow(len(self.slots))
        e = embeds.detach().float().contiguous()
        slot = torch.tensor(idx, dtype=torch.int32, device=self.device)
        rc = _lib.load().cm3p_segment_accumulate(e.data_ptr(), slot.data_ptr(), self.sums.data_ptr(),
                                                 self.counts.data_ptr(), e.shape[0], self.P, _stream())
        _lib.check(rc, "cm3p_segment_accumulate")

    def finalize(self):
        """-> (beatmap ids in first-seen order, (n, P) fp32 unit-length mean embeddings on the GPU)."""
        n = len(self.slots)
        out = torch.empty((n, self.P), device=self.device, dtype=torch.float32)
        rc = _lib.load().cm3p_mean_renormalize(self.sums.data_ptr(), self.counts.data_ptr(), out.data_ptr(), n, self.P,
                                               _stream())
        _lib.check(rc, "cm3p_mean_renormalize")
        return list(self.slots.keys()), out


@torch.no_grad()
def extract_beatmap_embeddings(model, batches, device=None):
    """`batches`: iterable of dicts with input_ids, attention_mask, [input_features], beatmap_id (the reference's
    DataLoader output).  -> (ids, (n_beatmaps, P) embeddings)."""
    device = device or next(model.parameters()).device
    dtype = next(model.parameters()).dtype
    acc = BeatmapEmbeddingAccumulator(model.config.projection_dim, device)
    for batch in batches:
        if len(batch.get("input_ids", [])) == 0:
            continue
        inputs = {"input_ids": batch["input_ids"].to(device), "attention_mask": batch["attention_mask"].to(device)}
        if batch.get("input_features") is not None:
            inputs["input_features"] = batch["input_features"].to(device=device, dtype=dtype)
        out = model(**inputs, return_loss=False)
        if batch.get("beatmap_id") is None:
            continue
        acc.add(out.beatmap_embeds, batch["beatmap_id"])
    return acc.finalize()
