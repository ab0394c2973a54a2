"""train.py — drop-in for the reference's training entry (`/root/reference/train.py:164-397`).

    python train.py -cn synthetic [-cd configs/train] key=value ...
    torchrun --nproc-per-node 8 train.py -cn synthetic training.per_device_train_batch_size=256

Same command line and config tree as the reference (`python train.py -cn v7 training.max_steps=...`):
Hydra-style composition of `configs/train/*.yaml` + `configs/model/*.yaml` (cm3p_b200/hydra_lite.py;
hydra-core / omegaconf / accelerate are not needed), the same top-level keys (`model_cls`,
`attn_implementation`, `from_pretrained`, `freeze_*`, `unfreeze_beatmap_model_at_step`, `training.*`,
`dataset.*`, `model.*`), the reference's Muon-vs-AdamW parameter split (train.py:325-352) and HF-style
checkpoints (`output_dir/checkpoint-N/` with `save_pretrained` safetensors + optimizer state,
auto-resume from the last checkpoint unless `overwrite_output_dir`).

What runs underneath is this framework: `CM3PModel` / `CM3PForMaskedLM` / `CM3PForBeatmapClassification`
on the sm_100a kernels, the explicit CUDA backward, `cm3p_b200.Muon`, and — under `torchrun` — one
process per GPU with a single NCCL all-reduce of the flat gradient buffer per step (optionally global
negatives through an embedding all-gather).  Out of scope: the MMRS dataset, .osu
parsing, tokenizers and WandB/hub plumbing; `dataset.synthetic=true` (default when the reference's
data stack is not importable) trains on seeded synthetic windows with the processor's output schema.
"""
from __future__ import annotations

import copy
import json
import logging
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

from cm3p_b200 import hydra_lite  # noqa: E402
from cm3p_b200.configuration_cm3p import CM3PConfig  # noqa: E402
from cm3p_b200.synthetic import synthetic_batch  # noqa: E402

logger = logging.getLogger("cm3p_b200.train")


def _last_checkpoint(output_dir: str):
    if not os.path.isdir(output_dir):
        return None
    steps = []
    for d in os.listdir(output_dir):
        if d.startswith("checkpoint-") and d.split("-", 1)[1].isdigit():
            steps.append(int(d.split("-", 1)[1]))
    return os.path.join(output_dir, f"checkpoint-{max(steps)}") if steps else None


class SyntheticWindows:
    """Seeded synthetic batches with the processor's output schema."""

    def __init__(self, config: CM3PConfig, ds, batch: int, rank: int, seed: int, model_cls: str):
        self.config, self.ds, self.batch, self.rank, self.seed, self.model_cls = config, ds, batch, rank, seed, model_cls

    def get(self, step: int, micro: int) -> dict:
        ds = self.ds
        V = int(ds.get("train_metadata_variations", 1))
        if ds.get("fixed_batch", False):  # overfit-one-batch mode (smoke tests)
            step = micro = 0
        b = synthetic_batch(self.config, batch=self.batch, seq_len=int(ds.get("seq_len", 2000)), variations=max(V, 1),
                            seed=(self.seed + 1000003 * step + 101 * micro + 7 * self.rank) % (2 ** 32 - 1),
                            min_len=ds.get("min_len"), with_labels=(ds.get("labels") == "masked_lm"))
        if self.model_cls == "CM3PForMaskedLM":
            return {k: b[k] for k in ("input_ids", "attention_mask", "input_features", "labels")}
        if self.model_cls == "CM3PForBeatmapClassification":
            g = torch.Generator().manual_seed(self.seed + step)
            out = {k: b[k] for k in ("input_ids", "attention_mask", "input_features")}
            out["labels"] = torch.randint(0, max(int(self.config.num_labels), 2), (self.batch,), generator=g)
            return out
        return b


def _build_optimizer(model, tr):
    optim = str(tr.get("optim", "adamw_torch"))
    lr = float(tr.get("learning_rate", 1e-4))
    betas = (float(tr.get("adam_beta1", 0.9)), float(tr.get("adam_beta2", 0.999)))
    eps, wd = float(tr.get("adam_epsilon", 1e-8)), float(tr.get("weight_decay", 0.0))
    if optim == "muon":
        from cm3p_b200.muon import Muon, split_muon_adamw
        muon_params, adamw_params = split_muon_adamw(model)
        logger.info("Number of parameters for Muon: %d, for AdamW: %d", len(muon_params), len(adamw_params))
        return Muon(muon_params=muon_params, lr=lr, adamw_lr=lr / 4, adamw_params=adamw_params, adamw_betas=betas,
                    adamw_wd=wd, adamw_eps=eps)
    return torch.optim.AdamW([p for p in model.parameters()], lr=lr, betas=betas, eps=eps, weight_decay=wd)


def _lr_at(step: int, tr) -> float:
    """HF Trainer default schedule: linear warm-up then linear decay to 0 at max_steps."""
    base = float(tr.get("learning_rate", 1e-4))
    warm, total = int(tr.get("warmup_steps", 0)), int(tr.get("max_steps", 1))
    kind = str(tr.get("lr_scheduler_type", "linear"))
    if step < warm:
        return base * (step + 1) / max(1, warm)
    if kind == "constant":
        return base
    return base * max(0.0, (total - step) / max(1, total - warm))


def main(argv=None) -> dict:
    argv = sys.argv[1:] if argv is None else argv
    cfg_dir, cfg_name, overrides = hydra_lite.parse_cli(argv, os.path.join(ROOT, "configs", "train"), "synthetic")
    args = hydra_lite.compose(cfg_dir, cfg_name, overrides)
    tr, ds = args.training, args.get("dataset", hydra_lite.Cfg())
    logging.basicConfig(format="%(asctime)s - %(levelname)s - %(name)s - %(message)s", level=logging.INFO,
                        handlers=[logging.StreamHandler(sys.stdout)])

    import torch.distributed as dist
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("train.py: cm3p_b200 trains on CUDA sm_100a devices only (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    if rank != 0:
        logger.setLevel(logging.WARNING)

    seed = int(tr.get("seed", 42))
    torch.manual_seed(seed)
    if tr.get("deterministic", False):
        # bit-reproducible weight gradients (ordered split-K accumulation instead of fp32 atomics); slower
        from cm3p_b200 import ops as _ops
        _ops.set_option(_ops.OPT_WGRAD_DETERMINISTIC, 1)

    # ---- model (train.py:274-321)
    from cm3p_b200.modeling_cm3p import CM3PForBeatmapClassification, CM3PForMaskedLM, CM3PModel
    model_config = CM3PConfig(**copy.deepcopy(hydra_lite.to_container(args.model)))
    model_config._attn_implementation = args.get("attn_implementation", "flash_attention_2")
    model_cls = str(args.get("model_cls", "CM3PModel"))
    klass = {"CM3PForMaskedLM": CM3PForMaskedLM, "CM3PForBeatmapClassification": CM3PForBeatmapClassification}.get(
        model_cls, CM3PModel)
    sub_config = model_config if klass is CM3PModel else model_config.beatmap_config
    output_dir = str(tr.get("output_dir", "runs/default"))
    checkpoint = tr.get("resume_from_checkpoint")
    if checkpoint is None and not tr.get("overwrite_output_dir", False):
        checkpoint = _last_checkpoint(output_dir)
    if checkpoint is not None:
        logger.info("Checkpoint detected, resuming training at %s", checkpoint)
        model = klass.from_pretrained(checkpoint, config=sub_config)
    elif args.get("from_pretrained") is not None:
        logger.warning("Loading model from %s", args.from_pretrained)
        model = klass.from_pretrained(args.from_pretrained, config=sub_config)
    else:
        model = klass(sub_config)
    model = model.to(dev).float().train()
    if args.get("freeze_beatmap_model", False):
        for p in model.beatmap_model.parameters():
            p.requires_grad = False
    if args.get("freeze_metadata_model", False) and hasattr(model, "metadata_model"):
        for p in model.metadata_model.parameters():
            p.requires_grad = False

    if world > 1:
        from cm3p_b200 import distributed as D
        dp = D.enable_data_parallel(model, global_negatives=bool(tr.get("global_negatives", False)))
        D.broadcast_parameters(model, dp)

    optimizer = _build_optimizer(model, tr)
    start_step = 0
    if checkpoint is not None and os.path.isfile(os.path.join(checkpoint, "optimizer.pt")):
        optimizer.load_state_dict(torch.load(os.path.join(checkpoint, "optimizer.pt"), map_location=dev))
        with open(os.path.join(checkpoint, "trainer_state.json")) as f:
            start_step = int(json.load(f)["global_step"])

    # ---- data
    if not ds.get("synthetic", True):
        raise RuntimeError("train.py: only dataset.synthetic=true is available here; the reference's MMRS dataset / "
                           "processor stack (slider, librosa, tokenizers) is outside this framework's scope")
    per_dev = int(tr.get("per_device_train_batch_size", 8))
    accum = int(tr.get("gradient_accumulation_steps", 1))
    data = SyntheticWindows(model_config, ds, per_dev, rank, seed, model_cls)

    # ---- loop (what transformers.Trainer.train does for this model: train.py:360-375)
    max_steps = int(tr.get("max_steps", 100))
    log_every, save_every = int(tr.get("logging_steps", 10)), int(tr.get("save_steps", 0) or 0)
    keep = int(tr.get("save_total_limit", 0) or 0)
    unfreeze_at = args.get("unfreeze_beatmap_model_at_step")
    history, t_log, seen = [], time.perf_counter(), 0

    def save(step: int):
        if rank != 0:
            return
        path = os.path.join(output_dir, f"checkpoint-{step}")
        os.makedirs(path, exist_ok=True)
        model.save_pretrained(path)
        torch.save(optimizer.state_dict(), os.path.join(path, "optimizer.pt"))
        with open(os.path.join(path, "trainer_state.json"), "w") as f:
            json.dump({"global_step": step, "log_history": history}, f)
        if keep > 0:
            steps = sorted(int(d.split("-", 1)[1]) for d in os.listdir(output_dir) if d.startswith("checkpoint-"))
            for old in steps[:-keep]:
                shutil.rmtree(os.path.join(output_dir, f"checkpoint-{old}"), ignore_errors=True)

    # ---- evaluation (transformers.Trainer.evaluate with batch_eval_metrics + compute_metrics: train.py:38-160,
    #      :366-371, :376-386): eval loss, zero-shot variation accuracies, masked-LM / classification accuracy
    from cm3p_b200.metrics import EvalPrediction, compute_metrics
    eval_every = int(tr.get("eval_steps", 0) or 0) if str(tr.get("eval_strategy", "no")) != "no" else 0
    eval_batch = int(tr.get("per_device_eval_batch_size", per_dev))
    eval_batches = int(ds.get("eval_batches", 2))
    eval_ds = hydra_lite.Cfg(dict(hydra_lite.to_container(ds)))
    eval_ds["train_metadata_variations"] = int(ds.get("test_metadata_variations", ds.get("train_metadata_variations", 1)))
    eval_ds["fixed_batch"] = False
    eval_data = SyntheticWindows(model_config, eval_ds, eval_batch, rank, seed + 7919, model_cls)

    def evaluate(step: int) -> dict:
        was_training = model.training
        model.eval()
        total = torch.zeros((), device=dev)
        metrics = {}
        with torch.no_grad():
            for i in range(eval_batches):
                batch = {k: v.to(dev, non_blocking=True) for k, v in eval_data.get(4000 + i, 0).items()}
                out = model(**batch)
                total += out.loss.detach().float()
                if hasattr(out, "logits_per_beatmap"):
                    # what transformers.Trainer.prediction_step hands to compute_metrics: every output but the loss,
                    # in CM3POutput's field order (predictions[0] = logits_per_beatmap, [4] = logits; train.py:77,101)
                    preds = tuple(v for k, v in out.items() if k != "loss")
                else:
                    preds = out.logits
                labels = batch.get("labels")
                if labels is not None and not isinstance(preds, tuple) and preds.dim() == 2 and labels.dim() == 2:
                    labels = labels[labels != -100]  # sparse prediction returns only the labelled rows
                metrics = compute_metrics(EvalPrediction(preds, labels, batch), i + 1 == eval_batches) or {}
        if world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.AVG)
        if was_training:
            model.train()
        rec = {"step": step, "eval_loss": float(total) / max(1, eval_batches)}
        rec.update({f"eval_{k}": v for k, v in metrics.items()})
        history.append(rec)
        logger.info(json.dumps(rec))
        return rec

    loss_acc = torch.zeros((), device=dev)
    for step in range(start_step, max_steps):
        lr = _lr_at(step, tr)
        for group in optimizer.param_groups:
            group["lr"] = lr
        optimizer.zero_grad(set_to_none=True)
        for micro in range(accum):
            batch = {k: v.to(dev, non_blocking=True) for k, v in data.get(step, micro).items()}
            out = model(**batch)
            (out.loss / accum).backward()
            loss_acc += out.loss.detach() / accum
            seen += per_dev * world
        optimizer.step()
        if unfreeze_at is not None and step + 1 == int(unfreeze_at):
            logger.info("Unfreezing beatmap_model at step %d", step + 1)
            for p in model.beatmap_model.parameters():
                p.requires_grad = True
        if (step + 1) % log_every == 0 or step + 1 == max_steps:
            n = (step + 1 - start_step) % log_every or log_every
            dt = time.perf_counter() - t_log
            rec = {"step": step + 1, "loss": float(loss_acc) / n, "learning_rate": lr,
                   "pairs_per_sec": round(seen / dt, 2)}
            history.append(rec)
            logger.info(json.dumps(rec))
            loss_acc.zero_()
            t_log, seen = time.perf_counter(), 0
        if eval_every and (step + 1) % eval_every == 0:
            evaluate(step + 1)
        if save_every and (step + 1) % save_every == 0:
            save(step + 1)
    if tr.get("do_eval", False) and not (eval_every and max_steps % eval_every == 0 and max_steps > start_step):
        evaluate(max_steps)
    if tr.get("do_train", True) and max_steps > start_step:
        save(max_steps)
        if rank == 0:
            model.save_pretrained(output_dir)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return {"log_history": history, "output_dir": output_dir}


if __name__ == "__main__":
    main()
