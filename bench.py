"""bench.py — headline benchmark of the CM3P hot path on B200 (contract: see the build brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload all|train|infer|mlm] [--variations V] [--train-batch B] [--global-negatives]

One "step" = one pass of the hot path over one batch of synthetic input per GPU.

  workload train (BASELINE.json configs[2], THE HEADLINE LINE): the contrastive train step, fp32 master
      weights / bf16 compute, batch 256 windows/GPU of 16 s (L = 2000 padded, real lengths U{600..2000},
      200 audio tokens + 80x1600 log-mel per window), V = 8 metadata variations (`--variations 256` = the value
      the reference trains with, configs/train/v7.yaml:40), forward + explicit backward + bucketed gradient
      all-reduce (NCCL, overlapped with the backward pass) when N > 1 -> pairs/s.  The optimizer step is
      excluded from the step and reported separately as `optimizer_step_ms` (Muon + AdamW).
      `--global-negatives --train-batch 512 --variations 1` = configs[3] (embedding all-gather, global loss).
  workload infer (configs[1]): base CM3P bf16, batch 64 windows/GPU, beatmap tower + audio encoder +
      metadata tower (V = 1) + projections + logits, `return_loss=False` -> beatmap embeds/s.
  workload mlm (configs[4]): CM3PForMaskedLM train step on 8192-token windows, 8 windows/GPU, 15 % masked
      -> real tokens/s.
  workload all (default): the train line with the inference result attached under the key "infer".

Prints ONE JSON line on rank 0.  `value` is measured with inputs resident in HBM; `e2e` goes through the
public `CM3PModel.__call__` with pinned host inputs (H2D inside the timed region) and a D2H read of the
result.  `--impl reference` times the reference's own CPU path (the unmodified reference model from
oracle/_ref when present, else the oracle port) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

from cm3p_b200.configuration_cm3p import CM3PConfig, base_config_dict  # noqa: E402
from cm3p_b200.synthetic import synthetic_batch, synthetic_state_dict  # noqa: E402

BATCH_PER_GPU = {"infer": 64, "train": 256}
SEQ_LEN = 2000
MIN_LEN = 600
TRAIN_VARIATIONS = 8
METRIC = {"infer": ("beatmap_embeds_per_sec", "embeds/s"), "train": ("train_pairs_per_sec", "pairs/s"),
          "mlm": ("mlm_train_tokens_per_sec", "tokens/s")}
MLM_SEQ_LEN, MLM_BATCH = 8192, 8
# dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch (average over the GEMM launches of one step), from
# the ncu launch lists committed under profiles/ (it cannot be measured from inside this script)
GEMM_TRAFFIC = {
    "infer": (271.4e6, "profiles/r2_infer_launch_summary_v4.txt (ncu, batch 64)"),
    "train": (1334.4e6, "profiles/r2_train256_launch_summary_v4.txt (ncu, batch 256, V=8)"),
}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tflops=p["bf16_tflops_sustained"], source="MEASURED_PEAKS.json (sustained)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def _algorithmic_flops_infer(cfg: CM3PConfig, batch: dict) -> float:
    """Forward FLOPs of one step, real tokens only."""
    bc, mc, ac = cfg.beatmap_config, cfg.metadata_config, cfg.beatmap_config.audio_config

    def tower_linear(c):
        H, I = c.hidden_size, int(c.intermediate_size)
        return c.num_hidden_layers * 2 * (H * 3 * H + H * H + H * 2 * I + I * H)

    def attn(c, ln):
        n_glob = sum(1 for i in range(c.num_hidden_layers) if c.layer_is_global(i))
        n_loc = c.num_hidden_layers - n_glob
        return n_glob * 4 * ln * ln * c.hidden_size + n_loc * 4 * ln * min(ln, 2 * c.window_half + 1) * c.hidden_size

    lens = batch["attention_mask"].sum(-1).tolist()
    B = len(lens)
    total = sum(lens) * tower_linear(bc) + sum(attn(bc, n) for n in lens)
    frames = batch["input_features"].shape[-1]
    t2 = frames // 2
    total += B * (2 * frames * ac.hidden_size * 3 * ac.n_mels + 2 * t2 * ac.hidden_size * 3 * ac.hidden_size)
    total += B * (t2 * tower_linear(ac) + attn(ac, t2))
    total += B * (t2 // 4) * 2 * (ac.projector_intermediate_size * ac.projector_dim + ac.projector_dim ** 2)
    mlens = batch["metadata_attention_mask"].reshape(-1, batch["metadata_attention_mask"].shape[-1]).sum(-1)
    mg = sum(1 for i in range(mc.num_hidden_layers) if mc.layer_is_global(i))
    total += float(mlens.sum()) * tower_linear(mc)
    total += float((mlens.double() ** 2).sum()) * 4 * mc.hidden_size * mg  # all metadata layers are global in the base config
    total += B * 2 * bc.hidden_size * cfg.projection_dim + len(mlens) * 2 * mc.hidden_size * cfg.projection_dim
    total += 2 * len(mlens) * B * cfg.projection_dim
    return float(total)


def _make_batch(cfg, workload, rank, batch, variations):
    if workload == "mlm":
        b = synthetic_batch(cfg, batch=batch, seq_len=MLM_SEQ_LEN, seed=1 + rank, min_len=MLM_SEQ_LEN // 2,
                            with_labels=True)
        return {k: b[k] for k in ("input_ids", "attention_mask", "input_features", "labels")}
    V = 1 if workload == "infer" else variations
    return synthetic_batch(cfg, batch=batch, seq_len=SEQ_LEN, variations=V, seed=1 + rank, min_len=MIN_LEN)


WORKLOAD_TEXT = {
    "infer": "BASELINE.json configs[1]: CM3P base inference, beatmap tower + audio encoder + metadata tower (V=1) "
             "bf16, batch {B} synthetic 16 s windows per GPU (L=2000 padded, real lengths U{{600..2000}}), "
             "return_loss=False",
    "train": "BASELINE.json configs[{cfgno}]: CM3P base contrastive train step (audio fusion + V={V} metadata "
             "variations), fp32 master weights / bf16 compute, batch {B} synthetic 16 s windows per GPU (L=2000 "
             "padded, real lengths U{{600..2000}}), forward + backward + bucketed overlapped gradient all-reduce, "
             "{neg} negatives; optimizer step excluded (reported as optimizer_step_ms)",
    "mlm": "BASELINE.json configs[4]: CM3PForMaskedLM train step (beatmap tower + audio encoder + MLM head), fp32 "
           "master weights / bf16 compute, {B} windows of L=8192 per GPU (real lengths U{{4096..8192}}), 15 % masked, "
           "forward + backward + gradient all-reduce ({neg}); value = real tokens/s; optimizer step excluded",
}


def _bench_workload(workload, args, dist, dev, world, rank, local_rank) -> dict | None:
    """Times one workload on this rank's GPU; returns the result dict on rank 0."""
    from cm3p_b200 import distributed as dp_utils
    from cm3p_b200 import ops
    from cm3p_b200 import training as _training
    from cm3p_b200.modeling_cm3p import CM3PModel

    train = workload in ("train", "mlm")
    V = args.variations if workload == "train" else 1
    B = MLM_BATCH if workload == "mlm" else (args.train_batch if train else BATCH_PER_GPU["infer"])
    cfg = CM3PConfig(attn_implementation="flash_attention_2",
                     **copy.deepcopy(base_config_dict(has_decoder_head=(workload == "mlm"))))
    if workload == "mlm":
        from cm3p_b200.modeling_cm3p import CM3PForMaskedLM
        model = CM3PForMaskedLM(cfg.beatmap_config)
        model.load_state_dict({k: v for k, v in synthetic_state_dict(cfg, seed=0).items()
                               if k.startswith(("beatmap_model.", "head.", "decoder."))}, strict=True)
    else:
        model = CM3PModel(cfg)
        model.load_state_dict(synthetic_state_dict(cfg, seed=0), strict=True)
    model = model.to(dev)
    dp = None
    if train:
        model.train()
        if world > 1:
            dp = dp_utils.enable_data_parallel(model, global_negatives=args.global_negatives)
    else:
        model = model.to(torch.bfloat16).eval()

    host = _make_batch(cfg, workload, rank, B, V)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}

    def step(feed):
        if train:
            model.zero_grad(set_to_none=True)
            out = model(**feed)
            out.loss.backward()  # explicit CUDA backward; includes the gradient all-reduce when world > 1
            return out
        with torch.no_grad():
            return model(**feed, return_loss=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    steps = args.steps
    n_warm = args.warmup if args.quick else max(args.warmup, 3)
    torch.cuda.reset_peak_memory_stats()
    for _ in range(n_warm):
        step(resident)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ops.LAUNCH_COUNT
    ms_step = timed(lambda: step(resident), steps)
    launches = (ops.LAUNCH_COUNT - launches0) // steps
    clocks = sampler.stop() if rank == 0 else None
    if args.quick:
        return {"quick": True, "workload": workload, "ms_per_step": round(ms_step, 3),
                "gpu_launches": int(launches)} if rank == 0 else None

    # ---- end-to-end through the public API with pinned host inputs and a D2H read of the result.
    # Every step copies ITS inputs host -> device and its result device -> host inside the timed region; as
    # in any serving / training input pipeline the copies are double-buffered: the inputs of step i+1 travel on
    # a copy stream while step i computes, and the host reads the result of step i (from pinned memory, after
    # its copy event) once step i+1 has been enqueued, so the host never idles the GPU.
    copy_stream = torch.cuda.Stream(device=dev)
    result_shape = () if train else (B, cfg.projection_dim)
    result_host = [torch.empty(result_shape, dtype=torch.float32).pin_memory() for _ in range(2)]
    pipe = {"i": 0, "pending": None, "last": None}

    def upload():
        with torch.cuda.stream(copy_stream):
            feed = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return feed, ev

    pipe["next"] = upload()

    def e2e_step():
        feed, ev = pipe["next"]
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        for v in feed.values():
            v.record_stream(cur)
        pipe["next"] = upload()  # inputs of the following step, under this step's compute
        out = step(feed)
        res = out.loss.detach().float() if train else out.beatmap_embeds.float()
        slot = pipe["i"] & 1
        result_host[slot].copy_(res, non_blocking=True)
        done = torch.cuda.Event()
        done.record(cur)
        if pipe["pending"] is not None:  # read the previous step's result on the host
            prev_slot, prev_done = pipe["pending"]
            prev_done.synchronize()
            pipe["last"] = float(result_host[prev_slot].flatten()[0])
        pipe["pending"] = (slot, done)
        pipe["i"] += 1

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, steps)
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    d2h = 4 if train else B * cfg.projection_dim * 4

    # ---- communication (N > 1, train): the gradient all-reduce alone, and how much of it the step still shows
    comm = None
    if train and dp is not None:
        n_params = sum((p.numel() + 3) // 4 * 4 for p in model.parameters())
        flat = torch.zeros(n_params, device=dev, dtype=torch.float32)
        for _ in range(2):
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        ar_ms = timed(lambda: dist.all_reduce(flat, op=dist.ReduceOp.AVG), 5)
        del flat
        # step time with and without the collectives, INTERLEAVED step by step (the power-capped step drifts by a few
        # per cent between runs, far more than the reduction costs): medians of per-step CUDA-event times
        def one(with_dp):
            model._dp = dp if with_dp else None
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step(resident)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b)
        one(False)
        one(True)
        t_with, t_without = [], []
        for _ in range(max(6, min(steps, 12))):
            t_without.append(one(False))
            t_with.append(one(True))
        model._dp = dp
        # the exposed tail measured directly: CUDA events around the compute stream's wait for the collectives at the end
        # of the backward pass (the last bucket can start only when the backward pass is over)
        _training.GradStore.TAIL_EVENTS = []
        for _ in range(5):
            step(resident)
        torch.cuda.synchronize()
        tails = [a.elapsed_time(b) for a, b in _training.GradStore.TAIL_EVENTS]
        _training.GradStore.TAIL_EVENTS = None
        # MIN over ranks: every rank but the one that finishes its backward pass last also waits for the slower ranks
        # (ragged batches: the ranks' token counts differ by a few per cent), which is load imbalance, not exposed
        # communication; the last rank's wait is the collective's own tail.  The MAX is reported next to it.
        tail_t = torch.tensor([statistics.median(tails) if tails else 0.0], device=dev, dtype=torch.float64)
        tail_max_t = tail_t.clone()
        dist.all_reduce(tail_t, op=dist.ReduceOp.MIN)
        dist.all_reduce(tail_max_t, op=dist.ReduceOp.MAX)
        tail_ms, tail_max_ms = float(tail_t.item()), float(tail_max_t.item())
        med = torch.tensor([statistics.median(t_with), statistics.median(t_without)], device=dev, dtype=torch.float64)
        dist.all_reduce(med, op=dist.ReduceOp.MAX)
        ms_comm, ms_nocomm = float(med[0]), float(med[1])
        exposed = max(0.0, ms_comm - ms_nocomm)
        comm = {"all_reduce_ms": round(ar_ms, 3), "all_reduce_bytes": int(n_params * 4),
                "all_reduce_busbw_gbs": round(2 * (world - 1) / world * n_params * 4 / (ar_ms * 1e-3) / 1e9, 1),
                "step_ms_without_collectives": round(ms_nocomm, 3), "step_ms_with_collectives": round(ms_comm, 3),
                "exposed_ms": round(exposed, 3), "exposed_tail_ms": round(tail_ms, 3),
                "wait_incl_rank_skew_ms": round(tail_max_ms, 3), "grad_overlap": args.grad_overlap,
                "nccl_max_ctas": args.nccl_max_ctas or None,
                "overlap": round(min(1.0, max(0.0, 1.0 - tail_ms / ar_ms)), 3) if ar_ms > 0 else None,
                "how": "all_reduce_ms = the whole fp32 gradient buffer reduced alone (CUDA events, max over ranks, "
                       "5 reps); exposed_tail_ms = CUDA events around the compute stream's wait for the bucketed "
                       "collectives at the end of the backward pass (median of 5 steps, MIN over ranks = the rank that "
                       "finishes last; wait_incl_rank_skew_ms = MAX over ranks: the faster ranks also wait for the slower "
                       "ones, whose ragged batches hold a few per cent more tokens); overlap = 1 - exposed_tail / all_reduce; exposed_ms = median step time with minus without the collectives, "
                       "steps interleaved one by one (a cross-check: the power-capped step drifts by more than the "
                       "collective costs)"}

    # ---- optimizer step (Muon for the matrices, its internal AdamW for embeddings / vectors), reported separately
    opt_ms = None
    if workload == "train":
        from cm3p_b200.muon import Muon, split_muon_adamw
        muon_params, adamw_params = split_muon_adamw(model)
        opt = Muon(muon_params, lr=4e-4, adamw_params=adamw_params)
        step(resident)  # fresh gradients
        # lr = 0 steps would still do all the work, but keep the weights finite and realistic: tiny lr
        for group in opt.param_groups:
            group["lr"] = 1e-6
        opt.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            opt.step()
        e1.record()
        torch.cuda.synchronize()
        opt_ms = e0.elapsed_time(e1) / 3
        del opt

    # ---- roofline of the dominant kernel family (the tcgen05 GEMM), timed live with CUDA events around
    #      every GEMM launch of one extra step on the launching stream
    gemm_events = []
    gemm_bytes_box = [0.0]
    orig_gemm = ops.gemm

    def timed_gemm(a, b, **kw):
        M, K = (a.shape[1], a.shape[0]) if kw.get("trans_a") else a.shape
        N = b.shape[1] if kw.get("trans_b") else b.shape[0]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_gemm(a, b, **kw)
        e1.record()
        gemm_events.append((e0, e1, 2.0 * M * N * K))
        gemm_bytes_box[0] += a.numel() * 2 + b.numel() * 2 + out.numel() * out.element_size()
        return out

    attn_events = []
    orig_fwd, orig_bwd = ops.attn_varlen_fwd, ops.attn_varlen_bwd

    def _attn_flops(cu_seqlens, heads, window, factor):
        lens = (cu_seqlens[1:] - cu_seqlens[:-1]).double()
        keys = lens if window < 0 else torch.clamp(lens, max=2 * window + 1)
        return factor * float((lens * keys).sum()) * 64 * heads

    def timed_attn_fwd(qkv, cu_seqlens, max_seqlen, heads, window=-1, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_fwd(qkv, cu_seqlens, max_seqlen, heads, window, **kw)
        e1.record()
        attn_events.append((e0, e1, cu_seqlens, heads, window, 4.0))
        return out

    def timed_attn_bwd(qkv, out, dout, lse, cu_seqlens, max_seqlen, heads, window=-1, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig_bwd(qkv, out, dout, lse, cu_seqlens, max_seqlen, heads, window, **kw)
        e1.record()
        attn_events.append((e0, e1, cu_seqlens, heads, window, 10.0))
        return r

    ops.gemm, ops.attn_varlen_fwd, ops.attn_varlen_bwd = timed_gemm, timed_attn_fwd, timed_attn_bwd
    try:
        step(resident)
        torch.cuda.synchronize()
    finally:
        ops.gemm, ops.attn_varlen_fwd, ops.attn_varlen_bwd = orig_gemm, orig_fwd, orig_bwd
    attn_ms = sum(e0.elapsed_time(e1) for e0, e1, *_ in attn_events)
    attn_flops = sum(_attn_flops(cu, h, w, f) for _, _, cu, h, w, f in attn_events)
    gemm_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in gemm_events)
    gemm_flops = sum(f for _, _, f in gemm_events)
    gemm_bytes = gemm_bytes_box[0]
    peaks = _peaks()
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0

    units = float(host["attention_mask"].sum()) if workload == "mlm" else float(B)  # tokens or windows per step
    total_flops = 0.0 if workload == "mlm" else _algorithmic_flops_infer(cfg, host) * (3.0 if train else 1.0)
    metric, unit = METRIC[workload]
    glob = bool(train and args.global_negatives and world > 1)
    neg = "global (embedding all-gather)" if glob else "local (per-rank)"
    default_cfg = workload != "train" or (V == TRAIN_VARIATIONS and B == BATCH_PER_GPU["train"])
    result = {
        "metric": metric, "value": round(world * units / (ms_step * 1e-3), 2), "unit": unit, "n_gpus": world,
        "steps": steps, "warmup": n_warm, "ms_per_step": round(ms_step, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "impl": "ours",
        "config": {
            "workload": WORKLOAD_TEXT[workload].format(B=B, neg=neg, V=V, cfgno=3 if args.global_negatives else 2),
            "batch_per_gpu": B, "variations": V, "global_batch": B * world,
            "seq_len": MLM_SEQ_LEN if workload == "mlm" else SEQ_LEN,
            "real_tokens_per_step": int(host["attention_mask"].sum()),
            "metadata_sequences_per_step": int(host["metadata_attention_mask"].numel() // host["metadata_attention_mask"].shape[-1]) if "metadata_attention_mask" in host else 0,
            "parallelism": f"dp{world}",
            "weights": "random init (seeded), 136.9 M params",
            "l2": "per-step working set (GBs of activations) exceeds the 126 MB L2; no explicit flush",
        },
        "e2e": {"value": round(world * units / (ms_e2e * 1e-3), 2), "unit": unit, "ms_per_step": round(ms_e2e, 3),
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "model_tflops": round(total_flops / (ms_step * 1e-3) / 1e12, 1),
        "mfu": round(total_flops / (ms_step * 1e-3) / 1e12 / peaks["tflops"], 4),
        "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
        "roofline": {"kernel": "gemm_bf16_sm100_kernel (all epilogues)", "bound": "tensor",
                     "achieved": round(achieved, 1), "peak": peaks["tflops"], "unit": "TFLOP/s",
                     "frac": round(achieved / peaks["tflops"], 4),
                     "traffic": GEMM_TRAFFIC[workload][0] if (workload in GEMM_TRAFFIC and default_cfg) else None,
                     "traffic_source": GEMM_TRAFFIC[workload][1] if workload in GEMM_TRAFFIC else None,
                     "algorithmic_bytes_per_launch": round(gemm_bytes / max(1, len(gemm_events))),
                     "launches_per_step": len(gemm_events), "share_of_step": round(gemm_ms / ms_step, 3),
                     "peak_source": peaks["source"]},
        "roofline_attention": {"kernel": "attn_fwd_v2 / attn_bwd_dq + attn_bwd_dkv (global layers) / attn_bwd_win band walk "
                                         "(window layers), varlen, D=64; packed short-sequence kernels for the metadata tower", "bound": "tensor",
                               "achieved": round(attn_flops / (attn_ms * 1e-3) / 1e12, 1) if attn_ms > 0 else 0.0,
                               "peak": peaks["tflops"], "unit": "TFLOP/s",
                               "frac": round(attn_flops / (attn_ms * 1e-3) / 1e12 / peaks["tflops"], 4) if attn_ms > 0 else 0.0,
                               "flops": "algorithmic: 4*l*keys*64 per head forward, 10*l*keys*64 backward (5 GEMMs)",
                               "note": ("head_dim 64: a 128x128 tile needs 16384 exp2 = 1024 clk of the 16/clk/SM "
                                        "MUFU against 512 clk (forward) / 1280 clk (backward, 5 GEMMs) of tcgen05 "
                                        "MMA, so the exp unit caps the forward at 0.5 of the tensor peak; window "
                                        "layers additionally compute 256 keys per row for a 129-key band"),
                               "launches_per_step": len(attn_events), "share_of_step": round(attn_ms / ms_step, 3)},
        "clocks": clocks,
    }
    if comm is not None:
        result.update(all_reduce_ms=comm["all_reduce_ms"], overlap=comm["overlap"], comm=comm)
    elif train:
        result.update(all_reduce_ms=0.0, overlap=None)
    if opt_ms is not None:
        result["optimizer_step_ms"] = round(opt_ms, 3)
        result["optimizer"] = ("Muon (Newton-Schulz-5 x6, grouped over same-shape matrices) + internal AdamW; 3 timed "
                               "steps, CUDA events; not part of ms_per_step")
    if rank == 0 and world == 1 and not args.no_cpu_baseline and workload != "mlm":
        result["cpu_baseline"] = cpu_baseline(cfg, workload, sample_batch=2, reps=1, variations=V)
    del model, resident, pinned
    torch.cuda.empty_cache()
    return result if rank == 0 else None


def run_ours(args) -> dict | None:
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA sm_100a device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if args.nccl_max_ctas > 0:
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = args.nccl_max_ctas
            dist.init_process_group("nccl", device_id=dev, pg_options=opts)
        else:
            dist.init_process_group("nccl", device_id=dev)
    from cm3p_b200 import training as _training
    _training.GradStore.OVERLAP = args.grad_overlap == "on"
    try:
        order = ["train", "infer"] if args.workload == "all" else [args.workload]
        results = {w: _bench_workload(w, args, dist, dev, world, rank, local_rank) for w in order}
    finally:
        if world > 1:
            dist.destroy_process_group()
    if rank != 0:
        return None
    head = results[order[0]]
    if len(order) > 1:
        head["infer"] = results["infer"]
    return head


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation of the path on the host cores

_REF = {}


def _reference_model(cfg_dict):
    """The UNMODIFIED reference CM3PModel (oracle/_ref, vendored by oracle/build_ref.py; /root/reference in the
    build container) with the benchmark's seeded weights, or None when it is not available on this box."""
    if "model" in _REF:
        return _REF["model"]
    model = None
    try:
        from oracle.ref_shim import build_reference_model, reference_available
        if reference_available():
            model, rcfg = build_reference_model(cfg_dict, attn_implementation="sdpa")
            cfg = CM3PConfig(**copy.deepcopy(cfg_dict))
            missing, unexpected = model.load_state_dict(synthetic_state_dict(cfg, seed=0), strict=False)
            bad = [k for k in missing if "tok_embeddings" not in k or "audio_encoder" not in k]
            if bad or unexpected:
                raise RuntimeError(f"reference state dict mismatch: missing {bad[:4]} unexpected {list(unexpected)[:4]}")
    except Exception as exc:  # noqa: BLE001  (no reference on this box: fall back to the oracle port, say so)
        print(f"[bench] reference model unavailable ({type(exc).__name__}: {exc}); timing the oracle port", file=sys.stderr)
        model = None
    _REF["model"] = model
    return model


def _cpu_run(cfg, workload, sample_batch, sd, variations):
    """One pass of the reference's CPU path over `sample_batch` windows; returns (seconds, kind)."""
    from oracle import cm3p_oracle as O

    V = 1 if workload == "infer" else variations
    batch = synthetic_batch(cfg, batch=sample_batch, seq_len=SEQ_LEN, variations=V, seed=1, min_len=MIN_LEN)
    ref = _reference_model(base_config_dict())
    t0 = time.perf_counter()
    if ref is not None:
        if workload == "infer":
            with torch.no_grad():
                ref(**{k: v.clone() for k, v in batch.items()}, return_loss=False)
        else:
            ref.zero_grad(set_to_none=True)
            out = ref(**{k: v.clone() for k, v in batch.items()})
            out.loss.backward()
        return time.perf_counter() - t0, "reference"
    if workload == "infer":
        with torch.no_grad():
            O.model_forward(sd, cfg, **batch, return_loss=False)
    else:
        O.forward_backward(sd, cfg, batch)
    return time.perf_counter() - t0, "port"


def cpu_baseline(cfg, workload, sample_batch, reps, variations=TRAIN_VARIATIONS):
    """The reference's CPU `sdpa` path (fp32) on the host cores, on a bounded sample of the same workload."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synthetic_state_dict(cfg, seed=0)
    runs = [_cpu_run(cfg, workload, sample_batch, sd, variations) for _ in range(reps + 1)]  # first = warm-up
    best = min(t for t, _ in runs[1:])
    kind = runs[-1][1]
    metric, unit = METRIC[workload]
    what = "forward" if workload == "infer" else "forward + autograd backward"
    impl = "unmodified reference CM3PModel (oracle/_ref), sdpa" if kind == "reference" else "oracle port, torch CPU SDPA"
    return {"value": round(sample_batch / best, 4), "unit": unit, "cores": cores, "kind": kind,
            "sample": f"{sample_batch} windows of the same workload (L={SEQ_LEN}), {what}, fp32, {impl}, "
                      f"{cores} threads, best of {reps} after a warm-up pass; {best:.2f} s"}


def _reference_workload(workload, args) -> dict:
    cfg = CM3PConfig(**copy.deepcopy(base_config_dict()))
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synthetic_state_dict(cfg, seed=0)
    V = args.variations if workload == "train" else 1
    # bounded sample: 2 windows per step unless a probe pass says the whole run would exceed ~3 minutes
    probe, kind = _cpu_run(cfg, workload, 1, sd, V)   # also the lazy-initialisation pass
    probe, kind = _cpu_run(cfg, workload, 1, sd, V)
    sample = 2 if 2 * probe * (args.warmup + args.steps) <= 180.0 else 1
    times = []
    for i in range(args.warmup + args.steps):
        dt, kind = _cpu_run(cfg, workload, sample, sd, V)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    metric, unit = METRIC[workload]
    value = round(sample / (ms * 1e-3), 4)
    what = "forward" if workload == "infer" else "forward + autograd backward"
    impl = "unmodified reference CM3PModel (oracle/_ref), sdpa" if kind == "reference" else "oracle port, torch CPU SDPA"
    return {
        "impl": "reference", "metric": metric, "value": value, "unit": unit,
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[workload].format(B=sample, neg="local (per-rank)", V=V, cfgno=2)
                   + f" -- CPU arm: each step = a bounded sample of {sample} windows, {what}", "seq_len": SEQ_LEN,
                   "variations": V},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": kind,
                         "sample": f"{sample} windows/step (L={SEQ_LEN}), {what}, fp32, {impl}, {cores} threads"},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def run_reference(args) -> dict | None:
    """The reference's own CPU implementation of the path on the host cores.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return None
    order = ["train", "infer"] if args.workload == "all" else [args.workload]
    if "mlm" in order:
        return {"impl": "reference", "unavailable": "the CPU arm covers the train and infer workloads only"}
    res = {w: _reference_workload(w, args) for w in order}
    head = res[order[0]]
    if len(order) > 1:
        head["infer"] = res["infer"]
    return head


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["all", "infer", "train", "mlm"], default="all")
    ap.add_argument("--train-batch", type=int, default=BATCH_PER_GPU["train"], help="train windows per GPU")
    ap.add_argument("--variations", type=int, default=TRAIN_VARIATIONS,
                    help="metadata variations per beatmap in the train step (reference v7: 256)")
    ap.add_argument("--global-negatives", action="store_true",
                    help="train: all-gather embeddings and use the global loss (BASELINE.json configs[3]: "
                         "--train-batch 512 --variations 1 on 8 GPUs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--grad-overlap", choices=["on", "off"], default="on",
                    help="train, N > 1: bucketed all-reduce during the backward pass (on) or one all-reduce after it (off)")
    ap.add_argument("--nccl-max-ctas", type=int, default=0,
                    help="cap the CTAs NCCL may use per collective (0 = NCCL's default); fewer CTAs leave more SMs to the "
                         "persistent compute kernels the collective overlaps with")
    ap.add_argument("--opt", action="append", default=[],
                    help="library option KEY=VALUE (cm3p_set_option, include/cm3p_b200.h CM3P_OPT_*), for A/B runs")
    ap.add_argument("--quick", action="store_true",
                    help="profiling aid: only the device-resident timed loop (no e2e / roofline / CPU legs)")
    args = ap.parse_args()
    if args.opt and args.impl == "ours":
        from cm3p_b200 import ops
        for kv in args.opt:
            k, v = kv.split("=")
            ops.set_option(int(k), int(v))
    res = run_reference(args) if args.impl == "reference" else run_ours(args)
    if res is not None:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
