"""The UNMODIFIED reference model on the B200 itself — the library path our kernels have to beat
(SURVEY.md §8d "also time the reference on the B200").  TEST INFRASTRUCTURE.

    python tools/bench_reference_gpu.py [--steps 5] [--infer-batch 64] [--train-batch 32] [--attn sdpa|fa2]

Imports the reference through oracle/ref_shim.py (oracle/_ref on the GPU box), loads the benchmark's seeded weights,
and times (CUDA events, after warm-up):
  infer : model.to(bfloat16), batch 64 x L=2000 padded, return_loss=False      (extract_beatmap_embeddings.py:161-232)
  train : fp32 master weights under torch.autocast(bfloat16) like HF Trainer `bf16: true`, forward + backward,
          V=8, the largest batch that is asked for (the reference pads every window to L, so its activation
          memory is ~1.5x ours per real token and it keeps every autograd intermediate)
`--attn fa2` switches the three towers' sub-configs to flash_attention_2 while the top-level config stays `sdpa`
(the reference's own FA2 unpadding path cannot run on transformers 5.5).
Prints one JSON line per workload; nothing of cm3p_b200's kernels is on this path.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from cm3p_b200.configuration_cm3p import CM3PConfig, base_config_dict  # noqa: E402
from cm3p_b200.synthetic import synthetic_batch, synthetic_state_dict  # noqa: E402
from oracle.ref_shim import build_reference_model  # noqa: E402


def _timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--infer-batch", type=int, default=64)
    ap.add_argument("--train-batch", type=int, default=32)
    ap.add_argument("--variations", type=int, default=8)
    ap.add_argument("--attn", choices=["sdpa", "fa2"], default="sdpa")
    ap.add_argument("--workload", choices=["all", "infer", "train"], default="all")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg_dict = base_config_dict()
    cfg = CM3PConfig(**copy.deepcopy(cfg_dict))
    sd = synthetic_state_dict(cfg, seed=0)

    def build():
        model, rcfg = build_reference_model(cfg_dict, attn_implementation="sdpa")
        model.load_state_dict(sd, strict=False)
        if args.attn == "fa2":
            for c in (rcfg.metadata_config, rcfg.beatmap_config, rcfg.beatmap_config.audio_config):
                c._attn_implementation = "flash_attention_2"
        return model

    for workload in (["infer", "train"] if args.workload == "all" else [args.workload]):
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        try:
            if workload == "infer":
                B = args.infer_batch
                model = build().to(dev).to(torch.bfloat16).eval()
                batch = synthetic_batch(cfg, batch=B, seq_len=2000, variations=1, seed=1, min_len=600)
                feed = {k: (v.to(dev).to(torch.bfloat16) if v.is_floating_point() else v.to(dev)) for k, v in batch.items()}

                def step():
                    with torch.no_grad():
                        model(**{k: v.clone() for k, v in feed.items()}, return_loss=False)
                unit, metric = "embeds/s", "beatmap_embeds_per_sec"
            else:
                B = args.train_batch
                model = build().to(dev).train()
                batch = synthetic_batch(cfg, batch=B, seq_len=2000, variations=args.variations, seed=1, min_len=600)
                feed = {k: v.to(dev) for k, v in batch.items()}

                def step():
                    model.zero_grad(set_to_none=True)
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        out = model(**{k: v.clone() for k, v in feed.items()})
                    out.loss.backward()
                unit, metric = "pairs/s", "train_pairs_per_sec"
            ms = _timed(step, args.steps, args.warmup)
            line = {"impl": "reference-on-gpu", "metric": metric, "value": round(B / (ms * 1e-3), 2), "unit": unit,
                    "ms_per_step": round(ms, 2), "batch": B, "attn": args.attn, "dtype": "bf16",
                    "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
                    "what": "unmodified reference CM3PModel (transformers ModernBERT, torch/cuBLAS/" +
                            ("flash-attn 2" if args.attn == "fa2" else "SDPA") + ") on one B200, padded L=2000"}
        except Exception as exc:  # noqa: BLE001
            line = {"impl": "reference-on-gpu", "workload": workload, "attn": args.attn,
                    "error": f"{type(exc).__name__}: {str(exc)[:300]}"}
        print(json.dumps(line), flush=True)
        model = None


if __name__ == "__main__":
    main()
