"""Probe of the train step (fwd + explicit bwd) on one B200: ms/step, peak memory and a per-kernel
device-time table from torch.profiler (CUPTI), for a list of batch sizes.

    python tools/train_probe.py [--batches 32,64] [--variations 8] [--seq-len 2000] [--profile] [--infer]
"""
import argparse
import copy
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from cm3p_b200.configuration_cm3p import CM3PConfig, base_config_dict  # noqa: E402
from cm3p_b200.modeling_cm3p import CM3PModel  # noqa: E402
from cm3p_b200.synthetic import synthetic_batch, synthetic_state_dict  # noqa: E402


def kernel_table(prof, top=25):
    rows = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
            name = name.replace("void ", "").replace("cm3p::", "").replace("at::native::", "")
            name = name.split("(")[0][:70]
            r = rows.setdefault(name, [0.0, 0])
            r[0] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
            r[1] += 1
    tot = sum(v[0] for v in rows.values())
    out = []
    for name, (us, n) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:top]:
        out.append(f"  {100 * us / tot:5.1f}%  {us / 1e3:9.3f} ms  {n:5d}  {name}")
    return f"total device time {tot / 1e3:.3f} ms\n" + "\n".join(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="32")
    ap.add_argument("--variations", type=int, default=8)
    ap.add_argument("--seq-len", type=int, default=2000)
    ap.add_argument("--min-len", type=int, default=600)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--infer", action="store_true", help="probe the no-grad forward instead of the train step")
    ap.add_argument("--opt", action="append", default=[], help="library option KEY=VALUE (cm3p_set_option), repeatable")
    args = ap.parse_args()
    from cm3p_b200 import ops
    for kv in args.opt:
        k, v = kv.split("=")
        ops.set_option(int(k), int(v))
    dev = torch.device("cuda", 0)
    cfg = CM3PConfig(attn_implementation="flash_attention_2", **copy.deepcopy(base_config_dict()))
    model = CM3PModel(cfg)
    model.load_state_dict(synthetic_state_dict(cfg, seed=0), strict=True)
    model = model.to(dev)
    model = model.to(torch.bfloat16).eval() if args.infer else model.train()

    for B in [int(b) for b in args.batches.split(",")]:
        V = 1 if args.infer else args.variations
        batch = {k: v.to(dev) for k, v in synthetic_batch(cfg, batch=B, seq_len=args.seq_len, variations=V, seed=1,
                                                          min_len=args.min_len).items()}

        def step():
            if args.infer:
                with torch.no_grad():
                    return model(**batch, return_loss=False)
            model.zero_grad(set_to_none=True)
            out = model(**batch)
            out.loss.backward()
            return out

        torch.cuda.reset_peak_memory_stats()
        for _ in range(2):
            out = step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            out = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        wall = (time.perf_counter() - t0) * 1e3 / args.steps
        rec = {"mode": "infer" if args.infer else "train", "batch": B, "variations": V, "seq_len": args.seq_len,
               "tokens": int(batch["attention_mask"].sum()), "ms_per_step": round(ms, 2), "wall_ms": round(wall, 2),
               "per_s": round(B / ms * 1e3, 1), "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2),
               "loss": None if args.infer else float(out.loss.detach())}
        print(json.dumps(rec), flush=True)
        if args.profile:
            with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA,
                                                    torch.profiler.ProfilerActivity.CPU]) as prof:
                step()
                torch.cuda.synchronize()
            print(kernel_table(prof), flush=True)


if __name__ == "__main__":
    main()
