"""Print the headline numbers of a bench.py JSON line (last line of the file)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])


def line(tag, r):
    print(tag, r["metric"], r["value"], r["unit"], "ms/step", r["ms_per_step"], "| e2e", r["e2e"]["value"],
          "| GEMM TF/s", r["roofline"]["achieved"], "frac", r["roofline"]["frac"], "| attn TF/s",
          r["roofline_attention"]["achieved"], "share", r["roofline_attention"]["share_of_step"], "| clocks",
          r["clocks"]["sm_mhz"], r["clocks"]["reasons"], "| launches", r["gpu_launches"])


line("infer", d)
if "train" in d:
    line("train", d["train"])
    print("train peak_mem_gb", d["train"].get("peak_mem_gb"))
