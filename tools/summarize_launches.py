"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import csv
import re
import sys
from collections import defaultdict


def main(path, skip_until_id=0):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((int(r["ID"]), r["Kernel Name"], val * scale))
    rows = [r for r in rows if r[0] >= skip_until_id]
    agg = defaultdict(lambda: [0, 0.0])
    for _, name, us in rows:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)
        m = re.search(r"gemm_bf16_sm100_kernel<(?:\(int\))?(\d+)>", name)
        if m:
            short = f"cm3p::gemm_bf16_sm100_kernel<EPI={m.group(1)}>"
        agg[short][0] += 1
        agg[short][1] += us
    total = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {total/1e3:.3f} ms summed kernel time (ncu: cold caches, serialised)")
    print(f"{'share':>7} {'ms':>9} {'count':>6} {'us/launch':>10}  kernel")
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"{100*us/total:6.1f}% {us/1e3:9.3f} {cnt:6d} {us/cnt:10.1f}  {name[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
