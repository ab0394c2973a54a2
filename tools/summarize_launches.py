"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]
--csv`): time share, launches and (when present) DRAM traffic per kernel.  `--second-half` keeps only the
launches of the last (timed) step when the command ran 1 warm-up + 1 timed step."""
import csv
import re
import sys
from collections import defaultdict

SCALE_T = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
SCALE_B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short_name(name):
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    m = re.search(r"gemm_bf16_sm100_kernel<(?:\(int\))?(\d+)", name)
    if m:
        short = f"cm3p::gemm_bf16_sm100_kernel<EPI={m.group(1)}>"
    return short.replace("<unnamed>::", "")


def main(path, second_half=False):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    per = defaultdict(dict)
    for r in csv.DictReader(lines):
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        if r["Metric Name"] == "gpu__time_duration.sum":
            per[int(r["ID"])].update(name=r["Kernel Name"], us=val * SCALE_T.get(unit, 1e-3))
        elif r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            d = per[int(r["ID"])]
            d["bytes"] = d.get("bytes", 0.0) + val * SCALE_B.get(unit, 1.0)
    ids = sorted(per)
    if second_half:
        # both steps launch the same number of cm3p kernels (the first one adds torch kernels: weight packing, caches):
        # the timed step starts at the middle cm3p launch
        ours = [i for i in ids if "cm3p::" in per[i].get("name", "")]
        ids = [i for i in ids if i >= ours[len(ours) // 2]] if ours else ids[len(ids) // 2:]
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for i in ids:
        d = per[i]
        a = agg[short_name(d["name"])]
        a[0] += 1
        a[1] += d["us"]
        a[2] += d.get("bytes", 0.0)
    total = sum(v[1] for v in agg.values())
    has_bytes = any(v[2] for v in agg.values())
    print(f"# {path}: {len(ids)} launches, {total / 1e3:.3f} ms summed kernel time (ncu: cold caches, serialised)")
    print(f"{'share':>7} {'ms':>9} {'count':>6} {'us/launch':>10}" + (f" {'DRAM MB/launch':>15} {'GB/s':>8}" if has_bytes else "")
          + "  kernel")
    for name, (cnt, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        extra = f" {by / cnt / 1e6:15.1f} {by / us / 1e3:8.0f}" if has_bytes else ""
        print(f"{100 * us / total:6.1f}% {us / 1e3:9.3f} {cnt:6d} {us / cnt:10.1f}{extra}  {name[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], "--second-half" in sys.argv[2:])
