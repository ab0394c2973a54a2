"""Top instructions of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K` (SASS view):
stall samples with their dominant reasons, shared-memory excess wavefronts.  Optional second argument: an
`nvdisasm -g` listing of the same function, used to print the source line of each instruction (matched by order)."""
import csv
import re
import sys


def main(path, disasm=None, top=40):
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body = rows[hdr_i + 1:]
    for j, r in enumerate(body):  # several launches of the kernel in the report: keep the first
        if not r or r[0] in ("Kernel Name", "Address") or len(r) < len(hdr):
            body = body[:j]
            break
    col = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    lines = {}
    if disasm:
        cur, k = None, 0
        for ln in open(disasm):
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = f"{m.group(1).split('/')[-1]}:{m.group(2)}"
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
                lines[k] = cur
                k += 1
    total = sum(int(r[col["# Samples"]] or 0) for r in body)
    print(f"# {path}: {len(body)} instructions, {total} samples")
    by_line = {}
    order = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))
    for i in order[:top]:
        r = body[i]
        n = int(r[col["# Samples"]] or 0)
        why = sorted(((int(r[col[s]] or 0), s[6:]) for s in stalls), reverse=True)[:3]
        why = " ".join(f"{s}={v}" for v, s in why if v)
        exc = r[col["L1 Wavefronts Shared Excessive"]]
        print(f"{100.0 * n / total:5.1f}% {n:6d} [{i:5d}] {lines.get(i, ''):28s} {r[col['Source']].strip()[:70]:70s} {why}  exc={exc}")
    if lines:
        for i, r in enumerate(body):
            key = lines.get(i)
            by_line[key] = by_line.get(key, 0) + int(r[col["# Samples"]] or 0)
        print("# by source line")
        for key, n in sorted(by_line.items(), key=lambda kv: -kv[1])[:30]:
            print(f"{100.0 * n / total:5.1f}% {n:6d} {key}")
    exc_rows = sorted(body, key=lambda r: -int(r[col["L1 Wavefronts Shared Excessive"]] or 0))[:8]
    print("# shared-memory excess wavefronts")
    for r in exc_rows:
        print(r[col["L1 Wavefronts Shared Excessive"]], r[col["L1 Wavefronts Shared"]], r[col["Source"]].strip()[:90])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
