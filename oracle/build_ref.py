"""Recipe for oracle/_ref: the reference's own implementation of the path, made available to the GPU box.

TEST INFRASTRUCTURE ONLY.  The reference is pure Python; its hot path lives in two files that import nothing
else from its tree (`cm3p/__init__.py:1-2` imports only `configuration_cm3p` and `modeling_cm3p`).  This script
places UNMODIFIED copies of them (plus `utils/muon_utils.py`, the optimizer the goldens of `oracle/muon_oracle.py`
were made with) under `oracle/_ref/`, which is git-ignored (no reference source enters the history) but travels to
the GPU box with the repo snapshot, like the built `.so` files.  There the CPU baseline of `bench.py`
(`cpu_baseline.kind = "reference"`, `--impl reference`) and `tools/bench_reference_gpu.py` import it through
`oracle/ref_shim.py`.  Nothing under `cm3p_b200/` ever imports it.

    python oracle/build_ref.py            # needs /root/reference; `__graft_entry__.build()` calls it when present
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["cm3p/__init__.py", "cm3p/configuration_cm3p.py", "cm3p/modeling_cm3p.py", "utils/muon_utils.py"]


def build(reference_root: str = "/root/reference") -> str | None:
    if not os.path.isfile(os.path.join(reference_root, "cm3p", "modeling_cm3p.py")):
        return DEST if os.path.isfile(os.path.join(DEST, "cm3p", "modeling_cm3p.py")) else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(reference_root, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": reference_root, "sha256": manifest}, f, indent=1)
    return DEST


if __name__ == "__main__":
    out = build(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print(out or "reference checkout not found; oracle/_ref not built")
