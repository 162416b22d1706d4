#!/usr/bin/env python
"""Vendors the reference's own inference classes into the git-ignored `baseline/_ref/` so that `bench.py --impl
reference` can time THEM (not a port) on the GPU box's host cores.

    python scripts/vendor_reference.py [--src /root/reference]

What is copied (verbatim, unmodified): `Net/`, `Config/`, `Util/Universal_Util/Utils.py` -- the files behind
`IMUNet`, `UpperNet`, `LowerNet` (+ `GCN.Model`) and `Transform2H/2R`, i.e. the chain of
Processor/Test/Demo_test.py:111-123.  Nothing else of the reference is needed to run that chain; the loader, the
trainers and the plotting code stay behind.

`baseline/_ref/` is listed in .gitignore (reference sources never enter this repository's history) but NOT in
.gpurunignore, so the directory travels to the GPU box with the snapshot, exactly like the built .so files.
/root/reference exists only in the build container: `__graft_entry__.build()` calls this script there; on the GPU box
the already vendored copy is used as is.  A MANIFEST with the sha256 of every copied file is written next to them.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
ITEMS = ["Net", "Config", os.path.join("Util", "Universal_Util", "Utils.py")]


def vendor(src: str = "/root/reference", quiet: bool = False) -> str | None:
    if not os.path.isdir(src):
        if not quiet:
            print(f"{src} not present: keeping whatever is already under {DST}")
        return DST if os.path.isdir(os.path.join(DST, "Net")) else None
    manifest = {}
    for item in ITEMS:
        s, d = os.path.join(src, item), os.path.join(DST, item)
        if os.path.isdir(s):
            if os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copy2(s, d)
    for base, _, files in os.walk(DST):
        for f in sorted(files):
            if f == "MANIFEST.json" or f.endswith(".pyc"):
                continue
            p = os.path.join(base, f)
            manifest[os.path.relpath(p, DST)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(dict(source=src, files=manifest), f, indent=1, sort_keys=True)
    if not quiet:
        print(f"vendored {len(manifest)} files from {src} into {DST}")
    return DST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    sys.exit(0 if vendor(a.src) else 1)
