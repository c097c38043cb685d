#!/usr/bin/env python
"""Recipe that makes the live sBayes reference travel to the GPU box.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference is pure Python:
there is nothing to compile.  This script places an UNMODIFIED copy of the reference's
`sbayes/` package, its `test/test_files/` and the `experiments/south_america` inputs under
`oracle/_ref/` -- a directory that is git-ignored (the reference's sources never enter this
repository's history) but is NOT gpurun-ignored, so it ships to the GPU box with the built
`.so` files.  `oracle/refharness.py` imports the reference from `/root/reference` when that
exists (the build container) and from `oracle/_ref/` otherwise (the GPU box), always together
with the five stub modules of `oracle/refstubs/`.

What it enables on the B200: `bench.py --impl reference` / `cpu_baseline` timing the reference's
own `MCMCChain.run` (kind "reference"), the `ref`-marked drop-in tests, and the YAML
end-to-end run of `sbayes.cli.main` with the device backend installed.

    python oracle/make_ref.py            # called by __graft_entry__.build()
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SBAYES_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
PARTS = ["sbayes", os.path.join("test", "test_files"), os.path.join("experiments", "south_america")]
SKIP = shutil.ignore_patterns("__pycache__", "*.pyc", "*.pdf", "*.png", "*.jpg")


def _same_tree(a: str, b: str) -> bool:
    if not os.path.isdir(b):
        return False
    cmp = filecmp.dircmp(a, b, ignore=["__pycache__"])

    def ok(c):
        if c.left_only and any(not n.endswith((".pyc", ".pdf", ".png", ".jpg")) for n in c.left_only):
            return False
        if c.diff_files or c.funny_files:
            return False
        return all(ok(s) for s in c.subdirs.values())
    return ok(cmp)


def make(verbose: bool = False) -> str | None:
    """Copy the reference into oracle/_ref (idempotent).  Returns the destination, or None
    when the reference is not present (GPU box: the prebuilt copy is used as it is)."""
    if not os.path.isdir(os.path.join(SRC, "sbayes")):
        return DST if os.path.isdir(os.path.join(DST, "sbayes")) else None
    os.makedirs(DST, exist_ok=True)
    for part in PARTS:
        s, d = os.path.join(SRC, part), os.path.join(DST, part)
        if not os.path.exists(s):
            continue
        if _same_tree(s, d):
            continue
        if os.path.exists(d):
            shutil.rmtree(d)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copytree(s, d, ignore=SKIP)
        if verbose:
            print(f"oracle/_ref: copied {part}")
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Unmodified copy of the sBayes reference made by oracle/make_ref.py (git-ignored; test infrastructure).\n")
    return DST


if __name__ == "__main__":
    out = make(verbose=True)
    print(out if out else "reference not available", file=sys.stderr if not out else sys.stdout)
