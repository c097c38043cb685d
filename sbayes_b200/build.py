"""Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# SB_LIB_OUT: profiling variants (e.g. SB_CFLAGS=-DSB_PHASE_PROF) are built beside the product library
OUT = os.environ.get("SB_LIB_OUT") or os.path.join(CSRC, "libsbayes_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-shared", "-I" + os.path.join(ROOT, "include")]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))) + \
        [os.path.join(ROOT, "include", "sbayes_b200.h")]


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in sources()):
        return OUT
    # two translation units (sb_api.cu with the 4-component step kernel, sb_uc8.cu with the 8-component one),
    # compiled in parallel, then linked into one shared library
    common = [NVCC] + [f for f in FLAGS if f != "-shared"] + os.environ.get("SB_CFLAGS", "").split() + (["-Xptxas", "-v"] if verbose else [])
    objs, procs = [], []
    tag = os.path.splitext(os.path.basename(OUT))[0]
    for src in ("sb_api.cu", "sb_uc8.cu"):
        obj = os.path.join(CSRC, f"{tag}.{src[:-3]}.o")
        objs.append(obj)
        cmd = common + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [NVCC, "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-lcudart", "-o", OUT]
    r = subprocess.run(link, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc link failed: " + " ".join(link))
    for o in objs:
        try:
            os.remove(o)
        except OSError:
            pass
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
