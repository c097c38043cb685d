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


# translation units and the headers each one sees: an object file is rebuilt only when one of them changed, so an
# edit of the stand-alone kernels (seconds) does not recompile the step kernel (minutes)
COMMON = ["sb_platform.cuh", "sb_types.h", "sb_plans.h", "sb_philox.cuh"]
UNITS = {
    "sb_api.cu": COMMON + ["sb_ctx.h", "sb_core.cuh"],                     # the 4-component step kernel + the C ABI
    "sb_uc8.cu": COMMON + ["sb_core.cuh"],                                  # the 8-component step kernel
    "sb_uc3.cu": COMMON + ["sb_core.cuh"],                                  # the step kernel for exactly 3 components
    "sb_api_eval.cu": COMMON + ["sb_ctx.h", "sb_eval.cuh", "sb_counts.cuh", "sb_init.cuh"],   # the streaming kernels
}


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))) + \
        [os.path.join(ROOT, "include", "sbayes_b200.h")]


def _newer(target, deps):
    return os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(d) for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    import fcntl
    with open(os.path.join(CSRC, ".build.lock"), "w") as lock:      # one build at a time (tests and tools may race)
        fcntl.flock(lock, fcntl.LOCK_EX)
        return _build(force, verbose)


def _build(force: bool, verbose: bool) -> str:
    if not force and _newer(OUT, sources()):
        return OUT
    extra = os.environ.get("SB_CFLAGS", "").split()
    common = [NVCC] + [f for f in FLAGS if f != "-shared"] + extra
    objs, procs = [], []
    tag = os.path.splitext(os.path.basename(OUT))[0]
    flagfile = os.path.join(CSRC, f"{tag}.flags")
    flags_now = " ".join(common)
    same_flags = os.path.exists(flagfile) and open(flagfile).read() == flags_now
    for src, hdrs in UNITS.items():
        obj = os.path.join(CSRC, f"{tag}.{src[:-3]}.o")
        objs.append(obj)
        deps = [os.path.join(CSRC, src), os.path.join(ROOT, "include", "sbayes_b200.h")] + [os.path.join(CSRC, h) for h in hdrs]
        if not force and same_flags and _newer(obj, deps):
            continue
        cmd = common + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    with open(flagfile, "w") as f:
        f.write(flags_now)
    link = [NVCC, "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-lcudart", "-o", OUT]
    r = subprocess.run(link, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc link failed: " + " ".join(link))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
