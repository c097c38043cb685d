"""Build the fibre-emulator flavour of the device code (TEST INFRASTRUCTURE, see hostsim.h):
the same sbayes_b200/csrc sources compiled with g++ -DSB_HOSTSIM into tests/_hostsim/."""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "tests", "_hostsim", "libsbayes_hostsim.so")
SRC = os.path.join(ROOT, "sbayes_b200", "csrc")


def build(force: bool = False) -> str:
    srcs = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs += [os.path.join(ROOT, "tests", "hostsim", "hostsim.h"), os.path.join(ROOT, "include", "sbayes_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-DSB_HOSTSIM", "-x", "c++",
           "-I" + os.path.join(ROOT, "tests", "hostsim"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(SRC, "sb_api.cu"), os.path.join(SRC, "sb_uc8.cu"), os.path.join(SRC, "sb_uc3.cu"), os.path.join(SRC, "sb_api_eval.cu"), "-o", OUT]
    subprocess.run(cmd, check=True)
    return OUT


def load():
    from sbayes_b200 import backend
    return backend._bind(ctypes.CDLL(build()))
