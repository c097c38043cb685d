"""Short run of the stand-alone evaluation kernels for an ncu capture (profiling helper).
usage: tools_evalprof.py <case> <hist|mma> [chains]"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root (bench.py, sbayes_b200/, oracle/)
import sys
from tools import scale as TS
from sbayes_b200 import synth
from sbayes_b200.backend import DeviceContext
from sbayes_b200.packed import ChainConfig

w = dict(TS.CASES[sys.argv[1]])
if len(sys.argv) > 3:
    w["chains"] = int(sys.argv[3])
prob, _ = synth.make_problem(N=w["N"], F=w["F"], S=w["S"], K=w["K"], n_fam=w["n_fam"], seed=w["seed"])
states = [synth.initial_state(prob, seed=100 + c) for c in range(min(w["chains"], 2))]
ctx = DeviceContext(prob, w["chains"])
for c in range(w["chains"]):
    ctx.configure(c, ChainConfig(seed=c))
    ctx.set_state(c, states[c % len(states)])
ctx.eval_set_mode(sys.argv[2])
for _ in range(3):
    r, ms = ctx.eval_batch(timing=True)
print(sys.argv[1], sys.argv[2], ms, r[0]["ll"])
ctx.close()
