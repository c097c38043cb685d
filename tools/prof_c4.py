"""Per-operator device time of the step kernel at the c4 shape (profiling helper).
usage: tools_prof_c4.py <chains> <steps> [team]"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root (bench.py, sbayes_b200/, oracle/)
import sys, time
import numpy as np, torch
from sbayes_b200 import synth
from sbayes_b200.backend import DeviceContext
from sbayes_b200.packed import ChainConfig, OP_NAMES
chains, nsteps = int(sys.argv[1]), int(sys.argv[2])
team = int(sys.argv[3]) if len(sys.argv) > 3 else 0
prob, _ = synth.make_problem(N=5000, F=200, S=6, K=5, n_fam=20, seed=4)
sts = [synth.initial_state(prob, seed=100 + c) for c in range(min(chains, 4))]
ctx = DeviceContext(prob, chains, team=team)
for c in range(chains):
    ctx.configure(c, ChainConfig(seed=c)); ctx.set_state(c, sts[c % len(sts)])
ctx.run(20)
a0, r0 = ctx.op_stats(0); c0 = ctx.op_cycles(0)
torch.cuda.synchronize(); t = time.perf_counter()
ctx.run(nsteps)
dt = time.perf_counter() - t
a1, r1 = ctx.op_stats(0); c1 = ctx.op_cycles(0)
n = (a1 + r1) - (a0 + r0); cy = c1 - c0
print(f"chains={chains} info={ctx.info()}  {dt/nsteps*1e6:.1f} us/step  {chains*nsteps/dt:.0f} steps/s")
for i in range(10):
    if n[i]:
        print(f"   {OP_NAMES[i]:36s} n={n[i]:5d} acc={(a1-a0)[i]/n[i]:.2f}  {cy[i]/n[i]/1.965e3:8.1f} us/step")
ctx.close()
