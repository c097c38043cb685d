"""Per-operator device time of the step kernel at several team sizes (profiling helper)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root (bench.py, sbayes_b200/, oracle/)
import sys, time, json
import numpy as np
import bench
from sbayes_b200.backend import DeviceContext
from sbayes_b200.packed import ChainConfig, OP_NAMES
import torch

def main():
    teams = [int(t) for t in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 4, 8, 16, 25, 37]
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    prob, states = bench.make_workload(0)
    for team in teams:
        ctx = DeviceContext(prob, 4, team=team)
        for c in range(4):
            ctx.configure(c, ChainConfig(temperature=1 + 0.05 * c, seed=c))
            ctx.set_state(c, states[c])
        ctx.run(200)
        a0, r0 = ctx.op_stats(0); c0 = ctx.op_cycles(0)
        torch.cuda.synchronize(); t = time.perf_counter()
        ctx.run(nsteps)
        dt = time.perf_counter() - t
        a1, r1 = ctx.op_stats(0); c1 = ctx.op_cycles(0)
        n = (a1 + r1) - (a0 + r0); cy = c1 - c0
        print(f"team={team} grid={ctx.info()['grid']}  {dt/nsteps*1e6:.1f} us/step/chain  {4*nsteps/dt:.0f} steps/s")
        for i in range(10):
            if n[i]:
                print(f"   {OP_NAMES[i]:36s} n={n[i]:5d} acc={(a1-a0)[i]/n[i]:.2f}  {cy[i]/n[i]/1.965e3:8.1f} us/step")
        ctx.close()

main()
