"""Short run of the step kernel at the config-4 shape for an ncu capture (profiling helper).
usage: tools_profrun_c4.py <chains> <steps>"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root (bench.py, sbayes_b200/, oracle/)
import sys
from sbayes_b200 import synth
from sbayes_b200.backend import DeviceContext
from sbayes_b200.packed import ChainConfig
chains, nsteps = int(sys.argv[1]), int(sys.argv[2])
prob, _ = synth.make_problem(N=5000, F=200, S=6, K=5, n_fam=20, seed=4)
sts = [synth.initial_state(prob, seed=100 + c) for c in range(min(chains, 4))]
ctx = DeviceContext(prob, chains)
for c in range(chains):
    ctx.configure(c, ChainConfig(seed=c)); ctx.set_state(c, sts[c % len(sts)])
ctx.run(10)
ctx.run(nsteps)
print(ctx.info())
ctx.close()
