"""Short run of the step kernel for an ncu capture (profiling helper).
usage: tools_profrun.py <team> <steps> [chains]"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root (bench.py, sbayes_b200/, oracle/)
import sys
import bench
from sbayes_b200.backend import DeviceContext
from sbayes_b200.packed import ChainConfig

team, nsteps = int(sys.argv[1]), int(sys.argv[2])
chains = int(sys.argv[3]) if len(sys.argv) > 3 else 4
prob, states = bench.make_workload(0)
ctx = DeviceContext(prob, chains, team=team)
for c in range(chains):
    ctx.configure(c, ChainConfig(temperature=1 + 0.05 * c, seed=c))
    ctx.set_state(c, states[c % len(states)])
ctx.run(50)
ctx.run(nsteps)
print(ctx.info())
ctx.close()
