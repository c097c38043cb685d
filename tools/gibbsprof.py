"""Short run of the full-sweep source Gibbs kernel for an ncu capture (profiling helper).
usage: tools/gibbsprof.py <c2|c4|c5s> [chains]"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root (bench.py, sbayes_b200/, oracle/)
import sys
from sbayes_b200 import synth
from sbayes_b200.backend import DeviceContext
from sbayes_b200.packed import ChainConfig
SHAPES = {"c2": dict(N=1000, F=100, S=4, K=5, n_fam=8, seed=2), "c4": dict(N=5000, F=200, S=6, K=5, n_fam=20, seed=4),
          "c5s": dict(N=25000, F=1000, S=8, K=16, n_fam=64, seed=5)}
case = sys.argv[1]
chains = int(sys.argv[2]) if len(sys.argv) > 2 else {"c2": 4, "c4": 128, "c5s": 1}[case]
prob, _ = synth.make_problem(**SHAPES[case])
sts = [synth.initial_state(prob, seed=100 + c) for c in range(min(chains, 4))]
ctx = DeviceContext(prob, chains)
for c in range(chains):
    ctx.configure(c, ChainConfig(seed=c)); ctx.set_state(c, sts[c % len(sts)])
for i in range(3):
    print(ctx.gibbs_sweep(step=1 + i))
ctx.close()
