"""One-off extended lockstep of the device chains with the fp64 oracle (same Philox streams) at
the bench shape and at the config-3 shape (checker: tests/parity.py)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root (bench.py, sbayes_b200/, oracle/)
import sys, time
sys.path.insert(0, "tests")
import parity
from sbayes_b200 import synth
from sbayes_b200.packed import ChainConfig
t = time.time()
prob, _ = synth.make_problem(N=1000, F=100, S=4, K=5, n_fam=8, seed=2)
cfgs = [ChainConfig(temperature=1.0 + 0.05 * c, prior_temperature=1.0, seed=40 + c) for c in range(4)]
st0 = [synth.initial_state(prob, seed=50 + c) for c in range(4)]
nt = parity.philox_lockstep(prob, st0, cfgs, 1200, team=0)
print(f"c2: 4 chains x 1200 steps in lockstep with the oracle OK (near ties {nt}, {time.time()-t:.0f} s)", flush=True)
t = time.time()
prob, _ = synth.make_problem(N=100, F=36, S=5, K=3, n_fam=6, seed=3)
cfgs = [ChainConfig(temperature=1.0 + 0.05 * c, prior_temperature=1.0 + 0.03 * c, seed=60 + c) for c in range(8)]
st0 = [synth.initial_state(prob, seed=70 + c) for c in range(8)]
nt = parity.philox_lockstep(prob, st0, cfgs, 3000, team=0)
print(f"c3: 8 chains x 3000 steps in lockstep with the oracle OK (near ties {nt}, {time.time()-t:.0f} s)", flush=True)
