"""GPU debugging helper: Philox lockstep against the oracle on small problems, reporting the first mismatch."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root (bench.py, sbayes_b200/, oracle/)
import sys
sys.path.insert(0, "tests")
import numpy as np
import parity
from sbayes_b200 import synth
from sbayes_b200.packed import ChainConfig

def run(N, F, S, K, nf, team, steps, seed=7):
    prob, _ = synth.make_problem(N=N, F=F, S=S, K=K, n_fam=nf, seed=seed, max_size=max(8, N // K))
    cfgs = [ChainConfig(seed=100), ChainConfig(temperature=1.2, prior_temperature=1.1, seed=101)]
    st0 = [synth.initial_state(prob, seed=5 + i, objects_per_cluster=5) for i in range(2)]
    try:
        parity.philox_lockstep(prob, st0, cfgs, steps, team=team)
        print(f"N={N} F={F} team={team}: OK", flush=True)
    except AssertionError as e:
        print(f"N={N} F={F} team={team}: FAIL {e}", flush=True)

if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "all"
    if mode == "tiny":
        run(40, 8, 3, 2, 3, 1, 30)
    else:
        for (N, F, S, K, nf) in [(40, 8, 3, 2, 3), (50, 12, 4, 3, 4), (200, 40, 4, 3, 4), (1000, 100, 4, 5, 8)]:
            for team in (1, 2, 4):
                run(N, F, S, K, nf, team, 60)
