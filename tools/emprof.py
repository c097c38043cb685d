"""Short run of the initializer EM for an ncu launch list (profiling helper).  usage: tools/emprof.py <case> [steps]"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import sys
import numpy as np
from tools import scale as TS
from sbayes_b200 import synth, initializer as INI
from sbayes_b200.backend import DeviceContext

w = dict(TS.CASES[sys.argv[1]])
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prob, _ = synth.make_problem(N=w["N"], F=w["F"], S=w["S"], K=w["K"], n_fam=w["n_fam"], seed=w["seed"])
ctx = DeviceContext(prob, 1)
z0 = INI.initial_soft_assignment(prob, np.random.default_rng(1))
z, ms = ctx.init_em(z0, steps, timing=True)
print(sys.argv[1], f"{ms / steps * 1e3:.1f} us per EM step")
ctx.close()
