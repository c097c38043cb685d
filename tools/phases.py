"""Per-phase cycle breakdown of the step kernel (needs a library built with SB_CFLAGS=-DSB_PHASE_PROF).
usage: tools_phases.py <team> <steps> [c2|c4] [chains]"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root (bench.py, sbayes_b200/, oracle/)
import ctypes as C, sys
import numpy as np
import bench
from sbayes_b200.backend import DeviceContext
from sbayes_b200.packed import ChainConfig, OP_NAMES

NPH = 48
NAMES = {1: "op choice", 2: "grow_candidates", 20: "build_tables", 21: "pass (batch)", 4: "team_pick", 5: "cand sizes/pre-open",
         6: "proposal_open", 7: "R1 remove", 8: "R2 draw", 9: "R3 add+sum", 10: "proposal_rows", 11: "geo", 13: "bwd team_sum",
         14: "subset", 15: "G1", 16: "G delta+sum", 17: "G2", 18: "jump fwd", 19: "jump bwd+sum", 22: "weights loop",
         23: "weight_tables", 24: "sp+wp", 32: "tile build", 33: "tile pass", 34: "ratio rows", 29: "op tail", 30: "posterior+MH", 31: "commit/close"}
def main():
    team, nsteps = int(sys.argv[1]), int(sys.argv[2])
    shape = sys.argv[3] if len(sys.argv) > 3 else "c2"
    if shape == "c4":      # config 4: 5000 x 200 x 6, many chains per GPU
        from sbayes_b200 import synth
        nch = int(sys.argv[4]) if len(sys.argv) > 4 else 128
        prob, _ = synth.make_problem(N=5000, F=200, S=6, K=5, n_fam=20, seed=4)
        states = [synth.initial_state(prob, seed=100 + c) for c in range(4)]
    else:
        nch = 4
        prob, states = bench.make_workload(0)
    ctx = DeviceContext(prob, nch, team=team)
    for c in range(nch):
        ctx.configure(c, ChainConfig(temperature=1 + 0.05 * (c % 4) if shape == "c2" else 1.0, seed=c))
        ctx.set_state(c, states[c % len(states)])
    ctx.run(100)
    ph, n = np.zeros((10, NPH), np.int64), np.zeros(10, np.int64)
    ctx.lib.sb_debug_phases.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    ctx.lib.sb_debug_phases(ctx.h, ph.ctypes.data, n.ctypes.data, 1)
    ctx.run(nsteps)
    ctx.lib.sb_debug_phases(ctx.h, ph.ctypes.data, n.ctypes.data, 0)
    print(f"team={team} info={ctx.info()}")
    for op in range(10):
        if not n[op]:
            continue
        tot = ph[op].sum() / n[op]
        print(f"{OP_NAMES[op]:34s} n={n[op]:6d}  {tot/1965:8.1f} us/step")
        for i in range(NPH):
            if ph[op, i]:
                print(f"      {NAMES.get(i, str(i)):22s} {ph[op, i]/n[op]/1965:8.2f} us")
    ctx.close()


if __name__ == "__main__":
    main()
