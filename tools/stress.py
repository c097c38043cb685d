"""Long-run invariants of the step kernel (stress helper): after many steps the incrementally
maintained counts equal a bulk rebuild and the cached posterior equals the uncached evaluation,
for several shapes / team sizes.  usage: tools_stress.py"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root (bench.py, sbayes_b200/, oracle/)
import numpy as np
from oracle import model as M
from sbayes_b200 import synth
from sbayes_b200.backend import DeviceContext
from sbayes_b200.packed import ChainConfig

def rel(a, b): return abs(a - b) / max(1.0, abs(a))

CASES = [("c2 team16", dict(N=1000, F=100, S=4, K=5, n_fam=8, seed=2), 4, 0, 100000),
         ("c2 team3", dict(N=1000, F=100, S=4, K=5, n_fam=8, seed=2), 4, 3, 20000),
         ("c3", dict(N=100, F=36, S=5, K=3, n_fam=6, seed=3), 8, 0, 100000),
         ("c4 team1", dict(N=5000, F=200, S=6, K=5, n_fam=20, seed=4), 64, 1, 1500),
         ("c4 team18", dict(N=5000, F=200, S=6, K=5, n_fam=20, seed=4), 8, 0, 3000)]
for name, shape, chains, team, steps in CASES:
    prob, _ = synth.make_problem(**shape)
    ctx = DeviceContext(prob, chains, team=team)
    for c in range(chains):
        ctx.configure(c, ChainConfig(temperature=1.0 + 0.05 * (c % 4), prior_temperature=1.0 + 0.02 * (c % 3), seed=c))
        ctx.set_state(c, synth.initial_state(prob, seed=c % 4))
    ctx.run(steps)
    worst = 0.0
    for c in range(min(chains, 8)):
        st = ctx.get_state(c)
        assert np.array_equal(st.counts, M.rebuild_counts(prob, st)), (name, c, "counts")
        assert np.all((st.src == 255) == (prob.x == 255)), (name, c, "src NA pattern")
        post = ctx.posterior(c)
        ll, pr = M.log_likelihood(prob, st), M.log_prior(prob, st)
        worst = max(worst, rel(ll, post["ll"]), rel(pr, post["prior"]))
        assert rel(ll, post["ll"]) < 1e-10 and rel(pr, post["prior"]) < 1e-10, (name, c, ll, post)
        a, r = ctx.op_stats(c)
        assert int(a.sum() + r.sum()) == steps
    ev = ctx.eval_batch()
    for c in range(min(chains, 8)):
        post = ctx.posterior(c)
        assert rel(ev[c]["ll"], post["ll"]) < 1e-10 and rel(ev[c]["prior"], post["prior"]) < 1e-10
    print(f"{name}: {steps} steps x {chains} chains OK  info={ctx.info()}  worst rel err vs oracle {worst:.2e}  acc rate {float(a.sum())/steps:.3f}", flush=True)
    ctx.close()
print("STRESS OK")
