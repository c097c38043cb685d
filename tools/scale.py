"""Exploratory timing of the step kernel and the full evaluation at the other BASELINE shapes
(profiling helper).  usage: tools_scale.py <case>[,<case>...]   cases: c3 c4 c4s c5q c5"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root (bench.py, sbayes_b200/, oracle/)
import sys, time
import numpy as np
import torch
from sbayes_b200 import synth
from sbayes_b200.backend import DeviceContext
from sbayes_b200.packed import ChainConfig

CASES = {
    "c3": dict(N=100, F=36, S=5, K=3, n_fam=6, seed=3, chains=8, steps=2000),
    "c4s": dict(N=5000, F=200, S=6, K=5, n_fam=20, seed=4, chains=8, steps=200),
    "c4": dict(N=5000, F=200, S=6, K=5, n_fam=20, seed=4, chains=128, steps=200),
    "c2": dict(N=1000, F=100, S=4, K=5, n_fam=8, seed=2, chains=4, steps=1000),
    "c5s": dict(N=25000, F=1000, S=8, K=16, n_fam=64, seed=5, chains=1, steps=20),
    "c5q": dict(N=50000, F=1000, S=8, K=16, n_fam=64, seed=5, chains=1, steps=20),
    "c5": dict(N=200000, F=1000, S=8, K=16, n_fam=64, seed=5, chains=1, steps=20),
}

def main():
    for name in sys.argv[1].split(","):
        w = CASES[name]
        t = time.time()
        prob, _ = synth.make_problem(N=w["N"], F=w["F"], S=w["S"], K=w["K"], n_fam=w["n_fam"], seed=w["seed"])
        nd = min(w["chains"], 4)
        states = [synth.initial_state(prob, seed=100 + c) for c in range(nd)]
        print(f"[{name}] host gen {time.time()-t:.1f}s", flush=True)
        try:
            ctx = DeviceContext(prob, w["chains"], team=0)
        except Exception as e:
            print(f"[{name}] create failed: {e}", flush=True); continue
        print(f"[{name}] info {ctx.info()}", flush=True)
        for c in range(w["chains"]):
            ctx.configure(c, ChainConfig(temperature=1.0, seed=c))
            ctx.set_state(c, states[c % nd])
        try:
            ctx.eval(0)
            torch.cuda.synchronize(); t = time.perf_counter()
            for _ in range(5): post = ctx.eval(0)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
            fe = 2 * prob.N * prob.F * w["chains"]
            print(f"[{name}] full eval of {w['chains']} chains: {dt*1e3:.3f} ms  -> {fe/dt/1e9:.1f} GB/s algorithmic, {w['chains']/dt:.0f} evals/s  post={post}", flush=True)
            ctx.eval_batch()
            ts = [ctx.eval_batch(timing=True) for _ in range(10)]
            h = min(t[1][0] for t in ts); tm = min(t[1][1] for t in ts)
            print(f"[{name}] statistics pass: {ctx.eval_info()}", flush=True)
            print(f"[{name}] eval_batch: stats {h*1e3:.1f} us ({fe/h/1e6:.1f} GB/s algorithmic) terms {tm*1e3:.1f} us  -> {w['chains']/((h+tm)*1e-3):.0f} evals/s  "
                  f"ll={ts[0][0][0]['ll']} prior={ts[0][0][0]['prior']}", flush=True)
            try:
                if "--no-mma" in sys.argv:
                    raise RuntimeError("skipped")
                ctx.eval_set_mode("mma")
                ctx.eval_batch()
                ts = [ctx.eval_batch(timing=True) for _ in range(10)]
                h = min(t[1][0] for t in ts); tm = min(t[1][1] for t in ts)
                print(f"[{name}] eval_batch (tcgen05): stats {h*1e3:.1f} us ({fe/h/1e6:.1f} GB/s algorithmic) terms {tm*1e3:.1f} us  -> {w['chains']/((h+tm)*1e-3):.0f} evals/s  "
                      f"ll={ts[0][0][0]['ll']} prior={ts[0][0][0]['prior']}", flush=True)
            except Exception as e:
                print(f"[{name}] mma failed: {e}", flush=True)
            ctx.eval_set_mode("hist")
            try:
                ctx.obs_loglik(want_lh=False, want_objects=False)
                tms = [ctx.obs_loglik(want_lh=False, want_objects=False)[3] for _ in range(8)]
                t = min(tms)
                print(f"[{name}] obs_loglik: {t*1e3:.1f} us  ({6*prob.N*prob.F*w['chains']/t/1e6:.1f} GB/s algorithmic at 6 B/obs, {prob.N*prob.F*w['chains']/t/1e6:.1f} G obs/s)", flush=True)
            except Exception as e:
                print(f"[{name}] obs_loglik failed: {e}", flush=True)
            try:
                ctx.gibbs_sweep(seed=1)
                tms = [ctx.gibbs_sweep(seed=2 + i) for i in range(6)]
                t = min(tms)
                print(f"[{name}] gibbs_sweep: {t*1e3:.1f} us  ({3*prob.N*prob.F*w['chains']/t/1e6:.1f} GB/s algorithmic at 3 B/obs)", flush=True)
            except Exception as e:
                print(f"[{name}] gibbs_sweep failed: {e}", flush=True)
            try:      # initializer EM: device ms per EM step against the float32 numpy einsum pair of the reference on this box's CPU
                from sbayes_b200 import initializer as INI
                from sbayes_b200.packed import NA
                z0 = INI.initial_soft_assignment(prob, np.random.default_rng(1))
                ctx.init_em(z0, 2)
                _, ms = ctx.init_em(z0, 10, timing=True)
                line = f"[{name}] init_em: {ms/10*1e3:.1f} us per EM step on the device (G_tot={prob.G_tot}, {prob.G_tot*prob.N*prob.F/1e6:.0f} M multiply-adds per contraction)"
                if prob.N * prob.F * prob.S <= 8e6:
                    X = np.zeros((prob.N, prob.F, prob.S), np.float32)
                    o, f = np.nonzero(prob.x != NA); X[o, f, prob.x[o, f]] = 1.0
                    t0 = time.perf_counter()
                    cnt = np.einsum("ij,jkl->ikl", z0, X, optimize=True)
                    p = np.log(cnt / cnt.sum(axis=-1, keepdims=True).clip(1e-30) + 1e-30)
                    ll = np.einsum("ikl,jkl->ij", p, X, optimize=True)
                    line += f"; numpy einsum pair of the reference (float32, this box's CPU) {(time.perf_counter()-t0)*1e3:.0f} ms"
                print(line, flush=True)
            except Exception as e:
                print(f"[{name}] init_em failed: {e}", flush=True)
            if "--no-steps" in sys.argv: ctx.close(); continue
            ctx.run(10)
            torch.cuda.synchronize(); t = time.perf_counter()
            ctx.run(w["steps"])
            torch.cuda.synchronize(); dt = time.perf_counter() - t
            print(f"[{name}] {w['steps']} steps x {w['chains']} chains: {dt/w['steps']*1e6:.1f} us/step  {w['chains']*w['steps']/dt:.0f} steps/s  "
                  f"-> {prob.N*prob.F*w['chains']*w['steps']/dt/1e9:.1f} GB/s algorithmic", flush=True)
            a, r = ctx.op_stats(0)
            print(f"[{name}] acc {a.tolist()} rej {r.tolist()}", flush=True)
        except Exception as e:
            print(f"[{name}] failed: {e}", flush=True)
        ctx.close()

if __name__ == "__main__":
    main()
