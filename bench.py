#!/usr/bin/env python
"""bench.py -- MCMC steps/s and likelihood evals/s of the sBayes inner loop on B200 (BASELINE.json metric).

    python bench.py [--config c2|c3|c4|c5] [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Configurations (BASELINE.json `configs`; all data synthetic):

  c2 (default, the configuration the metric is quoted on)  1,000 x 100 x 4, K=5, family(8) + universal,
      MC3 with 4 tempered chains per GPU (temperature_diff 0.05, swap_interval 1000), operators 70/10/20.
      One bench step = one MC3 round: swap_interval MCMC steps on every chain of the rank + the swap.
      N GPUs: a ladder of 4N chains, interleaved over the ranks, swap scalars all-gathered over NCCL (weak).
  c3  south_america shape 100 x 36 x 5, K=3, 6 families + isolates, 8 tempered chains in total, 8/N per GPU
      (strong scaling: "one per GPU" at N=8).
  c4  replicates 5,000 x 200 x 6, K=5, 20 families, 128 independent chains per GPU (1,024 over 8 GPUs, weak);
      one bench step = 200 MCMC steps on every chain.
  c5  200,000 x 1,000 x 8, K=16, 64 families, objects sharded 25,000 per GPU (weak): ONE chain advanced by all
      GPUs together (team barriers / reductions / statistics cross GPUs inside the step kernel through NVLink
      peer memory); one bench step = 100 MCMC steps.  c5e: the uncached evaluation of the same matrix
      (statistics pass + int32 all-reduce + terms + fp64 all-reduce); value = evaluations/s.

`--impl reference` times the reference's own CPU implementation (the unmodified sBayes copy of
oracle/make_ref.py, `MCMCChain.run`, one process per chain like its MC3) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "c2": dict(N=1000, F=100, S=4, K=5, n_fam=8, seed=2, kind="mc3", chains_per_gpu=4, dT=0.05, swap_interval=1000, scaling="weak"),
    "c3": dict(N=100, F=36, S=5, K=3, n_fam=6, seed=3, kind="mc3", chains_total=8, dT=0.05, swap_interval=1000, scaling="strong"),
    "c4": dict(N=5000, F=200, S=6, K=5, n_fam=20, seed=4, kind="replicates", chains_per_gpu=128, steps_per_launch=200, scaling="weak"),
    "c5": dict(N=25000, F=1000, S=8, K=16, n_fam=64, seed=5, kind="sharded_steps", chains_per_gpu=1, steps_per_launch=100, scaling="weak"),
    "c5e": dict(N=25000, F=1000, S=8, K=16, n_fam=64, seed=5, kind="sharded_eval", chains_per_gpu=1, scaling="weak"),
}
METRIC = "mcmc_steps_per_sec"
UNIT = "steps/s"


def chains_on_rank(w: dict, world: int) -> int:
    if "chains_total" in w:
        if w["chains_total"] % world:
            raise SystemExit(f"{w['chains_total']} chains do not divide over {world} GPUs")
        return w["chains_total"] // world
    return w["chains_per_gpu"]


def make_problem(w: dict, rank: int = 0):
    from sbayes_b200 import synth
    sharded = w["kind"] in ("sharded_eval", "sharded_steps")
    seed = w["seed"] + (rank if sharded else 0)        # c5: every rank generates its own block of objects
    return synth.make_problem(N=w["N"], F=w["F"], S=w["S"], K=w["K"], n_fam=w["n_fam"], seed=seed, with_adjacency=not sharded)


def make_states(prob, n: int, first: int):
    from sbayes_b200 import synth
    distinct = min(n, 8)                      # initial states are cheap to share: the chains differ by their Philox keys
    sts = [synth.initial_state(prob, seed=1000 + first + c) for c in range(distinct)]
    return [sts[c % distinct] for c in range(n)]


def make_workload(rank: int = 0, name: str = "c2"):
    """(problem, initial states of the rank's chains) of a configuration (used by the tools/ helpers)."""
    w = CONFIGS[name]
    prob, _ = make_problem(w, rank)
    return prob, make_states(prob, chains_on_rank(w, 1), 0)


def step_bytes(prob) -> int:
    """Algorithmic HBM bytes of one MCMC step, SURVEY.md section 8(d): STEP ~ N*F (one pass
    over the packed observations of the candidate objects; tables and weights are O(G*F*S))."""
    return prob.N * prob.F


def workload_text(name: str, w: dict, world: int) -> str:
    base = f"{name}: synthetic {w['N']}x{w['F']}x{w['S']}, K={w['K']}, family({w['n_fam']})+universal"
    if w["kind"] == "mc3":
        return base + f", MC3 {chains_on_rank(w, world)} chains/GPU ({chains_on_rank(w, world) * world} in the ladder), swap_interval {w['swap_interval']}, operators 70/10/20"
    if w["kind"] == "replicates":
        return base + f", {w['chains_per_gpu']} independent chains/GPU, {w['steps_per_launch']} MCMC steps per bench step, operators 70/10/20"
    if w["kind"] == "sharded_steps":
        return base + (f" per GPU: ONE chain over {w['N'] * world} objects sharded over {world} GPU(s), {w['steps_per_launch']} MCMC steps per "
                       f"bench step, operators 70/10/20")
    return base + f" per GPU ({w['N'] * world} objects sharded over {world} GPU(s)), uncached likelihood + prior evaluation"


def config_dict(name: str, w: dict, world: int) -> dict:
    n_local = chains_on_rank(w, world)
    per = w.get("swap_interval", w.get("steps_per_launch", 1))
    return {"workload": workload_text(name, w, world), "config": name,
            "mcmc_steps_per_bench_step": {"sharded_eval": 0, "sharded_steps": per}.get(w["kind"], per * n_local * world),
            "chains": 1 if w["kind"].startswith("sharded") else n_local * world,
            "parallelism": {"mc3": f"chains x{world} (temperature ladder interleaved over ranks, swap scalars over NCCL all_gather)",
                            "replicates": f"chains x{world} (no collective)",
                            "sharded_steps": f"objects x{world} (team barriers, reductions and count statistics through NVLink peer memory inside the step kernel)",
                            "sharded_eval": f"objects x{world} (int32 all-reduce of counts, fp64 all-reduce of the source prior)"}[w["kind"]],
            "l2": "flushed between timed bench steps (256 MiB write); the step kernel's state is L2-resident inside a step by design"}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ==========================================================================================
# reference arm: the reference's own CPU implementation (oracle/_ref), one process per chain
# ==========================================================================================
def _ref_worker(conn, w, chain_ids):
    """Child process: builds the reference Data / Model of the workload and one MCMCChain per
    assigned chain; then serves ("run", n_steps) requests and answers the seconds it took."""
    try:
        import warnings
        warnings.filterwarnings("ignore")
        from collections import OrderedDict
        from oracle import refharness as H, model as M
        from sbayes_b200 import synth
        r = H.ref()
        prob0, truth = make_problem(w)
        confs = OrderedDict(universal=np.ones((1, prob0.N), bool))
        if w["n_fam"]:
            confs["family"] = np.stack([prob0.conf_gid[1] == g for g in range(prob0.conf_ngroups[1])])
        ce = dict(universal={"<ALL>": dict(type="uniform")})
        if w["n_fam"]:
            ce["family"] = {"<DEFAULT>": dict(type="uniform")}
        prior_cfg = dict(objects_per_cluster=dict(type="uniform_area", min=prob0.min_size, max=prob0.max_size),
                         geo=dict(type="uniform"), weights=dict(type="uniform"), cluster_effect=dict(type="uniform"),
                         confounding_effects=ce)
        data, model = H.build_reference_model(prob0.x, prob0.applicable, truth["locations"], confs, prob0.K, prior_cfg)
        prob = H.problem_from_reference(data, model)
        r.mcmc_chain.MCMCChain.CHECK_CACHING = False
        chains, samples = [], []
        for c in chain_ids:
            H.seed_reference(1000 + c)
            T = 1.0 + w.get("dT", 0.0) * c if w["kind"] == "mc3" else 1.0
            ch = r.mcmc_chain.MCMCChain(model=model, data=data, operators=r.config.OperatorsConfig(clusters=70, weights=10, source=20),
                                        sample_loggers=[], temperature=T, prior_temperature=1.0)
            st = synth.initial_state(prob, seed=1000 + c % 8)
            st.counts = M.rebuild_counts(prob, st)
            s = H.sample_from_state(prob, model, data, st)
            s = ch.run(n_steps=20, logging_interval=10 ** 9, initial_sample=s)      # numba JIT + caches warm
            chains.append(ch); samples.append(s)
        conn.send(("ready", len(chain_ids)))
        while True:
            msg = conn.recv()
            if msg[0] == "stop":
                break
            t0 = time.perf_counter()
            for i, ch in enumerate(chains):
                samples[i] = ch.run(n_steps=msg[1], logging_interval=10 ** 9, initial_sample=samples[i])
            conn.send(("done", time.perf_counter() - t0))
    except Exception as e:            # report instead of leaving the parent waiting
        import traceback
        conn.send(("error", f"{e!r}\n{traceback.format_exc()}"))


class ReferenceChains:
    """`n_chains` reference chains on min(n_chains, host cores) worker processes."""

    def __init__(self, w: dict, n_chains: int):
        import multiprocessing as mp
        self.n_chains = n_chains
        self.cores = max(1, min(n_chains, os.cpu_count() or 1))
        ctx = mp.get_context("fork")
        self.procs, self.conns = [], []
        for p in range(self.cores):
            ids = list(range(p, n_chains, self.cores))
            a, b = ctx.Pipe()
            pr = ctx.Process(target=_ref_worker, args=(b, w, ids), daemon=True)
            pr.start()
            self.procs.append(pr); self.conns.append(a)
        for c in self.conns:
            m = c.recv()
            if m[0] != "ready":
                raise RuntimeError("reference worker failed: " + str(m[1]))

    def run(self, n_steps: int) -> float:
        """every chain advances n_steps; returns the wall seconds of the slowest worker"""
        t0 = time.perf_counter()
        for c in self.conns:
            c.send(("run", n_steps))
        for c in self.conns:
            m = c.recv()
            if m[0] != "done":
                raise RuntimeError("reference worker failed: " + str(m[1]))
        return time.perf_counter() - t0

    def close(self):
        for c in self.conns:
            try:
                c.send(("stop",))
            except Exception:
                pass
        for p in self.procs:
            p.join(5)


def reference_available() -> bool:
    try:
        from oracle import refharness as H
        return H.reference_available()
    except Exception:
        return False


def port_steps_per_sec(w, n_chains, budget_s):
    """Fallback when the reference copy is absent: the oracle port (numpy restatement), one process per chain."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    prob, _ = make_problem(w)
    states = make_states(prob, n_chains, 0)
    cores = max(1, min(n_chains, os.cpu_count() or 1))

    def work(c):
        from oracle import chain as CH
        from oracle.draws import PhiloxDraws
        from sbayes_b200.packed import ChainConfig
        cfg = ChainConfig(temperature=1.0 + w.get("dT", 0.0) * c, seed=c)
        oc = CH.OracleChain(prob, cfg, states[c].copy(), draws=PhiloxDraws(c))
        for _ in range(3):
            oc.step()
        t0, n = time.perf_counter(), 0
        while time.perf_counter() - t0 < budget_s:
            oc.step(); n += 1
        q.put((n, time.perf_counter() - t0))
    procs = [ctx.Process(target=work, args=(c,)) for c in range(cores)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    return sum(n / t for n, t in res), sum(n for n, _ in res), cores


def cpu_baseline(name: str, w: dict, n_chains: int, budget_s: float = 15.0) -> dict:
    """The reference's CPU path on this box's host cores, on a bounded sample of the same workload."""
    if w["kind"] == "sharded_eval":
        return cpu_eval_baseline(w, budget_s)
    if w["kind"] == "sharded_steps":
        return cpu_sharded_steps_baseline(w, w["N"], budget_s)
    if reference_available():
        rc = ReferenceChains(w, n_chains)
        try:
            probe = rc.run(5)
            per_step = max(probe / 5, 1e-4)
            n = int(max(5, min(400, budget_s / per_step)))
            dt = rc.run(n)
            value = rc.n_chains * n / dt
            return {"value": value, "unit": UNIT, "cores": rc.cores, "kind": "reference",
                    "sample": f"{n} MCMC steps on each of {rc.n_chains} chains of the same workload ({dt:.1f} s), the reference's own "
                              f"MCMCChain.run (sbayes v1.1, unmodified copy of oracle/make_ref.py), one process per chain, CHECK_CACHING off"}
        finally:
            rc.close()
    v, n, cores = port_steps_per_sec(w, n_chains, budget_s)
    return {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} MCMC steps of the same workload: {cores} chains x {budget_s:.0f} s, oracle port (numpy), one process per chain"}


def cpu_eval_baseline(w: dict, budget_s: float) -> dict:
    """c5: the reference cannot construct this shape (O(N^2) distance matrix) nor run its cluster
    operators at F = 1000 (F6); the timed CPU path is the oracle port of Likelihood(caching=False) +
    Prior(caching=False) on a 2,000-object block of the same matrix, scaled linearly in N (flagged)."""
    from oracle import model as M
    from sbayes_b200 import synth
    n_small = 2000
    prob, _ = synth.make_problem(N=n_small, F=w["F"], S=w["S"], K=w["K"], n_fam=w["n_fam"], seed=w["seed"], with_adjacency=False)
    st = synth.initial_state(prob, seed=1000)
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < budget_s and n < 50:
        st.counts = M.rebuild_counts(prob, st)
        M.log_likelihood(prob, st); M.log_prior(prob, st)
        n += 1
    dt = (time.perf_counter() - t0) / n
    return {"value": 1.0 / (dt * w["N"] / n_small), "unit": "evals/s", "cores": 1, "kind": "port",
            "sample": f"{n} uncached evaluations of a {n_small}-object block ({dt * 1e3:.0f} ms each), extrapolated linearly to "
                      f"{w['N']} objects per GPU; the reference itself cannot build or run this shape (SURVEY F6, F12)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name, w = args.config, CONFIGS[args.config]
    world = args.gpus
    n_chains = chains_on_rank(w, world) * world          # the repo arm runs this many chains at N GPUs: like for like
    cfg = config_dict(name, w, world)
    if w["kind"] == "sharded_steps":
        cb = cpu_sharded_steps_baseline(w, w["N"] * world, 12.0)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * w["steps_per_launch"] / cb["value"], "higher_is_better": True, "scaling": w["scaling"],
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return
    if w["kind"] == "sharded_eval":
        cb = cpu_eval_baseline(w, 10.0)
        line = {"impl": "reference", "metric": "uncached_likelihood_evals_per_sec", "value": cb["value"] / world, "unit": "evals/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / (cb["value"] / world),
                "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg, "cpu_baseline": cb,
                "e2e": {"value": cb["value"] / world, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return
    if reference_available():
        rc = ReferenceChains(w, n_chains)
        try:
            probe = rc.run(3)
            per_step = max(probe / 3, 1e-4)
            # one bench step = a bounded sample: about 2 s of MCMC steps on every chain
            n = int(max(3, min(w.get("swap_interval", w.get("steps_per_launch", 200)), 2.0 / per_step)))
            for _ in range(args.warmup):
                rc.run(n)
            t = [rc.run(n) for _ in range(args.steps)]
            total = float(sum(t))
            value = rc.n_chains * n * args.steps / total
            cores, kind = rc.cores, "reference"
            sample = (f"{n} MCMC steps per bench step on each of {rc.n_chains} chains (the repo arm runs the same number of chains at "
                      f"{world} GPU(s)); the reference's own MCMCChain.run (sbayes v1.1, unmodified), {rc.cores} processes, CHECK_CACHING off")
            ms = total / args.steps * 1e3
        finally:
            rc.close()
    else:
        value, n_tot, cores = port_steps_per_sec(w, n_chains, 12.0)
        kind, ms = "port", 12e3
        sample = f"oracle port (numpy), {cores} chains x 12 s (the reference copy oracle/_ref is absent)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
            "dtype": "f64 (f32 at the reference's rounding points)", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ==========================================================================================
# our arm
# ==========================================================================================
def ncu_traffic(kernel: str, tag: str):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t[f"{kernel}|{tag}"]["bytes"]
    except Exception:
        return None


def roofline(kernel, algo_bytes, seconds, peak, traffic, note=None):
    ach = algo_bytes / seconds / 1e9
    out = {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic}
    if traffic:
        out["dram_frac"] = traffic / seconds / 1e9 / peak        # measured DRAM bytes (ncu) over the same time
    if note:
        out["note"] = note
    return out


def full_eval_line(ctx, prob, n_chains, label, tag, flush, peak, reps=7):
    """Stand-alone uncached evaluation (statistics pass + terms) of every chain of `ctx`:
    device times from CUDA events around the kernels (inside the C ABI call), L2 flushed before
    every timed call.  Algorithmic bytes of the statistics pass: 2 B per observation (state
    byte + source byte), SURVEY.md section 8(d) FULL_EVAL."""
    import torch
    out = {"workload": label, "chains": n_chains}
    for mode in ("hist", "mma"):
        try:
            ctx.eval_set_mode(mode)
            ctx.eval_batch()
            ts = []
            for _ in range(reps):
                flush.fill_(1)
                torch.cuda.synchronize()
                ts.append(ctx.eval_batch(timing=True)[1])
            st, tm = float(np.median([t[0] for t in ts])), float(np.median([t[1] for t in ts]))
            algo = 2.0 * prob.N * prob.F * n_chains
            info = ctx.eval_info()          # which statistics pass the library planned for this shape
            kern = {"hist": "sb_eval_hist_kernel", "fine": "sb_eval_fine_kernel", "mma": "sb_counts_mma_kernel"}[info["pass"]]
            out[mode] = {"stats_us": st * 1e3, "terms_us": tm * 1e3, "evals_per_sec": n_chains / ((st + tm) * 1e-3),
                         "pass": info, "roofline": roofline(kern, algo, st * 1e-3, peak, ncu_traffic(kern, tag))}
        except Exception as e:       # the tensor-core pass refuses problems with > 256 group columns
            out[mode] = {"unavailable": str(e)}
    ctx.eval_set_mode("hist")
    try:      # observation likelihood (LikelihoodLogger's product): 2 B read + 4 B written per observation
        ctx.obs_loglik(False, False)
        ts = []
        for _ in range(reps):
            flush.fill_(1)
            torch.cuda.synchronize()
            ts.append(ctx.obs_loglik(False, False)[3])
        t = float(np.median(ts))
        out["obs_lik"] = {"us": t * 1e3, "observations_per_sec": prob.N * prob.F * n_chains / (t * 1e-3),
                          "roofline": roofline("sb_obs_lik_kernel", 6.0 * prob.N * prob.F * n_chains, t * 1e-3, peak,
                                               ncu_traffic("sb_obs_lik_kernel", tag))}
    except Exception as e:
        out["obs_lik"] = {"unavailable": str(e)}
    try:      # full-sweep source Gibbs (GibbsSampleSource over ALL objects, operators.py:496-552): 3 B per observation
        if hasattr(ctx, "gibbs_sweep"):
            ctx.gibbs_sweep(seed=1)
            ts = []
            for i in range(reps):
                flush.fill_(1)
                torch.cuda.synchronize()
                ts.append(ctx.gibbs_sweep(seed=2 + i))
            t = float(np.median(ts))
            out["gibbs_sweep"] = {"us": t * 1e3, "observations_per_sec": prob.N * prob.F * n_chains / (t * 1e-3),
                                  "roofline": roofline("sb_gibbs_sweep_kernel", 3.0 * prob.N * prob.F * n_chains, t * 1e-3, peak,
                                                       ncu_traffic("sb_gibbs_sweep_kernel", tag))}
    except Exception as e:
        out["gibbs_sweep"] = {"unavailable": str(e)}
    try:      # effect draws at log time (conditional_effect_sample of the parameter logger) for every chain
        ctx.effect_sample(step=1)
        ts = [ctx.effect_sample(step=2 + i, timing=True)[1] for i in range(3)]
        out["effect_sample"] = {"us": float(np.median(ts)) * 1e3, "draws": int(prob.G_tot * prob.F * n_chains)}
    except Exception as e:
        out["effect_sample"] = {"unavailable": str(e)}
    return out


def stepping_line(name, w, local, flush, peak, n_steps=200, chains=None):
    """MCMC steps/s of a configuration on ONE GPU (used for the extra c4 line of the default run)."""
    import torch
    from sbayes_b200.backend import DeviceContext
    from sbayes_b200.packed import ChainConfig
    prob, _ = make_problem(w)
    nch = chains or w["chains_per_gpu"]
    sts = make_states(prob, nch, 0)
    ctx = DeviceContext(prob, nch, device=local)
    for c in range(nch):
        ctx.configure(c, ChainConfig(seed=c))
        ctx.set_state(c, sts[c])
    ctx.run(20)
    st2 = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    ms = []
    for _ in range(3):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.fill_(1)
        torch.cuda.synchronize()
        ev0.record(st2)
        ctx.run_async(n_steps)
        ev1.record(st2)
        ctx.sync()
        ms.append(ev0.elapsed_time(ev1))
    t = float(np.median(ms)) * 1e-3
    line = full_eval_line(ctx, prob, nch, workload_text(name, w, 1), name, flush, peak)
    line["steps"] = {"mcmc_steps_per_sec": nch * n_steps / t, "us_per_step_and_chain": t * 1e6 / n_steps, "team_ctas": ctx.info()["team"],
                     "grid": ctx.info()["grid"], "smem": ctx.info()["smem"],
                     "roofline": roofline("sb_run_kernel", float(step_bytes(prob)) * nch * n_steps, t, peak, ncu_traffic("sb_run_kernel", name))}
    ctx.close()
    return line


def run_ours(args):
    import torch
    import torch.distributed as dist
    from sbayes_b200.packed import ChainConfig

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # (NCCL prints its version banner to stdout otherwise: the line must be the JSON)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    name, w = args.config, CONFIGS[args.config]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def gather(vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world == 1:
            return [t.tolist()]
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [o.tolist() for o in out]

    if w["kind"] == "sharded_eval":
        line = run_sharded_eval(args, name, w, rank, world, local, flush, peak, barrier, allmax, peaks)
    elif w["kind"] == "sharded_steps":
        line = run_sharded_steps(args, name, w, rank, world, local, flush, peak, barrier, allmax, gather, peaks)
    else:
        line = run_stepping(args, name, w, rank, world, local, flush, peak, barrier, allmax, gather, peaks)
    barrier()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_stepping(args, name, w, rank, world, local, flush, peak, barrier, allmax, gather, peaks):
    import torch
    from sbayes_b200.mc3 import DeviceMC3
    from sbayes_b200.packed import ChainConfig
    n_local = chains_on_rank(w, world)
    prob, _ = make_problem(w)
    is_mc3 = w["kind"] == "mc3"
    SI = w["swap_interval"] if is_mc3 else w["steps_per_launch"]
    # replicates are an MC3 ladder with temperature_diff 0 that never swaps: same driver object
    mc3 = DeviceMC3(prob, n_local, ChainConfig(seed=77), dT=w.get("dT", 0.0), device=local, team=args.team,
                    swap_seed=5, rank=rank, world=world)
    states = make_states(prob, n_local, 0)
    states = [states[(mc3.global_chain(l)) % len(states)] for l in range(n_local)]
    mc3.set_states(states)
    info = mc3.ctx.info()
    stream = torch.cuda.ExternalStream(mc3.ctx.stream(), device=torch.device("cuda", local))

    def one_round():
        mc3.ctx.run_async(SI)
        mc3.ctx.sync()
        if is_mc3:
            mc3.swap()

    for _ in range(args.warmup):
        one_round()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident timing: K bench steps, CUDA events on the launching stream ------
    kern_ms, step_ms = [], []
    launches0 = mc3.ctx.info()["launches"]
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        mc3.ctx.run_async(SI)
        e1.record(stream)
        mc3.ctx.sync()
        if is_mc3:
            mc3.swap()
        e2.record(stream)
        e2.synchronize()
        kern_ms.append(e0.elapsed_time(e1))
        step_ms.append(e0.elapsed_time(e2))
    barrier()
    launches = mc3.ctx.info()["launches"] - launches0
    total_ms, total_kern_ms = allmax([sum(step_ms), sum(kern_ms)])
    per_rank = gather([float(np.mean(kern_ms)), float(np.mean(step_ms)) - float(np.mean(kern_ms))])
    clocks = sampler.stop() if rank == 0 else None
    n_mcmc = SI * n_local * world * args.steps
    value = n_mcmc / (total_ms / 1000.0)

    # ---- end to end through the public API with host buffers -----------------------------
    h2d = sum(s.cluster_of.nbytes + s.W.nbytes + s.src.nbytes for s in states)
    d2h = h2d + n_local * (prob.G_tot * prob.F * prob.S * 4 + 48)
    cur = [mc3.ctx.get_state(c) for c in range(n_local)]
    e2e_steps = max(1, min(args.steps, 5)) if n_local > 16 else args.steps
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mc3.set_states(cur)                       # host -> device: clusters, weights, sources of every chain
        one_round()                               # the MCMC steps (+ swap: reads ll / prior back)
        cur = [mc3.ctx.get_state(c) for c in range(n_local)]   # device -> host: the samples
    barrier()
    e2e_s = allmax([time.perf_counter() - t0])[0]
    e2e_value = SI * n_local * world * e2e_steps / e2e_s

    line = None
    if rank == 0:
        kern_s = total_kern_ms / 1000.0 / args.steps
        kr = [p[0] for p in per_rank]
        cfg = dict(config_dict(name, w, world), team_ctas=info["team"], grid=info["grid"], smem=info["smem"])
        note = ("persistent step kernel; algorithmic bytes = N*F per MCMC step (SURVEY 8d). " +
                ("At 1000 x 100 the state is L1/L2 resident and the kernel is latency bound (team barriers), so the HBM fraction "
                 "is small by construction." if name in ("c2", "c3") else
                 "One CTA per chain; the pass tables are built tile by tile in shared memory (sb_core.cuh: tiled_pass)."))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "likelihood_evals_per_sec": value,     # one candidate posterior evaluation per MCMC step
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps},
            "gpu_launches": int(launches), "clocks": clocks,
            "per_rank": {"kern_ms": {"min": min(kr), "median": float(np.median(kr)), "max": max(kr)},
                         "swap_path_ms": {"min": min(p[1] for p in per_rank), "max": max(p[1] for p in per_rank)},
                         "note": "mean per bench step of every rank: step kernel (CUDA events) and everything after it (sync, swap scalars, all_gather)"},
            "roofline": dict(roofline("sb_run_kernel", float(step_bytes(prob)) * SI * n_local, kern_s, peak, ncu_traffic("sb_run_kernel", name), note),
                             peak_source="MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"),
        }
        # ---- likelihood evals/s: stand-alone uncached evaluation (rank 0 reports its own GPU) ----
        if not args.no_full_eval:
            fe = [full_eval_line(mc3.ctx, prob, n_local, workload_text(name, w, world), name, flush, peak)]
            line["full_uncached_evals_per_sec"] = fe[0]["hist"]["evals_per_sec"] * world
            if name == "c2" and world == 1:
                # the HBM-meaningful shapes next to the (latency-bound) headline: config 4 at its full single-GPU
                # size with its own steps/s, and the config-5 object shard
                fe.append(stepping_line("c4", CONFIGS["c4"], local, flush, peak))
                from sbayes_b200.backend import DeviceContext
                w5 = CONFIGS["c5e"]
                p5, _ = make_problem(w5)
                c5 = DeviceContext(p5, 1, device=local)
                c5.configure(0, ChainConfig(seed=0))
                c5.set_state(0, make_states(p5, 1, 0)[0])
                fe.append(full_eval_line(c5, p5, 1, workload_text("c5e", w5, 8), "c5", flush, peak))
                c5.close()
            line["full_eval"] = fe
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline(name, w, n_local * world)
    mc3.close()
    return line


def assemble_sharded_problem(w, rank, world, local):
    """every rank generates its block of objects; the packed observations and group ids are all-gathered so that
    every rank holds the whole matrix (replicated read-only: 200 MB at 200 000 x 1 000)"""
    import torch
    import torch.distributed as dist
    from dataclasses import replace
    prob, _ = make_problem(w, rank)
    if world == 1:
        return prob
    dev = torch.device("cuda", local)

    def gather(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return np.concatenate([o.cpu().numpy() for o in out], axis=0)
    x = gather(prob.x)
    gids = [gather(g) for g in prob.conf_gid]
    n_tot = x.shape[0]
    return replace(prob, x=x, conf_gid=gids, max_size=max(prob.min_size + 1, n_tot // prob.K),
                   adj_indptr=np.zeros(n_tot + 1, np.int32), adj_indices=np.zeros(0, np.int32))


def cpu_sharded_steps_baseline(w, n_total, budget_s=15.0):
    """c5 stepping: the reference can neither build this shape (O(N^2) distance matrix) nor run its
    cluster operators at F = 1000 (underflow, F6); the timed CPU path is the oracle port on a 1,000-object block of
    the same matrix, scaled linearly in the number of objects (flagged as extrapolated)."""
    from oracle import chain as CH
    from oracle.draws import PhiloxDraws
    from sbayes_b200 import synth
    from sbayes_b200.packed import ChainConfig
    n_small = 1000
    prob, _ = synth.make_problem(N=n_small, F=w["F"], S=w["S"], K=w["K"], n_fam=w["n_fam"], seed=w["seed"], with_adjacency=False)
    oc = CH.OracleChain(prob, ChainConfig(seed=1), synth.initial_state(prob, seed=1000), draws=PhiloxDraws(1))
    oc.step()
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < budget_s:
        oc.step(); n += 1
    dt = (time.perf_counter() - t0) / n
    return {"value": 1.0 / (dt * n_total / n_small), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n} MCMC steps of a {n_small}-object block ({dt * 1e3:.0f} ms per step), extrapolated linearly to {n_total} objects; "
                      f"the reference itself cannot build or run this shape (SURVEY F6, F12)"}


def run_sharded_steps(args, name, w, rank, world, local, flush, peak, barrier, allmax, gather, peaks):
    import torch
    from sbayes_b200 import synth
    from sbayes_b200.packed import ChainConfig
    from sbayes_b200.shard import ShardedChains
    prob = assemble_sharded_problem(w, rank, world, local)
    SPL = w["steps_per_launch"]
    sc = ShardedChains(prob, 1, rank=rank, world=world, device=local, team=args.team)
    sc.configure(0, ChainConfig(seed=77))
    st = synth.initial_state(prob, seed=1000) if rank == 0 else None
    sc.set_state(0, st)
    info = sc.ctx.info()
    stream = torch.cuda.ExternalStream(sc.ctx.stream(), device=torch.device("cuda", local))
    for _ in range(args.warmup):
        sc.run(SPL)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = sc.ctx.info()["launches"]
    kern_ms = []
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sc.run_async(SPL)
        e1.record(stream)
        sc.sync()
        kern_ms.append(e0.elapsed_time(e1))
    barrier()
    launches = sc.ctx.info()["launches"] - launches0
    total_ms = allmax([sum(kern_ms)])[0]
    per_rank = gather([float(np.mean(kern_ms))])
    clocks = sampler.stop() if rank == 0 else None
    value = SPL * args.steps / (total_ms / 1e3)
    # end to end: rank 0 uploads the whole state from host memory, all ranks step, rank 0 reads the sample back
    e2e_steps = max(1, min(args.steps, 3))
    cur = sc.get_state(0)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sc.set_state(0, cur)
        sc.run(SPL)
        cur = sc.get_state(0)
    barrier()
    e2e_s = allmax([time.perf_counter() - t0])[0]
    line = None
    if rank == 0:
        kr = [p[0] for p in per_rank]
        h2d = st.cluster_of.nbytes + st.W.nbytes + st.src.nbytes
        d2h = h2d + prob.G_tot * prob.F * prob.S * 4 + 48
        kern_s = total_ms / 1e3 / args.steps
        post = sc.posterior(0)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": dict(config_dict(name, w, world), team_ctas_per_gpu=info["team"], grid=info["grid"], smem=info["smem"],
                                                    objects=int(prob.N)),
                "likelihood_evals_per_sec": value,
                "e2e": {"value": SPL * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps},
                "gpu_launches": int(launches), "clocks": clocks,
                "per_rank": {"kern_ms": {"min": min(kr), "median": float(np.median(kr)), "max": max(kr)}},
                "roofline": dict(roofline("sb_run_kernel", float(step_bytes(prob)) * SPL, kern_s, peak * world, None,
                                          "one chain over all GPUs: algorithmic bytes = N_total*F per MCMC step, peak = GPUs x measured HBM peak"),
                                 peak_source="MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"),
                "posterior": {"ll": post["ll"], "prior": post["prior"]}}
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_sharded_steps_baseline(w, prob.N)
    sc.close()
    return line


def run_sharded_eval(args, name, w, rank, world, local, flush, peak, barrier, allmax, peaks):
    import torch
    from sbayes_b200.packed import ChainConfig
    from sbayes_b200.shard import ShardedEvaluator
    prob, _ = make_problem(w, rank)
    st = make_states(prob, 1, rank)[0]
    ev = ShardedEvaluator.from_shard(prob, w["N"] * world, 1, rank, world, device=local)
    ev.configure(0, ChainConfig(seed=0))
    ev.set_local_state(0, st)
    for _ in range(args.warmup):
        ev.evaluate()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ev.ctx.info()["launches"]
    ts, stats_ms = [], []
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()      # (the evaluation spans the ctx stream and the collective's stream: host clock between two device syncs)
        ev.evaluate()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    barrier()
    launches = ev.ctx.info()["launches"] - launches0
    for _ in range(5):
        flush.fill_(1)
        torch.cuda.synchronize()
        stats_ms.append(ev.ctx.eval_stats())
    total_ms = allmax([sum(ts)])[0]
    clocks = sampler.stop() if rank == 0 else None
    # end to end: the state block of the rank comes from host memory every time
    h2d = st.cluster_of.nbytes + st.W.nbytes + st.src.nbytes
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ev.set_local_state(0, st)
        res = ev.evaluate()
    barrier()
    e2e_s = allmax([time.perf_counter() - t0])[0]
    line = None
    if rank == 0:
        sm = float(np.median(stats_ms)) * 1e-3
        line = {"metric": "uncached_likelihood_evals_per_sec", "value": args.steps / (total_ms / 1e3), "unit": "evals/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": w["scaling"], "vs_baseline": None, "dtype": "i32 counts / f64 terms", "data": "synthetic",
                "config": config_dict(name, w, world),
                "e2e": {"value": args.steps / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 48},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": dict(roofline("sb_eval_hist_kernel", 2.0 * prob.N * prob.F, sm, peak, ncu_traffic("sb_eval_hist_kernel", "c5")),
                                 peak_source="MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"),
                "posterior": {"ll": res[0]["ll"], "prior": res[0]["prior"]},
                "note": "object-sharded MCMC stepping of this configuration is not built; a config-5 chain steps on one GPU (148-CTA team)"}
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline(name, w, 1)
    ev.close()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--team", type=int, default=0, help="CTAs per chain (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-full-eval", action="store_true", help="skip the stand-alone evaluation measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
