#!/usr/bin/env python
"""bench.py -- MCMC steps/s of the sBayes inner loop on B200 (BASELINE.json metric).

Workload (configs[1] of BASELINE.json): synthetic 1,000 objects x 100 features x 4 states,
5 clusters + family (8 groups) + universal, MC3 with 4 tempered chains per GPU
(temperature_diff 0.05, swap_interval 1000), default operator mix 70/10/20.  One bench
"step" is one MC3 round: `swap_interval` MCMC steps on every chain of the rank plus the
swap.  With N GPUs the temperature ladder has 4N chains, 4 per GPU, and the swap scalars are
all-gathered over NCCL (weak scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
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

WORKLOAD = dict(N=1000, F=100, S=4, K=5, n_fam=8, seed=2, chains_per_gpu=4, dT=0.05, swap_interval=1000)
METRIC = "mcmc_steps_per_sec"
UNIT = "steps/s"


def make_workload(rank: int):
    from sbayes_b200 import synth
    w = WORKLOAD
    prob, _ = synth.make_problem(N=w["N"], F=w["F"], S=w["S"], K=w["K"], n_fam=w["n_fam"], seed=w["seed"])
    states = [synth.initial_state(prob, seed=1000 + rank * w["chains_per_gpu"] + c) for c in range(w["chains_per_gpu"])]
    return prob, states


def step_bytes(prob) -> int:
    """Algorithmic HBM bytes of one MCMC step, SURVEY.md section 8(d): STEP ~ N*F (one pass
    over the packed observations of the candidate objects; tables and weights are O(G*F*S))."""
    return prob.N * prob.F


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


def cpu_port_steps_per_sec(prob, states, budget_s: float, n_procs: int):
    """The oracle (numpy restatement of the reference algorithm) timed on host cores: the same
    workload, `n_procs` chains in `n_procs` processes (the reference runs one process per MC3
    chain, sbayes/mcmc_setup.py:271-299), each for about `budget_s` seconds."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    q = ctx.Queue()

    def work(c):
        from oracle import chain as CH
        from oracle.draws import PhiloxDraws
        from sbayes_b200.packed import ChainConfig
        cfg = ChainConfig(temperature=1.0 + WORKLOAD["dT"] * c, seed=c)
        oc = CH.OracleChain(prob, cfg, states[c % len(states)].copy(), draws=PhiloxDraws(c))
        for _ in range(3):
            oc.step()
        t0, n = time.perf_counter(), 0
        while time.perf_counter() - t0 < budget_s:
            oc.step(); n += 1
        q.put((n, time.perf_counter() - t0))

    procs = [ctx.Process(target=work, args=(c,)) for c in range(n_procs)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    return sum(n / t for n, t in res), sum(n for n, _ in res)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is
    pure Python and does not exist on the GPU box, so the timed code is the oracle port of
    its algorithm (oracle/, validated step by step against the live reference), one process
    per chain like the reference's MC3."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prob, states = make_workload(0)
    cores = min(WORKLOAD["chains_per_gpu"], os.cpu_count() or 1)
    budget = 12.0
    vals = []
    for _ in range(max(1, min(args.steps, 3))):
        v, n = cpu_port_steps_per_sec(prob, states, budget, cores)
        vals.append(v)
    value = float(np.mean(vals))
    sample = (f"{cores} chains x ~{budget:.0f} s of MCMC steps of the same workload per bench step, oracle port "
              f"(numpy); live reference measured in the build container: ~31-35 steps/s per chain/core (BASELINE.md)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * WORKLOAD["swap_interval"] * cores / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def config_dict(n_gpus: int) -> dict:
    w = WORKLOAD
    return {"workload": f"sbayes.simulation synthetic {w['N']}x{w['F']}x{w['S']}, K={w['K']}, family({w['n_fam']})+universal, "
                        f"MC3 {w['chains_per_gpu']} chains/GPU, swap_interval {w['swap_interval']}, operators 70/10/20",
            "mcmc_steps_per_bench_step": w["swap_interval"] * w["chains_per_gpu"] * n_gpus,
            "chains": w["chains_per_gpu"] * n_gpus, "parallelism": f"chains x{n_gpus} (temperature swap over NCCL all_gather)",
            "l2": "flushed between timed bench steps (256 MiB write); state is L2-resident inside a step by design"}


FULL_EVAL_SHAPES = [
    # (label, problem, chains per GPU): BASELINE configs 4 and 5 as one GPU sees them at 8 GPUs
    ("c4 replicates 5000x200x6, K=5, 128 chains/GPU", dict(N=5000, F=200, S=6, K=5, n_fam=20, seed=4), 128),
    ("c5 object shard 25000x1000x8, K=16 (200000 objects over 8 GPUs)", dict(N=25000, F=1000, S=8, K=16, n_fam=64, seed=5), 1),
]


def ncu_traffic(kernel: str, tag: str):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t[f"{kernel}|{tag}"]["bytes"]
    except Exception:
        return None


def full_eval_line(ctx, prob, n_chains, label, flush, peak, reps=7):
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
            out[mode] = {"stats_us": st * 1e3, "terms_us": tm * 1e3, "evals_per_sec": n_chains / ((st + tm) * 1e-3),
                         "roofline": {"bound": "hbm", "kernel": "sb_eval_hist_kernel" if mode == "hist" else "sb_counts_mma_kernel",
                                      "achieved": algo / (st * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                      "frac": algo / (st * 1e-3) / 1e9 / peak,
                                      "traffic": ncu_traffic("sb_eval_hist_kernel", label[:2]) if mode == "hist" else None}}
        except Exception as e:       # the tensor-core pass refuses problems with > 256 group columns
            out[mode] = {"unavailable": str(e)}
    ctx.eval_set_mode("hist")
    # observation likelihood (LikelihoodLogger's product): 2 B read + 4 B written per observation
    try:
        ctx.obs_loglik(False, False)
        ts = []
        for _ in range(reps):
            flush.fill_(1)
            torch.cuda.synchronize()
            ts.append(ctx.obs_loglik(False, False)[3])
        t = float(np.median(ts))
        algo = 6.0 * prob.N * prob.F * n_chains
        out["obs_lik"] = {"us": t * 1e3, "observations_per_sec": prob.N * prob.F * n_chains / (t * 1e-3),
                          "roofline": {"bound": "hbm", "kernel": "sb_obs_lik_kernel", "achieved": algo / (t * 1e-3) / 1e9, "peak": peak,
                                       "unit": "GB/s", "frac": algo / (t * 1e-3) / 1e9 / peak, "traffic": None}}
    except Exception as e:
        out["obs_lik"] = {"unavailable": str(e)}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from sbayes_b200.mc3 import DeviceMC3
    from sbayes_b200.packed import ChainConfig

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = WORKLOAD
    prob, states = make_workload(rank)
    mc3 = DeviceMC3(prob, w["chains_per_gpu"], ChainConfig(seed=77), dT=w["dT"], device=local, team=args.team,
                    swap_seed=5, rank=rank, world=world)
    mc3.set_states(states)
    info = mc3.ctx.info()
    stream = torch.cuda.ExternalStream(mc3.ctx.stream(), device=torch.device("cuda", local))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    SI = w["swap_interval"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        mc3.round_trip(SI)
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
        mc3.swap()
        e2.record(stream)
        e2.synchronize()
        kern_ms.append(e0.elapsed_time(e1))
        step_ms.append(e0.elapsed_time(e2))
    barrier()
    launches = mc3.ctx.info()["launches"] - launches0
    tot = torch.tensor([sum(step_ms), sum(kern_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    total_ms, total_kern_ms = tot.tolist()
    clocks = sampler.stop() if rank == 0 else None
    n_mcmc = SI * w["chains_per_gpu"] * world * args.steps
    value = n_mcmc / (total_ms / 1000.0)

    # ---- end to end through the public API with host buffers -----------------------------
    h2d = sum(s.cluster_of.nbytes + s.W.nbytes + s.src.nbytes for s in states)
    d2h = h2d + states[0].src.shape[0] * 0 + w["chains_per_gpu"] * (prob.G_tot * prob.F * prob.S * 4 + 48)
    cur = [mc3.ctx.get_state(c) for c in range(w["chains_per_gpu"])]
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        mc3.set_states(cur)                       # host -> device: clusters, weights, sources of every chain
        mc3.round_trip(SI)                        # swap_interval steps + swap (reads ll / prior back)
        cur = [mc3.ctx.get_state(c) for c in range(w["chains_per_gpu"])]   # device -> host: the samples
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = n_mcmc / float(e2e_s.item())

    from json import load
    peaks = {}
    try:
        peaks = load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # ---- likelihood evals/s: stand-alone uncached evaluation (rank 0 reports its own GPU) ----
    full_eval = []
    if rank == 0 and not args.no_full_eval:
        from sbayes_b200 import synth
        from sbayes_b200.backend import DeviceContext
        full_eval.append(full_eval_line(mc3.ctx, prob, w["chains_per_gpu"], config_dict(world)["workload"], flush, peak))
        for label, shape, nch in FULL_EVAL_SHAPES:
            p2, _ = synth.make_problem(**shape)
            sts = [synth.initial_state(p2, seed=100 + c) for c in range(min(nch, 2))]
            c2 = DeviceContext(p2, nch, device=local)
            for c in range(nch):
                c2.configure(c, ChainConfig(seed=c))
                c2.set_state(c, sts[c % len(sts)])
            line2 = full_eval_line(c2, p2, nch, label, flush, peak)
            if nch > 1:
                # MCMC steps/s of the replicate-chain configuration itself (one CTA per chain)
                n_st = 200
                c2.run(20)
                st2 = torch.cuda.ExternalStream(c2.stream(), device=torch.device("cuda", local))
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                flush.fill_(1)
                torch.cuda.synchronize()
                ev0.record(st2)
                c2.run_async(n_st)
                ev1.record(st2)
                c2.sync()
                ms = ev0.elapsed_time(ev1)
                algo = float(p2.N) * p2.F * nch * n_st
                line2["steps"] = {"mcmc_steps_per_sec": nch * n_st / (ms * 1e-3), "us_per_step_and_chain": ms * 1e3 / n_st, "team_ctas": c2.info()["team"],
                                  "roofline": {"bound": "hbm", "kernel": "sb_run_kernel", "achieved": algo / (ms * 1e-3) / 1e9, "peak": peak,
                                               "unit": "GB/s", "frac": algo / (ms * 1e-3) / 1e9 / peak, "traffic": None}}
            full_eval.append(line2)
            c2.close()
    barrier()

    if rank == 0:
        per_launch_bytes = step_bytes(prob) * SI * w["chains_per_gpu"]
        kern_s = total_kern_ms / 1000.0 / args.steps
        achieved = per_launch_bytes / kern_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": dict(config_dict(world), team_ctas=info["team"], grid=info["grid"]),
            "likelihood_evals_per_sec": value,     # one candidate posterior evaluation per MCMC step
            "full_uncached_evals_per_sec": (full_eval[0]["hist"]["evals_per_sec"] * world) if full_eval else None,
            "full_eval": full_eval,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "sb_run_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic("sb_run_kernel", "c2"),
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "note": "persistent step kernel at 1000x100 is latency bound (state is L2/L1 resident); "
                                 "algorithmic bytes = N*F per MCMC step (SURVEY 8d)"},
        }
        if not args.no_cpu:
            cores = min(w["chains_per_gpu"], os.cpu_count() or 1)
            v, n = cpu_port_steps_per_sec(prob, states, 15.0, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n} MCMC steps of the same workload: {cores} chains x 15 s, oracle port (numpy), "
                                              f"one process per chain"}
        print(json.dumps(line))
    mc3.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
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
