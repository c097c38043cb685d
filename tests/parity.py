"""Shared parity checks: a DeviceContext (CUDA library on the GPU box, fibre-emulator build in
the GPU-less container) against the golden fixtures recorded from the live reference and
against the fp64 oracle on identical Philox streams."""
import os

import numpy as np

from oracle import fixtures as FX, chain as CH, model as M
from oracle.draws import PhiloxDraws
from sbayes_b200.backend import DeviceContext
from sbayes_b200.packed import ChainConfig

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# vs the live reference (float32 at its rounding points)
TOL_REF_LOGQ, TOL_REF_LL, TOL_REF_PRIOR = 2e-6, 5e-7, 5e-7
# vs the fp64 oracle: BASELINE.json north_star asks for 1e-10 relative on fp64 log-likelihoods
TOL_ORACLE = 1e-10


def rel(a, b):
    if not (np.isfinite(a) and np.isfinite(b)):
        return 0.0 if (a == b or (np.isnan(a) and np.isnan(b))) else np.inf
    return abs(a - b) / max(1.0, abs(a))


def replay_golden(case: str, lib=None, team: int = 1, n_steps=None):
    """The device path replays the reference's recorded draws and must reproduce the
    reference's decisions, proposal probabilities, posterior values and final state."""
    fx = FX.load_replay(os.path.join(GOLDEN, case + ".npz"))
    prob, cfg, outs, raw = fx["prob"], fx["cfg"], fx["outs"], fx["raw"]
    n = len(outs["op"]) if n_steps is None else n_steps
    ctx = DeviceContext(prob, 1, team=team, _lib=lib)
    try:
        ctx.configure(0, cfg)
        ctx.set_state(0, fx["st0"])
        ev = ctx.eval(0)
        st = ctx.get_state(0)
        assert np.array_equal(st.counts, fx["st0"].counts), "bulk count rebuild differs from the reference's counts"
        assert rel(float(raw["ref_ll0"]), ev["ll"]) < TOL_REF_LL
        assert rel(float(raw["ref_prior0"]), ev["prior"]) < TOL_REF_PRIOR
        # ... and against the fp64 oracle on the same state
        assert rel(M.log_likelihood(prob, fx["st0"]), ev["ll"]) < TOL_ORACLE
        assert rel(M.log_prior(prob, fx["st0"]), ev["prior"]) < TOL_ORACLE
        tape = dict(vals=raw["tape_vals"], off=raw["tape_off"], len=raw["tape_len"], wide_rows=raw["wide_rows"])
        rec = ctx.run(n, tape=tape, records=True)[0]
        for i in range(n):
            r = rec[i]
            assert r["op"] == int(outs["op"][i]), f"step {i}: operator"
            assert bool(r["accept"]) == bool(outs["accept"][i]), f"step {i}: accept decision"
            assert rel(outs["log_q"][i], r["log_q"]) < TOL_REF_LOGQ, f"step {i}: log_q"
            assert rel(outs["log_q_back"][i], r["log_q_back"]) < TOL_REF_LOGQ, f"step {i}: log_q_back"
            assert rel(outs["ll_new"][i], r["ll_new"]) < TOL_REF_LL, f"step {i}: ll"
            assert rel(outs["prior_new"][i], r["prior_new"]) < TOL_REF_PRIOR, f"step {i}: prior"
        if n == len(outs["op"]):
            fin, ref = ctx.get_state(0), fx["st1"]
            assert np.array_equal(fin.cluster_of, ref.cluster_of)
            assert np.array_equal(fin.src, ref.src)
            assert np.array_equal(fin.counts, ref.counts)          # incremental counts, bit-exact
            assert np.array_equal(fin.W, ref.W)
            # incremental == bulk rebuild (the reference's verify_counts, operators.py:2018-2025)
            assert np.array_equal(M.rebuild_counts(prob, fin), fin.counts)
            post = ctx.posterior(0)
            ev2 = ctx.eval(0)
            assert rel(ev2["ll"], post["ll"]) < 1e-12 and rel(ev2["prior"], post["prior"]) < 1e-12
    finally:
        ctx.close()
    return n


def oracle_trajectory(args):
    """One oracle chain on its Philox stream: per-step records and the final state.  A plain
    function of picklable arguments, so long runs can be spread over host processes."""
    prob, cfg, st0, n_steps = args
    s = st0.copy()
    s.counts = None
    oc = CH.OracleChain(prob, cfg, s, draws=PhiloxDraws(cfg.seed))
    recs = []
    for _ in range(n_steps):
        o = oc.step()
        recs.append((o.op, bool(o.accept), o.log_q, o.log_q_back, o.ll_new, o.prior_new, o.mh_ratio, o.log_u))
    return dict(recs=recs, state=oc.state, ll=oc.ll, prior=oc.prior, accepts=oc.accepts, rejects=oc.rejects)


def oracle_trajectories(jobs, procs: int = 1):
    """`jobs` = [(prob, cfg, st0, n_steps)]; with procs > 1 the chains run in forked workers
    (numpy only: the children never touch CUDA)."""
    if procs <= 1 or len(jobs) <= 1:
        return [oracle_trajectory(j) for j in jobs]
    import multiprocessing as mp
    with mp.get_context("fork").Pool(min(procs, len(jobs))) as pool:
        return pool.map(oracle_trajectory, jobs, chunksize=1)


def compare_with_trajectories(ctx, rec, trajs, tol=TOL_ORACLE):
    """device records / final states of every chain of `ctx` against oracle trajectories"""
    near_ties = 0
    for c, tr in enumerate(trajs):
        for i, (op, acc, lq, lqb, ll_new, pr_new, ratio, log_u) in enumerate(tr["recs"]):
            r = rec[c, i]
            assert r["op"] == op, f"chain {c} step {i}: operator {r['op']} vs {op}"
            if np.isfinite(ratio) and abs(log_u - ratio) < 1e-9:
                near_ties += 1
            assert bool(r["accept"]) == acc, f"chain {c} step {i} op {op}: accept"
            assert rel(lq, r["log_q"]) < tol, f"chain {c} step {i} op {op}: log_q {lq} {r['log_q']}"
            assert rel(lqb, r["log_q_back"]) < tol, f"chain {c} step {i} op {op}: log_q_back"
            assert rel(ll_new, r["ll_new"]) < tol, f"chain {c} step {i} op {op}: ll"
            assert rel(pr_new, r["prior_new"]) < tol, f"chain {c} step {i} op {op}: prior"
        fin = ctx.get_state(c)
        assert np.array_equal(fin.cluster_of, tr["state"].cluster_of)
        assert np.array_equal(fin.src, tr["state"].src)
        assert np.array_equal(fin.counts, tr["state"].counts)
        assert np.array_equal(fin.W, tr["state"].W)
        post = ctx.posterior(c)
        assert rel(tr["ll"], post["ll"]) < tol and rel(tr["prior"], post["prior"]) < tol
        a, rj = ctx.op_stats(c)
        assert np.array_equal(a, tr["accepts"]) and np.array_equal(rj, tr["rejects"])
    return near_ties


def philox_lockstep(prob, st0, cfgs, n_steps, lib=None, team: int = 1, tol=TOL_ORACLE, procs: int = 1):
    """Device chains and oracle chains on the same Philox streams: identical operator
    choices, accept decisions and states; fp64 values within `tol` relative."""
    ctx = DeviceContext(prob, len(cfgs), team=team, _lib=lib)
    try:
        for c, cfg in enumerate(cfgs):
            ctx.configure(c, cfg)
            ctx.set_state(c, st0[c])
        rec = ctx.run(n_steps, records=True)
        trajs = oracle_trajectories([(prob, cfg, st0[c], n_steps) for c, cfg in enumerate(cfgs)], procs)
        return compare_with_trajectories(ctx, rec, trajs, tol)
    finally:
        ctx.close()
