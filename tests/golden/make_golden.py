#!/usr/bin/env python
"""Generate the golden replay fixtures in this directory from the LIVE reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case the reference `MCMCChain.step` is driven with recording RNG proxies
(oracle/refrecord.py) while the fp64 oracle replays the recorded draws in lockstep.  The
script fails if the two ever disagree on a proposal, an accept decision or the resulting
state, so a committed fixture is (a) output of the reference itself and (b) a trajectory
on which oracle == reference held step by step (tolerances below).
"""
from __future__ import annotations

import os
import sys
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import refharness as H, model as M, problem as P, refrecord as R, chain as CH, fixtures as FX  # noqa: E402
from oracle.draws import TapeDraws, SITE_CLUSTER  # noqa: E402
from oracle.operators import OP_WIDE  # noqa: E402
from sbayes_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TOL_LOGQ = 2e-6      # relative (max(1,|x|)); the reference's proposal probabilities are float32
TOL_LL = 5e-7
TOL_PRIOR = 5e-7


def lockstep(name, data, model, st0_fn, T, Tp, n_steps, seed, opcfg=None):
    prob = H.problem_from_reference(data, model)
    st = st0_fn(prob)
    st.counts = M.rebuild_counts(prob, st)
    sample = H.sample_from_state(prob, model, data, st)
    opcfg = opcfg or dict(clusters=70, weights=10, source=20)
    rec = R.ReferenceRecorder(data, model, opcfg, T=T, Tp=Tp, seed=seed)
    rec.start(sample)
    cfg = P.ChainConfig(temperature=T, prior_temperature=Tp, op_clusters=opcfg["clusters"],
                        op_weights=opcfg["weights"], op_source=opcfg["source"], seed=seed)
    oc = CH.OracleChain(prob, cfg, st.copy(), draws=TapeDraws())
    ref_ll0, ref_prior0 = rec.chain._ll * T, rec.chain._prior * Tp
    ref_parts0 = [float(sample.cache.cluster_size_prior.value), float(np.sum(sample.cache.geo_prior.value)) if prob.geo else 0.0,
                  float(sample.cache.weights_prior.value), float(np.sum(sample.cache.source_prior.value))]

    def rel(a, b):
        if not (np.isfinite(a) and np.isfinite(b)):
            return 0.0 if a == b else np.inf
        return abs(a - b) / max(1.0, abs(a))

    tapes, outs, wide_rows, near_ties = [], [], [], 0
    worst = dict(lq=0.0, lqb=0.0, ll=0.0, pr=0.0)
    for i in range(n_steps):
        tape, out = rec.step()
        n_avail = 0
        if out["op"] == OP_WIDE:
            k = int(tape[SITE_CLUSTER][0])
            n_avail = int(np.count_nonzero((oc.state.cluster_of < 0) | (oc.state.cluster_of == k)))
        r = oc.step(tape)
        ref_state = H.state_from_sample(prob, rec.sample)
        same = (np.array_equal(ref_state.cluster_of, oc.state.cluster_of)
                and np.array_equal(ref_state.src, oc.state.src)
                and np.array_equal(ref_state.counts, oc.state.counts)
                and np.array_equal(ref_state.W, oc.state.W))
        e = dict(lq=rel(out["log_q"], r.log_q), lqb=rel(out["log_q_back"], r.log_q_back),
                 ll=rel(out["ll_new"], r.ll_new), pr=rel(out["prior_new"], r.prior_new))
        for kk in worst:
            worst[kk] = max(worst[kk], e[kk])
        if np.isfinite(r.mh_ratio) and abs(r.log_u - r.mh_ratio) < 1e-3:
            near_ties += 1
        ok = (out["op"] == r.op and out["accept"] == r.accept and same and e["lq"] < TOL_LOGQ
              and e["lqb"] < TOL_LOGQ and e["ll"] < TOL_LL and e["pr"] < TOL_PRIOR)
        if not ok:
            raise SystemExit(f"[{name}] oracle != reference at step {i}: {out} vs {r} same={same} err={e}")
        tapes.append(tape)
        outs.append(out)
        wide_rows.append(n_avail)
    final = H.state_from_sample(prob, rec.sample)
    ops = np.bincount([o["op"] for o in outs], minlength=P.N_OPS)
    print(f"[{name}] {n_steps} steps OK  N={prob.N} F={prob.F} S={prob.S} K={prob.K} C={prob.C}  worst rel err {worst}  "
          f"near-ties {near_ties}  ops {ops.tolist()}  accepts {int(sum(o['accept'] for o in outs))}")
    extra = dict(ref_ll0=np.float64(ref_ll0), ref_prior0=np.float64(ref_prior0), ref_prior_parts0=np.array(ref_parts0),
                 worst_rel_err=np.array([worst["lq"], worst["lqb"], worst["ll"], worst["pr"]]))
    FX.save_replay(os.path.join(HERE, name + ".npz"), prob, cfg, st, final, tapes, outs, wide_rows, extra)


def synth_case(name, N, F, S, K, n_fam, seed, T, Tp, n_steps, geo=None, size_type="uniform_area", min_size=3,
               max_size=40, opcfg=None, per_cluster=6, family_prior="jeffreys", weights_prior=None):
    prob0, truth = synth.make_problem(N=N, F=F, S=S, K=K, n_fam=n_fam, seed=seed, jagged_states=True)
    confs = OrderedDict(universal=P.bool_from_gid(prob0.conf_gid[0], 1))
    ce = dict(universal={"<ALL>": dict(type="uniform")})
    if n_fam:
        confs["family"] = P.bool_from_gid(prob0.conf_gid[1], n_fam)
        ce["family"] = {"<DEFAULT>": dict(type=family_prior) if family_prior != "symmetric_dirichlet"
                        else dict(type="symmetric_dirichlet", prior_concentration=0.7)}
    prior_cfg = dict(objects_per_cluster=dict(type=size_type, min=min_size, max=max_size),
                     geo=geo or dict(type="uniform"), weights=weights_prior or dict(type="uniform"),
                     cluster_effect=dict(type="uniform"), confounding_effects=ce)
    data, model = H.build_reference_model(prob0.x, prob0.applicable, truth["locations"], confs, K, prior_cfg)
    lockstep(name, data, model, lambda p: synth.initial_state(p, seed=seed + 100, objects_per_cluster=per_cluster),
             T, Tp, n_steps, seed, opcfg)


def config_case(name, config_path, K, T, Tp, n_steps, seed, per_cluster):
    r = H.ref()
    cfg = r.config.SBayesConfig.from_config_file(config_path, custom_settings={"model": {"clusters": K}})
    import sbayes.load_data as LD
    objects, features, confounders = LD.read_features_from_csv(
        data_path=cfg.data.features, feature_states_path=cfg.data.feature_states,
        confounder_names=cfg.model.confounders)
    data = LD.Data(objects=objects, features=features, confounders=confounders, projection=None,
                   geo_costs=cfg.model.prior.geo.costs)
    data.network.lat_lon = data.network.locations
    model = r.model.Model(data, cfg.model)
    opcfg = dict(clusters=cfg.mcmc.operators.clusters, weights=cfg.mcmc.operators.weights, source=cfg.mcmc.operators.source)
    lockstep(name, data, model, lambda p: synth.initial_state(p, seed=seed + 100, objects_per_cluster=per_cluster),
             T, Tp, n_steps, seed, opcfg)


if __name__ == "__main__":
    only = sys.argv[1:]

    def want(n):
        return not only or n in only

    # c1: the reference's own test fixture (test/test_files), K=2 override, cost-based geo prior
    if want("c1_testfiles"):
        config_case("c1_testfiles", os.path.join(H.REFERENCE_ROOT, "test/test_files/config.yaml"), K=2, T=1.0, Tp=1.0,
                    n_steps=400, seed=11, per_cluster=1)
    # c3: the south_america experiment (100 x 36, Dirichlet priors from JSON, families + isolates)
    if want("c3_south_america"):
        config_case("c3_south_america", os.path.join(H.REFERENCE_ROOT, "experiments/south_america/config.yaml"), K=3,
                    T=1.0, Tp=1.0, n_steps=300, seed=12, per_cluster=10)
    if want("c3_south_america_hot"):
        config_case("c3_south_america_hot", os.path.join(H.REFERENCE_ROOT, "experiments/south_america/config.yaml"), K=3,
                    T=1.14, Tp=1.07, n_steps=200, seed=13, per_cluster=10)
    if want("synth_small"):
        synth_case("synth_small", 60, 10, 4, 3, 4, seed=1, T=1.0, Tp=1.0, n_steps=400, size_type="uniform_size")
    if want("synth_tempered"):
        synth_case("synth_tempered", 60, 10, 4, 3, 4, seed=2, T=1.3, Tp=1.2, n_steps=400, size_type="uniform_size",
                   weights_prior=dict(type="symmetric_dirichlet", prior_concentration=2.0))
    if want("synth_geo"):
        synth_case("synth_geo", 40, 8, 3, 2, 3, seed=3, T=1.0, Tp=1.0, n_steps=300,
                   geo=dict(type="cost_based", rate=0.3, aggregation="mean"), size_type="quadratic")
    if want("synth_geo_sigmoid"):
        synth_case("synth_geo_sigmoid", 40, 8, 3, 2, 3, seed=5, T=1.1, Tp=1.0, n_steps=200,
                   geo=dict(type="cost_based", rate=0.2, aggregation="max", probability_function="sigmoid",
                            inflection_point=0.4), family_prior="symmetric_dirichlet")
    if want("synth_tiny_all"):
        synth_case("synth_tiny_all", 8, 5, 3, 1, 2, seed=4, T=1.0, Tp=1.0, n_steps=200, min_size=1, max_size=6,
                   opcfg=dict(clusters=40, weights=10, source=10), per_cluster=2)
    if want("synth_c2_width"):
        synth_case("synth_c2_width", 200, 100, 4, 4, 6, seed=6, T=1.0, Tp=1.0, n_steps=120, max_size=60)
