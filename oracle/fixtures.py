"""Golden-fixture (de)serialisation: problems, states, draw tapes and reference outputs.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The flat tape layout written here is also
what the CUDA path's replay mode consumes (`sb_run` with a tape, include/sbayes_b200.h):

  tape_vals  f64 [total]              all recorded draws of all steps, concatenated
  tape_off   i32 [n_steps, N_SITES]   offset of a site's payload in tape_vals, -1 = absent
  tape_len   i32 [n_steps, N_SITES]   payload length (SITE_WIDE: attempts * n_avail)
"""
from __future__ import annotations

import numpy as np

from .problem import Problem, ChainState, ChainConfig, GeoPriorSpec
from .draws import N_SITES, SITE_WIDE, SITE_OP


def flatten_tapes(tapes: list) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    vals, off, ln = [], np.full((len(tapes), N_SITES), -1, np.int32), np.zeros((len(tapes), N_SITES), np.int32)
    pos = 0
    for i, t in enumerate(tapes):
        for site, payload in t.items():
            if site == SITE_WIDE:
                v = np.concatenate([np.asarray(a, dtype=np.float64).reshape(-1) for a in payload])
            else:
                v = np.asarray(payload, dtype=np.float64).reshape(-1)
            off[i, site], ln[i, site] = pos, v.size
            vals.append(v)
            pos += v.size
    return (np.concatenate(vals) if vals else np.zeros(0)), off, ln


def unflatten_tapes(vals, off, ln, wide_rows=None) -> list:
    """Inverse of flatten_tapes.  SITE_WIDE payloads come back as a list with one array
    per attempt when `wide_rows[i]` (n_avail of step i) is known, else a single attempt."""
    tapes = []
    for i in range(off.shape[0]):
        t = {}
        for site in range(N_SITES):
            if off[i, site] < 0:
                continue
            v = vals[off[i, site]: off[i, site] + ln[i, site]]
            if site == SITE_WIDE:
                n = int(wide_rows[i]) if wide_rows is not None and wide_rows[i] > 0 else v.size
                t[site] = [v[a * n:(a + 1) * n] for a in range(max(1, v.size // n))]
            elif site == SITE_OP:
                t[site] = int(v[0])
            else:
                t[site] = v
        tapes.append(t)
    return tapes


def problem_to_dict(prob: Problem, prefix="p_") -> dict:
    d = {
        "x": prob.x, "applicable": prob.applicable,
        "conf_gid": np.stack(prob.conf_gid) if prob.conf_gid else np.zeros((0, prob.N), np.int32),
        "conf_ngroups": np.array(prob.conf_ngroups, np.int32),
        "alpha_cluster": prob.alpha_cluster,
        "alpha_conf": np.concatenate(prob.alpha_conf, axis=0) if prob.alpha_conf else np.zeros((0, prob.F, prob.S)),
        "alpha_w": prob.alpha_w,
        "scalars": np.array([prob.n_clusters, prob.min_size, prob.max_size, prob.size_prior,
                             int(prob.weights_prior_uniform), int(prob.geo is not None)], np.int64),
        "adj_indptr": prob.adj_indptr if prob.adj_indptr is not None else np.zeros(prob.N + 1, np.int32),
        "adj_indices": prob.adj_indices if prob.adj_indices is not None else np.zeros(0, np.int32),
        "conf_names": np.array(prob.conf_names),
    }
    if prob.geo is not None:
        d["geo_cost"] = prob.geo.cost
        d["geo_params"] = np.array([prob.geo.aggregation, prob.geo.prob_fn, prob.geo.rate, prob.geo.inflection], np.float64)
    return {prefix + k: v for k, v in d.items()}


def problem_from_dict(d, prefix="p_") -> Problem:
    g = lambda k: d[prefix + k]
    sc = g("scalars")
    ng = [int(v) for v in g("conf_ngroups")]
    ac, pos = [], 0
    for n in ng:
        ac.append(np.array(g("alpha_conf")[pos:pos + n], dtype=np.float64))
        pos += n
    geo = None
    if int(sc[5]):
        gp = g("geo_params")
        geo = GeoPriorSpec(cost=np.array(g("geo_cost"), np.float64), aggregation=int(gp[0]), prob_fn=int(gp[1]),
                           rate=float(gp[2]), inflection=float(gp[3]))
    return Problem(
        x=np.array(g("x"), np.uint8), applicable=np.array(g("applicable"), bool),
        conf_gid=[np.array(r, np.int32) for r in g("conf_gid")], conf_ngroups=ng,
        alpha_cluster=np.array(g("alpha_cluster"), np.float64), alpha_conf=ac,
        alpha_w=np.array(g("alpha_w"), np.float64), n_clusters=int(sc[0]), min_size=int(sc[1]),
        max_size=int(sc[2]), size_prior=int(sc[3]), weights_prior_uniform=bool(sc[4]), geo=geo,
        adj_indptr=np.array(g("adj_indptr"), np.int32), adj_indices=np.array(g("adj_indices"), np.int32),
        conf_names=[str(s) for s in g("conf_names")])


def state_to_dict(st: ChainState, prefix) -> dict:
    return {prefix + "cluster_of": st.cluster_of, prefix + "W": st.W, prefix + "src": st.src,
            prefix + "counts": st.counts}


def state_from_dict(d, prefix) -> ChainState:
    return ChainState(np.array(d[prefix + "cluster_of"], np.int32), np.array(d[prefix + "W"], np.float32),
                      np.array(d[prefix + "src"], np.uint8), np.array(d[prefix + "counts"], np.int32))


def cfg_to_array(cfg: ChainConfig) -> np.ndarray:
    return np.array([cfg.temperature, cfg.prior_temperature, float(cfg.sample_from_prior), cfg.op_clusters,
                     cfg.op_weights, cfg.op_source, cfg.p_grow, float(cfg.seed)], np.float64)


def cfg_from_array(a) -> ChainConfig:
    return ChainConfig(temperature=float(a[0]), prior_temperature=float(a[1]), sample_from_prior=bool(a[2]),
                       op_clusters=float(a[3]), op_weights=float(a[4]), op_source=float(a[5]), p_grow=float(a[6]),
                       seed=int(a[7]))


OUT_KEYS = ["op", "log_q", "log_q_back", "accept", "ll_new", "prior_new", "ll", "prior"]


def save_replay(path, prob, cfg, st0, st_final, tapes, outs, wide_rows, extra=None):
    vals, off, ln = flatten_tapes(tapes)
    d = {}
    d.update(problem_to_dict(prob))
    d.update(state_to_dict(st0, "s0_"))
    d.update(state_to_dict(st_final, "s1_"))
    d["cfg"] = cfg_to_array(cfg)
    d["tape_vals"], d["tape_off"], d["tape_len"] = vals, off, ln
    d["wide_rows"] = np.array(wide_rows, np.int32)
    for k in OUT_KEYS:
        d["out_" + k] = np.array([o[k] for o in outs], dtype=np.float64)
    if extra:
        d.update(extra)
    np.savez_compressed(path, **d)


def load_replay(path):
    d = np.load(path, allow_pickle=False)
    prob = problem_from_dict(d)
    cfg = cfg_from_array(d["cfg"])
    st0, st1 = state_from_dict(d, "s0_"), state_from_dict(d, "s1_")
    tapes = unflatten_tapes(d["tape_vals"], d["tape_off"], d["tape_len"], d["wide_rows"])
    outs = {k: d["out_" + k] for k in OUT_KEYS}
    return dict(prob=prob, cfg=cfg, st0=st0, st1=st1, tapes=tapes, outs=outs, raw=d)
