"""fp64 numpy restatement of the ten scheduled sBayes MCMC operators.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each function cites the reference code it
follows.  Differences from the reference, all deliberate and documented in DESIGN.md:

* float64 throughout (the reference rounds proposal probabilities to float32);
* products over features are accumulated as sums of logs; where the reference then adds
  the float32 EPS in *linear* space (AlterClusterWide, ClusterJump) the sum is
  exponentiated first so the degenerate wide-feature behaviour is reproduced, and where it
  divides m1/(m0+m1) without EPS (AlterCluster) the logistic of the log difference is
  used, which is finite where the reference underflows to NaN.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
from scipy.special import gammaln, expit

from .problem import Problem, ChainState, ChainConfig, NA
from . import model as M
from .draws import (SITE_CLUSTER, SITE_GROW, SITE_OBJECT, SITE_COMPONENT, SITE_JUMP, SITE_WPAIR,
                    cdf_pick)

NEG_INF = float("-inf")
Q_GIBBS, Q_BACK_GIBBS = NEG_INF, 0.0          # operators.py:241-243
Q_REJECT, Q_BACK_REJECT = 0.0, NEG_INF        # operators.py:245-247
EPS = M.EPS32

# operator ids (problem.OP_NAMES)
(OP_NAIVE_N1, OP_NAIVE_N1_GEO, OP_NAIVE_N2_GEO, OP_GIBBSISH, OP_GIBBSISH_GEO, OP_WIDE, OP_JUMP,
 OP_SRC_RANDOM, OP_SRC_GROUPS, OP_WEIGHTS) = range(10)

NB_EVERYWHERE, NB_DIRECT, NB_TWOSTEP = 0, 1, 2

ADDITIVE_SMOOTHING = 1e-6     # operators.py:978
W_STAY = 0.15                 # operators.py:133
GEO_SCALER = 2.0              # operators.py:1390
SRC_MIN_SIZE = 10             # operators.py:440
SRC_MAX_RANDOM = 20           # operators.py:185
SRC_MAX_GROUPS = 30           # operators.py:194


@dataclass
class Proposal:
    state: ChainState
    log_q: float
    log_q_back: float
    changed: bool = True          # False for direct rejects (candidate is the input state)
    info: dict = field(default_factory=dict)


def direct_reject(st: ChainState, **info) -> Proposal:
    return Proposal(st, Q_REJECT, Q_BACK_REJECT, changed=False, info=info)


def heat_binary_probability(p, T):
    """sbayes/util.py:1457-1469."""
    pw = 1.0 / T
    a = np.asarray(p, dtype=np.float64) ** pw
    return a / (a + (1.0 - np.asarray(p, dtype=np.float64)) ** pw)


# ---------------------------------------------------------------------------------------
# shared pieces
# ---------------------------------------------------------------------------------------

def neighbours(prob: Problem, mask: np.ndarray) -> np.ndarray:
    """adj.dot(mask) > 0 on the CSR adjacency (sbayes/util.py:121-143)."""
    out = np.zeros(prob.N, dtype=bool)
    ip, ix = prob.adj_indptr, prob.adj_indices
    for o in range(prob.N):
        nb = ix[ip[o]:ip[o + 1]]
        out[o] = bool(np.any(mask[nb]))
    return out


def grow_candidates(prob: Problem, st: ChainState, k: int, neighbourhood: int) -> np.ndarray:
    """AlterCluster.grow_candidates (operators.py:1097-1107), bool [N]."""
    occupied = st.cluster_of >= 0
    if neighbourhood == NB_EVERYWHERE:
        return ~occupied
    reach = neighbours(prob, st.cluster_of == k)
    if neighbourhood == NB_TWOSTEP:
        reach = neighbours(prob, reach)
    return reach & ~occupied


def weights_with_and_without(prob: Problem, st: ChainState, objs: np.ndarray, Tp: float):
    """compute_feature_weights_with_and_without (operators.py:1075-1095):
    w_z[o,f,c] = h_z[o,c] W[f,c]^(1/Tp) / sum_c' h_z[o,c'] W[f,c']^(1/Tp),  h_z[o,0] = z."""
    has = M.has_components(prob, st)[objs]
    h0, h1 = has.copy(), has.copy()
    h0[:, 0] = False
    h1[:, 0] = True
    return (M.normalize_weights(st.W, h0, 1.0 / Tp), M.normalize_weights(st.W, h1, 1.0 / Tp))


def cluster_log_marginals(prob: Problem, st: ChainState, k: int, objs: np.ndarray,
                          T: float, Tp: float, wide: bool = False):
    """log m_0[o], log m_1[o] of compute_cluster_posterior / compute_raw_cluster_probs
    (operators.py:1048-1061, 1441-1451): column 0 of the plain cached component
    likelihoods is replaced by the tempered table of cluster k (an extra 1/T power in the
    wide operator, operators.py:1444)."""
    plain = M.plain_table(prob, st.counts)
    temp = M.tempered_table(prob, st.counts, T, Tp)
    ell = M.gather_table(prob, st, plain, objs)
    col0 = M.gather_table(prob, st, temp, objs, force_cluster=k)[:, :, 0]
    if wide:
        col0 = col0 ** (1.0 / T)
    ell[:, :, 0] = col0
    w0, w1 = weights_with_and_without(prob, st, objs, Tp)
    with np.errstate(divide="ignore"):
        lm0 = np.log((ell * w0).sum(axis=-1)).sum(axis=-1) / T
        lm1 = np.log((ell * w1).sum(axis=-1)).sum(axis=-1) / T
    return lm0, lm1


def cluster_posterior(prob: Problem, st: ChainState, k: int, objs: np.ndarray, cfg: ChainConfig,
                      gibbsish: bool, use_geo: bool) -> np.ndarray:
    """AlterCluster.compute_cluster_posterior (operators.py:1035-1073), f64 [len(objs)]."""
    if cfg.sample_from_prior or not gibbsish:
        return np.full(len(objs), 0.5)
    T, Tp = cfg.temperature, cfg.prior_temperature
    lm0, lm1 = cluster_log_marginals(prob, st, k, objs, T, Tp)
    if use_geo and prob.geo is not None:
        lm1 = lm1 + M.geo_costs_per_object(prob, st, k)[objs] / Tp
    post = expit(lm1 - lm0)
    a = ADDITIVE_SMOOTHING
    return (post + a) / (1.0 + 2.0 * a)


def tables_given_unchanged(prob: Problem, st_old: ChainState, S: np.ndarray, T: float, Tp: float):
    """Tempered tables built from the counts with the observations of the changed objects
    S removed (component_likelihood_given_unchanged, operators.py:863-928)."""
    cnt_u = st_old.counts - M.rebuild_counts(prob, st_old, subset=S)
    return M.tempered_table(prob, cnt_u, T, Tp)


def lik_given_unchanged(prob: Problem, st_membership: ChainState, table_u: np.ndarray, S: np.ndarray,
                        k: int, T: float) -> np.ndarray:
    """l^U[o,f,c] for o in S: every object is scored against cluster k's row in column 0
    (operators.py:887), against its own group's row for confounders; NA -> 1; ^(1/T)."""
    ell = M.gather_table(prob, st_membership, table_u, S, force_cluster=k)
    return ell ** (1.0 / T)


def sample_sources(p: np.ndarray, u: np.ndarray, na: np.ndarray) -> np.ndarray:
    """sample_categorical (sbayes/preprocessing.py:224-256): first c with u < cdf[c];
    NA observations get source 255.  p [M,F,C] -> uint8 [M,F]."""
    cdf = np.cumsum(p, axis=-1)
    cdf = cdf / cdf[..., [-1]]
    new = np.argmax(u[..., None] < cdf, axis=-1).astype(np.uint8)
    new[na] = NA
    return new


def log_prob_of(p: np.ndarray, src: np.ndarray) -> float:
    """sum log p[o,f,src[o,f]] over non-NA observations."""
    valid = src != NA
    pick = np.take_along_axis(p, np.where(valid, src, 0)[..., None].astype(np.int64), axis=-1)[..., 0]
    with np.errstate(divide="ignore"):
        return float(np.log(pick[valid]).sum())


def normalize_last(q: np.ndarray) -> np.ndarray:
    return q / q.sum(axis=-1, keepdims=True)


def resample_sources_cluster_op(prob: Problem, st_old: ChainState, st_new: ChainState, S: np.ndarray,
                                k_new: int, k_old: int, cfg: ChainConfig, draws, jump: bool):
    """ClusterOperator.gibbs_sample_source (operators.py:796-849) and
    ClusterJump.gibbs_sample_source_jump (operators.py:1775-1827).

    `st_new` holds the new cluster membership but still the old sources and counts; it is
    updated in place.  Returns (log_q, log_q_back)."""
    T, Tp = cfg.temperature, cfg.prior_temperature
    S = np.sort(np.asarray(S, dtype=np.int64))
    na = prob.x[S] == NA
    w_new = M.normalize_weights(st_new.W, M.has_components(prob, st_new)[S])
    w_old = M.normalize_weights(st_old.W, M.has_components(prob, st_old)[S])
    if cfg.sample_from_prior:
        if jump:
            p, p_back = w_new, w_new
        else:
            p, p_back = w_new ** (1.0 / Tp), w_old ** (1.0 / Tp)
    else:
        table_u = tables_given_unchanged(prob, st_old, S, T, Tp)
        lh_new = lik_given_unchanged(prob, st_new, table_u, S, k_new, T)
        if jump:
            lh_old = lik_given_unchanged(prob, st_old, table_u, S, k_old, T)
            p = normalize_last(w_new * lh_new)
            p_back = normalize_last(w_new * lh_old)
        else:
            p = normalize_last(w_new ** (1.0 / Tp) * lh_new)
            p_back = normalize_last(w_old ** (1.0 / Tp) * lh_new)

    u = draws.gibbs_uniforms(len(S), prob.F)
    old_contrib = M.rebuild_counts(prob, st_old, subset=S)
    src_old = st_old.src[S].copy()
    st_new.src[S] = sample_sources(p, u, na)
    st_new.counts = st_old.counts - old_contrib + M.rebuild_counts(prob, st_new, subset=S)
    log_q = log_prob_of(p, st_new.src[S])
    log_q_back = log_prob_of(p_back, src_old)
    return log_q, log_q_back


# ---------------------------------------------------------------------------------------
# AlterCluster: grow / shrink (operators.py:969-1249)
# ---------------------------------------------------------------------------------------

def alter_cluster(prob: Problem, st: ChainState, cfg: ChainConfig, draws, gibbsish: bool,
                  use_geo: bool, neighbourhood: int) -> Proposal:
    T = cfg.temperature
    k = draws.index(SITE_CLUSTER, 0, prob.K)
    size = int(np.count_nonzero(st.cluster_of == k))
    log_q_back_extra = 0.0
    if size == prob.min_size:
        grow = True
        log_q_back_extra = -np.log(2.0)
    elif size == prob.max_size:
        grow = False
        log_q_back_extra = -np.log(2.0)
    else:
        grow = draws.uniform(SITE_GROW) < cfg.p_grow

    info = {"k": k, "grow": bool(grow)}
    if grow:
        cand = grow_candidates(prob, st, k, neighbourhood)
        if not cand.any() or size == prob.max_size:
            return direct_reject(st, **info)
        cand_idx = np.flatnonzero(cand)
        post = heat_binary_probability(cluster_posterior(prob, st, k, cand_idx, cfg, gibbsish, use_geo), T)
        if post.sum() == 0:
            return direct_reject(st, **info)
        p_add = post / post.sum()
        o = int(cand_idx[cdf_pick(p_add, draws.uniform(SITE_OBJECT))])
        log_q_obj = np.log(p_add[np.searchsorted(cand_idx, o)])

        new = st.copy()
        new.cluster_of[o] = k
        lq_s, lqb_s = resample_sources_cluster_op(prob, st, new, [o], k, k, cfg, draws, jump=False)

        members = np.flatnonzero(new.cluster_of == k)
        post_b = heat_binary_probability(cluster_posterior(prob, new, k, members, cfg, gibbsish, use_geo), T)
        p_remove = (1.0 - post_b) / (1.0 - post_b).sum()
        log_q_back_obj = np.log(p_remove[np.searchsorted(members, o)])
        log_q = log_q_obj + lq_s + np.log(cfg.p_grow)
        log_q_back = log_q_back_obj + lqb_s + np.log(1.0 - cfg.p_grow) + log_q_back_extra
    else:
        members = np.flatnonzero(st.cluster_of == k)
        if members.size == prob.min_size:
            return direct_reject(st, **info)
        post = heat_binary_probability(cluster_posterior(prob, st, k, members, cfg, gibbsish, use_geo), T)
        xr = 1.0 - post
        if xr.sum() == 0:
            return direct_reject(st, **info)
        p_remove = xr / xr.sum()
        o = int(members[cdf_pick(p_remove, draws.uniform(SITE_OBJECT))])
        log_q_obj = np.log(p_remove[np.searchsorted(members, o)])

        new = st.copy()
        new.cluster_of[o] = -1
        lq_s, lqb_s = resample_sources_cluster_op(prob, st, new, [o], k, k, cfg, draws, jump=False)

        cand = grow_candidates(prob, new, k, neighbourhood)
        if not cand[o]:
            return direct_reject(st, **info)
        cand_idx = np.flatnonzero(cand)
        post_b = heat_binary_probability(cluster_posterior(prob, new, k, cand_idx, cfg, gibbsish, use_geo), T)
        p_add = post_b / post_b.sum()
        log_q_back_obj = np.log(p_add[np.searchsorted(cand_idx, o)])
        log_q = log_q_obj + lq_s + np.log(1.0 - cfg.p_grow)
        log_q_back = log_q_back_obj + lqb_s + np.log(cfg.p_grow) + log_q_back_extra

    info["object"] = o
    return Proposal(new, float(log_q), float(log_q_back), info=info)


# ---------------------------------------------------------------------------------------
# AlterClusterWide (operators.py:1382-1568)
# ---------------------------------------------------------------------------------------

def wide_cluster_probs(prob: Problem, st: ChainState, k: int, avail: np.ndarray, cfg: ChainConfig,
                       use_geo: bool) -> np.ndarray:
    """compute_cluster_probs (operators.py:1399-1418) over the available objects."""
    T, Tp = cfg.temperature, cfg.prior_temperature
    eps = 0.01 / prob.N                       # operators.py:135
    if cfg.sample_from_prior:
        raw = np.full(len(avail), 0.5)
    else:
        lm0, lm1 = cluster_log_marginals(prob, st, k, avail, T, Tp, wide=True)
        m0, m1 = np.exp(lm0), np.exp(lm1)
        if use_geo and prob.geo is not None:
            m1 = m1 * np.exp(M.geo_costs_per_object(prob, st, k)[avail] / Tp / GEO_SCALER)
        raw = m1 / (m0 + m1 + EPS)
    member = (st.cluster_of[avail] == k).astype(np.float64)
    p = raw + EPS
    p = p / p.sum()
    p = p + eps
    p = (1.0 - W_STAY) * (p / p.sum()) + W_STAY * (member / member.sum())
    old_size = member.sum()
    new_expected = p.sum()
    for _ in range(10):
        p = np.clip(p * old_size / new_expected, eps, 1.0 - eps)
        new_expected = p.sum()
        if new_expected > 0.975 * old_size:
            break
    return p


def alter_cluster_wide(prob: Problem, st: ChainState, cfg: ChainConfig, draws, use_geo: bool) -> Proposal:
    k = draws.index(SITE_CLUSTER, 0, prob.K)
    avail = np.flatnonzero((st.cluster_of < 0) | (st.cluster_of == k))
    old = st.cluster_of[avail] == k
    p = wide_cluster_probs(prob, st, k, avail, cfg, use_geo)

    attempt = 0
    while True:
        new_mem = draws.wide_uniforms(len(avail), attempt) < p
        attempt += 1
        if not np.array_equal(new_mem, old):
            break
    info = {"k": k, "attempts": attempt}
    n_new = int(np.count_nonzero(new_mem))
    if not (prob.min_size <= n_new <= prob.max_size):
        return direct_reject(st, **info)

    with np.errstate(divide="ignore"):
        log_q = float(np.log(np.where(new_mem, p, 1.0 - p)).sum())
        p_stand = np.prod(np.where(old, p, 1.0 - p))
        log_q -= np.log(1.0 - p_stand)

    new = st.copy()
    new.cluster_of[avail] = np.where(new_mem, k, -1)
    changed = avail[new_mem != old]
    lq_s, lqb_s = resample_sources_cluster_op(prob, st, new, changed, k, k, cfg, draws, jump=False)

    p_back = wide_cluster_probs(prob, new, k, avail, cfg, use_geo)
    with np.errstate(divide="ignore"):
        log_q_back = float(np.log(np.where(old, p_back, 1.0 - p_back)).sum())
        p_stand_back = np.prod(np.where(new_mem, p_back, 1.0 - p_back))
        log_q_back -= np.log(1.0 - p_stand_back)
    info["n_changed"] = int(len(changed))
    return Proposal(new, log_q + lq_s, log_q_back + lqb_s, info=info)


# ---------------------------------------------------------------------------------------
# ClusterJump (operators.py:1668-1827)
# ---------------------------------------------------------------------------------------

def jump_lh(prob: Problem, st: ChainState, k_src: int, k_tgt: int, cfg: ChainConfig) -> np.ndarray:
    """get_jump_lh (operators.py:1679-1722) over the members of cluster k_src."""
    T, Tp = cfg.temperature, cfg.prior_temperature
    members = np.flatnonzero(st.cluster_of == k_src)
    w = M.normalize_weights(st.W, M.has_components(prob, st)[members])
    wh = normalize_last(w ** (1.0 / Tp))
    temp = M.tempered_table(prob, st.counts, T, Tp)
    ell_src = M.gather_table(prob, st, temp, members, force_cluster=k_src)
    ell_tgt = M.gather_table(prob, st, temp, members, force_cluster=k_tgt)
    valid = prob.x[members] != NA
    conf = (wh[:, :, 1:] * ell_src[:, :, 1:]).sum(axis=-1)       # expected_confounder_features at x
    stay_f = conf + wh[:, :, 0] * ell_src[:, :, 0]
    jump_f = conf + wh[:, :, 0] * ell_tgt[:, :, 0]
    with np.errstate(divide="ignore"):
        ls = np.where(valid, np.log(np.where(valid, stay_f, 1.0)), 0.0).sum(axis=-1)
        lj = np.where(valid, np.log(np.where(valid, jump_f, 1.0)), 0.0).sum(axis=-1)
    lh_stay = np.exp(ls / T) + EPS
    lh_jump = np.exp(lj / T) + EPS
    return lh_jump / (lh_jump + lh_stay)


def cluster_jump(prob: Problem, st: ChainState, cfg: ChainConfig, draws) -> Proposal:
    k_src, k_tgt = draws.pair(SITE_JUMP, prob.K)
    sizes = st.sizes(prob.K)
    info = {"k_src": k_src, "k_tgt": k_tgt}
    if sizes[k_src] <= prob.min_size or sizes[k_tgt] >= prob.max_size:
        return direct_reject(st, **info)
    members = np.flatnonzero(st.cluster_of == k_src)
    # gibbsish=True in the schedule (operators.py:163-170); note that the reference does not
    # consult sample_from_prior here (operators.py:1744-1747)
    pj = jump_lh(prob, st, k_src, k_tgt, cfg)
    p_jump = pj / pj.sum()
    i_fwd = cdf_pick(p_jump, draws.uniform(SITE_OBJECT))
    o = int(members[i_fwd])

    new = st.copy()
    new.cluster_of[o] = k_tgt
    lq_s, lqb_s = resample_sources_cluster_op(prob, st, new, [o], k_tgt, k_src, cfg, draws, jump=True)

    members_b = np.flatnonzero(new.cluster_of == k_tgt)
    pb = jump_lh(prob, new, k_tgt, k_src, cfg)
    p_back = pb / pb.sum()
    i_back = int(np.searchsorted(members_b, o))
    info["object"] = o
    return Proposal(new, float(np.log(p_jump[i_fwd]) + lq_s), float(np.log(p_back[i_back]) + lqb_s), info=info)


# ---------------------------------------------------------------------------------------
# GibbsSampleSource (operators.py:423-579)
# ---------------------------------------------------------------------------------------

def source_posterior(prob: Problem, st: ChainState, S: np.ndarray, cfg: ChainConfig) -> np.ndarray:
    """calculate_source_posterior (operators.py:554-574) for objects S, f64 [M,F,C]."""
    T, Tp = cfg.temperature, cfg.prior_temperature
    w = M.normalize_weights(st.W, M.has_components(prob, st)[S])
    if cfg.sample_from_prior:
        return normalize_last(w ** (1.0 / Tp))
    ell = M.likelihood_per_component(prob, st, S)
    return normalize_last(ell ** (1.0 / T) * w ** (1.0 / Tp))


def gibbs_sweep(prob: Problem, st: ChainState, cfg: ChainConfig, step: int, block: int = 512):
    """GibbsSampleSource._propose with ObjectSelector.ALL (operators.py:496-552): every source of
    every object resampled from its conditional given the CURRENT counts; the uniforms are the
    Philox stream (cfg.seed, step, SITE_GIBBS, o * ceil(F/2) + f // 2), two features per call.  Returns (new state with rebuilt counts,
    log_q, log_q_back, number of reassigned observations)."""
    from .draws import PhiloxDraws, SITE_GIBBS
    from . import philox
    d = PhiloxDraws(cfg.seed)
    d.begin_step(step)
    if st.counts is None:
        st = ChainState(st.cluster_of, st.W, st.src, M.rebuild_counts(prob, st))
    new_src = np.empty_like(st.src)
    log_q = 0.0
    for a in range(0, prob.N, block):
        S = np.arange(a, min(prob.N, a + block))
        p = source_posterior(prob, st, S, cfg)
        # one Philox call per pair of adjacent features: words 0,1 -> even feature, words 2,3 -> odd feature
        hF = (prob.F + 1) // 2
        idx = (S[:, None].astype(np.int64) * hF + (np.arange(prob.F)[None, :] // 2)).reshape(-1)
        w = philox.words(d.seed, d.step, SITE_GIBBS, 0, idx)
        odd = np.tile(np.arange(prob.F) % 2 == 1, len(S))
        u = np.where(odd, philox.u01_double(w[2], w[3]), philox.u01_double(w[0], w[1])).reshape(len(S), prob.F)
        new = sample_sources(p, u, prob.na[S])
        new_src[S] = new
        log_q += log_prob_of(p, new)
    st_new = ChainState(st.cluster_of.copy(), st.W.copy(), new_src, None)
    st_new.counts = M.rebuild_counts(prob, st_new)
    log_q_back = 0.0
    for a in range(0, prob.N, block):
        S = np.arange(a, min(prob.N, a + block))
        log_q_back += log_prob_of(source_posterior(prob, st_new, S, cfg), st.src[S])
    n_changed = int(np.count_nonzero((new_src != st.src) & ~prob.na))
    return st_new, log_q, log_q_back, n_changed


def select_source_subset(prob: Problem, st: ChainState, draws, groups_mode: bool) -> np.ndarray:
    """select_object_subset (operators.py:461-494); returns sorted object indices."""
    N = prob.N
    if N <= SRC_MIN_SIZE:
        return np.arange(N)
    if not groups_mode:
        return draws.subset_of(np.arange(N), min(SRC_MAX_RANDOM, N))
    # components are enumerated as [confounders..., clusters] (state.py:626-634)
    n_comp = prob.C
    ci = draws.index(SITE_COMPONENT, 0, n_comp)
    if ci == n_comp - 1:
        g = draws.index(SITE_COMPONENT, 1, prob.K)
        members = np.flatnonzero(st.cluster_of == g)
    else:
        g = draws.index(SITE_COMPONENT, 1, prob.conf_ngroups[ci])
        members = np.flatnonzero(prob.conf_gid[ci] == g)
    kmax = min(SRC_MAX_GROUPS, N)
    if members.size > kmax:
        members = draws.subset_of(members, kmax)
    return members


def gibbs_sample_source(prob: Problem, st: ChainState, cfg: ChainConfig, draws, groups_mode: bool) -> Proposal:
    S = select_source_subset(prob, st, draws, groups_mode)
    na = prob.x[S] == NA
    p = source_posterior(prob, st, S, cfg)
    u = draws.gibbs_uniforms(len(S), prob.F)
    new = st.copy()
    new.src[S] = sample_sources(p, u, na)
    new.counts = st.counts - M.rebuild_counts(prob, st, subset=S) + M.rebuild_counts(prob, new, subset=S)
    log_q = log_prob_of(p, new.src[S])
    p_back = p if cfg.sample_from_prior else source_posterior(prob, new, S, cfg)
    log_q_back = log_prob_of(p_back, st.src[S])
    return Proposal(new, log_q, log_q_back, info={"subset": S})


# ---------------------------------------------------------------------------------------
# GibbsSampleWeights (operators.py:582-693)
# ---------------------------------------------------------------------------------------

def beta_logpdf(x, a, b):
    """scipy.stats.beta.logpdf: (a-1) log x + (b-1) log1p(-x) - betaln(a, b)."""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = np.where(a == 1.0, 0.0, (a - 1.0) * np.log(x))
        t2 = np.where(b == 1.0, 0.0, (b - 1.0) * np.log1p(-x))
    return t1 + t2 - (gammaln(a) + gammaln(b) - gammaln(a + b))


def source_lh_by_feature(prob: Problem, st: ChainState, W: np.ndarray) -> np.ndarray:
    """source_lh_by_feature (operators.py:678-686): sum_o log w~[o,f,src[o,f]] per feature."""
    w = M.normalize_weights(W, M.has_components(prob, st))
    valid = st.src != NA
    pick = np.take_along_axis(w, np.where(valid, st.src, 0)[..., None].astype(np.int64), axis=-1)[..., 0]
    with np.errstate(divide="ignore"):
        return np.where(valid, np.log(np.where(valid, pick, 1.0)), 0.0).sum(axis=0)


def rebalance_weights_f32(W: np.ndarray, i1: int, i2: int, a2: np.ndarray):
    """The float32 arithmetic of resample_weight_for_two_components (operators.py:664-672),
    reproduced operation by operation because W is chain *state*.  Returns (W_new f32, a2_old f64)."""
    W = W.astype(np.float32)
    w02 = (W[:, i1] + W[:, i2]).astype(np.float32)
    Wn = W.copy()
    Wn[:, i1] = ((1.0 - a2) * w02.astype(np.float64)).astype(np.float32)
    Wn[:, i2] = (a2 * w02.astype(np.float64)).astype(np.float32)
    s = np.zeros(W.shape[0], dtype=np.float32)
    for c in range(W.shape[1]):
        s = (s + Wn[:, c]).astype(np.float32)
    Wn = (Wn / s[:, None]).astype(np.float32)
    a2_old = (W[:, i2] / w02).astype(np.float32).astype(np.float64)
    return Wn, a2_old


def gibbs_sample_weights(prob: Problem, st: ChainState, cfg: ChainConfig, draws) -> Proposal:
    Tp = cfg.prior_temperature
    F = prob.F
    i1, i2 = draws.pair(SITE_WPAIR, prob.C)
    has = M.has_components(prob, st)
    both = has[:, i1] & has[:, i2]
    cnt = np.zeros((F, prob.C))
    for c in range(prob.C):
        cnt[:, c] = np.count_nonzero(st.src[both] == c, axis=0)
    cnt += prob.alpha_w
    c1, c2 = cnt[:, i1] / Tp, cnt[:, i2] / Tp
    a, b = 1.0 + c2, 1.0 + c1
    a2 = draws.beta(a, b)
    W_new, a2_old = rebalance_weights_f32(st.W, i1, i2, a2)
    log_q = beta_logpdf(a2, a, b)
    log_q_back = beta_logpdf(a2_old, a, b)

    lp_old = source_lh_by_feature(prob, st, st.W) + M.weights_prior_pointwise(prob, st.W)
    lp_new = source_lh_by_feature(prob, st, W_new) + M.weights_prior_pointwise(prob, W_new)
    with np.errstate(over="ignore", invalid="ignore"):
        p_accept = np.exp((lp_new - lp_old + log_q_back - log_q) / Tp)
    accept = draws.accept_uniforms_f32(F) < p_accept
    new = st.copy()
    new.W = np.where(accept[:, None], W_new, st.W).astype(np.float32)
    return Proposal(new, Q_GIBBS, Q_BACK_GIBBS,
                    info={"pair": (i1, i2), "accept_rate": float(accept.mean()), "accept": accept})


# ---------------------------------------------------------------------------------------
# dispatch (operators.py:28-209)
# ---------------------------------------------------------------------------------------

def propose(prob: Problem, st: ChainState, cfg: ChainConfig, draws, op: int) -> Proposal:
    geo = prob.geo is not None
    if op == OP_NAIVE_N1:
        return alter_cluster(prob, st, cfg, draws, False, False, NB_DIRECT)
    if op == OP_NAIVE_N1_GEO:
        return alter_cluster(prob, st, cfg, draws, False, geo, NB_DIRECT)
    if op == OP_NAIVE_N2_GEO:
        return alter_cluster(prob, st, cfg, draws, False, geo, NB_TWOSTEP)
    if op == OP_GIBBSISH:
        return alter_cluster(prob, st, cfg, draws, True, False, NB_EVERYWHERE)
    if op == OP_GIBBSISH_GEO:
        return alter_cluster(prob, st, cfg, draws, True, geo, NB_EVERYWHERE)
    if op == OP_WIDE:
        return alter_cluster_wide(prob, st, cfg, draws, geo)
    if op == OP_JUMP:
        return cluster_jump(prob, st, cfg, draws)
    if op == OP_SRC_RANDOM:
        return gibbs_sample_source(prob, st, cfg, draws, groups_mode=False)
    if op == OP_SRC_GROUPS:
        return gibbs_sample_source(prob, st, cfg, draws, groups_mode=True)
    if op == OP_WEIGHTS:
        return gibbs_sample_weights(prob, st, cfg, draws)
    raise ValueError(op)
