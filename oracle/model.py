"""fp64 numpy restatement of the sBayes target density and the per-component tables.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Everything is evaluated in float64; the
reference rounds to float32 at several points, which is why parity
with the live reference is checked to ~1e-8 (likelihood) / ~1e-6 (priors, proposal
probabilities) while the CUDA path is checked against this file to 1e-10.
"""
from __future__ import annotations

import numpy as np
from scipy.special import gammaln

from .problem import (Problem, ChainState, NA, SIZE_UNIFORM_AREA, SIZE_UNIFORM_SIZE, SIZE_QUADRATIC,
                      AGG_MEAN, AGG_SUM, AGG_MAX, PF_EXPONENTIAL, PF_SIGMOID)

EPS32 = float(np.finfo(np.float32).eps)   # sbayes/util.py:34


# ---------------------------------------------------------------------------------------
# sufficient statistics (sbayes/sampling/counts.py:10-95)
# ---------------------------------------------------------------------------------------

def component_gid(prob: Problem, st: ChainState, c: int) -> np.ndarray:
    """group id of every object in mixture component c (component 0 = clusters)."""
    return st.cluster_of if c == 0 else prob.conf_gid[c - 1]


def rebuild_counts(prob: Problem, st: ChainState, subset=None) -> np.ndarray:
    """counts[g,f,s] = #{o in g : src[o,f]==c(g) and x[o,f]==s}, stacked over components.

    Restates compute_effect_counts / recalculate_feature_counts
    (sbayes/sampling/counts.py:10-52).  int32 [G_tot, F, S]."""
    N, F, S = prob.N, prob.F, prob.S
    offs = prob.group_offsets()
    out = np.zeros((prob.G_tot, F, S), dtype=np.int32)
    objs = np.arange(N) if subset is None else np.asarray(subset)
    x = prob.x[objs].astype(np.int64)
    src = st.src[objs]
    f_idx = np.arange(F, dtype=np.int64)[None, :]
    for c in range(prob.C):
        gid = component_gid(prob, st, c)[objs].astype(np.int64)
        m = (src == c) & (gid >= 0)[:, None] & (x != NA)
        flat = ((gid[:, None] + offs[c]) * F + f_idx) * S + x
        np.add.at(out.reshape(-1), flat[m], 1)
    return out


# ---------------------------------------------------------------------------------------
# likelihood (sbayes/model/likelihood.py:47-101, sbayes/util.py:1373-1394)
# ---------------------------------------------------------------------------------------

def dircat_logpdf(counts: np.ndarray, a: np.ndarray) -> np.ndarray:
    """Dirichlet-categorical log marginal per feature (last axis = states).

    lgamma(sum a) - lgamma(n + sum a) + sum_{a>0} [lgamma(c+a) - lgamma(a)]
    (sbayes/util.py:1390-1394; the reference casts the per-feature value to float32)."""
    counts = counts.astype(np.float64)
    n = counts.sum(axis=-1)
    sum_a = a.sum(axis=-1)
    const = gammaln(sum_a) - gammaln(n + sum_a)
    pos = a > 0
    a_safe = np.where(pos, a, 1.0)
    series = np.where(pos, gammaln(counts + a_safe) - gammaln(a_safe), 0.0)
    return const + series.sum(axis=-1)


def group_loglik(prob: Problem, counts: np.ndarray) -> np.ndarray:
    """per-group log-likelihood, f64 [G_tot] (cache.group_likelihoods in the reference)."""
    return dircat_logpdf(counts, prob.alpha_stacked()).sum(axis=-1)


def log_likelihood(prob: Problem, st: ChainState) -> float:
    """Likelihood.__call__ (sbayes/model/likelihood.py:47-63) with the state's counts."""
    return float(group_loglik(prob, st.counts).sum())


# ---------------------------------------------------------------------------------------
# weights (sbayes/model/likelihood.py:153-190, sbayes/sampling/state.py:353-376)
# ---------------------------------------------------------------------------------------

def has_components(prob: Problem, st: ChainState) -> np.ndarray:
    has = np.empty((prob.N, prob.C), dtype=bool)
    has[:, 0] = st.cluster_of >= 0
    for i, gid in enumerate(prob.conf_gid, start=1):
        has[:, i] = gid >= 0
    return has


def normalize_weights(W: np.ndarray, has: np.ndarray, power: float = 1.0) -> np.ndarray:
    """w[o,f,c] = has[o,c] W[f,c]^power / sum_c' has[o,c'] W[f,c']^power  (f64 [N,F,C])."""
    Wp = W.astype(np.float64) ** power
    w = has[:, None, :] * Wp[None, :, :]
    tot = w.sum(axis=-1, keepdims=True)
    # an object with no component at all has weight 0 everywhere (the reference yields NaN)
    return np.divide(w, tot, out=np.zeros_like(w), where=tot > 0)


# ---------------------------------------------------------------------------------------
# priors (sbayes/model/prior.py)
# ---------------------------------------------------------------------------------------

def source_prior_per_object(prob: Problem, st: ChainState) -> np.ndarray:
    """SourcePrior (sbayes/model/prior.py:568-611): sum_f log w~[o,f,src[o,f]] over valid f."""
    w = normalize_weights(st.W, has_components(prob, st))
    valid = st.src != NA
    pick = np.take_along_axis(w, np.where(valid, st.src, 0)[..., None].astype(np.int64), axis=-1)[..., 0]
    with np.errstate(divide="ignore"):
        lw = np.where(valid, np.log(np.where(valid, pick, 1.0)), 0.0)
    return lw.sum(axis=-1)


def log_multinom(n: int, ks) -> float:
    """log multinomial coefficient (sbayes/util.py:1125-1173), via lgamma."""
    ks = np.asarray(ks, dtype=np.int64)
    if ks.sum() == 0:
        return 0.0
    rest = n - ks.sum()
    return float(gammaln(n + 1) - gammaln(ks + 1).sum() - gammaln(rest + 1))


def size_prior(prob: Problem, st: ChainState) -> float:
    """ClusterSizePrior (sbayes/model/prior.py:647-681)."""
    sizes = st.sizes(prob.K)
    if prob.size_prior == SIZE_UNIFORM_SIZE:
        return -log_multinom(prob.N, sizes)
    if prob.size_prior == SIZE_QUADRATIC:
        with np.errstate(divide="ignore"):
            return float(-np.sum(np.log(sizes.astype(np.float64) ** 2)))
    return 0.0


def dirichlet_logpdf(x: np.ndarray, alpha: np.ndarray) -> np.ndarray:
    """scipy.stats.dirichlet._logpdf row-wise (sbayes/util.py:49): -lnB(alpha) + sum (alpha-1) log x."""
    x = x.astype(np.float64)
    lnB = gammaln(alpha).sum(axis=-1) - gammaln(alpha.sum(axis=-1))
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(alpha == 1.0, 0.0, (alpha - 1.0) * np.log(x))
    return -lnB + t.sum(axis=-1)


def weights_prior_pointwise(prob: Problem, W: np.ndarray) -> np.ndarray:
    """WeightsPrior.pointwise_prior (sbayes/model/prior.py:553-558): per-feature Dirichlet log-pdf."""
    return dirichlet_logpdf(W, prob.alpha_w)


def weights_prior(prob: Problem, st: ChainState) -> float:
    """WeightsPrior.__call__ (sbayes/model/prior.py:522-551): 0 for the uniform type."""
    if prob.weights_prior_uniform:
        return 0.0
    return float(weights_prior_pointwise(prob, st.W).sum())


def mst_edge_costs(cost: np.ndarray) -> np.ndarray:
    """Edge weights of a minimum spanning tree of the complete graph with weight matrix
    `cost` (Prim).  Restates compute_mst_distances (sbayes/model/prior.py:950-962), which
    returns the *non-zero* MST edge weights (scipy's sparse result drops explicit zeros),
    or a single 0 if there are none."""
    n = cost.shape[0]
    if n <= 1:
        return np.zeros(1)
    in_tree = np.zeros(n, dtype=bool)
    in_tree[0] = True
    best = cost[0].astype(np.float64).copy()
    best[0] = np.inf
    edges = []
    for _ in range(n - 1):
        j = int(np.argmin(np.where(in_tree, np.inf, best)))
        edges.append(best[j])
        in_tree[j] = True
        best = np.minimum(best, cost[j])
    edges = np.array([e for e in edges if e != 0.0 and np.isfinite(e)])
    if edges.size == 0:
        return np.zeros(1)
    return edges


def _aggregate(prob: Problem, d: np.ndarray) -> float:
    a = prob.geo.aggregation
    if a == AGG_MEAN:
        return float(np.mean(d))
    if a == AGG_SUM:
        return float(np.sum(d))
    return float(np.max(d))


def _log_expit(z):
    return -np.logaddexp(0.0, -np.asarray(z, dtype=np.float64))


def geo_prob_fn(prob: Problem, x):
    """GeoPrior.get_probability_function (sbayes/model/prior.py:754-767)."""
    g = prob.geo
    if g.prob_fn == PF_EXPONENTIAL:
        return -np.asarray(x, dtype=np.float64) / g.rate
    return _log_expit(-(np.asarray(x) - g.inflection) / g.rate) - _log_expit(g.inflection / g.rate)


def geo_prior_per_cluster(prob: Problem, st: ChainState) -> np.ndarray:
    """GeoPrior.__call__ (sbayes/model/prior.py:769-805), cost_based with MST skeleton."""
    out = np.zeros(prob.K)
    if prob.geo is None:
        return out
    for k in range(prob.K):
        members = np.flatnonzero(st.cluster_of == k)
        d = mst_edge_costs(prob.geo.cost[np.ix_(members, members)])
        out[k] = geo_prob_fn(prob, _aggregate(prob, d))
    return out


def geo_prior(prob: Problem, st: ChainState) -> float:
    return float(geo_prior_per_cluster(prob, st).sum())


def geo_costs_per_object(prob: Problem, st: ChainState, k: int) -> np.ndarray:
    """GeoPrior.get_costs_per_object (sbayes/model/prior.py:824-852): change of the geo
    prior of cluster k when each object is added.  f64 [N]; zeros for a uniform geo prior."""
    if prob.geo is None:
        return np.zeros(prob.N)
    members = np.flatnonzero(st.cluster_of == k)
    m = members.size
    cost_to_cluster = prob.geo.cost[members].min(axis=0)
    before = _aggregate(prob, mst_edge_costs(prob.geo.cost[np.ix_(members, members)]))
    a = prob.geo.aggregation
    if a == AGG_MEAN:
        after = (cost_to_cluster + m * before) / (1 + m)
    elif a == AGG_SUM:
        after = cost_to_cluster + before
    else:
        after = np.maximum(cost_to_cluster, before)
    return geo_prob_fn(prob, after) - geo_prob_fn(prob, before)


def prior_parts(prob: Problem, st: ChainState) -> dict:
    """Prior.__call__ (sbayes/model/prior.py:74-88) split into its four terms."""
    return {
        "size": size_prior(prob, st),
        "geo": geo_prior(prob, st),
        "weights": weights_prior(prob, st),
        "source": float(source_prior_per_object(prob, st).sum()),
    }


def log_prior(prob: Problem, st: ChainState) -> float:
    p = prior_parts(prob, st)
    return p["size"] + p["geo"] + p["weights"] + p["source"]


# ---------------------------------------------------------------------------------------
# effect tables (sbayes/sampling/conditionals.py:105-122,152-223)
# ---------------------------------------------------------------------------------------

def plain_table(prob: Problem, counts: np.ndarray) -> np.ndarray:
    """p[g,f,s] = norm_s(counts + alpha)   (conditionals.py:174-179,201-206)."""
    c = counts.astype(np.float64) + prob.alpha_stacked()
    return c / c.sum(axis=-1, keepdims=True)


def tempered_table(prob: Problem, counts: np.ndarray, T: float, Tp: float) -> np.ndarray:
    """p[g,f,s] = norm_s(u + (alpha-u)/Tp + counts/T)   (conditionals.py:105-122,
    operators.py:1255-1259).  `u` is the uniform concentration (1 on applicable states)
    for every component (operators.py:1352, prior.py:184-186)."""
    u = prob.unif[None]
    c = u + (prob.alpha_stacked() - u) / Tp + counts.astype(np.float64) / T
    return c / c.sum(axis=-1, keepdims=True)


def gather_table(prob: Problem, st: ChainState, table: np.ndarray, objs: np.ndarray,
                 force_cluster: int | None = None) -> np.ndarray:
    """l[o,f,c] = table[g_c(o), f, x[o,f]] for objects `objs`; 0 where o is in no group of
    c, 1 where the observation is NA (conditionals.py:152-223, likelihood.py:104-133).
    With `force_cluster=k` column 0 scores every object against cluster k's row."""
    F = prob.F
    offs = prob.group_offsets()
    x = prob.x[objs]
    valid = x != NA
    xs = np.where(valid, x, 0).astype(np.int64)
    f_idx = np.arange(F)[None, :]
    out = np.zeros((len(objs), F, prob.C))
    for c in range(prob.C):
        if c == 0 and force_cluster is not None:
            gid = np.full(len(objs), force_cluster, dtype=np.int64)
        else:
            gid = component_gid(prob, st, c)[objs].astype(np.int64)
        ing = gid >= 0
        rows = np.where(ing, gid, 0) + offs[c]
        val = table[rows[:, None], f_idx, xs]
        out[:, :, c] = np.where(ing[:, None], val, 0.0)
    out[~valid] = 1.0
    return out


def likelihood_per_component(prob: Problem, st: ChainState, objs=None) -> np.ndarray:
    """likelihood_per_component (conditionals.py:152-223): plain tables, f64 [M,F,C]."""
    objs = np.arange(prob.N) if objs is None else np.asarray(objs)
    return gather_table(prob, st, plain_table(prob, st.counts), objs)


# ---------------------------------------------------------------------------------------
# observation likelihood (what LikelihoodLogger writes, sbayes/sampling/loggers.py:354-359)
# ---------------------------------------------------------------------------------------

def observation_likelihood(prob: Problem, st: ChainState) -> np.ndarray:
    """lh[o,f] = sum_c w~[o,f,c] * l_c[o,f]  with the *leave-one-out* component likelihoods of
    `likelihood_per_component_exact` (sbayes/sampling/conditionals.py:300-367): the table row of
    the group of o in component c with o's own observation removed when o's source is c,
        l_c[o,f] = (counts[g,f,x] + alpha[g,f,x] - [src == c]) / (sum_s(counts + alpha)[g,f] - [src == c]),
    0 where o is in no group of c, 1 where the observation is missing (so lh = 1 there); the
    weights are `update_weights` (sbayes/model/likelihood.py:153-190).  f64 [N,F]."""
    N, F = prob.N, prob.F
    offs = prob.group_offsets()
    num = st.counts.astype(np.float64) + prob.alpha_stacked()          # [G_tot,F,S]
    den = num.sum(axis=-1)                                             # [G_tot,F]
    w = normalize_weights(st.W, has_components(prob, st))              # [N,F,C]
    valid = prob.x != NA
    xs = np.where(valid, prob.x, 0).astype(np.int64)
    f_idx = np.arange(F)[None, :]
    lh = np.zeros((N, F))
    for c in range(prob.C):
        gid = component_gid(prob, st, c).astype(np.int64)
        ing = gid >= 0
        rows = np.where(ing, gid, 0) + offs[c]
        d = (st.src == c).astype(np.float64)
        l = (num[rows[:, None], f_idx, xs] - d) / (den[rows[:, None], f_idx] - d)
        lh += np.where(ing[:, None], w[:, :, c] * l, 0.0)
    return np.where(valid, lh, 1.0)
