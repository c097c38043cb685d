"""Conversions between the reference's objects (`Data`, `Model`, `Sample`) and the packed
containers the CUDA path uploads (`packed.Problem`, `packed.ChainState`).

Host-side only, no compute.  Nothing here imports `sbayes`: the objects are the caller's
(the reference package the user already runs) and are read by attribute, so this module
works with whatever sBayes installation the drop-in is plugged into.

What is read from where:
  Data.features.values   bool [N,F,S] one-hot, all-zero row = NA   sbayes/load_data.py:85-135
  Data.features.states   bool [F,S] applicable states               sbayes/load_data.py:90
  Data.confounders       name -> Confounder(group_assignment bool [G,N])  sbayes/load_data.py:138-184
  Data.network.adj_mat   scipy sparse adjacency                      sbayes/preprocessing.py:136-141
  Model.prior.*          concentration arrays, size / geo / weights prior settings
                         sbayes/model/prior.py:155-255,325-354,438-455,522-558,614-681,704-853
  Sample                 clusters bool [K,N], weights f32 [F,C], source bool [N,F,C],
                         feature_counts[name] f32 [G,F,S]            sbayes/sampling/state.py:492-552
"""
from __future__ import annotations

import numpy as np

from . import packed as P


def problem_from_reference(data, model) -> P.Problem:
    """Packed `Problem` equivalent to a reference (Data, Model) pair."""
    prior = model.prior
    x = P.pack_onehot(np.asarray(data.features.values))
    conf_gid, conf_ng, alpha_conf, names = [], [], [], []
    for name, conf in data.confounders.items():
        conf_gid.append(P.gid_from_bool(np.asarray(conf.group_assignment)))
        conf_ng.append(int(conf.n_groups))
        cp = prior.prior_confounding_effects[name]
        if cp.any_dynamic_priors:
            # rejected by the reference's own config validation (sbayes/config/config.py:226-232)
            raise NotImplementedError("hierarchical 'universal' prior type is not a valid sBayes configuration")
        alpha_conf.append(np.array(cp._concentration_array, dtype=np.float64))
        names.append(name)
    sp = prior.size_prior
    size_type = {sp.PriorType.UNIFORM_AREA: P.SIZE_UNIFORM_AREA, sp.PriorType.UNIFORM_SIZE: P.SIZE_UNIFORM_SIZE,
                 sp.PriorType.QUADRATIC_SIZE: P.SIZE_QUADRATIC}[sp.prior_type]
    gp = prior.geo_prior
    geo = None
    if gp.prior_type is gp.PriorTypes.COST_BASED:
        GC = type(gp.config)
        if gp.config.skeleton is not GC.Skeleton.MST:
            raise NotImplementedError("sbayes_b200: only the MST skeleton of the cost-based geo prior is supported")
        geo = P.GeoPriorSpec(
            cost=np.array(gp.cost_matrix, dtype=np.float64),
            aggregation={GC.AggregationStrategies.MEAN: P.AGG_MEAN, GC.AggregationStrategies.SUM: P.AGG_SUM,
                         GC.AggregationStrategies.MAX: P.AGG_MAX}[gp.aggregation_policy],
            prob_fn={GC.ProbabilityFunction.EXPONENTIAL: P.PF_EXPONENTIAL,
                     GC.ProbabilityFunction.SIGMOID: P.PF_SIGMOID}[gp.probability_function],
            rate=float(gp.scale), inflection=float(gp.inflection_point or 0.0))
    elif gp.prior_type is not gp.PriorTypes.UNIFORM:
        raise NotImplementedError(f"sbayes_b200: geo prior {gp.prior_type} is not supported")
    wp = prior.prior_weights
    adj = data.network.adj_mat.tocsr()
    return P.Problem(
        x=x, applicable=np.array(data.features.states, dtype=bool),
        conf_gid=conf_gid, conf_ngroups=conf_ng,
        alpha_cluster=np.array(prior.prior_cluster_effect.concentration_array, dtype=np.float64),
        alpha_conf=alpha_conf,
        alpha_w=np.array(wp.concentration_array, dtype=np.float64),
        n_clusters=int(model.n_clusters), min_size=int(model.min_size), max_size=int(model.max_size),
        size_prior=size_type, weights_prior_uniform=(wp.prior_type is wp.PriorType.UNIFORM),
        geo=geo, adj_indptr=np.array(adj.indptr, dtype=np.int32), adj_indices=np.array(adj.indices, dtype=np.int32),
        conf_names=names)


def state_from_sample(prob: P.Problem, sample) -> P.ChainState:
    """Reference `Sample` -> packed `ChainState` (counts stacked [G_tot,F,S] int32)."""
    src = P.pack_onehot(np.asarray(sample.source.value))
    parts = [sample.feature_counts["clusters"].value]
    parts += [sample.feature_counts[n].value for n in prob.conf_names]
    counts = np.concatenate(parts, axis=0)
    if not np.all(counts == np.round(counts)):
        raise ValueError("feature counts of the sample are not integers")
    return P.ChainState(cluster_of=P.cluster_of_from_bool(np.asarray(sample.clusters.value)),
                        W=np.array(sample.weights.value, dtype=np.float32),
                        src=src, counts=counts.astype(np.int32))


def sample_from_state(prob: P.Problem, model, data, st: P.ChainState, sample_cls, i_step: int = 0, chain: int = 0):
    """Packed `ChainState` (with counts) -> reference `Sample` with fresh, outdated caches.
    `sample_cls` is the reference's `Sample` class (e.g. `type(initial_sample)`)."""
    if st.counts is None:
        raise ValueError("state has no counts; read it back with DeviceContext.get_state")
    offs = prob.group_offsets()
    fc = {"clusters": st.counts[offs[0]:offs[1]].astype(np.float32)}
    for i, n in enumerate(prob.conf_names):
        fc[n] = st.counts[offs[i + 1]:offs[i + 2]].astype(np.float32)
    s = sample_cls.from_numpy_arrays(
        clusters=st.clusters_bool(prob.K), weights=st.W.copy(), confounders=data.confounders,
        source=P.unpack_onehot(st.src, prob.C), feature_counts=fc, model_shapes=model.shapes, chain=chain)
    s.i_step = i_step
    s.everything_changed()
    return s


def fill_cache_from_device(prob: P.Problem, sample, exp: dict, sample_from_prior: bool = False):
    """Bring the reference's cache nodes of `sample` up to date with the values the DEVICE computed
    (`DeviceContext.export_samples`), instead of re-evaluating `model.prior` / `model.likelihood` on the CPU:
    `cluster_size_prior`, `geo_prior` [K], `weights_prior`, `source_prior` [N] (read by ParametersCSVLogger,
    sbayes/sampling/loggers.py:249-252, and by `swap_chains` through `Model.prior`, sbayes/mcmc_setup.py:389-390)
    and `group_likelihoods[component]` (what `Likelihood.__call__` sums, sbayes/model/likelihood.py:47-101).
    `CacheNode.update_value` stamps a node with the current versions of its inputs (state.py:264-273)."""
    c = sample.cache
    post = exp["post"]
    c.cluster_size_prior.update_value(post["size_prior"])
    c.weights_prior.update_value(post["weights_prior"])
    c.source_prior.update_value(np.array(exp["source_prior"], dtype=np.float64))
    if prob.geo is not None:
        c.geo_prior.update_value(np.array(exp["geo_prior"], dtype=np.float64))
    if not sample_from_prior:
        offs = prob.group_offsets()
        names = ["clusters"] + list(prob.conf_names)
        for i, name in enumerate(names):
            c.group_likelihoods[name].update_value(np.array(exp["group_ll"][offs[i]:offs[i + 1]], dtype=np.float64))
    sample.last_lh = 0.0 if sample_from_prior else post["ll"]
    sample.last_prior = post["prior"]
    return sample
