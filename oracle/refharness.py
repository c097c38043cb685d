"""Import the live sBayes reference (from /root/reference) and convert between its objects
and the packed containers of `oracle/problem.py`.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Only usable where /root/reference exists
(the build container).  Used by `tests/golden/make_golden.py` to generate the committed
fixtures and by the optional `-m ref` tests.  Recipe: SURVEY.md Appendix B.
"""
from __future__ import annotations

import os
import random
import sys
from collections import OrderedDict

import numpy as np

from . import problem as P

REFERENCE_ROOT = os.environ.get("SBAYES_REFERENCE", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refstubs")
_ref = None


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sbayes"))


def ref():
    """Import the reference once and return a namespace of its modules."""
    global _ref
    if _ref is not None:
        return _ref
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, _STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    warnings.filterwarnings("ignore")
    import sbayes.model as model                      # must come first (circular imports)
    import sbayes.sampling.state as state
    import sbayes.sampling.operators as operators
    import sbayes.sampling.counts as counts
    import sbayes.sampling.conditionals as conditionals
    import sbayes.sampling.mcmc_chain as mcmc_chain
    import sbayes.load_data as load_data
    import sbayes.config.config as config
    import sbayes.util as util
    import sbayes.preprocessing as preprocessing
    import sbayes.mcmc_setup as mcmc_setup
    import sbayes.model.prior as prior
    import sbayes.model.likelihood as likelihood

    class NS:
        pass
    ns = NS()
    for k, v in list(locals().items()):
        if k not in ("ns", "NS", "p", "warnings"):
            setattr(ns, k, v)
    _ref = ns
    return ns


def seed_reference(seed: int):
    """Make the reference bit-reproducible: seed the legacy numpy
    generator, Python's `random`, and rebind the module-level `RNG` names."""
    r = ref()
    np.random.seed(seed)
    random.seed(seed)
    rng = np.random.default_rng(seed)
    for mod in (r.util, r.operators, r.mcmc_setup):
        mod.RNG = rng
    try:
        import sbayes.sampling.initializers as ini
        ini.RNG = rng
    except Exception:
        pass
    return rng


def build_reference_model(x: np.ndarray, applicable: np.ndarray, locations: np.ndarray,
                          confounders: "OrderedDict[str, np.ndarray]", n_clusters: int,
                          prior_cfg: dict, group_names: dict | None = None):
    """Build reference `Data` and `Model` objects from packed arrays.

    x            uint8 [N,F] state index (255 = NA)
    applicable   bool [F,S]
    confounders  name -> bool [G,N] group assignment
    prior_cfg    dict in the format of the `model.prior` section of an sBayes config
    """
    r = ref()
    N, F = x.shape
    S = applicable.shape[1]
    values = P.unpack_onehot(x, S)
    names = np.array([f"f{i}" for i in range(F)])
    state_names = [[f"s{j}" for j in range(int(applicable[i].sum()))] for i in range(F)]
    features = r.load_data.Features(values=values, names=names, states=applicable.copy(),
                                    state_names=state_names, na_number=int((x == P.NA).sum()))
    objects = r.load_data.Objects(id=[f"o{i}" for i in range(N)], locations=np.asarray(locations, dtype=float),
                                  names=[f"o{i}" for i in range(N)])
    confs = OrderedDict()
    for name, grp in confounders.items():
        gn = (group_names or {}).get(name) or ([f"g{i}" for i in range(grp.shape[0])] if grp.shape[0] > 1 or name != "universal" else ["<ALL>"])
        confs[name] = r.load_data.Confounder(name=name, group_assignment=grp.astype(bool), group_names=list(gn))
    data = r.load_data.Data(objects=objects, features=features, confounders=confs, projection=None)
    # with the CRS stub the network uses Euclidean distances and has no lat_lon; the
    # cost_based geo prior dereferences it (sbayes/model/prior.py:812)
    data.network.lat_lon = data.network.locations
    cfg = r.config.ModelConfig(clusters=n_clusters, confounders=list(confounders.keys()), prior=prior_cfg)
    model = r.model.Model(data, cfg)
    return data, model


def problem_from_reference(data, model) -> P.Problem:
    """Packed `Problem` equivalent to a reference (Data, Model) pair."""
    r = ref()
    prior = model.prior
    x = P.pack_onehot(data.features.values)
    conf_gid, conf_ng, alpha_conf, names = [], [], [], []
    for name, conf in data.confounders.items():
        conf_gid.append(P.gid_from_bool(conf.group_assignment))
        conf_ng.append(conf.n_groups)
        cp = prior.prior_confounding_effects[name]
        if cp.any_dynamic_priors:
            raise NotImplementedError("hierarchical 'universal' prior type is rejected by the reference config")
        alpha_conf.append(np.array(cp._concentration_array, dtype=np.float64))
        names.append(name)
    sp = prior.size_prior
    size_type = {sp.PriorType.UNIFORM_AREA: P.SIZE_UNIFORM_AREA, sp.PriorType.UNIFORM_SIZE: P.SIZE_UNIFORM_SIZE,
                 sp.PriorType.QUADRATIC_SIZE: P.SIZE_QUADRATIC}[sp.prior_type]
    gp = prior.geo_prior
    geo = None
    if gp.prior_type is gp.PriorTypes.COST_BASED:
        GC = r.config.GeoPriorConfig
        if gp.config.skeleton is not GC.Skeleton.MST:
            raise NotImplementedError("only the MST skeleton is supported")
        geo = P.GeoPriorSpec(
            cost=np.array(gp.cost_matrix, dtype=np.float64),
            aggregation={GC.AggregationStrategies.MEAN: P.AGG_MEAN, GC.AggregationStrategies.SUM: P.AGG_SUM,
                         GC.AggregationStrategies.MAX: P.AGG_MAX}[gp.aggregation_policy],
            prob_fn={GC.ProbabilityFunction.EXPONENTIAL: P.PF_EXPONENTIAL,
                     GC.ProbabilityFunction.SIGMOID: P.PF_SIGMOID}[gp.probability_function],
            rate=float(gp.scale), inflection=float(gp.inflection_point or 0.0))
    elif gp.prior_type is not gp.PriorTypes.UNIFORM:
        raise NotImplementedError(f"geo prior {gp.prior_type}")
    wp = prior.prior_weights
    adj = data.network.adj_mat.tocsr()
    return P.Problem(
        x=x, applicable=np.array(data.features.states, dtype=bool),
        conf_gid=conf_gid, conf_ngroups=conf_ng,
        alpha_cluster=np.array(prior.prior_cluster_effect.concentration_array, dtype=np.float64),
        alpha_conf=alpha_conf,
        alpha_w=np.array(wp.concentration_array, dtype=np.float64),
        n_clusters=model.n_clusters, min_size=int(model.min_size), max_size=int(model.max_size),
        size_prior=size_type, weights_prior_uniform=(wp.prior_type is wp.PriorType.UNIFORM),
        geo=geo, adj_indptr=np.array(adj.indptr, dtype=np.int32), adj_indices=np.array(adj.indices, dtype=np.int32),
        conf_names=names)


def state_from_sample(prob: P.Problem, sample) -> P.ChainState:
    """Reference `Sample` -> packed `ChainState` (counts stacked [G_tot,F,S] int32)."""
    src = P.pack_onehot(sample.source.value)
    parts = [sample.feature_counts["clusters"].value]
    parts += [sample.feature_counts[n].value for n in prob.conf_names]
    counts = np.concatenate(parts, axis=0)
    assert np.all(counts == np.round(counts))
    return P.ChainState(cluster_of=P.cluster_of_from_bool(sample.clusters.value),
                        W=np.array(sample.weights.value, dtype=np.float32),
                        src=src, counts=counts.astype(np.int32))


def sample_from_state(prob: P.Problem, model, data, st: P.ChainState, i_step: int = 0):
    """Packed `ChainState` -> reference `Sample` (fresh caches)."""
    r = ref()
    offs = prob.group_offsets()
    counts = st.counts
    if counts is None:
        from .model import rebuild_counts
        counts = rebuild_counts(prob, st)
    fc = {"clusters": counts[offs[0]:offs[1]].astype(np.float32)}
    for i, n in enumerate(prob.conf_names):
        fc[n] = counts[offs[i + 1]:offs[i + 2]].astype(np.float32)
    s = r.state.Sample.from_numpy_arrays(
        clusters=st.clusters_bool(prob.K), weights=st.W.copy(), confounders=data.confounders,
        source=P.unpack_onehot(st.src, prob.C), feature_counts=fc, model_shapes=model.shapes)
    s.i_step = i_step
    s.everything_changed()
    return s
