"""Device version of the EM-style cluster initializer.

`generate_clusters_em(initializer, ctx)` stands in for `SbayesInitializer.generate_clusters_em`
(sbayes/sampling/initializers.py:93-169): the random start, the draw of the total cluster size and the final
discretisation are the reference's own code (host, O(G N)); the `n_em_steps` EM iterations -- two contractions
of the soft assignment with the one-hot feature tensor per step, the part that dominates start-up at large N --
run on the device (`sb_init_em`, csrc/sb_init.cuh).  `install_initializer()` rebinds the method on the
reference's class, so `SbayesInitializer.generate_sample` (initializers.py:204-237) uses it unchanged.
"""
from __future__ import annotations

import numpy as np

from . import refbridge
from .backend import DeviceContext


def initial_soft_assignment(prob, rng=None) -> np.ndarray:
    """z0 = normalize(random * groups_available, axis=0) as float32 (initializers.py:108-119); rows: clusters,
    then every confounder's groups."""
    rnd = np.random.random if rng is None else rng.random
    avail = groups_available(prob)
    z = rnd(avail.shape) * avail
    return (z / z.sum(axis=0, keepdims=True)).astype(np.float32)


def groups_available(prob) -> np.ndarray:
    rows = [np.ones((prob.K, prob.N), bool)]
    for gid, ng in zip(prob.conf_gid, prob.conf_ngroups):
        rows.append(np.arange(ng)[:, None] == gid[None, :])
    return np.concatenate(rows, axis=0)


def generate_clusters_em(initializer, ctx: DeviceContext = None, _lib=None):
    """Drop-in for `SbayesInitializer.generate_clusters_em`: returns bool clusters [K, N]."""
    self = initializer
    model, data = self.model, self.data
    own = ctx is None
    if own:
        prob = refbridge.problem_from_reference(data, model)
        ctx = DeviceContext(prob, 1, _lib=_lib)
    prob = ctx.prob
    try:
        n_objects, n_clusters = prob.N, prob.K
        total_size = self.sample_n_objects_in_all_clusters(
            mid=n_clusters * self.initial_size, lower_bound=n_clusters * model.min_size,
            upper_bound=min(n_objects, n_clusters * model.max_size))
        z0 = initial_soft_assignment(prob)                       # np.random.random, like the reference
        z = ctx.init_em(z0, int(self.n_em_steps))
        return self.discretize_fuzzy_cluster_2(z, total_size=total_size)
    finally:
        if own:
            ctx.close()


def install_initializer(_lib=None):
    """Rebind `SbayesInitializer.generate_clusters_em` to the device version (uniform geo prior; with a
    cost-based geo prior the reference's own method keeps running: its geo term is not on the device)."""
    import sbayes.model  # noqa: F401
    import sbayes.sampling.initializers as ini
    orig = ini.SbayesInitializer.generate_clusters_em

    def method(self):
        gp = self.model.prior.geo_prior
        if gp.prior_type is gp.PriorTypes.COST_BASED:
            return orig(self)
        return generate_clusters_em(self, _lib=_lib)
    ini.SbayesInitializer.generate_clusters_em = method
    return orig
