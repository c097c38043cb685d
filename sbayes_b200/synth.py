"""Synthetic sBayes problems and initial chain states (host-side, numpy only).

Follows the reference simulator (sbayes/simulation.py:132-257, sbayes/preprocessing.py:323-394)
as specified in SURVEY.md section 8(d): locations uniform in the unit square, spatially
compact true clusters, families ~ multinomial with 20 % isolates, one universal group,
weights ~ Dir(2,..,2) per feature, effects ~ Dir(0.3) for clusters and Dir(0.4) for
confounders, features drawn from the explicit mixture, 2.5 % missing values.
Used by `bench.py` and the tests; not part of the timed path.
"""
from __future__ import annotations

import numpy as np

from .packed import Problem, ChainState, GeoPriorSpec, NA, SIZE_UNIFORM_AREA, AGG_MEAN, PF_EXPONENTIAL


def delaunay_csr(locations: np.ndarray):
    """Symmetric CSR adjacency of the Delaunay triangulation (the reference's
    ComputeNetwork.adj_mat, sbayes/preprocessing.py:136-141, sbayes/util.py:146-175)."""
    from scipy.spatial import Delaunay
    n = locations.shape[0]
    if n < 4:
        ip = np.arange(0, n * (n - 1) + 1, max(n - 1, 1), dtype=np.int32)[: n + 1]
        ix = np.array([j for i in range(n) for j in range(n) if j != i], dtype=np.int32)
        return ip, ix
    tri = Delaunay(locations, qhull_options="QJ Pp")
    indptr, indices = tri.vertex_neighbor_vertices
    # sort neighbour lists for determinism
    out_ix = np.empty_like(indices, dtype=np.int32)
    for i in range(n):
        out_ix[indptr[i]:indptr[i + 1]] = np.sort(indices[indptr[i]:indptr[i + 1]])
    return indptr.astype(np.int32), out_ix


def make_problem(N: int, F: int, S: int, K: int, n_fam: int, seed: int, na_frac: float = 0.025,
                 geo_cost_based: bool = False, min_size: int = 3, max_size: int | None = None,
                 with_adjacency: bool = True, jagged_states: bool = False) -> tuple[Problem, dict]:
    """Returns (problem, truth) where truth holds the generating clusters / weights."""
    rng = np.random.default_rng(seed)
    loc = rng.random((N, 2))
    # compact true clusters: nearest neighbours of random centres, disjoint
    true_cluster = np.full(N, -1, dtype=np.int32)
    per = max(min_size, int(np.ceil(N / (4 * K))))
    for k in range(K):
        free = np.flatnonzero(true_cluster < 0)
        centre = loc[rng.choice(free)]
        d = np.linalg.norm(loc[free] - centre, axis=1)
        true_cluster[free[np.argsort(d)[:per]]] = k
    # families: 20 % isolates, the rest multinomial over n_fam groups (every family non-empty)
    fam = np.full(N, -1, dtype=np.int32)
    if n_fam > 0:
        non_iso = np.flatnonzero(rng.random(N) >= 0.2)
        fam[non_iso] = rng.integers(0, n_fam, size=non_iso.size)
        for g in range(n_fam):                      # guarantee every group has a member
            if not np.any(fam == g):
                fam[rng.integers(0, N)] = g
    universal = np.zeros(N, dtype=np.int32)
    conf_gid = [universal] + ([fam] if n_fam > 0 else [])
    conf_ng = [1] + ([n_fam] if n_fam > 0 else [])
    names = ["universal"] + (["family"] if n_fam > 0 else [])
    C = 1 + len(conf_gid)

    applicable = np.ones((F, S), dtype=bool)
    if jagged_states:
        ns = rng.integers(2, S + 1, size=F)
        ns[0] = S
        applicable = np.arange(S)[None, :] < ns[:, None]

    def dirichlet_rows(alpha, shape):
        p = rng.gamma(alpha, size=shape + (S,)) * applicable
        return p / p.sum(axis=-1, keepdims=True)

    W_true = rng.dirichlet(np.full(C, 2.0), size=F)
    p_cluster = dirichlet_rows(0.3, (K, F))
    p_conf = [dirichlet_rows(0.4, (g, F)) for g in conf_ng]

    has = np.stack([true_cluster >= 0] + [g >= 0 for g in conf_gid], axis=1)
    def draw_block(sl):
        """features of the objects in `sl` from the explicit mixture"""
        n = has[sl].shape[0]
        w = has[sl][:, None, :] * W_true[None]
        w = w / w.sum(axis=-1, keepdims=True)
        comp = (rng.random((n, F, 1)) < np.cumsum(w, axis=-1)).argmax(axis=-1)
        xb = np.zeros((n, F), dtype=np.uint8)
        for c in range(C):
            gid = (true_cluster if c == 0 else conf_gid[c - 1])[sl]
            table = p_cluster if c == 0 else p_conf[c - 1]
            rows = table[np.where(gid >= 0, gid, 0)]            # [n,F,S]
            draw = (rng.random((n, F, 1)) < np.cumsum(rows, axis=-1)).argmax(axis=-1)
            xb = np.where(comp == c, draw, xb).astype(np.uint8)
        xb[rng.random((n, F)) < na_frac] = NA
        return xb

    if N * F * S <= (1 << 24):
        x = draw_block(slice(0, N))
    else:   # large shapes (BASELINE configs 4 and 5): blocks of objects bound the [n,F,S] temporaries
        blk = max(1, (1 << 22) // (F * S))
        x = np.concatenate([draw_block(slice(a, min(N, a + blk))) for a in range(0, N, blk)], axis=0)

    geo = None
    if geo_cost_based:
        cost = np.linalg.norm(loc[:, None] - loc[None], axis=-1)
        geo = GeoPriorSpec(cost=cost, aggregation=AGG_MEAN, prob_fn=PF_EXPONENTIAL, rate=0.1)
    ip, ix = delaunay_csr(loc) if with_adjacency else (np.zeros(N + 1, np.int32), np.zeros(0, np.int32))
    prob = Problem(
        x=x, applicable=applicable, conf_gid=conf_gid, conf_ngroups=conf_ng,
        alpha_cluster=applicable.astype(np.float64),
        alpha_conf=[np.broadcast_to(applicable.astype(np.float64), (g, F, S)).copy() for g in conf_ng],
        alpha_w=np.ones((F, C)), n_clusters=K, min_size=min_size,
        max_size=int(max_size if max_size is not None else max(min_size + 1, N // K)),
        size_prior=SIZE_UNIFORM_AREA, weights_prior_uniform=True, geo=geo,
        adj_indptr=ip, adj_indices=ix, conf_names=names)
    return prob, {"clusters": true_cluster, "W": W_true, "locations": loc}


def initial_state(prob: Problem, seed: int, objects_per_cluster: int = 10) -> ChainState:
    """A valid starting state: disjoint random clusters of `objects_per_cluster` objects,
    W ~ Dir(1), sources drawn from the prior w~ (sample_source_from_prior,
    sbayes/sampling/conditionals.py:370-375).  Counts are left to the backend."""
    rng = np.random.default_rng(seed)
    N, F, C, K = prob.N, prob.F, prob.C, prob.K
    size = int(np.clip(objects_per_cluster, prob.min_size, min(prob.max_size, max(prob.min_size, N // max(K, 1)))))
    perm = rng.permutation(N)
    cluster_of = np.full(N, -1, dtype=np.int32)
    for k in range(K):
        cluster_of[perm[k * size:(k + 1) * size]] = k
    W = rng.dirichlet(np.ones(C), size=F).astype(np.float32)
    has = np.stack([cluster_of >= 0] + [g >= 0 for g in prob.conf_gid], axis=1)

    def draw_block(sl):
        w = has[sl][:, None, :] * W[None].astype(np.float64)
        w = w / w.sum(axis=-1, keepdims=True)
        return (rng.random((w.shape[0], F, 1)) < np.cumsum(w, axis=-1)).argmax(axis=-1).astype(np.uint8)

    if N * F * C <= (1 << 24):
        src = draw_block(slice(0, N))
    else:
        blk = max(1, (1 << 22) // (F * C))
        src = np.concatenate([draw_block(slice(a, min(N, a + blk))) for a in range(0, N, blk)], axis=0)
    src[prob.x == NA] = NA
    return ChainState(cluster_of=cluster_of, W=W, src=src, counts=None)
