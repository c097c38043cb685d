"""Packed problem / chain-state containers: the host-side image of what the CUDA path keeps
in HBM ("Data layout").  Pure numpy, no compute.

* `x`   uint8 [N, F]  state index of object o in feature f, 255 = missing (NA).  The
  reference stores the same information as a one-hot bool tensor [N, F, S] with an
  all-zero row for NA (sbayes/load_data.py:85-105).
* `src` uint8 [N, F]  mixture component the observation is assigned to, 255 iff NA.
  Reference: one-hot bool [N, F, C] (sbayes/sampling/state.py:492-552).
* `cluster_of` int32 [N]  cluster index or -1.  Reference: bool [K, N]; clusters are
  disjoint by construction (sbayes/sampling/operators.py:724-725,1097-1107).
* confounder membership as one int32 group id per object (-1 = in no group).  Reference:
  bool [G, N], a partition of the non-missing objects (sbayes/load_data.py:160-184).
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Optional

import numpy as np

NA = 255

# operator ids, in the order of the reference schedule (sbayes/sampling/operators.py:50-205)
OP_NAMES = [
    "cluster_naive_n1",
    "cluster_naive_n1_geo",
    "cluster_naive_n2_geo",
    "cluster_gibbsish",
    "cluster_gibbsish_geo",
    "gibbsish_sample_cluster_wide_geo",
    "cluster_jump_gibbsish",
    "gibbs_sample_sources",
    "gibbs_sample_sources_groups",
    "gibbs_sample_weights",
]
N_OPS = len(OP_NAMES)

SIZE_UNIFORM_AREA, SIZE_UNIFORM_SIZE, SIZE_QUADRATIC = 0, 1, 2
GEO_UNIFORM, GEO_COST_BASED = 0, 1
AGG_MEAN, AGG_SUM, AGG_MAX = 0, 1, 2
PF_EXPONENTIAL, PF_SIGMOID = 0, 1


@dataclass
class GeoPriorSpec:
    """cost_based geo prior (sbayes/model/prior.py:704-853), MST skeleton only."""
    cost: np.ndarray                 # f64 [N, N]
    aggregation: int = AGG_MEAN
    prob_fn: int = PF_EXPONENTIAL
    rate: float = 1.0
    inflection: float = 0.0


@dataclass
class Problem:
    x: np.ndarray                    # u8 [N, F]
    applicable: np.ndarray           # bool [F, S]
    conf_gid: list                   # per confounder: int32 [N]
    conf_ngroups: list               # per confounder: int
    alpha_cluster: np.ndarray        # f64 [F, S]   cluster-effect Dirichlet concentration
    alpha_conf: list                 # per confounder: f64 [G, F, S]
    alpha_w: np.ndarray              # f64 [F, C]   weights-prior concentration
    n_clusters: int
    min_size: int
    max_size: int
    size_prior: int = SIZE_UNIFORM_AREA
    weights_prior_uniform: bool = True
    geo: Optional[GeoPriorSpec] = None
    adj_indptr: Optional[np.ndarray] = None   # CSR adjacency (Delaunay graph in the reference)
    adj_indices: Optional[np.ndarray] = None
    conf_names: list = field(default_factory=list)

    @property
    def N(self): return self.x.shape[0]
    @property
    def F(self): return self.x.shape[1]
    @property
    def S(self): return self.applicable.shape[1]
    @property
    def C(self): return 1 + len(self.conf_gid)
    @property
    def K(self): return self.n_clusters
    @property
    def G_tot(self): return self.n_clusters + int(sum(self.conf_ngroups))

    @property
    def unif(self) -> np.ndarray:
        """uniform_concentration_array: 1 on applicable states (sbayes/model/prior.py:184-186)."""
        return self.applicable.astype(np.float64)

    @property
    def na(self) -> np.ndarray:
        return self.x == NA

    def group_offsets(self):
        """Row offsets of every component's groups inside the stacked [G_tot, F, S] tensors."""
        offs = [0, self.n_clusters]
        for g in self.conf_ngroups:
            offs.append(offs[-1] + int(g))
        return offs

    def alpha_stacked(self) -> np.ndarray:
        """alpha [G_tot, F, S]: cluster rows repeat the shared cluster-effect prior."""
        parts = [np.broadcast_to(self.alpha_cluster, (self.n_clusters, self.F, self.S))]
        parts += list(self.alpha_conf)
        return np.ascontiguousarray(np.concatenate(parts, axis=0), dtype=np.float64)

    def conf_pattern(self) -> np.ndarray:
        """bitmask over confounders c>=1 of `has[o, c]` (bit c-1), int32 [N]."""
        pat = np.zeros(self.N, dtype=np.int32)
        for i, gid in enumerate(self.conf_gid):
            pat |= (gid >= 0).astype(np.int32) << i
        return pat


@dataclass
class ChainState:
    cluster_of: np.ndarray           # int32 [N]
    W: np.ndarray                    # f32 [F, C]
    src: np.ndarray                  # u8 [N, F]
    counts: Optional[np.ndarray] = None   # int32 [G_tot, F, S]

    def copy(self) -> "ChainState":
        return ChainState(self.cluster_of.copy(), self.W.copy(), self.src.copy(),
                          None if self.counts is None else self.counts.copy())

    def clusters_bool(self, K: int) -> np.ndarray:
        z = np.zeros((K, self.cluster_of.shape[0]), dtype=bool)
        m = self.cluster_of >= 0
        z[self.cluster_of[m], np.flatnonzero(m)] = True
        return z

    def sizes(self, K: int) -> np.ndarray:
        return np.bincount(self.cluster_of[self.cluster_of >= 0], minlength=K).astype(np.int64)


def cluster_of_from_bool(z: np.ndarray) -> np.ndarray:
    if np.any(z.sum(axis=0) > 1):
        raise ValueError("clusters overlap; the packed layout needs disjoint clusters")
    out = np.full(z.shape[1], -1, dtype=np.int32)
    k, o = np.nonzero(z)
    out[o] = k
    return out


def pack_onehot(values: np.ndarray) -> np.ndarray:
    """bool [.., S] one-hot (all-zero = NA) -> uint8 index with 255 = NA."""
    idx = values.argmax(axis=-1).astype(np.uint8)
    idx[values.sum(axis=-1) == 0] = NA
    return idx


def unpack_onehot(idx: np.ndarray, n: int) -> np.ndarray:
    out = np.zeros(idx.shape + (n,), dtype=bool)
    valid = idx != NA
    np.put_along_axis(out, np.where(valid, idx, 0)[..., None].astype(np.int64), valid[..., None], axis=-1)
    return out


def gid_from_bool(groups: np.ndarray) -> np.ndarray:
    if np.any(groups.sum(axis=0) > 1):
        raise ValueError("confounder groups overlap")
    out = np.full(groups.shape[1], -1, dtype=np.int32)
    g, o = np.nonzero(groups)
    out[o] = g
    return out


def bool_from_gid(gid: np.ndarray, n_groups: int) -> np.ndarray:
    out = np.zeros((n_groups, gid.shape[0]), dtype=bool)
    m = gid >= 0
    out[gid[m], np.flatnonzero(m)] = True
    return out


@dataclass
class ChainConfig:
    """Per-chain run parameters (sbayes/sampling/mcmc_chain.py:30-72, operators.py:28-209)."""
    temperature: float = 1.0
    prior_temperature: float = 1.0
    sample_from_prior: bool = False
    op_clusters: float = 70.0       # OperatorsConfig defaults (sbayes/config/config.py:311-318)
    op_weights: float = 10.0
    op_source: float = 20.0
    p_grow: float = 0.5
    seed: int = 0

    def op_weight_vector(self, n_clusters: int) -> np.ndarray:
        """Normalised operator weights (sbayes/sampling/operators.py:50-215)."""
        c, s, w = self.op_clusters, self.op_source, self.op_weights
        v = np.array([
            0.025 * c, 0.025 * c, 0.025 * c, 0.025 * c, 0.6 * c, 0.05 * c,
            0.25 * c if n_clusters > 1 else 0.0,
            0.4 * s, 0.6 * s, w,
        ], dtype=np.float64)
        return v / v.sum()
