"""Draw sources for the oracle: Philox streams (identical to the CUDA path) or a tape of
draws recorded from the live reference.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Every random decision of one MCMC step is addressed by a *site* (which decision), a
*sub-stream* (retry number where a decision loops) and an *index* (which element).  The
site list follows the RNG call sites of the reference:

  SITE_OP        operator choice                     sbayes/sampling/mcmc_chain.py:183
  SITE_MH        Metropolis-Hastings uniform         sbayes/sampling/mcmc_chain.py:217
  SITE_CLUSTER   cluster index                       sbayes/sampling/operators.py:998,1507
  SITE_GROW      grow/shrink uniform                 sbayes/sampling/operators.py:1007
  SITE_OBJECT    object to add/remove/jump           sbayes/sampling/operators.py:1151,1211,1749
  SITE_SUBSET    object subset of the source Gibbs   sbayes/sampling/operators.py:456
  SITE_COMPONENT component + group of GROUPS selector sbayes/sampling/operators.py:472-473
  SITE_GIBBS     categorical uniforms [M, F]         sbayes/preprocessing.py:247
  SITE_WIDE      wide-flip uniforms [n_avail] (f32)  sbayes/sampling/operators.py:1515-1517
  SITE_JUMP      source/target cluster pair          sbayes/sampling/operators.py:1732
  SITE_WPAIR     the two components to rebalance     sbayes/sampling/operators.py:646
  SITE_BETA      Beta draws [F]                      sbayes/sampling/operators.py:661
  SITE_WACCEPT   per-feature accept uniforms (f32)   sbayes/sampling/operators.py:626
  SITE_SWAP      MC3 swap pairs / uniforms           sbayes/mcmc_setup.py:386,403
"""
from __future__ import annotations

import math

import numpy as np

from . import philox

(SITE_OP, SITE_MH, SITE_CLUSTER, SITE_GROW, SITE_OBJECT, SITE_SUBSET, SITE_COMPONENT, SITE_GIBBS,
 SITE_WIDE, SITE_JUMP, SITE_WPAIR, SITE_BETA, SITE_WACCEPT, SITE_SWAP) = range(14)
N_SITES = 14


def cdf_pick(p: np.ndarray, u: float) -> int:
    """Inverse-CDF draw used by numpy's `choice(p=...)`: first i with cumsum(p)[i]/total > u."""
    cdf = np.cumsum(np.asarray(p, dtype=np.float64))
    cdf = cdf / cdf[-1]
    i = int(np.searchsorted(cdf, u, side="right"))
    return min(i, len(cdf) - 1)


class PhiloxDraws:
    """Counter-based streams; bit-identical to `sb_rng.cuh`."""

    def __init__(self, seed: int):
        self.seed = int(seed)
        self.step = 0

    def begin_step(self, step: int):
        self.step = int(step)

    # -- primitive draws ---------------------------------------------------------------
    def uniform(self, site: int, idx=0, sub: int = 0):
        w = philox.words(self.seed, self.step, site, sub, idx)
        u = philox.u01_double(w[0], w[1])
        return float(u) if np.ndim(idx) == 0 else u

    def uniform_f32(self, site: int, idx=0, sub: int = 0):
        w = philox.words(self.seed, self.step, site, sub, idx)
        u = philox.u01_float(w[0])
        return float(u) if np.ndim(idx) == 0 else u

    def index(self, site: int, idx: int, n: int, sub: int = 0) -> int:
        return min(int(self.uniform(site, idx, sub) * n), n - 1)

    # -- composite draws ---------------------------------------------------------------
    def operator(self, weights: np.ndarray) -> int:
        return cdf_pick(weights, self.uniform(SITE_OP))

    def pair(self, site: int, n: int):
        """ordered pair of distinct indices in [0, n)."""
        i = self.index(site, 0, n)
        j = self.index(site, 1, n - 1)
        return i, j + (1 if j >= i else 0)

    def subset_of(self, population: np.ndarray, k: int) -> np.ndarray:
        """uniform random k-subset of `population` (Floyd's algorithm), sorted."""
        n = len(population)
        chosen = []
        for j in range(n - k, n):
            t = self.index(SITE_SUBSET, j, j + 1)
            chosen.append(j if t in chosen else t)
        return np.sort(np.asarray(population)[np.array(chosen, dtype=np.int64)])

    def gibbs_uniforms(self, M: int, F: int) -> np.ndarray:
        return self.uniform(SITE_GIBBS, np.arange(M * F, dtype=np.int64)).reshape(M, F)

    def wide_uniforms(self, n: int, attempt: int) -> np.ndarray:
        return self.uniform_f32(SITE_WIDE, np.arange(n, dtype=np.int64), sub=attempt)

    def accept_uniforms_f32(self, F: int) -> np.ndarray:
        return self.uniform_f32(SITE_WACCEPT, np.arange(F, dtype=np.int64))

    def _gamma(self, shape: float, f: int, which: int) -> float:
        """Marsaglia-Tsang (2000) for shape >= 1; attempt t uses sub-streams 4*t + 2*which + {0, 1}."""
        d = shape - 1.0 / 3.0
        c = 1.0 / math.sqrt(9.0 * d)
        t = 0
        while True:
            w = philox.words(self.seed, self.step, SITE_BETA, 4 * t + 2 * which, f)
            u1 = float(philox.u01_double(w[0], w[1]))
            u2 = float(philox.u01_double(w[2], w[3]))
            w2 = philox.words(self.seed, self.step, SITE_BETA, 4 * t + 2 * which + 1, f)
            u3 = float(philox.u01_double(w2[0], w2[1]))
            # Box-Muller; 1-u1 in (0,1] keeps the log finite
            xn = math.sqrt(-2.0 * math.log(1.0 - u1)) * math.cos(2.0 * math.pi * u2)
            v = 1.0 + c * xn
            t += 1
            if v <= 0.0:
                continue
            v = v * v * v
            if math.log(1.0 - u3) < 0.5 * xn * xn + d - d * v + d * math.log(v):
                return d * v

    def beta(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        """Beta(a[f], b[f]) for every feature, a,b >= 1, as X/(X+Y) of two Gammas."""
        out = np.empty(len(a))
        for f in range(len(a)):
            gx = self._gamma(float(a[f]), f, 0)
            gy = self._gamma(float(b[f]), f, 1)
            out[f] = gx / (gx + gy)
        return out

    def swap_pairs(self, pairs: list, attempts: int):
        """choose `attempts` of the candidate pairs without replacement (partial Fisher-Yates)."""
        idx = list(range(len(pairs)))
        out = []
        for a in range(attempts):
            j = a + self.index(SITE_SWAP, a, len(idx) - a)
            idx[a], idx[j] = idx[j], idx[a]
            out.append(pairs[idx[a]])
        return out

    def swap_uniform(self, a: int) -> float:
        return self.uniform(SITE_SWAP, a, sub=1)


class TapeDraws:
    """Replays the draws of one recorded reference step.

    `tape` maps site -> recorded payload (see oracle/refrecord.py)."""

    def __init__(self, tape: dict | None = None):
        self.tape = tape or {}
        self.step = 0

    def begin_step(self, step: int, tape: dict | None = None):
        self.step = step
        if tape is not None:
            self.tape = tape

    def _get(self, site):
        if site not in self.tape:
            raise KeyError(f"tape has no draw for site {site} at step {self.step}")
        return self.tape[site]

    def uniform(self, site, idx=0, sub=0):
        v = np.asarray(self._get(site), dtype=np.float64).reshape(-1)
        return float(v[idx]) if np.ndim(idx) == 0 else v[idx]

    uniform_f32 = uniform

    def index(self, site, idx, n, sub=0) -> int:
        return int(np.asarray(self._get(site)).reshape(-1)[idx])

    def operator(self, weights) -> int:
        return int(self._get(SITE_OP))

    def pair(self, site, n):
        v = np.asarray(self._get(site)).reshape(-1)
        return int(v[0]), int(v[1])

    def subset_of(self, population, k):
        return np.sort(np.asarray(self._get(SITE_SUBSET), dtype=np.int64))

    def gibbs_uniforms(self, M, F):
        return np.asarray(self._get(SITE_GIBBS), dtype=np.float64).reshape(M, F)

    def wide_uniforms(self, n, attempt):
        v = self._get(SITE_WIDE)          # list of attempts
        return np.asarray(v[min(attempt, len(v) - 1)], dtype=np.float64)

    def accept_uniforms_f32(self, F):
        return np.asarray(self._get(SITE_WACCEPT), dtype=np.float64)

    def beta(self, a, b):
        return np.asarray(self._get(SITE_BETA), dtype=np.float64)
