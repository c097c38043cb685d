"""TEST INFRASTRUCTURE (oracle): effect draws at log time.

Restates `conditional_effect_sample` (sbayes/sampling/conditionals.py:125-149) as the reference's parameter
logger calls it (sbayes/sampling/loggers.py:160-224): for every group g (clusters first, then the groups of every
confounder) and feature f

    p[g, f, applicable states] ~ Dirichlet(counts[g, f, :] + prior_counts[g, f, :])

with the counts of `conditional_effect_concentration` (conditionals.py:68-86: observations of the group's objects
whose source is the group's component).  The reference draws from numpy's global generator; device and oracle draw
Gamma variates by Marsaglia-Tsang (2000) from the Philox streams (chain seed, step, SITE_EFFECT, attempt sub-stream,
(g*F + f)*S + s), shapes below one boosted by U^(1/shape), so both produce the same numbers.  Pinned against the
reference by the Dirichlet moments of its own function (tests/test_effect_sample.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module."""
from __future__ import annotations

import math

import numpy as np

from . import philox

SITE_EFFECT = 14


def concentration(prob, counts: np.ndarray) -> np.ndarray:
    """counts + prior_counts, float64 [G_tot, F, S] (conditionals.py:68-86 with the stacked prior rows)"""
    return counts.astype(np.float64) + prob.alpha_stacked()


def _gamma(seed: int, step: int, shape: float, idx: int) -> float:
    boost = shape < 1.0
    a = shape + 1.0 if boost else shape
    d = a - 1.0 / 3.0
    c = 1.0 / math.sqrt(9.0 * d)
    t = 0
    while True:
        p = philox.words(seed, step, SITE_EFFECT, 2 * t, idx)
        q = philox.words(seed, step, SITE_EFFECT, 2 * t + 1, idx)
        u1, u2 = float(philox.u01_double(p[0], p[1])), float(philox.u01_double(p[2], p[3]))
        u3, u4 = float(philox.u01_double(q[0], q[1])), float(philox.u01_double(q[2], q[3]))
        t += 1
        xn = math.sqrt(-2.0 * math.log(1.0 - u1)) * math.cos(2.0 * 3.141592653589793 * u2)
        v = 1.0 + c * xn
        if v <= 0.0:
            continue
        v = v * v * v
        if math.log(1.0 - u3) < 0.5 * xn * xn + d - d * v + d * math.log(v):
            g = d * v
            if boost:
                g *= math.exp(math.log(1.0 - u4) / shape)
            return g


def effect_sample(prob, counts: np.ndarray, seed: int, step: int) -> np.ndarray:
    """float64 [G_tot, F, S]; states with zero prior concentration are not applicable and stay zero"""
    alpha = prob.alpha_stacked()
    G, F, S = alpha.shape
    out = np.zeros((G, F, S))
    for g in range(G):
        for f in range(F):
            app = alpha[g, f] > 0.0
            tot = 0.0
            for s in range(S):
                if app[s]:
                    out[g, f, s] = _gamma(seed, step, float(alpha[g, f, s]) + float(counts[g, f, s]), (g * F + f) * S + s)
                    tot += out[g, f, s]
            if tot > 0.0:
                out[g, f] /= tot
            else:
                out[g, f] = app / max(int(app.sum()), 1)
    return out
