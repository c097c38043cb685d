"""Metropolis-Hastings step, chain run loop and MC3 swap: numpy restatement.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows sbayes/sampling/mcmc_chain.py:128-255
and sbayes/mcmc_setup.py:263-269,365-418.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .problem import Problem, ChainState, ChainConfig, N_OPS
from . import model as M
from . import operators as O
from .draws import SITE_MH, PhiloxDraws


@dataclass
class StepRecord:
    op: int
    log_q: float
    log_q_back: float
    ll_new: float          # un-tempered log-likelihood of the candidate
    prior_new: float       # un-tempered log-prior of the candidate
    accept: bool
    changed: bool
    mh_ratio: float = float("nan")
    log_u: float = float("nan")
    info: dict = field(default_factory=dict)


class OracleChain:
    """One tempered chain (MCMCChain, sbayes/sampling/mcmc_chain.py:23-275)."""

    def __init__(self, prob: Problem, cfg: ChainConfig, state: ChainState, draws=None, i_step: int = 0):
        self.prob = prob
        self.cfg = cfg
        self.state = state
        if self.state.counts is None:
            self.state.counts = M.rebuild_counts(prob, self.state)
        self.draws = draws if draws is not None else PhiloxDraws(cfg.seed)
        self.i_step = i_step
        self.op_weights = cfg.op_weight_vector(prob.K)
        self.accepts = np.zeros(N_OPS, dtype=np.int64)
        self.rejects = np.zeros(N_OPS, dtype=np.int64)
        self.ll = self.likelihood(self.state)          # un-tempered
        self.prior = M.log_prior(prob, self.state)     # un-tempered

    def likelihood(self, st: ChainState) -> float:
        """MCMCChain.likelihood (mcmc_chain.py:90-109), un-tempered."""
        if self.cfg.sample_from_prior:
            return 0.0
        return M.log_likelihood(self.prob, st)

    def step(self, tape: dict | None = None) -> StepRecord:
        """MCMCChain.step (mcmc_chain.py:186-238)."""
        cfg = self.cfg
        self.i_step += 1
        if tape is not None:
            self.draws.begin_step(self.i_step, tape)
        else:
            self.draws.begin_step(self.i_step)
        op = self.draws.operator(self.op_weights)
        prop = O.propose(self.prob, self.state, cfg, self.draws, op)
        cand = prop.state
        if prop.changed:
            ll_new = self.likelihood(cand)
            prior_new = M.log_prior(self.prob, cand)
        else:
            ll_new, prior_new = self.ll, self.prior
        rec = StepRecord(op, prop.log_q, prop.log_q_back, ll_new, prior_new, False, prop.changed, info=prop.info)
        T, Tp = cfg.temperature, cfg.prior_temperature
        if prop.log_q_back == O.NEG_INF:
            accept = False
        elif prop.log_q == O.NEG_INF:
            accept = True
        else:
            ratio = (ll_new / T + prior_new / Tp) - (self.ll / T + self.prior / Tp) - (prop.log_q - prop.log_q_back)
            u = self.draws.uniform(SITE_MH)
            log_u = math.log(u) if u > 0 else O.NEG_INF
            accept = log_u < ratio
            rec.mh_ratio, rec.log_u = float(ratio), log_u
        rec.accept = bool(accept)
        if accept:
            self.state, self.ll, self.prior = cand, ll_new, prior_new
            self.accepts[op] += 1
            if not (np.isfinite(self.ll) and np.isfinite(self.prior)):
                raise ValueError(f"non-finite posterior accepted after operator {op}")   # mcmc_chain.py:231-236
        else:
            self.rejects[op] += 1
        return rec

    def run(self, n_steps: int) -> list:
        return [self.step() for _ in range(n_steps)]


def temperature_ladder(n_chains: int, dT: float, dTp: float, exponential: bool):
    """sbayes/mcmc_setup.py:263-269."""
    idx = np.arange(n_chains)
    if exponential:
        return (1.0 + dT) ** idx, (1.0 + dTp) ** idx
    return 1.0 + dT * idx, 1.0 + dTp * idx


def swap_ratio(ll_i, prior_i, ll_j, prior_j, T_i, T_j, Tp_i, Tp_j) -> float:
    """MH ratio of an MC3 swap between chains i and j (sbayes/mcmc_setup.py:389-395)."""
    return -((prior_i - prior_j) * (1.0 / Tp_i - 1.0 / Tp_j) + (ll_i - ll_j) * (1.0 / T_i - 1.0 / T_j))


def swap_chains(lls, priors, temps, prior_temps, slot_of_chain, draws, attempts: int,
                only_neighbours: bool = False):
    """swap_chains (sbayes/mcmc_setup.py:365-418) on scalars.  `slot_of_chain[c]` says which
    state currently sits at temperature slot c; states are swapped by swapping those
    entries.  Returns the list of accepted (from, to) pairs."""
    n = len(slot_of_chain)
    if only_neighbours:
        pairs = [(i, i + 1) for i in range(n - 1)]
    else:
        pairs = [(i, j) for i in range(n - 1) for j in range(i + 1, n)]
    attempts = min(attempts, len(pairs))
    accepted = []
    for a, (i, j) in enumerate(draws.swap_pairs(pairs, attempts)):
        si, sj = slot_of_chain[i], slot_of_chain[j]
        r = swap_ratio(lls[si], priors[si], lls[sj], priors[sj], temps[i], temps[j], prior_temps[i], prior_temps[j])
        u = draws.swap_uniform(a)
        if (math.log(u) if u > 0 else O.NEG_INF) < r:
            slot_of_chain[i], slot_of_chain[j] = sj, si
            accepted.append((i, j))
    return accepted
