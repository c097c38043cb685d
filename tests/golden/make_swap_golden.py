#!/usr/bin/env python
"""Generate `mc3_swap.npz` from the LIVE reference: the MC3 swap of
`MCMCSetup.swap_chains` (sbayes/mcmc_setup.py:365-418), UNMODIFIED, called on reference
`Sample` objects, with recording proxies at its two RNG call sites (SURVEY.md Appendix D:
`RNG.choice(possible_swaps, size=attempts, replace=False)` :386 and `random.random()` :403).

Run in the build container only (needs /root/reference):

    python tests/golden/make_swap_golden.py

Stored per round: the un-tempered (log-likelihood, log-prior) the reference's own model
returns for the sample sitting at every temperature slot before the swap, the ladder, the
drawn pairs and uniforms, the swaps the reference accepted and the resulting order of the
samples.  The script replays every round through `oracle.chain.swap_chains` and fails on
any difference, so the committed fixture is reference output on which oracle == reference.
"""
from __future__ import annotations

import logging
import os
import random as _pyrandom
import sys
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import refharness as H, model as M, problem as P, chain as CH, fixtures as FX  # noqa: E402
from sbayes_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


class _RNGProxy:
    def __init__(self, seed, log):
        self.inner, self.log = np.random.default_rng(seed), log

    def choice(self, a, size=None, replace=True, p=None):
        out = self.inner.choice(a, size=size, replace=replace, p=p)
        self.log["pairs"] = np.array(out, dtype=np.int32).reshape(-1, 2)
        return out


class _RandomProxy:
    def __init__(self, seed, log):
        self.inner, self.log = _pyrandom.Random(seed), log

    def random(self):
        u = self.inner.random()
        self.log.setdefault("u", []).append(u)
        return u


class _TapeSwapDraws:
    """Draw source of oracle.chain.swap_chains replaying one recorded round."""

    def __init__(self, pairs, u):
        self.pairs, self.u = pairs, u

    def swap_pairs(self, pairs, attempts):
        assert attempts == len(self.pairs)
        return [tuple(int(v) for v in p) for p in self.pairs]

    def swap_uniform(self, a):
        return float(self.u[a])


class _Setup:
    """The attributes `swap_chains` touches on its MCMCSetup (mcmc_setup.py:365-418,420-425)."""

    def __init__(self, model, n):
        self.model = model
        self.swap_attempts = 0
        self.swap_accepts = 0
        self.swap_matrix = np.zeros((n, n), dtype=int)
        self.logger = logging.getLogger("swap_golden")
        self.t_start = 0.0

    def print_screen_log(self, i_step, likelihood):
        pass


def main():
    r = H.ref()
    n = 5
    prob0, truth = synth.make_problem(N=60, F=10, S=3, K=2, n_fam=3, seed=11, max_size=25)
    confs = OrderedDict(universal=np.ones((1, prob0.N), bool),
                        family=np.stack([prob0.conf_gid[1] == g for g in range(prob0.conf_ngroups[1])]))
    prior_cfg = dict(objects_per_cluster=dict(type="uniform_size", min=prob0.min_size, max=prob0.max_size),
                     geo=dict(type="uniform"), weights=dict(type="symmetric_dirichlet", prior_concentration=2.0),
                     cluster_effect=dict(type="uniform"),
                     confounding_effects=dict(universal={"<ALL>": dict(type="uniform")}, family={"<DEFAULT>": dict(type="jeffreys")}))
    data, model = H.build_reference_model(prob0.x, prob0.applicable, truth["locations"], confs, prob0.K, prior_cfg)
    prob = H.problem_from_reference(data, model)
    # diverse states: different seeds and cluster sizes, so the posteriors differ by O(10)
    states = []
    for c in range(n):
        st = synth.initial_state(prob, seed=40 + c, objects_per_cluster=3 + 3 * c)
        st.counts = M.rebuild_counts(prob, st)
        states.append(st)
    samples = [H.sample_from_state(prob, model, data, st) for st in states]
    for c, s in enumerate(samples):
        s.chain = c
    ladders = [(0.05, 0.0, False, False), (0.3, 0.1, False, False), (0.5, 0.2, True, False), (0.4, 0.15, False, True),
               (0.02, 0.0, False, False), (1.0, 0.5, False, False)]
    setup = _Setup(model, n)
    rounds = []
    order = list(range(n))          # order[slot] = index of the state sitting at that temperature slot
    for rnd, (dT, dTp, expo, neigh) in enumerate(ladders * 2):
        T, Tp = CH.temperature_ladder(n, dT, dTp, expo)
        n_pairs = n - 1 if neigh else n * (n - 1) // 2
        attempts = min(100, n_pairs) if rnd % 3 else min(4, n_pairs)
        ll = np.array([model.likelihood(samples[s]) for s in range(n)], dtype=np.float64)
        pr = np.array([model.prior(samples[s]) for s in range(n)], dtype=np.float64)
        log = {}
        r.mcmc_setup.RNG = _RNGProxy(1000 + rnd, log)
        r.mcmc_setup.random = _RandomProxy(2000 + rnd, log)
        before = list(samples)
        acc0 = setup.swap_accepts
        r.mcmc_setup.MCMCSetup.swap_chains(setup, samples, list(T), list(Tp), attempts=attempts, only_swap_neighbours=neigh)
        perm = [before.index(s) for s in samples]                 # perm[slot] = slot the sample came from
        n_acc = setup.swap_accepts - acc0
        # replay through the oracle on the same scalars and draws
        slot = list(range(n))
        acc = CH.swap_chains(ll, pr, T, Tp, slot, _TapeSwapDraws(log["pairs"], log["u"]), attempts, neigh)
        if slot != perm or len(acc) != n_acc:
            raise SystemExit(f"round {rnd}: oracle {slot} {acc} != reference {perm} ({n_acc} accepted)")
        order = [order[p] for p in perm]
        rounds.append(dict(ll=ll, prior=pr, T=np.array(T), Tp=np.array(Tp), pairs=log["pairs"], u=np.array(log["u"]),
                           perm=np.array(perm, np.int32), n_acc=n_acc, attempts=attempts, neigh=int(neigh),
                           accepted=np.array(acc, np.int32).reshape(-1, 2)))
        print(f"round {rnd}: dT={dT} dTp={dTp} exp={expo} neigh={neigh} attempts={attempts} accepted {acc}")
    tot = sum(r_["n_acc"] for r_ in rounds)
    tot_att = sum(r_["attempts"] for r_ in rounds)
    assert 0 < tot < tot_att, "fixture must contain accepted and rejected swaps"
    out = dict(n_rounds=np.int64(len(rounds)), n_chains=np.int64(n))
    for i, r_ in enumerate(rounds):
        for k, v in r_.items():
            out[f"r{i}_{k}"] = np.asarray(v)
    out.update(FX.problem_to_dict(prob))
    for c, st in enumerate(states):
        out[f"s{c}_cluster_of"], out[f"s{c}_W"], out[f"s{c}_src"] = st.cluster_of, st.W, st.src
    np.savez_compressed(os.path.join(HERE, "mc3_swap.npz"), **out)
    print(f"mc3_swap.npz: {len(rounds)} rounds, {tot} of {tot_att} swap attempts accepted")


if __name__ == "__main__":
    main()
