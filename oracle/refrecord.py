"""Replay recorder: runs the *live* reference `MCMCChain.step` with recording proxies at
every RNG call site and emits, per step, the draw tape in the
format `oracle/draws.py:TapeDraws` consumes plus the reference's own outputs.

TEST INFRASTRUCTURE (see oracle/__init__.py); only usable in the build container.

The proxies do not alter the reference's algorithms: discrete draws with computed
probabilities are made by the same inverse-CDF rule numpy's `Generator.choice(p=...)` uses
(cdf = cumsum(p)/cumsum(p)[-1]; searchsorted(u, 'right')) so that the uniform itself can
be recorded; everything else is recorded as the outcome.
"""
from __future__ import annotations

import random as _pyrandom

import numpy as np

from . import refharness as H
from .draws import (SITE_OP, SITE_MH, SITE_CLUSTER, SITE_GROW, SITE_OBJECT, SITE_SUBSET, SITE_COMPONENT,
                    SITE_GIBBS, SITE_WIDE, SITE_JUMP, SITE_WPAIR, SITE_BETA, SITE_WACCEPT)
from .problem import OP_NAMES


class _Tape(dict):
    def put(self, site, value):
        self[site] = value

    def append(self, site, value):
        self.setdefault(site, []).append(value)


class _RNGProxy:
    """Stands in for the module-level `RNG` of sbayes.sampling.operators."""

    def __init__(self, rec, seed):
        self.rec = rec
        self.inner = np.random.default_rng(seed)

    def random(self, size=None, dtype=np.float64):
        v = self.inner.random(size, dtype=dtype)
        op = self.rec.current_op
        if op == "gibbsish_sample_cluster_wide_geo":
            self.rec.tape.append(SITE_WIDE, np.array(v, dtype=np.float64))
        elif op == "gibbs_sample_weights":
            self.rec.tape.put(SITE_WACCEPT, np.array(v, dtype=np.float64))
        else:
            raise RuntimeError(f"unexpected RNG.random in {op}")
        return v

    def choice(self, a, size=None, replace=True, p=None):
        tape = self.rec.tape
        if p is not None:
            p64 = np.asarray(p, dtype=np.float64)
            u = float(self.inner.random())
            cdf = np.cumsum(p64)
            cdf /= cdf[-1]
            idx = int(np.searchsorted(cdf, u, side="right"))
            tape.put(SITE_OBJECT, np.array([u]))
            pop = np.arange(a) if np.ndim(a) == 0 else np.asarray(a)
            val = pop[idx]
            return val if size is None else np.array([val])
        if isinstance(a, range):
            n = len(a)
            if size is None:
                k = int(self.inner.integers(0, n))
                tape.put(SITE_CLUSTER, np.array([k]))
                return a[k]
            assert size == 2 and not replace
            i = int(self.inner.integers(0, n))
            j = int(self.inner.integers(0, n - 1))
            j += (j >= i)
            tape.put(SITE_JUMP, np.array([i, j]))
            return np.array([a[i], a[j]])
        if isinstance(a, list) and len(a) > 0 and isinstance(a[0], str):
            k = int(self.inner.integers(0, len(a)))
            tape.setdefault(SITE_COMPONENT, [0, 0])[0] = k
            return a[k]
        arr = np.asarray(a)
        if arr.ndim == 2:
            k = int(self.inner.integers(0, arr.shape[0]))
            tape.setdefault(SITE_COMPONENT, [0, 0])[1] = k
            return arr[k]
        # subset without replacement from an int population or an index array
        assert size is not None and not replace
        out = self.inner.choice(a, size=size, replace=False)
        tape.put(SITE_SUBSET, np.sort(np.asarray(out, dtype=np.int64)))
        return out


class _PyRandomProxy:
    """Stands in for the `random` module inside sbayes.sampling.operators / mcmc_chain."""

    def __init__(self, rec, seed, mh: bool):
        self.rec, self.mh = rec, mh
        self.inner = _pyrandom.Random(seed)

    def random(self):
        u = self.inner.random()
        self.rec.tape.put(SITE_MH if self.mh else SITE_GROW, np.array([u]))
        return u

    def sample(self, population, k):
        out = self.inner.sample(population, k)
        self.rec.tape.put(SITE_WPAIR, np.array(out))
        return out


class _FrozenBeta:
    def __init__(self, rec, dist):
        self.rec, self.dist = rec, dist

    def rvs(self, *a, **kw):
        v = self.dist.rvs(*a, random_state=self.rec.beta_rng, **kw)
        self.rec.tape.put(SITE_BETA, np.array(v, dtype=np.float64))
        return v

    def logpdf(self, x):
        return self.dist.logpdf(x)


class _StatsProxy:
    def __init__(self, rec, stats):
        self.rec, self._stats = rec, stats

    def beta(self, a, b):
        return _FrozenBeta(self.rec, self._stats.beta(a, b))

    def __getattr__(self, name):
        return getattr(self._stats, name)


class ReferenceRecorder:
    """Drives one reference `MCMCChain` step by step and records the draws."""

    def __init__(self, data, model, op_cfg: dict, T: float = 1.0, Tp: float = 1.0, seed: int = 0,
                 sample_from_prior: bool = False):
        r = H.ref()
        self.r = r
        self.tape = _Tape()
        self.current_op = None
        self.beta_rng = np.random.default_rng(seed + 7)
        self._np_rng = np.random.default_rng(seed + 11)
        r.mcmc_chain.MCMCChain.CHECK_CACHING = True
        self.chain = r.mcmc_chain.MCMCChain(
            model=model, data=data, operators=r.config.OperatorsConfig(**op_cfg), sample_loggers=[],
            sample_from_prior=sample_from_prior, temperature=T, prior_temperature=Tp)
        # install proxies
        r.operators.RNG = _RNGProxy(self, seed + 1)
        r.operators.random = _PyRandomProxy(self, seed + 2, mh=False)
        r.mcmc_chain._random = _PyRandomProxy(self, seed + 3, mh=True)
        r.operators.stats = _StatsProxy(self, __import__("scipy.stats", fromlist=["stats"]))
        self._orig_np_random = np.random.random
        self._orig_np_choice = np.random.choice
        self.op_names = list(self.chain.callable_operators.keys())
        assert self.op_names == OP_NAMES, self.op_names

    def _np_random(self, size=None):
        v = self._np_rng.random(size)
        self.tape.put(SITE_GIBBS, np.array(v, dtype=np.float64).reshape(v.shape[0], -1) if np.ndim(v) > 1 else np.array(v))
        return v

    def _np_choice(self, a, size=None, replace=True, p=None):
        u = float(self._np_rng.random())
        cdf = np.cumsum(np.asarray(p, dtype=np.float64))
        cdf /= cdf[-1]
        idx = int(np.searchsorted(cdf, u, side="right"))
        self.current_op = a[idx]
        self.tape.put(SITE_OP, idx)
        return np.array([a[idx]])

    def start(self, sample):
        self.sample = sample
        self.chain._ll = self.chain.likelihood(sample)
        self.chain._prior = self.chain.prior(sample)

    def step(self):
        """One reference step.  Returns (tape, outputs dict)."""
        self.tape = _Tape()
        ch = self.chain
        before = self.sample
        acc0 = {k: (op.accepts, op.rejects) for k, op in ch.callable_operators.items()}
        np.random.random = self._np_random
        np.random.choice = self._np_choice
        try:
            # capture the proposal outputs by wrapping the chosen operator's function
            captured = {}
            ops = ch.callable_operators

            def wrap(op):
                orig = op.function

                def f(sample, **kw):
                    out = orig(sample, **kw)
                    captured["log_q"], captured["log_q_back"] = float(out[1]), float(out[2])
                    captured["candidate"] = out[0]
                    return out
                return orig, f
            saved = {}
            for k, op in ops.items():
                saved[k], op.function = wrap(op)
            try:
                after = ch.step(before)
            finally:
                for k, op in ops.items():
                    op.function = saved[k]
        finally:
            np.random.random = self._orig_np_random
            np.random.choice = self._orig_np_choice
        after.i_step = before.i_step + 1
        name = self.current_op
        a1, r1 = ops[name].accepts, ops[name].rejects
        accept = (a1 - acc0[name][0]) == 1
        cand = captured["candidate"]
        out = dict(op=OP_NAMES.index(name), log_q=captured["log_q"], log_q_back=captured["log_q_back"],
                   accept=bool(accept), ll_new=float(cand.last_lh), prior_new=float(cand.last_prior),
                   ll=float(ch._ll * ch.temperature), prior=float(ch._prior * ch.prior_temperature))
        self.sample = after
        return dict(self.tape), out
