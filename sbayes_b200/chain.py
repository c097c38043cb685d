"""Drop-in for the reference's chain driver (boundary B1 of SURVEY.md section 8b).

`MCMCChain` has the constructor and the `run / likelihood / prior / close_loggers /
reset_posterior_cache / callable_operators` surface of
`sbayes.sampling.mcmc_chain.MCMCChain` (sbayes/sampling/mcmc_chain.py:23-274), takes and
returns the reference's own `Sample` objects, and runs every MCMC step on the B200 through
the C ABI (`sb_run`, include/sbayes_b200.h).  Host code runs only at the points where the
reference touches the host as well: chain start (state upload) and every
`logging_interval` steps (state download, the loggers' `write_sample`).

`install()` substitutes this class for the reference's in `sbayes.sampling.mcmc_chain` and
`sbayes.mcmc_setup`, so `sbayes config.yaml` (MC3 included: one `MCMCChainProcess` per
chain, each with its own device context) runs unchanged with only the backend swapped.

There is no CPU fallback: without the CUDA library / a B200 the constructor raises.
"""
from __future__ import annotations

import logging
import os
import time as _time
import typing as typ

import numpy as np

from . import refbridge
from .backend import DeviceContext, SbError
from .packed import ChainConfig, OP_NAMES, N_OPS


class MCMCChain:
    """Tempered Metropolis-Hastings chain whose steps run on the device.

    Same arguments as the reference class (sbayes/sampling/mcmc_chain.py:30-42); the extra
    keyword-only ones select the device and the Philox seed (the reference is unseeded)."""

    CHECK_CACHING = __debug__

    def __init__(self, model, data, operators, sample_loggers: typ.List, sample_from_prior: bool = False,
                 show_screen_log: bool = False, logger: logging.Logger = None, screen_log_interval: int = 1000,
                 temperature: float = 1.0, prior_temperature: float = 1.0, *,
                 device: typ.Optional[int] = None, seed: typ.Optional[int] = None, team: int = 0, _lib=None):
        self.model, self.data = model, data
        self.temperature, self.prior_temperature = temperature, prior_temperature
        self.sample_from_prior = sample_from_prior
        self.callable_operators = self.get_operators(operators)
        self.sample_loggers = sample_loggers
        self._ll = -np.inf
        self._prior = -np.inf
        self.show_screen_log = show_screen_log
        self.screen_log_interval = screen_log_interval
        self.t_start = None
        self.screen_logger = logger
        for lg in self.sample_loggers:
            if type(lg).__name__ == "OperatorStatsLogger":
                lg.operators = list(self.callable_operators.values())
        self.i_step_start = 0
        self.previous_operator = None

        # ---- device side -------------------------------------------------------------
        missing = [n for n in OP_NAMES if n not in self.callable_operators]
        extra = [n for n in self.callable_operators if n not in OP_NAMES]
        if missing or extra:
            raise SbError(f"operator schedule differs from the one the device implements: missing {missing}, unknown {extra}")
        self.prob = refbridge.problem_from_reference(data, model)
        if device is None:
            device = int(os.environ.get("SBAYES_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little") >> 1
        self.seed = seed
        self.ctx = DeviceContext(self.prob, 1, device=device, team=team, _lib=_lib)
        w = np.array([self.callable_operators[n].weight for n in OP_NAMES], dtype=np.float64)
        self._cfg = _ExplicitWeights(temperature=float(temperature), prior_temperature=float(prior_temperature),
                                     sample_from_prior=bool(sample_from_prior), seed=seed, weights=w / w.sum(),
                                     p_grow=float(getattr(self.callable_operators["cluster_gibbsish_geo"], "p_grow", 0.5)))
        self.ctx.configure(0, self._cfg)
        self._stats0 = (np.zeros(N_OPS, np.int64), np.zeros(N_OPS, np.int64))
        self._uploaded = None          # the Sample object whose state currently sits on the device

    # ---- reference API ---------------------------------------------------------------------
    def get_operators(self, operators) -> dict:
        """The reference's own operator objects (sbayes/sampling/operators.py:28-209): they
        carry the schedule weights and receive the accept / reject statistics."""
        from sbayes.sampling.operators import get_operator_schedule
        return get_operator_schedule(operators_config=operators, model=self.model, data=self.data,
                                     temperature=self.temperature, prior_temperature=self.prior_temperature,
                                     sample_from_prior=self.sample_from_prior)

    def _ensure_uploaded(self, sample):
        if self._uploaded is not sample:
            st = refbridge.state_from_sample(self.prob, sample)
            self.ctx.set_state(0, st, i_step=int(sample.i_step))
            self._eval = self.ctx.eval(0)
            self._uploaded = sample
        return self._eval

    def prior(self, sample) -> float:
        """log-prior / prior_temperature (sbayes/sampling/mcmc_chain.py:74-88)."""
        ev = self._ensure_uploaded(sample)
        sample.last_prior = ev["prior"]
        return ev["prior"] / self.prior_temperature

    def likelihood(self, sample) -> float:
        """log-likelihood / temperature (sbayes/sampling/mcmc_chain.py:90-109)."""
        if self.sample_from_prior:
            sample.last_lh = 0.0
            return 0.0
        ev = self._ensure_uploaded(sample)
        sample.last_lh = ev["ll"]
        return ev["ll"] / self.temperature

    def run(self, n_steps: int, logging_interval: int, initial_sample, log_memory_usage: bool = True):
        """sbayes/sampling/mcmc_chain.py:128-172: steps i_step+1 .. i_step+n_steps; the
        loggers see the state after every step whose index is a multiple of `logging_interval`."""
        sample = initial_sample
        self._uploaded = None
        self._ll = self.likelihood(sample)
        self._prior = self.prior(sample)
        self.t_start = _time.time()
        self.i_step_start = sample.i_step + 1
        i, end = sample.i_step, sample.i_step + n_steps
        while i < end:
            nxt = min(end, (i // logging_interval + 1) * logging_interval)
            try:
                self.ctx.run(nxt - i)
            except SbError as e:
                if "non-finite" in str(e):
                    raise ValueError(str(e))          # mcmc_chain.py:231-236
                raise
            crossed_screen = (nxt + 1) // self.screen_log_interval > (i + 1) // self.screen_log_interval
            i = nxt
            if i % logging_interval == 0 or i == end:
                sample = self._download(sample, i)
                if i % logging_interval == 0:
                    for lg in self.sample_loggers:
                        lg.write_sample(sample)
            if crossed_screen:
                self.print_screen_log(i + 1, sample)
        return sample

    def _download(self, template, i_step: int):
        st = self.ctx.get_state(0)
        post = self.ctx.posterior(0)
        s = refbridge.sample_from_state(self.prob, self.model, self.data, st, type(template), i_step=i_step,
                                        chain=getattr(template, "chain", 0))
        # fill the reference's cache nodes (read by ParametersCSVLogger, loggers.py:157-262)
        self.model.prior(sample=s)
        if not self.sample_from_prior:
            self.model.likelihood(sample=s, caching=True)
        s.last_lh = 0.0 if self.sample_from_prior else post["ll"]
        s.last_prior = post["prior"]
        self._ll = s.last_lh / self.temperature
        self._prior = s.last_prior / self.prior_temperature
        self._uploaded, self._eval = s, post
        self._sync_operator_stats()
        return s

    def _sync_operator_stats(self):
        a, r = self.ctx.op_stats(0)
        cyc = self.ctx.op_cycles(0)
        for i, name in enumerate(OP_NAMES):
            op = self.callable_operators[name]
            da, dr = int(a[i] - self._stats0[0][i]), int(r[i] - self._stats0[1][i])
            op.accepts += da
            op.rejects += dr
            n = da + dr
            if n:   # mean device time per step of this operator since the last sync
                op.step_times.extend([float(cyc[i]) / 1.965e9 / max(int(a[i] + r[i]), 1)] * n)
        self._stats0 = (a, r)

    def close_loggers(self):
        for lg in self.sample_loggers:
            lg.close()

    def print_screen_log(self, i_step, sample):
        if self.screen_logger is None:
            return
        t = (_time.time() - self.t_start) / (i_step + 1 - self.i_step_start) * 1000000
        self.screen_logger.info(f"{i_step:<12}log-likelihood:  {self._ll:<19.2f}{t:.0f} seconds / million steps")

    def reset_posterior_cache(self):
        self._ll = -np.inf
        self._prior = -np.inf
        self._uploaded = None

    def close(self):
        self.ctx.close()


class _ExplicitWeights(ChainConfig):
    """ChainConfig whose operator weights come from the reference's schedule objects."""

    def __init__(self, weights, **kw):
        super().__init__(**kw)
        self._w = np.asarray(weights, dtype=np.float64)

    def op_weight_vector(self, n_clusters: int) -> np.ndarray:
        return self._w


def install():
    """Make the reference use the device chain: `sbayes.sampling.mcmc_chain.MCMCChain` and the
    name imported into `sbayes.mcmc_setup` (sbayes/mcmc_setup.py:25) are rebound."""
    import sbayes.model  # noqa: F401  (must be imported first: circular imports in the reference)
    import sbayes.sampling.mcmc_chain as mc
    import sbayes.mcmc_setup as ms
    mc.MCMCChain = MCMCChain
    ms.MCMCChain = MCMCChain
    return MCMCChain
