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
import threading
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
        # The device addresses every draw by (Philox key, step index).  The reference restarts the
        # step index on the SAME chain object (every warm-up attempt of MCMCChainProcess.initialize_sample
        # starts at i_step 0, and the best sample is reset to 0 before sampling, mcmc_setup.py:503-523),
        # so the key must change whenever the step index does not continue where the last run ended:
        # `_epoch` counts those restarts and is mixed into the key (epoch 0 = the seed itself).
        self._epoch = 0
        self._next_i_step = None
        self._restart = False

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
        self._rekey(int(sample.i_step))
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
        self._next_i_step = end
        return sample

    def _rekey(self, i_step: int):
        """new Philox key when the step index restarts (see __init__)"""
        if self._restart or (self._next_i_step is not None and i_step != self._next_i_step):
            self._epoch += 1
            z = (self.seed + 0x9E3779B97F4A7C15 * self._epoch) & 0xFFFFFFFFFFFFFFFF      # splitmix64
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
            self._cfg.seed = (z ^ (z >> 31)) >> 1
            self.ctx.configure(0, self._cfg)
        self._restart = False

    def _download(self, template, i_step: int):
        # one device round trip: state + the values of the reference's cache nodes (read by
        # ParametersCSVLogger, loggers.py:157-262) as the device computed them -- no CPU re-evaluation
        exp = self.ctx.export_samples([0])[0]
        st, post = exp["state"], exp["post"]
        s = refbridge.sample_from_state(self.prob, self.model, self.data, st, type(template), i_step=i_step,
                                        chain=getattr(template, "chain", 0))
        refbridge.fill_cache_from_device(self.prob, s, exp, self.sample_from_prior)
        _attach_effects(self.ctx, self.prob, s, 0, i_step)
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
                op.step_times.extend([float(cyc[i]) / self.ctx.clock_hz() / max(int(a[i] + r[i]), 1)] * n)
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
        self._restart = True

    def close(self):
        self.ctx.close()


class MCMC:
    """Drop-in for `sbayes.sampling.mcmc.MCMC` (boundary B1', sbayes/sampling/mcmc.py:36-361): the
    non-MC3 sampler and the warm-up of `MCMCSetup.sample` (sbayes/mcmc_setup.py:80-142).

    The `n_chains` chains (all at temperature 1, independent Philox streams) live in ONE device
    context and advance together in one kernel launch per logging interval; `generate_samples`
    keeps the reference's contract: warm-up returns the sample of the chain with the highest
    log-likelihood, sampling logs chain 0 after every step whose index is a multiple of
    `ceil(n_steps / n_samples)` and closes the loggers at the end."""

    CHECK_CACHING = __debug__

    def __init__(self, model, data, operators, sample_loggers: typ.List, n_chains: int = 1, sample_from_prior: bool = False,
                 show_screen_log: bool = False, logger: logging.Logger = None, screen_log_interval: int = 1000, *,
                 device: typ.Optional[int] = None, seed: typ.Optional[int] = None, team: int = 0, _lib=None, **kwargs):
        self.model, self.data = model, data
        self.n_chains = n_chains
        self.chain_idx = list(range(n_chains))
        self.sample_from_prior = sample_from_prior
        self.callable_operators = self.get_operators(operators)
        self.sample_loggers = sample_loggers
        self._ll = np.full(n_chains, -np.inf)
        self._prior = np.full(n_chains, -np.inf)
        self.show_screen_log = show_screen_log
        self.screen_log_interval = screen_log_interval
        self.logger = logger if logger is not None else logging.getLogger("sbayes_b200")
        self.t_start = _time.time()
        self.i_step_start = 0
        for lg in self.sample_loggers:
            if type(lg).__name__ == "OperatorStatsLogger":
                lg.operators = list(self.callable_operators.values())
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
        self.ctx = DeviceContext(self.prob, n_chains, device=device, team=team, _lib=_lib)
        w = np.array([self.callable_operators[n].weight for n in OP_NAMES], dtype=np.float64)
        for c in self.chain_idx:
            self.ctx.configure(c, _ExplicitWeights(sample_from_prior=bool(sample_from_prior), seed=seed + c, weights=w / w.sum(),
                                                   p_grow=float(getattr(self.callable_operators["cluster_gibbsish_geo"], "p_grow", 0.5))))
        self._stats0 = (np.zeros(N_OPS, np.int64), np.zeros(N_OPS, np.int64))

    def get_operators(self, operators) -> dict:
        from sbayes.sampling.operators import get_operator_schedule
        return get_operator_schedule(operators_config=operators, model=self.model, data=self.data,
                                     sample_from_prior=self.sample_from_prior)

    # -- reference API: un-tempered posterior terms of a sample (mcmc.py:99-138) -----------------
    def _evaluate(self, sample, chain: int) -> dict:
        self.ctx.set_state(chain, refbridge.state_from_sample(self.prob, sample), i_step=int(sample.i_step))
        return self.ctx.eval(chain)

    def prior(self, sample, chain: int) -> float:
        ev = self._evaluate(sample, chain)
        sample.last_prior = ev["prior"]
        return ev["prior"]

    def likelihood(self, sample, chain: int) -> float:
        if self.sample_from_prior:
            sample.last_lh = 0.0
            return 0.0
        ev = self._evaluate(sample, chain)
        sample.last_lh = ev["ll"]
        return ev["ll"]

    def _download(self, template, chain: int, i_step: int):
        exp = self.ctx.export_samples([chain])[0]
        s = refbridge.sample_from_state(self.prob, self.model, self.data, exp["state"], type(template), i_step=i_step, chain=chain)
        refbridge.fill_cache_from_device(self.prob, s, exp, self.sample_from_prior)
        _attach_effects(self.ctx, self.prob, s, chain, i_step)
        self._ll[chain], self._prior[chain] = s.last_lh, s.last_prior
        return s

    def _sync_operator_stats(self):
        a = np.zeros(N_OPS, np.int64); r = np.zeros(N_OPS, np.int64); cyc = np.zeros(N_OPS, np.int64)
        for c in self.chain_idx:
            ac, rc = self.ctx.op_stats(c)
            a += ac; r += rc; cyc += self.ctx.op_cycles(c)
        for i, name in enumerate(OP_NAMES):
            op = self.callable_operators[name]
            da, dr = int(a[i] - self._stats0[0][i]), int(r[i] - self._stats0[1][i])
            op.accepts += da
            op.rejects += dr
            if da + dr:   # mean device time per step of this operator (Operator.step_times, operators.py:232-293)
                op.step_times.extend([float(cyc[i]) / self.ctx.clock_hz() / max(int(a[i] + r[i]), 1)] * (da + dr))
        self._stats0 = (a, r)

    def _run(self, n: int):
        try:
            self.ctx.run(n)
        except SbError as e:
            if "non-finite" in str(e):
                raise ValueError(str(e))              # mcmc.py:307-312
            raise

    def generate_samples(self, n_steps: int, n_samples: int, initial_sample=None, initializer_logger=None, initializer=None,
                         warm_up: bool = False, warm_up_steps: typ.Optional[int] = None, log_memory_usage: bool = True):
        """sbayes/sampling/mcmc.py:155-263."""
        from copy import deepcopy
        sample = []
        for c in self.chain_idx:
            if initial_sample is None:
                if initializer is None:
                    raise AttributeError("Either initial_sample or initializer needs to be defined.")
                s = initializer.generate_sample(c)
            else:
                s = deepcopy(initial_sample)
                s.chain = c
            sample.append(s)
            self._ll[c] = self.likelihood(s, c)       # uploads the state of chain c
            self._prior[c] = self.prior(s, c)
        self.t_start = _time.time()
        if warm_up:
            self.logger.info("Tuning parameters in warm-up...")
            if warm_up_steps:
                self._run(int(warm_up_steps))
            for c in self.chain_idx:
                p = self.ctx.posterior(c)
                self._ll[c] = 0.0 if self.sample_from_prior else p["ll"]
                self._prior[c] = p["prior"]
            best = int(np.argmax(self._ll))
            self.logger.info(f"Starting state taken from warmup chain {best} with log-likelihood {self._ll[best]} "
                             f"(log-lokelihoods of all chains: {list(self._ll)}).")
            self._sync_operator_stats()
            return self._download(sample[best], best, max(int(warm_up_steps or 0) - 1, 0))
        self.logger.info("Sampling from posterior...")
        steps_per_sample = int(np.ceil(n_steps / n_samples))
        self.i_step_start = start = int(sample[0].i_step)
        i = start                                     # steps with index < i are done
        out = sample[0]
        while i < n_steps:
            # run up to and including the next logged step index (a multiple of steps_per_sample)
            nxt = i if i % steps_per_sample == 0 else (i // steps_per_sample + 1) * steps_per_sample
            upto = min(nxt, n_steps - 1)
            self._run(upto - i + 1)
            crossed = (upto + 1) // self.screen_log_interval > i // self.screen_log_interval
            i = upto + 1
            if upto % steps_per_sample == 0 or i == n_steps:
                out = self._download(sample[0], 0, upto)
                self._sync_operator_stats()            # OperatorStatsLogger rewrites its file at every sample
                if upto % steps_per_sample == 0:
                    for lg in self.sample_loggers:
                        lg.write_sample(out)
            if crossed:
                self.print_screen_log(i, sample)
        self._sync_operator_stats()
        for lg in self.sample_loggers:
            lg.close()
        self.logger.info(f"MCMC run finished after {(_time.time() - self.t_start):.1f} seconds")
        return out

    def print_screen_log(self, i_step, sample):
        t = (_time.time() - self.t_start) / max(i_step - self.i_step_start, 1) * 1000000
        self.logger.info(f"{i_step:<12}log-likelihood of chain 0:  {self._ll[0]:<19.2f}{t:.0f} seconds / million steps")

    def close(self):
        self.ctx.close()


class DeviceModel:
    """`Model.likelihood(sample, caching=True)` / `Model.prior(sample, caching=True)`
    (sbayes/model/model.py:47-51) evaluated on the device: the calls `swap_chains`
    (sbayes/mcmc_setup.py:389-390,418) and `SbayesInitializer.generate_sample`
    (sbayes/sampling/initializers.py:231) make directly on the model (boundary B2).  Every
    call is a full uncached evaluation (`sb_eval_batch`); `caching` is accepted and ignored."""

    def __init__(self, model, data, *, device: int = 0, _lib=None):
        self.model, self.data = model, data
        self.prob = refbridge.problem_from_reference(data, model)
        self.ctx = DeviceContext(self.prob, 1, device=device, _lib=_lib)
        self.ctx.configure(0, ChainConfig())
        self._last = (None, None)

    def _evaluate(self, sample) -> dict:
        key = id(sample), getattr(sample, "i_step", None)
        if self._last[0] != key or not all(getattr(sample, p).version == v for p, v in self._last[2].items()):
            self.ctx.set_state(0, refbridge.state_from_sample(self.prob, sample), i_step=int(sample.i_step))
            ev = self.ctx.eval_batch()[0]
            self._last = (key, ev, {p: getattr(sample, p).version for p in ("clusters", "weights", "source")})
        return self._last[1]

    def likelihood(self, sample, caching: bool = True) -> float:
        ev = self._evaluate(sample)
        sample.last_lh = ev["ll"]
        return ev["ll"]

    def prior(self, sample, caching: bool = True) -> float:
        ev = self._evaluate(sample)
        sample.last_prior = ev["prior"]
        return ev["prior"]

    def __call__(self, sample, caching: bool = True) -> float:
        return self.likelihood(sample, caching) + self.prior(sample, caching)

    def close(self):
        self.ctx.close()


class _OperatorEngine:
    """One device context that executes single operator calls (boundary B2, operator API)."""

    def __init__(self, model, data, temperature, prior_temperature, sample_from_prior, p_grow, device, seed, _lib):
        self.model, self.data = model, data
        self.prob = refbridge.problem_from_reference(data, model)
        self.ctx = DeviceContext(self.prob, 1, device=device, _lib=_lib)
        self.base = dict(temperature=float(temperature), prior_temperature=float(prior_temperature),
                         sample_from_prior=bool(sample_from_prior), seed=int(seed), p_grow=float(p_grow))
        self.calls = 0
        self.last_record = None

    def propose(self, op_index: int, sample):
        """Operator.function(sample) (sbayes/sampling/operators.py:269-276): one step of operator
        `op_index` with the Metropolis-Hastings test switched off (SB_CFG_PROPOSE_ONLY)."""
        w = np.zeros(N_OPS)
        w[op_index] = 1.0
        cfg = _ExplicitWeights(weights=w, **self.base)
        cfg.propose_only = True
        self.ctx.configure(0, cfg)
        self.ctx.set_state(0, refbridge.state_from_sample(self.prob, sample), i_step=self.calls)   # draws of step calls + 1
        self.calls += 1
        rec = self.ctx.run(1, records=True)[0, 0]
        self.last_record = rec
        lq, lqb = float(rec["log_q"]), float(rec["log_q_back"])
        if not rec["changed"] or not rec["accept"]:
            return sample, 0, -np.inf                 # Q_REJECT, Q_BACK_REJECT (operators.py:245-247)
        st = self.ctx.get_state(0)
        new = refbridge.sample_from_state(self.prob, self.model, self.data, st, type(sample), i_step=int(sample.i_step),
                                          chain=getattr(sample, "chain", 0))
        return new, lq, lqb


def device_operator_schedule(operators, model, data, temperature: float = 1.0, prior_temperature: float = 1.0,
                             sample_from_prior: bool = False, *, device: int = 0, seed: typ.Optional[int] = None, _lib=None) -> dict:
    """`get_operator_schedule` (sbayes/sampling/operators.py:28-209) with every operator's proposal
    executed on the device.  The returned objects ARE the reference's operator objects (same
    classes, weights, statistics API, parameter strings); only `_propose` is rebound, so
    `op.function(sample) -> (sample', log_q, log_q_back)` keeps the reference's contract including
    the sentinels (Q_GIBBS / Q_BACK_GIBBS for the Gibbs operators, Q_REJECT / Q_BACK_REJECT when
    the operator rejects by itself)."""
    from sbayes.sampling.operators import get_operator_schedule
    ops = get_operator_schedule(operators_config=operators, model=model, data=data, temperature=temperature,
                                prior_temperature=prior_temperature, sample_from_prior=sample_from_prior)
    missing = [n for n in OP_NAMES if n not in ops]
    extra = [n for n in ops if n not in OP_NAMES]
    if missing or extra:
        raise SbError(f"operator schedule differs from the one the device implements: missing {missing}, unknown {extra}")
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little") >> 1
    engine = _OperatorEngine(model, data, temperature, prior_temperature, sample_from_prior,
                             getattr(ops["cluster_gibbsish_geo"], "p_grow", 0.5), device, seed, _lib)
    for i, name in enumerate(OP_NAMES):
        op = ops[name]

        def _propose(self, sample, _i=i, **kwargs):
            return engine.propose(_i, sample)
        op.__class__ = type(type(op).__name__, (type(op),), {"_propose": _propose, "device_engine": engine})
    return ops


class _ExplicitWeights(ChainConfig):
    """ChainConfig whose operator weights come from the reference's schedule objects."""

    def __init__(self, weights, **kw):
        super().__init__(**kw)
        self._w = np.asarray(weights, dtype=np.float64)

    def op_weight_vector(self, n_clusters: int) -> np.ndarray:
        return self._w


def _attach_effects(ctx, prob, sample, chain: int, i_step: int):
    """The effect draws the parameter logger asks for (conditional_effect_sample, sbayes/sampling/conditionals.py:125-149,
    called at sbayes/sampling/loggers.py:162 and :217) taken on the device from the chain's committed counts: one array per
    call the logger will make (cluster effect, then every confounder in order).  `install_effect_logging` hands them out."""
    p = ctx.effect_sample([chain], step=int(i_step))[0]
    offs = prob.group_offsets()
    sample._sb_effects = [p[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]


_effect_state = threading.local()
effect_logging_stats = {"device": 0, "host": 0}       # how many effect arrays the logger took from the device / drew on the host


def install_effect_logging():
    """Rebind `conditional_effect_sample` inside the reference's logger module: while a `ParametersCSVLogger` writes a sample
    that came from the device chain, its calls return the device's draws (in call order) instead of rebuilding a
    [groups, objects, features, states] boolean contraction per effect on the host (conditionals.py:68-86); any other
    sample takes the reference's own path."""
    import sbayes.sampling.loggers as lg
    if getattr(lg, "_sb_effects_installed", False):
        return
    orig_sample, orig_write = lg.conditional_effect_sample, lg.ParametersCSVLogger._write_sample

    def conditional_effect_sample(features, is_source_group, applicable_states, prior_counts=0.1):
        queue = getattr(_effect_state, "queue", None)
        if queue:
            effect_logging_stats["device"] += 1
            return queue.pop(0)
        effect_logging_stats["host"] += 1
        return orig_sample(features=features, is_source_group=is_source_group, applicable_states=applicable_states, prior_counts=prior_counts)

    def _write_sample(self, sample):
        eff = getattr(sample, "_sb_effects", None)
        _effect_state.queue = list(eff) if eff is not None else None
        try:
            return orig_write(self, sample)
        finally:
            _effect_state.queue = None

    lg.conditional_effect_sample = conditional_effect_sample
    lg.ParametersCSVLogger._write_sample = _write_sample
    lg._sb_effects_installed = True


def install(*, seed: typ.Optional[int] = None, device: typ.Optional[int] = None, _lib=None):
    """Make the reference use the device chain: `sbayes.sampling.mcmc_chain.MCMCChain` and the
    name imported into `sbayes.mcmc_setup` (sbayes/mcmc_setup.py:25) are rebound; so are
    `sbayes.sampling.mcmc.MCMC` and its import in `sbayes.mcmc_setup` (warm-up and non-MC3 runs).
    After this call `sbayes.cli.main(config.yaml)` (sbayes/cli.py:17-44) runs unchanged with the
    MCMC steps on the B200: `MCMCSetup.sample` builds the device `MCMC`, `MCMCSetup.sample_mc3`
    forks one `MCMCChainProcess` per chain and each child builds a device `MCMCChain` (its own
    context on the GPU; the parent never touches CUDA, so the fork is safe).

    `seed` fixes the Philox keys (reproducible runs; the reference itself is unseeded), `device`
    the GPU; `_lib` is for the tests of the GPU-less build container (fibre-emulator build of the
    same device code) -- the product default is the CUDA library and nothing else."""
    import sbayes.model  # noqa: F401  (must be imported first: circular imports in the reference)
    import sbayes.sampling.mcmc_chain as mc
    import sbayes.mcmc_setup as ms
    import sbayes.sampling.mcmc as mm
    defaults = dict(seed=seed, device=device, _lib=_lib)

    class InstalledMCMCChain(MCMCChain):
        def __init__(self, *a, **kw):
            for k, v in defaults.items():
                if v is not None:
                    kw.setdefault(k, v)
            if kw.get("seed") is not None:     # one stream per temperature of the ladder
                kw["seed"] = int(kw["seed"]) + 7919 * int(round(1e6 * float(kw.get("temperature", 1.0))))
            super().__init__(*a, **kw)

    class InstalledMCMC(MCMC):
        _instances = 0

        def __init__(self, *a, **kw):
            for k, v in defaults.items():
                if v is not None:
                    kw.setdefault(k, v)
            if kw.get("seed") is not None:     # the warm-up sampler and the sampler proper must not share streams
                kw["seed"] = int(kw["seed"]) + 1000003 * InstalledMCMC._instances
            InstalledMCMC._instances += 1
            super().__init__(*a, **kw)

    InstalledMCMCChain.__name__ = InstalledMCMCChain.__qualname__ = "MCMCChain"
    InstalledMCMC.__name__ = InstalledMCMC.__qualname__ = "MCMC"
    mc.MCMCChain = InstalledMCMCChain
    ms.MCMCChain = InstalledMCMCChain
    mm.MCMC = InstalledMCMC
    ms.MCMC = InstalledMCMC
    install_effect_logging()
    return InstalledMCMCChain
