"""ctypes binding of the C ABI in include/sbayes_b200.h (libsbayes_b200.so, sm_100a CUDA).

There is no CPU fallback: `load()` raises if the CUDA library has not been built and
`DeviceContext` raises if the library cannot create a context on a B200.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from .packed import Problem, ChainState, ChainConfig, NA, N_OPS

N_SITES = 14
_HERE = os.path.dirname(os.path.abspath(__file__))
# SBAYES_B200_LIB: another build of the SAME CUDA library (profiling variant with phase counters); there is no
# non-CUDA implementation to point it at
LIB_PATH = os.environ.get("SBAYES_B200_LIB") or os.path.join(_HERE, "csrc", "libsbayes_b200.so")


class SbError(RuntimeError):
    pass


class _ProblemDesc(C.Structure):
    _fields_ = [
        ("n_objects", C.c_int32), ("n_features", C.c_int32), ("n_states", C.c_int32), ("n_clusters", C.c_int32),
        ("n_confounders", C.c_int32),
        ("x", C.c_void_p), ("applicable", C.c_void_p), ("conf_gid", C.c_void_p), ("conf_ngroups", C.c_void_p),
        ("alpha", C.c_void_p), ("alpha_w", C.c_void_p),
        ("min_size", C.c_int32), ("max_size", C.c_int32), ("size_prior", C.c_int32), ("weights_prior_uniform", C.c_int32),
        ("geo_enabled", C.c_int32), ("geo_aggregation", C.c_int32), ("geo_prob_fn", C.c_int32),
        ("geo_rate", C.c_double), ("geo_inflection", C.c_double),
        ("geo_cost", C.c_void_p), ("adj_indptr", C.c_void_p), ("adj_indices", C.c_void_p),
    ]


class _ChainCfg(C.Structure):
    _fields_ = [
        ("temperature", C.c_double), ("prior_temperature", C.c_double), ("sample_from_prior", C.c_int32),
        ("flags", C.c_int32), ("op_weights", C.c_double * N_OPS), ("p_grow", C.c_double), ("seed", C.c_uint64),
    ]


class _Tape(C.Structure):
    _fields_ = [("vals", C.c_void_p), ("off", C.c_void_p), ("len", C.c_void_p), ("wide_rows", C.c_void_p),
                ("n_vals", C.c_int64)]


class _SampleOut(C.Structure):
    _fields_ = [("cluster_of", C.c_void_p), ("W", C.c_void_p), ("src", C.c_void_p), ("counts", C.c_void_p), ("post", C.c_void_p),
                ("group_ll", C.c_void_p), ("geo_prior", C.c_void_p), ("source_prior", C.c_void_p)]


class _Posterior(C.Structure):
    _fields_ = [("ll", C.c_double), ("prior", C.c_double), ("size_prior", C.c_double), ("geo_prior", C.c_double),
                ("weights_prior", C.c_double), ("source_prior", C.c_double)]


RECORD_DTYPE = np.dtype([("op", "<i4"), ("accept", "<i4"), ("changed", "<i4"), ("n_changed", "<i4"),
                         ("log_q", "<f8"), ("log_q_back", "<f8"), ("ll_new", "<f8"), ("prior_new", "<f8"),
                         ("mh_ratio", "<f8"), ("log_u", "<f8")])

EXPORTS = ["sb_create", "sb_destroy", "sb_last_error", "sb_chain_config", "sb_set_state", "sb_get_state", "sb_eval",
           "sb_posterior", "sb_run", "sb_run_async", "sb_sync", "sb_op_stats", "sb_swap", "sb_get_slots",
           "sb_set_slots", "sb_counts_rebuild_mma", "sb_counts_rebuild_hist", "sb_info", "sb_stream",
           "sb_eval_batch", "sb_eval_stats", "sb_eval_terms", "sb_stats_buffer", "sb_eval_set_mode", "sb_obs_loglik",
           "sb_swap_scalars", "sb_clock_khz", "sb_posteriors", "sb_set_temperatures", "sb_gibbs_sweep",
           "sb_export_samples", "sb_shard_handle", "sb_shard_attach", "sb_run_group", "sb_init_em", "sb_eval_info",
           "sb_effect_sample"]


def _bind(lib):
    vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
    lib.sb_create.argtypes = [C.POINTER(_ProblemDesc), i32, i32, i32, C.POINTER(vp)]
    lib.sb_destroy.argtypes = [vp]
    lib.sb_destroy.restype = None
    lib.sb_last_error.argtypes = [vp]
    lib.sb_last_error.restype = C.c_char_p
    lib.sb_chain_config.argtypes = [vp, i32, C.POINTER(_ChainCfg)]
    lib.sb_set_state.argtypes = [vp, i32, vp, vp, vp, i64]
    lib.sb_get_state.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.sb_eval.argtypes = [vp, i32, C.POINTER(_Posterior)]
    lib.sb_posterior.argtypes = [vp, i32, C.POINTER(_Posterior)]
    lib.sb_gibbs_sweep.argtypes = [vp, u64, i32, vp, vp, vp]
    lib.sb_export_samples.argtypes = [vp, i32, vp, C.POINTER(_SampleOut)]
    lib.sb_shard_handle.argtypes = [vp, vp]
    lib.sb_shard_attach.argtypes = [vp, i32, i32, vp]
    lib.sb_run_group.argtypes = [vp, i32, i64, vp]
    lib.sb_init_em.argtypes = [vp, i32, vp, vp, vp]
    lib.sb_eval_info.argtypes = [vp, vp]
    lib.sb_effect_sample.argtypes = [vp, i32, vp, u64, vp, vp]
    lib.sb_posteriors.argtypes = [vp, vp]
    lib.sb_set_temperatures.argtypes = [vp, vp, vp]
    lib.sb_run.argtypes = [vp, i64, C.POINTER(_Tape), vp]
    lib.sb_run_async.argtypes = [vp, i64]
    lib.sb_sync.argtypes = [vp]
    lib.sb_op_stats.argtypes = [vp, i32, vp, vp, vp]
    lib.sb_swap.argtypes = [vp, i32, i32, u64, i64, vp, vp]
    lib.sb_swap_scalars.argtypes = [i32, vp, vp, vp, vp, i32, i32, u64, i64, vp, vp, vp, vp, vp]
    lib.sb_get_slots.argtypes = [vp, vp]
    lib.sb_set_slots.argtypes = [vp, vp]
    lib.sb_counts_rebuild_mma.argtypes = [vp, i32, vp, vp]
    lib.sb_counts_rebuild_hist.argtypes = [vp, i32, vp, vp]
    lib.sb_info.argtypes = [vp, vp]
    lib.sb_stream.argtypes = [vp]
    lib.sb_stream.restype = u64
    lib.sb_clock_khz.argtypes = [vp]
    lib.sb_clock_khz.restype = i64
    lib.sb_eval_batch.argtypes = [vp, vp, vp]
    lib.sb_eval_stats.argtypes = [vp, vp]
    lib.sb_eval_terms.argtypes = [vp, i64, vp, vp]
    lib.sb_stats_buffer.argtypes = [vp, vp, vp, vp, vp]
    lib.sb_eval_set_mode.argtypes = [vp, i32]
    lib.sb_obs_loglik.argtypes = [vp, vp, vp, vp, vp]
    return lib


_lib = None


def load():
    """Open libsbayes_b200.so (built by `__graft_entry__.build()`); no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SbError(f"{LIB_PATH} is missing: build the CUDA library first (python -c 'import __graft_entry__ as g; "
                          f"g.build()'); sbayes_b200 has no CPU fallback")
        _lib = _bind(C.CDLL(LIB_PATH))
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def run_group(ctxs: Sequence["DeviceContext"], n_steps: int, records: bool = False):
    """All ranks of a sharded chain held by ONE process advance together (`sb_run_group`)."""
    lib = ctxs[0].lib
    arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    rec = np.zeros((ctxs[0].n_chains, n_steps), RECORD_DTYPE) if records else None
    rc = lib.sb_run_group(C.cast(arr, C.c_void_p), len(ctxs), int(n_steps), _ptr(rec))
    if rc != 0:
        msgs = [lib.sb_last_error(c.h).decode() for c in ctxs]
        raise SbError(f"sb_run_group failed ({rc}): {msgs}")
    return rec


def swap_scalars(ll, prior, temps, prior_temps, state_at_slot, seed: int, round_: int, attempts: int = 100,
                 only_neighbours: bool = False, tape_pairs=None, tape_u=None, _lib=None):
    """swap_chains (sbayes/mcmc_setup.py:365-418) on host scalars through the C ABI
    (`sb_swap_scalars`).  `state_at_slot` (list or int32 array) is updated in place; returns
    the accepted (from, to) slot pairs."""
    lib = _lib if _lib is not None else load()
    n = len(state_at_slot)
    a = lambda v: np.ascontiguousarray(v, np.float64)
    ll, prior, temps, prior_temps = a(ll), a(prior), a(temps), a(prior_temps)
    sas = np.ascontiguousarray(state_at_slot, np.int32).copy()
    n_pairs = n - 1 if only_neighbours else n * (n - 1) // 2
    attempts = min(int(attempts), n_pairs)
    tp = None if tape_pairs is None else np.ascontiguousarray(tape_pairs, np.int32)
    tu = None if tape_u is None else np.ascontiguousarray(tape_u, np.float64)
    if tp is not None and (tp.shape != (attempts, 2) or tu is None or tu.shape != (attempts,)):
        raise SbError("swap tape must hold `attempts` pairs and uniforms")
    acc, n_acc = np.zeros((max(attempts, 1), 2), np.int32), np.zeros(1, np.int32)
    rc = lib.sb_swap_scalars(n, _ptr(ll), _ptr(prior), _ptr(temps), _ptr(prior_temps), attempts, int(only_neighbours), int(seed),
                             int(round_), _ptr(tp), _ptr(tu), _ptr(sas), _ptr(acc), _ptr(n_acc))
    if rc != 0:
        raise SbError(f"sb_swap_scalars failed ({rc})")
    for i in range(n):
        state_at_slot[i] = int(sas[i])
    return [tuple(int(v) for v in p) for p in acc[:int(n_acc[0])]]


class DeviceContext:
    """One GPU: the packed problem plus `n_chains` chain states resident in HBM.

    Mirrors what a reference chain sees of `Data` and `Model`
    (sbayes/load_data.py:187-256, sbayes/model/model.py:24-45)."""

    def __init__(self, prob: Problem, n_chains: int = 1, device: int = 0, team: int = 0, _lib=None):
        self.lib = _lib if _lib is not None else load()
        self.prob = prob
        self.n_chains = n_chains
        self._keep = []

        def keep(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            self._keep.append(a)
            return a

        d = _ProblemDesc()
        d.n_objects, d.n_features, d.n_states, d.n_clusters = prob.N, prob.F, prob.S, prob.K
        d.n_confounders = len(prob.conf_gid)
        d.x = _ptr(keep(prob.x, np.uint8))
        d.applicable = _ptr(keep(prob.applicable, np.uint8))
        gid = np.stack(prob.conf_gid) if prob.conf_gid else np.zeros((0, prob.N), np.int32)
        d.conf_gid = _ptr(keep(gid, np.int32))
        d.conf_ngroups = _ptr(keep(np.array(prob.conf_ngroups, np.int32), np.int32))
        d.alpha = _ptr(keep(prob.alpha_stacked(), np.float64))
        d.alpha_w = _ptr(keep(prob.alpha_w, np.float64))
        d.min_size, d.max_size, d.size_prior = prob.min_size, prob.max_size, prob.size_prior
        d.weights_prior_uniform = int(prob.weights_prior_uniform)
        if prob.geo is not None:
            d.geo_enabled, d.geo_aggregation, d.geo_prob_fn = 1, prob.geo.aggregation, prob.geo.prob_fn
            d.geo_rate, d.geo_inflection = prob.geo.rate, prob.geo.inflection
            d.geo_cost = _ptr(keep(prob.geo.cost, np.float64))
        if prob.adj_indptr is not None:
            d.adj_indptr = _ptr(keep(prob.adj_indptr, np.int32))
            d.adj_indices = _ptr(keep(prob.adj_indices if len(prob.adj_indices) else np.zeros(1, np.int32), np.int32))
        h = C.c_void_p()
        rc = self.lib.sb_create(C.byref(d), device, n_chains, team, C.byref(h))
        if rc != 0:
            raise SbError(f"sb_create failed ({rc}): {self.lib.sb_last_error(None).decode()}")
        self.h = h
        self._keep.clear()

    # ------------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise SbError(f"sbayes_b200 error {rc}: {self.lib.sb_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.sb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        out = np.zeros(8, np.int32)
        self._ck(self.lib.sb_info(self.h, _ptr(out)))
        return dict(zip(["device", "n_chains", "team", "grid", "block", "smem", "n_sm", "launches"], out.tolist()))

    def stream(self) -> int:
        return int(self.lib.sb_stream(self.h))

    def clock_hz(self) -> float:
        """SM clock the per-operator cycle counters are converted with (cudaDevAttrClockRate)."""
        if getattr(self, "_clock_hz", None) is None:
            self._clock_hz = float(self.lib.sb_clock_khz(self.h)) * 1e3
        return self._clock_hz

    def configure(self, chain: int, cfg: ChainConfig):
        c = _ChainCfg()
        c.temperature, c.prior_temperature = cfg.temperature, cfg.prior_temperature
        c.sample_from_prior = int(cfg.sample_from_prior)
        c.flags = 1 if getattr(cfg, "propose_only", False) else 0
        w = cfg.op_weight_vector(self.prob.K)
        for i in range(N_OPS):
            c.op_weights[i] = float(w[i])
        c.p_grow, c.seed = cfg.p_grow, int(cfg.seed)
        self._ck(self.lib.sb_chain_config(self.h, chain, C.byref(c)))

    # ---- objects of a chain sharded over GPUs (include/sbayes_b200.h: sb_shard_*) -----------------
    def shard_handle(self) -> np.ndarray:
        h = np.zeros(128, np.uint8)
        self._ck(self.lib.sb_shard_handle(self.h, _ptr(h)))
        return h

    def shard_attach(self, rank: int, world: int, handles: np.ndarray):
        hs = np.ascontiguousarray(handles, np.uint8).reshape(world, 128)
        self._ck(self.lib.sb_shard_attach(self.h, int(rank), int(world), _ptr(hs)))
        self.shard_rank, self.shard_world = int(rank), int(world)

    def mark_active(self, chain: int):
        """rank > 0 of a sharded chain: the state itself is rank 0's"""
        self._ck(self.lib.sb_set_state(self.h, chain, None, None, None, 0))

    def set_state(self, chain: int, st: ChainState, i_step: int = 0):
        co = np.ascontiguousarray(st.cluster_of, np.int32)
        W = np.ascontiguousarray(st.W, np.float32)
        src = np.ascontiguousarray(st.src, np.uint8)
        assert co.shape == (self.prob.N,) and W.shape == (self.prob.F, self.prob.C) and src.shape == (self.prob.N, self.prob.F)
        self._ck(self.lib.sb_set_state(self.h, chain, _ptr(co), _ptr(W), _ptr(src), int(i_step)))

    def get_state(self, chain: int) -> ChainState:
        p = self.prob
        co, W = np.empty(p.N, np.int32), np.empty((p.F, p.C), np.float32)
        src, counts = np.empty((p.N, p.F), np.uint8), np.empty((p.G_tot, p.F, p.S), np.int32)
        self._ck(self.lib.sb_get_state(self.h, chain, _ptr(co), _ptr(W), _ptr(src), _ptr(counts)))
        return ChainState(co, W, src, counts)

    def _host_buffer(self, shape, dtype):
        """pinned host memory when torch sees a GPU (asynchronous device-to-host copies), plain numpy otherwise"""
        try:
            import torch
            if torch.cuda.is_available():
                return torch.empty(tuple(shape), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True).numpy()
        except Exception:
            pass
        return np.empty(shape, dtype)

    def export_samples(self, chains: Optional[Sequence[int]] = None, src: bool = True, counts: bool = True, caches: bool = True) -> list:
        """The committed states of several chains in ONE device round trip (`sb_export_samples`): a list of dicts
        {"state": ChainState, "post": {...}, "group_ll" [G_tot], "geo_prior" [K], "source_prior" [N], "i_step"}
        -- the cache-node values are what the reference's loggers read from a Sample (loggers.py:249-252)."""
        p = self.prob
        chains = list(range(self.n_chains)) if chains is None else [int(c) for c in chains]
        n = len(chains)
        key = (n, src, counts, caches)
        if getattr(self, "_exp_key", None) != key:          # the (pinned) buffers are reused from call to call
            B = self._host_buffer
            self._exp = dict(cluster_of=B((n, p.N), np.int32), W=B((n, p.F, p.C), np.float32), post=B((n, 8), np.float64),
                             src=B((n, p.N, p.F), np.uint8) if src else None,
                             counts=B((n, p.G_tot, p.F, p.S), np.int32) if counts else None,
                             group_ll=B((n, p.G_tot), np.float64) if caches else None,
                             geo_prior=B((n, p.K), np.float64) if caches else None,
                             source_prior=B((n, p.N), np.float64) if caches else None)
            self._exp_key = key
        b = self._exp
        o = _SampleOut()
        for k in ("cluster_of", "W", "src", "counts", "post", "group_ll", "geo_prior", "source_prior"):
            setattr(o, k, _ptr(b[k]))
        ch = np.ascontiguousarray(chains, np.int32)
        self._ck(self.lib.sb_export_samples(self.h, n, _ptr(ch), C.byref(o)))
        out = []
        for i in range(n):
            post = b["post"][i]
            st = ChainState(b["cluster_of"][i].copy(), b["W"][i].copy(), b["src"][i].copy() if src else None,
                            b["counts"][i].copy() if counts else None)
            out.append(dict(state=st, i_step=int(post[5]),
                            post=dict(ll=float(post[0]), size_prior=float(post[1]), geo_prior=float(post[2]), weights_prior=float(post[3]),
                                      source_prior=float(post[4]), prior=float(post[1] + post[2] + post[3] + post[4])),
                            group_ll=b["group_ll"][i].copy() if caches else None, geo_prior=b["geo_prior"][i].copy() if caches else None,
                            source_prior=b["source_prior"][i].copy() if caches else None))
        return out

    def eval(self, chain: int) -> dict:
        """Full uncached evaluation (count rebuild + likelihood + all prior terms)."""
        out = _Posterior()
        self._ck(self.lib.sb_eval(self.h, chain, C.byref(out)))
        return {k: getattr(out, k) for k, _ in _Posterior._fields_}

    @staticmethod
    def _posts(arr) -> list:
        return [{k: getattr(o, k) for k, _ in _Posterior._fields_} for o in arr]

    def eval_batch(self, timing: bool = False):
        """Uncached likelihood + prior of EVERY chain with the stand-alone evaluation kernels
        (statistics pass over the packed observations + terms).  Returns a list of dicts, and
        with timing=True also the device times in ms of (statistics pass, terms)."""
        out = (_Posterior * self.n_chains)()
        ms = np.zeros(2, np.float32)
        self._ck(self.lib.sb_eval_batch(self.h, C.cast(out, C.c_void_p), _ptr(ms) if timing else None))
        res = self._posts(out)
        return (res, (float(ms[0]), float(ms[1]))) if timing else res

    def eval_set_mode(self, mode: str):
        """Statistics pass of eval_batch / eval_stats: "hist" (shared-memory histogram) or "mma"
        (int8 tcgen05 contraction)."""
        self._ck(self.lib.sb_eval_set_mode(self.h, {"hist": 0, "mma": 1}[mode]))

    def eval_stats(self) -> float:
        """Statistics pass only (counts, source histogram, cluster sizes of every chain)."""
        ms = np.zeros(1, np.float32)
        self._ck(self.lib.sb_eval_stats(self.h, _ptr(ms)))
        return float(ms[0])

    def eval_terms(self, n_global: int = 0) -> list:
        """Posterior terms from the statistics currently held by the ctx."""
        out = (_Posterior * self.n_chains)()
        self._ck(self.lib.sb_eval_terms(self.h, int(n_global), C.cast(out, C.c_void_p), None))
        return self._posts(out)

    def obs_loglik(self, want_lh: bool = True, want_objects: bool = True):
        """Observation likelihood of every chain (LikelihoodLogger's product): returns
        (lh float32 [n_chains,N,F] or None, per-object sums of log lh [n_chains,N] or None,
        per-chain sums [n_chains], device ms)."""
        p = self.prob
        lh = np.empty((self.n_chains, p.N, p.F), np.float32) if want_lh else None
        ob = np.empty((self.n_chains, p.N), np.float64) if want_objects else None
        sums, ms = np.empty(self.n_chains, np.float64), np.zeros(1, np.float32)
        self._ck(self.lib.sb_obs_loglik(self.h, _ptr(lh), _ptr(ob), _ptr(sums), _ptr(ms)))
        return lh, ob, sums, float(ms[0])

    def gibbs_sweep(self, seed=None, step: int = 1, mode: str = "draw", full: bool = False):
        """Full-sweep source Gibbs of every chain (GibbsSampleSource over ALL objects).  mode "draw" resamples
        and commits every source (draws addressed by (chain seed, `step`)), "score_previous" / "score" return
        sum log q of the previous / committed sources under the current counts.  Returns the device ms, or with
        full=True (log_q [n_chains], n_changed [n_chains], ms).  `seed` is accepted for symmetry with the
        bench helper and offsets `step`."""
        lq, ch, ms = np.zeros(self.n_chains, np.float64), np.zeros(self.n_chains, np.int64), np.zeros(1, np.float32)
        st = int(step) + (int(seed) if seed is not None else 0)
        self._ck(self.lib.sb_gibbs_sweep(self.h, st, {"draw": 0, "score_previous": 1, "score": 2}[mode], _ptr(lq), _ptr(ch), _ptr(ms)))
        return (lq, ch, float(ms[0])) if full else float(ms[0])

    def init_em(self, z0: np.ndarray, n_steps: int, timing: bool = False):
        """EM-style cluster initializer (SbayesInitializer.generate_clusters_em): z0 float32 [G_tot, N] ->
        final soft assignment float32 [G_tot, N]."""
        z0 = np.ascontiguousarray(z0, np.float32)
        assert z0.shape == (self.prob.G_tot, self.prob.N)
        out, ms = np.empty_like(z0), np.zeros(1, np.float32)
        self._ck(self.lib.sb_init_em(self.h, int(n_steps), _ptr(z0), _ptr(out), _ptr(ms)))
        return (out, float(ms[0])) if timing else out

    def eval_info(self) -> dict:
        """which statistics pass `eval_stats` runs ("hist", "mma", "fine") and its launch geometry"""
        o = np.zeros(8, np.int32)
        self._ck(self.lib.sb_eval_info(self.h, _ptr(o)))
        return {"pass": ("hist", "mma", "fine")[int(o[0])], "ft": int(o[1]), "n_ftiles": int(o[2]), "n_oblk": int(o[3]),
                "smem": int(o[4]), "fine_rows": int(o[5]), "grid": int(o[6]), "block": int(o[7])}

    def effect_sample(self, chains=None, step: int = 0, timing: bool = False):
        """conditional_effect_sample of the reference's parameter logger for the listed chains: float64
        [n, G_tot, F, S] Dirichlet draws from (committed counts + prior concentration), rows = clusters, then every
        confounder's groups."""
        chains = np.arange(self.n_chains, dtype=np.int32) if chains is None else np.ascontiguousarray(chains, np.int32)
        out = np.empty((len(chains), self.prob.G_tot, self.prob.F, self.prob.S), np.float64)
        ms = np.zeros(1, np.float32)
        self._ck(self.lib.sb_effect_sample(self.h, len(chains), _ptr(chains), int(step), _ptr(out), _ptr(ms)))
        return (out, float(ms[0])) if timing else out

    def stats_buffer(self):
        """(device address, int32 elements, per-state stride, (n_counts, n_hist, n_sizes))."""
        ptr, n, stride, lay = C.c_uint64(), C.c_int64(), C.c_int64(), (C.c_int64 * 3)()
        self._ck(self.lib.sb_stats_buffer(self.h, C.byref(ptr), C.byref(n), C.byref(stride), C.cast(lay, C.c_void_p)))
        return int(ptr.value), int(n.value), int(stride.value), tuple(int(v) for v in lay)

    def posterior(self, chain: int) -> dict:
        out = _Posterior()
        self._ck(self.lib.sb_posterior(self.h, chain, C.byref(out)))
        return {k: getattr(out, k) for k, _ in _Posterior._fields_}

    def posteriors(self) -> list:
        """cached posterior terms of every chain (one device-to-host copy)"""
        out = (_Posterior * self.n_chains)()
        self._ck(self.lib.sb_posteriors(self.h, C.cast(out, C.c_void_p)))
        return self._posts(out)

    def set_temperatures(self, temps, prior_temps):
        t, tp = np.ascontiguousarray(temps, np.float64), np.ascontiguousarray(prior_temps, np.float64)
        assert t.shape == (self.n_chains,) and tp.shape == (self.n_chains,)
        self._ck(self.lib.sb_set_temperatures(self.h, _ptr(t), _ptr(tp)))

    def run(self, n_steps: int, tape: Optional[dict] = None, records: bool = False):
        """Advance every chain by n_steps.  `tape` = dict(vals, off, len, wide_rows) replays
        recorded draws (single chain); records=True returns one record per chain and step."""
        t = None
        keep = []
        if tape is not None:
            t = _Tape()
            vals = np.ascontiguousarray(tape["vals"], np.float64)
            off = np.ascontiguousarray(tape["off"][:n_steps], np.int32)
            ln = np.ascontiguousarray(tape["len"][:n_steps], np.int32)
            wr = np.ascontiguousarray(tape["wide_rows"][:n_steps], np.int32)
            keep += [vals, off, ln, wr]
            t.vals, t.off, t.len, t.wide_rows, t.n_vals = _ptr(vals), _ptr(off), _ptr(ln), _ptr(wr), vals.size
        rec = np.zeros((self.n_chains, n_steps), RECORD_DTYPE) if records else None
        self._ck(self.lib.sb_run(self.h, int(n_steps), C.byref(t) if t is not None else None, _ptr(rec)))
        return rec

    def run_async(self, n_steps: int):
        self._ck(self.lib.sb_run_async(self.h, int(n_steps)))

    def sync(self):
        self._ck(self.lib.sb_sync(self.h))

    def op_stats(self, chain: int):
        a, r = np.zeros(N_OPS, np.int64), np.zeros(N_OPS, np.int64)
        self._ck(self.lib.sb_op_stats(self.h, chain, _ptr(a), _ptr(r), None))
        return a, r

    def op_cycles(self, chain: int) -> np.ndarray:
        """SM cycles spent per operator (Operator.step_times, operators.py:232-293)."""
        a, r, cy = np.zeros(N_OPS, np.int64), np.zeros(N_OPS, np.int64), np.zeros(N_OPS, np.int64)
        self._ck(self.lib.sb_op_stats(self.h, chain, _ptr(a), _ptr(r), _ptr(cy)))
        return cy

    def swap(self, attempts: int, only_neighbours: bool, seed: int, round_: int):
        sos, n = np.zeros(self.n_chains, np.int32), np.zeros(1, np.int32)
        self._ck(self.lib.sb_swap(self.h, int(attempts), int(only_neighbours), int(seed), int(round_), _ptr(sos), _ptr(n)))
        return sos, int(n[0])

    def slots(self) -> np.ndarray:
        s = np.zeros(self.n_chains, np.int32)
        self._ck(self.lib.sb_get_slots(self.h, _ptr(s)))
        return s

    def counts_rebuild(self, chain: int, mma: bool = False):
        p = self.prob
        counts, ms = np.empty((p.G_tot, p.F, p.S), np.int32), np.zeros(1, np.float32)
        fn = self.lib.sb_counts_rebuild_mma if mma else self.lib.sb_counts_rebuild_hist
        self._ck(fn(self.h, chain, _ptr(counts), _ptr(ms)))
        return counts, float(ms[0])
