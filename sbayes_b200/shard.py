"""Object-sharded full evaluation (BASELINE config 5: a matrix too large for one pass of one
GPU, objects partitioned across the ranks of one box).

Every sum of the uncached `Model.likelihood + Model.prior` (sbayes/model/model.py:47-51) is a
sum over objects: the count tensors (`recalculate_feature_counts`, sbayes/sampling/counts.py:
10-52), the cluster sizes, and the source prior (`SourcePrior.__call__`, sbayes/model/prior.py:
568-611).  Each rank holds a contiguous block of objects in its own `DeviceContext` and runs the
statistics pass (`sb_eval_stats`) on it; then

  1. ONE all-reduce (sum, int32) of the leading `counts | sizes` part of every chain's statistics block
  2. the terms kernel on every rank (`sb_eval_terms`, identical counts -> identical log-likelihood
     and size prior on every rank; the source prior it returns is the shard's partial sum)
  3. ONE all-reduce (sum, fp64) of the per-chain partial source priors.

The collectives go through `torch.distributed` (NCCL over NVLink on the GPU box; gloo in the CPU
tests).  The int32 statistics are wrapped zero-copy from the device address the C ABI returns
(`sb_stats_buffer`); the (pattern, feature, component) histogram part of the block is indexed by
the *shard's own* has-patterns, so it is excluded from the all-reduce and only its linear
functional (the partial source prior) crosses ranks.
"""
from __future__ import annotations

import ctypes
from dataclasses import replace
from typing import List, Optional, Sequence

import numpy as np

from .backend import DeviceContext
from .packed import Problem, ChainState, ChainConfig


def shard_bounds(n_objects: int, world: int) -> List[int]:
    """Contiguous, balanced object blocks: rank r owns [b[r], b[r+1])."""
    return [(n_objects * r) // world for r in range(world + 1)]


def shard_problem(prob: Problem, lo: int, hi: int) -> Problem:
    """The sub-problem of objects [lo, hi): group tables and priors are global, everything
    object-indexed is sliced.  Neighbourhoods / geo costs are dropped (evaluation only)."""
    if prob.geo is not None:
        raise ValueError("the cost-based geo prior is not supported with sharded objects")
    return replace(prob, x=np.ascontiguousarray(prob.x[lo:hi]), conf_gid=[np.ascontiguousarray(g[lo:hi]) for g in prob.conf_gid],
                   adj_indptr=np.zeros(hi - lo + 1, np.int32), adj_indices=np.zeros(0, np.int32))


def shard_state(st: ChainState, lo: int, hi: int) -> ChainState:
    return ChainState(cluster_of=np.ascontiguousarray(st.cluster_of[lo:hi]), W=st.W.copy(),
                      src=np.ascontiguousarray(st.src[lo:hi]), counts=None)


class _DeviceArray:
    """__cuda_array_interface__ view of device memory owned by the ctx (zero-copy torch tensor)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}


class ShardedEvaluator:
    """`n_chains` states of ONE problem whose objects are split over `world` ranks."""

    def __init__(self, prob: Problem, n_chains: int, rank: int = 0, world: int = 1, device: int = 0, _lib=None,
                 stats_mode: str = "hist"):
        self.N = prob.N
        self.rank, self.world, self.n_chains = rank, world, n_chains
        b = shard_bounds(prob.N, world)
        self.lo, self.hi = b[rank], b[rank + 1]
        self.local = shard_problem(prob, self.lo, self.hi) if world > 1 else prob
        self.ctx = DeviceContext(self.local, n_chains, device=device, _lib=_lib)
        self._host = _lib is not None
        if stats_mode != "hist":
            self.ctx.eval_set_mode(stats_mode)
        ptr, n, stride, lay = self.ctx.stats_buffer()
        self._stride, self._lay = stride, lay
        if self._host:      # emulator build: "device" memory is host memory
            buf = (ctypes.c_int32 * n).from_address(ptr)
            self._stats_np = np.frombuffer(buf, dtype=np.int32)
            self._stats = None
        else:
            import torch
            self._stats = torch.as_tensor(_DeviceArray(ptr, n), device=torch.device("cuda", device))
            self._stats_np = None

    @classmethod
    def from_shard(cls, local_prob: Problem, n_global: int, n_chains: int, rank: int, world: int, device: int = 0, _lib=None,
                   stats_mode: str = "hist") -> "ShardedEvaluator":
        """Build from a rank's own block of objects (e.g. loaded or generated per rank: the whole
        matrix never has to exist in one process); states are then set with `set_local_state`."""
        self = cls.__new__(cls)
        self.N, self.rank, self.world, self.n_chains = int(n_global), rank, world, n_chains
        self.lo, self.hi = 0, local_prob.N
        self.local = local_prob
        self.ctx = DeviceContext(local_prob, n_chains, device=device, _lib=_lib)
        self._host = _lib is not None
        if stats_mode != "hist":
            self.ctx.eval_set_mode(stats_mode)
        ptr, n, stride, lay = self.ctx.stats_buffer()
        self._stride, self._lay = stride, lay
        if self._host:
            self._stats_np = np.frombuffer((ctypes.c_int32 * n).from_address(ptr), dtype=np.int32)
            self._stats = None
        else:
            import torch
            self._stats = torch.as_tensor(_DeviceArray(ptr, n), device=torch.device("cuda", device))
            self._stats_np = None
        return self

    def set_local_state(self, chain: int, st: ChainState):
        """`st` covers only this rank's block of objects."""
        self.ctx.set_state(chain, st)

    def configure(self, chain: int, cfg: ChainConfig):
        self.ctx.configure(chain, cfg)

    def set_state(self, chain: int, st: ChainState):
        """`st` is the state of the WHOLE problem; the rank keeps its block of objects."""
        self.ctx.set_state(chain, shard_state(st, self.lo, self.hi) if self.world > 1 else st)

    def _allreduce_stats(self):
        if self.world == 1:
            return
        import torch
        import torch.distributed as dist
        n_counts, n_hist, n_sizes = self._lay
        if self._host:
            t = torch.from_numpy(self._stats_np).view(self.n_chains, self._stride)
        else:
            t = self._stats.view(self.n_chains, self._stride)
        # counts | sizes lead every state's block and are reduced in place; the histogram that
        # follows stays local (shard-specific pattern order)
        n_red = n_counts + n_sizes
        if self.n_chains == 1:
            dist.all_reduce(t[0, :n_red], op=dist.ReduceOp.SUM)
        else:
            packed = t[:, :n_red].contiguous()
            dist.all_reduce(packed, op=dist.ReduceOp.SUM)
            t[:, :n_red] = packed
        if not self._host:      # the ctx launches on its own stream: the collective must have landed
            torch.cuda.current_stream().synchronize()

    def evaluate(self) -> List[dict]:
        """Uncached log-likelihood and prior terms of every chain; identical on every rank."""
        self.ctx.eval_stats()
        self._allreduce_stats()
        res = self.ctx.eval_terms(n_global=self.N)
        if self.world > 1:
            import torch
            import torch.distributed as dist
            dev = "cpu" if self._host else self._stats.device
            sp = torch.tensor([r["source_prior"] for r in res], dtype=torch.float64, device=dev)
            dist.all_reduce(sp, op=dist.ReduceOp.SUM)
            sp = sp.cpu().numpy()
            for r, v in zip(res, sp):
                r["source_prior"] = float(v)
                r["prior"] = r["size_prior"] + r["geo_prior"] + r["weights_prior"] + r["source_prior"]
        return res

    def counts(self, chain: int) -> np.ndarray:
        """the (all-reduced) count tensor of a chain, [G_tot, F, S]"""
        p = self.local
        n_counts = self._lay[0]
        if self._host:
            flat = self._stats_np[chain * self._stride: chain * self._stride + n_counts].copy()
        else:
            flat = self._stats[chain * self._stride: chain * self._stride + n_counts].cpu().numpy()
        return flat.reshape(p.G_tot, p.F, p.S)

    def close(self):
        self.ctx.close()


class ShardedChains:
    """`n_chains` chains whose OBJECTS are sharded over the `world` GPUs of one box: every MCMC step is taken
    by all GPUs together, each walking its block of the objects in the proposal passes, with the team barriers,
    the reductions and the bulk statistics crossing GPUs inside the persistent step kernel through NVLink peer
    memory (include/sbayes_b200.h: sb_shard_attach) -- no host collective per step.

    One process per GPU: `torch.distributed` is used once, to exchange the 128-byte arena handles, and as a host
    barrier around state uploads.  Every rank constructs this object with the SAME whole problem (the packed
    observations are replicated: 200 MB at 200 000 x 1 000); the chains' state lives on rank 0.  `run` is
    collective.  Replaces, for matrices too large for one GPU's pass, the per-step sums over objects of
    sbayes/sampling/operators.py:1035-1073 (candidate posterior), :496-552 (source Gibbs) and
    sbayes/model/prior.py:596-611 (source prior)."""

    def __init__(self, prob: Problem, n_chains: int, rank: int = 0, world: int = 1, device: int = 0, team: int = 0, _lib=None):
        self.rank, self.world, self.n_chains = rank, world, n_chains
        self.ctx = DeviceContext(prob, n_chains, device=device, team=team, _lib=_lib)
        if world > 1:
            import torch
            import torch.distributed as dist
            h = self.ctx.shard_handle()
            dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
            mine = torch.from_numpy(h.copy()).to(dev)
            out = torch.empty(world * 128, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(out, mine)
            self.ctx.shard_attach(rank, world, out.cpu().numpy())
            dist.barrier()

    def _barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()

    def configure(self, chain: int, cfg: ChainConfig):
        self.ctx.configure(chain, cfg)          # the same configuration on every rank

    def set_state(self, chain: int, st: Optional[ChainState], i_step: int = 0):
        """collective; `st` is needed on rank 0 only"""
        if self.rank == 0:
            self.ctx.set_state(chain, st, i_step)
            self.ctx.sync()
        else:
            self.ctx.mark_active(chain)
        self._barrier()

    def run(self, n_steps: int, records: bool = False):
        """collective: the kernels of all ranks meet on the device"""
        self._barrier()
        rec = self.ctx.run(n_steps, records=records)
        self._barrier()
        return rec if self.rank == 0 else None

    def run_async(self, n_steps: int):
        self.ctx.run_async(n_steps)

    def sync(self):
        self.ctx.sync()

    def get_state(self, chain: int) -> Optional[ChainState]:
        return self.ctx.get_state(chain) if self.rank == 0 else None

    def posterior(self, chain: int) -> Optional[dict]:
        return self.ctx.posterior(chain) if self.rank == 0 else None

    def close(self):
        self._barrier()
        self.ctx.close()
