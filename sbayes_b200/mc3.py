"""Host-side mirror of the reference's chain driver for the device backend.

`DeviceMC3` stands in for `MCMCSetup.sample_mc3` + `MCMCChainProcess` + `swap_chains`
(sbayes/mcmc_setup.py:247-418,428-570): a ladder of tempered chains that live in HBM, are
advanced `swap_interval` steps at a time by one launch of the persistent step kernel, and
exchange temperatures between launches.  With `torch.distributed` initialised the ladder
spans all ranks (one process per GPU): every rank advances its local chains, the ranks
all-gather the per-chain (log-likelihood, log-prior) scalars (16 B per chain) and take the
same swap decisions from a shared Philox stream; an accepted swap exchanges the two
chains' *temperatures*, so no state ever crosses NVLink.
"""
from __future__ import annotations

from dataclasses import replace
from typing import List, Optional

import numpy as np

from .backend import DeviceContext, swap_scalars
from .packed import Problem, ChainState, ChainConfig


def temperature_ladder(n_chains: int, dT: float, dTp: float, exponential: bool):
    """sbayes/mcmc_setup.py:263-269."""
    idx = np.arange(n_chains)
    if exponential:
        return (1.0 + dT) ** idx, (1.0 + dTp) ** idx
    return 1.0 + dT * idx, 1.0 + dTp * idx


class DeviceMC3:
    """A temperature ladder of `n_local * world` chains, `n_local` of them on this GPU."""

    def __init__(self, prob: Problem, n_local: int, base_cfg: ChainConfig, dT: float = 0.05, dTp: float = 0.0,
                 exponential: bool = False, device: int = 0, team: int = 0, swap_seed: int = 0, rank: int = 0,
                 world: int = 1, attempts: int = 100, only_neighbours: bool = False, _lib=None):
        self.prob, self.n_local, self.rank, self.world = prob, n_local, rank, world
        self.n_total = n_local * world
        self.temps, self.prior_temps = temperature_ladder(self.n_total, dT, dTp, exponential)
        self.base_cfg = base_cfg
        self.swap_seed, self.attempts, self.only_neighbours = swap_seed, attempts, only_neighbours
        self.round = 0
        # global chain g = local index * world + rank (initially chain g sits at slot g): every rank holds
        # an interleaved sample of the ladder, so the hot chains -- which accept more and do more work per
        # step -- are spread over the GPUs instead of sitting on the last rank
        self.chain_at_slot = list(range(self.n_total))
        self.ctx = DeviceContext(prob, n_local, device=device, team=team, _lib=_lib)
        self.swap_matrix = np.zeros((self.n_total, self.n_total), np.int64)
        self._apply_temperatures()

    def global_chain(self, local: int) -> int:
        return local * self.world + self.rank

    def _slot_of_chain(self):
        s = [0] * self.n_total
        for slot, ch in enumerate(self.chain_at_slot):
            s[ch] = slot
        return s

    def _apply_temperatures(self, first: bool = True):
        soc = self._slot_of_chain()
        gl = [self.global_chain(l) for l in range(self.n_local)]
        if not first:       # after a swap only the temperatures move: one call, no host synchronisation
            self.ctx.set_temperatures([float(self.temps[soc[g]]) for g in gl], [float(self.prior_temps[soc[g]]) for g in gl])
            return
        for l, g in enumerate(gl):
            cfg = replace(self.base_cfg, temperature=float(self.temps[soc[g]]),
                          prior_temperature=float(self.prior_temps[soc[g]]), seed=self.base_cfg.seed + g)
            self.ctx.configure(l, cfg)

    def set_states(self, states: List[ChainState], i_step: int = 0):
        for l, st in enumerate(states):
            self.ctx.set_state(l, st, i_step)

    def local_scalars(self) -> np.ndarray:
        return np.array([(p["ll"], p["prior"]) for p in self.ctx.posteriors()], dtype=np.float64).reshape(self.n_local, 2)

    def gather_scalars(self, local: np.ndarray) -> np.ndarray:
        if self.world == 1:
            return local
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.from_numpy(np.ascontiguousarray(local)).to(dev).reshape(-1)
        out = torch.empty(self.world * t.numel(), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(out, t)
        # rank-major order of the collective -> global chain order (g = local * world + rank)
        return out.cpu().numpy().reshape(self.world, self.n_local, 2).transpose(1, 0, 2).reshape(self.n_total, 2)

    def round_trip(self, swap_interval: int):
        """One MC3 round: `swap_interval` steps on every local chain, then the swap."""
        self.ctx.run_async(swap_interval)
        self.ctx.sync()
        return self.swap()

    def swap(self):
        allv = self.gather_scalars(self.local_scalars())
        self.round += 1
        # the decision procedure lives once, in the library (sb_swap_scalars); every rank calls it
        # on the same gathered scalars and the same Philox stream -> the same decisions everywhere
        acc = swap_scalars(allv[:, 0], allv[:, 1], self.temps, self.prior_temps, self.chain_at_slot,
                           self.swap_seed, self.round, self.attempts, self.only_neighbours, _lib=self.ctx.lib)
        for i, j in acc:
            self.swap_matrix[i, j] += 1
        if acc:
            self._apply_temperatures(first=False)
        return acc

    def cold_chain(self) -> int:
        """global index of the chain currently at T = 1 (the one the reference logs)."""
        return self.chain_at_slot[0]

    def close(self):
        self.ctx.close()
