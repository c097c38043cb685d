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

import math
from dataclasses import replace
from typing import List, Optional

import numpy as np

from .backend import DeviceContext
from .packed import Problem, ChainState, ChainConfig

SITE_SWAP = 13


def temperature_ladder(n_chains: int, dT: float, dTp: float, exponential: bool):
    """sbayes/mcmc_setup.py:263-269."""
    idx = np.arange(n_chains)
    if exponential:
        return (1.0 + dT) ** idx, (1.0 + dTp) ** idx
    return 1.0 + dT * idx, 1.0 + dTp * idx


# --- Philox4x32-10, the swap stream (same addressing as the device code, sb_core.cuh) -------
def _philox(c0, c1, c2, c3, k0, k1):
    M0, M1, MASK = 0xD2511F53, 0xCD9E8D57, 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + 0x9E3779B9) & MASK, (k1 + 0xBB67AE85) & MASK
    return c0, c1, c2, c3


def _swap_uniform(seed: int, round_: int, idx: int, sub: int) -> float:
    w = _philox(idx, SITE_SWAP | (sub << 8), round_ & 0xFFFFFFFF, (round_ >> 32) & 0xFFFFFFFF,
                seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return ((w[0] >> 5) * 67108864.0 + (w[1] >> 6)) / 9007199254740992.0


def swap_decisions(ll, prior, temps, prior_temps, chain_at_slot: List[int], seed: int, round_: int,
                   attempts: int = 100, only_neighbours: bool = False):
    """swap_chains (sbayes/mcmc_setup.py:365-418) on scalars.  `chain_at_slot[s]` is the chain
    currently running at temperature slot s; accepted swaps are applied to it in place.
    Deterministic in (seed, round_): every rank computes the same result."""
    n = len(chain_at_slot)
    if only_neighbours:
        pairs = [(i, i + 1) for i in range(n - 1)]
    else:
        pairs = [(i, j) for i in range(n - 1) for j in range(i + 1, n)]
    attempts = min(attempts, len(pairs))
    idx = list(range(len(pairs)))
    accepted = []
    for a in range(attempts):
        rem = len(idx) - a
        j = a + min(int(_swap_uniform(seed, round_, a, 0) * rem), rem - 1)
        idx[a], idx[j] = idx[j], idx[a]
        i0, j0 = pairs[idx[a]]
        ci, cj = chain_at_slot[i0], chain_at_slot[j0]
        r = -((prior[ci] - prior[cj]) * (1.0 / prior_temps[i0] - 1.0 / prior_temps[j0])
              + (ll[ci] - ll[cj]) * (1.0 / temps[i0] - 1.0 / temps[j0]))
        u = _swap_uniform(seed, round_, a, 1)
        if (math.log(u) if u > 0 else -math.inf) < r:
            chain_at_slot[i0], chain_at_slot[j0] = cj, ci
            accepted.append((i0, j0))
    return accepted


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
        # global chain g = rank * n_local + local index; initially chain g sits at slot g
        self.chain_at_slot = list(range(self.n_total))
        self.ctx = DeviceContext(prob, n_local, device=device, team=team, _lib=_lib)
        self.swap_matrix = np.zeros((self.n_total, self.n_total), np.int64)
        self._apply_temperatures()

    def _slot_of_chain(self):
        s = [0] * self.n_total
        for slot, ch in enumerate(self.chain_at_slot):
            s[ch] = slot
        return s

    def _apply_temperatures(self):
        soc = self._slot_of_chain()
        for l in range(self.n_local):
            g = self.rank * self.n_local + l
            cfg = replace(self.base_cfg, temperature=float(self.temps[soc[g]]),
                          prior_temperature=float(self.prior_temps[soc[g]]), seed=self.base_cfg.seed + g)
            self.ctx.configure(l, cfg)

    def set_states(self, states: List[ChainState], i_step: int = 0):
        for l, st in enumerate(states):
            self.ctx.set_state(l, st, i_step)

    def local_scalars(self) -> np.ndarray:
        out = np.zeros((self.n_local, 2))
        for l in range(self.n_local):
            p = self.ctx.posterior(l)
            out[l] = (p["ll"], p["prior"])
        return out

    def gather_scalars(self, local: np.ndarray) -> np.ndarray:
        if self.world == 1:
            return local
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.from_numpy(np.ascontiguousarray(local)).to(dev).reshape(-1)
        out = torch.empty(self.world * t.numel(), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(out, t)
        return out.cpu().numpy().reshape(self.n_total, 2)

    def round_trip(self, swap_interval: int):
        """One MC3 round: `swap_interval` steps on every local chain, then the swap."""
        self.ctx.run_async(swap_interval)
        self.ctx.sync()
        return self.swap()

    def swap(self):
        allv = self.gather_scalars(self.local_scalars())
        self.round += 1
        acc = swap_decisions(allv[:, 0], allv[:, 1], self.temps, self.prior_temps, self.chain_at_slot,
                             self.swap_seed, self.round, self.attempts, self.only_neighbours)
        for i, j in acc:
            self.swap_matrix[i, j] += 1
        if acc:
            self._apply_temperatures()
        return acc

    def cold_chain(self) -> int:
        """global index of the chain currently at T = 1 (the one the reference logs)."""
        return self.chain_at_slot[0]

    def close(self):
        self.ctx.close()
