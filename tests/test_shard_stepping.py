"""Object-sharded MCMC stepping: ONE chain advanced by several ranks, every rank
processing its block of the objects; team barriers, reductions and the bulk statistics cross ranks inside the
step kernel through peer memory (no host collective per step).

Here: `world` ranks inside one process in the fibre emulator (all ranks' CTAs in one fibre pool, peer memory =
host memory).  The sharded chain must be IDENTICAL, decision by decision and bit by bit, to the same chain
advanced by one rank with the same total number of CTAs, and in lockstep with the fp64 oracle.  On the B200s:
tests/multigpu_worker.py (one process per GPU, CUDA IPC) and the single-process variant below (`-m gpu`)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
import build as HB  # noqa: E402
import parity  # noqa: E402
from sbayes_b200 import synth  # noqa: E402
from sbayes_b200.backend import DeviceContext, run_group  # noqa: E402
from sbayes_b200.packed import ChainConfig  # noqa: E402
from test_hostsim_parity import _ClusterHeavy  # noqa: E402


def sharded_vs_single(prob, states, cfgs, n_steps, world, team_local, lib=None, devices=None, rounds=2):
    n = len(states)
    ctxs = [DeviceContext(prob, n, device=(devices[r] if devices else 0), team=team_local, _lib=lib) for r in range(world)]
    single = DeviceContext(prob, n, device=(devices[0] if devices else 0), team=team_local * world, _lib=lib)
    try:
        hs = np.stack([c.shard_handle() for c in ctxs])
        for r, c in enumerate(ctxs):
            c.shard_attach(r, world, hs)
        for ch in range(n):
            for r, c in enumerate(ctxs):
                c.configure(ch, cfgs[ch])
                if r == 0:
                    c.set_state(ch, states[ch])
                else:
                    c.mark_active(ch)
            single.configure(ch, cfgs[ch])
            single.set_state(ch, states[ch])
        trajs = parity.oracle_trajectories([(prob, cfgs[ch], states[ch], n_steps * rounds) for ch in range(n)], procs=1)
        recs_s, recs_1 = [], []
        for _ in range(rounds):             # several launches: the flag epochs carry over, the replicas are rebuilt
            recs_s.append(run_group(ctxs, n_steps, records=True))
            recs_1.append(single.run(n_steps, records=True))
        rs, r1 = np.concatenate(recs_s, axis=1), np.concatenate(recs_1, axis=1)
        for f in ("op", "accept", "changed", "n_changed"):
            assert np.array_equal(rs[f], r1[f]), f
        for f in ("log_q", "log_q_back", "ll_new", "prior_new"):
            assert np.array_equal(rs[f], r1[f], equal_nan=True), f"{f}: the sharded chain must be bit-identical to the single-rank chain"
        for ch in range(n):
            a, b = ctxs[0].get_state(ch), single.get_state(ch)
            assert np.array_equal(a.cluster_of, b.cluster_of) and np.array_equal(a.src, b.src)
            assert np.array_equal(a.counts, b.counts) and np.array_equal(a.W, b.W)
            pa, pb = ctxs[0].posterior(ch), single.posterior(ch)
            assert pa == pb
        parity.compare_with_trajectories(ctxs[0], rs, trajs)
    finally:
        for c in ctxs:
            c.close()
        single.close()


@pytest.fixture(scope="module")
def lib():
    return HB.load()


@pytest.fixture
def env():
    saved = dict(os.environ)
    yield os.environ
    os.environ.clear()
    os.environ.update(saved)


@pytest.mark.parametrize("world,team_local,tiled", [(2, 2, False), (3, 1, False), (2, 3, True)])
def test_emulated_sharded_chain(lib, env, world, team_local, tiled):
    env["SB_HOSTSIM_BLOCK"] = "32"
    if tiled:
        env["SB_HOSTSIM_TILED"] = "1"
    else:
        env["SB_HOSTSIM_SMEM"] = str(16 * 1024)      # per-object replicas not in shared memory: the rank-local replicas serve
    prob, _ = synth.make_problem(N=90, F=18, S=4, K=3, n_fam=4, seed=29, max_size=35, jagged_states=True)
    cfgs = [ChainConfig(seed=900), _ClusterHeavy(temperature=1.2, prior_temperature=1.1, seed=901)]
    st0 = [synth.initial_state(prob, seed=35 + i, objects_per_cluster=7) for i in range(2)]
    sharded_vs_single(prob, st0, cfgs, 36, world, team_local, lib=lib)


@pytest.mark.parametrize("world,team_local,block", [(2, 3, 64)])
def test_emulated_sharded_chain_split_reductions(lib, env, world, team_local, block):
    """the reduction path of large teams (the warps of a CTA split the team's slots), forced at a small team size"""
    env["SB_HOSTSIM_BLOCK"] = str(block)
    env["SB_HOSTSIM_SPLIT_MIN"] = "1"
    env["SB_HOSTSIM_SMEM"] = str(16 * 1024)
    prob, _ = synth.make_problem(N=90, F=18, S=4, K=3, n_fam=4, seed=31, max_size=35, jagged_states=True)
    cfgs = [ChainConfig(seed=910), _ClusterHeavy(temperature=1.2, prior_temperature=1.1, seed=911)]
    st0 = [synth.initial_state(prob, seed=45 + i, objects_per_cluster=7) for i in range(2)]
    sharded_vs_single(prob, st0, cfgs, 20, world, team_local, lib=lib)


@pytest.mark.gpu
def test_gpu_sharded_chain_single_process():
    """all GPUs of the box driven by one process (peer access); skipped on a one-GPU box"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    world = 2
    prob, _ = synth.make_problem(N=3000, F=150, S=6, K=6, n_fam=12, seed=5)
    cfgs = [ChainConfig(seed=910)]
    st0 = [synth.initial_state(prob, seed=45)]
    sharded_vs_single(prob, st0, cfgs, 40, world, 8, devices=list(range(world)))
