"""Object sharding on CPU: two processes (gloo, world_size 2) each hold half of the objects in
the fibre-emulator build of the device code, all-reduce the int32 statistics and the partial
source priors, and must reproduce the single-process evaluation and the fp64 oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
    import build as HB
    from sbayes_b200 import synth
    from sbayes_b200.packed import SIZE_UNIFORM_SIZE
    prob, _ = synth.make_problem(N=90, F=21, S=4, K=3, n_fam=4, seed=5, max_size=25, jagged_states=True)
    prob.size_prior = SIZE_UNIFORM_SIZE          # the size prior needs the GLOBAL number of objects
    prob.weights_prior_uniform = False
    prob.alpha_w = 1.0 + np.random.default_rng(1).random((prob.F, prob.C))
    # make the two halves see different confounder patterns first (shard-local pattern order)
    states = [synth.initial_state(prob, seed=10 + c, objects_per_cluster=6) for c in range(3)]
    return HB.load(), prob, states


def _evaluate(lib, prob, states, rank, world):
    from sbayes_b200.shard import ShardedEvaluator
    from sbayes_b200.packed import ChainConfig
    ev = ShardedEvaluator(prob, len(states), rank=rank, world=world, _lib=lib)
    for c, st in enumerate(states):
        ev.configure(c, ChainConfig(seed=c))
        ev.set_state(c, st)
    res = ev.evaluate()
    counts = [ev.counts(c) for c in range(len(states))]
    ev.close()
    return res, counts


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib, prob, states = _setup()
        q.put((rank, _evaluate(lib, prob, states, rank, world)))
    except Exception as e:
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    sys.path.insert(0, ROOT)
    from sbayes_b200.shard import shard_bounds
    for n, w in [(5, 2), (200000, 8), (7, 8), (90, 3)]:
        b = shard_bounds(n, w)
        assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(w))
        assert max(b[i + 1] - b[i] for i in range(w)) - min(b[i + 1] - b[i] for i in range(w)) <= 1


@pytest.mark.timeout(600)
def test_two_rank_sharded_eval_equals_single_process_and_oracle():
    import torch.multiprocessing as mp
    from oracle import model as M
    lib, prob, states = _setup()
    single, counts1 = _evaluate(lib, prob, states, 0, 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    assert all(isinstance(v, tuple) for v in res.values()), res
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for c, st in enumerate(states):
        ref = st.copy()
        ref.counts = M.rebuild_counts(prob, st)
        ll, pr = M.log_likelihood(prob, ref), M.log_prior(prob, ref)
        for r in range(2):
            got, counts = res[r]
            assert np.array_equal(counts[c], ref.counts), "all-reduced counts must be bit-exact"
            assert np.array_equal(counts[c], counts1[c])
            assert got[c]["ll"] == single[c]["ll"], "identical integer statistics -> identical fp64 log-likelihood"
            assert got[c]["size_prior"] == single[c]["size_prior"] and got[c]["weights_prior"] == single[c]["weights_prior"]
            assert abs(got[c]["source_prior"] - single[c]["source_prior"]) <= 1e-12 * abs(single[c]["source_prior"])
            assert abs(got[c]["ll"] - ll) <= 1e-10 * abs(ll) and abs(got[c]["prior"] - pr) <= 1e-10 * abs(pr)
        assert res[0][0][c] == res[1][0][c], "every rank must hold the same posterior"
