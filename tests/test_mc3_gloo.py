"""N>1 path on CPU: two processes (gloo, world_size 2), each advancing its half of an MC3
temperature ladder in the fibre-emulator build of the device code, all-gathering the swap
scalars and taking the same swap decisions.  The result must equal one process holding the
whole ladder."""
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
    from sbayes_b200.packed import ChainConfig
    prob, _ = synth.make_problem(N=40, F=8, S=3, K=2, n_fam=3, seed=5, max_size=15)
    states = [synth.initial_state(prob, seed=10 + c, objects_per_cluster=5) for c in range(4)]
    return HB.load(), prob, states, ChainConfig(seed=900)


def _run_ladder(lib, prob, states, cfg, rank, world, rounds=4, interval=25):
    from sbayes_b200.mc3 import DeviceMC3
    n_local = 4 // world
    mc3 = DeviceMC3(prob, n_local, cfg, dT=0.4, dTp=0.2, swap_seed=17, rank=rank, world=world, _lib=lib)
    mc3.set_states([states[mc3.global_chain(l)] for l in range(n_local)])      # chain g = local * world + rank
    log = []
    for _ in range(rounds):
        log.append(mc3.round_trip(interval))
    fin = [mc3.ctx.get_state(c) for c in range(n_local)]
    post = mc3.local_scalars()
    out = dict(acc=log, slots=list(mc3.chain_at_slot), swaps=mc3.swap_matrix.copy(),
               cluster_of=[f.cluster_of for f in fin], src=[f.src for f in fin], post=post)
    mc3.close()
    return out


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib, prob, states, cfg = _setup()
        q.put((rank, _run_ladder(lib, prob, states, cfg, rank, world)))
    except Exception as e:      # report instead of leaving the parent waiting
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_ladder_equals_single_process():
    import torch.multiprocessing as mp
    lib, prob, states, cfg = _setup()
    single = _run_ladder(lib, prob, states, cfg, 0, 1)
    assert sum(len(a) for a in single["acc"]) > 0, "test needs at least one accepted swap to be meaningful"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    assert all(isinstance(v, dict) for v in res.values()), res
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for r in range(2):
        assert res[r]["acc"] == single["acc"] and res[r]["slots"] == single["slots"]
        assert np.array_equal(res[r]["swaps"], single["swaps"])
        for l in range(2):
            g = l * 2 + r
            assert np.array_equal(res[r]["cluster_of"][l], single["cluster_of"][g])
            assert np.array_equal(res[r]["src"][l], single["src"][g])
            assert np.array_equal(res[r]["post"][l], single["post"][g])
