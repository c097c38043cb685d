"""Worker of tests/test_multi_gpu.py: run under torchrun with one rank per GPU (NCCL).
(a) object-sharded evaluation == single-GPU evaluation of the whole problem == fp64 oracle;
(b) an MC3 ladder split over the ranks == the same ladder on one GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import model as M  # noqa: E402
from sbayes_b200 import synth  # noqa: E402
from sbayes_b200.backend import DeviceContext  # noqa: E402
from sbayes_b200.mc3 import DeviceMC3  # noqa: E402
from sbayes_b200.packed import ChainConfig, SIZE_UNIFORM_SIZE  # noqa: E402
from sbayes_b200.shard import ShardedEvaluator  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # ---- (a) sharded evaluation -------------------------------------------------------------
    prob, _ = synth.make_problem(N=3001, F=150, S=8, K=16, n_fam=64, seed=5)
    prob.size_prior = SIZE_UNIFORM_SIZE
    states = [synth.initial_state(prob, seed=3 + c) for c in range(2)]
    for mode in ("hist", "mma"):
        ev = ShardedEvaluator(prob, 2, rank=rank, world=world, device=local, stats_mode=mode)
        for c, st in enumerate(states):
            ev.configure(c, ChainConfig(seed=c))
            ev.set_state(c, st)
        res = ev.evaluate()
        counts = [ev.counts(c) for c in range(2)]
        ev.close()
        ctx = DeviceContext(prob, 2, device=local)
        for c, st in enumerate(states):
            ctx.configure(c, ChainConfig(seed=c))
            ctx.set_state(c, st)
        whole = ctx.eval_batch()
        for c, st in enumerate(states):
            ref = st.copy()
            ref.counts = M.rebuild_counts(prob, st)
            assert np.array_equal(counts[c], ref.counts), "all-reduced counts"
            assert res[c]["ll"] == whole[c]["ll"] and res[c]["size_prior"] == whole[c]["size_prior"]
            assert abs(res[c]["source_prior"] - whole[c]["source_prior"]) <= 1e-12 * abs(whole[c]["source_prior"])
            ll, pr = M.log_likelihood(prob, ref), M.log_prior(prob, ref)
            assert abs(res[c]["ll"] - ll) <= 1e-10 * abs(ll) and abs(res[c]["prior"] - pr) <= 1e-10 * abs(pr)
        ctx.close()
    # ---- (b) MC3 ladder over the ranks --------------------------------------------------------
    n_total = 2 * world
    prob, _ = synth.make_problem(N=200, F=20, S=4, K=3, n_fam=4, seed=9, max_size=60)
    sts = [synth.initial_state(prob, seed=20 + c, objects_per_cluster=8) for c in range(n_total)]

    def ladder(r, w):
        n_local = n_total // w
        mc3 = DeviceMC3(prob, n_local, ChainConfig(seed=500), dT=0.3, dTp=0.1, swap_seed=11, rank=r, world=w, device=local)
        mc3.set_states([sts[mc3.global_chain(l)] for l in range(n_local)])      # chain g = local * world + rank
        acc = [mc3.round_trip(40) for _ in range(5)]
        post = mc3.local_scalars()
        fin = [mc3.ctx.get_state(c) for c in range(n_local)]
        mc3.close()
        return acc, post, fin
    acc, post, fin = ladder(rank, world)
    acc1, post1, fin1 = ladder(0, 1)
    assert acc == acc1, "swap decisions differ between the split and the single-GPU ladder"
    for l in range(n_total // world):
        g = l * world + rank
        assert np.array_equal(post[l], post1[g])
        assert np.array_equal(fin[l].cluster_of, fin1[g].cluster_of) and np.array_equal(fin[l].src, fin1[g].src)
    dist.barrier()
    # ---- (c) ONE chain advanced by all ranks, objects sharded (in-kernel exchange over peer memory) -----
    from sbayes_b200.shard import ShardedChains
    import parity
    prob, _ = synth.make_problem(N=3000, F=150, S=6, K=6, n_fam=12, seed=5)
    cfgs = [ChainConfig(seed=910), ChainConfig(temperature=1.2, prior_temperature=1.1, seed=911)]
    st0 = [synth.initial_state(prob, seed=45 + c) for c in range(2)]
    # (at eight ranks the team has 80 CTAs: above the size where the warps of a CTA split the reduction slots)
    team_local, n_steps = (8 if world < 8 else 10), 60
    sc = ShardedChains(prob, 2, rank=rank, world=world, device=local, team=team_local)
    for ch in range(2):
        sc.configure(ch, cfgs[ch])
        sc.set_state(ch, st0[ch] if rank == 0 else None)
    recs = [sc.run(n_steps, records=True) for _ in range(2)]
    if rank == 0:
        rs = np.concatenate(recs, axis=1)
        single = DeviceContext(prob, 2, device=local, team=team_local * world)
        for ch in range(2):
            single.configure(ch, cfgs[ch]); single.set_state(ch, st0[ch])
        r1 = np.concatenate([single.run(n_steps, records=True) for _ in range(2)], axis=1)
        for f in ("op", "accept", "changed", "n_changed"):
            assert np.array_equal(rs[f], r1[f]), f
        for f in ("log_q", "log_q_back", "ll_new", "prior_new"):
            assert np.array_equal(rs[f], r1[f], equal_nan=True), f"{f}: sharded chain != single-GPU chain"
        for ch in range(2):
            a, b = sc.get_state(ch), single.get_state(ch)
            assert np.array_equal(a.cluster_of, b.cluster_of) and np.array_equal(a.src, b.src) and np.array_equal(a.counts, b.counts)
            assert sc.posterior(ch) == single.posterior(ch)
        trajs = parity.oracle_trajectories([(prob, cfgs[ch], st0[ch], 2 * n_steps) for ch in range(2)], procs=2)
        parity.compare_with_trajectories(sc.ctx, rs, trajs)
        single.close()
        print(f"SHARDED STEPPING OK world={world}: {2 * n_steps} steps x 2 chains identical to the single-GPU chain and the oracle")
    sc.close()
    dist.barrier()
    if rank == 0:
        print(f"MULTIGPU OK world={world} swaps={sum(len(a) for a in acc)}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
