"""Object-sharded full evaluation of BASELINE config 5 (200 000 x 1000 x 8, K=16, 64 families)
under torchrun: every rank generates and holds N/world objects, evaluate() = statistics pass +
int32 all-reduce + terms + fp64 all-reduce.  usage: torchrun --nproc-per-node W tools_shard_bench.py [N_total]"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root (bench.py, sbayes_b200/, oracle/)
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
from sbayes_b200 import synth
from sbayes_b200.packed import ChainConfig
from sbayes_b200.shard import ShardedEvaluator, shard_bounds

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
b = shard_bounds(N, world)
n_loc = b[rank + 1] - b[rank]
prob, _ = synth.make_problem(N=n_loc, F=1000, S=8, K=16, n_fam=64, seed=5 + rank, with_adjacency=False)
st = synth.initial_state(prob, seed=100 + rank)
ev = ShardedEvaluator.from_shard(prob, N, 1, rank, world, device=local)
ev.configure(0, ChainConfig(seed=0))
ev.set_local_state(0, st)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    res = ev.evaluate()
ts, stats = [], []
for _ in range(10):
    flush.fill_(1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = ev.evaluate()
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
t = torch.tensor([float(np.median(ts))], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
stats_ms = ev.ctx.eval_stats()
if rank == 0:
    n_counts = ev._lay[0]
    print(json.dumps({"workload": f"c5 {N}x1000x8, K=16, 64 families, objects sharded over {world} GPU(s)", "n_gpus": world,
                      "objects_per_gpu": n_loc, "evaluate_ms": float(t.item()) * 1e3, "evals_per_sec": 1.0 / float(t.item()),
                      "stats_pass_ms_rank0": stats_ms, "allreduce_int32_bytes": 4 * (n_counts + ev._lay[2]),
                      "algorithmic_GBps": 2.0 * N * 1000 / float(t.item()) / 1e9, "ll": res[0]["ll"], "prior": res[0]["prior"]}))
ev.close()
if world > 1:
    dist.destroy_process_group()
