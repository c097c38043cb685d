"""Per-phase cycle breakdown of the step kernel with the objects of one chain sharded over the ranks (torchrun;
needs a library built with SB_CFLAGS=-DSB_PHASE_PROF, SBAYES_B200_LIB pointing at it).
usage: torchrun --nproc-per-node W tools/phases_sharded.py <objects per GPU> <steps>"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root (bench.py, sbayes_b200/, oracle/)
import ctypes as C, os, sys
import numpy as np
import torch
import torch.distributed as dist
import bench
from sbayes_b200 import synth
from sbayes_b200.packed import ChainConfig, OP_NAMES
from sbayes_b200.shard import ShardedChains
from tools.phases import NAMES, NPH

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = dict(bench.CONFIGS["c5"], N=int(sys.argv[1]))
nsteps = int(sys.argv[2])
prob = bench.assemble_sharded_problem(w, rank, world, local)
sc = ShardedChains(prob, 1, rank=rank, world=world, device=local)
sc.configure(0, ChainConfig(seed=77))
sc.set_state(0, synth.initial_state(prob, seed=1000) if rank == 0 else None)
sc.run(30)
ph, n = np.zeros((10, NPH), np.int64), np.zeros(10, np.int64)
lib = sc.ctx.lib
lib.sb_debug_phases.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
lib.sb_debug_phases(sc.ctx.h, ph.ctypes.data, n.ctypes.data, 1)
sc.run(nsteps)
lib.sb_debug_phases(sc.ctx.h, ph.ctypes.data, n.ctypes.data, 0)
if rank == 0:
    print(f"world={world} objects={prob.N} info={sc.ctx.info()}")
    for op in range(10):
        if not n[op]:
            continue
        print(f"{OP_NAMES[op]:34s} n={n[op]:6d}  {ph[op].sum() / n[op] / 1965:8.1f} us/step")
        for i in range(NPH):
            if ph[op, i]:
                print(f"      {NAMES.get(i, str(i)):22s} {ph[op, i] / n[op] / 1965:8.2f} us")
sc.close()
if world > 1:
    dist.destroy_process_group()
