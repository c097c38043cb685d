"""Fine-row statistics pass: time against tile width (budget), block size and object blocks (tuning helper).
usage: python tools/finesweep.py c4,c5s"""
import os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
for kb, blk, ob in ((32, 256, 0), (60, 256, 0), (32, 512, 0)):
    env = dict(os.environ, SB_FINE_SMEM_KB=str(kb), SB_FINE_BLOCK=str(blk))
    if ob:
        env["SB_FINE_OBLK"] = str(ob)
    out = subprocess.run([sys.executable, os.path.join(HERE, "scale.py"), sys.argv[1], "--no-steps", "--no-mma"], env=env, capture_output=True, text=True).stdout
    for l in out.splitlines():
        if "statistics pass:" in l:
            info = l.split("statistics pass:")[1].strip()
        if "eval_batch:" in l:
            print(f"budget {kb:3d} KB block {blk:4d} oblk {ob:2d}  {l.split('terms')[0]}  {info}", flush=True)
