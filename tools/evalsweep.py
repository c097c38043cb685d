"""Statistics-pass time against the shared-memory budget of the histogram kernel (tuning helper)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root (bench.py, sbayes_b200/, oracle/)
import os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
for kb in (28, 40, 56, 75, 100, 150):
    env = dict(os.environ, SB_EVAL_SMEM_KB=str(kb), SB_FINE_SMEM_KB=str(kb))      # two-increment pass / fine-row pass
    out = subprocess.run([sys.executable, os.path.join(HERE, "scale.py"), sys.argv[1], "--no-steps", "--no-mma"], env=env, capture_output=True, text=True).stdout
    for l in out.splitlines():
        if "eval_batch:" in l:
            print(f"smem {kb:3d} KB  {l}", flush=True)
