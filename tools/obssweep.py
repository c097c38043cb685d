"""Observation-likelihood kernel time against the feature-tile width (tuning helper)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root (bench.py, sbayes_b200/, oracle/)
import os, subprocess, sys
for ft in (16, 32, 64, 128):
    env = dict(os.environ, SB_OBS_FT=str(ft))
    out = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "scale.py"), sys.argv[1], "--no-steps", "--no-mma"], env=env, capture_output=True, text=True).stdout
    for l in out.splitlines():
        if "obs_loglik" in l or "gibbs_sweep" in l:
            print(f"ft {ft:3d}  {l}", flush=True)
