"""BASELINE north_star: "a config YAML runs unchanged and only the backend differs".

`sbayes.cli.main` (sbayes/cli.py:60-115) -> `run_experiment` -> `MCMCSetup.sample` /
`sample_mc3` (sbayes/mcmc_setup.py:80-142,247-344) on the reference's OWN test config
(test/test_files/config.yaml, copied byte for byte together with its two CSV files into a
scratch directory), K=2, once without and once with MC3 -- with the device drop-in installed
(`sbayes_b200.chain.install()`).  The reference's loggers must produce `stats_*.txt`,
`clusters_*.txt` and `operator_stats_*.txt` with the configured number of samples and sane
contents.  Needs the reference (`ref`): /root/reference in the build container (device code in
the fibre emulator), the copy of oracle/make_ref.py on the B200 box (CUDA library, `-m gpu`).
"""
import filecmp
import json
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from oracle import refharness as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.ref


def _run(tmp_path, backend, mc3, steps, samples):
    src = os.path.join(H.REFERENCE_ROOT, "test", "test_files")
    for f in ("config.yaml", "features.csv", "feature_states_expected.csv"):
        shutil.copy(os.path.join(src, f), tmp_path / f)
        assert filecmp.cmp(os.path.join(src, f), tmp_path / f, shallow=False)       # the YAML runs UNCHANGED
    # settings a user passes on top of the file (the `custom_settings` argument of sbayes.cli.main):
    # PyTables is not installed here (likelihood .h5 logger off); fewer steps than the file's 10 000
    custom = {"results": {"log_likelihood": False}, "mcmc": {"steps": steps, "samples": samples,
                                                             "warmup": {"warmup_steps": 60, "warmup_chains": 2}}}
    if mc3:
        custom["mcmc"]["mc3"] = {"activate": True, "chains": 3, "swap_interval": steps // 4, "temperature_diff": 0.2}
    cmd = [sys.executable, os.path.join(ROOT, "tests", "yaml_e2e_runner.py"), str(tmp_path / "config.yaml"), "e2e", backend, "2",
           json.dumps(custom)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=str(tmp_path))
    assert r.returncode == 0 and "E2E_DONE" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    out = tmp_path / "results" / "e2e" / "K2"
    assert out.is_dir(), list((tmp_path / "results").rglob("*"))
    return out


def _check_outputs(out, samples, steps, mc3):
    stats = out / "stats_K2_0.txt"
    clusters = out / "clusters_K2_0.txt"
    opstats = out / "operator_stats_K2_0.txt"
    assert stats.exists() and clusters.exists() and opstats.exists(), sorted(p.name for p in out.iterdir())
    rows = [ln.rstrip("\n").split("\t") for ln in open(stats)]
    head, body = rows[0], rows[1:]
    assert len(body) == samples, (len(body), samples)
    col = {h: i for i, h in enumerate(head)}
    v = np.array([[float(x) for x in r] for r in body])
    # posterior = likelihood + prior; prior = sum of its four parts (ParametersCSVLogger, loggers.py:176-262)
    assert np.allclose(v[:, col["posterior"]], v[:, col["likelihood"]] + v[:, col["prior"]], atol=1e-5)
    parts = v[:, col["cluster_size_prior"]] + v[:, col["geo_prior"]] + v[:, col["source_prior"]] + v[:, col["weights_prior"]]
    assert np.allclose(v[:, col["prior"]], parts, atol=1e-5)
    assert np.all(np.isfinite(v)) and np.all(v[:, col["likelihood"]] < 0)
    step_ids = v[:, col["Sample"]].astype(int)
    assert np.all(np.diff(step_ids) == steps // samples)
    for k in range(2):
        assert np.all(v[:, col[f"size_a{k}"]] >= 1)
    # weights rows are normalised
    for f in ("F1", "F2"):
        w = v[:, col[f"w_areal_{f}"]] + v[:, col[f"w_universal_{f}"]] + v[:, col[f"w_family_{f}"]]
        assert np.allclose(w, 1.0, atol=1e-5)
    lines = [ln.strip() for ln in open(clusters) if ln.strip()]
    assert len(lines) == samples
    for ln in lines:
        bits = ln.split("\t")
        assert len(bits) == 2 and all(len(b) == 5 and set(b) <= {"0", "1"} for b in bits)
        assert not any(a == "1" and b == "1" for a, b in zip(*bits)), "clusters must stay disjoint"
    # operator_stats: the reference's operator objects received the device's accept / reject counts
    tot = 0
    for ln in open(opstats).read().splitlines()[1:]:
        f = ln.split()
        if len(f) > 3 and f[3].isdigit():
            assert int(f[1]) + int(f[2]) == int(f[3])
            tot += int(f[3])
    # (under MC3 the warm-up attempts run on the same chain object, mcmc_setup.py:503-523: 2 x 60 more steps)
    warm = 120 if mc3 else 0
    assert steps - steps // samples + warm <= tot <= steps + 1 + warm, tot
    if mc3:
        assert (out / "stats_K2_0_chain1.txt").exists() or any("chain" in p.name for p in out.iterdir()), \
            "hot chains are logged too (results.log_hot_chains)"


@pytest.mark.parametrize("backend", ["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
@pytest.mark.parametrize("mc3", [False, True])
def test_cli_main_on_reference_test_config(tmp_path, backend, mc3):
    steps, samples = (400, 20) if backend == "emu" else (4000, 100)
    out = _run(tmp_path, backend, mc3, steps, samples)
    _check_outputs(out, samples, steps, mc3)
