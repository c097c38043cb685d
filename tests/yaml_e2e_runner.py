#!/usr/bin/env python
"""Runs the reference's command-line entry point `sbayes.cli.main` (sbayes/cli.py:60-115) on a
config YAML, with the MCMC backend either the reference's own or the device drop-in
(`sbayes_b200.chain.install()`).  TEST INFRASTRUCTURE: executed as a subprocess by
tests/test_yaml_e2e.py so that the MC3 child processes fork from a parent that never
initialised CUDA.

    python tests/yaml_e2e_runner.py <config.yaml> <name> <backend: reference|cuda|emu> <K> <json custom settings>
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    config, name, backend, K, custom = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), json.loads(sys.argv[5])
    from oracle import refharness as H
    H.ref()
    H.patch_network_for_stubs()
    H.seed_reference(int(os.environ.get("SB_E2E_SEED", "5")))
    if backend != "reference":
        from sbayes_b200 import chain as C
        lib = None
        if backend == "emu":
            sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
            import build as HB
            lib = HB.load()
        C.install(seed=int(os.environ.get("SB_E2E_SEED", "5")), _lib=lib)
    from sbayes.cli import main as sbayes_main
    sbayes_main(config=config, experiment_name=name, custom_settings=custom, n_clusters=K)
    print("E2E_DONE")


if __name__ == "__main__":
    main()
