#!/usr/bin/env python
"""Golden vectors of the observation likelihood (what `LikelihoodLogger._write_sample` appends,
sbayes/sampling/loggers.py:354-359) from the LIVE reference.  Build container only:

    python tests/golden/make_obs_golden.py

For each case a state is advanced by the fp64 oracle for some MCMC steps (so the counts are
not the initial ones), converted to a reference `Sample`, and the reference computes
`sum_c update_weights(sample) * likelihood_per_component_exact(model, sample)`.  The script fails
unless oracle/model.py:observation_likelihood agrees with it to float32 accuracy (the reference's
`normalize` rounds the effect tables and the weights to float32)."""
from __future__ import annotations

import os
import sys
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refharness as H, model as M, chain as CH, fixtures as FX  # noqa: E402
from oracle.draws import PhiloxDraws  # noqa: E402
from sbayes_b200 import synth  # noqa: E402
from sbayes_b200.packed import ChainConfig  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {
    "obs_lik_small": dict(N=60, F=10, S=3, K=2, n_fam=3, seed=11, max_size=25, steps=150),
    "obs_lik_jagged": dict(N=150, F=33, S=5, K=3, n_fam=6, seed=12, max_size=50, jagged_states=True, steps=300),
}


def main():
    r = H.ref()
    from sbayes.sampling.conditionals import likelihood_per_component_exact
    from sbayes.model.likelihood import update_weights
    for name, kw in CASES.items():
        kw = dict(kw)
        steps = kw.pop("steps")
        prob0, truth = synth.make_problem(**kw)
        confs = OrderedDict()
        confs["universal"] = np.ones((1, prob0.N), bool)
        confs["family"] = np.stack([prob0.conf_gid[1] == g for g in range(prob0.conf_ngroups[1])])
        prior_cfg = dict(
            objects_per_cluster=dict(type="uniform_area", min=prob0.min_size, max=prob0.max_size),
            geo=dict(type="uniform"), weights=dict(type="uniform"), cluster_effect=dict(type="uniform"),
            confounding_effects=dict(universal={"<ALL>": dict(type="uniform")},
                                     family={f"g{i}": dict(type="uniform") for i in range(prob0.conf_ngroups[1])}))
        data, model = H.build_reference_model(prob0.x, prob0.applicable, truth["locations"], confs, prob0.K, prior_cfg)
        prob = H.problem_from_reference(data, model)
        st = synth.initial_state(prob, seed=3, objects_per_cluster=6)
        oc = CH.OracleChain(prob, ChainConfig(seed=5), st, draws=PhiloxDraws(5))
        for _ in range(steps):
            oc.step()
        st = oc.state
        sample = H.sample_from_state(prob, model, data, st)
        w = update_weights(sample, caching=False)
        lpc = likelihood_per_component_exact(model=model, sample=sample)
        ref_lh = np.sum(w * lpc, axis=2)                    # what the logger stores (as float32)
        ora = M.observation_likelihood(prob, st)
        err = np.max(np.abs(ref_lh - ora) / np.maximum(np.abs(ora), 1e-30))
        print(f"[{name}] N={prob.N} F={prob.F} max rel err oracle vs reference {err:.2e}  sum log {np.log(ora).sum():.6f}")
        if not err < 2e-6:
            raise SystemExit(f"[{name}] oracle != reference")
        d = FX.problem_to_dict(prob)
        d.update(FX.state_to_dict(st, "st_"))
        d["ref_lh"] = ref_lh.astype(np.float64)
        d["ref_lh_f32"] = ref_lh.astype(np.float32)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)


if __name__ == "__main__":
    main()
