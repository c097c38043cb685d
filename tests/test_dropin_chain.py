"""Boundary B1: `sbayes_b200.chain.MCMCChain` behind the reference's own API, driven with the
reference's `Data` / `Model` / `Sample` / loggers.  Needs the live reference (`-m ref`, build
container only); the device code runs in the fibre emulator here and on the B200 in the
`gpu`-marked variant."""
import os
import sys
from collections import OrderedDict

import numpy as np
import pytest

from oracle import refharness as H, chain as CH, model as M
from oracle.draws import PhiloxDraws
from sbayes_b200 import synth
from sbayes_b200.packed import ChainConfig

pytestmark = pytest.mark.ref


def _lib():
    import torch
    if torch.cuda.is_available():
        return None
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    import build as HB
    return HB.load()


class CountingLogger:
    """Stands in for a ResultsLogger: reads what ParametersCSVLogger reads (loggers.py:157-262)."""

    def __init__(self):
        self.rows = []

    def write_sample(self, sample):
        self.rows.append((sample.i_step, sample.last_lh, sample.last_prior,
                          float(sample.cache.cluster_size_prior.value), float(np.sum(sample.cache.source_prior.value)),
                          sample.clusters.value.copy(), sample.weights.value.copy()))

    def close(self):
        pass


def _reference_problem():
    prob, truth = synth.make_problem(N=60, F=10, S=3, K=2, n_fam=3, seed=11, max_size=25)
    confs = OrderedDict()
    confs["universal"] = np.ones((1, prob.N), bool)
    confs["family"] = np.stack([prob.conf_gid[1] == g for g in range(prob.conf_ngroups[1])])
    prior_cfg = dict(
        objects_per_cluster=dict(type="uniform_area", min=prob.min_size, max=prob.max_size),
        geo=dict(type="uniform"), weights=dict(type="uniform"), cluster_effect=dict(type="uniform"),
        confounding_effects=dict(universal={"<ALL>": dict(type="uniform")},
                                 family={f"g{i}": dict(type="uniform") for i in range(prob.conf_ngroups[1])}))
    data, model = H.build_reference_model(prob.x, prob.applicable, truth["locations"], confs, prob.K, prior_cfg)
    return prob, data, model


@pytest.mark.parametrize("T,Tp", [(1.0, 1.0), (1.2, 1.1)])
def test_dropin_chain_matches_oracle(T, Tp):
    from sbayes_b200.chain import MCMCChain
    r = H.ref()
    prob0, data, model = _reference_problem()
    opcfg = r.config.OperatorsConfig(clusters=70, weights=10, source=20)
    lg = CountingLogger()
    chain = MCMCChain(model, data, opcfg, [lg], temperature=T, prior_temperature=Tp, seed=321, team=2, _lib=_lib())
    prob = chain.prob
    st0 = synth.initial_state(prob, seed=3, objects_per_cluster=6)
    st0.counts = M.rebuild_counts(prob, st0)
    sample0 = H.sample_from_state(prob, model, data, st0)
    # the reference's own evaluation of the start state agrees with the device's
    ll_ref, pr_ref = model.likelihood(sample0, caching=False), model.prior(sample0, caching=False)
    assert abs(chain.likelihood(sample0) * T - ll_ref) < 5e-7 * max(1, abs(ll_ref))
    assert abs(chain.prior(sample0) * Tp - pr_ref) < 5e-7 * max(1, abs(pr_ref))

    out = chain.run(n_steps=90, logging_interval=40, initial_sample=sample0)
    assert out.i_step == 90 and [row[0] for row in lg.rows] == [40, 80]
    # same trajectory as the fp64 oracle on the same Philox stream
    oc = CH.OracleChain(prob, ChainConfig(temperature=T, prior_temperature=Tp, seed=321), st0.copy(), draws=PhiloxDraws(321))
    for _ in range(90):
        oc.step()
    fin = H.state_from_sample(prob, out)
    assert np.array_equal(fin.cluster_of, oc.state.cluster_of)
    assert np.array_equal(fin.src, oc.state.src)
    assert np.array_equal(fin.counts, oc.state.counts)
    assert np.array_equal(fin.W, oc.state.W)
    assert abs(out.last_lh - oc.ll) <= 1e-10 * max(1, abs(oc.ll))
    assert abs(out.last_prior - oc.prior) <= 1e-10 * max(1, abs(oc.prior))
    assert abs(chain._ll - oc.ll / T) <= 1e-10 * max(1, abs(oc.ll))
    # the returned Sample is a valid reference Sample: the reference's own model agrees
    assert abs(model.likelihood(out, caching=False) - out.last_lh) < 5e-7 * max(1, abs(out.last_lh))
    assert abs(model.prior(out, caching=False) - out.last_prior) < 5e-7 * max(1, abs(out.last_prior))
    # operator statistics arrive on the reference's operator objects
    acc = sum(op.accepts for op in chain.callable_operators.values())
    rej = sum(op.rejects for op in chain.callable_operators.values())
    assert acc + rej == 90 and acc == int(oc.accepts.sum())
    # a second run continues from the returned sample (MC3 hands samples back and forth)
    out2 = chain.run(n_steps=10, logging_interval=40, initial_sample=out)
    assert out2.i_step == 100
    chain.close()


def test_install_rebinds_the_reference_names():
    from sbayes_b200 import chain as C
    r = H.ref()
    orig = r.mcmc_chain.MCMCChain
    try:
        C.install()
        assert r.mcmc_chain.MCMCChain is C.MCMCChain and r.mcmc_setup.MCMCChain is C.MCMCChain
    finally:
        r.mcmc_chain.MCMCChain = orig
        r.mcmc_setup.MCMCChain = orig
