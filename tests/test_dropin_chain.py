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


# every test runs twice: in the fibre emulator (GPU-less container) and, `gpu`-marked, on the
# B200 through the CUDA library (the reference then comes from oracle/_ref, see oracle/make_ref.py)
BACKENDS = ["emu", pytest.param("cuda", marks=pytest.mark.gpu)]


def _lib(backend):
    if backend == "cuda":
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


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("T,Tp", [(1.0, 1.0), (1.2, 1.1)])
def test_dropin_chain_matches_oracle(T, Tp, backend):
    from sbayes_b200.chain import MCMCChain
    r = H.ref()
    prob0, data, model = _reference_problem()
    opcfg = r.config.OperatorsConfig(clusters=70, weights=10, source=20)
    lg = CountingLogger()
    chain = MCMCChain(model, data, opcfg, [lg], temperature=T, prior_temperature=Tp, seed=321, team=2, _lib=_lib(backend))
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
    # the cache nodes of the returned Sample carry the DEVICE's values: the reference's cached calls return them
    # without recomputing (this is what swap_chains and ParametersCSVLogger read), and they agree with its own
    # uncached evaluation
    c = out.cache
    assert not c.source_prior.is_outdated() and not c.cluster_size_prior.is_outdated() and not c.weights_prior.is_outdated()
    assert all(not n.is_outdated() for n in c.group_likelihoods.values())
    assert abs(model.prior(out) - out.last_prior) <= 1e-9 * max(1, abs(out.last_prior))
    assert abs(model.likelihood(out) - out.last_lh) <= 1e-9 * max(1, abs(out.last_lh))
    ref_sp = H.ref().prior.SourcePrior.__call__(model.prior.source_prior, H.sample_from_state(prob, model, data, fin), caching=False)
    assert abs(float(np.sum(c.source_prior.value)) - ref_sp) < 5e-7 * max(1, abs(ref_sp))
    for row in lg.rows:
        assert abs(row[3]) < 1e-12 or np.isfinite(row[3])
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
        cls = C.install()
        assert issubclass(cls, C.MCMCChain) and r.mcmc_chain.MCMCChain is cls and r.mcmc_setup.MCMCChain is cls
    finally:
        r.mcmc_chain.MCMCChain = orig
        r.mcmc_setup.MCMCChain = orig


@pytest.mark.parametrize("backend", BACKENDS)
def test_dropin_mcmc_generate_samples_matches_oracle(backend):
    """Boundary B1': `MCMC.generate_samples` (sbayes/sampling/mcmc.py:155-263) -- warm-up picks the
    best of several chains, sampling logs chain 0 at the reference's step indices."""
    import logging
    from sbayes_b200.chain import MCMC
    r = H.ref()
    prob0, data, model = _reference_problem()
    opcfg = r.config.OperatorsConfig(clusters=70, weights=10, source=20)
    st0 = synth.initial_state(prob0, seed=3, objects_per_cluster=6)

    # ---- warm-up: 3 chains from the same sample, 25 steps, best log-likelihood wins ----------
    warm = MCMC(model, data, opcfg, [], n_chains=3, logger=logging.getLogger("t"), seed=50, team=1, _lib=_lib(backend))
    prob = warm.prob
    st0.counts = M.rebuild_counts(prob, st0)
    sample0 = H.sample_from_state(prob, model, data, st0)
    best = warm.generate_samples(n_steps=0, n_samples=0, initial_sample=sample0, warm_up=True, warm_up_steps=25)
    ocs = [CH.OracleChain(prob, ChainConfig(seed=50 + c), st0.copy(), draws=PhiloxDraws(50 + c)) for c in range(3)]
    for oc in ocs:
        for _ in range(25):
            oc.step()
    b = int(np.argmax([oc.ll for oc in ocs]))
    fin = H.state_from_sample(prob, best)
    assert best.chain == b
    assert np.array_equal(fin.cluster_of, ocs[b].state.cluster_of) and np.array_equal(fin.src, ocs[b].state.src)
    assert np.array_equal(fin.counts, ocs[b].state.counts) and np.array_equal(fin.W, ocs[b].state.W)
    assert abs(best.last_lh - ocs[b].ll) <= 1e-10 * abs(ocs[b].ll)
    warm.close()

    # ---- sampling: 50 steps, 5 samples -> logged after the steps with index 0, 10, 20, 30, 40 --
    lg = CountingLogger()
    smp = MCMC(model, data, opcfg, [lg], logger=logging.getLogger("t"), seed=77, team=2, _lib=_lib(backend))
    sample0.i_step = 0
    out = smp.generate_samples(n_steps=50, n_samples=5, initial_sample=sample0)
    assert [row[0] for row in lg.rows] == [0, 10, 20, 30, 40] and out.i_step == 49
    oc = CH.OracleChain(prob, ChainConfig(seed=77), st0.copy(), draws=PhiloxDraws(77))
    logged = {}
    for i in range(50):
        oc.step()
        if i % 10 == 0:
            logged[i] = (oc.ll, oc.prior, oc.state.cluster_of.copy())
    for row in lg.rows:
        ll, pr, co = logged[row[0]]
        assert abs(row[1] - ll) <= 1e-10 * abs(ll) and abs(row[2] - pr) <= 1e-10 * abs(pr)
        assert np.array_equal(np.where(row[5].any(axis=0), row[5].argmax(axis=0), -1), co)
    fin = H.state_from_sample(prob, out)
    assert np.array_equal(fin.src, oc.state.src) and np.array_equal(fin.counts, oc.state.counts)
    acc = sum(op.accepts for op in smp.callable_operators.values())
    rej = sum(op.rejects for op in smp.callable_operators.values())
    assert acc + rej == 50 and acc == int(oc.accepts.sum())
    smp.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_device_model_matches_reference_model(backend):
    """Boundary B2 (model API): Model.likelihood / Model.prior of a reference Sample on the device"""
    from sbayes_b200.chain import DeviceModel
    prob0, data, model = _reference_problem()
    dm = DeviceModel(model, data, _lib=_lib(backend))
    for seed in (3, 4):
        st = synth.initial_state(dm.prob, seed=seed, objects_per_cluster=6)
        st.counts = M.rebuild_counts(dm.prob, st)
        s = H.sample_from_state(dm.prob, model, data, st)
        ll_ref, pr_ref = model.likelihood(s, caching=False), model.prior(s, caching=False)
        assert abs(dm.likelihood(s) - ll_ref) < 5e-7 * abs(ll_ref) and abs(dm.prior(s) - pr_ref) < 5e-7 * abs(pr_ref)
        assert abs(dm.likelihood(s) - M.log_likelihood(dm.prob, st)) <= 1e-10 * abs(ll_ref)
        assert abs(dm.prior(s) - M.log_prior(dm.prob, st)) <= 1e-10 * abs(pr_ref)
        assert s.last_lh == dm.likelihood(s) and s.last_prior == dm.prior(s)
    dm.close()


def test_install_rebinds_mcmc_too():
    from sbayes_b200 import chain as C
    r = H.ref()
    import sbayes.sampling.mcmc as mm
    orig = (r.mcmc_chain.MCMCChain, mm.MCMC)
    try:
        C.install()
        assert issubclass(mm.MCMC, C.MCMC) and r.mcmc_setup.MCMC is mm.MCMC
    finally:
        r.mcmc_chain.MCMCChain = orig[0]; r.mcmc_setup.MCMCChain = orig[0]
        mm.MCMC = orig[1]; r.mcmc_setup.MCMC = orig[1]


@pytest.mark.parametrize("backend", BACKENDS)
def test_device_operator_api_matches_oracle_proposals(backend):
    """Boundary B2 (operator API): `op.function(sample) -> (sample', log_q, log_q_back)` of every
    scheduled operator, executed on the device, against the oracle's proposal on the same draws."""
    from oracle import operators as O
    from sbayes_b200.chain import device_operator_schedule
    from sbayes_b200.packed import OP_NAMES
    r = H.ref()
    prob0, data, model = _reference_problem()
    opcfg = r.config.OperatorsConfig(clusters=70, weights=10, source=20)
    ops = device_operator_schedule(opcfg, model, data, seed=4242, _lib=_lib(backend))
    ref_ops = r.operators.get_operator_schedule(operators_config=opcfg, model=model, data=data)
    engine = next(iter(ops.values())).device_engine
    prob = engine.prob
    st = synth.initial_state(prob, seed=3, objects_per_cluster=6)
    st.counts = M.rebuild_counts(prob, st)
    sample = H.sample_from_state(prob, model, data, st)
    n_changed = 0
    for rnd in range(3):
        for i, name in enumerate(OP_NAMES):
            op = ops[name]
            assert type(op).__name__ == type(ref_ops[name]).__name__ and op.weight == ref_ops[name].weight
            assert op.get_parameters_string() == ref_ops[name].get_parameters_string()
            step = engine.calls + 1
            new, lq, lqb = op.function(sample)
            draws = PhiloxDraws(4242)
            draws.begin_step(step)
            cur = H.state_from_sample(prob, sample)
            prop = O.propose(prob, cur, ChainConfig(seed=4242), draws, i)
            if not prop.changed or prop.log_q_back == O.NEG_INF:
                assert new is sample and lq == op.Q_REJECT and lqb == op.Q_BACK_REJECT
                continue
            n_changed += 1
            got = H.state_from_sample(prob, new)
            assert np.array_equal(got.cluster_of, prop.state.cluster_of) and np.array_equal(got.src, prop.state.src)
            assert np.array_equal(got.W, prop.state.W) and np.array_equal(got.counts, prop.state.counts)
            if prop.log_q == O.NEG_INF:
                assert lq == op.Q_GIBBS and lqb == op.Q_BACK_GIBBS
            else:
                assert abs(lq - prop.log_q) <= 1e-10 * max(1, abs(prop.log_q))
                assert abs(lqb - prop.log_q_back) <= 1e-10 * max(1, abs(prop.log_q_back))
            # the proposal is a valid reference Sample
            assert abs(model.likelihood(new, caching=False) - M.log_likelihood(prob, prop.state)) < 5e-7 * abs(M.log_likelihood(prob, prop.state))
            sample = new
    assert n_changed >= 15
    engine.ctx.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_restarted_step_index_gets_fresh_draws(backend):
    """The reference restarts `i_step` on the SAME chain object (every warm-up attempt of
    MCMCChainProcess.initialize_sample starts at 0, mcmc_setup.py:503-523).  The device addresses
    its draws by (key, step), so the key must change with every restart: two runs from the same
    sample must not replay the same operator choices and uniforms."""
    from sbayes_b200.chain import MCMCChain
    r = H.ref()
    prob0, data, model = _reference_problem()
    opcfg = r.config.OperatorsConfig(clusters=70, weights=10, source=20)
    chain = MCMCChain(model, data, opcfg, [], seed=11, team=1, _lib=_lib(backend))
    st0 = synth.initial_state(chain.prob, seed=3, objects_per_cluster=6)
    st0.counts = M.rebuild_counts(chain.prob, st0)
    sample0 = H.sample_from_state(chain.prob, model, data, st0)
    outs = []
    for attempt in range(3):
        out = chain.run(n_steps=60, logging_interval=10 ** 9, initial_sample=sample0)
        outs.append(H.state_from_sample(chain.prob, out))
        chain.reset_posterior_cache()
    assert chain._epoch == 2
    for a in range(3):
        for b in range(a + 1, 3):
            assert not (np.array_equal(outs[a].src, outs[b].src) and np.array_equal(outs[a].W, outs[b].W)), \
                "restarted runs replayed the same draws"
    # a continuation (i_step goes on where the last run ended) keeps its key
    cont = chain.run(n_steps=10, logging_interval=10 ** 9, initial_sample=sample0)
    e = chain._epoch
    chain.run(n_steps=10, logging_interval=10 ** 9, initial_sample=cont)
    assert chain._epoch == e
    chain.close()
