"""Effect draws at log time (`sb_effect_sample`: conditional_effect_sample, sbayes/sampling/conditionals.py:125-149, as the
reference's parameter logger calls it, sbayes/sampling/loggers.py:160-224).

* device == oracle (oracle/effects.py) on the same Philox streams: 1e-10 relative (fp64; the emulator is exact), with
  concentrations below and above one, jagged state sets and counts taken after a few hundred MCMC steps;
* the draws have the Dirichlet moments of (counts + prior concentration);
* `ref`: the concentration the device draws from IS the reference's `conditional_effect_concentration` on the same
  sample, and the reference's own `conditional_effect_sample` has the same mean."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
import build as HB  # noqa: E402
from oracle import effects as E, model as M, refharness as H  # noqa: E402
from sbayes_b200 import synth  # noqa: E402
from sbayes_b200.backend import DeviceContext  # noqa: E402
from sbayes_b200.packed import ChainConfig  # noqa: E402

BACKENDS = ["emu", pytest.param("cuda", marks=pytest.mark.gpu)]


def _lib(backend):
    return None if backend == "cuda" else HB.load()


def _problem(N, F, S, K, nf, seed, low=0.3, high=2.5):
    prob, truth = synth.make_problem(N=N, F=F, S=S, K=K, n_fam=nf, seed=seed, jagged_states=True, max_size=max(20, N // 3))
    # concentrations below one for the cluster effect (boosted gamma draws), above one for a confounder
    prob.alpha_cluster = prob.alpha_cluster * low
    prob.alpha_conf[-1] = prob.alpha_conf[-1] * high
    return prob, truth


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("shape", [(40, 9, 3, 2, 3), (120, 30, 5, 3, 5)])
def test_effect_draws_match_oracle(backend, shape):
    N, F, S, K, nf = shape
    prob, _ = _problem(N, F, S, K, nf, seed=11)
    cfgs = [ChainConfig(seed=501), ChainConfig(seed=502, temperature=1.2, prior_temperature=1.1)]
    ctx = DeviceContext(prob, 2, team=2, _lib=_lib(backend))
    try:
        for c in range(2):
            ctx.configure(c, cfgs[c])
            ctx.set_state(c, synth.initial_state(prob, seed=70 + c, objects_per_cluster=5))
        ctx.run(150 if backend == "cuda" else 60)
        for step, chains in ((7, [0, 1]), (1 << 33, [1])):
            got = ctx.effect_sample(chains, step=step)
            for i, c in enumerate(chains):
                st = ctx.get_state(c)
                assert np.array_equal(st.counts, M.rebuild_counts(prob, st))
                want = E.effect_sample(prob, st.counts, cfgs[c].seed, step)
                assert got[i].shape == want.shape
                assert np.all(got[i][prob.alpha_stacked() == 0.0] == 0.0)
                assert np.allclose(got[i].sum(axis=-1), 1.0, atol=1e-12)
                assert np.abs(got[i] - want).max() < 1e-10, f"chain {c} step {step}"
    finally:
        ctx.close()


def test_effect_draws_have_dirichlet_moments():
    prob, _ = _problem(30, 6, 3, 2, 2, seed=5, low=0.4, high=1.0)
    ctx = DeviceContext(prob, 1, _lib=HB.load())
    try:
        ctx.configure(0, ChainConfig(seed=9))
        st = synth.initial_state(prob, seed=3, objects_per_cluster=4)
        ctx.set_state(0, st)
        conc = E.concentration(prob, M.rebuild_counts(prob, st))
        n = 400
        draws = np.stack([ctx.effect_sample([0], step=1000 + i)[0] for i in range(n)])
    finally:
        ctx.close()
    a0 = conc.sum(axis=-1, keepdims=True)
    mean, var = conc / a0, conc * (a0 - conc) / (a0 * a0 * (a0 + 1.0))
    z = (draws.mean(axis=0) - mean) / np.sqrt(np.maximum(var, 1e-300) / n)
    assert np.abs(z[conc > 0]).max() < 5.0
    # second moment: sample variance within 35 % of the Dirichlet variance wherever it is not tiny
    big = var > 1e-3
    assert np.all(np.abs(draws.var(axis=0)[big] / var[big] - 1.0) < 0.35)


@pytest.mark.ref
def test_concentration_is_the_references():
    from test_dropin_chain import _reference_problem
    r = H.ref()
    prob0, data, model = _reference_problem()
    prob = H.problem_from_reference(data, model)
    st = synth.initial_state(prob, seed=3, objects_per_cluster=6)
    st.counts = M.rebuild_counts(prob, st)
    sample = H.sample_from_state(prob, model, data, st)
    from sbayes.sampling.conditionals import conditional_effect_concentration, conditional_effect_sample
    feats = data.features
    offs = prob.group_offsets()
    ours = E.concentration(prob, st.counts)
    # cluster effect (loggers.py:162-167) and every confounding effect (loggers.py:217-222)
    ref_cl = conditional_effect_concentration(
        features=feats.values, is_source_group=sample.clusters.value[..., np.newaxis] & sample.source.value[np.newaxis, ..., 0],
        applicable_states=feats.states, prior_counts=model.prior.prior_cluster_effect.concentration_array)
    assert np.array_equal(ref_cl, ours[:prob.K])
    for i_conf, conf in enumerate(data.confounders.values(), start=1):
        isg = conf.group_assignment[..., np.newaxis] & sample.source.value[np.newaxis, ..., i_conf]
        pc = model.prior.prior_confounding_effects[conf.name].concentration_array(sample)
        ref_c = conditional_effect_concentration(features=feats.values, is_source_group=isg, applicable_states=feats.states, prior_counts=pc)
        assert np.array_equal(ref_c, ours[offs[i_conf]:offs[i_conf + 1]])
    # the reference's own sampler has the mean the device draws have (Dirichlet mean of that concentration)
    np.random.seed(4)
    isg = sample.clusters.value[..., np.newaxis] & sample.source.value[np.newaxis, ..., 0]
    draws = np.stack([conditional_effect_sample(feats.values, isg, feats.states, model.prior.prior_cluster_effect.concentration_array)
                      for _ in range(200)])
    conc = ours[:prob.K]
    a0 = conc.sum(axis=-1, keepdims=True)
    var = conc * (a0 - conc) / (a0 * a0 * (a0 + 1.0))
    z = (draws.mean(axis=0) - conc / a0) / np.sqrt(np.maximum(var, 1e-300) / 200)
    assert np.abs(z[conc > 0]).max() < 5.0
    ctx = DeviceContext(prob, 1, _lib=HB.load())
    try:
        ctx.configure(0, ChainConfig(seed=77))
        ctx.set_state(0, st)
        dev = np.stack([ctx.effect_sample([0], step=i)[0][:prob.K] for i in range(200)])
    finally:
        ctx.close()
    z = (dev.mean(axis=0) - conc / a0) / np.sqrt(np.maximum(var, 1e-300) / 200)
    assert np.abs(z[conc > 0]).max() < 5.0


@pytest.mark.ref
@pytest.mark.parametrize("backend", BACKENDS)
def test_parameter_logger_takes_device_draws(backend, tmp_path):
    """the reference's own ParametersCSVLogger, fed by the drop-in chain: its effect columns are the device's draws"""
    from test_dropin_chain import _reference_problem
    from sbayes_b200 import chain as CHN
    r = H.ref()
    import sbayes.sampling.loggers as lg
    prob0, data, model = _reference_problem()
    CHN.install_effect_logging()
    logger = lg.ParametersCSVLogger(str(tmp_path / "stats.txt"), data, model, False, match_clusters=False)
    opcfg = r.config.OperatorsConfig(clusters=70, weights=10, source=20)
    ch = CHN.MCMCChain(model, data, opcfg, [logger], temperature=1.0, prior_temperature=1.0, seed=99, team=2, _lib=_lib(backend))
    prob = ch.prob
    st0 = synth.initial_state(prob, seed=3, objects_per_cluster=6)
    st0.counts = M.rebuild_counts(prob, st0)
    before = dict(CHN.effect_logging_stats)
    out = ch.run(n_steps=60, logging_interval=30, initial_sample=H.sample_from_state(prob, model, data, st0))
    logger.close()
    n_eff = 1 + len(prob.conf_gid)
    assert CHN.effect_logging_stats["device"] - before["device"] == 2 * n_eff      # two logged samples
    assert CHN.effect_logging_stats["host"] == before["host"]
    # the last row's areal columns are the oracle's draws from the final state's counts (same Philox streams)
    rows = [l.rstrip("\n").split("\t") for l in open(tmp_path / "stats.txt")]
    head, last = rows[0], rows[-1]
    st = H.state_from_sample(prob, out)
    want = E.effect_sample(prob, M.rebuild_counts(prob, st), ch._cfg.seed, 60)
    for name, k, f, s in (("areal_a0_f0_s0", 0, 0, 0), ("areal_a1_f3_s1", 1, 3, 1)):
        assert abs(float(last[head.index(name)]) - want[k, f, s]) < 1e-7        # (the logger prints 8 significant digits)
    off = prob.group_offsets()[2]
    assert abs(float(last[head.index("family_g1_f2_s0")]) - want[off + 1, 2, 0]) < 1e-7
