"""Full-sweep source Gibbs (`sb_gibbs_sweep`: GibbsSampleSource with ObjectSelector.ALL,
sbayes/sampling/operators.py:496-574) against the oracle's restatement of the same proposal:
new sources bit-exact, log_q / log_q_back within 1e-10, counts of the new state bit-exact.
Emulated in the GPU-less container, on the B200 under `-m gpu` (also at the config-4 / 5 widths)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
import build as HB  # noqa: E402
import parity  # noqa: E402
from oracle import model as M, operators as O  # noqa: E402
from sbayes_b200 import synth  # noqa: E402
from sbayes_b200.backend import DeviceContext  # noqa: E402
from sbayes_b200.packed import ChainConfig  # noqa: E402


def check_sweep(prob, states, cfgs, lib=None, steps=(3, 4)):
    ctx = DeviceContext(prob, len(states), _lib=lib)
    try:
        for c, st in enumerate(states):
            ctx.configure(c, cfgs[c])
            ctx.set_state(c, st)
        cur = [s.copy() for s in states]
        for step in steps:
            lq, ch, _ = ctx.gibbs_sweep(step=step, mode="draw", full=True)
            lqb, _, _ = ctx.gibbs_sweep(step=step, mode="score_previous", full=True)
            lq2, _, _ = ctx.gibbs_sweep(step=step, mode="score", full=True)
            ev = ctx.eval_batch()
            for c in range(len(states)):
                ref, rq, rqb, rch = O.gibbs_sweep(prob, cur[c], cfgs[c], step)
                got = ctx.get_state(c)
                assert np.array_equal(got.src, ref.src), f"chain {c} step {step}: resampled sources differ"
                assert np.array_equal(got.cluster_of, ref.cluster_of) and np.array_equal(got.W, ref.W)
                assert np.array_equal(got.counts, ref.counts)          # statistics of the new state
                assert int(ch[c]) == rch
                assert parity.rel(rq, lq[c]) < 1e-10 and parity.rel(rqb, lqb[c]) < 1e-10
                # scoring the committed (new) sources under the new counts
                want = sum(O.log_prob_of(O.source_posterior(prob, ref, np.arange(a, min(prob.N, a + 512)), cfgs[c]), ref.src[a:a + 512])
                           for a in range(0, prob.N, 512))
                assert parity.rel(want, lq2[c]) < 1e-10
                assert parity.rel(M.log_likelihood(prob, ref), ev[c]["ll"]) < 1e-10 or cfgs[c].sample_from_prior
                cur[c] = ref
    finally:
        ctx.close()


@pytest.fixture(scope="module")
def lib():
    return HB.load()


def test_emulated_gibbs_sweep(lib):
    prob, _ = synth.make_problem(N=70, F=37, S=4, K=3, n_fam=4, seed=7, jagged_states=True, max_size=20)
    states = [synth.initial_state(prob, seed=5 + i, objects_per_cluster=5) for i in range(3)]
    cfgs = [ChainConfig(seed=11), ChainConfig(temperature=1.3, prior_temperature=1.2, seed=12), ChainConfig(seed=13, sample_from_prior=True)]
    check_sweep(prob, states, cfgs, lib=lib)


def test_emulated_gibbs_sweep_six_components(lib):
    from test_hostsim_parity import _many_confounders
    prob = _many_confounders(3)
    states = [synth.initial_state(prob, seed=40 + i, objects_per_cluster=4) for i in range(2)]
    check_sweep(prob, states, [ChainConfig(seed=7), ChainConfig(temperature=1.15, prior_temperature=1.05, seed=8)], lib=lib)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(100, 36, 5, 3, 6), (1000, 100, 4, 5, 8), (5000, 200, 6, 5, 20), (3000, 1000, 8, 16, 64)])
def test_gpu_gibbs_sweep(shape):
    N, F, S, K, nf = shape
    prob, _ = synth.make_problem(N=N, F=F, S=S, K=K, n_fam=nf, seed=1)
    states = [synth.initial_state(prob, seed=2 + i) for i in range(2)]
    check_sweep(prob, states, [ChainConfig(seed=21), ChainConfig(temperature=1.2, prior_temperature=1.1, seed=22)], steps=(5,))
