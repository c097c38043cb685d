"""GPU parity at the BENCHMARK shapes (run on the B200 box): every shape `bench.py` times has a
parity test here, so no performance number is quoted on a shape the device code was never
checked on.

* c5 object shard 25 000 x 1000 x 8, K=16, 64 families: statistics pass (shared-memory
  histogram AND int8 tcgen05 contraction), terms, observation likelihood vs the fp64 oracle
* a lockstep at a 1000-feature width (the underflow width of the reference, SURVEY F6: the
  device and the oracle work in log space)
* long lockstep at c2 (4 tempered chains x 1200 steps) and c4 (2 chains x 320 steps): identical
  decisions and states, fp64 values to 1e-10, oracle chains spread over host processes
"""
import os

import numpy as np
import pytest

import parity
from oracle import model as M
from sbayes_b200 import synth
from sbayes_b200.backend import DeviceContext
from sbayes_b200.packed import ChainConfig

pytestmark = pytest.mark.gpu


def test_c5_shard_shape_full_eval_and_obs_lik():
    prob, _ = synth.make_problem(N=25000, F=1000, S=8, K=16, n_fam=64, seed=5)
    st = synth.initial_state(prob, seed=2)
    counts = M.rebuild_counts(prob, st)
    st.counts = counts
    ll, pr = M.log_likelihood(prob, st), M.log_prior(prob, st)
    ctx = DeviceContext(prob, 1)
    try:
        ctx.configure(0, ChainConfig(seed=1))
        ctx.set_state(0, st)
        for mode in ("hist", "mma"):
            ctx.eval_set_mode(mode)
            ev = ctx.eval_batch()[0]
            got = ctx.get_state(0).counts          # the statistics block the pass just filled
            assert np.array_equal(got, counts), f"{mode}: counts differ from the oracle"
            assert parity.rel(ll, ev["ll"]) < 1e-10 and parity.rel(pr, ev["prior"]) < 1e-10, mode
        ctx.eval_set_mode("hist")
        # the same statistics through the single-chain rebuild entry points
        for mma in (False, True):
            c2, _ = ctx.counts_rebuild(0, mma=mma)
            assert np.array_equal(c2, counts)
        # uncached evaluation through the persistent kernel (148-CTA team)
        ev2 = ctx.eval(0)
        assert parity.rel(ll, ev2["ll"]) < 1e-10 and parity.rel(pr, ev2["prior"]) < 1e-10
        # observation likelihood: float32 [N, F] exact to rounding, fp64 sums 1e-10
        lh, ob, sums, _ = ctx.obs_loglik()
        ref = M.observation_likelihood(prob, st)
        assert np.allclose(lh[0], ref.astype(np.float32), rtol=3e-7, atol=0)
        ref_ll = np.log(ref).sum(axis=1)
        assert np.allclose(ob[0], ref_ll, rtol=1e-10, atol=1e-9)
        assert parity.rel(float(ref_ll.sum()), float(sums[0])) < 1e-10
    finally:
        ctx.close()


def test_lockstep_thousand_features():
    """F = 1000, S = 8: products over 1000 features underflow in linear space."""
    prob, _ = synth.make_problem(N=400, F=1000, S=8, K=4, n_fam=6, seed=9)
    cfgs = [ChainConfig(seed=500), ChainConfig(temperature=1.2, prior_temperature=1.1, seed=501)]
    st0 = [synth.initial_state(prob, seed=70 + c) for c in range(2)]
    parity.philox_lockstep(prob, st0, cfgs, 60, team=0, procs=2)


def test_lockstep_c5_like_shard_stepping():
    """a 16-cluster, 64-family, 1000-feature problem stepped by a large team (the single-GPU
    placement of config 5)"""
    prob, _ = synth.make_problem(N=3000, F=1000, S=8, K=16, n_fam=64, seed=5)
    cfgs = [ChainConfig(seed=600)]
    st0 = [synth.initial_state(prob, seed=80)]
    parity.philox_lockstep(prob, st0, cfgs, 25, team=0)


@pytest.mark.timeout(1500)
def test_long_lockstep_c2_and_c4():
    """>= 1000 steps at config 2 (4 MC3 temperatures) and >= 300 at config 4 (one CTA per chain,
    replicas in HBM slices).  The six oracle chains run in six host processes while the device
    runs its launches."""
    p2, _ = synth.make_problem(N=1000, F=100, S=4, K=5, n_fam=8, seed=2)
    cfg2 = [ChainConfig(temperature=1.0 + 0.05 * c, prior_temperature=1.0, seed=40 + c) for c in range(4)]
    st2 = [synth.initial_state(p2, seed=50 + c) for c in range(4)]
    p4, _ = synth.make_problem(N=5000, F=200, S=6, K=5, n_fam=20, seed=4)
    cfg4 = [ChainConfig(seed=300), ChainConfig(temperature=1.1, prior_temperature=1.05, seed=301)]
    st4 = [synth.initial_state(p4, seed=100 + c) for c in range(2)]
    n2, n4 = 1200, 320
    # c4 jobs first: they are the long ones
    jobs = [(p4, cfg4[c], st4[c], n4) for c in range(2)] + [(p2, cfg2[c], st2[c], n2) for c in range(4)]
    trajs = parity.oracle_trajectories(jobs, procs=min(6, os.cpu_count() or 1))
    for prob, cfgs, sts, n, tr, team in ((p2, cfg2, st2, n2, trajs[2:], 0), (p4, cfg4, st4, n4, trajs[:2], 1)):
        ctx = DeviceContext(prob, len(cfgs), team=team)
        try:
            for c, cfg in enumerate(cfgs):
                ctx.configure(c, cfg)
                ctx.set_state(c, sts[c])
            rec = ctx.run(n, records=True)
            parity.compare_with_trajectories(ctx, rec, tr)
            acc = sum(int(t["accepts"].sum()) for t in tr)
            assert acc > 0
        finally:
            ctx.close()
