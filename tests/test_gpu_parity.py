"""GPU parity tests (run on the B200 box): the CUDA library, through the C ABI, against
(a) the golden fixtures recorded from the live reference and (b) the fp64 oracle."""
import numpy as np
import pytest

import parity
from conftest import golden_cases
from oracle import model as M
from sbayes_b200 import synth
from sbayes_b200.backend import DeviceContext
from sbayes_b200.packed import ChainConfig

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", golden_cases())
@pytest.mark.parametrize("team", [1, 3])
def test_replay_reference_tape(case, team):
    parity.replay_golden(case, team=team)


def test_replay_team_auto():
    parity.replay_golden("synth_c2_width", team=0)


@pytest.mark.parametrize("team", [1, 8])
def test_philox_lockstep_small(team):
    prob, _ = synth.make_problem(N=80, F=16, S=4, K=3, n_fam=4, seed=7, jagged_states=True, max_size=30)
    cfgs = [ChainConfig(seed=10), ChainConfig(temperature=1.3, prior_temperature=1.2, seed=11),
            ChainConfig(seed=12, sample_from_prior=True)]
    st0 = [synth.initial_state(prob, seed=20 + i, objects_per_cluster=6) for i in range(3)]
    parity.philox_lockstep(prob, st0, cfgs, 400, team=team)


def test_philox_lockstep_geo():
    prob, _ = synth.make_problem(N=48, F=10, S=3, K=2, n_fam=3, seed=3, geo_cost_based=True, max_size=20)
    cfgs = [ChainConfig(seed=5), ChainConfig(temperature=1.1, seed=6)]
    st0 = [synth.initial_state(prob, seed=30 + i, objects_per_cluster=5) for i in range(2)]
    parity.philox_lockstep(prob, st0, cfgs, 300, team=2)


def test_philox_lockstep_c2_shape():
    """BASELINE config 2 shape (1000 x 100 x 4, K=5, 8 families), 4 MC3 temperatures."""
    prob, _ = synth.make_problem(N=1000, F=100, S=4, K=5, n_fam=8, seed=2)
    cfgs = [ChainConfig(temperature=1.0 + 0.05 * c, prior_temperature=1.0, seed=40 + c) for c in range(4)]
    st0 = [synth.initial_state(prob, seed=50 + c) for c in range(4)]
    parity.philox_lockstep(prob, st0, cfgs, 60, team=0)


@pytest.mark.parametrize("shape", [(5, 2, 2, 2, 2), (100, 36, 5, 3, 6), (1000, 100, 4, 5, 8), (5000, 200, 6, 5, 20)])
def test_full_eval_matches_oracle(shape):
    """Count rebuild bit-exact, log-likelihood and priors 1e-10 relative (north_star)."""
    N, F, S, K, nf = shape
    prob, _ = synth.make_problem(N=N, F=F, S=S, K=K, n_fam=nf, seed=1, min_size=1 if N < 10 else 3)
    st = synth.initial_state(prob, seed=2, objects_per_cluster=1 if N < 10 else 10)
    ctx = DeviceContext(prob, 1)
    try:
        ctx.set_state(0, st)
        ev = ctx.eval(0)
        got = ctx.get_state(0)
        assert np.array_equal(got.counts, M.rebuild_counts(prob, st))
        hist, _ = ctx.counts_rebuild(0)
        assert np.array_equal(hist, got.counts)
        st.counts = got.counts
        assert parity.rel(M.log_likelihood(prob, st), ev["ll"]) < 1e-10
        assert parity.rel(M.log_prior(prob, st), ev["prior"]) < 1e-10
        assert np.array_equal(got.src, st.src) and np.array_equal(got.cluster_of, st.cluster_of)
    finally:
        ctx.close()


def test_incremental_counts_after_long_run():
    """size-independent property at the bench shape: after many steps the incrementally
    maintained counts equal a bulk rebuild, and cached posterior == uncached evaluation."""
    prob, _ = synth.make_problem(N=1000, F=100, S=4, K=5, n_fam=8, seed=2)
    ctx = DeviceContext(prob, 4)
    try:
        for c in range(4):
            ctx.configure(c, ChainConfig(temperature=1.0 + 0.05 * c, seed=c))
            ctx.set_state(c, synth.initial_state(prob, seed=c))
        ctx.run(3000)
        for c in range(4):
            st = ctx.get_state(c)
            assert np.array_equal(st.counts, M.rebuild_counts(prob, st))
            hist, _ = ctx.counts_rebuild(c)
            assert np.array_equal(hist, st.counts)
            assert np.all((st.src == 255) == (prob.x == 255))
            post = ctx.posterior(c)
            assert parity.rel(M.log_likelihood(prob, st), post["ll"]) < 1e-10
            assert parity.rel(M.log_prior(prob, st), post["prior"]) < 1e-10
            a, r = ctx.op_stats(c)
            assert int(a.sum() + r.sum()) == 3000
    finally:
        ctx.close()


def test_mc3_swap_matches_oracle():
    from oracle import chain as CH
    from oracle.draws import PhiloxDraws
    prob, _ = synth.make_problem(N=100, F=20, S=4, K=3, n_fam=4, seed=9, max_size=30)
    n = 4
    T, Tp = CH.temperature_ladder(n, 0.3, 0.1, False)
    ctx = DeviceContext(prob, n)
    try:
        sts = [synth.initial_state(prob, seed=60 + c) for c in range(n)]
        for c in range(n):
            ctx.configure(c, ChainConfig(temperature=float(T[c]), prior_temperature=float(Tp[c]), seed=70 + c))
            ctx.set_state(c, sts[c])
        ctx.run(50)
        posts = [ctx.posterior(c) for c in range(n)]
        slot = list(range(n))
        d = PhiloxDraws(1234)
        d.begin_step(7)
        acc = CH.swap_chains([p["ll"] for p in posts], [p["prior"] for p in posts], T, Tp, slot, d, attempts=100)
        sos, n_acc = ctx.swap(100, False, 1234, 7)
        assert n_acc == len(acc) and sos.tolist() == slot
        # the chain at slot c now runs the state that was swapped in
        for c in range(n):
            assert parity.rel(ctx.posterior(c)["ll"], posts[slot[c]]["ll"]) == 0.0
        ctx.run(20)
    finally:
        ctx.close()


def test_philox_lockstep_six_components():
    """more than 4 mixture components: the wide instantiation of the step kernel."""
    from test_hostsim_parity import _many_confounders
    prob = _many_confounders(3)
    cfgs = [ChainConfig(seed=7), ChainConfig(temperature=1.15, prior_temperature=1.05, seed=8)]
    st0 = [synth.initial_state(prob, seed=40 + i, objects_per_cluster=4) for i in range(2)]
    parity.philox_lockstep(prob, st0, cfgs, 200, team=2)


@pytest.mark.parametrize("team", [1, 0])
def test_philox_lockstep_c4_shape(team):
    """BASELINE config 4 shape (5000 x 200 x 6, K=5, 20 families): the replicas do not fit in
    shared memory (private HBM slices) and, at team 1, one CTA walks a whole chain -- the
    placement the 128-chains-per-GPU replicate runs use.  Lockstep with the fp64 oracle."""
    prob, _ = synth.make_problem(N=5000, F=200, S=6, K=5, n_fam=20, seed=4)
    cfgs = [ChainConfig(seed=300), ChainConfig(temperature=1.1, prior_temperature=1.05, seed=301)]
    st0 = [synth.initial_state(prob, seed=100 + c) for c in range(2)]
    parity.philox_lockstep(prob, st0, cfgs, 40, team=team)


def test_replicate_chains_are_independent_of_batching():
    """independent replicates (config 4): a chain's trajectory depends on its seed only, not on
    how many other chains share the launch or on the team size"""
    prob, _ = synth.make_problem(N=300, F=40, S=5, K=4, n_fam=6, seed=8)
    sts = [synth.initial_state(prob, seed=c) for c in range(6)]

    def run(chains, team):
        ctx = DeviceContext(prob, len(chains), team=team)
        try:
            for i, c in enumerate(chains):
                ctx.configure(i, ChainConfig(seed=1000 + c))
                ctx.set_state(i, sts[c])
            ctx.run(300)
            return {c: (ctx.get_state(i), ctx.posterior(i)) for i, c in enumerate(chains)}
        finally:
            ctx.close()
    a, b = run([0, 1, 2, 3, 4, 5], 1), run([4, 2], 4)
    for c in (4, 2):
        assert np.array_equal(a[c][0].cluster_of, b[c][0].cluster_of) and np.array_equal(a[c][0].src, b[c][0].src)
        assert np.array_equal(a[c][0].counts, b[c][0].counts) and np.array_equal(a[c][0].W, b[c][0].W)
        assert parity.rel(a[c][1]["ll"], b[c][1]["ll"]) < 1e-12 and parity.rel(a[c][1]["prior"], b[c][1]["prior"]) < 1e-12


def test_propose_only_operator_calls():
    """SB_CFG_PROPOSE_ONLY (operator API of the reference, operators.py:269-276): one step of a
    forced operator commits its proposal without the MH test; state, log_q and log_q_back equal
    the oracle's proposal on the same draws."""
    from dataclasses import replace
    from oracle import operators as O
    from oracle.draws import PhiloxDraws
    from sbayes_b200.packed import N_OPS

    class OneHot(ChainConfig):
        def op_weight_vector(self, n_clusters):
            return self._w
    prob, _ = synth.make_problem(N=120, F=24, S=4, K=3, n_fam=4, seed=7, max_size=40)
    st = synth.initial_state(prob, seed=5, objects_per_cluster=8)
    st.counts = M.rebuild_counts(prob, st)
    ctx = DeviceContext(prob, 1, team=2)
    try:
        step, changed = 0, 0
        for rnd in range(3):
            for op in range(N_OPS):
                cfg = OneHot(seed=99)
                cfg._w = np.eye(N_OPS)[op]
                cfg.propose_only = True
                ctx.configure(0, cfg)
                ctx.set_state(0, st, i_step=step)
                rec = ctx.run(1, records=True)[0, 0]
                step += 1
                draws = PhiloxDraws(99)
                draws.begin_step(step)
                prop = O.propose(prob, st, ChainConfig(seed=99), draws, op)
                assert rec["op"] == op
                if not prop.changed or prop.log_q_back == O.NEG_INF:
                    assert not rec["accept"]
                    continue
                assert rec["accept"], "propose-only must commit every proposal the operator does not reject"
                assert parity.rel(prop.log_q, rec["log_q"]) < 1e-10 and parity.rel(prop.log_q_back, rec["log_q_back"]) < 1e-10
                got = ctx.get_state(0)
                assert np.array_equal(got.cluster_of, prop.state.cluster_of) and np.array_equal(got.src, prop.state.src)
                assert np.array_equal(got.counts, prop.state.counts) and np.array_equal(got.W, prop.state.W)
                st = got
                changed += 1
        assert changed >= 15
    finally:
        ctx.close()
