"""Stand-alone evaluation kernels (sb_eval.cuh: statistics pass + terms, sb_eval_batch) against
the fp64 oracle and the golden fixtures: count tensors bit-exact, log-likelihood and prior terms
within 1e-10 relative (BASELINE.json north_star).  Emulated here, on the B200 under `-m gpu`."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
import build as HB  # noqa: E402
import parity  # noqa: E402
from oracle import fixtures as FX, model as M  # noqa: E402
from sbayes_b200 import synth  # noqa: E402
from sbayes_b200.backend import DeviceContext  # noqa: E402
from sbayes_b200.packed import ChainConfig, SIZE_UNIFORM_SIZE, SIZE_QUADRATIC  # noqa: E402

TOL = parity.TOL_ORACLE      # 1e-10 relative


def check_batch(prob, states, lib=None, cfgs=None, want_pass=None, geometry=None):
    ctx = DeviceContext(prob, len(states), _lib=lib)
    try:
        for c, st in enumerate(states):
            ctx.configure(c, cfgs[c] if cfgs else ChainConfig(seed=c))
            ctx.set_state(c, st)
        if want_pass:
            info = ctx.eval_info()
            assert info["pass"] == want_pass, info
            if geometry:
                assert (info["n_ftiles"], info["n_oblk"]) == geometry, info
        res = ctx.eval_batch()
        for c, st in enumerate(states):
            ref = st.copy()
            ref.counts = M.rebuild_counts(prob, st)
            got = ctx.get_state(c)
            assert np.array_equal(got.counts, ref.counts), "count tensors must be bit-exact"
            assert parity.rel(M.log_likelihood(prob, ref), res[c]["ll"]) < TOL
            assert parity.rel(M.log_prior(prob, ref), res[c]["prior"]) < TOL
            for k, v in M.prior_parts(prob, ref).items():
                assert parity.rel(v, res[c][k + "_prior"]) < TOL, k
            # the persistent kernel's own rebuild agrees to rounding
            ev = ctx.eval(c)
            assert parity.rel(ev["ll"], res[c]["ll"]) < 1e-12 and parity.rel(ev["prior"], res[c]["prior"]) < 1e-12
        return res
    finally:
        ctx.close()


def small_problem(**kw):
    args = dict(N=70, F=37, S=4, K=3, n_fam=4, seed=7, jagged_states=True, max_size=20)
    args.update(kw)
    prob, _ = synth.make_problem(**args)
    return prob, [synth.initial_state(prob, seed=5 + i, objects_per_cluster=5) for i in range(3)]


# ------------------------------------------------------------------ emulated (GPU-less container)
@pytest.fixture(scope="module")
def lib():
    return HB.load()


@pytest.fixture
def env():
    saved = dict(os.environ)
    yield os.environ
    os.environ.clear()
    os.environ.update(saved)


def test_emulated_eval_batch(lib):
    prob, sts = small_problem()
    check_batch(prob, sts, lib, want_pass="fine", geometry=(2, 1))      # one increment per observation, plain-store flush


def test_emulated_eval_fine_tiles_and_object_blocks(lib, env):
    """fine-row pass with several feature tiles (the last one narrower) and several object blocks (atomic flush)"""
    prob, sts = small_problem(F=70)
    env["SB_HOSTSIM_FINE_FT"] = "32"
    check_batch(prob, sts, lib, want_pass="fine", geometry=(3, 1))
    env["SB_HOSTSIM_EVAL_OBLK"] = "3"
    check_batch(prob, sts, lib, want_pass="fine", geometry=(3, 3))


def test_emulated_eval_two_increment_pass(lib, env):
    """the two-increment histogram (what problems with more than 254 fine rows fall back to)"""
    prob, sts = small_problem()
    env["SB_EVAL_NO_FINE"] = "1"
    check_batch(prob, sts, lib, want_pass="hist")


def test_emulated_eval_feature_tiles_and_object_blocks(lib, env):
    prob, sts = small_problem(F=70)
    env["SB_HOSTSIM_EVAL_SMEM"] = str(16 * (prob.G_tot * prob.S + 2 * 2 * prob.C) * 4 + prob.K * 4 + 64 + 200)   # 16-feature tiles
    env["SB_HOSTSIM_EVAL_OBLK"] = "3"
    env["SB_EVAL_NO_FINE"] = "1"
    check_batch(prob, sts, lib, want_pass="hist")


def test_emulated_eval_priors(lib):
    prob, sts = small_problem(geo_cost_based=True)
    prob.size_prior = SIZE_UNIFORM_SIZE
    prob.weights_prior_uniform = False
    prob.alpha_w = 1.0 + np.random.default_rng(0).random((prob.F, prob.C))
    check_batch(prob, sts, lib)
    prob.size_prior = SIZE_QUADRATIC
    check_batch(prob, sts, lib)


def test_emulated_eval_sample_from_prior(lib):
    prob, sts = small_problem()
    res = check_batch(prob, sts[:1], lib)
    ctx = DeviceContext(prob, 1, _lib=lib)
    try:
        ctx.configure(0, ChainConfig(seed=0, sample_from_prior=True))
        ctx.set_state(0, sts[0])
        r = ctx.eval_batch()[0]
        assert r["ll"] == 0.0 and parity.rel(res[0]["prior"], r["prior"]) < 1e-15
    finally:
        ctx.close()


@pytest.mark.parametrize("case", ["c1_testfiles", "c3_south_america", "synth_geo_sigmoid"])
def test_emulated_eval_golden(lib, case):
    """initial state of a fixture recorded from the live reference: the reference's own counts,
    likelihood and prior"""
    fx = FX.load_replay(os.path.join(parity.GOLDEN, case + ".npz"))
    prob, raw = fx["prob"], fx["raw"]
    ctx = DeviceContext(prob, 1, _lib=lib)
    try:
        ctx.configure(0, fx["cfg"])
        ctx.set_state(0, fx["st0"])
        r = ctx.eval_batch()[0]
        assert np.array_equal(ctx.get_state(0).counts, fx["st0"].counts)
        assert parity.rel(float(raw["ref_ll0"]), r["ll"]) < parity.TOL_REF_LL
        assert parity.rel(float(raw["ref_prior0"]), r["prior"]) < parity.TOL_REF_PRIOR
    finally:
        ctx.close()


# ------------------------------------------------------------------ B200
@pytest.mark.gpu
def test_gpu_eval_batch_small():
    prob, sts = small_problem()
    check_batch(prob, sts)
    prob, sts = small_problem(geo_cost_based=True)
    prob.size_prior = SIZE_UNIFORM_SIZE
    prob.weights_prior_uniform = False
    prob.alpha_w = 1.0 + np.random.default_rng(0).random((prob.F, prob.C))
    check_batch(prob, sts)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["c1_testfiles", "c3_south_america", "synth_geo", "synth_c2_width"])
def test_gpu_eval_golden(case):
    fx = FX.load_replay(os.path.join(parity.GOLDEN, case + ".npz"))
    prob, raw = fx["prob"], fx["raw"]
    ctx = DeviceContext(prob, 1)
    try:
        ctx.configure(0, fx["cfg"])
        ctx.set_state(0, fx["st0"])
        r = ctx.eval_batch()[0]
        assert np.array_equal(ctx.get_state(0).counts, fx["st0"].counts)
        assert parity.rel(float(raw["ref_ll0"]), r["ll"]) < parity.TOL_REF_LL
        assert parity.rel(float(raw["ref_prior0"]), r["prior"]) < parity.TOL_REF_PRIOR
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_eval_c2_against_oracle():
    """BASELINE config 2 (1000 x 100 x 4, K=5, 4 chains): oracle-checked at full size"""
    prob, _ = synth.make_problem(N=1000, F=100, S=4, K=5, n_fam=8, seed=2)
    sts = [synth.initial_state(prob, seed=1000 + c) for c in range(4)]
    check_batch(prob, sts)


@pytest.mark.gpu
def test_gpu_eval_c4_shape_properties():
    """BASELINE config 4 shape (5000 x 200 x 6, several feature tiles / object blocks): numpy
    count rebuild bit-exact, histogram invariants, agreement with the step kernel's rebuild"""
    prob, _ = synth.make_problem(N=5000, F=200, S=6, K=5, n_fam=20, seed=4)
    sts = [synth.initial_state(prob, seed=100 + c) for c in range(3)]
    ctx = DeviceContext(prob, 3)
    try:
        for c, st in enumerate(sts):
            ctx.configure(c, ChainConfig(seed=c))
            ctx.set_state(c, st)
        res = ctx.eval_batch()
        valid = int((prob.x != 255).sum())
        for c, st in enumerate(sts):
            got = ctx.get_state(c)
            assert np.array_equal(got.counts, M.rebuild_counts(prob, st))
            # every non-missing observation is counted exactly once (universal covers all objects)
            assert int(got.counts.sum()) == valid
            ev = ctx.eval(c)
            assert parity.rel(ev["ll"], res[c]["ll"]) < 1e-12 and parity.rel(ev["prior"], res[c]["prior"]) < 1e-12
        ref = sts[0].copy(); ref.counts = M.rebuild_counts(prob, sts[0])
        assert parity.rel(M.log_likelihood(prob, ref), res[0]["ll"]) < TOL
        assert parity.rel(M.log_prior(prob, ref), res[0]["prior"]) < TOL
    finally:
        ctx.close()


# ------------------------------------------------------------------ tcgen05 statistics pass
def _mma_vs_hist(prob, states):
    ctx = DeviceContext(prob, len(states))
    try:
        for c, st in enumerate(states):
            ctx.configure(c, ChainConfig(seed=c))
            ctx.set_state(c, st)
        ctx.eval_set_mode("hist")
        ref = ctx.eval_batch()
        ptr, n, stride, lay = ctx.stats_buffer()
        hist_stats = [ctx.get_state(c).counts.copy() for c in range(len(states))]
        ctx.eval_set_mode("mma")
        got = ctx.eval_batch()
        for c, st in enumerate(states):
            counts = ctx.get_state(c).counts
            assert np.array_equal(counts, M.rebuild_counts(prob, st)), "tcgen05 counts must be bit-exact"
            assert np.array_equal(counts, hist_stats[c])
            # identical integer statistics -> identical posterior terms
            for k in ("ll", "prior", "source_prior", "size_prior", "weights_prior"):
                assert got[c][k] == ref[c][k], k
            one, _ = ctx.counts_rebuild(c, mma=True)
            assert np.array_equal(one, counts)
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(70, 37, 4, 3, 4), (5, 2, 2, 2, 2), (100, 36, 5, 3, 6), (1000, 100, 4, 5, 8),
                                   (777, 130, 7, 9, 30)])
def test_gpu_mma_statistics_small(shape):
    N, F, S, K, nf = shape
    prob, _ = synth.make_problem(N=N, F=F, S=S, K=K, n_fam=nf, seed=11, min_size=1 if N < 10 else 3,
                                 jagged_states=N > 10)
    sts = [synth.initial_state(prob, seed=5 + i, objects_per_cluster=1 if N < 10 else 5) for i in range(2)]
    _mma_vs_hist(prob, sts)


@pytest.mark.gpu
def test_gpu_mma_statistics_c4_shape():
    prob, _ = synth.make_problem(N=5000, F=200, S=6, K=5, n_fam=20, seed=4)
    _mma_vs_hist(prob, [synth.initial_state(prob, seed=100 + c) for c in range(2)])


@pytest.mark.gpu
def test_gpu_mma_statistics_wide_groups():
    """many group columns (N tile of 96) and few features per CTA, the config-5 regime"""
    prob, _ = synth.make_problem(N=3000, F=150, S=8, K=16, n_fam=64, seed=5)
    _mma_vs_hist(prob, [synth.initial_state(prob, seed=3)])


# ------------------------------------------------------------------ observation likelihood
def check_obs(prob, states, lib=None):
    ctx = DeviceContext(prob, len(states), _lib=lib)
    try:
        for c, st in enumerate(states):
            ctx.configure(c, ChainConfig(seed=c))
            ctx.set_state(c, st)
        lh, ob, sums, _ = ctx.obs_loglik()
        for c, st in enumerate(states):
            ref = st.copy()
            ref.counts = M.rebuild_counts(prob, st)
            want = M.observation_likelihood(prob, ref)
            assert np.array_equal(lh[c], want.astype(np.float32)) or np.max(np.abs(lh[c] - want) / want) < 1.2e-7   # float32 output
            assert np.max(np.abs(ob[c] - np.log(want).sum(axis=1)) / np.maximum(1.0, np.abs(ob[c]))) < TOL
            assert parity.rel(float(np.log(want).sum()), float(sums[c])) < TOL
        return lh, sums
    finally:
        ctx.close()


def test_emulated_observation_likelihood(lib, env):
    prob, sts = small_problem()
    check_obs(prob, sts, lib)
    prob, sts = small_problem(F=70, N=200)
    check_obs(prob, sts[:2], lib)


@pytest.mark.parametrize("case", ["obs_lik_small", "obs_lik_jagged"])
def test_emulated_observation_likelihood_golden(lib, case):
    _obs_golden(case, lib)


def _obs_golden(case, lib=None):
    """vectors computed by the live reference (tests/golden/make_obs_golden.py); the reference
    rounds tables and weights to float32, so 2e-6 relative"""
    d = np.load(os.path.join(parity.GOLDEN, case + ".npz"), allow_pickle=False)
    prob, st = FX.problem_from_dict(d), FX.state_from_dict(d, "st_")
    lh, sums = check_obs(prob, [st], lib)
    assert np.max(np.abs(lh[0] - d["ref_lh"]) / d["ref_lh"]) < 2e-6
    assert abs(sums[0] - np.log(d["ref_lh"]).sum()) < 2e-6 * abs(sums[0])


@pytest.mark.gpu
def test_gpu_observation_likelihood():
    for case in ("obs_lik_small", "obs_lik_jagged"):
        _obs_golden(case)
    prob, sts = small_problem(F=70, N=200)
    check_obs(prob, sts[:2])
    prob, _ = synth.make_problem(N=1000, F=100, S=4, K=5, n_fam=8, seed=2)
    check_obs(prob, [synth.initial_state(prob, seed=1000 + c) for c in range(2)])
    prob, _ = synth.make_problem(N=5000, F=200, S=6, K=5, n_fam=20, seed=4)
    check_obs(prob, [synth.initial_state(prob, seed=100)])


def test_emulated_eval_and_obs_six_components(lib):
    """more than four mixture components: the generic (non-packed) paths of the statistics pass
    and of the observation-likelihood kernel"""
    from test_hostsim_parity import _many_confounders
    prob = _many_confounders(3)
    assert prob.C == 6
    sts = [synth.initial_state(prob, seed=40 + i, objects_per_cluster=4) for i in range(2)]
    check_batch(prob, sts, lib)
    check_obs(prob, sts, lib)


@pytest.mark.gpu
def test_gpu_eval_and_obs_six_components():
    from test_hostsim_parity import _many_confounders
    prob = _many_confounders(3)
    sts = [synth.initial_state(prob, seed=40 + i, objects_per_cluster=4) for i in range(2)]
    check_batch(prob, sts)
    check_obs(prob, sts)
    _mma_vs_hist(prob, sts)
