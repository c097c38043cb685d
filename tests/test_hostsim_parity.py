"""CPU-side check of the *device code's control flow*: the CUDA sources compiled for the fibre
emulator (tests/hostsim) replay the reference's recorded steps and run in lockstep with
the oracle.  The emulator is test infrastructure; the real parity tests are the `-m gpu`
ones, which call the CUDA library through the same C ABI."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
import build as HB  # noqa: E402
import parity  # noqa: E402
from sbayes_b200 import synth  # noqa: E402
from sbayes_b200.packed import ChainConfig  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    return HB.load()


@pytest.mark.parametrize("case,team,n", [("synth_small", 1, 150), ("synth_tempered", 2, 150), ("synth_geo", 2, 120),
                                         ("synth_geo_sigmoid", 1, 80), ("synth_tiny_all", 3, 200), ("c1_testfiles", 2, 200),
                                         ("c3_south_america", 3, 60), ("c3_south_america_hot", 1, 60)])
def test_emulated_replay(lib, case, team, n):
    parity.replay_golden(case, lib=lib, team=team, n_steps=n)


def test_emulated_philox_lockstep(lib):
    prob, _ = synth.make_problem(N=50, F=12, S=4, K=3, n_fam=4, seed=7, jagged_states=True, max_size=20)
    cfgs = [ChainConfig(seed=100), ChainConfig(temperature=1.2, prior_temperature=1.1, seed=101)]
    st0 = [synth.initial_state(prob, seed=5 + i, objects_per_cluster=5) for i in range(2)]
    parity.philox_lockstep(prob, st0, cfgs, 120, lib=lib, team=2)


def _many_confounders(n_conf_extra=3):
    """universal + family + extra partitions: C = 3 + n_conf_extra mixture components (the wide
    instantiation of the step kernel, more than 4 components)."""
    import numpy as np
    prob, _ = synth.make_problem(N=40, F=8, S=3, K=2, n_fam=3, seed=21, max_size=15)
    rng = np.random.default_rng(3)
    for i in range(n_conf_extra):
        gid = rng.integers(-1, 3, size=prob.N).astype(np.int32)
        for g in range(3):
            gid[g + 3 * i] = g
        prob.conf_gid.append(gid)
        prob.conf_ngroups.append(3)
        prob.alpha_conf.append(np.broadcast_to(prob.applicable.astype(np.float64), (3, prob.F, prob.S)).copy())
        prob.conf_names.append(f"extra{i}")
    prob.alpha_w = np.ones((prob.F, prob.C))
    return prob


def test_emulated_lockstep_six_components(lib):
    prob = _many_confounders(3)
    assert prob.C == 6
    cfgs = [ChainConfig(seed=7), ChainConfig(temperature=1.15, prior_temperature=1.05, seed=8)]
    st0 = [synth.initial_state(prob, seed=40 + i, objects_per_cluster=4) for i in range(2)]
    parity.philox_lockstep(prob, st0, cfgs, 100, lib=lib, team=2)


@pytest.fixture
def env():
    saved = dict(os.environ)
    yield os.environ
    os.environ.clear()
    os.environ.update(saved)


@pytest.mark.parametrize("smem,no_tables", [(24 * 1024, False), (24 * 1024, True), (200 * 1024, True)])
def test_emulated_lockstep_replicas_in_hbm_slice(lib, env, smem, no_tables):
    """the placement the large shapes (BASELINE config 4 / 5) run with: the group-indexed replica
    does not fit in shared memory and lives in the CTA's private HBM slice; with and without the
    per-combo pass tables"""
    env["SB_HOSTSIM_SMEM"] = str(smem)
    if no_tables:
        env["SB_NO_TABLES"] = "1"
    prob, _ = synth.make_problem(N=60, F=14, S=4, K=3, n_fam=4, seed=17, jagged_states=True, max_size=25)
    cfgs = [ChainConfig(seed=200), ChainConfig(temperature=1.25, prior_temperature=1.1, seed=201)]
    st0 = [synth.initial_state(prob, seed=15 + i, objects_per_cluster=6) for i in range(2)]
    parity.philox_lockstep(prob, st0, cfgs, 150, lib=lib, team=2)


class _WideHeavy(ChainConfig):
    """operator mix dominated by the wide cluster operator and the source operators (the paths
    whose per-item work is split over the team)"""

    def op_weight_vector(self, n_clusters):
        import numpy as np
        w = np.array([0.02, 0.02, 0.02, 0.02, 0.12, 0.4, 0.1, 0.12, 0.12, 0.06])
        return w / w.sum()


def test_emulated_lockstep_team_split_items(lib, env):
    """many flipped objects x features: the source resampling of the wide operator and the source
    Gibbs passes split their items over the team (threshold: 4 items per thread)"""
    env["SB_HOSTSIM_BLOCK"] = "32"
    prob, _ = synth.make_problem(N=70, F=40, S=4, K=3, n_fam=4, seed=23, max_size=30)
    cfgs = [_WideHeavy(seed=400), _WideHeavy(temperature=1.2, prior_temperature=1.1, seed=401)]
    st0 = [synth.initial_state(prob, seed=25 + i, objects_per_cluster=6) for i in range(2)]
    parity.philox_lockstep(prob, st0, cfgs, 120, lib=lib, team=3)


def test_emulated_lockstep_lane_per_object_pass(lib, env):
    """a CTA with at least 32 objects per warp walks the posterior pass with one lane per object
    (the large-N placement: config 4 with one CTA per chain)"""
    env["SB_HOSTSIM_BLOCK"] = "32"
    prob, _ = synth.make_problem(N=90, F=18, S=4, K=3, n_fam=4, seed=29, max_size=35, jagged_states=True)
    cfgs = [ChainConfig(seed=500), _WideHeavy(temperature=1.3, prior_temperature=1.15, seed=501)]
    st0 = [synth.initial_state(prob, seed=35 + i, objects_per_cluster=7) for i in range(2)]
    parity.philox_lockstep(prob, st0, cfgs, 150, lib=lib, team=2)
    env["SB_HOSTSIM_SMEM"] = str(24 * 1024)          # and with the tables in the HBM slice
    parity.philox_lockstep(prob, st0, cfgs, 80, lib=lib, team=1)


class _ClusterHeavy(ChainConfig):
    """mostly gibbsish grow / shrink, wide and jump steps: the operators with passes over many objects"""

    def op_weight_vector(self, n_clusters):
        import numpy as np
        w = np.array([0.02, 0.02, 0.02, 0.1, 0.3, 0.2, 0.2, 0.05, 0.05, 0.04])
        return w / w.sum()


@pytest.mark.parametrize("team,block,pairs", [(1, 32, True), (1, 64, False), (2, 32, True), (5, 32, False), (3, 64, True)])
def test_emulated_lockstep_tiled_passes(lib, env, team, block, pairs):
    """the pass tables do not fit in shared memory as a whole (BASELINE configs 4 / 5): objects in
    combo order, tables built tile by tile in shared memory, ratio form for AlterCluster, dual
    tables for the wide operator and the jump; one or several lanes per object"""
    import ctypes
    env["SB_HOSTSIM_TILED"] = "1"
    env["SB_HOSTSIM_BLOCK"] = str(block)
    if not pairs:
        env["SB_NO_PAIRS"] = "1"       # single-feature ratio tables (what problems with more than 15 states use)
    prob, _ = synth.make_problem(N=110, F=21, S=4, K=3, n_fam=5, seed=31, max_size=60, jagged_states=True)
    cfgs = [_ClusterHeavy(seed=700), _ClusterHeavy(temperature=1.3, prior_temperature=1.15, seed=701)]
    st0 = [synth.initial_state(prob, seed=45 + i, objects_per_cluster=25) for i in range(2)]
    lib.sb_hostsim_tiled_calls.restype = ctypes.c_long
    before = [lib.sb_hostsim_tiled_calls(k) for k in range(4)]
    parity.philox_lockstep(prob, st0, cfgs, 150, lib=lib, team=team)
    after = [lib.sb_hostsim_tiled_calls(k) for k in range(4)]
    ran = [a > b for a, b in zip(after, before)]
    assert ran[1] and ran[2] and ran[3 if pairs else 0] and not ran[0 if pairs else 3], (before, after)   # alter, wide and jump all ran tiled
