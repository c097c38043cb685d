"""EM-style cluster initializer on the device (`sb_init_em`, SbayesInitializer.generate_clusters_em,
sbayes/sampling/initializers.py:93-169) against (a) a float64 numpy restatement of the loop and (b) the LIVE
reference's own method on the same random start (`ref`): the reference is float32 with numpy's summation orders,
so the soft assignment agrees to 2e-4 absolute and the discretised clusters are identical."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
import build as HB  # noqa: E402
from oracle import refharness as H  # noqa: E402
from sbayes_b200 import synth, initializer as INI  # noqa: E402
from sbayes_b200.backend import DeviceContext  # noqa: E402
from sbayes_b200.packed import NA  # noqa: E402

BACKENDS = ["emu", pytest.param("cuda", marks=pytest.mark.gpu)]


def _lib(backend):
    return None if backend == "cuda" else HB.load()


def em_numpy(prob, z0, n_steps):
    """float64 restatement of initializers.py:122-155 (one-hot contractions as segmented sums / gathers)"""
    N, F, S = prob.N, prob.F, prob.S
    avail = INI.groups_available(prob)
    X = np.zeros((N, F, S))
    valid = prob.x != NA
    o, f = np.nonzero(valid)
    X[o, f, prob.x[o, f]] = 1.0
    z = z0.astype(np.float64)
    for i in range(n_steps):
        counts = np.einsum("ij,jkl->ikl", z, X).astype(np.float32).astype(np.float64) + 0.5 * prob.applicable
        p = (counts / counts.sum(axis=-1, keepdims=True)).astype(np.float32).astype(np.float64)
        with np.errstate(divide="ignore"):
            logp = np.log(p).astype(np.float32).astype(np.float64)
        ll = np.zeros((z.shape[0], N))
        xs = np.where(valid, prob.x, 0)
        for g in range(z.shape[0]):
            ll[g] = np.where(valid, logp[g][np.arange(F)[None, :], xs], 0.0).sum(axis=1)
        T = (n_steps / (1 + i)) ** 3
        lh = np.where(avail, (ll.astype(np.float32) / np.float32(T)).astype(np.float64), -np.inf)
        e = np.exp(lh - lh.max(axis=0, keepdims=True))
        z = (e / e.sum(axis=0, keepdims=True)).astype(np.float32).astype(np.float64)
    return z.astype(np.float32)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("shape", [(60, 10, 3, 2, 3), (300, 45, 5, 4, 7)])
def test_em_matches_numpy_restatement(backend, shape):
    N, F, S, K, nf = shape
    prob, _ = synth.make_problem(N=N, F=F, S=S, K=K, n_fam=nf, seed=3, jagged_states=True)
    z0 = INI.initial_soft_assignment(prob, np.random.default_rng(5))
    ctx = DeviceContext(prob, 1, _lib=_lib(backend))
    try:
        z = ctx.init_em(z0, 8)
    finally:
        ctx.close()
    ref = em_numpy(prob, z0, 8)
    assert np.all(z[~INI.groups_available(prob)] == 0)
    assert np.allclose(z.sum(axis=0), 1.0, atol=1e-5)
    assert np.abs(z - ref).max() < 2e-5


@pytest.mark.gpu
def test_em_c4_shape_properties():
    prob, _ = synth.make_problem(N=5000, F=200, S=6, K=5, n_fam=20, seed=4)
    z0 = INI.initial_soft_assignment(prob, np.random.default_rng(1))
    ctx = DeviceContext(prob, 1)
    try:
        z, ms = ctx.init_em(z0, 50, timing=True)
    finally:
        ctx.close()
    assert np.all(np.isfinite(z)) and np.allclose(z.sum(axis=0), 1.0, atol=1e-5) and np.all(z[~INI.groups_available(prob)] == 0)
    assert np.abs(z - em_numpy(prob, z0, 50)).max() < 1e-4


@pytest.mark.ref
@pytest.mark.parametrize("backend", BACKENDS)
def test_em_matches_live_reference(backend):
    """the reference's own generate_clusters_em on the same random start: same soft assignment (float32 tolerance),
    same discretised clusters; and the rebound method runs inside SbayesInitializer.generate_sample"""
    from collections import OrderedDict
    H.ref()                                                   # puts the reference (and its stub modules) on sys.path
    import sbayes.sampling.initializers as ini
    prob0, truth = synth.make_problem(N=80, F=14, S=4, K=3, n_fam=4, seed=11, max_size=30)
    confs = OrderedDict(universal=np.ones((1, prob0.N), bool), family=np.stack([prob0.conf_gid[1] == g for g in range(4)]))
    prior_cfg = dict(objects_per_cluster=dict(type="uniform_area", min=prob0.min_size, max=prob0.max_size), geo=dict(type="uniform"),
                     weights=dict(type="uniform"), cluster_effect=dict(type="uniform"),
                     confounding_effects=dict(universal={"<ALL>": dict(type="uniform")}, family={"<DEFAULT>": dict(type="uniform")}))
    data, model = H.build_reference_model(prob0.x, prob0.applicable, truth["locations"], confs, prob0.K, prior_cfg)
    init = ini.SbayesInitializer(model=model, data=data, initial_size=8, attempts=1, initial_cluster_steps=False, n_em_steps=20)
    # capture the reference's final z by wrapping its discretisation
    seen = {}
    orig_disc = ini.SbayesInitializer.discretize_fuzzy_cluster_2

    def spy(self, z, total_size):
        seen.setdefault("z", []).append(np.array(z))
        seen.setdefault("size", []).append(total_size)
        return orig_disc(self, z, total_size)
    ini.SbayesInitializer.discretize_fuzzy_cluster_2 = spy
    try:
        H.seed_reference(42)
        clusters_ref = init.generate_clusters_em()
        H.seed_reference(42)
        clusters_dev = INI.generate_clusters_em(init, _lib=_lib(backend))
    finally:
        ini.SbayesInitializer.discretize_fuzzy_cluster_2 = orig_disc
    z_ref, z_dev = seen["z"]
    assert seen["size"][0] == seen["size"][1]
    assert np.abs(z_ref - z_dev).max() < 2e-4
    assert np.array_equal(clusters_ref, clusters_dev)
    # installed: SbayesInitializer.generate_sample runs with the device EM inside
    orig = INI.install_initializer(_lib=_lib(backend))
    try:
        H.seed_reference(7)
        s = init.generate_sample(c=0)
        assert s.clusters.value.shape == (prob0.K, prob0.N) and np.all(s.clusters.value.sum(axis=1) >= model.min_size)
        assert np.isfinite(model.likelihood(s, caching=False))
    finally:
        ini.SbayesInitializer.generate_clusters_em = orig
