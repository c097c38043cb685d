"""MC3 swap (sbayes/mcmc_setup.py:365-418) pinned against the LIVE reference:
`tests/golden/mc3_swap.npz` holds, per round, the pairs and uniforms the reference drew, the
scalars its model returned and the swaps it accepted (tests/golden/make_swap_golden.py).
Replayed through (a) the oracle's `swap_chains` and (b) the library's one swap implementation
(`sb_swap_scalars`, also what `sb_swap` and the multi-rank ladder of mc3.py call)."""
import os
import sys

import numpy as np
import pytest

from oracle import chain as CH, fixtures as FX, model as M
from oracle.draws import PhiloxDraws
from sbayes_b200 import backend
from sbayes_b200.packed import ChainState, ChainConfig

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mc3_swap.npz")


class _Tape:
    def __init__(self, pairs, u):
        self.pairs, self.u = pairs, u

    def swap_pairs(self, pairs, attempts):
        return [tuple(int(v) for v in p) for p in self.pairs]

    def swap_uniform(self, a):
        return float(self.u[a])


def _rounds():
    d = np.load(GOLDEN)
    out = []
    for i in range(int(d["n_rounds"])):
        out.append({k: d[f"r{i}_{k}"] for k in ("ll", "prior", "T", "Tp", "pairs", "u", "perm", "n_acc", "attempts", "neigh", "accepted")})
    return d, out


def _hostsim_lib():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    import build as HB
    return HB.load()


def test_oracle_swap_replays_reference():
    d, rounds = _rounds()
    n = int(d["n_chains"])
    for r in rounds:
        slot = list(range(n))
        acc = CH.swap_chains(r["ll"], r["prior"], r["T"], r["Tp"], slot, _Tape(r["pairs"], r["u"]), int(r["attempts"]), bool(r["neigh"]))
        assert slot == r["perm"].tolist() and len(acc) == int(r["n_acc"])
        assert np.array_equal(np.array(acc, np.int32).reshape(-1, 2), r["accepted"])


def _library_replay(lib):
    d, rounds = _rounds()
    n = int(d["n_chains"])
    total = 0
    for r in rounds:
        slot = list(range(n))
        acc = backend.swap_scalars(r["ll"], r["prior"], r["T"], r["Tp"], slot, seed=0, round_=0, attempts=int(r["attempts"]),
                                   only_neighbours=bool(r["neigh"]), tape_pairs=r["pairs"], tape_u=r["u"], _lib=lib)
        assert slot == r["perm"].tolist(), "library swap != reference swap"
        assert np.array_equal(np.array(acc, np.int32).reshape(-1, 2), r["accepted"])
        total += len(acc)
        # Philox mode: the library and the oracle take the same decisions from the same stream
        s1, s2 = list(range(n)), list(range(n))
        a1 = backend.swap_scalars(r["ll"], r["prior"], r["T"], r["Tp"], s1, seed=77, round_=5, attempts=int(r["attempts"]),
                                  only_neighbours=bool(r["neigh"]), _lib=lib)
        dr = PhiloxDraws(77)
        dr.begin_step(5)
        a2 = CH.swap_chains(r["ll"], r["prior"], r["T"], r["Tp"], s2, dr, int(r["attempts"]), bool(r["neigh"]))
        assert s1 == s2 and a1 == [tuple(p) for p in a2]
    assert total > 0


def test_library_swap_replays_reference_hostsim():
    _library_replay(_hostsim_lib())


@pytest.mark.gpu
def test_library_swap_replays_reference_cuda():
    _library_replay(None)


@pytest.mark.gpu
def test_device_scalars_take_the_reference_swaps():
    """The fixture's states evaluated ON THE DEVICE give scalars within 5e-7 of the reference's
    and, fed to the swap with the reference's recorded draws, the reference's decisions (a
    decision may only differ where log u and the ratio are closer than the scalar tolerance)."""
    from sbayes_b200.backend import DeviceContext
    d, rounds = _rounds()
    n = int(d["n_chains"])
    prob = FX.problem_from_dict(d)
    ctx = DeviceContext(prob, n)
    try:
        for c in range(n):
            ctx.configure(c, ChainConfig(seed=c))
            ctx.set_state(c, ChainState(d[f"s{c}_cluster_of"], d[f"s{c}_W"], d[f"s{c}_src"]))
        ev = ctx.eval_batch()
        ll0 = np.array([e["ll"] for e in ev]); pr0 = np.array([e["prior"] for e in ev])
        order = list(range(n))     # order[slot] = state index at that slot before the round
        for r in rounds:
            ll, pr = ll0[order], pr0[order]
            assert np.allclose(ll, r["ll"], rtol=5e-7, atol=0) and np.allclose(pr, r["prior"], rtol=5e-7, atol=0)
            slot = list(range(n))
            backend.swap_scalars(ll, pr, r["T"], r["Tp"], slot, seed=0, round_=0, attempts=int(r["attempts"]),
                                 only_neighbours=bool(r["neigh"]), tape_pairs=r["pairs"], tape_u=r["u"])
            assert slot == r["perm"].tolist()
            order = [order[p] for p in r["perm"]]
    finally:
        ctx.close()
