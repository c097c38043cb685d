"""The oracle replays every committed golden fixture (draw tapes and outputs recorded from
the live reference, tests/golden/make_golden.py) and must reproduce the reference's
proposals, accept decisions, posterior values and final state."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases
from oracle import fixtures as FX, chain as CH, model as M
from oracle.draws import TapeDraws

TOL_LOGQ, TOL_LL, TOL_PRIOR = 2e-6, 5e-7, 5e-7     # reference is float32 at these points


def rel(a, b):
    if not (np.isfinite(a) and np.isfinite(b)):
        return 0.0 if a == b else np.inf
    return abs(a - b) / max(1.0, abs(a))


@pytest.mark.parametrize("case", golden_cases())
def test_replay_matches_reference(case):
    fx = FX.load_replay(os.path.join(GOLDEN, case + ".npz"))
    prob, cfg, outs = fx["prob"], fx["cfg"], fx["outs"]
    st = fx["st0"].copy()
    # integer sufficient statistics: bit-exact against the reference's counts
    assert np.array_equal(M.rebuild_counts(prob, st), fx["st0"].counts)
    oc = CH.OracleChain(prob, cfg, st, draws=TapeDraws())
    assert rel(float(fx["raw"]["ref_ll0"]), oc.ll) < TOL_LL
    assert rel(float(fx["raw"]["ref_prior0"]), oc.prior) < TOL_PRIOR
    for i, tape in enumerate(fx["tapes"]):
        r = oc.step(tape)
        assert r.op == int(outs["op"][i])
        assert r.accept == bool(outs["accept"][i]), f"step {i}"
        assert rel(outs["log_q"][i], r.log_q) < TOL_LOGQ
        assert rel(outs["log_q_back"][i], r.log_q_back) < TOL_LOGQ
        assert rel(outs["ll_new"][i], r.ll_new) < TOL_LL
        assert rel(outs["prior_new"][i], r.prior_new) < TOL_PRIOR
    fin = fx["st1"]
    assert np.array_equal(oc.state.cluster_of, fin.cluster_of)
    assert np.array_equal(oc.state.src, fin.src)
    assert np.array_equal(oc.state.counts, fin.counts)
    assert np.array_equal(oc.state.W, fin.W)
    # incremental counts == bulk rebuild (the reference's verify_counts, operators.py:2018-2025)
    assert np.array_equal(M.rebuild_counts(prob, oc.state), oc.state.counts)
