"""Known-answer tests: the reference's own doctest values (sbayes/util.py) and the
survey's hand-checked Dirichlet-categorical values against the oracle."""
import math

import numpy as np

from oracle import model as M, operators as O, philox


def test_dircat_kat():
    # counts [2,1,0,0], a [1,1,0,0] -> log(1/12); counts [0,3,1,0], a [1,1,1,0] -> log(1/60)
    v = M.dircat_logpdf(np.array([[2, 1, 0, 0], [0, 3, 1, 0]]), np.array([[1., 1., 0., 0.], [1., 1., 1., 0.]]))
    assert np.allclose(v, [math.log(1 / 12), math.log(1 / 60)], rtol=1e-13)
    assert abs(v[0] - (-2.4849066)) < 1e-7 and abs(v[1] - (-4.0943446)) < 1e-7


def test_log_multinom_kat():
    # sbayes/util.py:1137-1143
    assert abs(M.log_multinom(5, [1, 1, 1, 1]) - 4.787491742782046) < 1e-12
    assert abs(M.log_multinom(13, [4]) - 6.572282542694008) < 1e-12
    assert abs(M.log_multinom(13, [3, 2]) - 9.462654300590172) < 1e-12
    # permutation invariance (test/test_model.py:288-306, test/test_util.py:11-49)
    assert abs(M.log_multinom(20, [3, 5, 2]) - M.log_multinom(20, [5, 2, 3])) < 1e-12
    assert abs(M.log_multinom(10, [3]) - math.log(math.comb(10, 3))) < 1e-12


def test_heat_binary_probability_kat():
    # sbayes/util.py:1461-1465
    assert O.heat_binary_probability(0.5, 2) == 0.5
    assert abs(O.heat_binary_probability(1 / 3, 0.5) - 0.2) < 1e-12


def test_philox_kat():
    # Random123 known-answer vectors for philox4x32-10
    out = philox.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(v) for v in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = philox.philox4x32_10(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(v) for v in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = philox.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(v) for v in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_mst_edge_costs():
    cost = np.array([[0, 1, 4, 3], [1, 0, 2, 5], [4, 2, 0, 6], [3, 5, 6, 0]], dtype=float)
    assert sorted(M.mst_edge_costs(cost).tolist()) == [1.0, 2.0, 3.0]
    assert M.mst_edge_costs(np.zeros((1, 1))).tolist() == [0.0]
    assert M.mst_edge_costs(np.zeros((3, 3))).tolist() == [0.0]
