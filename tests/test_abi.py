"""The C-ABI library loads and exports every symbol include/sbayes_b200.h declares (no
compute calls: there is no GPU in the build container)."""
import ctypes
import os
import re

from sbayes_b200 import backend, build


def declared_symbols():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "sbayes_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_the_header():
    path = build.build()
    lib = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/sbayes_b200.h but not exported"
    assert sorted(backend.EXPORTS) == syms


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(backend._ChainCfg) == 8 * 2 + 4 * 2 + 8 * 10 + 8 + 8
    assert backend.RECORD_DTYPE.itemsize == 4 * 4 + 8 * 6
    assert ctypes.sizeof(backend._Posterior) == 48


def test_create_without_gpu_fails_loudly():
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sbayes_b200 import synth
    prob, _ = synth.make_problem(N=20, F=4, S=3, K=2, n_fam=2, seed=0, with_adjacency=False)
    with pytest.raises(backend.SbError):
        backend.DeviceContext(prob, 1)
