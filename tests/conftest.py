import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs the live reference at /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    from oracle import refharness
    if refharness.reference_available():
        return
    skip = pytest.mark.skip(reason="live reference not present")
    for item in items:
        if "ref" in item.keywords:
            item.add_marker(skip)


def golden_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith(("obs_lik_", "mc3_swap")))
