"""Import the live sBayes reference (from /root/reference) and convert between its objects
and the packed containers of `oracle/problem.py`.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Usable where /root/reference exists (the build
container) or where oracle/make_ref.py placed its copy (oracle/_ref, shipped to the GPU box).  Used by `tests/golden/make_golden.py` to generate the committed
fixtures and by the optional `-m ref` tests.  Recipe: SURVEY.md Appendix B.
"""
from __future__ import annotations

import os
import random
import sys
from collections import OrderedDict

import numpy as np

from . import problem as P

_HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = os.path.join(_HERE, "refstubs")
_ref = None


def _find_reference() -> str:
    """/root/reference in the build container; the unmodified copy made by oracle/make_ref.py
    (git-ignored oracle/_ref, travels to the GPU box) otherwise."""
    cands = [os.environ.get("SBAYES_REFERENCE"), "/root/reference", os.path.join(_HERE, "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "sbayes")):
            return c
    return "/root/reference"


REFERENCE_ROOT = _find_reference()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sbayes"))


def reference_paths() -> list:
    """sys.path entries that make `import sbayes` resolve to the reference (stubs first)."""
    return [_STUBS, REFERENCE_ROOT]


def ref():
    """Import the reference once and return a namespace of its modules."""
    global _ref
    if _ref is not None:
        return _ref
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, _STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    warnings.filterwarnings("ignore")
    import sbayes.model as model                      # must come first (circular imports)
    import sbayes.sampling.state as state
    import sbayes.sampling.operators as operators
    import sbayes.sampling.counts as counts
    import sbayes.sampling.conditionals as conditionals
    import sbayes.sampling.mcmc_chain as mcmc_chain
    import sbayes.load_data as load_data
    import sbayes.config.config as config
    import sbayes.util as util
    import sbayes.preprocessing as preprocessing
    import sbayes.mcmc_setup as mcmc_setup
    import sbayes.model.prior as prior
    import sbayes.model.likelihood as likelihood

    class NS:
        pass
    ns = NS()
    for k, v in list(locals().items()):
        if k not in ("ns", "NS", "p", "warnings"):
            setattr(ns, k, v)
    _ref = ns
    return ns


def patch_network_for_stubs():
    """With the `pyproj` stub (CRS -> None) `ComputeNetwork` uses Euclidean distances and leaves
    `lat_lon = None` (sbayes/preprocessing.py:136-155), which the cost_based geo prior
    dereferences (sbayes/model/prior.py:812).  Give the network an identity projection
    (`lat_lon = locations`) so that configs with `geo: cost_based` (test/test_files/config.yaml)
    run through `sbayes.cli.main` in this environment.  Harness-side only; idempotent."""
    r = ref()
    CN = r.preprocessing.ComputeNetwork
    if getattr(CN, "_sb_patched", False):
        return
    orig = CN.__init__

    def init(self, *a, **kw):
        orig(self, *a, **kw)
        if getattr(self, "lat_lon", None) is None:
            self.lat_lon = self.locations
    CN.__init__ = init
    CN._sb_patched = True


def seed_reference(seed: int):
    """Make the reference bit-reproducible: seed the legacy numpy
    generator, Python's `random`, and rebind the module-level `RNG` names."""
    r = ref()
    np.random.seed(seed)
    random.seed(seed)
    rng = np.random.default_rng(seed)
    for mod in (r.util, r.operators, r.mcmc_setup):
        mod.RNG = rng
    try:
        import sbayes.sampling.initializers as ini
        ini.RNG = rng
    except Exception:
        pass
    return rng


def build_reference_model(x: np.ndarray, applicable: np.ndarray, locations: np.ndarray,
                          confounders: "OrderedDict[str, np.ndarray]", n_clusters: int,
                          prior_cfg: dict, group_names: dict | None = None):
    """Build reference `Data` and `Model` objects from packed arrays.

    x            uint8 [N,F] state index (255 = NA)
    applicable   bool [F,S]
    confounders  name -> bool [G,N] group assignment
    prior_cfg    dict in the format of the `model.prior` section of an sBayes config
    """
    r = ref()
    N, F = x.shape
    S = applicable.shape[1]
    values = P.unpack_onehot(x, S)
    names = np.array([f"f{i}" for i in range(F)])
    state_names = [[f"s{j}" for j in range(int(applicable[i].sum()))] for i in range(F)]
    features = r.load_data.Features(values=values, names=names, states=applicable.copy(),
                                    state_names=state_names, na_number=int((x == P.NA).sum()))
    objects = r.load_data.Objects(id=[f"o{i}" for i in range(N)], locations=np.asarray(locations, dtype=float),
                                  names=[f"o{i}" for i in range(N)])
    confs = OrderedDict()
    for name, grp in confounders.items():
        gn = (group_names or {}).get(name) or ([f"g{i}" for i in range(grp.shape[0])] if grp.shape[0] > 1 or name != "universal" else ["<ALL>"])
        confs[name] = r.load_data.Confounder(name=name, group_assignment=grp.astype(bool), group_names=list(gn))
    data = r.load_data.Data(objects=objects, features=features, confounders=confs, projection=None)
    # with the CRS stub the network uses Euclidean distances and has no lat_lon; the
    # cost_based geo prior dereferences it (sbayes/model/prior.py:812)
    data.network.lat_lon = data.network.locations
    cfg = r.config.ModelConfig(clusters=n_clusters, confounders=list(confounders.keys()), prior=prior_cfg)
    model = r.model.Model(data, cfg)
    return data, model


def problem_from_reference(data, model) -> P.Problem:
    """Packed `Problem` equivalent to a reference (Data, Model) pair (the product's own
    converter, sbayes_b200/refbridge.py)."""
    from sbayes_b200 import refbridge
    return refbridge.problem_from_reference(data, model)


def state_from_sample(prob: P.Problem, sample) -> P.ChainState:
    """Reference `Sample` -> packed `ChainState` (counts stacked [G_tot,F,S] int32)."""
    from sbayes_b200 import refbridge
    return refbridge.state_from_sample(prob, sample)


def sample_from_state(prob: P.Problem, model, data, st: P.ChainState, i_step: int = 0):
    """Packed `ChainState` -> reference `Sample` (fresh caches)."""
    from sbayes_b200 import refbridge
    if st.counts is None:
        from .model import rebuild_counts
        st = P.ChainState(st.cluster_of, st.W, st.src, rebuild_counts(prob, st))
    return refbridge.sample_from_state(prob, model, data, st, ref().state.Sample, i_step=i_step)
