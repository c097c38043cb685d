"""CPU oracle for the sBayes MCMC hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `sbayes_b200/` may import this package; only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` use it, and there only as the checker or the timed CPU baseline.

Contents
--------
* `problem.py`   packed problem / chain-state containers shared by the oracle and the tests
* `philox.py`    Philox4x32-10 counter-based streams (same streams as the CUDA path)
* `draws.py`     draw sources: Philox (production streams) and Tape (replayed reference draws)
* `model.py`     fp64 numpy restatement of the target density (likelihood + priors + tables)
* `operators.py` fp64 numpy restatement of the ten scheduled MCMC operators
* `chain.py`     Metropolis-Hastings step / run loop / MC3 swap
* `refharness.py`, `refrecord.py`  import the live reference (only where /root/reference
  exists, i.e. in the build container) to validate the restatement and to generate the
  golden fixtures committed under `tests/golden/`.

Parity pinning: every function here was checked against the live reference
(`tests/golden/make_golden.py`, run in the build container); the committed fixtures under
`tests/golden/` are outputs of the reference itself.  The reference's own known-answer
values (doctests in sbayes/util.py) are asserted in
`tests/test_oracle_kat.py`.
"""
