"""CPU oracle for the syke-pic inference hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in numpy (byte / integer work) and plain torch-CPU fp32
functional calls (the CNN, which is floating point), what the reference
`sykepic prob` / `sykepic class` path computes.  Every function cites the
reference file:line it follows.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it, and only as the checker.
Nothing under `sykepic_b200/` imports it; the product path fails loudly if the
CUDA library is missing.

Parity pin: the restatement is checked (tests/test_oracle_*.py, `-m "not gpu"`)
against
  * the reference's own fixtures (tests/data/raw/valid, tests/data/prob,
    tests/model/thresholds-*.txt), committed as
    tests/golden/ref_fixture/*, and
  * golden vectors produced by running the reference itself
    (sykepic.compute.probability.main / prediction.prediction_dataframe,
    imported read-only from /root/reference in the build container) by
    tests/golden/make_golden.py; decoded bytes, preprocessed tensors,
    probabilities, CSV text and labels are committed under tests/golden/.
"""
