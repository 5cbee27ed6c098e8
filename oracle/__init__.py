"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

`oracle/` is a CPU restatement of the brute-force ranking path of AdamCodd/local-hyperDB
(`hyperdb/ranking_algorithm.py` + the brute-force tail of `HyperDB._execute_query`,
`hyperdb/hyperdb.py:1554-1575`).  Two layers:

* `oracle.reference_port`  -- the same algorithm expressed with the same NumPy primitives the
  reference calls (np.dot / np.linalg.norm / np.sum / argpartition ...).  This is the checker for
  the parity tests and the `cpu_baseline` / `--impl reference` arm of `bench.py`.
* `oracle.canonical`       -- the arithmetic of those NumPy primitives spelled out operation by
  operation (sequential fp32 chain of HALF_dot, NumPy's pairwise summation, per-element fp16
  rounding ...).  It is the *specification* of the CUDA "certify" kernel and is itself checked
  against `reference_port` (bit-exact where NumPy's arithmetic is deterministic).

Parity pinning: both layers are checked (tests/test_oracle_*.py) against
  (1) the reference's own KATs (`tests/test_ranking_algorithm.py:6-123`, restated in
      tests/test_reference_kats.py),
  (2) golden vectors produced by importing the *real* reference module in the build container
      (`tests/golden/make_golden.py`, outputs committed under tests/golden/),
  (3) the demo fixture (`demo/pokemon_hyperdb.pickle`, BASELINE config C1).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.
"""
