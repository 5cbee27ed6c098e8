"""The NumPy port (oracle/reference_port.py) against outputs of the REAL reference
(tests/golden/*.npz, generated in the build container by tests/golden/make_golden.py).
Same NumPy primitives => bit-identical similarity vectors and identical (indices, scores)."""
import numpy as np
import pytest

import golden_io as G
from oracle import reference_port as P

GOLDEN = G.load_sort_golden()
FN = {"dot_product": P.dot_scores, "cosine_similarity": P.cosine_scores, "euclidean_metric": P.euclidean_scores,
      "manhattan_distance": P.manhattan_scores, "hamming_distance": P.hamming_scores, "jaccard_similarity": P.jaccard_scores,
      "pearson_correlation": P.pearson_scores}


@pytest.mark.parametrize("entry", GOLDEN, ids=[G.case_id(e[0]) for e in GOLDEN])
def test_port_matches_reference(entry):
    case, ref_sims, ref_idx, ref_sc = entry
    V, q, ts = G.inputs(case)
    with np.errstate(all="ignore"):
        sims = np.asarray(FN[case["metric"]](V, q))
    assert sims.dtype == ref_sims.dtype
    blas = case["metric"] in ("dot_product", "cosine_similarity") and str(sims.dtype) != "float16"
    if blas:
        # OpenBLAS sgemv/dgemv: the summation order depends on the CPU model and buffer alignment,
        # so only the deterministic NumPy loops are compared bit for bit
        tol = {"float32": 1e-5, "float64": 1e-12}[str(sims.dtype)]
        cond = np.linalg.norm(V.astype(float), axis=1) * np.linalg.norm(q.astype(float)) if case["metric"] == "dot_product" else 1.0
        assert np.all(np.abs(sims.astype(float) - ref_sims.astype(float)) <= tol * np.maximum(cond, 1e-300))
        return
    assert np.array_equal(sims, ref_sims, equal_nan=True) if case["metric"] == "pearson_correlation" else sims.tobytes() == ref_sims.tobytes()
    with np.errstate(all="ignore"):
        idx, sc = P.rank(V, q, case["k"], case["metric"], ts, case["bias"] if ts is not None else 0, canonical=False)
    assert np.array_equal(np.asarray(sc, float).reshape(-1), ref_sc)
    if len(set(ref_sc.tolist())) == len(ref_sc):      # the reference's order among equal scores is unspecified
        assert list(idx) == list(ref_idx)
    # canonical mode: same score multiset, ties resolved to the lower index
    with np.errstate(all="ignore"):
        cidx, csc = P.rank(V, q, case["k"], case["metric"], ts, case["bias"] if ts is not None else 0, canonical=True)
    assert np.array_equal(np.asarray(csc, float).reshape(-1), ref_sc)
    if len(ref_idx):
        full = P.similarities_f64(V, q, case["metric"]) + P.recency_term(ts, case["bias"] if ts is not None else 0, len(V))
        assert list(cidx) == list(G.canon_order(full, case["k"]))


def test_inputs_not_mutated():
    V = np.array([[0.5, -2.0], [3.0, 0.0]])
    q = np.array([0.25, 4.0])
    V0, q0 = V.copy(), q.copy()
    P.hamming_scores(V, q)
    assert np.array_equal(V, V0) and np.array_equal(q, q0)


def test_chunked_equals_whole():
    rng = np.random.default_rng(3)
    V = rng.standard_normal((1000, 24)).astype(np.float32)
    q = rng.standard_normal(24).astype(np.float32)
    ts = 1.7e9 + rng.uniform(0, 9, 1000)
    keep = rng.random(1000) < 0.5
    chunks = [(o, V[o:o + 128]) for o in range(0, 1000, 128)]
    ids, sc = P.chunked_rank(chunks, q, 15, "euclidean_metric", ts, 0.3, keep)
    sub = np.flatnonzero(keep)
    idx, sc2 = P.rank(V[sub], q, 15, "euclidean_metric", ts[sub], 0.3, canonical=True)
    assert list(ids) == list(sub[idx]) and np.array_equal(sc, sc2)


def test_tail_matches_real_hyperdb():
    z, specs = G.load_tail()
    V, ts, queries, groups = z["V"], z["ts"], z["queries"], z["groups"]
    n = len(V)
    for si, spec in enumerate(specs):
        keep = np.ones(n, bool)
        for name, par in spec["filters"] or []:
            if name == "skip_doc":
                keep &= (np.arange(n) >= par) if par > 0 else (np.arange(n) < n + par)
            elif name == "metadata":
                keep &= groups == par["group"]
        for qi in range(len(queries)):
            ids, sc = P.hyperdb_bruteforce_tail(V, queries[qi], spec["top_k"], spec["metric"], ts, spec["recency_bias"], keep)
            assert list(ids) == list(z[f"ids_{si}_{qi}"]), (si, qi)
            blas = spec["metric"] in ("dot_product", "cosine_similarity")
            np.testing.assert_allclose(sc, z[f"sc_{si}_{qi}"], rtol=1e-5 if blas else 0, atol=0)


DIST = G.load_dist_golden()


@pytest.mark.parametrize("entry", DIST, ids=[G.case_id(e[0]) for e in DIST])
def test_euclidean_distance_form(entry):
    """euclidean_metric(..., get_similarity_score=False), hyperdb/ranking_algorithm.py:49-52: the port and the
    explicit-arithmetic spec against the REAL reference's distances, bit for bit (NumPy loops only, no BLAS)."""
    from oracle import canonical as K
    case, ref_dist = entry
    V, q, _ts = G.inputs(case)
    with np.errstate(all="ignore"):
        d_port = np.asarray(P.euclidean_scores(V, q, get_similarity_score=False))
        d_spec = K.euclidean_distance(V, q)
    assert d_port.dtype == ref_dist.dtype and d_port.tobytes() == ref_dist.tobytes()
    assert d_spec.dtype == ref_dist.dtype and d_spec.tobytes() == ref_dist.tobytes()
