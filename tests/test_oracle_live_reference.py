"""The oracle against the LIVE reference (imported from /root/reference when it is there -- the build container; skipped on the
GPU box, where only the committed golden vectors of tests/golden/ travel).  A randomized differential run that widens the pin
beyond the golden cases: other shapes (n = 1 .. 400, d = 1 .. 300), every metric and dtype pair, scaled / shifted / constant /
duplicated rows, time decay, top_k beyond N.  Same comparison rules as tests/test_oracle_port.py: NumPy-loop metrics bit for bit,
BLAS-backed dot products within the stated tolerance, the reference's unspecified order among equal scores not compared."""
import contextlib
import importlib.util
import io
import os
import zlib

import numpy as np
import pytest

from oracle import canonical as K
from oracle import reference_port as P

REF = "/root/reference/hyperdb/ranking_algorithm.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="the reference sources are not on this machine")

METRICS = ["dot_product", "cosine_similarity", "euclidean_metric", "manhattan_distance", "hamming_distance", "jaccard_similarity",
           "pearson_correlation"]
DTS = (np.float16, np.float32, np.float64)
PORT = {"dot_product": P.dot_scores, "cosine_similarity": P.cosine_scores, "euclidean_metric": P.euclidean_scores,
        "manhattan_distance": P.manhattan_scores, "hamming_distance": P.hamming_scores, "jaccard_similarity": P.jaccard_scores,
        "pearson_correlation": P.pearson_scores}


@pytest.fixture(scope="module")
def ref():
    spec = importlib.util.spec_from_file_location("live_ref_ranking", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _inputs(rng, trial):
    n = int(rng.choice([1, 2, 3, 7, 33, 150, 400]))
    d = int(rng.choice([1, 2, 5, 16, 31, 96, 129, 300]))
    vdt, qdt = DTS[int(rng.integers(0, 3))], DTS[int(rng.integers(0, 3))]
    V = rng.standard_normal((n, d)) * float(rng.choice([0.02, 1.0, 30.0])) + float(rng.choice([0.0, 0.0, 2.5]))
    q = rng.standard_normal(d) * float(rng.choice([0.02, 1.0, 30.0])) + float(rng.choice([0.0, 0.0, 2.5]))
    if n > 3 and trial % 5 == 0:
        V[n // 2] = V[0]                                      # duplicates: equal scores
    if n > 2 and trial % 7 == 0:
        V[1] = 0.75                                           # a constant row (pearson: NaN, cosine: fine)
    if trial % 11 == 0:
        V[n - 1] = 0.0                                        # a zero row (norm 0 -> divided by 1)
    if trial % 13 == 0:
        q = V[0].astype(np.float64) * 1.0                     # an exact hit
    return np.ascontiguousarray(V.astype(vdt)), np.ascontiguousarray(q.astype(qdt))


@pytest.mark.parametrize("metric", METRICS)
def test_port_and_spec_equal_the_live_reference(ref, metric):
    rng = np.random.default_rng(zlib.crc32(("live" + metric).encode()))
    compared = 0
    for trial in range(120):
        V, q = _inputs(rng, trial)
        n = len(V)
        with np.errstate(all="ignore"), contextlib.redirect_stdout(io.StringIO()):
            want = np.asarray(getattr(ref, metric)(V.copy(), q.copy()))
            got = np.asarray(PORT[metric](V, q))
            spec = np.asarray(K.scores(V, q, metric))
        assert got.dtype == want.dtype and got.shape == want.shape, (trial, got.dtype, want.dtype)
        blas = metric in ("dot_product", "cosine_similarity") and want.dtype != np.float16
        if blas:
            tol = {"float32": 1e-5, "float64": 1e-12}[str(want.dtype)]
            cond = (np.linalg.norm(V.astype(float), axis=1) * np.linalg.norm(q.astype(float))) if metric == "dot_product" else 1.0
            for x in (got, spec):
                assert np.all(np.abs(x.astype(float) - want.astype(float)) <= tol * np.maximum(cond, 1e-300)), trial
        else:
            assert np.array_equal(got, want, equal_nan=True), (trial, V.dtype, q.dtype)
            assert np.array_equal(spec.astype(want.dtype), want, equal_nan=True), (trial, V.dtype, q.dtype)
        # the sort: time decay on every third trial, top_k around and beyond N
        ts = 1.7e9 + rng.uniform(0, 50, n) if trial % 3 == 0 else None
        bias = float(rng.choice([0.3, 1.0, -0.5])) if ts is not None else 0
        k = int(rng.choice([1, 3, 10, n, n + 5]))
        with np.errstate(all="ignore"), contextlib.redirect_stdout(io.StringIO()):
            r_idx, r_sc = ref.hyperDB_ranking_algorithm_sort(V.copy(), q.copy(), top_k=k, metric=metric,
                                                             timestamps=None if ts is None else ts.copy(), recency_bias=bias)
            p_idx, p_sc = P.rank(V, q, k, metric, ts, bias, canonical=False)
            c_idx, c_sc = P.rank(V, q, k, metric, ts, bias, canonical=True)
        r_sc = np.asarray(r_sc, float).reshape(-1)
        if blas:
            np.testing.assert_allclose(np.asarray(p_sc, float).reshape(-1), r_sc, rtol=1e-5 if want.dtype == np.float32 else 1e-12, atol=1e-30)
        else:
            assert np.array_equal(np.asarray(p_sc, float).reshape(-1), r_sc), trial
            assert np.array_equal(np.asarray(c_sc, float).reshape(-1), r_sc), trial
            with np.errstate(all="ignore"):
                full = P.similarities_f64(V, q, metric) + P.recency_term(ts, bias, n)
            if n > 1 and len(set(full.tolist())) == n:                  # no two rows tie: the reference's order is defined
                assert list(np.asarray(p_idx).reshape(-1)) == list(np.asarray(r_idx).reshape(-1)), trial
                assert list(np.asarray(c_idx).reshape(-1)) == list(np.asarray(r_idx).reshape(-1)), trial
        compared += 1
    assert compared == 120


def test_norm_vector_and_distance_form_equal_the_live_reference(ref):
    """get_norm_vector (hyperdb/ranking_algorithm.py:8-21) and euclidean_metric(get_similarity_score=False) (:44-52): NumPy loops
    only, so bit for bit."""
    rng = np.random.default_rng(2024)
    for trial in range(150):
        V, q = _inputs(rng, trial)
        with np.errstate(all="ignore"), contextlib.redirect_stdout(io.StringIO()):
            want = ref.get_norm_vector(V.copy())
            got = K.unit_rows(V)
            assert got.dtype == want.dtype and np.array_equal(got, want, equal_nan=True), trial
            wd = np.asarray(ref.euclidean_metric(V.copy(), q.copy(), get_similarity_score=False))
            gd = np.asarray(P.euclidean_scores(V, q, get_similarity_score=False))
            sd = K.euclidean_distance(V, q)
        assert gd.dtype == wd.dtype and gd.tobytes() == wd.tobytes(), trial
        assert sd.dtype == wd.dtype and sd.tobytes() == wd.tobytes(), trial
