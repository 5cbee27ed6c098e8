"""The reference's own known-answer tests (tests/test_ranking_algorithm.py:6-123 of the reference,
plus SURVEY.md appendix B's probed outputs) restated against the oracle (CPU) and, under -m gpu,
against the CUDA drop-in module.  These are the reference's only pins at the function boundary."""
import numpy as np
import pytest

from oracle import reference_port as P
from oracle import canonical as K

V3 = np.array([[1, 0], [0, 1], [0.5, 0.5]])
Q3 = np.array([1, 0])
TS3 = [1627825200.0, 1627911600.0, 1627998000.0]

SORT_TABLE = [                                   # tests/test_ranking_algorithm.py:82-98
    ("cosine_similarity", 0, [0, 2, 1], [1.0, 0.70710678, 0.0]),
    ("cosine_similarity", 1, [2, 0, 1], [1.70710678, 1.0, 0.0]),
    ("euclidean_metric", 0, [0, 2, 1], [1.0, 0.58578644, 0.41421356]),
    ("manhattan_distance", 0, [0, 2, 1], [1.0, 0.5, 0.33333333]),
    ("hamming_distance", 0, [0, 2, 1], [2.0, 1.0, 0.0]),
    ("jaccard_similarity", 0, [0, 2, 1], [1.0, 0.5, 0.0]),
    ("pearson_correlation", 0, [0, 1, 2], [1.0, -1.0, -np.inf]),
]


def _impls():
    yield "port", lambda V, q, k, m, ts, b: P.rank(V, q, k, m, ts, b, canonical=True)
    yield "canonical", lambda V, q, k, m, ts, b: K.rank(V, q, k, m, ts, b)


@pytest.mark.parametrize("metric,bias,idx,sc", SORT_TABLE)
def test_sort_table(metric, bias, idx, sc):
    for name, fn in _impls():
        i, s = fn(V3.copy(), Q3.copy(), 5, metric, TS3, bias)
        assert list(i) == idx, name
        np.testing.assert_allclose(np.asarray(s, float).reshape(-1), sc, atol=1e-8, err_msg=name)


def test_metric_kats():
    # :24-29 cosine, :32-37 manhattan, :74-79 hamming, :7-14 euclid shape
    for mod_scores in (lambda V, q, m: np.asarray({"cosine_similarity": P.cosine_scores, "manhattan_distance": P.manhattan_scores,
                                                    "hamming_distance": P.hamming_scores, "euclidean_metric": P.euclidean_scores}[m](V, q)),
                       lambda V, q, m: K.scores(V, q, m)):
        assert np.array_equal(mod_scores(np.array([[1, 0], [0, 1]]), np.array([1, 0]), "cosine_similarity"), [1.0, 0.0])
        assert np.allclose(mod_scores(np.array([[1, 0], [0, 1]]), np.array([1, 0]), "manhattan_distance"), [1.0, 1 / 3])
        assert np.array_equal(mod_scores(np.array([[1, 1], [0, 1], [1, 0]]), np.array([1, 1]), "hamming_distance"), [2, 1, 1])
        r = mod_scores(np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]]), np.array([1, 1, 1]), "euclidean_metric")
        assert r.shape == (3,) and np.all(r > 0)


def test_error_kats():
    with pytest.raises(ValueError):                                   # :100-105 unknown metric
        P.rank(np.array([[1, 0], [0, 1]]), np.array([1, 0]), metric="unknown_metric")
    with pytest.raises(ValueError):                                   # :107-114 1-D vectors, euclid
        P.rank(np.array([1, 0]), np.array([1, 0]), metric="euclidean_metric")
    with pytest.raises(ValueError):                                   # :116-123 NaN
        P.rank(np.array([[1, 0], [0, 1], [np.nan, np.nan]]), np.array([1, 0]))
    with pytest.raises(ValueError):                                   # :16-21 empty arrays
        P.euclidean_scores(np.array([]), np.array([]))


def test_edge_quirks():
    # SURVEY.md section 3.4 item 8 / appendix B
    V = np.eye(5)
    assert len(P.rank(V, V[0], 50)[0]) == 5
    assert P.rank(V, V[0], 0) == ([], [])
    assert P.rank(V, V[0], -1) == ([], [])
    i, s = P.rank(V[:1], V[0], 3)
    assert list(i) == [0] and np.asarray(s).shape == (1, 1)
    dec = P.recency_term(1.7e9 + np.array([0, 1, 2, -5, 2.0]), 0.5, 5)
    np.testing.assert_allclose(dec, [0.0676676, 0.1839397, 0.5, 0.000455941, 0.5], rtol=1e-6)


def test_pokemon_c1_oracle():
    import golden_io as G
    z = G.load_pokemon()
    V32 = z["vectors"]
    V64 = V32.astype(np.float64)
    for row in (0, 25, 77, 150):
        for tag, V, q in (("f32", V32, V32[row]), ("f64", V64, V64[row]), ("mixed", V64, V32[row])):
            i, s = P.rank(V, q, 5, "cosine_similarity", canonical=True)
            assert list(i) == list(z[f"idx_{tag}_{row}"])
            np.testing.assert_allclose(s, z[f"sc_{tag}_{row}"], rtol=1e-6 if tag == "f32" else 1e-12)
            ci, cs = K.rank(V, q, 5, "cosine_similarity")
            assert list(ci) == list(z[f"idx_{tag}_{row}"])
            np.testing.assert_allclose(cs, z[f"sc_{tag}_{row}"], rtol=1e-6 if tag == "f32" else 1e-12)
    assert list(z["idx_f32_0"]) == [0, 119, 91, 104, 100]
