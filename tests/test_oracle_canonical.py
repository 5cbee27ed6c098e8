"""The explicit-arithmetic spec (oracle/canonical.py) against the REAL reference's outputs.

Bit-exact wherever NumPy's arithmetic is deterministic: every float16 metric, and euclidean /
manhattan / hamming in every dtype.  float32/float64 dot and cosine go through OpenBLAS in the
reference (unknowable summation order): scores within north_star's tolerances (rel 1e-5 / 1e-12)
and index lists identical unless the two rows involved are within that tolerance of each other."""
import numpy as np
import pytest

import golden_io as G
from oracle import canonical as K

GOLDEN = G.load_sort_golden()
TOL = {"float16": 1e-3, "float32": 1e-5, "float64": 1e-12, "uint64": 0, "int64": 0}


def exact_expected(case, dtype):
    return dtype == "float16" or case["metric"] in ("euclidean_metric", "manhattan_distance", "hamming_distance", "jaccard_similarity",
                                                 "pearson_correlation")


@pytest.mark.parametrize("entry", GOLDEN, ids=[G.case_id(e[0]) for e in GOLDEN])
def test_canonical_matches_reference(entry):
    case, ref_sims, ref_idx, ref_sc = entry
    V, q, ts = G.inputs(case)
    sims = K.scores(V, q, case["metric"])
    dt = str(ref_sims.dtype)
    assert str(sims.dtype) == dt
    if case["metric"] == "pearson_correlation":
        assert np.array_equal(sims, ref_sims.reshape(-1), equal_nan=True)
    elif exact_expected(case, dt):
        assert sims.tobytes() == ref_sims.reshape(-1).tobytes()
    else:
        a, b = sims.astype(np.float64), ref_sims.reshape(-1).astype(np.float64)
        # a dot product's rounding error scales with |v||q| (cancellation), not with the result
        cond = np.ones(len(b)) if case["metric"] == "cosine_similarity" else \
            np.linalg.norm(V.astype(np.float64), axis=1) * np.linalg.norm(q.astype(np.float64))
        assert np.all(np.abs(a - b) <= TOL[dt] * np.maximum(np.abs(b), cond))
    bias = case["bias"] if ts is not None else 0
    idx, sc = K.rank(V, q, case["k"], case["metric"], ts, bias)
    if case["n"] == 1 or len(ref_idx) == 0:
        assert len(idx) == len(ref_idx)
        return
    assert len(idx) == len(ref_idx)
    if exact_expected(case, dt):
        # same score multiset; canonical order == reference scores canonicalised
        full = K.total_scores(V, q, case["metric"], ts, bias)
        assert np.array_equal(sc, ref_sc)
        assert list(idx) == list(G.canon_order(full, case["k"]))
    else:
        np.testing.assert_allclose(sc, ref_sc, rtol=TOL[dt] * 10, atol=TOL[dt] * 1e-2)
        # identical index lists, except swaps certified as tolerance-ties
        ref_full = ref_sims.reshape(-1).astype(np.float64)
        if ts is not None:
            ref_full = ref_full + bias * np.exp(-np.max(ts) + ts)
        ref_canon = G.canon_order(ref_full, case["k"])
        for mine, theirs in zip(idx, ref_canon):
            if mine != theirs:
                assert abs(ref_full[mine] - ref_full[theirs]) <= TOL[dt] * max(1.0, abs(ref_full[theirs]))


def test_pairwise_sum_matches_numpy_reduce():
    rng = np.random.default_rng(0)
    for d in (1, 7, 8, 9, 127, 128, 129, 255, 256, 257, 1000, 4097):
        for dt in (np.float16, np.float32, np.float64):
            x = (rng.standard_normal((50, d)) * 3).astype(dt)
            assert K.row_sum(x).tobytes() == np.add.reduce(x, axis=1).tobytes()
            assert K.row_norm(x).tobytes() == np.linalg.norm(x, axis=1).tobytes()


def test_mean_std_match_numpy():
    """np.mean / np.std spelled out (numpy/_core/_methods.py::_mean, _var, _std) -- the per-row statistics of
    pearson_correlation (hyperdb/ranking_algorithm.py:90-94), array form (axis=1) and scalar form (1-D query)."""
    rng = np.random.default_rng(2)
    for d in (1, 2, 7, 8, 9, 100, 128, 129, 384, 1000):
        for dt in (np.float16, np.float32, np.float64):
            for scale, shift in ((1.0, 0.0), (0.03, 0.5), (4.0, -20.0)):
                x = (rng.standard_normal((40, d)) * scale + shift).astype(dt)
                with np.errstate(all="ignore"):
                    assert K.row_mean(x).tobytes() == np.mean(x, axis=1).tobytes(), (d, dt, scale)
                    assert K.row_std(x).tobytes() == np.std(x, axis=1).tobytes(), (d, dt, scale)
                    for r in (0, 17):
                        assert K.row_mean(x[r][None, :], scalar=True)[0].tobytes() == np.mean(x[r]).tobytes()
                        assert K.row_std(x[r][None, :])[0].tobytes() == np.std(x[r]).tobytes()


def test_pearson_edge_cases_match_port():
    from oracle import reference_port as P
    rng = np.random.default_rng(3)
    for vdt, qdt in ((np.float16, np.float16), (np.float16, np.float64), (np.float32, np.float32), (np.float32, np.float64), (np.float64, np.float32)):
        V = (rng.standard_normal((50, 24)) * 2 + 1).astype(vdt)
        V[3] = 0.75                                   # constant row -> NaN
        V[4] = 0
        V[5] = V[5] * 1e-4                            # tiny std: float16 denominators can underflow to zero -> 0.0, not NaN
        for q in ((rng.standard_normal(24) + 0.5).astype(qdt), np.full(24, 0.5, qdt)):
            with np.errstate(all="ignore"):
                ref = P.pearson_scores(V, q)
            assert np.array_equal(K.pearson(V, q), ref, equal_nan=True), (vdt, qdt)


def test_spec_equals_numpy_on_random_shapes():
    """Beyond the golden cases: the operation-by-operation spec against the NumPy calls of the reference (the port) on
    seeded random shapes / scales / dtype pairs, bit for bit wherever NumPy is deterministic."""
    from oracle import reference_port as P
    rng = np.random.default_rng(20241018)
    fns = {"euclidean_metric": P.euclidean_scores, "manhattan_distance": P.manhattan_scores, "hamming_distance": P.hamming_scores,
           "jaccard_similarity": P.jaccard_scores, "pearson_correlation": P.pearson_scores, "dot_product": P.dot_scores,
           "cosine_similarity": P.cosine_scores}
    dts = (np.float16, np.float32, np.float64)
    checked = 0
    for trial in range(120):
        n, d = int(rng.integers(1, 40)), int(rng.integers(1, 300))
        vdt, qdt = dts[int(rng.integers(0, 3))], dts[int(rng.integers(0, 3))]
        scale, shift = float(rng.choice([0.01, 1.0, 30.0])), float(rng.choice([0.0, 0.0, 2.5]))
        V = (rng.standard_normal((n, d)) * scale + shift).astype(vdt)
        q = (rng.standard_normal(d) * scale + shift).astype(qdt)
        if trial % 7 == 0:
            V[0] = 0
        for metric, fn in fns.items():
            rdt = np.promote_types(vdt, qdt)
            if metric in ("dot_product", "cosine_similarity") and rdt != np.float16:
                continue                                   # OpenBLAS order: covered by tolerance tests
            with np.errstate(all="ignore"):
                want = np.asarray(fn(V.copy(), q.copy())).reshape(-1)
                got = K.scores(V, q, metric)
            assert got.dtype == want.dtype, (metric, vdt, qdt)
            assert np.array_equal(got, want, equal_nan=True), (trial, metric, vdt.__name__, qdt.__name__, n, d, scale, shift)
            checked += 1
    assert checked > 500
