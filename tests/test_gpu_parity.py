"""Parity of the CUDA path (through the C ABI, via the drop-in module) against the oracle and the
committed outputs of the real reference.  Bar (north_star): top-k indices identical with ties to the
lower index; scores bit-equal where NumPy's arithmetic is deterministic (all float16 metrics,
euclidean/manhattan/hamming in every dtype), else within rel 1e-5 (fp32) / 1e-12 (fp64)."""
import os

import numpy as np
import pytest

import golden_io as G
from oracle import canonical as K
from oracle import reference_port as P

pytestmark = pytest.mark.gpu

GOLDEN = G.load_sort_golden()
TOL = {"float16": 1e-3, "float32": 1e-5, "float64": 1e-12, "uint64": 0}
NATIVE_DTYPE = {"hamming_distance": "uint64", "jaccard_similarity": "float64", "pearson_correlation": "float64"}


@pytest.fixture(scope="module")
def hb():
    import hyperdb_b200
    return hyperdb_b200


def exact_expected(metric, dtype):
    return dtype == "float16" or metric in ("euclidean_metric", "manhattan_distance", "hamming_distance", "jaccard_similarity",
                                            "pearson_correlation")


def check_against_oracle(V, q, ts, bias, k, metric, idx, sc, keep=None):
    """idx/sc from the CUDA path vs the canonical oracle (same arithmetic spec)."""
    oi, os_ = K.rank(V, q, k, metric, ts, bias, keep)
    rdt = str(np.promote_types(np.asarray(V).dtype if np.asarray(V).dtype.kind == "f" else np.float64,
                               np.asarray(q).dtype if np.asarray(q).dtype.kind == "f" else np.float64))
    assert len(idx) == len(oi)
    if exact_expected(metric, rdt):
        assert list(idx) == list(oi)
        if ts is None:
            assert np.array_equal(sc, os_)
        else:
            np.testing.assert_allclose(sc, os_, rtol=1e-14)       # exp() of CUDA vs NumPy: <= 1 ulp apart
    else:
        np.testing.assert_allclose(sc, os_, rtol=TOL[rdt], atol=TOL[rdt] * 1e-3)
        full = K.total_scores(V, q, metric, ts, bias, keep)
        for mine, theirs in zip(idx, oi):
            if mine != theirs:      # only acceptable as a certified tolerance-tie
                assert abs(full[mine] - full[theirs]) <= TOL[rdt] * max(1.0, abs(full[theirs]))


@pytest.mark.parametrize("mode", [0, 1], ids=["auto", "exact-path"])
@pytest.mark.parametrize("entry", GOLDEN, ids=[G.case_id(e[0]) for e in GOLDEN])
def test_sort_golden(hb, entry, mode):
    case, ref_sims, ref_idx, ref_sc = entry
    V, q, ts = G.inputs(case)
    bias = case["bias"] if ts is not None else 0
    hb.ranking_algorithm.set_path_mode(mode)
    try:
        idx, sc = hb.hyperDB_ranking_algorithm_sort(V, q, top_k=case["k"], metric=case["metric"], timestamps=ts, recency_bias=bias)
    finally:
        hb.ranking_algorithm.set_path_mode(0)
    if case["k"] <= 0:
        assert idx == [] and sc == []
        return
    idx, sc = np.asarray(idx), np.asarray(sc, float).reshape(-1)
    check_against_oracle(V, q, ts, bias, case["k"], case["metric"], idx, sc)
    # and against what the REAL reference returned (tests/golden/sort_golden.npz)
    dt = str(ref_sims.dtype)
    assert len(idx) == len(ref_idx)
    if exact_expected(case["metric"], dt) and ts is None:
        assert np.array_equal(sc, ref_sc)
    else:
        np.testing.assert_allclose(sc, ref_sc, rtol=max(TOL[dt], 1e-13) * 10, atol=TOL[dt] * 1e-2)


METRIC_FN = {"dot_product": "dot_product", "cosine_similarity": "cosine_similarity", "euclidean_metric": "euclidean_metric",
             "manhattan_distance": "manhattan_distance", "hamming_distance": "hamming_distance",
             "jaccard_similarity": "jaccard_similarity", "pearson_correlation": "pearson_correlation"}


@pytest.mark.parametrize("entry", GOLDEN[::3], ids=[G.case_id(e[0]) for e in GOLDEN[::3]])
def test_metric_functions_golden(hb, entry):
    case, ref_sims, _, _ = entry
    V, q, _ts = G.inputs(case)
    V0, q0 = V.copy(), q.copy()
    out = getattr(hb, METRIC_FN[case["metric"]])(V, q)
    assert np.array_equal(V, V0) and np.array_equal(q, q0)        # no in-place binarisation (quirk 9)
    dt = str(ref_sims.dtype)
    assert str(out.dtype) == dt and out.shape == ref_sims.reshape(-1).shape
    if case["metric"] in ("jaccard_similarity", "pearson_correlation"):
        assert np.array_equal(out, ref_sims.reshape(-1), equal_nan=True)      # 0/0: NaN payloads may differ between x86 and CUDA
    elif exact_expected(case["metric"], dt):
        assert out.tobytes() == ref_sims.reshape(-1).tobytes()
    else:
        a, b = out.astype(float), ref_sims.reshape(-1).astype(float)
        cond = np.ones(len(b)) if case["metric"] == "cosine_similarity" else \
            np.linalg.norm(V.astype(float), axis=1) * np.linalg.norm(q.astype(float))
        assert np.all(np.abs(a - b) <= TOL[dt] * np.maximum(np.abs(b), cond))


DIST = G.load_dist_golden()


@pytest.mark.parametrize("entry", DIST, ids=[G.case_id(e[0]) for e in DIST])
def test_euclidean_distance_form_golden(hb, entry):
    """euclidean_metric(V, q, get_similarity_score=False) (hyperdb/ranking_algorithm.py:49-52): the kernel's own distance,
    bit-equal to the REAL reference's output in every dtype -- not reconstructed from 1/(1+d)."""
    case, ref_dist = entry
    V, q, _ts = G.inputs(case)
    out = hb.euclidean_metric(V, q, get_similarity_score=False)
    assert out.dtype == ref_dist.dtype and out.shape == ref_dist.shape
    assert out.tobytes() == ref_dist.tobytes()
    sims = hb.euclidean_metric(V, q)                         # default form unchanged
    one = out.dtype.type(1)
    with np.errstate(all="ignore"):
        assert sims.tobytes() == (one / (one + ref_dist)).tobytes()


def test_euclidean_distance_small_and_zero(hb):
    """exact matches give exactly 0 and tiny distances keep their relative accuracy (the 1/s - 1 round trip lost both)"""
    rng = np.random.default_rng(5)
    for dt in (np.float16, np.float32, np.float64):
        V = rng.standard_normal((64, 40)).astype(dt)
        q = V[7].copy()
        V[9] = q + np.asarray(1e-3, dt)
        d = hb.euclidean_metric(V, q, get_similarity_score=False)
        with np.errstate(all="ignore"):
            want = K.euclidean_distance(V, q)
        assert d.dtype == np.dtype(dt) and d.tobytes() == want.tobytes() and d[7] == 0


def test_get_norm_vector(hb):
    rng = np.random.default_rng(11)
    for dt in (np.float16, np.float32, np.float64):
        x = (rng.standard_normal((40, 77)) * 3).astype(dt)
        x[5] = 0
        assert hb.get_norm_vector(x).tobytes() == P.unit_rows(x).tobytes()
        assert hb.get_norm_vector(x[3]).tobytes() == P.unit_rows(x[3]).tobytes()


# ---- the reference's own KATs against the drop-in (tests/test_ranking_algorithm.py of the reference) ------
def test_reference_kats_on_gpu(hb):
    r = hb.euclidean_metric(np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]]), np.array([1, 1, 1]))
    assert r.shape == (3,) and np.all(r > 0)
    with pytest.raises(ValueError):
        hb.euclidean_metric(np.array([]), np.array([]))
    assert np.array_equal(hb.cosine_similarity(np.array([[1, 0], [0, 1]]), np.array([1, 0])), [1.0, 0.0])
    assert np.allclose(hb.manhattan_distance(np.array([[1, 0], [0, 1]]), np.array([1, 0])), [1.0, 1 / 3])
    assert np.array_equal(hb.hamming_distance(np.array([[1, 1], [0, 1], [1, 0]]), np.array([1, 1])), [2, 1, 1])
    assert np.array_equal(hb.jaccard_similarity(np.array([[1, 1], [1, 0], [0, 0]]), np.array([1, 1])), [1.0, 0.5, 0.0])   # :40-46
    assert np.array_equal(hb.jaccard_similarity(np.array([[2, 2], [2, 0], [0, 0]]), np.array([1, 1])), [1.0, 0.5, 0.0])   # :48-52
    r = hb.pearson_correlation(np.array([[1, 1], [0, 1], [1, 0]]), np.array([1, 1]))                        # :55-62
    assert np.isnan(r[0]) and r[1] != 0.0 and r[2] != 0.0
    assert np.all(np.isnan(hb.pearson_correlation(np.array([[1, 1], [0, 0], [1, 1]]), np.array([1, 1]))))       # :64-71
    V = np.array([[1, 0], [0, 1], [0.5, 0.5]])
    q = np.array([1, 0])
    ts = [1627825200.0, 1627911600.0, 1627998000.0]
    table = [("pearson_correlation", 0, [0, 1, 2]),("cosine_similarity", 0, [0, 2, 1]), ("cosine_similarity", 1, [2, 0, 1]), ("euclidean_metric", 0, [0, 2, 1]),
             ("manhattan_distance", 0, [0, 2, 1]), ("hamming_distance", 0, [0, 2, 1]), ("jaccard_similarity", 0, [0, 2, 1])]
    for metric, bias, want in table:
        idx, _ = hb.hyperDB_ranking_algorithm_sort(V, q, metric=metric, timestamps=ts, recency_bias=bias)
        assert list(idx) == want, metric
        idx, _ = hb.custom_ranking_algorithm_sort(V, q, metric=metric, timestamps=ts, recency_bias=bias)
        assert list(idx) == want, metric
    with pytest.raises(ValueError):
        hb.hyperDB_ranking_algorithm_sort(np.array([[1, 0], [0, 1]]), np.array([1, 0]), metric="unknown_metric")
    with pytest.raises(ValueError):
        hb.hyperDB_ranking_algorithm_sort(np.array([1, 0]), np.array([1, 0]), metric="euclidean_metric")
    with pytest.raises(ValueError):
        hb.hyperDB_ranking_algorithm_sort(np.array([[1, 0], [0, 1], [np.nan, np.nan]]), np.array([1, 0]))
    with pytest.raises(ValueError):
        hb.hyperDB_ranking_algorithm_sort(np.array([[1, 0], [0, 1]]), np.array([np.nan, 0]))


def test_edge_quirks_on_gpu(hb, capsys):
    V = np.eye(5)
    assert len(hb.hyperDB_ranking_algorithm_sort(V, V[0], 50)[0]) == 5
    assert hb.hyperDB_ranking_algorithm_sort(V, V[0], 0) == ([], [])
    assert hb.hyperDB_ranking_algorithm_sort(V, V[0], -1) == ([], [])
    i, s = hb.hyperDB_ranking_algorithm_sort(V[:1], V[0], 3)
    assert list(i) == [0] and np.asarray(s).shape == (1, 1) and s[0, 0] == 1.0
    assert "Only one document left" in capsys.readouterr().out
    # list of row arrays (what the filters hand over, hyperdb/hyperdb.py:1305)
    rows = [np.array([1.0, 0.0]), np.array([0.0, 1.0]), np.array([0.5, 0.5])]
    i, _ = hb.hyperDB_ranking_algorithm_sort(rows, np.array([1.0, 0.0]), timestamps=[1.0, 2.0, 3.0], recency_bias=0)
    assert list(i) == [0, 2, 1]


def test_pokemon_c1(hb):
    z = G.load_pokemon()
    V32 = z["vectors"]
    V64 = V32.astype(np.float64)
    for row in (0, 25, 77, 150):
        for tag, V, q in (("f32", V32, V32[row]), ("f64", V64, V64[row]), ("mixed", V64, V32[row])):
            i, s = hb.hyperDB_ranking_algorithm_sort(V, q, top_k=5, metric="cosine_similarity")
            assert list(i) == list(z[f"idx_{tag}_{row}"])
            np.testing.assert_allclose(s, z[f"sc_{tag}_{row}"], rtol=1e-5 if tag == "f32" else 1e-12)


# ---- fused sweep == exact path on larger inputs; certification statistics ----------------------------------
SWEEP_CASES = [(dt, m, n, d, kind) for dt in ("f16", "f32", "f64")
               for m in ("dot_product", "cosine_similarity", "euclidean_metric", "manhattan_distance", "hamming_distance",
                         "jaccard_similarity", "pearson_correlation")
               for (n, d, kind) in ((20000, 96, "unit"), (50021, 200, "scaled"), (30000, 33, "coarse"))]


@pytest.mark.parametrize("vdt,metric,n,d,kind", SWEEP_CASES)
def test_fused_equals_exact(hb, vdt, metric, n, d, kind):
    import cases as C
    import zlib
    case = dict(seed=zlib.crc32(f"{vdt}{metric}{n}{d}".encode()) % 100000, n=n, d=d, vdt=vdt, qdt=vdt, kind=kind, ts=True, ts_span=10.0)
    V, q, ts = C.make_inputs(case)
    m = hb.DeviceMatrix(V)
    rng = np.random.default_rng(1)
    keep = rng.random(n) < 0.6
    fallbacks = 0
    try:
        for (k, use_ts, use_mask) in ((10, False, False), (100, False, False), (16, True, False), (10, False, True), (7, True, True)):
            m.set_mask(keep if use_mask else None)
            m.set_timestamps(ts if use_ts else None)
            if use_ts:
                m.refresh_decay()
            bias = 0.4 if use_ts else 0.0
            m.set_path(1)
            ei, es, ec, _ = m.query(q, k, metric, bias)
            m.set_path(0)
            fi, fs, fc, ff = m.query(q, k, metric, bias)
            fallbacks += int(ff[0] & 1)
            assert ec[0] == fc[0] == min(k, int(keep.sum()) if use_mask else n)
            assert list(fi[0]) == list(ei[0]), (k, use_ts, use_mask)
            assert np.array_equal(fs[0], es[0])
            if k == 10:      # and the exact path itself against the oracle
                check_against_oracle(V, q, ts if use_ts else None, bias, k, metric, ei[0], es[0], keep if use_mask else None)
    finally:
        m.close()
    if kind == "unit":
        assert fallbacks == 0, "fused pass should certify on well-separated data"


MQ_CASES = [(vdt, metric, n, d) for vdt in ("f16", "f32", "f64")
            for metric in ("dot_product", "cosine_similarity", "euclidean_metric", "manhattan_distance", "hamming_distance",
                           "jaccard_similarity", "pearson_correlation")
            for (n, d) in ((30011, 136), (9001, 33))]


@pytest.mark.parametrize("vdt,metric,n,d", MQ_CASES)
def test_multi_query_sweep(hb, vdt, metric, n, d):
    """Small batches share ONE read of the matrix (1 / 2 / 4 / 8 queries per sweep pass, SURVEY.md section 7.2 "B <= ~8"): the
    batched answers must equal, bit for bit, one-query-per-pass answers (hdb_matrix_set_max_group(1)) for every batch size
    1..11, with and without row mask / time decay, for k = 10 and the wide class k = 100, on aligned (d = 136) and unaligned
    (d = 33) rows; the B = 5 batch is also checked against the oracle."""
    import cases as C
    import zlib
    case = dict(seed=zlib.crc32(f"mq{vdt}{metric}{n}{d}".encode()) % 100000, n=n, d=d, vdt=vdt, qdt=vdt, kind="gauss", ts=True, ts_span=10.0)
    V, _q, ts = C.make_inputs(case)
    rng = np.random.default_rng(case["seed"])
    Q = rng.standard_normal((11, d)).astype(V.dtype)
    Q[3] = V[17]                                            # an exact hit
    Q[6] = Q[2]                                             # two identical queries in one pass
    keep = rng.random(n) < 0.55
    m = hb.DeviceMatrix(V)
    try:
        for (k, use_ts, use_mask) in ((10, False, False), (100, False, False), (10, True, True)):
            m.set_mask(keep if use_mask else None)
            m.set_timestamps(ts if use_ts else None)
            if use_ts:
                m.refresh_decay()
            bias = 0.4 if use_ts else 0.0
            m.set_max_group(1)
            ref = m.query(Q, k, metric, bias)
            m.set_max_group(0)
            for b in (2, 3, 4, 5, 8, 9, 11):
                got = m.query(Q[:b], k, metric, bias)
                assert np.array_equal(got[0], ref[0][:b]), (k, use_ts, use_mask, b)
                assert np.array_equal(got[1], ref[1][:b]) and np.array_equal(got[2], ref[2][:b])
            if k == 10:
                for b in range(5):
                    check_against_oracle(V, Q[b], ts if use_ts else None, bias, k, metric, ref[0][b], ref[1][b], keep if use_mask else None)
    finally:
        m.close()


@pytest.mark.parametrize("d", [130, 600, 1100, 2500, 4200])
def test_hamming_wide_rows(hb, d):
    """every lanes-per-row class of the bit-packed sweep (d <= 4096) and the generic form above it"""
    rng = np.random.default_rng(d)
    n = 40000
    V = rng.standard_normal((n, d)).astype(np.float16)
    q = rng.standard_normal(d).astype(np.float16)
    keep = rng.random(n) < 0.8
    m = hb.DeviceMatrix(V)
    try:
        for use_mask in (False, True):
            m.set_mask(keep if use_mask else None)
            idx, sc, cnt, flags = m.query(q, 16, "hamming_distance")
            assert not (flags[0] & 1)
            oi, os_ = K.rank(V, q, 16, "hamming_distance", keep=keep if use_mask else None)
            assert list(idx[0]) == list(oi) and np.array_equal(sc[0], os_)
    finally:
        m.close()


def test_range_and_batch(hb):
    rng = np.random.default_rng(5)
    V = rng.standard_normal((5000, 64)).astype(np.float32)
    Q = rng.standard_normal((9, 64)).astype(np.float32)
    m = hb.DeviceMatrix(V)
    try:
        m.set_range(100, 4321)
        keep = np.zeros(5000, bool)
        keep[100:4321] = True
        idx, sc, cnt, _ = m.query(Q, 10, "cosine_similarity")
        for b in range(len(Q)):
            check_against_oracle(V, Q[b], None, 0.0, 10, "cosine_similarity", idx[b], sc[b], keep)
        m.set_range(0, 5000)
        idx, sc, cnt, _ = m.query(Q, 5000, "euclidean_metric")        # k = N: exact path
        assert cnt.tolist() == [5000] * 9
        oi, os_ = K.rank(V, Q[3], 5000, "euclidean_metric")
        assert list(idx[3]) == list(oi) and np.array_equal(sc[3], os_)
    finally:
        m.close()


# ---- tensor-core batched path (tcgen05) == streaming sweep == oracle ------------------------------------------------
@pytest.mark.parametrize("sdt", ["float16", "float32"])
@pytest.mark.parametrize("metric", ["dot_product", "cosine_similarity", "euclidean_metric"])
@pytest.mark.parametrize("nq", [17, 64, 130, 300])
def test_batched_tensor_path(hb, metric, nq, sdt):
    import torch
    n, d = 530_000, 136                      # row bytes are a multiple of 16 but not of 128: exercises the K tail
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(nq)
    V = torch.randn(n, d, generator=g, device=dev) * (0.5 + torch.rand(n, 1, generator=g, device=dev))
    if metric == "euclidean_metric":
        # the norm expansion |v|^2 + |q|^2 - 2 v.q cancels: its certificate only holds on embedding-like (unit-norm) data;
        # anything else is still answered exactly, by the sweep or the exact path (checked below through the equality)
        V = V / V.norm(dim=1, keepdim=True)
    V = V.half() if sdt == "float16" else V.float()
    Q = torch.randn(nq, d, generator=g, device=dev)
    if metric == "euclidean_metric":
        Q = Q / Q.norm(dim=1, keepdim=True)
    Q = Q.half() if sdt == "float16" else Q.float()
    ts = 1.7e9 + 10 * torch.rand(n, generator=g, device=dev, dtype=torch.float64)
    keep_bits = torch.randint(-2**31, 2**31 - 1, ((n + 31) // 32,), generator=g, device=dev, dtype=torch.int32)
    q_np = Q.cpu().numpy()
    m = hb.DeviceMatrix(V)
    try:
        for use_ts, use_mask, k in ((False, False, 10), (True, True, 10), (False, True, 100)):
            m.set_mask(keep_bits if use_mask else None)
            m.set_timestamps(ts if use_ts else None)
            if use_ts:
                m.refresh_decay()
            bias = 0.2 if use_ts else 0.0
            # reference: the streaming sweep (forced), or for euclidean -- whose float16 scores tie so densely on
            # this data that the sweep's own certificate may refuse -- the exact path on the first queries
            nref = nq if metric != "euclidean_metric" else min(nq, 24)
            m.set_path(2 if metric != "euclidean_metric" else 1)
            i0, s0, c0, f0 = m.query(q_np[:nref], k, metric, bias)
            m.set_path(0)                                       # automatic: tensor cores for >= 2 queries on shards of >= 65 536 rows
            i1, s1, c1, f1 = m.query(q_np, k, metric, bias)
            i1c, s1c, c1c = i1[:nref], s1[:nref], c1[:nref]
            # bit 4 = answered by the tensor-core pass; queries its certificate rejected are re-run by the sweep
            if not (metric == "euclidean_metric" and use_ts):     # 1/(1+d) + decay is not monotone in -d^2: sweeps
                assert sum(1 for f in f1 if f & 4) >= 0.7 * nq, ("tensor-core path was not taken", f1.tolist())
            if not (metric == "euclidean_metric" and sdt == "float16"):
                # (float16 euclidean scores are so coarse -- ulp 2.4e-4 at 0.44 -- that with k = 100 of 128 candidates the
                #  worst-case certificate legitimately refuses and the exact path answers; equality below still holds)
                assert sum(1 for f in f1 if f & 1) <= nq // 8, "too many exact-path fallbacks on the tensor path"
            assert np.array_equal(i0, i1c) and np.array_equal(s0, s1c) and np.array_equal(c0, c1c)
        # and against the oracle on a row subset small enough for it
        m.set_mask(None)
        m.set_timestamps(None)
        m.set_range(1000, 41000)
        i1, s1, _, f1 = m.query(q_np[:4], 10, metric)
        sub = V[1000:41000].cpu().numpy()
        for b in range(4):
            oi, os_ = K.rank(sub, q_np[b], 10, metric)
            assert list(i1[b] - 1000) == list(oi) and np.array_equal(s1[b], os_)
    finally:
        m.close()


# ---- batched pearson on the tensor cores (missing #6) ------------------------------------------------
@pytest.mark.parametrize("sdt", ["float16", "float32"])
@pytest.mark.parametrize("nq", [17, 300])
def test_batched_pearson_tensor_path(hb, nq, sdt):
    """pearson_correlation (hyperdb/ranking_algorithm.py:84-113) for a batch: V . (q - mean(q)) on tcgen05, 1 / (std_v d) in the
    epilogue, the query's 1 / std_q and the left-out mean correction in the certify step (csrc/certificate.cuh).  Rows carry
    a per-row offset (|mean| / std up to 1: the mean correction matters), a few rows are constant (NaN -> -inf, never a
    candidate) and one query is constant (every score NaN: exact path).  Must equal the streaming sweep bit for bit, and the
    oracle on a row range."""
    import torch
    n, d = 530_000, 136
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + nq)
    V = torch.randn(n, d, generator=g, device=dev) * (0.5 + torch.rand(n, 1, generator=g, device=dev))
    V = V + 1.0 * (torch.rand(n, 1, generator=g, device=dev) - 0.5)
    V[1234] = 0.75                                              # constant rows: std == 0
    V[400_001] = 0.0
    V = V.half() if sdt == "float16" else V.float()
    Q = torch.randn(nq, d, generator=g, device=dev) + 0.3
    Q[5] = 0.5 * (V[77_777].float() - 1.0) + 0.05 * Q[5]        # a strongly correlated row for one query
    Q[9] = 0.25                                                 # a constant query
    Q = Q.half() if sdt == "float16" else Q.float()
    ts = 1.7e9 + 10 * torch.rand(n, generator=g, device=dev, dtype=torch.float64)
    keep_bits = torch.randint(-2**31, 2**31 - 1, ((n + 31) // 32,), generator=g, device=dev, dtype=torch.int32)
    q_np = Q.cpu().numpy()
    metric = "pearson_correlation"
    m = hb.DeviceMatrix(V)
    try:
        for use_ts, use_mask, k in ((False, False, 10), (False, True, 100), (True, True, 10)):
            m.set_mask(keep_bits if use_mask else None)
            m.set_timestamps(ts if use_ts else None)
            if use_ts:
                m.refresh_decay()
            bias = 0.2 if use_ts else 0.0
            m.set_path(4)                                       # the streaming sweep with the exact-path repair: the reference answer
            i0, s0, c0, f0 = m.query(q_np, k, metric, bias)
            m.set_path(0)
            i1, s1, c1, f1 = m.query(q_np, k, metric, bias)
            if use_ts:       # a per-row decay term does not commute with the per-query 1 / std_q: sweeps
                assert not any(f & 4 for f in f1)
            elif k == 10:
                # (k = 100 of 128 candidates: the gap between the 100th and the 128th score is of the order of the tf32 band
                #  times max ||v|| / (std sqrt d) ~ 1.4 here, so many of those queries are legitimately re-run by the sweep,
                #  which clears bit 4; the equality below covers them)
                assert sum(1 for f in f1 if f & 4) >= 0.7 * nq, ("tensor-core path was not taken", f1.tolist())
                assert sum(1 for f in f1 if f & 1) <= max(2, nq // 8), "too many exact-path fallbacks on the tensor path"
            assert np.array_equal(i0, i1) and np.array_equal(c0, c1)
            assert np.array_equal(s0, s1, equal_nan=True)
        m.set_mask(None)
        m.set_timestamps(None)
        m.set_range(1000, 41000)
        i1, s1, _, f1 = m.query(q_np[:6], 10, metric)
        sub = V[1000:41000].cpu().numpy()
        for b in range(6):
            oi, os_ = K.rank(sub, q_np[b], 10, metric)
            assert list(i1[b] - 1000) == list(oi) and np.array_equal(s1[b], os_)
    finally:
        m.close()


@pytest.mark.skipif(os.environ.get("HDB_TC_MIXED") != "1",
                    reason="opt-in configuration (HDB_TC_MIXED=1, read once by the library): queries wider than the store on the tensor path")
@pytest.mark.parametrize("sdt,qdt", [("float16", "float32"), ("float16", "float64"), ("float32", "float64")])
@pytest.mark.parametrize("metric", ["dot_product", "cosine_similarity", "euclidean_metric", "pearson_correlation"])
def test_batched_tensor_path_wider_queries(hb, metric, sdt, qdt):
    """HDB_TC_MIXED=1: a float32 / float64 query tile over a float16 store (float64 over float32) -- what `HyperDB.query_batch`
    passes by default -- takes the tensor path with the canonical query ROUNDED to the storage precision as the B operand; the
    certificate carries that rounding (tests/test_emul_canonical.py proves the bound on the CPU).  Answers must equal the
    streaming sweep's bit for bit (same float64 / float32 result arithmetic), and the oracle's on a row range.
    NOT run on hardware yet (the round's GPU minutes were spent): the switch is off by default."""
    import torch
    n, d, nq = 530_000, 136, 40
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(77)
    V = torch.randn(n, d, generator=g, device=dev)
    V = V / V.norm(dim=1, keepdim=True)
    V = V.half() if sdt == "float16" else V.float()
    Q = torch.randn(nq, d, generator=g, device=dev, dtype=torch.float64)
    Q = Q / Q.norm(dim=1, keepdim=True)
    q_np = Q.cpu().numpy().astype(qdt)
    m = hb.DeviceMatrix(V)
    try:
        m.set_path(4)
        i0, s0, c0, _ = m.query(q_np, 10, metric)
        m.set_path(0)
        i1, s1, c1, f1 = m.query(q_np, 10, metric)
        assert sum(1 for f in f1 if f & 4) >= 0.7 * nq, ("tensor-core path was not taken", f1.tolist())
        assert np.array_equal(i0, i1) and np.array_equal(s0, s1) and np.array_equal(c0, c1)
        m.set_range(1000, 41000)
        i1, s1, _, _ = m.query(q_np[:4], 10, metric)
        sub = V[1000:41000].cpu().numpy()
        for b in range(4):
            oi, os_ = K.rank(sub, q_np[b], 10, metric)
            assert list(i1[b] - 1000) == list(oi) and np.array_equal(s1[b], os_)
    finally:
        m.close()


@pytest.mark.parametrize("vdt", ["f16", "f32", "f64"])
@pytest.mark.parametrize("metric", ["dot_product", "cosine_similarity", "euclidean_metric", "manhattan_distance", "hamming_distance",
                                    "jaccard_similarity", "pearson_correlation"])
def test_row_order_clustered_storage(hb, vdt, metric):
    """hdb_matrix_set_row_order: the rows are STORED clustered by a metadata key (here also: in a random order) while ids and
    ties follow the caller's numbering.  Every answer -- fused sweeps with 1 and 5 queries per pass, the wide candidate
    class, the exact path, top_k > 100, with mask + time decay, with a row offset -- must equal, bit for bit, the answer of
    the same rows stored in the caller's order; duplicates across clusters pin the tie rule (lower ORIGINAL index first)."""
    import zlib
    rng = np.random.default_rng(zlib.crc32(f"ro{vdt}{metric}".encode()))
    n, d = 20_011, 72
    dt = {"f16": np.float16, "f32": np.float32, "f64": np.float64}[vdt]
    V = rng.standard_normal((n, d)).astype(dt)
    for a, b in ((17, 15_000), (18, 9_001), (19_000, 23)):      # duplicates in different clusters
        V[b] = V[a]
    Q = rng.standard_normal((5, d)).astype(dt)
    Q[1] = V[17]
    Q[2] = V[23]
    ts = 1.7e9 + rng.uniform(0, 8, n)
    keep = rng.random(n) < 0.5
    cat = rng.integers(0, 6, n)
    for perm in (np.argsort(cat, kind="stable"), rng.permutation(n)):
        perm = perm.astype(np.uint32)                            # physical row p holds the caller's row perm[p]
        ref = hb.DeviceMatrix(V, row_offset=1000)
        m = hb.DeviceMatrix(np.ascontiguousarray(V[perm]), row_offset=1000)
        m.set_row_order(perm)
        try:
            for (k, use_ts, use_mask, path) in ((10, False, False, 0), (100, False, False, 0), (10, True, True, 0), (10, False, True, 1),
                                                (150, True, False, 0)):
                ref.set_mask(keep if use_mask else None)
                m.set_mask(keep[perm] if use_mask else None)
                ref.set_timestamps(ts if use_ts else None)
                m.set_timestamps(ts[perm] if use_ts else None)
                if use_ts:
                    ref.refresh_decay()
                    m.refresh_decay()
                bias = 0.25 if use_ts else 0.0
                ref.set_path(path)
                m.set_path(path)
                for q in (Q[0], Q[1], Q):
                    want = ref.query(q, k, metric, bias)
                    got = m.query(q, k, metric, bias)
                    assert np.array_equal(got[0], want[0]), (k, use_ts, use_mask, path, np.shape(q))
                    assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
            ref.set_path(0)
            m.set_path(0)
            with pytest.raises(Exception):
                m.append(V[:2])
            with pytest.raises(Exception):
                m.set_row_order(np.zeros(n, np.uint32))           # not a permutation
            m.set_row_order(None)                                 # back to the physical numbering
            got = m.query(Q[0], 5, metric)
            ref.set_mask(None); ref.set_timestamps(None)
            m.set_mask(None); m.set_timestamps(None)
            want = ref.query(Q[0], 5, metric)
            assert np.array_equal(np.sort(perm[got[0][0] - 1000].astype(np.int64)), np.sort(want[0][0] - 1000)) or metric in ("hamming_distance", "jaccard_similarity")
        finally:
            m.close()
            ref.close()


def test_tensor_path_record_overflow_is_repaired(hb):
    """when the CTA-private record buffers of the batched contraction overflow, every query of the batch must come
    back UNCERTIFIED (the poison survives concurrent appends) and the host call must repair it -- never a silently short candidate
    list.  HDB_TC_REC_CAP forces the overflow; the answers must equal the unforced run and the oracle."""
    import os
    rng = np.random.default_rng(4)
    n, d, b = 540_000, 64, 24
    V = rng.standard_normal((n, d)).astype(np.float16)
    Q = rng.standard_normal((b, d)).astype(np.float16)
    m = hb.DeviceMatrix(V)
    try:
        want = m.query(Q, 10, "cosine_similarity")
        assert (want[3] & hb._native.FLAG_TENSOR).any()                       # the batch did take the tensor-core path
        os.environ["HDB_TC_REC_CAP"] = "16"
        try:
            got = m.query(Q, 10, "cosine_similarity")
        finally:
            del os.environ["HDB_TC_REC_CAP"]
        assert not (got[3] & hb._native.FLAG_UNCERTIFIED).any()               # repaired by the host call ...
        assert not (got[3] & hb._native.FLAG_TENSOR).all()                    # ... i.e. (most of) the batch did not keep its tensor-path answer
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        for qi in (0, 7, 23):
            check_against_oracle(V, Q[qi], None, 0.0, 10, "cosine_similarity", got[0][qi], got[1][qi], None)
        again = m.query(Q, 10, "cosine_similarity")                          # the workspace recovers once the cap is lifted
        assert (again[3] & hb._native.FLAG_TENSOR).any() and np.array_equal(again[0], want[0])
    finally:
        m.close()
