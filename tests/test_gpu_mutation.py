"""Incremental ingest of a resident shard: hdb_matrix_append / hdb_matrix_remove_rows /
hdb_matrix_reserve against a freshly uploaded matrix of the same rows and against the oracle.  The reference
re-materialises `self.vectors` on every add (hyperdb/hyperdb.py:504-509) and remove_document (:718-728); the resident form
must give the same answers as a rebuild."""
import numpy as np
import pytest

from oracle import canonical as K

pytestmark = pytest.mark.gpu

METRICS = ("dot_product", "cosine_similarity", "euclidean_metric", "manhattan_distance", "hamming_distance",
           "jaccard_similarity", "pearson_correlation")


@pytest.fixture(scope="module")
def hb():
    import hyperdb_b200
    return hyperdb_b200


def _data(n, d, dt, seed):
    rng = np.random.default_rng(seed)
    V = (rng.standard_normal((n, d)) * rng.uniform(0.3, 2.0, (n, 1)) + rng.uniform(-0.5, 0.5, (n, 1))).astype(dt)
    q = rng.standard_normal(d).astype(dt)
    return V, q


def _same_as_fresh(hb, m, V, q, k=10, exact_scores=True):
    fresh = hb.DeviceMatrix(V)
    try:
        assert m.shape == fresh.shape == V.shape
        for metric in METRICS:
            i0, s0, c0, _ = fresh.query(q, k, metric)
            i1, s1, c1, _ = m.query(q, k, metric)
            assert list(i1[0]) == list(i0[0]), metric
            assert np.array_equal(s1[0], s0[0]), metric
            assert c1[0] == c0[0]
            assert np.array_equal(m.scores(q, metric), fresh.scores(q, metric), equal_nan=True), metric
        oi, os_ = K.rank(V, q, k, "euclidean_metric")
        i1, s1, _, _ = m.query(q, k, "euclidean_metric")
        assert list(i1[0]) == list(oi) and np.array_equal(s1[0], os_)
    finally:
        fresh.close()


@pytest.mark.parametrize("dt", [np.float16, np.float32, np.float64])
@pytest.mark.parametrize("warm", [False, True], ids=["cold-columns", "bits+pearson-built"])
def test_append_matches_rebuild(hb, dt, warm):
    V, q = _data(5000, 72, dt, 5)
    m = hb.DeviceMatrix(V[:1200])
    try:
        if warm:       # lazily built sign bits and pearson columns must be extended for the new rows
            m.query(q, 5, "hamming_distance")
            m.query(q, 5, "pearson_correlation")
        m.append(V[1200:1300])                 # exceeds the capacity: geometric growth
        _same_as_fresh(hb, m, V[:1300], q)
        m.reserve(4000)
        m.append(V[1300:3999])
        m.append(V[3999:4000])                 # exactly full
        m.append(V[4000:])                     # grows again
        _same_as_fresh(hb, m, V, q)
    finally:
        m.close()


def test_append_resets_row_state_and_rejects_nan(hb):
    V, q = _data(900, 40, np.float32, 6)
    m = hb.DeviceMatrix(V[:600])
    try:
        m.set_mask(np.arange(600) % 2 == 0)
        m.set_range(10, 500)
        bad = V[600:700].copy()
        bad[17, 3] = np.nan
        with pytest.raises(ValueError):
            m.append(bad)
        assert m.shape == (600, 40)
        i, s, c, _ = m.query(q, 10, "cosine_similarity")            # unchanged shard, filters still in force
        keep = (np.arange(600) % 2 == 0) & (np.arange(600) >= 10) & (np.arange(600) < 500)
        oi, _ = K.rank(V[:600], q, 10, "cosine_similarity", keep=keep)
        assert list(i[0]) == list(oi)
        m.append(V[600:])
        assert m.n_kept == 900                                       # mask and range were reset
        _same_as_fresh(hb, m, V, q)
    finally:
        m.close()


@pytest.mark.parametrize("dt,d", [(np.float16, 33), (np.float16, 64), (np.float32, 100), (np.float64, 7)])
def test_remove_rows_matches_rebuild(hb, dt, d):
    V, q = _data(7000, d, dt, 7)
    rng = np.random.default_rng(8)
    m = hb.DeviceMatrix(V)
    try:
        m.query(q, 5, "jaccard_similarity")
        m.query(q, 5, "pearson_correlation")
        keep = np.ones(len(V), bool)
        for frac in (0.001, 0.3, 0.5):
            alive = np.flatnonzero(keep)
            drop_local = rng.choice(len(alive), max(1, int(frac * len(alive))), replace=False)
            drop_local = np.concatenate([drop_local, drop_local[:3]])        # duplicates are allowed
            m.remove_rows(drop_local)
            keep[alive[drop_local]] = False
            _same_as_fresh(hb, m, V[keep], q)
        with pytest.raises(Exception):
            m.remove_rows([m.shape[0]])                                       # out of range
        m.remove_rows([0, -1])                                                # first and last (negative = from the end)
        alive = np.flatnonzero(keep)
        keep[alive[[0, -1]]] = False
        _same_as_fresh(hb, m, V[keep], q)
        m.append(V[:50])                                                      # append after a removal
        _same_as_fresh(hb, m, np.concatenate([V[keep], V[:50]]), q)
    finally:
        m.close()


def test_remove_with_timestamps_keeps_them_aligned(hb):
    V, q = _data(3000, 48, np.float32, 9)
    ts = 1.7e9 + np.random.default_rng(10).uniform(0, 5.0, len(V))
    m = hb.DeviceMatrix(V)
    try:
        m.set_timestamps(ts)
        drop = np.arange(0, len(V), 3)
        m.remove_rows(drop)
        keep = np.ones(len(V), bool)
        keep[drop] = False
        m.refresh_decay()                      # the decay reference changed with the row set
        i, s, _, _ = m.query(q, 10, "cosine_similarity", 0.5)
        oi, os_ = K.rank(V[keep], q, 10, "cosine_similarity", ts[keep], 0.5)
        assert list(i[0]) == list(oi)
        np.testing.assert_allclose(s[0], os_, rtol=1e-5)
    finally:
        m.close()


def test_adopted_memory_is_immutable(hb):
    import torch
    t = torch.randn(100, 16, device="cuda")
    m = hb.DeviceMatrix(t)
    try:
        with pytest.raises(ValueError):
            m.append(np.zeros((1, 16), np.float32))
        with pytest.raises(ValueError):
            m.remove_rows([0])
    finally:
        m.close()


def test_hyperdb_add_remove_incremental(hb):
    from hyperdb_b200.hyperdb import HyperDB
    V, q = _data(400, 24, np.float32, 11)
    docs = [{"id": i} for i in range(len(V))]
    db = HyperDB(documents=docs[:100], vectors=V[:100])
    try:
        db.add(docs[100:300], vectors=V[100:300])
        db.add(docs[300:], vectors=V[300:])
        db.remove_document([5, 250, 399])
        keep = np.ones(len(V), bool)
        keep[[5, 250, 399]] = False
        res = db.query(q, top_k=7, metric="euclidean_metric")
        want = np.flatnonzero(keep)
        oi, os_ = K.rank(V[keep].astype(np.float64), q.astype(np.float64), 7, "euclidean_metric")
        assert [r[0]["id"] for r in res] == [int(want[j]) for j in oi]
        np.testing.assert_allclose([r[1] for r in res], os_, rtol=1e-12)
        assert db.size() == 397
    finally:
        db.close()
