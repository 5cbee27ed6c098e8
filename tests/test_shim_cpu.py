"""The HOST logic of hyperdb_b200.hyperdb.HyperDB without a GPU: the device matrix is replaced by an oracle-backed stand-in
(tests/shim_fakes.py), everything else is the product's code.  Pinned to runs of the REAL HyperDB class
(tests/golden/hyperdb_tail.npz: recency applied twice, skip_doc ranges, metadata masks, float64 query)."""
import numpy as np
import pytest

import golden_io as G
from shim_fakes import FakeDeviceMatrix


@pytest.fixture
def shim(monkeypatch):
    import hyperdb_b200.hyperdb as H
    monkeypatch.setattr(H, "DeviceMatrix", FakeDeviceMatrix)
    FakeDeviceMatrix.calls = []
    return H


def _tail_db(H, **kw):
    z, specs = G.load_tail()
    V, ts, queries, groups = z["V"], z["ts"], z["queries"], z["groups"]
    docs = [{"id": i, "group": str(groups[i]), "timestamp": float(ts[i])} for i in range(len(V))]
    db = H.HyperDB(documents=docs, vectors=V, metadata_keys=["group", "timestamp"], fp_precision="float32", sharded=False, **kw)
    return db, z, specs, queries


@pytest.mark.parametrize("cluster_by", [None, "group"])
def test_shim_matches_real_hyperdb_on_cpu(shim, cluster_by, capsys):
    """Every spec x query of the fixtures produced by the real class: ids, scores, source indices -- plain and with the
    store clustered by the filtered key (row order bookkeeping, masks / ranges / timestamps in physical order)."""
    db, z, specs, queries = _tail_db(shim, cluster_by=cluster_by)
    assert (db._perm is not None) == (cluster_by is not None)
    for si, spec in enumerate(specs):
        filters = [tuple(f) if f[0] == "skip_doc" else (f[0], f[1]) for f in spec["filters"]] if spec["filters"] else None
        for qi in range(len(queries)):
            res = db.query(queries[qi], top_k=spec["top_k"], filters=filters, recency_bias=spec["recency_bias"],
                           timestamp_key="timestamp" if spec["recency_bias"] else None, metric=spec["metric"])
            assert [doc["id"] for doc, _s, _i in res] == list(z[f"ids_{si}_{qi}"]), (si, qi)
            blas = spec["metric"] in ("dot_product", "cosine_similarity")
            np.testing.assert_allclose([s for _d, s, _i in res], z[f"sc_{si}_{qi}"], rtol=1e-12 if blas else 1e-15)
            assert all(i == d["id"] for d, _s, i in res)
    assert "Bruteforce method used instead" in capsys.readouterr().out


def test_subset_and_decay_state_is_reused(shim):
    """The row subset and the two-stage decay are rebuilt only when (filters, timestamp_key, recency_bias) change; the
    timestamp column is extracted from the documents and staged once per key."""
    db, z, specs, queries = _tail_db(shim)
    kw = dict(top_k=5, filters=[("metadata", {"group": str(z["groups"][0])}), ("skip_doc", 3)], recency_bias=0.4, timestamp_key="timestamp")
    db.lru_cache.maxsize = 0
    db.query(queries[0], **kw)
    first = list(FakeDeviceMatrix.calls)
    assert [c[0] for c in first].count("set_mask") == 1 and [c[0] for c in first].count("stage1") == 1
    FakeDeviceMatrix.calls.clear()
    db.query(queries[1], **kw)
    assert [c[0] for c in FakeDeviceMatrix.calls] == ["query"]                       # nothing but the ranking
    db.query(queries[1], top_k=5, recency_bias=0.1, timestamp_key="timestamp")       # other settings: rebuilt
    names = [c[0] for c in FakeDeviceMatrix.calls]
    assert "set_mask" in names and "stage1" in names and "decay" in names
    assert len(db._ts_cache) == 1 and len(db._ts_dev) == 1


def test_mutations_and_errors_on_cpu(shim):
    db, z, specs, queries = _tail_db(shim)
    n = db.size()
    rng = np.random.default_rng(1)
    extra = rng.standard_normal((2, z["V"].shape[1])).astype(np.float32)
    db.add([{"id": n, "group": "zz", "timestamp": 1.0}, {"id": n + 1, "group": "zz", "timestamp": 2.0}], vectors=extra)
    assert db.size() == n + 2 and len(db._chunks) == 2                                # the host copy grew by one chunk
    assert db.query(extra[1], top_k=1)[0][0]["id"] == n + 1
    assert db.vectors.shape[0] == n + 2 and len(db._chunks) == 1                      # concatenated when somebody looks
    db.remove_document(n + 1)
    assert db.size() == n + 1 and db.query(extra[1], top_k=1)[0][0]["id"] != n + 1
    with pytest.raises(ValueError):
        db.query(queries[0], metric="nope")
    with pytest.raises(ValueError):
        db.query(queries[0][:5])
    with pytest.raises(ValueError):
        db.query(queries[0], recency_bias=0.5, timestamp_key="missing")
    with pytest.raises(ValueError, match="empty sequence"):
        db.query(queries[0], top_k=0)
    with pytest.raises(Exception):
        db.query(queries[0], filters=[("skip_doc", n + 5)])
    assert db.query(queries[0], top_k=3, filters=[("metadata", {"group": "no such group"})]) == []
    got = db.query_batch(queries[:3], top_k=4)
    for i in range(3):
        assert [x[0]["id"] for x in got[i]] == [x[0]["id"] for x in db.query(queries[i], top_k=4)]
    # native_dtype: the tile reaches the engine in its own dtype (float32 queries over the float32 store: the tensor-core
    # case) and the scores are those of the ranking function called with that array; default: float64, as `query`
    from shim_fakes import FakeDeviceMatrix
    from oracle import canonical as K
    Q32 = np.asarray(queries[:3], np.float32)
    seen = []
    real_query = FakeDeviceMatrix.query
    FakeDeviceMatrix.query = lambda self, q, *a, **kw: (seen.append(np.asarray(q).dtype), real_query(self, q, *a, **kw))[1]
    try:
        nat = db.query_batch(Q32, top_k=4, metric="euclidean_metric", native_dtype=True)
        dflt = db.query_batch(Q32, top_k=4, metric="euclidean_metric")
        db.query_batch(Q32.astype(np.int32), top_k=1, native_dtype=True)               # not a float tile: float64 as before
    finally:
        FakeDeviceMatrix.query = real_query
    assert seen == [np.dtype(np.float32), np.dtype(np.float64), np.dtype(np.float64)]
    V = np.asarray(db.vectors)
    for i in range(3):
        oi, os_ = K.rank(V, Q32[i], 4, "euclidean_metric")
        assert [x[2] for x in nat[i]] == list(oi) and [x[1] for x in nat[i]] == list(os_)
        oi, os_ = K.rank(V, Q32[i].astype(np.float64), 4, "euclidean_metric")
        assert [x[2] for x in dflt[i]] == list(oi) and [x[1] for x in dflt[i]] == list(os_)


def _sharded_shim_worker(rank, world, port, out_dir):
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.join(here, "golden"), os.path.dirname(here), os.path.join(os.path.dirname(here), "local-hyperdb_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import hyperdb_b200.hyperdb as H
    import hyperdb_b200.sharded as S
    from shim_fakes import FakeDeviceMatrix, FakeEngine
    H.DeviceMatrix = FakeDeviceMatrix
    S.CudaEngine = FakeEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(41)
    n, d = 1501, 24
    V = rng.standard_normal((n, d)).astype(np.float32)
    V[1200] = V[5]                                               # a cross-shard tie
    ts = 1.7e9 + rng.uniform(0, 20, n)
    docs = [{"id": i, "group": "xyz"[i % 3], "timestamp": float(ts[i])} for i in range(n)]
    Q = rng.standard_normal((3, d)).astype(np.float32)
    Q[1] = V[5]
    kw = dict(metadata_keys=["group", "timestamp"], fp_precision="float32", device=0)
    single = H.HyperDB(documents=docs, vectors=V, sharded=False, **kw)
    shard = H.HyperDB(documents=docs, vectors=V, **kw)
    clus = H.HyperDB(documents=docs, vectors=V, cluster_by="group", **kw)
    assert shard._world == world and shard._sm is not None and shard._matrix.shape[0] < n and clus._perm is not None
    cases = [dict(top_k=10), dict(top_k=10, metric="hamming_distance"), dict(top_k=120, metric="euclidean_metric"),
             dict(top_k=7, filters=[("metadata", {"group": "y"}), ("skip_doc", -200)], recency_bias=0.3, timestamp_key="timestamp"),
             dict(top_k=5, filters=[("skip_doc", 900)], metric="manhattan_distance")]
    for kwq in cases:
        for i in range(len(Q)):
            a = single.query(Q[i], **kwq)
            for other in (shard, clus):
                b = other.query(Q[i], **kwq)
                assert [x[0]["id"] for x in a] == [x[0]["id"] for x in b], (kwq, i)
                np.testing.assert_allclose([x[1] for x in a], [x[1] for x in b], rtol=1e-14)
    got = shard.query_batch(Q, top_k=4, metric="manhattan_distance")
    for i in range(len(Q)):
        assert [x[0]["id"] for x in got[i]] == [x[0]["id"] for x in single.query(Q[i], top_k=4, metric="manhattan_distance")]
    extra = rng.standard_normal((3, d)).astype(np.float32)
    for db in (single, shard, clus):
        db.add([{"id": n + j, "group": "x", "timestamp": 1.7e9} for j in range(3)], vectors=extra)
        db.remove_document(5)
    a = single.query(Q[1], top_k=6)
    for other in (shard, clus):
        b = other.query(Q[1], top_k=6)
        assert [x[0]["id"] for x in a] == [x[0]["id"] for x in b] and a[0][0]["id"] == 1200
        assert other.query(extra[2], top_k=1)[0][0]["id"] == n + 2
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


def test_sharded_shim_under_gloo(tmp_path):
    """HyperDB over a row-sharded store, world_size 2 under gloo on the CPU (oracle-backed shards): plain and clustered, every
    filter / recency combination, a batch, add and remove_document equal the single-shard store."""
    import os
    import torch.multiprocessing as mp
    mp.spawn(_sharded_shim_worker, args=(2, 29871 + (os.getpid() % 100), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
