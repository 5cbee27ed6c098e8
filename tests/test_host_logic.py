"""Host-side logic that needs no GPU: the query-cache key of the HyperDB shim (hyperdb/hyperdb.py:1368-1379), the lazy
result views of a sharded step, message sizes."""
import numpy as np

from hyperdb_b200.hyperdb import HyperDB
from hyperdb_b200.sharded import StepResult, packed_len


def test_cache_key_is_value_based_and_hashable():
    key = HyperDB._hashable_key
    q32 = np.array([0.25, -1.5, 3.0], np.float32)
    base = key(q32, 5, True, None, 0, None, "cosine_similarity", 5)
    assert base == key(q32.astype(np.float64), 5, True, None, 0, None, "cosine_similarity", 5)      # tuple(tolist()) would match too
    assert base == key(q32.tolist(), 5, True, None, 0, None, "cosine_similarity", 5)
    assert base != key(q32 + 1e-6, 5, True, None, 0, None, "cosine_similarity", 5)
    assert base != key(q32, 6, True, None, 0, None, "cosine_similarity", 5)
    assert base != key(q32.reshape(1, 3), 5, True, None, 0, None, "cosine_similarity", 5)           # shape is part of the key
    f1 = [("metadata", {"group": "a", "lang": "en"}), ("skip_doc", 3)]
    f2 = [("metadata", {"lang": "en", "group": "a"}), ("skip_doc", 3)]
    assert key(q32, 5, True, f1, 0.3, "timestamp", "dot_product", 5) == key(q32, 5, True, f2, 0.3, "timestamp", "dot_product", 5)
    hash(key(q32, 5, True, f1, 0.3, "timestamp", "dot_product", 5))
    assert key("some text", 5, True, None, 0, None, "cosine_similarity", 5)[0] == "some text"
    # -0.0 and +0.0 are the same tuple element in the reference's key; element order matters; float16 values compare by value
    z = np.array([0.0, 1.0, -2.0])
    assert key(z, 5, True, None, 0, None, "dot_product", 5) == key(np.array([-0.0, 1.0, -2.0]), 5, True, None, 0, None, "dot_product", 5)
    assert key(z, 5, True, None, 0, None, "dot_product", 5) != key(z[::-1].copy(), 5, True, None, 0, None, "dot_product", 5)
    assert key(z.astype(np.float16), 5, True, None, 0, None, "dot_product", 5) == key(z, 5, True, None, 0, None, "dot_product", 5)


def test_step_result_views():
    import torch
    b, k, w = 3, 4, 2
    block = torch.arange(2 * b * k + b + (w * b + 1) // 2, dtype=torch.int64)
    r = StepResult(block, b, k, w)
    idx, sc, cnt, flags = r
    assert idx.shape == (b, k) and sc.shape == (b, k) and sc.dtype == torch.float64 and cnt.shape == (b,) and flags.shape == (w, b)
    assert idx.data_ptr() + 8 * b * k == sc.data_ptr()          # ShardedMatrix.query copies [idx | score | count] in one piece
    assert r[3] is flags and len(r) == 4
    assert packed_len(b, k) == 2 * b * k + b + 2


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c2_cosine_b1",
                          "--rows", "20000", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "queries/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["steps"] == 1 and j["config"]["workload"] == "c2_cosine_b1"
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def _shim_without_gpu(cache_size=256):
    """A HyperDB shim whose ranking is stubbed out: only the cache / bookkeeping logic runs (no CUDA needed)."""
    db = HyperDB(cache_size=cache_size)
    db.documents = [{"id": i} for i in range(3)]
    db.vectors = np.zeros((3, 4), np.float32)
    calls = []

    def fake_execute(query_input, top_k, *rest):
        calls.append(query_input if isinstance(query_input, str) else tuple(np.asarray(query_input).ravel().tolist()))
        return [("doc", 1.0, 0)][:top_k]
    db._execute_query = fake_execute
    return db, calls


def test_cache_counters_like_the_reference_tests():
    """tests/test_hyperdb.py:708-736 of the reference, against the shim's cache (ranking stubbed)."""
    db, calls = _shim_without_gpu()
    db.query("Abra")
    assert db.get_cache_size_and_info()["cache_info"]["hits"] == 0 and db.get_cache_size_and_info()["cache_info"]["misses"] == 1
    db.query("Abra")
    info = db.get_cache_size_and_info()["cache_info"]
    assert info["hits"] == 1 and info["misses"] == 1 and calls == ["Abra"]
    # :724-728 -- the tests swap the cache object for a cachetools.LRUCache and read maxsize back
    cachetools = __import__("pytest").importorskip("cachetools")
    db.lru_cache = cachetools.LRUCache(maxsize=128)
    assert db.get_cache_size_and_info()["cache_info"]["maxsize"] == 128
    # :730-737 -- eviction
    db.lru_cache = cachetools.LRUCache(maxsize=2)
    for i in range(3):
        db.query(f"Query {i}")
    assert db.get_cache_size_and_info()["cache_info"]["currsize"] == 2


def test_builtin_lru_evicts_least_recently_used():
    db, calls = _shim_without_gpu(cache_size=2)
    q = [np.full(4, float(i), np.float32) for i in range(3)]
    db.query(q[0]); db.query(q[1]); db.query(q[0])            # q0 is now the most recently used
    db.query(q[2])                                             # evicts q1
    assert len(db.lru_cache) == 2 and db.get_cache_size_and_info()["cache_info"]["maxsize"] == 2
    n = len(calls)
    db.query(q[0])
    assert len(calls) == n                                     # still cached
    db.query(q[1])
    assert len(calls) == n + 1                                 # was evicted: ranked again
    db.clear_cache()
    assert len(db.lru_cache) == 0 and db.cache_hits == 0 and db.cache_misses == 0
    off, _ = _shim_without_gpu(cache_size=0)
    off.query(q[0]); off.query(q[0])
    assert len(off.lru_cache) == 0 and off.cache_hits == 0


def test_bench_mask_helpers():
    """bench.py's synthetic metadata: categories are reproducible per global id and uniform; the packed keep bits are NumPy's
    little-endian packbits (bit i of word i/32 = row i, the layout hdb_matrix_set_mask documents)."""
    import torch
    import bench
    cat = bench.gen_category_torch(1000, 201_000, "cpu")
    assert cat.min() >= 0 and cat.max() < bench.MASK_CATEGORIES
    counts = torch.bincount(cat, minlength=bench.MASK_CATEGORIES).double()
    assert (counts / counts.sum() - 1.0 / bench.MASK_CATEGORIES).abs().max() < 0.01
    assert torch.equal(cat[500:700], bench.gen_category_torch(1500, 1700, "cpu"))      # a shard sees the same documents
    keep = bench.gen_keep_torch(0, 100_003, "cpu")
    assert abs(keep.double().mean().item() - len(bench.MASK_KEPT) / bench.MASK_CATEGORIES) < 0.01
    bits = bench.pack_keep_bits(keep).numpy().view(np.uint32)
    want = np.packbits(keep.numpy(), bitorder="little")
    want = np.concatenate([want, np.zeros((-len(want)) % 4, np.uint8)]).view(np.uint32)
    assert np.array_equal(bits, want)


def test_shim_filter_semantics_without_gpu():
    """hyperdb/hyperdb.py:1474-1481 applies only the FIRST skip_doc filter; metadata filters intersect; the subset tag is
    stable under dict order (it keys the device-side subset / decay cache)."""
    db, _ = _shim_without_gpu()
    db.documents = [{"id": i, "g": "ab"[i % 2], "h": i % 3} for i in range(12)]
    db.metadata_keys = ["g", "h"]
    db._invalidate_columns()
    lo, hi, keep, tag = db._apply_filters([("skip_doc", 2), ("skip_doc", -3)])
    assert (lo, hi, keep) == (2, 12, None)
    lo, hi, keep, tag = db._apply_filters([("skip_doc", -3), ("metadata", {"g": "a"}), ("metadata", {"h": 0})])
    assert (lo, hi) == (0, 9) and list(np.flatnonzero(keep)) == [0, 6]
    assert tag == db._apply_filters([("skip_doc", -3), ("metadata", {"g": "a"}), ("metadata", {"h": 0})])[3]
    assert db._apply_filters([("metadata", {"g": "a", "h": 0})])[3] == db._apply_filters([("metadata", {"h": 0, "g": "a"})])[3]
    import pytest
    with pytest.raises(Exception):
        db._apply_filters([("skip_doc", 12)])
    with pytest.raises(ValueError):
        db._apply_filters([("metadata", {"missing": 1})])
    with pytest.raises(ValueError):
        db._apply_filters([("nope", 1)])
    # documents.index semantics: the first EQUAL document, whatever the dict order
    db.documents[7] = {"h": 1, "g": "b", "id": 1}
    db._invalidate_columns()
    assert db._first_index(7) == 1 and db._first_index(8) == 8


def test_shim_host_copy_is_lazy():
    """`HyperDB.vectors` (the attribute the reference exposes) is concatenated only when somebody reads it: an `add` appends
    a chunk instead of copying the whole matrix (hyperdb/hyperdb.py:504-509 does np.concatenate per call)."""
    db, _ = _shim_without_gpu()
    a, b = np.ones((3, 4), np.float32), np.zeros((2, 4), np.float32)
    db._chunks = [a, b]
    assert len(db._chunks) == 2
    v = db.vectors
    assert v.shape == (5, 4) and len(db._chunks) == 1 and np.array_equal(v[:3], a)
    db.vectors = None
    assert db.vectors is None and db._n == 0


def test_native_digest_equals_numpy_statement():
    """hdb_query_digest_host (the C++ twin of the device kernel's formula, no GPU needed) == hyperdb.query_digest_host (NumPy) for
    every query dtype, incl. -0.0, subnormals, infinities and long vectors; a batch is digested row by row."""
    import ctypes as C
    from hyperdb_b200 import _native as N
    from hyperdb_b200.hyperdb import query_digest_host
    rng = np.random.default_rng(8)
    for dt, code in ((np.float16, 0), (np.float32, 1), (np.float64, 2)):
        for d in (1, 7, 768, 4099):
            Q = rng.standard_normal((3, d)).astype(dt)
            Q[0, 0] = -0.0
            Q[1, -1] = np.inf
            Q[2, d // 2] = np.finfo(dt).smallest_subnormal
            out = (C.c_uint64 * 6)()
            N.check(N.lib().hdb_query_digest_host(C.c_void_p(Q.ctypes.data), code, 3, d, out))
            for b in range(3):
                assert (int(out[2 * b]), int(out[2 * b + 1])) == query_digest_host(Q[b]), (dt, d, b)


def test_bench_hyperdb_query_extra_on_the_stand_in(monkeypatch, capsys):
    """bench.py's `hyperdb_query_c2` extra (HyperDB.query end to end, plain and with a metadata predicate) run on the CPU with the
    oracle-backed stand-in for the device matrix and a small row count: the record it returns, and nothing on stdout (the
    driver parses ONE JSON line; the shim's INFO prints must stay in the buffer)."""
    import torch
    import bench
    import hyperdb_b200.hyperdb as H
    from shim_fakes import FakeDeviceMatrix
    monkeypatch.setattr(H, "DeviceMatrix", FakeDeviceMatrix)
    monkeypatch.setitem(bench.WORKLOADS, "c2_cosine_b1", dict(bench.WORKLOADS["c2_cosine_b1"], n=4000))
    r = bench.hyperdb_query_extra(torch.device("cpu"), steps=4, warmup=1)
    assert capsys.readouterr().out == ""
    assert r["workload"] == "hyperdb_query_c2" and r["planted_first"] is True
    assert r["value"] > 0 and r["filtered_value"] > 0 and r["unit"] == "queries/s"
    assert "of 4000 documents kept" in r["filter"]
    import json
    json.dumps(r)
