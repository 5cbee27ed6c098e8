"""The HyperDB shim's HOST logic against the LIVE HyperDB class of the reference (imported from /root/reference when it is there,
third-party imports stubbed as in tests/golden/make_golden.py; skipped elsewhere): randomized filters (skip_doc of both signs,
metadata, several filters at once), recency of both signs, every brute-force metric, the three storage precisions, top_k beyond
the filtered set, return_similarities on / off.  The device matrix is the oracle-backed stand-in of tests/shim_fakes.py, so this
runs on the CPU; the GPU twin of the comparison is tests/test_gpu_hyperdb_shim.py on the committed fixtures."""
import contextlib
import io
import os
import zlib

import numpy as np
import pytest

from shim_fakes import FakeDeviceMatrix

pytestmark = pytest.mark.skipif(not os.path.exists("/root/reference/hyperdb/hyperdb.py"), reason="the reference sources are not on this machine")

METRICS = ['dot_product', 'cosine_similarity', 'euclidean_metric', 'manhattan_distance', 'jaccard_similarity', 'pearson_correlation']


@pytest.fixture(scope="module")
def real_class():
    import make_golden
    with contextlib.redirect_stdout(io.StringIO()):
        return make_golden.load_hyperdb_class()


@pytest.mark.parametrize("precision", ["float16", "float32", "float64"])
def test_shim_equals_the_live_hyperdb_class(real_class, monkeypatch, precision):
    import hyperdb_b200.hyperdb as H
    monkeypatch.setattr(H, "DeviceMatrix", FakeDeviceMatrix)
    rng = np.random.default_rng(zlib.crc32(("liveshim" + precision).encode()))
    n, d = 90, 24
    V = rng.standard_normal((n, d)).astype(precision)
    ts = (1.7e9 + rng.uniform(0, 3.0, n)).round(3)
    groups = ["a", "b", "c", "d"]
    docs = [{"id": i, "group": groups[int(rng.integers(0, 4))], "meta": {"lang": ["en", "fr"][i % 2]}, "timestamp": float(ts[i])} for i in range(n)]

    def embed(documents, **kw):
        ids = [doc["id"] for doc in documents]
        return V[ids], list(ids), {i: 1 for i in ids}

    with contextlib.redirect_stdout(io.StringIO()):
        real = real_class.HyperDB(documents=None, embedding_function=embed, ann_metric="hamming", metadata_keys=["group", "meta.lang", "timestamp"],
                                  fp_precision=precision)
        real.add(docs)
        mine = H.HyperDB(documents=docs, vectors=V, metadata_keys=["group", "meta.lang", "timestamp"], fp_precision=precision, sharded=False)
    assert real.vectors.dtype == V.dtype and real.vectors.shape == V.shape
    compared = 0
    for trial in range(70):
        metric = METRICS[trial % len(METRICS)]
        filters = []
        r = rng.random()
        if r < 0.35:
            filters.append(("skip_doc", int(rng.integers(1, 40)) * (1 if rng.random() < 0.5 else -1)))
        if rng.random() < 0.5:
            filters.append(("metadata", {"group": groups[int(rng.integers(0, 4))]}))
        if rng.random() < 0.25:
            filters.append(("metadata", {"meta.lang": "fr"}))
        if rng.random() < 0.15:
            filters.append(("skip_doc", 5))                                # a second skip_doc: only the first one acts
        rng.shuffle(filters)
        bias = float(rng.choice([0, 0, 0.3, 1.5, -0.4]))
        kw = dict(top_k=int(rng.choice([1, 4, 9, 200])), filters=filters or None, recency_bias=bias,
                  timestamp_key="timestamp" if (bias or rng.random() < 0.2) else None, metric=metric,
                  return_similarities=bool(rng.random() < 0.8))
        q = rng.standard_normal(d).astype(rng.choice(["float32", "float64"]))
        out = []
        for db in (real, mine):
            with np.errstate(all="ignore"), contextlib.redirect_stdout(io.StringIO()):
                try:
                    out.append(("ok", db.query(q, **kw)))
                except Exception as e:                                    # noqa: BLE001 -- both sides must refuse alike
                    out.append(("err", type(e).__name__))
            db.lru_cache.clear()
        (ka, a), (kb, b) = out
        assert ka == kb, (trial, kw, a, b)
        if ka == "err":
            assert a == b, (trial, kw, a, b)
            continue
        assert len(a) == len(b), (trial, kw)
        if not kw["return_similarities"]:
            if metric != "jaccard_similarity":                             # (bit-count ratios tie; without scores the groups are unknown)
                assert [x["id"] for x in a] == [x["id"] for x in b], (trial, kw)
        else:
            sa = np.array([float(np.asarray(x[1]).reshape(-1)[0]) for x in a])
            sb = np.array([float(x[1]) for x in b])
            np.testing.assert_allclose(sb, sa, rtol=1e-12 if metric in ("dot_product", "cosine_similarity") else 0, atol=0, err_msg=str((trial, kw)))
            assert all(x[2] == x[0]["id"] for x in a) and all(x[2] == x[0]["id"] for x in b)
            # the reference's order among EQUAL scores is unspecified (argpartition / argsort); the shim's is ascending id.  Compare the
            # id SETS of every group of equal scores; the last group may be cut by top_k, where even the membership is unspecified
            ia, ib = [x[0]["id"] for x in a], [x[0]["id"] for x in b]
            cut = len(a) == kw["top_k"]
            start = 0
            while start < len(sa):
                end = start
                while end < len(sa) and sa[end] == sa[start]:
                    end += 1
                if not (cut and end == len(sa) and (end - start > 1 or metric == "jaccard_similarity")):    # (jaccard: ties with rows left out)
                    assert set(ia[start:end]) == set(ib[start:end]), (trial, kw, start, end)
                assert ib[start:end] == sorted(ib[start:end]), (trial, kw)       # ties to the lower id on our side
                start = end
        compared += 1
    assert compared >= 50
