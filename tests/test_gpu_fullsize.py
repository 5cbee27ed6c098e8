"""Parity at BASELINE.json's full sizes, through size-independent properties (the oracle cannot score 10M+ rows):
  * planted rows (q + growing noise, one duplicated at a higher index) must come back in exactly the planted
    order with the duplicate right after its twin (ties -> lower index);
  * the returned scores must equal, bit for bit where NumPy is deterministic, the oracle's scores of those rows
    copied back to the host, and no row of a random sample may beat the k-th returned score;
  * idempotence, descending order, sharded (2 shards on one GPU + hdb_merge_topk) == unsharded;
  * with a keep mask, dropped rows never appear; the decay reference is the max over kept rows only.
Data are generated on the device (torch) like bench.py does."""
import numpy as np
import pytest

from oracle import canonical as K

pytestmark = pytest.mark.gpu

CONFIGS = [  # name, rows, dim, dtype, metrics, k
    ("C2", 1_000_000, 384, "float32", ("cosine_similarity",), 10),
    ("C3", 10_000_000, 768, "float16", ("dot_product", "cosine_similarity"), 10),
    ("C5", 5_000_000, 1024, "float32", ("euclidean_metric", "manhattan_distance", "hamming_distance"), 10),
    ("C4", 100_000_000, 384, "float16", ("cosine_similarity",), 100),
]
EXACT = {"float16": True, "float32": False}


def plant(rows, q, k, rng, torch):
    """overwrite k+1 rows: planted[i] = unit(q + 0.02*(i+1)*noise); the last one duplicates planted[3] at a higher index"""
    n, d = rows.shape
    pos = np.sort(rng.choice(n - 1000, size=k + 1, replace=False))
    qf = torch.as_tensor(q.astype(np.float32), device=rows.device)
    order = rng.permutation(k)                      # rank i sits at row pos[order[i]]
    planted = []
    for i in range(k):
        noise = torch.as_tensor(rng.standard_normal(d).astype(np.float32), device=rows.device)
        noise = noise - (noise @ qf) / (qf @ qf) * qf       # orthogonal to q and unit length:
        noise = noise / noise.norm() * qf.norm()            # cos(q, v) = 1/sqrt(1+a^2), strictly decreasing in a
        v = qf + 0.05 * (i + 1) * noise
        v = v / v.norm()
        rows[int(pos[order[i]])] = v.to(rows.dtype)
        planted.append(int(pos[order[i]]))
    twin_of = planted[3]
    dup = int(pos[k]) if pos[k] > twin_of else None
    if dup is None:                                 # make sure the duplicate has the HIGHER index
        dup = n - 7
    rows[dup] = rows[twin_of]
    want = planted[:4] + [dup] + planted[4:]
    return want


@pytest.mark.parametrize("name,n,d,dtype,metrics,k", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_fullsize_properties(name, n, d, dtype, metrics, k):
    import torch
    import bench
    import hyperdb_b200 as hb
    from hyperdb_b200.sharded import CudaEngine, shard_bounds
    free, _total = torch.cuda.mem_get_info()
    if free < n * d * bench.ITEM[dtype] * 1.35 + (4 << 30):
        pytest.skip("not enough free HBM for this configuration")
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(12345)
    rows = bench.gen_rows_torch(0, n, d, dtype, dev, seed=0)
    q = bench.gen_queries(1, d, dtype, seed=99)[0]
    want = plant(rows, q, k - 1 if k <= 10 else 20, rng, torch)
    m = hb.DeviceMatrix(rows)
    try:
        for metric in metrics:
            idx, sc, cnt, flags = m.query(q, k, metric)
            idx, sc = idx[0], sc[0]
            assert cnt[0] == k
            if metric != "hamming_distance":        # sign bits of the planted rows are not ordered by construction
                assert list(idx[:len(want)]) == want, (metric, idx, want)
            assert np.all(np.diff(sc) <= 0)
            # scores: the oracle on the returned rows (copied back) must reproduce them
            got_rows = rows[torch.as_tensor(idx, device=dev)].cpu().numpy()
            osc = K.total_scores(got_rows, q, metric)
            if EXACT[dtype] or metric in ("euclidean_metric", "manhattan_distance", "hamming_distance"):
                assert np.array_equal(osc, sc), (metric, osc, sc)
            else:
                np.testing.assert_allclose(osc, sc, rtol=1e-5)
            # ties -> lower index
            for a in range(k - 1):
                if sc[a] == sc[a + 1]:
                    assert idx[a] < idx[a + 1]
            # a random sample of rows must not beat the k-th score
            samp = rng.choice(n, size=4096, replace=False)
            ssc = K.total_scores(rows[torch.as_tensor(samp, device=dev)].cpu().numpy(), q, metric)
            inside = np.isin(samp, idx)
            assert np.all(ssc[~inside] <= sc[-1])
            tied = (~inside) & (ssc == sc[-1])
            assert np.all(samp[tied] > idx[-1])
            # idempotence
            idx2, sc2, _, _ = m.query(q, k, metric)
            assert np.array_equal(idx2[0], idx) and np.array_equal(sc2[0], sc)
        # keep mask + time decay: dropped rows never appear, decay reference = max over KEPT rows
        metric = metrics[-1] if name != "C5" else "euclidean_metric"
        g = torch.Generator(device=dev)
        g.manual_seed(5)
        bits = torch.randint(-2**31, 2**31 - 1, ((n + 31) // 32,), generator=g, device=dev, dtype=torch.int32)
        ts = 1.7e9 + 3600.0 * torch.rand(n, generator=g, device=dev, dtype=torch.float64)
        m.set_mask(bits)
        m.set_timestamps(ts)
        ts_max, kept = m.kept_ts_max()
        keep_host = ((bits.cpu().numpy().view(np.uint32)[:, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool).reshape(-1)[:n]
        assert kept == int(keep_host.sum())
        assert ts_max == float(ts.cpu().numpy()[keep_host].max())
        m.set_decay_reference(ts_max)
        idx, sc, cnt, flags = m.query(q, k, metric, 0.3)
        idx, sc = idx[0], sc[0]
        assert np.all(keep_host[idx]) and np.all(np.diff(sc) <= 0)
        got_rows = rows[torch.as_tensor(idx, device=dev)].cpu().numpy()
        sims = K.scores(got_rows, q, metric).astype(np.float64)
        expect = sims + 0.3 * np.exp(-ts_max + ts.cpu().numpy()[idx])
        np.testing.assert_allclose(sc, expect, rtol=1e-5 if not EXACT[dtype] else 1e-14)
        # sharded == unsharded (two shards on this GPU; needs room for a second copy of the rows)
        m.set_mask(None)
        m.set_timestamps(None)
        if name in ("C2",):
            metric = metrics[0]
            ref_idx, ref_sc, _, _ = m.query(q, k, metric)
            parts, engines = [], []
            for r in range(2):
                lo, hi = shard_bounds(n, 2, r)
                engines.append(CudaEngine(hb.DeviceMatrix(rows[lo:hi], row_offset=lo)))
                parts.append(engines[-1].local_topk(torch.as_tensor(q[None, :]), k, metric, 0.0))
            midx, msc, mcnt, mflags = engines[0].merge(torch.stack(parts), 1, k)
            assert np.array_equal(midx.cpu().numpy()[0], ref_idx[0]) and np.array_equal(msc.cpu().numpy()[0], ref_sc[0])
            for e in engines:
                e.m.close()
    finally:
        m.close()
        del rows
        torch.cuda.empty_cache()


BATCH_CONFIGS = [  # name, rows, dim, dtype, metric, k, batch  (BASELINE.json's batched configurations + the multi-query sweep)
    ("C2_b1024", 1_000_000, 384, "float32", "cosine_similarity", 10, 1024),
    ("C3_b64", 10_000_000, 768, "float16", "cosine_similarity", 10, 64),
    ("C3_b4096", 10_000_000, 768, "float16", "cosine_similarity", 10, 4096),
    ("C3_dot_b8", 10_000_000, 768, "float16", "dot_product", 10, 8),
    ("C5_euclid_b1024", 5_000_000, 1024, "float32", "euclidean_metric", 10, 1024),
    ("C5_manhattan_b8", 5_000_000, 1024, "float32", "manhattan_distance", 10, 8),
    ("C5_hamming_b8", 5_000_000, 1024, "float32", "hamming_distance", 10, 8),
]


def plant_batch(rows, Q, per_query, rng, torch):
    """For every query b overwrite `per_query` rows with unit(q_b + 0.05*(i+1)*noise_i), i = 0..per_query-1, at random
    distinct positions (the positions of one query are NOT sorted by rank): cos(q_b, planted_i) = 1/sqrt(1 + (0.05(i+1))^2) is
    far above anything a random unit row reaches, so query b must return its plants first, in plant order."""
    n, d = rows.shape
    b = Q.shape[0]
    pos = rng.choice(n, size=b * per_query, replace=False).reshape(b, per_query)
    Qf = torch.as_tensor(Q.astype(np.float32), device=rows.device)
    for i in range(per_query):
        noise = torch.as_tensor(rng.standard_normal((b, d)).astype(np.float32), device=rows.device)
        noise = noise - ((noise * Qf).sum(1, keepdim=True) / (Qf * Qf).sum(1, keepdim=True)) * Qf
        noise = noise / noise.norm(dim=1, keepdim=True) * Qf.norm(dim=1, keepdim=True)
        v = Qf + 0.05 * (i + 1) * noise
        v = v / v.norm(dim=1, keepdim=True)
        rows[torch.as_tensor(pos[:, i], device=rows.device)] = v.to(rows.dtype)
    return pos


@pytest.mark.parametrize("name,n,d,dtype,metric,k,b", BATCH_CONFIGS, ids=[c[0] for c in BATCH_CONFIGS])
def test_fullsize_batches(name, n, d, dtype, metric, k, b):
    """BASELINE.json's batched configurations at full size (tensor-core path for the wide batches, multi-query sweep for the
    small ones): per-query planted rows come back first and in plant order, every returned score equals the oracle's score of
    that row (bit for bit where NumPy is deterministic), no row of a random sample beats a query's k-th score, ties go to the
    lower index, and the batched answer equals the single-query answer for a subset of the queries."""
    import torch
    import bench
    import hyperdb_b200 as hb
    free, _total = torch.cuda.mem_get_info()
    if free < n * d * bench.ITEM[dtype] * 1.35 + (4 << 30):
        pytest.skip("not enough free HBM for this configuration")
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(777)
    rows = bench.gen_rows_torch(0, n, d, dtype, dev, seed=0)
    Q = bench.gen_queries(b, d, dtype, seed=4242)
    per_query = 3
    pos = plant_batch(rows, Q, per_query, rng, torch)
    m = hb.DeviceMatrix(rows)
    try:
        idx, sc, cnt, flags = m.query(Q, k, metric)
        assert idx.shape == (b, k) and np.all(cnt == k)
        assert np.all(np.diff(sc, axis=1) <= 0)
        if metric != "hamming_distance":             # sign bits of the planted rows are not ordered by construction
            assert np.array_equal(idx[:, :per_query], pos), np.nonzero((idx[:, :per_query] != pos).any(axis=1))[0][:10]
        # ties -> lower index
        tie = sc[:, :-1] == sc[:, 1:]
        assert np.all(idx[:, :-1][tie] < idx[:, 1:][tie])
        # the oracle's score of every returned row (rows copied back in one gather)
        got = rows[torch.as_tensor(idx.reshape(-1), device=dev)].cpu().numpy().reshape(b, k, d)
        check = range(b) if b <= 64 else list(range(0, b, max(1, b // 192)))
        exact = EXACT[dtype] or metric in ("euclidean_metric", "manhattan_distance", "hamming_distance")
        for qi in check:
            osc = K.total_scores(got[qi], Q[qi], metric)
            if exact:
                assert np.array_equal(osc, sc[qi]), (qi, osc, sc[qi])
            else:
                np.testing.assert_allclose(osc, sc[qi], rtol=1e-5)
        # a random sample of rows must not beat any query's k-th score (oracle arithmetic on a subset of the queries)
        samp = rng.choice(n, size=2048, replace=False)
        srows = rows[torch.as_tensor(samp, device=dev)].cpu().numpy()
        for qi in list(check)[:24]:
            ssc = K.total_scores(srows, Q[qi], metric)
            inside = np.isin(samp, idx[qi])
            assert np.all(ssc[~inside] <= sc[qi, -1]), qi
            tied = (~inside) & (ssc == sc[qi, -1])
            assert np.all(samp[tied] > idx[qi, -1])
        # batched == one query at a time
        for qi in list(check)[:6]:
            i1, s1, _, _ = m.query(Q[qi], k, metric)
            assert np.array_equal(i1[0], idx[qi]) and np.array_equal(s1[0], sc[qi]), qi
    finally:
        m.close()
        del rows
        torch.cuda.empty_cache()
