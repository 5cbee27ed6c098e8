"""Host logic of the row-sharded path under torch.distributed/gloo, world_size 2, no GPU:
shard bounds, message packing, the single all-gather, the decay all-reduce(MAX) over kept rows, the
uncertified-flag repair branch.  The local engine is the oracle (tests/sharding_fakes.py)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hyperdb_b200.sharded import ShardedMatrix, packed_len, shard_bounds
from oracle import canonical as K


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 1000, 10_000_000, 100_000_003):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_packed_len():
    assert packed_len(1, 10) == 22 and packed_len(3, 4) == 24 + 3 + 2


def _worker(rank, world, port, fail_first, result_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.join(here, "golden"), os.path.dirname(here), os.path.join(os.path.dirname(here), "local-hyperdb_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from sharding_fakes import OracleEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(42)
    n, d = 1001, 24
    V = rng.standard_normal((n, d)).astype(np.float32)
    V[500] = V[10]                      # an exact tie across the shard boundary
    Q = rng.standard_normal((3, d)).astype(np.float32)
    Q[0] = V[10]
    ts = 1.7e9 + rng.uniform(0, 9, n)
    keep = rng.random(n) < 0.7
    lo, hi = shard_bounds(n, world, rank)
    eng = OracleEngine(V[lo:hi], lo, ts[lo:hi], keep[lo:hi], fail_first=fail_first)
    sm = ShardedMatrix(eng, n)
    ref = sm.refresh_decay()
    assert ref == ts[keep].max()
    assert sm.total_kept() == int(keep.sum())
    for metric in ("cosine_similarity", "euclidean_metric", "hamming_distance"):
        idx, sc, cnt = sm.query(Q, 7, metric, 0.3)
        for b in range(len(Q)):
            oi, os_ = K.rank(V, Q[b], 7, metric, ts, 0.3, keep)
            assert list(idx[b]) == list(oi), (metric, b)
            assert np.array_equal(sc[b], os_)
            assert cnt[b] == 7
    assert sm.exchanges == (6 if fail_first else 3)
    assert eng.calls == ([False, True] * 3 if fail_first else [False] * 3)
    # mutation: remove rows on both sides of the shard boundary, then append; global ids stay dense and ordered
    drop = np.array([0, 10, 499, 500, 501, 1000, 10])
    alive = np.ones(n, bool)
    alive[drop] = False
    assert sm.remove_rows(drop) == n - 6
    extra = rng.standard_normal((5, d)).astype(np.float32)
    assert sm.append(extra) == n - 1
    V2 = np.concatenate([V[alive], extra])
    idx, sc, cnt = sm.query(Q[1:], 9, "euclidean_metric")
    for b in range(2):
        oi, os_ = K.rank(V2, Q[1 + b], 9, "euclidean_metric")
        assert list(idx[b]) == list(oi) and np.array_equal(sc[b], os_)
    with pytest.raises(IndexError):
        sm.remove_rows([n + 100])
    # enable_peer_exchange is collective-safe: if ONE rank cannot map its peers, EVERY rank raises (nobody is left waiting
    # in a collective) and the all-gather path stays in force; if all succeed, all switch.  (PeerExchange itself needs CUDA
    # IPC: a stand-in with the same interface is patched in.)
    import hyperdb_b200.sharded as S

    class FakeExchange:
        fail_on = None

        def __init__(self, device_index, world_, rank_, max_words):
            self.rank, self.max_words, self.closed = rank_, max_words, False

        def handle(self):
            return bytes([self.rank]) * 8

        def connect(self, handles):
            assert [h[0] for h in handles] == list(range(world))
            if FakeExchange.fail_on == self.rank:
                raise RuntimeError("cudaIpcOpenMemHandle failed")

        def close(self):
            self.closed = True

    real, S.PeerExchange = S.PeerExchange, FakeExchange
    eng.device = torch.device("cpu")
    try:
        FakeExchange.fail_on = 1
        with pytest.raises(RuntimeError):
            sm.enable_peer_exchange(max_batch=4, max_k=8)
        assert sm.xchg is None
        idx, sc, cnt = sm.query(Q[1:], 9, "euclidean_metric")         # still answers through the all-gather
        assert list(idx[0]) == list(K.rank(V2, Q[1], 9, "euclidean_metric")[0])
        FakeExchange.fail_on = None
        assert sm.enable_peer_exchange(max_batch=4, max_k=8) is True and isinstance(sm.xchg, FakeExchange)
        sm.xchg = None
    finally:
        S.PeerExchange = real
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")


@pytest.mark.parametrize("fail_first", [False, True])
def test_world2_gloo(tmp_path, fail_first):
    port = 29600 + (os.getpid() % 200) + (50 if fail_first else 0)
    mp.spawn(_worker, args=(2, port, fail_first, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_query_cuts_batches_that_exceed_the_peer_exchange_buffer():
    """ShardedMatrix.query with a peer exchange attached: hdb_query refuses a step whose message exceeds the exchange buffer
    (csrc/api.cu), so a larger batch is cut into the largest pieces that fit, at most three in flight, and the answers are
    concatenated in order.  Host logic only: submit / collect are stand-ins that record what they were asked."""
    from hyperdb_b200.sharded import ShardedMatrix, packed_len

    class Xchg:
        max_words = packed_len(64, 128)

    class Engine:
        m = object()

    sm = ShardedMatrix.__new__(ShardedMatrix)
    sm.engine, sm.xchg, sm.world, sm.rank = Engine(), Xchg(), 2, 0
    log, in_flight = [], []

    def submit(q, k, metric, bias=0.0, _path=0):
        q = np.atleast_2d(np.asarray(q))
        in_flight.append(len(q))
        assert len(in_flight) <= 3
        log.append(len(q))
        return (q, k)

    def collect(t):
        in_flight.pop(0)
        q, k = t
        b = len(q)
        return (np.repeat(q[:, :1].astype(np.int64), k, axis=1), np.repeat(q[:, 1:2].astype(np.float64), k, axis=1), np.full(b, k, np.int64))

    sm.submit, sm.collect = submit, collect
    for k in (10, 128, 1000):
        fit = sm.exchange_batch_limit(k)
        assert packed_len(fit, k) <= Xchg.max_words < packed_len(fit + 1, k)
        b = 2 * fit + 5
        Q = np.stack([np.arange(b), np.arange(b) * 0.5, np.zeros(b)], axis=1)
        log.clear()
        idx, sc, cnt = sm.query(Q, k, "cosine_similarity")
        assert log == [fit, fit, 5] and not in_flight
        assert idx.shape == (b, k) and np.array_equal(idx[:, 0], np.arange(b)) and np.array_equal(sc[:, 0], np.arange(b) * 0.5)
        assert cnt.tolist() == [k] * b
        log.clear()
        sm.query(Q[:fit], k, "cosine_similarity")                          # fits: one step, as before
        sm.query(Q[0], k, "cosine_similarity")                             # a single 1-D query
        assert log == [fit, 1]
    with pytest.raises(ValueError, match="does not fit"):
        sm.query(Q, 10 ** 6, "cosine_similarity")
    sm.xchg = None
    assert sm.exchange_batch_limit(10) is None
