"""Test-side stand-in for hyperdb_b200.DeviceMatrix, backed by the oracle (NumPy): lets the HOST logic of the HyperDB shim --
filters -> row subset, skip_doc, the two-stage recency, clustered storage bookkeeping, result mapping, caching -- run in the
CPU test suite against the fixtures generated from the REAL HyperDB class.  Not part of the product."""
import numpy as np

from oracle import canonical as K


class FakeDeviceMatrix:
    calls = []                     # (method, ...) log shared by all instances of one test

    def __init__(self, vectors, device=None, row_offset=0):
        self.V = np.array(vectors)
        self.shape = self.V.shape
        self.np_dtype = self.V.dtype
        self.device = 0 if device is None else device
        self.row_offset = int(row_offset)
        self.order = None
        self.lo, self.hi, self.mask = 0, self.shape[0], None
        self.ts = self.decay = None
        FakeDeviceMatrix.calls.append(("create", self.shape[0]))

    def close(self):
        pass

    # -- row subset / order ------------------------------------------------------------------
    def set_row_order(self, order):
        self.order = None if order is None else np.asarray(order, np.int64)
        if self.order is not None:
            assert sorted(self.order.tolist()) == list(range(self.shape[0]))
        FakeDeviceMatrix.calls.append(("set_row_order",))

    def set_range(self, lo, hi):
        self.lo, self.hi = max(0, int(lo)), min(self.shape[0], int(hi))
        FakeDeviceMatrix.calls.append(("set_range", self.lo, self.hi))

    def set_mask(self, keep):
        self.mask = None if keep is None else np.asarray(keep, bool).copy()
        FakeDeviceMatrix.calls.append(("set_mask", None if keep is None else int(self.mask.sum())))

    def _kept(self):
        k = np.zeros(self.shape[0], bool)
        k[self.lo:self.hi] = True
        return k if self.mask is None else (k & self.mask)

    @property
    def n_kept(self):
        return int(self._kept().sum())

    # -- time decay ---------------------------------------------------------------------------
    def stage_column(self, values):
        return np.array(values, np.float64)

    def set_timestamps(self, ts):
        self.ts = None if ts is None else np.array(ts, np.float64)          # a COPY: stage 1 transforms it in place
        self.decay = None
        FakeDeviceMatrix.calls.append(("set_timestamps", ts is not None))

    def kept_ts_max(self):
        k = self._kept()
        return (float(self.ts[k].max()) if k.any() else float("-inf")), int(k.sum())

    def stage1_recency(self, bias1, ts_max):
        self.ts = bias1 * np.exp(-ts_max + self.ts)                          # hyperdb/hyperdb.py:1344
        FakeDeviceMatrix.calls.append(("stage1",))

    def set_decay_reference(self, ts_max):
        self.decay = np.exp(-ts_max + self.ts)                               # ranking_algorithm.py:183
        FakeDeviceMatrix.calls.append(("decay",))

    # -- mutation -----------------------------------------------------------------------------
    def append(self, rows):
        assert self.order is None
        self.V = np.concatenate([self.V, np.asarray(rows, self.V.dtype)])
        self.shape = self.V.shape
        self.lo, self.hi, self.mask, self.ts, self.decay = 0, self.shape[0], None, None, None

    def remove_rows(self, local_rows):
        assert self.order is None
        keep = np.ones(self.shape[0], bool)
        keep[np.asarray(local_rows)] = False
        self.V = self.V[keep]
        self.shape = self.V.shape
        self.lo, self.hi, self.mask, self.ts, self.decay = 0, self.shape[0], None, None, None

    # -- the ranking itself: the oracle, with the caller's numbering and tie rule -----------------
    def query(self, queries, top_k, metric, recency_bias=0.0):
        Q = np.atleast_2d(np.asarray(queries))
        k = max(int(top_k), 0)
        kept = self._kept()
        order = np.arange(self.shape[0]) if self.order is None else self.order
        idx = np.full((len(Q), k), -1, np.int64)
        sc = np.full((len(Q), k), -np.inf)
        cnt = np.zeros(len(Q), np.int64)
        for b, q in enumerate(Q):
            s = K.scores(self.V, q, metric).astype(np.float64)
            s[np.isnan(s)] = -np.inf
            if self.decay is not None and recency_bias != 0:
                s = s + recency_bias * self.decay
            rows = np.flatnonzero(kept)
            rows = rows[np.lexsort((order[rows], -s[rows]))][:k]
            idx[b, :len(rows)] = order[rows] + self.row_offset
            sc[b, :len(rows)] = s[rows]
            cnt[b] = len(rows)
        FakeDeviceMatrix.calls.append(("query", len(Q), k, metric))
        return idx, sc, cnt, np.zeros(len(Q), np.uint32)


class FakeEngine:
    """hyperdb_b200.sharded.CudaEngine's interface over a FakeDeviceMatrix (CPU, gloo tests of the sharded shim)."""
    device = "cpu"
    post = None

    def __init__(self, matrix):
        self.m = matrix

    def kept_ts_max(self):
        return self.m.kept_ts_max()

    def set_decay_reference(self, ts_max):
        self.m.set_decay_reference(ts_max)

    def n_kept(self):
        return self.m.n_kept

    def n_rows(self):
        return self.m.shape[0]

    def row_offset(self):
        return self.m.row_offset

    def append(self, rows):
        self.m.append(rows)

    def remove_local(self, local_rows):
        self.m.remove_rows(local_rows)

    def set_row_offset(self, off):
        self.m.row_offset = int(off)

    def local_topk(self, queries, k, metric, bias, exact=False):
        import torch
        from hyperdb_b200.sharded import packed_len
        q = np.asarray(queries)
        q = q[None, :] if q.ndim == 1 else q
        b = len(q)
        idx, sc, cnt, _flags = self.m.query(q, k, metric, bias)
        buf = torch.zeros(packed_len(b, k), dtype=torch.int64)
        buf[: b * k].view(torch.float64).view(b, k)[:] = torch.from_numpy(sc)
        buf[b * k: 2 * b * k].view(b, k)[:] = torch.from_numpy(idx)
        buf[2 * b * k: 2 * b * k + b] = torch.from_numpy(cnt)
        return buf

    def merge(self, gathered, b, k):
        from sharding_fakes import OracleEngine
        return OracleEngine.merge(self, gathered, b, k)
