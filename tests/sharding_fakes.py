"""Test-side engine for hyperdb_b200.sharded.ShardedMatrix: the local shard is scored by the ORACLE on the
CPU so the host logic (bounds, packing, exchange, flag handling, decay all-reduce) can run under gloo
without a GPU.  Never imported by the product."""
import numpy as np
import torch

from hyperdb_b200.sharded import packed_len
from oracle import canonical as K


class OracleEngine:
    device = "cpu"

    def __init__(self, V, row_offset, ts=None, keep=None, fail_first=False):
        self.V, self.off, self.ts, self.keep = V, row_offset, ts, keep
        self.ts_ref = None
        self.fail_first = fail_first        # report "uncertified" once to exercise the repair branch
        self.calls = []

    def kept_ts_max(self):
        keep = np.ones(len(self.V), bool) if self.keep is None else self.keep
        if self.ts is None or not keep.any():
            return float("-inf"), int(keep.sum())
        return float(self.ts[keep].max()), int(keep.sum())

    def set_decay_reference(self, ts_max):
        self.ts_ref = ts_max

    def n_kept(self):
        return len(self.V) if self.keep is None else int(self.keep.sum())

    def n_rows(self):
        return len(self.V)

    def row_offset(self):
        return self.off

    def append(self, rows):
        self.V = np.concatenate([self.V, np.asarray(rows, self.V.dtype)])
        self.ts, self.keep = None, None

    def remove_local(self, local_rows):
        alive = np.ones(len(self.V), bool)
        alive[np.asarray(local_rows)] = False
        self.V = self.V[alive]
        self.ts = None if self.ts is None else self.ts[alive]
        self.keep = None

    def set_row_offset(self, off):
        self.off = off

    def local_topk(self, queries, k, metric, bias, exact=False):
        q = np.asarray(queries)
        q = q[None, :] if q.ndim == 1 else q
        b = len(q)
        self.calls.append(exact)
        buf = torch.zeros(packed_len(b, k), dtype=torch.int64)
        sc = buf[: b * k].view(torch.float64).view(b, k)
        ids = buf[b * k: 2 * b * k].view(b, k)
        cnt = buf[2 * b * k: 2 * b * k + b]
        flags = buf[2 * b * k + b:].view(torch.int32)
        for i in range(b):
            s = K.scores(self.V, q[i], metric).astype(np.float64)
            s[np.isnan(s)] = -np.inf
            if self.ts is not None and self.ts_ref is not None:
                s = s + bias * np.exp(-self.ts_ref + self.ts)
            keep = np.ones(len(s), bool) if self.keep is None else self.keep
            rows = np.flatnonzero(keep)
            order = rows[np.lexsort((rows, -s[rows]))][:k]
            n = len(order)
            sc[i, :n] = torch.from_numpy(s[order])
            ids[i, :n] = torch.from_numpy(order + self.off)
            cnt[i] = n
            flags[i] = 8 if (self.fail_first and not exact) else 0
        return buf

    def merge(self, gathered, b, k):
        g = gathered.shape[0]
        idx = torch.full((b, k), -1, dtype=torch.int64)
        sc = torch.full((b, k), float("-inf"), dtype=torch.float64)
        cnt = torch.zeros(b, dtype=torch.int64)
        for i in range(b):
            ss, ii = [], []
            for l in range(g):
                row = gathered[l]
                n = int(row[2 * b * k + i])
                ss.append(row[: b * k].view(torch.float64).view(b, k)[i, :n].numpy())
                ii.append(row[b * k: 2 * b * k].view(b, k)[i, :n].numpy())
            ss, ii = np.concatenate(ss), np.concatenate(ii)
            order = np.lexsort((ii, -ss))[:k]
            idx[i, :len(order)] = torch.from_numpy(ii[order])
            sc[i, :len(order)] = torch.from_numpy(ss[order])
            cnt[i] = len(order)
        flags = gathered[:, 2 * b * k + b:].view(torch.int32)[:, :b]
        return idx, sc, cnt, flags
