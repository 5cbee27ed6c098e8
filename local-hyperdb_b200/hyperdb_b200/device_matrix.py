"""DeviceMatrix: the stored embedding matrix (`HyperDB.vectors`, hyperdb/hyperdb.py:127-135) living
row-major in the HBM of one B200, with the per-row state the ranking kernels need (norms as
get_norm_vector computes them, packed sign bits, time-decay column, row mask/range).

One DeviceMatrix is one row shard; `hyperdb_b200.sharded.ShardedMatrix` holds one per rank.
Everything here is a thin wrapper over the C ABI (include/hyperdb_b200.h).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

_NP2HDB = {np.dtype(np.float16): N.HDB_F16, np.dtype(np.float32): N.HDB_F32, np.dtype(np.float64): N.HDB_F64}
_HDB2NP = {N.HDB_F16: np.float16, N.HDB_F32: np.float32, N.HDB_F64: np.float64, 3: np.uint64}


def as_float_array(x):
    """NumPy view of `x` in one of the three stored dtypes.  Integer/bool input becomes float64
    (what NumPy's own promotion gives the reference once a float enters the expression)."""
    a = np.asarray(x)
    if a.dtype not in _NP2HDB:
        if a.dtype.kind in "iub":
            a = a.astype(np.float64)
        elif a.dtype.kind == "f":
            a = a.astype(np.float64)
        else:
            raise TypeError(f"unsupported dtype {a.dtype}")
    return np.ascontiguousarray(a)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(x):
    """(address, space) of a NumPy array or a torch tensor."""
    if _is_torch(x):
        return C.c_void_p(x.data_ptr()), (N.HDB_DEVICE if x.is_cuda else N.HDB_HOST)
    return C.c_void_p(x.ctypes.data), N.HDB_HOST


class DeviceMatrix:
    """vectors: (n, d) NumPy array (uploaded) or CUDA torch tensor (adopted without a copy;
    the tensor is kept alive).  dtype float16/32/64 is kept as given (the reference stores
    `fp_precision`, hyperdb/hyperdb.py:65-66)."""

    def __init__(self, vectors, device=None, row_offset=0):
        self._h = C.c_void_p()
        self._keep = None
        if _is_torch(vectors):
            import torch
            t = vectors.contiguous()
            if not t.is_cuda:
                t = t.cuda(device if device is not None else 0)
            dt = {torch.float16: N.HDB_F16, torch.float32: N.HDB_F32, torch.float64: N.HDB_F64}[t.dtype]
            if t.dim() != 2:
                raise ValueError("vectors must be 2-D")
            dev = t.device.index
            N.check(N.lib().hdb_matrix_create(dev, dt, t.shape[0], t.shape[1], row_offset, C.byref(self._h)))
            N.check(N.lib().hdb_matrix_adopt(self._h, C.c_void_p(t.data_ptr())))
            self._keep = t
            self.np_dtype = np.dtype(_HDB2NP[dt])
            self.shape = (int(t.shape[0]), int(t.shape[1]))
            self.device = dev
        else:
            a = as_float_array(vectors)
            if a.ndim != 2:
                raise np.exceptions.AxisError(1, a.ndim)          # as np.linalg.norm(axis=1) on 1-D input
            dev = 0 if device is None else int(device)
            N.check(N.lib().hdb_matrix_create(dev, _NP2HDB[a.dtype], a.shape[0], a.shape[1], row_offset, C.byref(self._h)))
            N.check(N.lib().hdb_matrix_upload(self._h, 0, a.shape[0], C.c_void_p(a.ctypes.data), N.HDB_HOST))
            self.np_dtype = a.dtype
            self.shape = a.shape
            self.device = dev
        self.row_offset = int(row_offset)
        self._has_ts = False
        N.check(N.lib().hdb_matrix_finalize(self._h))       # raises ValueError on NaN (ranking_algorithm.py:150-151)

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            N.lib().hdb_matrix_destroy(self._h)
            self._h = C.c_void_p()
            self._keep = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return self.shape[0]

    # -- mutation (hyperdb/hyperdb.py:504-509 add, :718-728 remove_document) ---------------------------
    def reserve(self, capacity_rows):
        N.check(N.lib().hdb_matrix_reserve(self._h, int(capacity_rows)))

    def append(self, rows):
        """Append rows (NumPy array or CUDA tensor of the stored dtype; other float dtypes are cast as
        `np.concatenate` would not: the stored dtype wins).  Resets mask / range / timestamps."""
        if self._keep is not None:
            raise ValueError("cannot append to a DeviceMatrix that adopted a CUDA tensor")
        if _is_torch(rows):
            import torch
            t = rows.contiguous().to({np.dtype(np.float16): torch.float16, np.dtype(np.float32): torch.float32,
                                      np.dtype(np.float64): torch.float64}[self.np_dtype])
            if t.dim() != 2 or t.shape[1] != self.shape[1]:
                raise ValueError("appended rows must be (m, d)")
            p, space = _ptr(t)
            cnt = int(t.shape[0])
        else:
            a = np.ascontiguousarray(np.asarray(rows), dtype=self.np_dtype)
            if a.ndim != 2 or a.shape[1] != self.shape[1]:
                raise ValueError("appended rows must be (m, d)")
            p, space, cnt = C.c_void_p(a.ctypes.data), N.HDB_HOST, a.shape[0]
        N.check(N.lib().hdb_matrix_append(self._h, cnt, p, space))
        self.shape = (self.shape[0] + cnt, self.shape[1])
        self._has_ts = False

    def remove_rows(self, local_rows):
        """Remove the listed local rows; survivors keep their order (stable device-side compaction)."""
        if self._keep is not None:
            raise ValueError("cannot remove rows of a DeviceMatrix that adopted a CUDA tensor")
        idx = np.ascontiguousarray(np.atleast_1d(np.asarray(local_rows)), dtype=np.int64)
        idx = np.where(idx < 0, idx + self.shape[0], idx)
        N.check(N.lib().hdb_matrix_remove_rows(self._h, C.c_void_p(idx.ctypes.data), len(idx), N.HDB_HOST))
        out = C.c_int64()
        N.check(N.lib().hdb_matrix_info(self._h, None, C.byref(out), None, None, None))
        self.shape = (int(out.value), self.shape[1])

    def set_row_offset(self, row_offset):
        N.check(N.lib().hdb_matrix_set_row_offset(self._h, int(row_offset)))
        self.row_offset = int(row_offset)

    # -- row subset (the filters' output) ------------------------------------------------------
    def set_mask(self, keep):
        """keep: bool[n] (True = row takes part), packed uint32 bits, CUDA tensor of packed bits, or None."""
        if keep is None:
            N.check(N.lib().hdb_matrix_set_mask(self._h, None, N.HDB_HOST))
            return
        if _is_torch(keep):
            p, space = _ptr(keep)
            N.check(N.lib().hdb_matrix_set_mask(self._h, p, space))
            return
        keep = np.asarray(keep)
        if keep.dtype == np.bool_:
            if keep.shape != (self.shape[0],):
                raise ValueError("mask must have one entry per row")
            bits = np.packbits(keep, bitorder="little")
            bits = np.concatenate([bits, np.zeros((-len(bits)) % 4, np.uint8)]).view(np.uint32)
        else:
            bits = np.ascontiguousarray(keep, np.uint32)
        N.check(N.lib().hdb_matrix_set_mask(self._h, C.c_void_p(bits.ctypes.data), N.HDB_HOST))

    def set_row_order(self, order):
        """order[p] = the caller's local index of physical row p (uint32 permutation; NumPy array or CUDA tensor), or None.
        Ids and ties then follow the caller's numbering although the rows are stored clustered (hdb_matrix_set_row_order)."""
        if order is None:
            N.check(N.lib().hdb_matrix_set_row_order(self._h, None, N.HDB_HOST))
        elif _is_torch(order):
            import torch
            t = order.contiguous()
            if t.dtype not in (torch.int32, torch.uint32) or t.numel() != self.shape[0]:
                raise ValueError("row order must be a 32-bit permutation with one entry per row")
            p, space = _ptr(t)
            N.check(N.lib().hdb_matrix_set_row_order(self._h, p, space))
        else:
            o = np.ascontiguousarray(order, np.uint32)
            if o.shape != (self.shape[0],):
                raise ValueError("row order must have one entry per row")
            N.check(N.lib().hdb_matrix_set_row_order(self._h, C.c_void_p(o.ctypes.data), N.HDB_HOST))

    def set_range(self, lo, hi):
        N.check(N.lib().hdb_matrix_set_range(self._h, int(lo), int(hi)))

    @property
    def n_kept(self):
        out = C.c_int64()
        N.check(N.lib().hdb_matrix_info(self._h, None, None, None, None, C.byref(out)))
        return out.value

    # -- time decay --------------------------------------------------------------------------
    def set_timestamps(self, ts):
        if ts is None:
            N.check(N.lib().hdb_matrix_set_timestamps(self._h, None, N.HDB_HOST))
            self._has_ts = False
            return
        if _is_torch(ts):
            p, space = _ptr(ts.contiguous())
        else:
            ts = np.ascontiguousarray(ts, np.float64)
            if ts.shape != (self.shape[0],):
                raise ValueError("timestamps must have one entry per row")
            p, space = C.c_void_p(ts.ctypes.data), N.HDB_HOST
        N.check(N.lib().hdb_matrix_set_timestamps(self._h, p, space))
        self._has_ts = True

    def stage_column(self, values):
        """A float64 per-row column (timestamps) as a device-resident tensor on this shard's GPU: the HyperDB shim keeps one per
        timestamp key and hands it to `set_timestamps` (device -> device) instead of re-uploading N doubles per recency query."""
        import torch
        return torch.as_tensor(np.ascontiguousarray(values, np.float64)).to(f"cuda:{self.device}")

    def kept_ts_max(self):
        """(max timestamp over kept rows of this shard, number of kept rows)."""
        mx, cnt = C.c_double(), C.c_int64()
        N.check(N.lib().hdb_matrix_kept_ts_max(self._h, C.byref(mx), C.byref(cnt)))
        return mx.value, cnt.value

    def set_decay_reference(self, ts_max):
        N.check(N.lib().hdb_matrix_set_decay_reference(self._h, float(ts_max)))

    def stage1_recency(self, bias1, ts_max):
        N.check(N.lib().hdb_matrix_stage1_recency(self._h, float(bias1), float(ts_max)))

    def refresh_decay(self):
        """Single-shard convenience: decay column from this shard's own kept maximum."""
        mx, cnt = self.kept_ts_max()
        if cnt > 0:
            self.set_decay_reference(mx)

    # -- queries --------------------------------------------------------------------------------
    def set_path(self, mode):
        """0 automatic, 1 exact full-vector path, 2 fused sweep only (error instead of fallback)."""
        N.check(N.lib().hdb_matrix_set_path(self._h, int(mode)))

    def set_max_group(self, n):
        """Cap on the queries that share one sweep pass (1 = one pass per query; 0 = as many as the shape allows)."""
        N.check(N.lib().hdb_matrix_set_max_group(self._h, int(n)))

    def set_post_stream(self, cuda_stream_ptr):
        """Pipelining of device-output queries: certify on this stream, overlapping the next query's sweep (0 = off)."""
        N.check(N.lib().hdb_matrix_set_post_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_sweep_overlap(self, on):
        """Pipelined mode: sweeps of consecutive queries may overlap (default) or run strictly one after the other."""
        N.check(N.lib().hdb_matrix_set_sweep_overlap(self._h, 1 if on else 0))

    def set_stream(self, cuda_stream_ptr):
        N.check(N.lib().hdb_matrix_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def query(self, queries, top_k, metric, recency_bias=0.0):
        """Host-resident results: (indices int64 [B,k], scores float64 [B,k], counts int64 [B], flags uint32 [B]).
        `queries`: (d,) or (B,d) array-like (NumPy float16/32/64; other dtypes -> float64) or CUDA tensor."""
        mid = N.METRIC_IDS.get(metric)
        if mid is None:
            raise ValueError(f"Unknown metric: {metric}")
        if _is_torch(queries):
            import torch
            q = queries.contiguous()
            qdt = {torch.float16: N.HDB_F16, torch.float32: N.HDB_F32, torch.float64: N.HDB_F64}[q.dtype]
            qshape = tuple(q.shape)
            p, space = _ptr(q)
        else:
            q = as_float_array(queries)
            qdt, qshape = _NP2HDB[q.dtype], q.shape
            p, space = C.c_void_p(q.ctypes.data), N.HDB_HOST
        if len(qshape) == 1:
            qshape = (1,) + qshape
        if len(qshape) != 2 or qshape[1] != self.shape[1]:
            raise ValueError(f"query dimension {qshape} does not match the stored dimension {self.shape[1]}")
        b, k = qshape[0], max(int(top_k), 0)
        idx = np.full((b, k), -1, np.int64)
        sc = np.full((b, k), -np.inf, np.float64)
        cnt = np.zeros(b, np.int64)
        flags = np.zeros(b, np.uint32)
        N.check(N.lib().hdb_query(self._h, mid, p, qdt, space, b, int(top_k), float(recency_bias),
                                  C.c_void_p(idx.ctypes.data), C.c_void_p(sc.ctypes.data), C.c_void_p(cnt.ctypes.data),
                                  C.c_void_p(flags.ctypes.data), N.HDB_HOST))
        return idx, sc, cnt, flags

    def query_device(self, queries, top_k, metric, recency_bias, out_idx, out_score, out_count, out_flags):
        """Asynchronous form: CUDA-tensor queries and outputs, nothing is synchronised.  A set
        FLAG_UNCERTIFIED bit in out_flags means that query must be repeated through `query`."""
        import torch
        mid = N.METRIC_IDS[metric]
        q = queries
        qdt = {torch.float16: N.HDB_F16, torch.float32: N.HDB_F32, torch.float64: N.HDB_F64}[q.dtype]
        b = 1 if q.dim() == 1 else q.shape[0]
        N.check(N.lib().hdb_query(self._h, mid, C.c_void_p(q.data_ptr()), qdt, N.HDB_DEVICE, b, int(top_k), float(recency_bias),
                                  C.c_void_p(out_idx.data_ptr()), C.c_void_p(out_score.data_ptr()),
                                  C.c_void_p(out_count.data_ptr()), C.c_void_p(out_flags.data_ptr()), N.HDB_DEVICE))

    def scores(self, query, metric, distance=False):
        """The metric function's own output for every stored row (NumPy result dtype; uint64 for hamming).
        distance=True (euclidean only): the distance itself, euclidean_metric(..., get_similarity_score=False)."""
        mid = N.METRIC_IDS.get(metric)
        if mid is None:
            raise ValueError(f"Unknown metric: {metric}")
        q = as_float_array(query).reshape(-1)
        if q.shape[0] != self.shape[1]:
            raise ValueError(f"operands could not be broadcast together with shapes {self.shape} {q.shape}")
        rdt = np.promote_types(self.np_dtype, q.dtype)
        out = np.empty(self.shape[0], np.uint64 if metric == "hamming_distance" else (np.float64 if metric in ("jaccard_similarity", "pearson_correlation") else rdt))
        got = C.c_int()
        N.check(N.lib().hdb_scores_ex(self._h, mid, C.c_void_p(q.ctypes.data), _NP2HDB[q.dtype], N.HDB_HOST,
                                      C.c_void_p(out.ctypes.data), N.HDB_HOST, C.byref(got), N.SCORES_DISTANCE if distance else 0))
        assert np.dtype(_HDB2NP[got.value]) == out.dtype
        return out

    def time_last_query(self, what=0, iters=20):
        ms = C.c_float()
        N.check(N.lib().hdb_time_last_query(self._h, int(what), int(iters), C.byref(ms)))
        return ms.value

    def profile_enable(self, max_pairs):
        N.check(N.lib().hdb_profile_enable(self._h, int(max_pairs)))

    def profile_read(self):
        """(launches of the dominant kernel recorded since the last read, their summed duration in ms)."""
        n, ms = C.c_int(), C.c_float()
        N.check(N.lib().hdb_profile_read(self._h, C.byref(n), C.byref(ms)))
        return n.value, ms.value
