"""Host shim that keeps `HyperDB.query`'s API (hyperdb/hyperdb.py:1584) over the device-resident matrix.

Only the brute-force branch of `_execute_query` (hyperdb/hyperdb.py:1429-1582) exists here: there is no Annoy
index, every query is an exact sweep on the GPU.  What the reference does with Python lists of rows becomes
state of the DeviceMatrix:

  skip_doc filter   (:1119-1134, :1474-1481)  -> kept row range            (hdb_matrix_set_range)
  metadata filter   (:1218-1257)              -> 1 bit per row             (hdb_matrix_set_mask)
  _handle_timestamps (:1310-1346)             -> stage-1 recency ON the device, then the sort's own
                                                 transform (recency is applied twice)
  row -> document   (:1565-1573)              -> global row id == document index (one row per document)

Per-query host work is O(top_k): metadata columns and timestamp columns are extracted from the documents ONCE per
(key, mutation) and cached; the timestamp column lives on the device; the row subset and the two-stage decay are only
recomputed when (filters, timestamp_key, recency_bias) differ from the previous query's.

Row-sharded over several GPUs: when torch.distributed is initialised with more than one rank (or
`group=` is given) every rank builds the HyperDB with the SAME documents and vectors and keeps rows
[rank*N/G, (rank+1)*N/G) on its GPU (hyperdb_b200.sharded.ShardedMatrix); `query` is then a collective -- every rank
calls it with the same arguments and gets the same answer.

Quirk kept on purpose: a vector query reaches the ranking as FLOAT64 whatever dtype the caller used (the LRU
key is `tuple(query.tolist())`, :1369-1370), so NumPy promotes the stored matrix and the scores are float64
arithmetic on the stored values.  Deviations from the reference, on purpose: the query-cache key is a 128-bit digest
of the query's float64 values (hdb_query_digest; computed on the device for CUDA-tensor queries) instead of the tuple of
Python floats; CUDA tensors are accepted as query_input; `query_batch` ranks a whole (B, d) tile in one call.
Out of scope: embedding model, chunking, 'key' and 'sentence' filters, persistence.
"""
from __future__ import annotations

import ctypes as C
import sys
import time
from collections import OrderedDict

import numpy as np

from . import _native as N
from .device_matrix import DeviceMatrix, _is_torch

_METRICS = ['dot_product', 'cosine_similarity', 'euclidean_metric', 'manhattan_distance', 'jaccard_similarity',
            'pearson_correlation', 'hamming_distance']

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_C0, _C1, _C2 = np.uint64(0x9e3779b97f4a7c15), np.uint64(0xd1b54a32d192ed03), np.uint64(0xa0761d6478bd642f)


def _mix(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xbf58476d1ce4e5b9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94d049bb133111eb)
    return z ^ (z >> np.uint64(31))


def query_digest_host(q):
    """The NumPy statement of hdb_query_digest (csrc/ingest.cu::query_digest_kernel): (h0, h1) of a 1-D query's float64
    values, -0.0 folded onto +0.0."""
    v = np.ascontiguousarray(np.asarray(q, dtype=np.float64).reshape(-1)) + 0.0          # -0.0 + 0.0 == +0.0
    bits = v.view(np.uint64)
    pos = np.arange(1, len(v) + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h0 = _mix(bits + pos * _C0).sum(dtype=np.uint64)
        h1 = _mix((bits ^ _C2) + pos * _C1).sum(dtype=np.uint64)
    return int(h0), int(h1)


class _LRUCache(OrderedDict):
    """The slice of cachetools.LRUCache the reference uses (hyperdb/hyperdb.py:60, :1381-1388, :1398-1404): `maxsize`,
    `in`, `[]`, `[]=` with least-recently-used eviction, `clear()`, `len()`.  A cachetools.LRUCache assigned to
    `HyperDB.lru_cache` (as the reference's tests do, tests/test_hyperdb.py:724-736) works as well."""

    def __init__(self, maxsize):
        super().__init__()
        self.maxsize = int(maxsize)

    def __getitem__(self, key):
        value = super().__getitem__(key)
        self.move_to_end(key)
        return value

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        self.move_to_end(key)
        while len(self) > self.maxsize:
            self.popitem(last=False)


def _nested(document, dotted):
    cur = document
    for part in dotted.split('.'):
        if not isinstance(cur, dict) or part not in cur:
            return None
        cur = cur[part]
    return cur


def _canon(x):
    """Hashable stand-in of a document under ==: equal documents map to equal keys (dict order does not matter)."""
    if isinstance(x, dict):
        return ("d", tuple(sorted(((k, _canon(v)) for k, v in x.items()), key=lambda kv: repr(kv[0]))))
    if isinstance(x, (list, tuple)):
        return ("l", tuple(_canon(v) for v in x))
    if isinstance(x, np.ndarray):
        return ("a", x.shape, x.tobytes())
    hash(x)
    return x


class HyperDB:
    def __init__(self, documents=None, vectors=None, select_keys=None, embedding_function=None, fp_precision="float32",
                 add_timestamp=False, metadata_keys=None, ann_metric="cosine", n_trees=10, cache_size=256, device=None,
                 group=None, sharded=None, cluster_by=None):
        if fp_precision not in ["float16", "float32", "float64"]:
            raise ValueError("Unsupported floating-point precision.")
        self.fp_precision = getattr(np, fp_precision)
        self.embedding_function = embedding_function
        self.add_timestamp = add_timestamp
        self.metadata_keys = [metadata_keys] if isinstance(metadata_keys, str) else list(metadata_keys or [])
        if add_timestamp and "timestamp" not in self.metadata_keys:
            self.metadata_keys.append("timestamp")
        self.ann_metric = ann_metric            # accepted for signature compatibility; there is no ANN index
        self.device = device
        self.documents = []
        self.source_indices = []
        self._chunks = []                        # host copy of the vectors as appended (concatenated only if `.vectors` is read)
        self._n, self._d, self._vdtype = 0, 0, None
        self._matrix = None                      # DeviceMatrix of the local shard
        self._sm = None                          # ShardedMatrix (world > 1)
        self._lo = 0                             # global id of the local shard's first row
        self._group = group
        # cluster_by: a metadata key whose equal values are STORED next to each other on the device (stable within a value),
        # so that a metadata filter on it keeps a few contiguous runs of rows and the masked sweep streams them at the full
        # HBM rate; ids, ties and every result are those of the unclustered store (hdb_matrix_set_row_order)
        self.cluster_by = cluster_by
        self._perm = None                        # physical local row -> local document index (None = identity)
        self._world, self._rank_id = 1, 0
        if sharded is not False:
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized() and (sharded or group is not None or dist.get_world_size() > 1):
                    self._world, self._rank_id = dist.get_world_size(group), dist.get_rank(group)
            except ImportError:
                pass
        self._invalidate_columns()
        # query cache (hyperdb/hyperdb.py:60-62, :1368-1388): same semantics, but the key is a 128-bit digest of the
        # query's values (hdb_query_digest) instead of tuple(query.tolist()) -- O(D) at C / GPU speed instead of D Python floats
        self.lru_cache = _LRUCache(cache_size)
        self.cache_hits = 0
        self.cache_misses = 0
        if vectors is not None:
            if documents is None or len(documents) != len(vectors):
                raise ValueError("documents and vectors must have the same length")
            self.documents = list(documents)
            v = np.asarray(vectors)              # stored as given (hyperdb/hyperdb.py:127-135)
            self.source_indices = list(range(len(self.documents)))
            self._chunks = [v]
            self._upload()
        elif documents:
            self.add(documents)

    # -- storage ---------------------------------------------------------------------------------
    @property
    def vectors(self):
        """The stored matrix as the reference keeps it (`HyperDB.vectors`): the host copy, concatenated on demand."""
        if not self._chunks:
            return None
        if len(self._chunks) > 1:
            self._chunks = [np.concatenate(self._chunks)]
        return self._chunks[0]

    @vectors.setter
    def vectors(self, value):
        self._chunks = [] if value is None else [np.asarray(value)]
        self._n = 0 if value is None else len(self._chunks[0])
        if self._n and self._chunks[0].ndim == 2:
            self._d, self._vdtype = int(self._chunks[0].shape[1]), self._chunks[0].dtype

    def _invalidate_columns(self):
        """Everything derived from the documents / the row set: dropped by every mutation."""
        self._mask_cache = {}                    # metadata spec -> bool[n]
        self._ts_cache = {}                      # timestamp_key -> (float64[n], missing bool[n])
        self._first_equal = None                 # canonical document -> first index (documents.index semantics)
        self._subset_tag = None                  # (lo, hi, mask tag) resident on the device
        self._decay_tag = None                   # (subset tag, timestamp_key, recency_bias) resident on the device
        self._ts_dev = {}                        # timestamp_key -> CUDA float64 tensor of the local rows

    def _close_device(self):
        if self._sm is not None and self._sm.xchg is not None:
            try:
                self._sm.engine.attach_exchange(None)
                self._sm.xchg.close()
            except Exception:
                pass
            self._sm.xchg = None
        self._sm = None
        if self._matrix is not None:
            self._matrix.close()
            self._matrix = None

    def _upload(self):
        self.clear_cache()
        self._close_device()
        self._invalidate_columns()
        v = self.vectors
        if v is None or len(v) == 0:
            self._n = 0
            return
        self._n, self._d, self._vdtype = int(v.shape[0]), int(v.shape[1]) if v.ndim == 2 else 0, v.dtype
        lo, hi = 0, self._n
        if self._world > 1:
            from .sharded import shard_bounds
            lo, hi = shard_bounds(self._n, self._world, self._rank_id)
        self._lo = lo
        local = v[lo:hi]
        self._perm = None
        if self.cluster_by is not None and hi - lo > 1:
            seen = {}
            code = np.fromiter((seen.setdefault(_canon(_nested(d, self.cluster_by)), len(seen)) for d in self.documents[lo:hi]),
                               np.int64, hi - lo)
            perm = np.argsort(code, kind="stable")
            if not np.array_equal(perm, np.arange(hi - lo)):
                self._perm = perm.astype(np.uint32)
                local = np.ascontiguousarray(local[perm])
        if self._world > 1:
            from .sharded import CudaEngine, ShardedMatrix
            import torch
            dev = self.device if self.device is not None else torch.cuda.current_device()
            self._matrix = DeviceMatrix(local, device=dev, row_offset=lo)
            self._sm = ShardedMatrix(CudaEngine(self._matrix), self._n, group=self._group)
            try:
                self._sm.enable_peer_exchange(max_batch=64, max_k=128)
            except Exception as e:               # noqa: BLE001 -- every rank fails alike; the all-gather path remains
                print(f"INFO: peer-memory exchange unavailable ({e}); candidates travel through the all-gather.")
        else:
            self._matrix = DeviceMatrix(local, device=self.device)
        if self._perm is not None:
            self._matrix.set_row_order(self._perm)

    def add(self, documents, vectors=None, add_timestamp=False):
        """hyperdb/hyperdb.py:496-545, :626-689 reduced to: embed (if needed), cast to fp_precision, append on the device
        (hdb_matrix_append: only the new rows are uploaded and ingested; sharded: they join the last rank's shard).  The
        host copy grows by one chunk -- no np.concatenate of the whole matrix per call (hyperdb/hyperdb.py:504-509)."""
        documents = [documents] if isinstance(documents, (dict, str)) else list(documents)
        if add_timestamp or self.add_timestamp:
            now = time.time()
            documents = [dict(d, timestamp=d.get("timestamp", now)) if isinstance(d, dict) else d for d in documents]
        if vectors is None:
            if self.embedding_function is None:
                raise ValueError("hyperdb_b200.HyperDB needs `vectors` or an `embedding_function` (no embedding model is bundled)")
            out = self.embedding_function(documents)
            vectors = out[0] if isinstance(out, tuple) else out
        vectors = np.asarray(vectors, dtype=self.fp_precision)
        if vectors.ndim != 2 or len(vectors) != len(documents):
            raise ValueError("one vector per document expected")
        base = len(self.documents)
        if self._matrix is None or self._n == 0 or self.cluster_by is not None:
            # first rows, or a clustered store (the layout is an ingest-time property: the shard is rebuilt and re-clustered)
            previous = (self._chunks, len(self.documents))
            self._chunks = list(self._chunks) + [vectors if self._vdtype is None else vectors.astype(self._vdtype)]
            self.documents.extend(documents)
            try:
                self._upload()                               # raises ValueError on NaN
            except Exception:
                self._chunks = previous[0]
                del self.documents[previous[1]:]
                raise
        else:
            vectors = vectors.astype(self._vdtype)
            if self._sm is not None:
                self._sm.append(vectors)                     # collective: the rows join the last rank's shard
                self._lo = self._matrix.row_offset
            else:
                self._matrix.append(vectors)                 # device side: only the new rows are copied and ingested
            self._chunks.append(vectors)
            self._n += len(vectors)
            self._invalidate_columns()
            self.clear_cache()                               # hyperdb/hyperdb.py:566
            self.documents.extend(documents)
        self.source_indices.extend(range(base, base + len(documents)))

    def remove_document(self, index):
        """hyperdb/hyperdb.py:691-766 for one-row-per-document stores."""
        keep = np.ones(len(self.documents), bool)
        keep[index] = False
        gone = np.flatnonzero(~keep)
        self.documents = [d for d, k in zip(self.documents, keep) if k]
        self._chunks = [self.vectors[keep]]
        self.source_indices = list(range(len(self.documents)))
        self._n = len(self.documents)
        if self._matrix is not None and self._n and self.cluster_by is None:
            if self._sm is not None:
                self._sm.remove_rows(gone)                   # collective; shards are renumbered
                self._lo = self._matrix.row_offset
            else:
                self._matrix.remove_rows(gone)               # stable compaction on the device, no re-upload
            self._invalidate_columns()
            self.clear_cache()                               # hyperdb/hyperdb.py:766
        else:
            self._upload()

    def size(self):
        return len(self.documents)

    # -- query cache ---------------------------------------------------------------------------------
    def clear_cache(self):
        """hyperdb/hyperdb.py:1390-1396."""
        self.lru_cache.clear()
        self.cache_hits = 0
        self.cache_misses = 0

    def get_cache_size_and_info(self):
        """hyperdb/hyperdb.py:1398-1427 (sizes from sys.getsizeof instead of pympler)."""
        nbytes = sys.getsizeof(self.lru_cache) + sum(sys.getsizeof(k) + sys.getsizeof(v) for k, v in self.lru_cache.items())
        if nbytes >= 1024 * 1024:
            size = f"{nbytes / (1024 * 1024):.2f} MB"
        elif nbytes >= 1024:
            size = f"{nbytes / 1024:.2f} KB"
        else:
            size = f"{int(nbytes)} bytes"
        return {"cache_info": {"hits": self.cache_hits, "misses": self.cache_misses, "maxsize": self.lru_cache.maxsize,
                               "currsize": len(self.lru_cache)}, "cache_memory_size": size}

    def _digest(self, query_input):
        """A CUDA-tensor query is hashed where it lives (hdb_query_digest: one tiny kernel, 16 bytes back); everything else
        is left to `_hashable_key` (same function on the host)."""
        if _is_torch(query_input) and query_input.is_cuda and self._matrix is not None and query_input.numel() == self._d:
            t = query_input.contiguous()
            qdt = {"torch.float16": N.HDB_F16, "torch.float32": N.HDB_F32, "torch.float64": N.HDB_F64}.get(str(t.dtype))
            if qdt is not None and t.device.index == self._matrix.device:
                out = (C.c_uint64 * 2)()
                N.check(N.lib().hdb_query_digest(self._matrix._h, C.c_void_p(t.data_ptr()), qdt, N.HDB_DEVICE, 1, out))
                return ("digest", tuple(t.shape), int(out[0]), int(out[1]))
        return query_input

    @staticmethod
    def _hashable_key(query_input, top_k, return_similarities, filters, recency_bias, timestamp_key, metric, ann_percent):
        """hyperdb/hyperdb.py:1368-1379 with the array turned into a digest of its float64 values (value-equal queries
        share a key whatever their dtype or residence, as tuple(tolist()) keys do).  A CUDA-tensor query arrives here
        already digested (`_digest`)."""
        if _is_torch(query_input):
            query_input = query_input.detach().cpu().numpy()
        if isinstance(query_input, (np.ndarray, list)) or (isinstance(query_input, tuple) and not (query_input and query_input[0] == "digest")):
            q = np.asarray(query_input, dtype=np.float64)
            query_input = ("digest", q.shape) + query_digest_host(q)
        if filters is None:
            hashable_filters = None
        else:
            hashable_filters = tuple(
                (name, tuple(sorted(params.items())) if isinstance(params, dict) else tuple(params) if isinstance(params, list) else params)
                for name, params in filters)
        return (query_input, top_k, return_similarities, hashable_filters, recency_bias, timestamp_key, metric, ann_percent)

    def close(self):
        self._close_device()

    # -- filters -> row subset ---------------------------------------------------------------------
    def _metadata_mask(self, spec):
        for key in spec:
            if key not in self.metadata_keys:
                raise ValueError(f"Invalid key '{key}' in metadata_filter: not found in metadata_keys")
        tag = tuple(sorted(spec.items(), key=repr))
        if tag not in self._mask_cache:
            keep = np.ones(len(self.documents), bool)
            for key, value in spec.items():
                keep &= np.fromiter((_nested(d, key) == value for d in self.documents), bool, len(self.documents))
            self._mask_cache[tag] = keep
        return tag, self._mask_cache[tag]

    def _apply_filters(self, filters):
        """-> (lo, hi, keep bool[n] or None, hashable tag of the subset)."""
        n = len(self.documents)
        lo, hi, keep, tags = 0, n, None, []
        skip_seen = False
        for name, params in filters or []:
            if name not in ['key', 'metadata', 'sentence', 'skip_doc']:
                raise ValueError(f"Invalid filter name {name}")
            if name == 'skip_doc':
                if skip_seen:
                    continue                                  # only the FIRST skip_doc filter acts (hyperdb/hyperdb.py:1474-1481: break)
                skip_seen = True
                if abs(params) >= n:
                    print(f"The absolute value of skip_doc ({abs(params)}) is equal or greater than the total number of documents ({n}).")
                    raise Exception("The absolute value of skip_doc is equal or greater than the total number of documents")
                if params > 0:
                    lo = params
                elif params < 0:
                    hi = n + params
            elif name == 'metadata':
                if not self.metadata_keys:
                    raise ValueError("The 'metadata_keys' parameter has not been set in HyperDB(). Cannot filter by metadata.")
                tag, m = self._metadata_mask(dict(params))
                tags.append(tag)
                keep = m if keep is None else (keep & m)
            else:
                raise NotImplementedError(f"filter '{name}' is string processing outside the B200 hot path (SURVEY.md section 2)")
        return lo, hi, keep, (lo, hi, tuple(tags))

    def _set_subset(self, lo, hi, keep, tag):
        """Row range + mask of the LOCAL shard; skipped when the previous query used the same subset.  Returns the kept
        count over all shards."""
        m = self._matrix
        if tag != self._subset_tag:
            n_loc = m.shape[0]
            if self._perm is None:
                m.set_range(max(0, lo - self._lo), max(0, min(n_loc, hi - self._lo)))
                m.set_mask(None if keep is None else keep[self._lo:self._lo + n_loc])
            else:
                # clustered storage: per-row inputs travel in PHYSICAL order, and a skip_doc range is no longer contiguous
                local = np.ones(n_loc, bool) if keep is None else keep[self._lo:self._lo + n_loc].copy()
                local[:max(0, min(n_loc, lo - self._lo))] = False
                local[max(0, min(n_loc, hi - self._lo)):] = False
                m.set_range(0, n_loc)
                m.set_mask(None if local.all() else local[self._perm])
            self._subset_tag = tag
            self._decay_tag = None
            self._kept = self._sm.total_kept() if self._sm is not None else m.n_kept
        return self._kept

    # -- time decay ----------------------------------------------------------------------------------
    def _timestamps(self, timestamp_key):
        if timestamp_key not in self._ts_cache:
            raw = [_nested(d, timestamp_key) for d in self.documents]
            missing = np.fromiter((t is None for t in raw), bool, len(raw))
            ts = np.array([0.0 if t is None else t for t in raw], dtype=float)
            self._ts_cache[timestamp_key] = (ts, missing)
        return self._ts_cache[timestamp_key]

    def _set_decay(self, lo, hi, keep, tag, timestamp_key, recency_bias):
        dtag = (tag, timestamp_key, float(recency_bias))
        if dtag == self._decay_tag:
            return
        m = self._matrix
        ts, missing = self._timestamps(timestamp_key)
        if missing.any():
            sub = missing[lo:hi] if keep is None else (missing[lo:hi] & keep[lo:hi])
            if sub.any():
                raise ValueError("All timestamps must be populated when recency_bias is not 0 or timestamp_key is provided.")
        n_loc = m.shape[0]
        if timestamp_key not in self._ts_dev:
            local_ts = ts[self._lo:self._lo + n_loc]
            if self._perm is not None:
                local_ts = local_ts[self._perm]
            self._ts_dev[timestamp_key] = m.stage_column(local_ts)
        m.set_timestamps(self._ts_dev[timestamp_key])           # device -> device: the column is transformed in place below

        def global_max():
            mx, cnt = m.kept_ts_max()
            if self._sm is None:
                return mx
            import torch
            t = torch.tensor([mx if cnt > 0 else float("-inf")], dtype=torch.float64, device=self._sm._comm_device())
            self._sm.dist.all_reduce(t, op=self._sm.dist.ReduceOp.MAX, group=self._sm.group)
            return t.item()

        m.stage1_recency(recency_bias, global_max())            # what _handle_timestamps returns (:1344-1346)
        m.set_decay_reference(global_max())                     # the sort's own transform (ranking_algorithm.py:183)
        self._decay_tag = dtag

    # -- the query ---------------------------------------------------------------------------------
    def _query_vector(self, query_input):
        if isinstance(query_input, str):
            if self.embedding_function is None:
                raise ValueError("text queries need an `embedding_function` (no embedding model is bundled)")
            out = self.embedding_function([query_input])
            q = out[0] if isinstance(out, tuple) else out
            if _is_torch(q):
                return self._tensor_query(q)
            q = np.asarray(q)
        elif _is_torch(query_input):
            return self._tensor_query(query_input)
        elif isinstance(query_input, (list, np.ndarray, tuple)):
            q = np.array(query_input)
            if q.dtype.kind not in "iuf":
                raise ValueError("Numeric array-like query_input expected.")
            if q.ndim > 2:
                raise ValueError("query_input must be a 1D or 2D array.")
        else:
            raise ValueError("query_input must be either a string or a numeric array-like object.")
        q = np.squeeze(q)
        if q.size == 0:
            raise ValueError("The generated query vector is empty.")
        if q.shape[-1] != self._d:
            raise ValueError(f"The dimension of the query_vector ({q.shape[-1]}) must match the dimension of the vectors in the database ({self._d}).")
        # the reference's LRU key round-trips an ndarray through tuple(tolist()): always float64
        return np.asarray(q, dtype=np.float64) if not isinstance(query_input, str) else q

    def _tensor_query(self, t):
        """A torch tensor stays where it is: a CUDA query is ranked without touching the host (float64 on the device,
        the dtype the reference's cache round trip gives every vector query)."""
        import torch
        if t.dim() > 2:
            raise ValueError("query_input must be a 1D or 2D array.")
        t = t.squeeze()
        if t.numel() == 0:
            raise ValueError("The generated query vector is empty.")
        if t.shape[-1] != self._d:
            raise ValueError(f"The dimension of the query_vector ({t.shape[-1]}) must match the dimension of the vectors in the database ({self._d}).")
        if not t.is_floating_point() or t.dtype == torch.bfloat16:
            t = t.to(torch.float64)
        return t.to(torch.float64) if t.is_cuda else t.to(torch.float64).numpy()

    def query(self, query_input=None, top_k=5, return_similarities=True, filters=None, recency_bias=0, timestamp_key=None,
              metric='cosine_similarity', ann_percent=5, query_vector=None):
        """hyperdb/hyperdb.py:1584-1586 -> :1429-1582 (brute-force branch).  `query_vector=` is accepted as an alias of
        the positional `query_input` (README.md:33 of the reference)."""
        if query_input is None:
            query_input = query_vector
        if self._n == 0 or not self.documents:
            raise Exception("The database is empty. Cannot proceed with the query.")
        if metric not in _METRICS:
            raise ValueError(f"Invalid metric '{metric}'. Supported: 'dot_product', 'cosine_similarity', 'euclidean_metric', 'manhattan_distance', 'jaccard_similarity', 'pearson_correlation', 'hamming_distance'")
        key = None
        if self.lru_cache.maxsize > 0:
            key = self._hashable_key(self._digest(query_input), top_k, return_similarities, filters, recency_bias, timestamp_key, metric, ann_percent)
            if key in self.lru_cache:                        # hyperdb/hyperdb.py:1381-1384
                self.cache_hits += 1
                return self.lru_cache[key]
            self.cache_misses += 1
        results = self._execute_query(query_input, top_k, return_similarities, filters, recency_bias, timestamp_key, metric)
        if key is not None:
            self.lru_cache[key] = results                    # evicts the least recently used entry beyond maxsize
        return results

    def _rank(self, q, top_k, metric, bias):
        if self._sm is not None:
            idx, sc, cnt = self._sm.query(q, int(top_k), metric, bias)[:3]
        else:
            idx, sc, cnt, _flags = self._matrix.query(q, int(top_k), metric, bias)
        return idx, sc, cnt

    def _first_index(self, row):
        """documents.index(documents[row]) (hyperdb/hyperdb.py:1567): the FIRST document equal to the hit, from a map built
        once per mutation instead of a linear scan per result."""
        if self._first_equal is None:
            table = {}
            try:
                for i, d in enumerate(self.documents):
                    table.setdefault(_canon(d), i)
            except TypeError:
                table = False                                # unhashable content: fall back to list.index per result
            self._first_equal = table
        if self._first_equal is False:
            return self.documents.index(self.documents[row])
        return self._first_equal[_canon(self.documents[row])]

    def _prepare(self, filters, recency_bias, timestamp_key, top_k):
        lo, hi, keep, tag = self._apply_filters(filters)
        kept = self._set_subset(lo, hi, keep, tag)
        if kept == 0:
            print("INFO: No document matches your query with the brute-force method and the current filters.")
            return None, 0.0
        if top_k > kept:
            print(f"Warning: top_k ({top_k}) is greater than the number of filtered documents ({kept}). Setting top_k to {kept}.")
            top_k = kept
        bias = 0.0
        if recency_bias != 0:
            timestamp_key = timestamp_key or "timestamp"
            if timestamp_key not in self.metadata_keys:
                raise ValueError(f"The timestamp_key '{timestamp_key}' must be present in metadata_keys when recency_bias is not 0.")
            self._set_decay(lo, hi, keep, tag, timestamp_key, recency_bias)
            bias = float(recency_bias)
        if top_k <= 0:
            raise ValueError("max() arg is an empty sequence")    # what max(ranked_results) raises in the reference (:1558)
        return top_k, bias

    def _execute_query(self, query_input, top_k, return_similarities, filters, recency_bias, timestamp_key, metric):
        """hyperdb/hyperdb.py:1429-1582, brute-force branch."""
        try:
            q = self._query_vector(query_input)
            print(f"INFO: Metric '{metric}': exact brute-force ranking on the GPU (hyperdb_b200 has no ANN index). Bruteforce method used instead.")
            top_k, bias = self._prepare(filters, recency_bias, timestamp_key, top_k)
            if top_k is None:
                return []
            idx, sc, cnt = self._rank(q, top_k, metric, bias)
            return self._results(idx[0], sc[0], int(cnt[0]), return_similarities)
        except (ValueError, TypeError) as e:
            print(f"An exception occurred due to invalid input: {e}")
            raise e
        except Exception as e:
            print(f"An unknown exception occurred: {e}")
            raise

    def _results(self, idx, sc, cnt, return_similarities):
        results = []
        for j in range(cnt):
            row = int(idx[j])
            doc = self.documents[row]
            results.append((doc, sc[j], self.source_indices[self._first_index(row)]) if return_similarities else doc)
        return results

    def query_batch(self, queries, top_k=5, return_similarities=True, filters=None, recency_bias=0, timestamp_key=None,
                    metric='cosine_similarity', native_dtype=False):
        """B queries in one call -- `queries` is a (B, d) array, a CUDA tensor (e.g. the embedding model's output, used as
        the query tile without a host round trip) or a list of texts for `embedding_function`.  One result list per query,
        each equal to what `query` returns for that row (the reference has no batched API: hyperdb/hyperdb.py:1584 is
        called once per query).  The query cache is not consulted.

        native_dtype=False (default): the tile is ranked as FLOAT64, the dtype `query`'s cache round trip gives every vector
        query (hyperdb/hyperdb.py:1369-1370) -- queries wider than the stored matrix: up to 8 of them share one read of the
        matrix (multi-query sweep), the scores are float64 arithmetic.
        native_dtype=True: a float16 / float32 / float64 tile is ranked in ITS OWN dtype -- the arithmetic of
        `hyperDB_ranking_algorithm_sort(vectors, q)` called with that array (hyperdb/ranking_algorithm.py:116-204) instead of
        `query`'s float64 detour.  A tile no wider than the stored matrix (the usual case: an fp16 / fp32 embedding model
        over an fp16 / fp32 store) then runs as ONE tensor-core contraction for dot / cosine / euclidean / pearson."""
        if self._n == 0 or not self.documents:
            raise Exception("The database is empty. Cannot proceed with the query.")
        if metric not in _METRICS:
            raise ValueError(f"Invalid metric '{metric}'. Supported: 'dot_product', 'cosine_similarity', 'euclidean_metric', 'manhattan_distance', 'jaccard_similarity', 'pearson_correlation', 'hamming_distance'")
        if isinstance(queries, (list, tuple)) and queries and isinstance(queries[0], str):
            if self.embedding_function is None:
                raise ValueError("text queries need an `embedding_function` (no embedding model is bundled)")
            out = self.embedding_function(list(queries))
            queries = out[0] if isinstance(out, tuple) else out
        if _is_torch(queries):
            import torch
            Q = queries if queries.dim() == 2 else queries.reshape(1, -1)
            if not (native_dtype and Q.dtype in (torch.float16, torch.float32, torch.float64)):
                Q = Q.to(torch.float64)
            Q = Q if Q.is_cuda else Q.numpy()
        else:
            Q = np.asarray(queries)
            if not (native_dtype and Q.dtype in (np.float16, np.float32, np.float64)):
                Q = np.asarray(Q, dtype=np.float64)
            Q = np.atleast_2d(Q)
        if Q.shape[-1] != self._d:
            raise ValueError(f"The dimension of the query_vector ({Q.shape[-1]}) must match the dimension of the vectors in the database ({self._d}).")
        top_k, bias = self._prepare(filters, recency_bias, timestamp_key, top_k)
        if top_k is None:
            return [[] for _ in range(Q.shape[0])]
        idx, sc, cnt = self._rank(Q, top_k, metric, bias)
        return [self._results(idx[b], sc[b], int(cnt[b]), return_similarities) for b in range(Q.shape[0])]
