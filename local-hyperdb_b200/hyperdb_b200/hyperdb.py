"""Host shim that keeps `HyperDB.query`'s API (hyperdb/hyperdb.py:1584) over the device-resident matrix.

Only the brute-force branch of `_execute_query` (hyperdb/hyperdb.py:1429-1582) exists here: there is no Annoy
index, every query is an exact sweep on the GPU.  What the reference does with Python lists of rows becomes
state of the DeviceMatrix:

  skip_doc filter   (:1119-1134, :1474-1481)  -> kept row range            (hdb_matrix_set_range)
  metadata filter   (:1218-1257)              -> 1 bit per row             (hdb_matrix_set_mask)
  _handle_timestamps (:1310-1346)             -> stage-1 recency ON the device, then the sort's own
                                                 transform (recency is applied twice)
  row -> document   (:1565-1573)              -> global row id == document index (one row per document)

Quirk kept on purpose: a vector query reaches the ranking as FLOAT64 whatever dtype the caller used (the LRU
key is `tuple(query.tolist())`, :1369-1370), so NumPy promotes the stored matrix and the scores are float64
arithmetic on the stored values.  Out of scope: embedding model, chunking, 'key' and
'sentence' filters, persistence, the LRU cache itself.
"""
from __future__ import annotations

import hashlib
import sys
import time
from collections import OrderedDict

import numpy as np

from .device_matrix import DeviceMatrix

_METRICS = ['dot_product', 'cosine_similarity', 'euclidean_metric', 'manhattan_distance', 'jaccard_similarity',
            'pearson_correlation', 'hamming_distance']


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


class HyperDB:
    def __init__(self, documents=None, vectors=None, select_keys=None, embedding_function=None, fp_precision="float32",
                 add_timestamp=False, metadata_keys=None, ann_metric="cosine", n_trees=10, cache_size=256, device=None):
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
        self.vectors = None                      # host copy, as the reference keeps it
        self.source_indices = []
        self._matrix = None
        self._mask_cache = {}
        # query cache (hyperdb/hyperdb.py:60-62, :1368-1388): same semantics, but the key is a 128-bit digest of the
        # query's bytes instead of tuple(query.tolist()) -- O(D) C speed instead of D Python floats per lookup
        self.lru_cache = _LRUCache(cache_size)
        self.cache_hits = 0
        self.cache_misses = 0
        if vectors is not None:
            if documents is None or len(documents) != len(vectors):
                raise ValueError("documents and vectors must have the same length")
            self.documents = list(documents)
            self.vectors = np.asarray(vectors)   # stored as given (hyperdb/hyperdb.py:127-135)
            self.source_indices = list(range(len(self.documents)))
            self._upload()
        elif documents:
            self.add(documents)

    # -- storage ---------------------------------------------------------------------------------
    def _upload(self):
        self.clear_cache()
        if self._matrix is not None:
            self._matrix.close()
        self._matrix = DeviceMatrix(self.vectors, device=self.device) if self.vectors is not None and len(self.vectors) else None
        self._mask_cache.clear()

    def add(self, documents, vectors=None, add_timestamp=False):
        """hyperdb/hyperdb.py:496-545, :626-689 reduced to: embed (if needed), cast to fp_precision, append on the device
        (hdb_matrix_append: only the new rows are uploaded and ingested)."""
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
        if self.vectors is None or self._matrix is None:
            previous, self.vectors = self.vectors, vectors
            try:
                self._upload()                               # raises ValueError on NaN
            except Exception:
                self.vectors = previous
                raise
        else:
            vectors = vectors.astype(self.vectors.dtype)
            self._matrix.append(vectors)                     # device side: only the new rows are copied and ingested
            self.vectors = np.concatenate([self.vectors, vectors])
            self._mask_cache.clear()
            self.clear_cache()                               # hyperdb/hyperdb.py:566
        base = len(self.documents)
        self.documents.extend(documents)
        self.source_indices.extend(range(base, base + len(documents)))

    def remove_document(self, index):
        """hyperdb/hyperdb.py:691-766 for one-row-per-document stores."""
        keep = np.ones(len(self.documents), bool)
        keep[index] = False
        self.documents = [d for d, k in zip(self.documents, keep) if k]
        self.vectors = self.vectors[keep]
        self.source_indices = list(range(len(self.documents)))
        if self._matrix is not None and len(self.vectors):
            self._matrix.remove_rows(np.flatnonzero(~keep))  # stable compaction on the device, no re-upload
            self._mask_cache.clear()
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

    @staticmethod
    def _hashable_key(query_input, top_k, return_similarities, filters, recency_bias, timestamp_key, metric, ann_percent):
        """hyperdb/hyperdb.py:1368-1379 with the array turned into a digest (value-equal float64 queries share a key, as
        tuple(tolist()) keys do)."""
        if isinstance(query_input, (np.ndarray, list, tuple)):
            q = np.ascontiguousarray(np.asarray(query_input, dtype=np.float64))
            query_input = (q.shape, hashlib.blake2b(q.tobytes(), digest_size=16).digest())
        if filters is None:
            hashable_filters = None
        else:
            hashable_filters = tuple(
                (name, tuple(sorted(params.items())) if isinstance(params, dict) else tuple(params) if isinstance(params, list) else params)
                for name, params in filters)
        return (query_input, top_k, return_similarities, hashable_filters, recency_bias, timestamp_key, metric, ann_percent)

    def close(self):
        if self._matrix is not None:
            self._matrix.close()
            self._matrix = None

    # -- filters -> row subset ---------------------------------------------------------------------
    def _metadata_mask(self, spec):
        for key in spec:
            if key not in self.metadata_keys:
                raise ValueError(f"Invalid key '{key}' in metadata_filter: not found in metadata_keys")
        tag = tuple(sorted(spec.items()))
        if tag not in self._mask_cache:
            keep = np.ones(len(self.documents), bool)
            for key, value in spec.items():
                keep &= np.fromiter((_nested(d, key) == value for d in self.documents), bool, len(self.documents))
            self._mask_cache[tag] = keep
        return self._mask_cache[tag]

    def _apply_filters(self, filters):
        n = len(self.documents)
        lo, hi, keep = 0, n, None
        for name, params in filters or []:
            if name not in ['key', 'metadata', 'sentence', 'skip_doc']:
                raise ValueError(f"Invalid filter name {name}")
            if name == 'skip_doc':
                if abs(params) >= n:
                    print(f"The absolute value of skip_doc ({abs(params)}) is equal or greater than the total number of documents ({n}).")
                    raise Exception("The absolute value of skip_doc is equal or greater than the total number of documents")
                if params > 0:
                    lo = max(lo, params)
                elif params < 0:
                    hi = min(hi, n + params)
            elif name == 'metadata':
                if not self.metadata_keys:
                    raise ValueError("The 'metadata_keys' parameter has not been set in HyperDB(). Cannot filter by metadata.")
                m = self._metadata_mask(dict(params))
                keep = m if keep is None else (keep & m)
            else:
                raise NotImplementedError(f"filter '{name}' is string processing outside the B200 hot path (SURVEY.md section 2)")
        return lo, hi, keep

    # -- the query ---------------------------------------------------------------------------------
    def _query_vector(self, query_input):
        if isinstance(query_input, str):
            if self.embedding_function is None:
                raise ValueError("text queries need an `embedding_function` (no embedding model is bundled)")
            out = self.embedding_function([query_input])
            q = np.asarray(out[0] if isinstance(out, tuple) else out)
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
        if q.shape[-1] != self.vectors.shape[1]:
            raise ValueError(f"The dimension of the query_vector ({q.shape[-1]}) must match the dimension of the vectors in the database ({self.vectors.shape[1]}).")
        # the reference's LRU key round-trips an ndarray through tuple(tolist()): always float64
        return np.array(tuple(np.asarray(q).tolist()), dtype=np.float64) if not isinstance(query_input, str) else q

    def query(self, query_input=None, top_k=5, return_similarities=True, filters=None, recency_bias=0, timestamp_key=None,
              metric='cosine_similarity', ann_percent=5, query_vector=None):
        """hyperdb/hyperdb.py:1584-1586 -> :1429-1582 (brute-force branch).  `query_vector=` is accepted as an alias of
        the positional `query_input` (README.md:33 of the reference)."""
        if query_input is None:
            query_input = query_vector
        if self.vectors is None or len(self.vectors) == 0 or not self.documents:
            raise Exception("The database is empty. Cannot proceed with the query.")
        if metric not in _METRICS:
            raise ValueError(f"Invalid metric '{metric}'. Supported: 'dot_product', 'cosine_similarity', 'euclidean_metric', 'manhattan_distance', 'jaccard_similarity', 'pearson_correlation', 'hamming_distance'")
        key = None
        if self.lru_cache.maxsize > 0:
            key = self._hashable_key(query_input, top_k, return_similarities, filters, recency_bias, timestamp_key, metric, ann_percent)
            if key in self.lru_cache:                        # hyperdb/hyperdb.py:1381-1384
                self.cache_hits += 1
                return self.lru_cache[key]
            self.cache_misses += 1
        results = self._execute_query(query_input, top_k, return_similarities, filters, recency_bias, timestamp_key, metric)
        if key is not None:
            self.lru_cache[key] = results                    # evicts the least recently used entry beyond maxsize
        return results

    def _execute_query(self, query_input, top_k, return_similarities, filters, recency_bias, timestamp_key, metric):
        """hyperdb/hyperdb.py:1429-1582, brute-force branch."""
        try:
            q = self._query_vector(query_input)
            lo, hi, keep = self._apply_filters(filters)
            print(f"INFO: Metric '{metric}': exact brute-force ranking on the GPU (hyperdb_b200 has no ANN index). Bruteforce method used instead.")
            m = self._matrix
            m.set_range(lo, hi)
            m.set_mask(keep)
            kept = m.n_kept
            if kept == 0:
                print("INFO: No document matches your query with the brute-force method and the current filters.")
                return []
            if top_k > kept:
                print(f"Warning: top_k ({top_k}) is greater than the number of filtered documents ({kept}). Setting top_k to {kept}.")
                top_k = kept
            bias = 0.0
            if recency_bias != 0:
                timestamp_key = timestamp_key or "timestamp"
                if timestamp_key not in self.metadata_keys:
                    raise ValueError(f"The timestamp_key '{timestamp_key}' must be present in metadata_keys when recency_bias is not 0.")
                ts = [_nested(d, timestamp_key) for d in self.documents]
                rows = np.arange(lo, hi) if keep is None else np.flatnonzero(keep[lo:hi]) + lo
                if any(ts[i] is None for i in rows):
                    raise ValueError("All timestamps must be populated when recency_bias is not 0 or timestamp_key is provided.")
                ts = np.array([0.0 if t is None else t for t in ts], dtype=float)
                m.set_timestamps(ts)
                stage1_max, _ = m.kept_ts_max()
                m.stage1_recency(recency_bias, stage1_max)          # what _handle_timestamps returns (:1344-1346)
                stage2_max, _ = m.kept_ts_max()
                m.set_decay_reference(stage2_max)                   # the sort's own transform (ranking_algorithm.py:183)
                bias = float(recency_bias)
            else:
                m.set_timestamps(None)
            idx, sc, cnt, _flags = m.query(q, int(top_k), metric, bias)
            results = []
            for j in range(int(cnt[0])):
                row = int(idx[0, j])
                doc = self.documents[row]
                results.append((doc, sc[0, j], self.source_indices[row]) if return_similarities else doc)
            return results
        except (ValueError, TypeError) as e:
            print(f"An exception occurred due to invalid input: {e}")
            raise e
        except Exception as e:
            print(f"An unknown exception occurred: {e}")
            raise
