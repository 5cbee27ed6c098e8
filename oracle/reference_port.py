"""NumPy port of the reference's brute-force ranking path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function cites the reference lines it restates (paths relative to /root/reference).
The arithmetic is delegated to the same NumPy primitives the reference calls, so on one machine
this port and the reference produce identical bits (checked by tests/golden/make_golden.py when
the golden vectors are generated, and by tests/test_oracle_port.py against those vectors).

Differences from the reference, all deliberate:
  * inputs are never mutated (the reference binarises float inputs in place,
    hyperdb/ranking_algorithm.py:122-124 -- SURVEY.md quirk 9);
  * `rank(...)` has a `canonical=True` mode that breaks ties by LOWER INDEX
    (north_star's definition; the reference's argpartition/argsort order among equal scores is
    unspecified);
  * a row-chunked driver (`chunked_rank`) for matrices that do not fit in host RAM.
"""
from __future__ import annotations

import numpy as np

METRICS = (
    "dot_product",
    "cosine_similarity",
    "euclidean_metric",
    "manhattan_distance",
    "hamming_distance",
    "jaccard_similarity",
    "pearson_correlation",
)


# ----------------------------------------------------------------------------------------------
# metric functions
# ----------------------------------------------------------------------------------------------
def unit_rows(x):
    """hyperdb/ranking_algorithm.py:8-21 (get_norm_vector).

    L2-normalise along the last axis in the array's own dtype; a zero norm divides by 1.
    (The reference also prints a warning when NaNs are present; the sort rejects NaNs earlier.)
    """
    x = np.asarray(x)
    length = np.linalg.norm(x, axis=-1, keepdims=True)
    length[length == 0] = 1
    return x / length


def dot_scores(vectors, query):
    """hyperdb/ranking_algorithm.py:24-30 (dot_product): np.dot(V, q.T), not flattened."""
    return np.dot(vectors, np.asarray(query).T)


def cosine_scores(vectors, query):
    """hyperdb/ranking_algorithm.py:32-42 (cosine_similarity)."""
    return np.dot(unit_rows(vectors), unit_rows(query).T).flatten()


def euclidean_scores(vectors, query, get_similarity_score=True):
    """hyperdb/ranking_algorithm.py:44-52 (euclidean_metric): 1/(1+||v-q||_2)."""
    dist = np.linalg.norm(vectors - query, axis=1)
    return 1 / (1 + dist) if get_similarity_score else dist


def manhattan_scores(vectors, query):
    """hyperdb/ranking_algorithm.py:54-61 (manhattan_distance): 1/(1+sum|v-q|)."""
    return 1 / (1 + np.sum(np.abs(vectors - query), axis=1))


def sign_bits(x):
    """hyperdb/ranking_algorithm.py:116-126 (check_and_binarize_vectors) without the mutation.

    Already-binary input ({0,1}) is returned as is; anything else becomes (x > 0).  Both cases
    are the predicate x > 0, so the np.unique scan of the reference is not needed for the value.
    """
    return (np.asarray(x) > 0).astype(np.uint8)


def hamming_scores(vectors, query):
    """hyperdb/ranking_algorithm.py:128-147 (hamming_distance): D - popcount(bits(v) xor bits(q)).

    Returned as uint64 like the reference (np.sum of unpackbits is uint64; `D - hd` keeps it).
    """
    vb, qb = sign_bits(vectors), sign_bits(query)
    differing = np.sum(np.unpackbits(np.bitwise_xor(vb, qb), axis=1), axis=1)
    return np.asarray(vectors).shape[-1] - differing


def jaccard_scores(vectors, query):
    """hyperdb/ranking_algorithm.py:63-76 (jaccard_similarity): popcount(v & q) / popcount(v | q) on the sign bits,
    float64; an empty union gives 0/0 = NaN (ranked last by the sort, :174)."""
    vb, qb = sign_bits(vectors), sign_bits(query)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.sum(np.bitwise_and(vb, qb), axis=1) / np.sum(np.bitwise_or(vb, qb), axis=1)


def pearson_scores(vectors, query):
    """hyperdb/ranking_algorithm.py:78-113 (pearson_correlation): covariance sum / (std_v * std_q * D) with np.mean / np.std
    taken in each operand's own dtype; float64 output (np.zeros(N)); NaN when either side is constant, 0 where the
    denominator underflowed to zero."""
    vectors = np.asarray(vectors)
    query = np.asarray(query).flatten()
    q_mean, v_mean = np.mean(query), np.mean(vectors, axis=1)
    q_std, v_std = np.std(query), np.std(vectors, axis=1)
    numerator = np.sum((vectors - v_mean[:, np.newaxis]) * (query - q_mean), axis=1)
    denominator = v_std * q_std * vectors.shape[1]
    ok = denominator != 0
    out = np.zeros(vectors.shape[0])
    out[ok] = numerator[ok] / denominator[ok]
    out[(q_std == 0) | (v_std == 0)] = np.nan          # `:108-111`: both constant, or exactly one of them
    return out


_DISPATCH = {
    "dot_product": dot_scores,
    "cosine_similarity": cosine_scores,
    "euclidean_metric": euclidean_scores,
    "manhattan_distance": manhattan_scores,
    "hamming_distance": hamming_scores,
    "jaccard_similarity": jaccard_scores,
    "pearson_correlation": pearson_scores,
}


# ----------------------------------------------------------------------------------------------
# the sort
# ----------------------------------------------------------------------------------------------
def similarities_f64(vectors, query, metric):
    """hyperdb/ranking_algorithm.py:150-174: NaN guard, metric dispatch, cast to float64, NaN -> -inf."""
    if np.isnan(vectors).any() or np.isnan(query).any():
        raise ValueError("Vectors and query_vector should not contain NaN values.")
    fn = _DISPATCH.get(metric)
    if fn is None:
        raise ValueError(f"Unknown metric: {metric}")
    sims = fn(np.array(vectors), query).astype(float)
    sims[np.isnan(sims)] = -np.inf
    return sims


def recency_term(timestamps, recency_bias, n, ts_max=None):
    """hyperdb/ranking_algorithm.py:179-183: bias * exp(ts - max ts); zeros if no timestamps.

    `ts_max` lets the chunked driver pass the GLOBAL maximum.
    """
    if timestamps is None or len(timestamps) == 0:
        return np.zeros(n)
    ts = np.asarray(timestamps)
    top = np.max(ts) if ts_max is None else ts_max
    return recency_bias * np.exp(-top + ts)


def select_top(scores, top_k, canonical):
    """hyperdb/ranking_algorithm.py:194-204.  canonical=True: (score desc, index asc)."""
    n = len(scores)
    k = max(0, min(top_k, n))
    if k == 0:
        return [], []
    scores = scores.flatten()
    if canonical:
        order = np.lexsort((np.arange(n), -scores))[:k]
    else:
        order = np.argpartition(scores, -k)[-k:]
        order = order[np.argsort(-scores[order])]
    return order, scores[order]


def rank(vectors, query, top_k=5, metric="cosine_similarity", timestamps=None, recency_bias=0,
         canonical=True):
    """hyperdb/ranking_algorithm.py:149-204 (hyperDB_ranking_algorithm_sort).

    Returns (indices int64[k], scores float64[k]) sorted by descending score.  With
    canonical=False the tie order is whatever argpartition/argsort give (as the reference);
    the N==1 special case (`:189-191`) is reproduced in both modes.
    """
    sims = similarities_f64(vectors, query, metric)
    scores = sims + recency_term(timestamps, recency_bias, len(sims))
    if np.array(scores).shape == () or (len(scores) == 1 and np.array(scores).ndim == 1):
        return np.array([0]), np.array([scores])
    if len(scores) == 0:
        raise UnboundLocalError("empty vectors (reference: ranking_algorithm.py:194-204)")
    return select_top(scores, top_k, canonical)


def chunked_rank(row_chunks, query, top_k, metric, timestamps=None, recency_bias=0, mask=None):
    """SURVEY.md section 8(c) 'chunked oracle': rows are independent except for max(ts) over the
    kept rows and the final select, so score chunk by chunk and select once.

    row_chunks: iterable of (row_offset, ndarray chunk).  mask: optional bool[N] of kept rows
    (the metadata filter of hyperdb/hyperdb.py:1492-1493 expressed on global row ids); indices
    returned are GLOBAL row ids.  The max over timestamps is over kept rows only
    (hyperdb/hyperdb.py:1334-1344).
    """
    parts, offsets = [], []
    for off, chunk in row_chunks:
        parts.append(similarities_f64(chunk, query, metric))
        offsets.append(off)
    sims = np.concatenate(parts)
    n = len(sims)
    keep = np.ones(n, bool) if mask is None else np.asarray(mask, bool)
    ids = np.flatnonzero(keep)
    ts = None if timestamps is None else np.asarray(timestamps, float)[ids]
    scores = sims[ids] + recency_term(ts, recency_bias, len(ids))
    order, top = select_top(scores, top_k, canonical=True)
    if len(order) == 0:
        return np.empty(0, np.int64), np.empty(0)
    return ids[order], top


# ----------------------------------------------------------------------------------------------
# brute-force tail of HyperDB._execute_query
# ----------------------------------------------------------------------------------------------
def stage1_recency(timestamps, recency_bias):
    """hyperdb/hyperdb.py:1334-1346 (_handle_timestamps): bias * exp(ts - max ts) over the
    filtered documents; this vector is then passed as `timestamps=` to the sort, which applies the
    transform a second time."""
    ts = np.asarray(timestamps, float)
    return recency_bias * np.exp(ts - np.max(ts))


def hyperdb_bruteforce_tail(vectors, query, top_k, metric, timestamps=None, recency_bias=0,
                            keep=None):
    """hyperdb/hyperdb.py:1541-1575 restated on row ids instead of documents.

    keep: optional bool[N] (filters).  Returns (global row ids, float64 scores).  top_k is clamped
    to the number of kept rows (`:1541-1543`); recency is the two-stage composition (quirk 1) and is
    only active when recency_bias != 0 and timestamps are given (`:1312-1313`).

    The query reaches _execute_query through the LRU key, where an ndarray was turned into
    `tuple(query.tolist())` (`hyperdb/hyperdb.py:1369-1370`) and is rebuilt by np.array (`:1201`):
    on this path the query is ALWAYS float64, whatever dtype the caller passed, so NumPy promotes the
    stored matrix to float64 for every query (probed against the real class: tests/golden/hyperdb_tail.npz).
    """
    query = np.array(tuple(np.asarray(query).tolist()))
    vectors = np.asarray(vectors)
    ids = np.arange(len(vectors)) if keep is None else np.flatnonzero(np.asarray(keep, bool))
    if len(ids) == 0:
        return np.empty(0, np.int64), np.empty(0)
    top_k = min(top_k, len(ids))
    stage1 = None
    if recency_bias != 0 and timestamps is not None:
        stage1 = stage1_recency(np.asarray(timestamps, float)[ids], recency_bias)
    order, top = rank(vectors[ids], query, top_k, metric, stage1, recency_bias, canonical=True)
    order = np.asarray(order, np.int64)
    return ids[order], np.asarray(top, float).reshape(-1)
