"""Operation-by-operation specification of the scores the reference produces
(TEST INFRASTRUCTURE, see oracle/__init__.py).

The reference delegates its arithmetic to NumPy (pinned numpy==1.26.3, requirements.txt:1; not
vendored under /root/reference).  The published algorithms of the NumPy loops on the path are
restated here so that the CUDA "certify" kernel (csrc/certify.cuh) has an exact spec:

  * HALF_dot  (np.dot on float16, hyperdb/ranking_algorithm.py:29,41): products and a SEQUENTIAL
    running sum in float32, one final round-to-nearest-even to float16.
  * @TYPE@_pairwise_sum (np.add.reduce along the contiguous axis, used by np.linalg.norm
    `:9,:49` and np.sum `:59`): blocks of <=128 elements with 8 interleaved accumulators combined as
    ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), recursive halving (n2 = n/2 rounded down to a multiple of
    8) above 128; float16 inputs are accumulated in float32 and rounded once at the end.
  * every other float16 ufunc (subtract, multiply, divide, abs, sqrt, 1+x, 1/x) = the float32
    operation followed by a round to float16.
  * float32/float64 np.dot goes to OpenBLAS sgemv/dgemv whose summation order is unknowable; the
    canonical value is DEFINED as the exact dot product of the (canonically normalised, for
    cosine) operands rounded once to the result dtype.  The reference lies within a few ulp of it
    (probed: <= 1.1e-7 abs for fp32, <= 1.8e-16 for fp64 on unit-scale data).

Result dtype R = np.promote_types(vectors.dtype, query.dtype), as NumPy promotes
(SURVEY.md quirk 4).  All functions are vectorised over rows and loop over the D columns in
Python, so they are for N <= ~1e5.
"""
from __future__ import annotations

import numpy as np

_F = {np.dtype(np.float16): np.float16, np.dtype(np.float32): np.float32, np.dtype(np.float64): np.float64}


def result_dtype(vectors, query):
    return np.promote_types(np.asarray(vectors).dtype, np.asarray(query).dtype)


def _as_float(x):
    x = np.asarray(x)
    if x.dtype.kind != "f":
        x = x.astype(np.float64)       # integer inputs behave as float64 after NumPy promotion
    return x


# ---------------------------------------------------------------------------------------------
# NumPy pairwise summation along axis 1
# ---------------------------------------------------------------------------------------------
def _pairwise(cols):
    """cols: (N, n) array already in the ACCUMULATION dtype.  Returns (N,) in that dtype."""
    n = cols.shape[1]
    if n < 8:
        acc = np.zeros(cols.shape[0], cols.dtype)
        for i in range(n):
            acc = acc + cols[:, i]
        return acc
    if n <= 128:
        r = [cols[:, j].copy() for j in range(8)]
        full = n - (n % 8)
        for i in range(8, full, 8):
            for j in range(8):
                r[j] = r[j] + cols[:, i + j]
        acc = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        for i in range(full, n):
            acc = acc + cols[:, i]
        return acc
    half = n // 2
    half -= half % 8
    return _pairwise(cols[:, :half]) + _pairwise(cols[:, half:])


def row_sum(x):
    """np.add.reduce(x, axis=1) bit for bit (float16: float32 accumulation, one final rounding)."""
    if x.dtype == np.float16:
        return _pairwise(x.astype(np.float32)).astype(np.float16)
    return _pairwise(x)


def _sqrt(x):
    if x.dtype == np.float16:
        return np.sqrt(x.astype(np.float32)).astype(np.float16)
    return np.sqrt(x)


def row_norm(x):
    """np.linalg.norm(x, axis=1): sqrt(add.reduce(x*x)) with the squares rounded to x.dtype."""
    with np.errstate(over="ignore", under="ignore"):
        return _sqrt(row_sum(x * x))


def unit_rows(x):
    """get_norm_vector (hyperdb/ranking_algorithm.py:8-21) with the arithmetic spelled out."""
    x2 = x if x.ndim == 2 else x[None, :]
    length = row_norm(x2)[:, None]
    length = np.where(length == 0, np.ones_like(length), length)
    with np.errstate(over="ignore", under="ignore", invalid="ignore"):
        out = x2 / length
    return out if x.ndim == 2 else out[0]


# ---------------------------------------------------------------------------------------------
# dot products
# ---------------------------------------------------------------------------------------------
def _dot(v, q):
    """v: (N, D), q: (D,), both already of the result dtype R."""
    if v.dtype == np.float16:                         # HALF_dot
        acc = np.zeros(v.shape[0], np.float32)
        vf, qf = v.astype(np.float32), q.astype(np.float32)
        with np.errstate(over="ignore", invalid="ignore"):
            for j in range(v.shape[1]):
                acc = acc + vf[:, j] * qf[j]
            return acc.astype(np.float16)
    # exact dot rounded once: long double accumulation of exact (fp32) / near-exact (fp64) products
    wide = np.longdouble
    acc = np.zeros(v.shape[0], wide)
    for j in range(v.shape[1]):
        acc = acc + v[:, j].astype(wide) * wide(q[j])
    return acc.astype(v.dtype)


# ---------------------------------------------------------------------------------------------
# pearson_correlation (hyperdb/ranking_algorithm.py:78-113): np.mean / np.std spelled out
# ---------------------------------------------------------------------------------------------
def _div_count(x, n, out_dtype):
    """um.true_divide(x, np.intp(n)) as numpy._core._methods._mean/_var call it: the count is a NumPy integer
    scalar, so the quotient is formed in float64 and cast to the output dtype (float32 sums first, for float16)."""
    return (x.astype(np.float64) / np.float64(n)).astype(out_dtype)


def row_mean(x, scalar=False):
    """np.mean(x, axis=1): float16 is summed in float32, divided, cast to float32 and only then to float16.
    scalar=True: np.mean of a 1-D array returns a NumPy scalar, whose float64 quotient is cast to float16 directly."""
    n = x.shape[1]
    if x.dtype == np.float16:
        s32 = _pairwise(x.astype(np.float32))
        return _div_count(s32, n, np.float16) if scalar else _div_count(s32, n, np.float32).astype(np.float16)
    return _div_count(_pairwise(x), n, x.dtype)


def row_std(x):
    """np.std(x, axis=1): mean of the squared deviations from np.add.reduce(x)/n, everything in x.dtype."""
    n = x.shape[1]
    with np.errstate(over="ignore", under="ignore", invalid="ignore"):
        arrmean = _div_count(row_sum(x), n, x.dtype)
        dev = x - arrmean[:, None]
        dev = dev * dev
        return _sqrt(_div_count(row_sum(dev), n, x.dtype))


def pearson(v, q):
    """pearson_correlation with each NumPy call replaced by its arithmetic.  v: (N, D), q: (D,) float arrays in their
    OWN dtypes (the statistics are taken before NumPy promotes).  Returns float64 (np.zeros(N) receives the quotients)."""
    q = q.reshape(-1)
    d = v.shape[1]
    R = _F[np.dtype(np.promote_types(v.dtype, q.dtype))]
    with np.errstate(over="ignore", under="ignore", invalid="ignore", divide="ignore"):
        q_mean, v_mean = row_mean(q[None, :], scalar=True)[0], row_mean(v)
        q_std, v_std = row_std(q[None, :])[0], row_std(v)
        prod = (v - v_mean[:, None]).astype(R) * (q - q_mean).astype(R)
        numerator = row_sum(prod)
        denominator = (v_std.astype(R) * R(q_std)) * R(d)
        out = np.zeros(v.shape[0])
        ok = denominator != 0
        out[ok] = numerator[ok] / denominator[ok]
        out[(v_std == 0) | (q_std == 0)] = np.nan
    return out


def euclidean_distance(vectors, query):
    """euclidean_metric(..., get_similarity_score=False), hyperdb/ranking_algorithm.py:49-52: np.linalg.norm(v - q, axis=1)
    in the promoted dtype (element-wise subtract and square, NumPy's pairwise sum, sqrt)."""
    v, q = _as_float(vectors), _as_float(query)
    R = _F[np.dtype(result_dtype(v, q))]
    with np.errstate(over="ignore", under="ignore", invalid="ignore", divide="ignore"):
        return row_norm(v.astype(R) - q.astype(R))


def scores(vectors, query, metric):
    """Canonical similarity vector in the reference's result dtype (uint64 for hamming)."""
    v, q = _as_float(vectors), _as_float(query)
    if metric == "hamming_distance":
        vb, qb = v > 0, q > 0
        return (v.shape[1] - np.sum(vb != qb[None, :], axis=1)).astype(np.uint64)
    if metric == "jaccard_similarity":
        vb, qb = v > 0, q > 0
        inter = np.sum(vb & qb[None, :], axis=1).astype(np.float64)
        union = np.sum(vb | qb[None, :], axis=1).astype(np.float64)
        with np.errstate(invalid="ignore", divide="ignore"):
            return inter / union
    if metric == "pearson_correlation":
        return pearson(v, q)
    R = _F[np.dtype(result_dtype(v, q))]
    if metric == "cosine_similarity":
        # each operand is normalised in ITS OWN dtype (ranking_algorithm.py:37-38); np.dot promotes after
        with np.errstate(over="ignore", under="ignore", invalid="ignore", divide="ignore"):
            return _dot(unit_rows(v).astype(R), unit_rows(q).astype(R))
    v, q = v.astype(R), q.astype(R)
    one = R(1)
    with np.errstate(over="ignore", under="ignore", invalid="ignore", divide="ignore"):
        if metric == "dot_product":
            return _dot(v, q)
        if metric == "euclidean_metric":
            return one / (one + row_norm(v - q))
        if metric == "manhattan_distance":
            return one / (one + row_sum(np.abs(v - q)))
    raise ValueError(f"Unknown metric: {metric}")


def total_scores(vectors, query, metric, timestamps=None, recency_bias=0.0, keep=None):
    """float64 ranking scores of hyperdb/ranking_algorithm.py:171-186 on the canonical similarities.
    keep (bool[N]) restricts both the rows and the max over timestamps.  Dropped rows get NaN."""
    s = scores(vectors, query, metric).astype(np.float64)
    s[np.isnan(s)] = -np.inf
    n = len(s)
    keep = np.ones(n, bool) if keep is None else np.asarray(keep, bool)
    if timestamps is not None and len(timestamps) > 0 and keep.any():
        ts = np.asarray(timestamps, np.float64)
        s = s + recency_bias * np.exp(-np.max(ts[keep]) + ts)
    s[~keep] = np.nan
    return s


def rank(vectors, query, top_k, metric, timestamps=None, recency_bias=0.0, keep=None):
    """Canonical top-k: (score desc, index asc) over kept rows; returns (global ids, f64 scores)."""
    s = total_scores(vectors, query, metric, timestamps, recency_bias, keep)
    ids = np.flatnonzero(~np.isnan(s))
    k = max(0, min(int(top_k), len(ids)))
    order = ids[np.lexsort((ids, -s[ids]))][:k]
    return order.astype(np.int64), s[order]
