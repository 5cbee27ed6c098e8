"""Seeded input generator shared by make_golden.py (which runs the REAL reference in the build
container) and the parity tests (which replay the same inputs against the oracle and the CUDA path).

Inputs are regenerated from a seed; `digest(...)` is stored next to the reference outputs so a drift
of NumPy's Generator stream would be detected instead of silently comparing different inputs.
"""
from __future__ import annotations

import hashlib
import itertools

import numpy as np

METRICS = ("dot_product", "cosine_similarity", "euclidean_metric", "manhattan_distance", "hamming_distance",
           "jaccard_similarity", "pearson_correlation")
DT = {"f16": np.float16, "f32": np.float32, "f64": np.float64}


def make_inputs(case):
    """case: dict(seed, n, d, vdt, qdt, kind, ts) -> (V, q, timestamps|None)."""
    rng = np.random.default_rng(case["seed"])
    n, d = case["n"], case["d"]
    kind = case.get("kind", "gauss")
    base = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal(d).astype(np.float32)
    if kind == "unit":                      # the bench distribution: unit-norm rows
        base /= np.maximum(np.linalg.norm(base, axis=1, keepdims=True), 1e-30)
        q /= max(np.linalg.norm(q), 1e-30)
    elif kind == "scaled":                  # rows of very different lengths, one zero row
        base *= rng.uniform(0.05, 4.0, (n, 1)).astype(np.float32)
        base[n // 3] = 0
        q *= 1.7
    elif kind == "ties":                    # duplicated rows -> exact score ties
        reps = rng.integers(0, max(1, n // 4), n)
        base = base[reps]
    elif kind == "binary":                  # already-binary {0,1} data (hamming fast path of :118-120)
        base = (base > 0.3).astype(np.float32)
        q = (q > 0).astype(np.float32)
    elif kind == "coarse":                  # few distinct values -> many ties in every metric
        base = np.round(base * 2) / 2
        q = np.round(q * 2) / 2
    elif kind == "shifted":                 # non-zero means, a constant row, a zero row (pearson: NaN -> ranked last)
        base = base * rng.uniform(0.2, 2.0, (n, 1)).astype(np.float32) + rng.uniform(-3.0, 3.0, (n, 1)).astype(np.float32)
        base[n // 4] = 0.75
        base[n // 2] = 0
        q = q * 0.5 + 1.25
    elif kind == "constq":                  # constant query: every pearson score is NaN
        q[:] = 0.5
    V = base.astype(DT[case["vdt"]])
    qv = q.astype(DT[case["qdt"]])
    ts = None
    if case.get("ts"):
        ts = 1.7e9 + rng.uniform(0, case.get("ts_span", 5.0), n)
    return V, qv, ts


def digest(V, q, ts):
    h = hashlib.sha1()
    h.update(np.ascontiguousarray(V).tobytes())
    h.update(np.ascontiguousarray(q).tobytes())
    if ts is not None:
        h.update(np.ascontiguousarray(ts).tobytes())
    return h.hexdigest()


def sort_cases():
    """The case matrix for hyperDB_ranking_algorithm_sort."""
    cases = []
    seed = 1000
    shapes = [(3, 2), (17, 3), (64, 7), (200, 8), (257, 33), (300, 100), (512, 128), (400, 129),
              (384, 384), (256, 768), (130, 1000), (96, 1536)]
    for (n, d), vdt, metric in itertools.product(shapes, ("f16", "f32", "f64"), METRICS):
        seed += 1
        kind = ("gauss", "unit", "scaled", "ties", "coarse")[seed % 5]
        cases.append(dict(seed=seed, n=n, d=d, vdt=vdt, qdt=vdt, kind=kind, metric=metric,
                          k=(1, 5, 10, 100, 7)[seed % 5], ts=(seed % 3 == 0), bias=(0.3, 1.0, 0.05)[seed % 3]))
    # mixed precision (NumPy promotes the whole matrix)
    for (vdt, qdt), metric in itertools.product((("f32", "f64"), ("f16", "f32"), ("f16", "f64"), ("f32", "f16")), METRICS):
        seed += 1
        cases.append(dict(seed=seed, n=150, d=96, vdt=vdt, qdt=qdt, kind="gauss", metric=metric, k=10,
                          ts=False, bias=0.0))
    # binary data for hamming, big-k, k > n, k <= 0
    for vdt in ("f16", "f32", "f64"):
        for metric in ("hamming_distance", "jaccard_similarity"):
            seed += 1
            cases.append(dict(seed=seed, n=500, d=256, vdt=vdt, qdt=vdt, kind="binary", metric=metric,
                              k=10, ts=False, bias=0.0))
    # an all-non-positive row and query: empty union -> 0/0 -> NaN -> ranked last
    seed += 1
    cases.append(dict(seed=seed, n=60, d=9, vdt="f32", qdt="f32", kind="coarse", metric="jaccard_similarity", k=60, ts=False, bias=0.0))
    # pearson on data with non-zero means / constant rows / a constant query, own and mixed dtypes
    for (vdt, qdt), (n, d), kind in itertools.product((("f16", "f16"), ("f32", "f32"), ("f64", "f64"), ("f16", "f64"), ("f32", "f64")),
                                                      ((120, 24), (90, 200)), ("shifted", "constq")):
        seed += 1
        cases.append(dict(seed=seed, n=n, d=d, vdt=vdt, qdt=qdt, kind=kind, metric="pearson_correlation", k=12,
                          ts=(seed % 2 == 0), bias=0.2))
    for k in (0, -1, 50, 1000):
        seed += 1
        cases.append(dict(seed=seed, n=40, d=16, vdt="f32", qdt="f32", kind="gauss", metric="cosine_similarity",
                          k=k, ts=False, bias=0.0))
    # decay-dominated ranking
    for metric in METRICS:
        seed += 1
        cases.append(dict(seed=seed, n=333, d=48, vdt="f32", qdt="f32", kind="unit", metric=metric, k=20,
                          ts=True, ts_span=3600.0, bias=0.3))
    return cases
