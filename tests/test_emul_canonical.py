"""The PRODUCT's canonical arithmetic (local-hyperdb_b200/csrc/canonical.cuh -- the code the certify step and the exact
path run on the GPU) compiled for the HOST (tests/emul/canonical_host.cpp maps the CUDA rounding intrinsics to IEEE
operations) and compared with NumPy / the reference port on seeded random inputs.  The GPU tests check the same code on
the device against the golden vectors; this widens the input space without needing a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import reference_port as P

HERE = os.path.dirname(os.path.abspath(__file__))
DT = {np.dtype(np.float16): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}
DTS = (np.float16, np.float32, np.float64)
METRIC = {"dot_product": 0, "cosine_similarity": 1, "euclidean_metric": 2, "manhattan_distance": 3, "hamming_distance": 4,
          "jaccard_similarity": 5, "pearson_correlation": 6}


def _same(a, b):
    return (a == b) or (np.isnan(a) and np.isnan(b))


def test_pairwise_sum_norm_mean_std(emul):
    rng = np.random.default_rng(7)
    for trial in range(900):
        d = int(rng.integers(1, 2100))
        dt = DTS[trial % 3]
        x = np.ascontiguousarray((rng.standard_normal(d) * float(rng.choice([0.01, 1.0, 40.0])) + float(rng.choice([0.0, 3.0]))).astype(dt))
        p = C.c_void_p(x.ctypes.data)
        with np.errstate(all="ignore"):
            assert _same(emul.emul_pairwise_sum(DT[x.dtype], p, d), float(np.add.reduce(x))), (trial, d, dt)
            # get_norm_vector calls np.linalg.norm(axis=-1, keepdims=True): sqrt(add.reduce(x*x)), NOT the BLAS dot of axis=None
            assert _same(emul.emul_norm(DT[x.dtype], p, d), float(np.linalg.norm(x, axis=-1, keepdims=True)[0])), (trial, d, dt)
            m, s = C.c_double(), C.c_double()
            emul.emul_mean_std(DT[x.dtype], p, d, 1, C.byref(m), C.byref(s))            # 1-D input: np.mean's scalar branch
            assert _same(m.value, float(np.mean(x))) and _same(s.value, float(np.std(x))), (trial, d, dt)
            emul.emul_mean_std(DT[x.dtype], p, d, 0, C.byref(m), C.byref(s))            # a row of a 2-D array (axis=1)
            x2 = np.stack([x, x])
            assert _same(m.value, float(np.mean(x2, axis=1)[0])) and _same(s.value, float(np.std(x2, axis=1)[0])), (trial, d, dt)


def test_pairwise_plan_matches_numpy(emul):
    """The host-built pairwise PLAN of csrc/rowwise.cu (leaves + post-order addition program) evaluates to np.add.reduce
    bit for bit for every length, and gives up (no plan) only beyond kPlanLeaves leaves."""
    rng = np.random.default_rng(19)
    lengths = list(range(1, 300)) + [383, 384, 385, 511, 512, 513, 767, 768, 769, 1000, 1023, 1024, 1025, 1536, 2047, 2048, 3000, 4096,
                                      4097, 5000, 8191, 8192]
    planned = 0
    for i, d in enumerate(lengths):
        dt = DTS[i % 3]
        x = np.ascontiguousarray((rng.standard_normal(d) * float(rng.choice([0.01, 1.0, 40.0])) + float(rng.choice([0.0, 3.0]))).astype(dt))
        nl = C.c_int()
        with np.errstate(all="ignore"):
            got = emul.emul_pairwise_sum_plan(DT[x.dtype], C.c_void_p(x.ctypes.data), d, C.byref(nl))
            if nl.value < 0:
                assert d > 4096
                continue
            planned += 1
            assert nl.value >= 1 and _same(got, float(np.add.reduce(x))), (d, dt)
    assert planned >= len(lengths) - 3


def _prepared_query(q, metric):
    """What prep_query stores in qc (float64 carrier) for this metric."""
    with np.errstate(all="ignore"):
        if metric == "cosine_similarity":
            return P.unit_rows(q).astype(np.float64)
        if metric == "pearson_correlation":
            return (q - np.mean(q)).astype(np.float64)
    return q.astype(np.float64)


def _bits(x, words):
    b = np.zeros(words * 32, bool)
    b[: len(x)] = np.asarray(x) > 0
    return np.packbits(b, bitorder="little").view(np.uint32).copy()


@pytest.mark.parametrize("metric", sorted(METRIC))
def test_similarity_of_the_product_source(emul, metric):
    import zlib
    rng = np.random.default_rng(zlib.crc32(metric.encode()))
    fn = {"dot_product": P.dot_scores, "cosine_similarity": P.cosine_scores, "euclidean_metric": P.euclidean_scores,
          "manhattan_distance": P.manhattan_scores, "hamming_distance": P.hamming_scores, "jaccard_similarity": P.jaccard_scores,
          "pearson_correlation": P.pearson_scores}[metric]
    exact_cases = 0
    for trial in range(150):
        n, d = 6, int(rng.integers(1, 400))
        vdt, qdt = DTS[int(rng.integers(0, 3))], DTS[int(rng.integers(0, 3))]
        scale, shift = float(rng.choice([0.02, 1.0, 25.0])), float(rng.choice([0.0, 0.0, 2.0]))
        V = np.ascontiguousarray((rng.standard_normal((n, d)) * scale + shift).astype(vdt))
        q = np.ascontiguousarray((rng.standard_normal(d) * scale + shift).astype(qdt))
        V[0] = 0
        if d > 1:
            V[1] = V[1, 0]                                     # a constant row
        rdt = np.promote_types(vdt, qdt)
        with np.errstate(all="ignore"):
            want = np.asarray(fn(V.copy(), q.copy())).reshape(-1)
            qc = np.ascontiguousarray(_prepared_query(q, metric))
            qstd = float(np.std(q))
            words = ((d + 31) // 32 + 3) // 4 * 4
            qb = _bits(q, words)
            for i in range(n):
                row = np.ascontiguousarray(V[i])
                nrm, aux2 = 1.0, 0.0
                if metric == "cosine_similarity":
                    nrm = float(np.linalg.norm(row, axis=-1, keepdims=True)[0]) or 1.0
                elif metric == "pearson_correlation":
                    V2 = np.stack([row, row])
                    nrm, aux2 = float(np.mean(V2, axis=1)[0]), float(np.std(V2, axis=1)[0])
                rb = _bits(row, words)
                got = emul.emul_similarity(DT[np.dtype(rdt)], DT[row.dtype], METRIC[metric], C.c_void_p(row.ctypes.data),
                                           C.c_void_p(qc.ctypes.data), d, nrm, aux2, qstd, C.c_void_p(rb.ctypes.data),
                                           C.c_void_p(qb.ctypes.data), words)
                ref = float(want[i])
                if metric in ("dot_product", "cosine_similarity") and rdt != np.float16:
                    # OpenBLAS summation order is not NumPy's to define: the canonical value is the exact dot rounded once
                    tol = (1e-5 if rdt == np.float32 else 1e-12)
                    cond = 1.0 if metric == "cosine_similarity" else float(np.linalg.norm(row.astype(float)) * np.linalg.norm(q.astype(float)))
                    assert abs(got - ref) <= tol * max(abs(ref), cond, 1e-300), (trial, i, vdt, qdt, d)
                else:
                    assert _same(got, ref), (trial, i, metric, vdt.__name__, qdt.__name__, d, scale, shift, got, ref)
                    exact_cases += 1
    if metric not in ("dot_product", "cosine_similarity"):
        assert exact_cases == 150 * 6


# ---------------------------------------------------------------------------------------------------------------------
# The certificate (csrc/certificate.cuh::outsider_bound, compiled for the host) against the oracle: for every row, the
# reference's total score must not exceed the bound derived from the sweep's own float32 score and the ingest
# statistics.  The sweep is emulated in NumPy with the kernel's accumulate type (fp32, or fp64 for fp64 storage).
# ---------------------------------------------------------------------------------------------------------------------
def _sweep_scores(V, q, metric, acc):
    """float32 score part of the selection key, as sweep_kernel's epilogue forms it, and what prep_query derives from q."""
    from oracle import canonical as K
    with np.errstate(all="ignore"):
        if metric == "cosine_similarity":
            c = K.unit_rows(q).astype(np.float64)
        else:
            c = q.astype(np.float64)
        qa = c.astype(acc)
        qnorm = float(np.sqrt(np.sum(c * c)))
        Va = V.astype(acc)
        if metric in ("dot_product", "cosine_similarity"):
            total = (Va @ qa).astype(acc)
            if metric == "cosine_similarity":
                norms = K.row_norm(V).astype(np.float64)
                norms[norms == 0] = 1
                total = total * (acc(1) / norms.astype(acc))
            sim = total
        else:
            df = Va - qa[None, :]
            dist = np.sqrt(np.sum(df * df, axis=1, dtype=acc)) if metric == "euclidean_metric" else np.sum(np.abs(df), axis=1, dtype=acc)
            sim = acc(1) / (acc(1) + dist)
    return sim.astype(np.float64), qnorm


@pytest.mark.parametrize("metric", ["dot_product", "cosine_similarity", "euclidean_metric", "manhattan_distance"])
@pytest.mark.parametrize("decay", [False, True], ids=["plain", "decay"])
def test_certificate_bound_covers_the_reference(emul, metric, decay):
    import zlib
    from oracle import canonical as K
    rng = np.random.default_rng(zlib.crc32((metric + str(decay)).encode()))
    worst_use, rows_checked = 0.0, 0
    for trial in range(90):
        n, d = 64, int(rng.choice([3, 16, 96, 384, 768, 1536]))
        vdt, qdt = DTS[int(rng.integers(0, 3))], DTS[int(rng.integers(0, 3))]
        kind = trial % 3
        V = rng.standard_normal((n, d))
        q = rng.standard_normal(d)
        if kind == 0:                                    # the bench distribution: unit rows, unit query
            V /= np.linalg.norm(V, axis=1, keepdims=True)
            q /= np.linalg.norm(q)
        elif kind == 1:                                  # rows of very different lengths, near-duplicates of the query
            V *= rng.uniform(0.05, 4.0, (n, 1))
            V[:8] = q[None, :] * rng.uniform(0.5, 1.5, (8, 1)) + rng.standard_normal((8, d)) * 1e-3
        else:                                            # shifted data (large common component: cancellation in dot products)
            V += 1.5
            q += 1.5
        V, q = np.ascontiguousarray(V.astype(vdt)), np.ascontiguousarray(q.astype(qdt))
        rdt = np.promote_types(vdt, qdt)
        acc = np.float64 if vdt == np.float64 else np.float32
        with np.errstate(all="ignore"):
            canon = K.scores(V, q, metric).astype(np.float64)
            sim, qnorm = _sweep_scores(V, q, metric, acc)
            bias, dec = 0.0, np.zeros(n)
            if decay:
                bias = float(rng.choice([0.3, 1.0, -0.5]))
                dec = np.exp(-rng.uniform(0, 5, n))
            sweep = (sim + bias * dec).astype(np.float32).astype(np.float64)
            total = canon + bias * dec
            Vd = V.astype(np.float64)
            true_norm = np.linalg.norm(Vd, axis=1)
            cn = K.row_norm(V).astype(np.float64)
            cn[cn == 0] = 1
            max_norm = float(np.float32(min(true_norm.max() * (1 + 1e-6), 3e38)))
            max_ratio = float(np.float32(min((true_norm / cn).max() * (1 + 1e-6), 3e38)))
        for i in range(n):
            if not (np.isfinite(total[i]) and np.isfinite(sweep[i])):
                continue
            b = emul.emul_outsider_bound(float(sweep[i]), METRIC[metric], DT[np.dtype(rdt)], DT[V.dtype], d, max_norm, max_ratio,
                                         0.0, 0.0, 0.0, int(decay), bias, 0, qnorm, 1.0)
            assert total[i] <= b, (trial, i, metric, vdt.__name__, qdt.__name__, d, kind, total[i], sweep[i], b)
            if b > sweep[i] and np.isfinite(b):
                worst_use = max(worst_use, (total[i] - sweep[i]) / (b - sweep[i]))
            rows_checked += 1
    assert rows_checked > 3000
    assert worst_use < 0.95, worst_use


# ---------------------------------------------------------------------------------------------------------------------
# The same certificate for the BATCHED TENSOR PASS (csrc/batched_tc.cu): operands as tcgen05 reads them (fp16 exact; fp32 cut
# to tf32), the accumulator pushed DOWN by half of the band the certificate grants the tensor cores' non-IEEE fp32
# accumulation (d * 2^-20 of sum |v_j q_j|) -- the worst case for a row left outside the candidates --, then the epilogue's own
# float32 arithmetic (fma with the per-row factor and the decay addend; the norm expansion for euclidean).
# ---------------------------------------------------------------------------------------------------------------------
def _trunc_tf32(x):
    return (np.asarray(x, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _tensor_pass_scores(V, q, metric, addend32):
    from oracle import canonical as K
    d = V.shape[1]
    with np.errstate(all="ignore"):
        c = (np.asarray(K.unit_rows(q)).reshape(-1) if metric == "cosine_similarity" else q).astype(np.float64)      # prep_query's canonical values
        qnorm = float(np.sqrt(np.sum(c * c)))
        qa = c.astype(np.float32)                                                             # qb.qa
        if V.dtype == np.float16:
            Vop, bop = V.astype(np.float64), qa.astype(np.float16).astype(np.float64)          # queries_to_half_kernel: round to nearest
        else:
            Vop, bop = _trunc_tf32(V).astype(np.float64), _trunc_tf32(qa).astype(np.float64)
        acc = (Vop @ bop - d * 2.0 ** -20 * (np.abs(Vop) @ np.abs(bop))).astype(np.float32)
        Vd = V.astype(np.float64)
        if metric == "euclidean_metric":
            sqn = np.sum(Vd * Vd, axis=1).astype(np.float32)                                  # ingest column ||v||^2
            qsq = np.float32(qnorm * qnorm)
            t = (np.float32(2) * acc.astype(np.float64) - sqn.astype(np.float64)).astype(np.float32)   # fma(acc, 2, -|v|^2)
            d2 = (qsq - t).astype(np.float32)
            key = (np.float32(1) / (np.float32(1) + np.sqrt(np.maximum(d2, np.float32(0))))).astype(np.float32)
        else:
            inv = np.ones(len(V), np.float32)
            if metric == "cosine_similarity":
                cn = K.row_norm(V).astype(np.float32)
                cn[cn == 0] = 1
                inv = (np.float32(1) / cn).astype(np.float32)
            key = (acc.astype(np.float64) * inv.astype(np.float64) + addend32.astype(np.float64)).astype(np.float32)    # one fma
    return key.astype(np.float64), qnorm


@pytest.mark.parametrize("metric", ["dot_product", "cosine_similarity", "euclidean_metric"])
@pytest.mark.parametrize("decay", [False, True], ids=["plain", "decay"])
@pytest.mark.parametrize("mixed", [False, True], ids=["same-precision", "wider-query"])
def test_certificate_bound_of_the_batched_tensor_pass(emul, metric, decay, mixed):
    """mixed: the opt-in HDB_TC_MIXED=1 configuration -- a query wider than the store, whose B operand is the canonical query
    ROUNDED to the storage precision (float copy, then fp16 round-to-nearest or the tf32 cut)."""
    import zlib
    from oracle import canonical as K
    if metric == "euclidean_metric" and decay:
        pytest.skip("batched_tc_supported: euclidean with time decay stays on the sweeps")
    rng = np.random.default_rng(zlib.crc32(("tc" + metric + str(decay) + str(mixed)).encode()))
    worst_use, rows_checked = 0.0, 0
    for trial in range(80):
        n, d = 64, int(rng.choice([8, 16, 96, 384, 768, 1536]))
        vdt = DTS[trial % 2]
        if mixed:
            qdt = DTS[int(rng.integers(DT[np.dtype(vdt)] + 1, 3))]      # fp32 / fp64 over fp16, fp64 over fp32
        else:
            qdt = DTS[int(rng.integers(0, DT[np.dtype(vdt)] + 1))]      # batched_tc_supported: the query is never wider than the store
        kind = (trial // 2) % 3
        V = rng.standard_normal((n, d))
        q = rng.standard_normal(d)
        if kind == 0 or metric == "euclidean_metric":     # unit rows, unit query (the norm expansion only certifies embedding-like data)
            V /= np.linalg.norm(V, axis=1, keepdims=True)
            q /= np.linalg.norm(q)
            if kind == 1:
                V[:8] = q[None, :] + rng.standard_normal((8, d)) * 3e-2
        elif kind == 1:
            V *= rng.uniform(0.05, 4.0, (n, 1))
            V[:8] = q[None, :] * rng.uniform(0.5, 1.5, (8, 1)) + rng.standard_normal((8, d)) * 1e-3
        else:
            V += 1.5
            q += 1.5
        V, q = np.ascontiguousarray(V.astype(vdt)), np.ascontiguousarray(q.astype(qdt))
        rdt = np.promote_types(vdt, qdt)
        with np.errstate(all="ignore"):
            canon = K.scores(V, q, metric).astype(np.float64)
            bias, dec = 0.0, np.zeros(n)
            if decay:
                bias = float(rng.choice([0.3, 1.0, -0.5]))
                dec = np.exp(-rng.uniform(0, 5, n))
            key, qnorm = _tensor_pass_scores(V, q, metric, (bias * dec).astype(np.float32))
            total = canon + bias * dec
            Vd = V.astype(np.float64)
            true_norm = np.linalg.norm(Vd, axis=1)
            cn = K.row_norm(V).astype(np.float64)
            cn[cn == 0] = 1
            max_norm = float(np.float32(min(true_norm.max() * (1 + 1e-6), 3e38)))
            max_ratio = float(np.float32(min((true_norm / cn).max() * (1 + 1e-6), 3e38)))
        for i in range(n):
            if not (np.isfinite(total[i]) and np.isfinite(key[i])):
                continue
            b = emul.emul_outsider_bound(float(key[i]), METRIC[metric], DT[np.dtype(rdt)], DT[V.dtype], d, max_norm, max_ratio,
                                         0.0, 0.0, 0.0, int(decay), bias, 1, qnorm, 1.0)
            assert total[i] <= b, (trial, i, metric, vdt.__name__, qdt.__name__, d, kind, total[i], key[i], b)
            if b > key[i] and np.isfinite(b):
                worst_use = max(worst_use, (total[i] - key[i]) / (b - key[i]))
            rows_checked += 1
    assert rows_checked > 2500
    assert worst_use < 0.95, worst_use
