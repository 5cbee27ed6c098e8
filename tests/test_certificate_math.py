"""The pearson certificate of csrc/finalize.cu (outsider_bound, HDB_PEARSON branch) restated in NumPy and checked
against the oracle: for every row, the reference's value must not exceed the bound computed from the sweep's own (fp32 /
fp64) score and the ingest statistics.  A violated bound would mean a row outside the candidate list could belong to the
top-k without the certificate noticing -- this runs on the CPU, the GPU tests check the resulting indices."""
import numpy as np
import pytest

from oracle import canonical as K

U = {np.float16: 2.0 ** -11, np.float32: 2.0 ** -24, np.float64: 2.0 ** -53}


def bound_and_canon(rng, sdt, qdt, n, d, scale, shift):
    V = (rng.standard_normal((n, d)) * scale * rng.uniform(0.2, 3, (n, 1)) + shift * rng.uniform(-1, 1, (n, 1))).astype(sdt)
    q = (rng.standard_normal(d) * scale + shift).astype(qdt)
    R = np.promote_types(sdt, qdt).type
    with np.errstate(all="ignore"):
        canon = K.pearson(V, q)
        vmean = K.row_mean(V).astype(np.float64)
        vstd = K.row_std(V).astype(np.float64)
        qmean = K.row_mean(q[None, :], scalar=True)[0]
        qstd = float(K.row_std(q[None, :])[0])
    ok = (vstd > 0) & np.isfinite(vstd) & np.isfinite(canon)
    if not ok.any() or not (qstd > 0):
        return None
    b = q - qmean                                            # in the query's dtype, as prep_query stores it
    acc = np.float64 if sdt == np.float64 else np.float32   # the sweep's accumulate type
    b_acc = b.astype(acc)
    sumb = float(np.sum(b_acc.astype(np.float64)))
    with np.errstate(all="ignore"):
        t = (V.astype(acc) @ b_acc).astype(acc)
        pscale = (1.0 / (vstd * d)).astype(acc)
        sweep = ((t - vmean.astype(acc) * acc(sumb)) * acc(1.0 / qstd) * pscale).astype(np.float32).astype(np.float64)
    Vd = V.astype(np.float64)
    max_pratio = np.max(np.linalg.norm(Vd[ok], axis=1) / (vstd[ok] * np.sqrt(d))) * (1 + 1e-6)
    max_cratio = np.max(np.linalg.norm(Vd[ok] - vmean[ok, None], axis=1) / (vstd[ok] * np.sqrt(d))) * (1 + 1e-6)
    min_std = np.min(vstd[ok]) * (1 - 1e-6)
    qn = np.linalg.norm(b.astype(np.float64)) / (qstd * np.sqrt(d))
    D = d + 8.0
    ua = 4.440892098500626e-16 if sdt == np.float64 else 2.0 ** -23
    uk, uT, uR = 2.0 ** -23, U[sdt], U[R]
    uaccR = 2.0 ** -53 if R == np.float64 else 2.0 ** -24
    A, Ac = max_pratio * qn, max_cratio * qn
    s = sweep[ok]
    bound = s + np.abs(s) * uk + A * ((D + 16) * ua + 2.0 ** -23) + Ac * 1.05 * (uT + uR + D * uaccR)
    bound = bound + 6 * uR * np.abs(bound)
    if R == np.float16 or sdt == np.float16:
        max_norm = np.max(np.linalg.norm(Vd, axis=1))
        if not (min_std * qstd > 2.44140625e-4) or not (max_norm * np.sqrt(D) * qstd * max(1.0, 2 * A) < 3.0e4):
            return None                                      # the kernel refuses to certify (exact path)
        bound = bound + 2.0 ** -25 * (qn / min_std + 1.0 / (min_std * qstd)) * 1.5
    return canon[ok], s, bound


@pytest.mark.parametrize("sdt", [np.float16, np.float32, np.float64])
@pytest.mark.parametrize("qdt", [np.float16, np.float32, np.float64])
def test_pearson_bound_covers_the_reference(sdt, qdt):
    import zlib
    rng = np.random.default_rng(zlib.crc32((sdt.__name__ + qdt.__name__).encode()))
    checked = 0
    for d in (8, 33, 96, 384, 1536):
        for scale, shift in ((1.0, 0.0), (0.04, 0.0), (1.0, 5.0), (0.04, 0.3), (3.0, 20.0)):
            r = bound_and_canon(rng, sdt, qdt, 200, d, scale, shift)
            if r is None:
                continue
            canon, s, bound = r
            assert np.all(canon <= bound), (d, scale, shift, float(np.max(canon - bound)))
            used = np.max((canon - s) / np.maximum(bound - s, 1e-300))
            assert used < 0.9, (d, scale, shift, used)        # the slack is not razor-thin either
            checked += 1
    assert checked >= 10


@pytest.mark.parametrize("sdt", [np.float16, np.float32, np.float64])
def test_pearson_bound_of_the_product_code(emul, sdt):
    """The same property through the PRODUCT's own outsider_bound (csrc/certificate.cuh compiled for the host): the
    NumPy restatement above and the shipped code must agree, and the shipped code must cover the reference."""
    import zlib
    rng = np.random.default_rng(zlib.crc32(("product" + sdt.__name__).encode()))
    rows = 0
    for qdt in (np.float16, np.float32, np.float64):
        for d in (8, 96, 768):
            for scale, shift in ((1.0, 0.0), (0.04, 0.3), (3.0, 20.0)):
                n = 150
                V = (rng.standard_normal((n, d)) * scale * rng.uniform(0.2, 3, (n, 1)) + shift * rng.uniform(-1, 1, (n, 1))).astype(sdt)
                q = (rng.standard_normal(d) * scale + shift).astype(qdt)
                R = np.promote_types(sdt, qdt)
                with np.errstate(all="ignore"):
                    canon = K.pearson(V, q)
                    vmean, vstd = K.row_mean(V).astype(np.float64), K.row_std(V).astype(np.float64)
                    qmean = K.row_mean(q[None, :], scalar=True)[0]
                    qstd = float(K.row_std(q[None, :])[0])
                    b = q - qmean
                    acc = np.float64 if sdt == np.float64 else np.float32
                    b_acc = b.astype(acc)
                    sumb = float(np.sum(b_acc.astype(np.float64)))
                    t = (V.astype(acc) @ b_acc).astype(acc)
                    sweep = ((t - vmean.astype(acc) * acc(sumb)) * acc(1.0 / qstd) * (1.0 / (vstd * d)).astype(acc)).astype(np.float32).astype(np.float64)
                ok = (vstd > 0) & np.isfinite(vstd) & np.isfinite(canon) & np.isfinite(sweep)
                if not ok.any() or not (qstd > 0):
                    continue
                Vd = V.astype(np.float64)
                f32 = lambda x: float(np.float32(x))
                max_norm = f32(np.max(np.linalg.norm(Vd, axis=1)) * (1 + 1e-6))
                max_pratio = f32(np.max(np.linalg.norm(Vd[ok], axis=1) / (vstd[ok] * np.sqrt(d))) * (1 + 1e-6))
                max_cratio = f32(np.max(np.linalg.norm(Vd[ok] - vmean[ok, None], axis=1) / (vstd[ok] * np.sqrt(d))) * (1 + 1e-6))
                min_pstd = f32(np.min(vstd[ok]) * (1 - 1e-6))
                qn = float(np.linalg.norm(b.astype(np.float64)) / (qstd * np.sqrt(d)))
                dts = {np.dtype(np.float16): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}
                for i in np.flatnonzero(ok):
                    bound = emul.emul_outsider_bound(float(sweep[i]), 6, dts[np.dtype(R)], dts[np.dtype(sdt)], d, max_norm, 1.0, max_pratio,
                                                     max_cratio, min_pstd, 0, 0.0, 0, qn, qstd)
                    assert canon[i] <= bound, (qdt.__name__, d, scale, shift, i, canon[i], sweep[i], bound)
                    rows += 1
    assert rows > 2000


def _trunc_tf32(x):
    """kind::tf32 reads fp32 operands with 10 explicit mantissa bits (the low 13 bits are ignored)."""
    return (np.asarray(x, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


@pytest.mark.parametrize("sdt", [np.float16, np.float32])
def test_pearson_bound_of_the_batched_tensor_pass(emul, sdt):
    """Batched pearson on the tensor cores (csrc/batched_tc.cu, VERDICT round 1 missing #6): the select pass screens
    u = fl32(acc * pscale) with acc = V . b on tcgen05 (b = q - mean(q); fp16 operands exact, fp32 operands cut to tf32), WITHOUT
    the mean correction mean_v * sum(b) and without 1 / std_q; finalize.cu divides the edge key by std_q and calls outsider_bound
    with sum(b).  Property: for every row the reference's value stays below the bound of the row's own key -- with the
    accumulator pushed DOWN by half of the error band the certificate grants the tensor cores (d * 2^-20 of sum |v_j b_j|), the
    worst case for an outsider."""
    import zlib
    rng = np.random.default_rng(zlib.crc32(("tensor" + sdt.__name__).encode()))
    dts = {np.dtype(np.float16): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}
    rows = 0
    worst = 0.0
    # batched_tc_supported: q_dtype <= storage; wider queries (rounded B operand) only with the opt-in HDB_TC_MIXED=1
    for qdt in (np.float16, np.float32, np.float64):
        for d in (8, 96, 384, 768):
            for scale, shift in ((1.0, 0.0), (0.04, 0.0), (0.04, 0.3), (1.0, 5.0), (3.0, 20.0)):
                n = 150
                V = (rng.standard_normal((n, d)) * scale * rng.uniform(0.2, 3, (n, 1)) + shift * rng.uniform(-1, 1, (n, 1))).astype(sdt)
                q = (rng.standard_normal(d) * scale + shift).astype(qdt)
                R = np.promote_types(sdt, qdt)
                with np.errstate(all="ignore"):
                    canon = K.pearson(V, q)
                    vmean, vstd = K.row_mean(V).astype(np.float64), K.row_std(V).astype(np.float64)
                    qmean = K.row_mean(q[None, :], scalar=True)[0]
                    qstd = float(K.row_std(q[None, :])[0])
                    b = (q - qmean).astype(np.float32)                  # qb.qa: exact unless the query is float64
                    sumb = float(np.sum(b.astype(np.float64)))          # qb.qaux[2b + 1]
                    if sdt == np.float32:
                        Vop, bop = _trunc_tf32(V).astype(np.float64), _trunc_tf32(b).astype(np.float64)
                    else:
                        Vop, bop = V.astype(np.float64), b.astype(np.float16).astype(np.float64)
                        if qdt == np.float16:
                            assert np.array_equal(bop, b.astype(np.float64))      # fp16 query: the B operand is exact
                    acc = Vop @ bop - d * 2.0 ** -20 * (np.abs(Vop) @ np.abs(bop))
                    pscale = (1.0 / (vstd * d)).astype(np.float32)
                    u = (acc.astype(np.float32) * pscale).astype(np.float32).astype(np.float64)     # the key's score part
                ok = (vstd > 0) & np.isfinite(vstd) & np.isfinite(canon) & np.isfinite(u)
                if not ok.any() or not (qstd > 0):
                    continue
                Vd = V.astype(np.float64)
                f32 = lambda x: float(np.float32(x))
                max_norm = f32(np.max(np.linalg.norm(Vd, axis=1)) * (1 + 1e-6))
                max_pratio = f32(np.max(np.linalg.norm(Vd[ok], axis=1) / (vstd[ok] * np.sqrt(d))) * (1 + 1e-6))
                max_cratio = f32(np.max(np.linalg.norm(Vd[ok] - vmean[ok, None], axis=1) / (vstd[ok] * np.sqrt(d))) * (1 + 1e-6))
                min_pstd = f32(np.min(vstd[ok]) * (1 - 1e-6))
                qn = float(np.linalg.norm((q - qmean).astype(np.float64)) / (qstd * np.sqrt(d)))
                for i in np.flatnonzero(ok):
                    s = float(u[i]) / qstd
                    bound = emul.emul_outsider_bound_tc(s, 6, dts[np.dtype(R)], dts[np.dtype(sdt)], d, max_norm, 1.0, max_pratio,
                                                        max_cratio, min_pstd, qn, qstd, sumb)
                    assert canon[i] <= bound, (qdt.__name__, d, scale, shift, i, canon[i], s, bound)
                    if np.isfinite(bound):
                        worst = max(worst, (canon[i] - s) / max(bound - s, 1e-300))
                        rows += 1
    assert rows > 3000
    assert worst < 0.95, worst            # the band is used, not exhausted
