"""The select -> certify PROTOCOL of the batched tensor path (csrc/batched_tc.cu + csrc/finalize.cu) and, below, of the streaming sweep, run end to end on the CPU.

The per-row property "reference value <= bound of the row's own key" is checked in test_emul_canonical.py /
test_certificate_math.py.  This file checks what the product builds on it: for a whole matrix and a query,

  1. keys      = the select pass's float32 scores (operands as tcgen05 reads them, accumulator error drawn at random inside half
                 of the band the certificate grants, epilogue arithmetic in float32),
  2. tau0      = the k'-th best key of a strided row SAMPLE (sample_threshold_kernel; a query that ties on more than 4 k' sample
                 rows gets an unreachable threshold),
  3. candidates = rows with key >= tau0, the best k' of them by (key desc, row asc) re-scored with the oracle (what finalize's
                 canonical re-scoring is pinned to by the golden vectors),
  4. certified = k-th canonical total > outsider_bound(edge key)  -- the SHIPPED csrc/certificate.cuh compiled for the host; the
                 edge key is the k'-th candidate's, or tau0 when fewer than k' rows passed; pearson: divided by std_q,

and asserts the protocol's contract: WHENEVER a query is certified, its top-k (row ids with ties to the lower id, float64 scores)
equals the oracle's ranking of the WHOLE matrix.  Uncertified queries are the repair path's business (streaming sweep, exact path);
the test also checks that well-separated data does certify, so the contract is not vacuous.  NumPy only; nothing here runs on a GPU."""
import zlib

import numpy as np
import pytest

from oracle import canonical as K

DT = {np.dtype(np.float16): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}
METRIC = {"dot_product": 0, "cosine_similarity": 1, "euclidean_metric": 2, "pearson_correlation": 6}


def _trunc_tf32(x):
    return (np.asarray(x, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _select_pass_keys(V, q, metric, rng):
    """-> (float32 key score per row as float64, dict of what prep_query / the ingest pass hand to the certificate)."""
    n, d = V.shape
    info = {"qstd": 1.0, "qsumb": 0.0, "max_pratio": 0.0, "max_cratio": 0.0, "min_pstd": 0.0}
    Vd = V.astype(np.float64)
    with np.errstate(all="ignore"):
        if metric == "cosine_similarity":
            c = np.asarray(K.unit_rows(q)).reshape(-1).astype(np.float64)
        elif metric == "pearson_correlation":
            qmean = K.row_mean(q[None, :], scalar=True)[0]
            info["qstd"] = float(K.row_std(q[None, :])[0])
            c = (q - qmean).astype(np.float64)                                   # in the query's dtype, widened
        else:
            c = q.astype(np.float64)
        qa = c.astype(np.float32)
        if V.dtype == np.float16:
            Vop, bop = Vd, qa.astype(np.float16).astype(np.float64)
        else:
            Vop, bop = _trunc_tf32(V).astype(np.float64), _trunc_tf32(qa).astype(np.float64)
        band = d * 2.0 ** -20 * (np.abs(Vop) @ np.abs(bop))
        acc = (Vop @ bop + rng.uniform(-1.0, 1.0, n) * band).astype(np.float32)
        true_norm = np.linalg.norm(Vd, axis=1)
        info["max_norm"] = float(np.float32(true_norm.max() * (1 + 1e-6)))
        cn = K.row_norm(V).astype(np.float64)
        cn[cn == 0] = 1
        info["max_ratio"] = float(np.float32((true_norm / cn).max() * (1 + 1e-6)))
        info["qnorm"] = float(np.sqrt(np.sum(c * c)))
        if metric == "euclidean_metric":
            sqn = np.sum(Vd * Vd, axis=1).astype(np.float32)
            qsq = np.float32(info["qnorm"] ** 2)
            t = (2.0 * acc.astype(np.float64) - sqn.astype(np.float64)).astype(np.float32)      # the screened value (+ |q|^2 in tau)
            d2 = (qsq - t).astype(np.float32)
            key = (np.float32(1) / (np.float32(1) + np.sqrt(np.maximum(d2, np.float32(0))))).astype(np.float32)
            screen = (t - qsq).astype(np.float32)                                                # -d^2: what tau0 is selected on
        else:
            inv = np.ones(n, np.float32)
            if metric == "cosine_similarity":
                inv = (np.float32(1) / cn.astype(np.float32)).astype(np.float32)
            elif metric == "pearson_correlation":
                vmean, vstd = K.row_mean(V).astype(np.float64), K.row_std(V).astype(np.float64)
                inv = np.where(vstd == 0, np.nan, 1.0 / (vstd * d)).astype(np.float32)           # NaN for a constant row
                ok = vstd > 0
                info["max_pratio"] = float(np.float32(np.max(true_norm[ok] / (vstd[ok] * np.sqrt(d))) * (1 + 1e-6)))
                info["max_cratio"] = float(np.float32(np.max(np.linalg.norm(Vd[ok] - vmean[ok, None], axis=1) / (vstd[ok] * np.sqrt(d))) * (1 + 1e-6)))
                info["min_pstd"] = float(np.float32(np.min(vstd[ok]) * (1 - 1e-6)))
                info["qnorm"] = info["qnorm"] / (info["qstd"] * np.sqrt(d)) if info["qstd"] > 0 else float("nan")
                info["qsumb"] = float(np.sum(qa.astype(np.float64)))
            key = (acc.astype(np.float64) * inv.astype(np.float64)).astype(np.float32)
            screen = key
    return key.astype(np.float64), screen.astype(np.float64), info


def _certified_topk(emul, V, q, metric, k, kp):
    """One query through the protocol.  -> None (uncertified) or (row ids, float64 scores)."""
    n, d = V.shape
    rng = np.random.default_rng(zlib.crc32(q.tobytes()) & 0xFFFFFFFF)
    key, screen, info = _select_pass_keys(V, q, metric, rng)
    rdt = np.promote_types(V.dtype, q.dtype)
    # 2. threshold from a strided sample of 128-row tiles (every 8th tile, as ensure_tc_workspace sizes it)
    tiles = np.arange(0, n, 128)[::8]
    sample = np.concatenate([np.arange(t, min(t + 128, n)) for t in tiles])
    sv = screen[sample]
    sv = np.where(np.isnan(sv), -np.inf, sv)
    finite_sorted = np.sort(sv)[::-1]
    tau0 = finite_sorted[kp - 1] if len(finite_sorted) >= kp else -np.inf
    if tau0 > -np.inf and np.sum(sv == tau0) > 4 * kp:
        tau0 = np.inf
    # 3. select + the best k' by (key desc, row asc)
    cand = np.flatnonzero(screen >= tau0)                       # NaN compares false: a constant pearson row is never appended
    order = cand[np.lexsort((cand, -key[cand]))]
    top = order[:kp]
    m = len(top)
    if m < k:
        return None
    with np.errstate(all="ignore"):
        canon = K.total_scores(V[top], q, metric)
    rk = np.lexsort((top, -canon))[:k]
    res_rows, res_sc = top[rk], canon[rk]
    if n > m:
        if m == kp:
            s_edge = key[top[-1]]
        else:                                                   # fewer than k' rows passed: every outsider lies below tau0
            s_edge = 1.0 / (1.0 + np.sqrt(max(-tau0, 0.0))) if metric == "euclidean_metric" else tau0
        if metric == "pearson_correlation":
            s_edge = s_edge / info["qstd"] if info["qstd"] != 0 else float("nan")
        bound = emul.emul_outsider_bound_tc(float(s_edge), METRIC[metric], DT[np.dtype(rdt)], DT[V.dtype], d, info["max_norm"],
                                            info["max_ratio"], info["max_pratio"], info["max_cratio"], info["min_pstd"],
                                            info["qnorm"], info["qstd"], info["qsumb"])
        if not (res_sc[k - 1] > bound):
            return None
    return res_rows, res_sc


def _matrix(rng, kind, n, d, vdt):
    V = rng.standard_normal((n, d))
    if kind == "unit":
        V /= np.linalg.norm(V, axis=1, keepdims=True)
    elif kind == "clustered":                                   # groups of near-duplicates: thin margins, many key swaps
        centres = rng.standard_normal((n // 50 + 1, d))
        V = centres[rng.integers(0, len(centres), n)] + 0.02 * rng.standard_normal((n, d))
        V /= np.linalg.norm(V, axis=1, keepdims=True)
    elif kind == "shifted":
        V = V * rng.uniform(0.5, 1.5, (n, 1)) + rng.uniform(-0.5, 0.5, (n, 1))
    V = V.astype(vdt)
    V[n // 3] = V[7]                                            # an exact duplicate: the tie must go to the lower id
    return np.ascontiguousarray(V)


@pytest.mark.parametrize("metric", ["dot_product", "cosine_similarity", "euclidean_metric", "pearson_correlation"])
@pytest.mark.parametrize("vdt,qdt", [(np.float16, np.float16), (np.float32, np.float32), (np.float32, np.float16),
                                     (np.float16, np.float32), (np.float16, np.float64), (np.float32, np.float64)],
                         ids=["f16", "f32", "f32-store-f16-query", "mixed-f16-f32", "mixed-f16-f64", "mixed-f32-f64"])
def test_certified_answers_equal_the_oracle(emul, metric, vdt, qdt):
    """ids `mixed-*`: the opt-in HDB_TC_MIXED=1 configuration (query wider than the store, rounded B operand)."""
    rng = np.random.default_rng(zlib.crc32((metric + np.dtype(vdt).name + np.dtype(qdt).name).encode()))
    n, d, k = 9000, 64, 10
    certified = {"unit": 0, "clustered": 0, "shifted": 0}
    asked = dict(certified)
    kinds = ("unit", "clustered") if metric == "euclidean_metric" else ("unit", "clustered", "shifted")
    for kind in kinds:
        V = _matrix(rng, kind, n, d, vdt)
        kp = 128 if V.dtype == np.float32 else 32               # hdb_query: the tf32 select certifies the wide class
        queries = [rng.standard_normal(d) for _ in range(5)] + [V[7].astype(np.float64) + 0.01 * rng.standard_normal(d),
                                                                 V[n // 2].astype(np.float64)]
        if metric in ("euclidean_metric", "cosine_similarity") or kind != "shifted":
            queries = [x / np.linalg.norm(x) for x in queries]
        queries.append(np.zeros(d) if metric != "pearson_correlation" else np.full(d, 0.25))     # degenerate: ties on every row
        for x in queries:
            q = np.ascontiguousarray(x.astype(qdt))
            asked[kind] += 1
            got = _certified_topk(emul, V, q, metric, k, kp)
            if got is None:
                continue
            certified[kind] += 1
            with np.errstate(all="ignore"):
                want_rows, want_sc = K.rank(V, q, k, metric)
            assert list(got[0]) == list(want_rows), (kind, metric, got[0], want_rows)
            assert np.array_equal(got[1], want_sc), (kind, metric)
    # the contract is not vacuous: embedding-like data certifies (all but the degenerate query, which never may)
    assert certified["unit"] >= asked["unit"] - 2, (certified, asked)
    assert sum(certified.values()) < sum(asked.values())        # ... and the degenerate queries did go to the repair path


# ---------------------------------------------------------------------------------------------------------------------
# The same contract for the STREAMING SWEEP (csrc/sweep_float.cuh + finalize.cu), the path of the headline workload: the union of
# the per-CTA candidate lists always contains the k' best keys overall, so the protocol is: k' best rows by (float32 key, row) ->
# oracle re-scoring -> shipped certificate on the k'-th key.  Keys: the sweep's arithmetic in its accumulate type (NumPy's
# summation order differs from the kernel's FMA chains; both lie inside the D * u_acc band the certificate grants).
# ---------------------------------------------------------------------------------------------------------------------
SWEEP_METRIC = dict(METRIC, manhattan_distance=3)


@pytest.mark.parametrize("metric", ["dot_product", "cosine_similarity", "euclidean_metric", "manhattan_distance"])
@pytest.mark.parametrize("vdt,qdt", [(np.float16, np.float16), (np.float32, np.float32), (np.float64, np.float64),
                                     (np.float16, np.float64), (np.float32, np.float16)],
                         ids=["f16", "f32", "f64", "f16-store-f64-query", "f32-store-f16-query"])
@pytest.mark.parametrize("decay", [False, True], ids=["plain", "decay"])
def test_sweep_certified_answers_equal_the_oracle(emul, metric, vdt, qdt, decay):
    from test_emul_canonical import _sweep_scores
    rng = np.random.default_rng(zlib.crc32(("sweep" + metric + np.dtype(vdt).name + np.dtype(qdt).name + str(decay)).encode()))
    n, d, k = 6000, 48, 10
    certified = asked = 0
    for kind in ("unit", "clustered", "shifted"):
        V = _matrix(rng, kind, n, d, vdt)
        ts = 1.7e9 + rng.uniform(0, 5, n) if decay else None
        bias = 0.3 if decay else 0.0
        acc = np.float64 if V.dtype == np.float64 else np.float32
        Vd = V.astype(np.float64)
        true_norm = np.linalg.norm(Vd, axis=1)
        cn = K.row_norm(V).astype(np.float64)
        cn[cn == 0] = 1
        max_norm = float(np.float32(true_norm.max() * (1 + 1e-6)))
        max_ratio = float(np.float32((true_norm / cn).max() * (1 + 1e-6)))
        for kp in (32, 128):
            for x in [rng.standard_normal(d) for _ in range(3)] + [V[7].astype(np.float64) + 0.01 * rng.standard_normal(d)]:
                if kind != "shifted":
                    x = x / np.linalg.norm(x)
                q = np.ascontiguousarray(x.astype(qdt))
                rdt = np.promote_types(V.dtype, q.dtype)
                with np.errstate(all="ignore"):
                    sim, qnorm = _sweep_scores(V, q, metric, acc)
                    dec = np.exp(-np.max(ts) + ts) if decay else np.zeros(n)
                    key = (sim + bias * dec).astype(np.float32).astype(np.float64)
                rows = np.arange(n)
                top = rows[np.lexsort((rows, -key))][:kp]
                with np.errstate(all="ignore"):
                    canon = K.total_scores(V[top], q, metric) + (bias * dec[top] if decay else 0.0)
                rk = np.lexsort((top, -canon))[:k]
                asked += 1
                bound = emul.emul_outsider_bound(float(key[top[-1]]), SWEEP_METRIC[metric], DT[np.dtype(rdt)], DT[V.dtype], d, max_norm,
                                                 max_ratio, 0.0, 0.0, 0.0, int(decay), bias, 0, qnorm, 1.0)
                if not (canon[rk][k - 1] > bound):
                    continue
                certified += 1
                with np.errstate(all="ignore"):
                    want_rows, want_sc = K.rank(V, q, k, metric, ts, bias)
                assert list(top[rk]) == list(want_rows), (kind, kp, metric)
                assert np.array_equal(canon[rk], want_sc), (kind, kp, metric)
    assert certified >= asked // 3, (certified, asked)
