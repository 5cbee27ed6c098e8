// The certificate of the fused path: an upper bound of the CANONICAL total score of any
// row that is NOT among the candidates, given the fp32 score part of the weakest candidate key.  Pure arithmetic on the
// ingest statistics -- kept in a header so that the host-side tests (tests/emul/) can compile exactly this code and check
// it against NumPy without a GPU.
#pragma once
#include "hdb_common.cuh"
#include "hdb_internal.h"
#include "../../include/hyperdb_b200.h"

namespace hdb {

// Upper bound of the CANONICAL total score of any row whose selection key is <= the KP-th key
// (score part s).  See DESIGN.md "certificate"; every term is a worst-case rounding bound.
// qsumb: pearson on the batched tensor pass only -- sum_j (q_j - mean_q), the factor of the mean correction that pass leaves out.
__device__ inline double outsider_bound(double s, const FinalizeArgs& a, double qnorm, double qstd, double qsumb = 0.0) {
  const double D = (double)a.m.d + 8.0;
  const double uk = 1.1920928955078125e-7;                    // 2^-23: float32 key rounding (2x slack)
  // accumulate type of the select pass; the tensor cores' fp32 accumulation is not IEEE round-to-nearest
  const double ua = a.cand_count ? 1.9073486328125e-6 /*2^-19*/ : (a.m.dtype == 2) ? 4.440892098500626e-16 : 1.1920928955078125e-7;
  const double uR = unit_roundoff(a.rdt);
  const double uT = unit_roundoff(a.m.dtype);
  const bool decay = a.f.decay != nullptr;
  const double chain16 = (a.rdt == 0) ? D * 1.1920928955078125e-7 : 0.0;   // HALF_dot / f16 pairwise run in float32
  if (a.metric == HDB_HAMMING || a.metric == HDB_JACCARD) return s + fabs(s) * uk;     // exact on both sides: only the key rounding
  // Batched pass with a query WIDER than the stored matrix (HDB_TC_MIXED=1): the B operand of the contraction is the canonical query
  // rounded to the storage precision.  fp16 store: round to nearest (queries_to_half_kernel) -- 2^-11 relative (2^-24 more for a
  // float64 query's float copy), or 2^-25 absolute below the normal range, and no overflow while |q_j| <= ||q|| < 65504 (checked per
  // metric below).  fp32 store: only the float copy of a float64 query is new; the tf32 cut is in the 2 * 2^-9 band already.
  const bool qcut = a.cand_count && a.rdt > a.m.dtype;
  const double qrel = !qcut ? 0.0 : (a.m.dtype == 0 ? 4.8835e-4 : 1.1920928955078125e-7);
  const double qabs = (qcut && a.m.dtype == 0) ? 2.98023223876953125e-8 : 0.0;          // per unit of sum_j |v_j| * (row factor)
  if (a.metric == HDB_PEARSON) {
    // Both sides evaluate rho~ = sum_j (v_j - mean_v) b_j / (std_v std_q d) from the SAME rounded statistics (b = q - mean_q).
    // Magnitudes: sum_j |v_j b_j| / (std_v std_q d) <= A := max_i ||v_i|| / (std_i sqrt d) * ||b|| / (std_q sqrt d)   (qnorm slot),
    // and ||v - mean_v|| <= 2 ||v||, so |rho~| <= 2A.
    //             sum_j |(v_j - mean_v) b_j| / (std_v std_q d) <= Ac := max_i ||v_i - mean_i|| / (std_i sqrt d) * (same query factor), ~1.
    const double A = (double)a.m.max_pratio * qnorm, Ac = (double)a.m.max_cratio * qnorm;
    const double uaccR = (a.rdt == 2) ? 1.1102230246251565e-16 : 5.9604644775390625e-8;
    double b = s + fabs(s) * uk
             + A * ((D + 16.0) * ua + 1.1920928955078125e-7)   // sweep: FMA chains, mean correction, scalings; query rounded to the accumulate type
             + Ac * 1.05 * (uT + uR + D * uaccR);              // reference: (v - mean) rounded to S, products to R, pairwise sum
    if (a.cand_count) {
      // batched pass (batched_tc.cu): v.b / (std_v d) on the tensor cores, divided by std_q by the caller.  It leaves out the mean
      // correction mean_v * sum(b) / (std_v std_q d), at most max_pratio * |sum(b)| / (d std_q) because |mean_v| <= ||v|| / sqrt(d);
      // kind::tf32 keeps 10 mantissa bits of each fp32 operand (2 * 2^-9 of sum_j |v_j b_j|, as for dot / cosine).
      b += (double)a.m.max_pratio * fabs(qsumb) / ((double)a.m.d * qstd) * 1.000001;
      if (a.m.dtype == 1) b += A * 3.90625e-3;
      if (qcut) {
        // ||b|| = qnorm * std_q * sqrt(d) bounds every |b_j|;  sum_j |v_j| / (std_v d std_q) <= max_pratio / std_q
        if (a.m.dtype == 0 && !(qnorm * qstd * sqrt(D) < 6.0e4)) return INFINITY;
        b += A * qrel + qabs * (double)a.m.max_pratio / qstd;
      }
    }
    // reference: rounding of the sum, the denominator (3 roundings) and the quotient, relative to the similarity itself
    b += 6.0 * uR * (decay ? Ac : fabs(b));
    if (a.rdt == 0 || a.m.dtype == 0) {
      // float16: differences / products are multiples of 2^-24 (absolute errors), and the denominator must stay normal and finite
      const double smin = (double)a.m.min_pstd;
      if (!(smin * qstd > 2.44140625e-4) || !((double)a.m.max_norm * sqrt(D) * qstd * fmax(1.0, 2.0 * A) < 3.0e4)) return INFINITY;
      b += 2.98023223876953125e-8 * (qnorm / smin + 1.0 / (smin * qstd)) * 1.5;
    }
    return decay ? b + fabs(b) * 1e-15 : b;
  }
  if (a.metric == HDB_DOT || a.metric == HDB_COSINE) {
    const double A = (double)(a.metric == HDB_DOT ? a.m.max_norm : a.m.max_ratio) * qnorm;   // >= sum |v_i q_i|
    double e = D * ua + chain16;
    if (a.cand_count && a.m.dtype == 1) e += 3.90625e-3;       // kind::tf32 keeps 10 mantissa bits of each fp32 operand (2 * 2^-9)
    if (a.metric == HDB_COSINE) e += 3.0 * uT + (a.m.dtype == 0 ? sqrt(D) * 5.9604644775390625e-8 : 0.0);
    double b = s + fabs(s) * uk + A * e;
    if (qcut) {
      // sum_j |v_j| * (row factor) <= sqrt(d) * max_norm (dot) or sqrt(d) * max ||v|| / norm (cosine: the unit query cannot overflow)
      if (a.m.dtype == 0 && a.metric == HDB_DOT && !(qnorm < 6.0e4)) return INFINITY;
      b += A * qrel + qabs * sqrt(D) * (double)(a.metric == HDB_DOT ? a.m.max_norm : a.m.max_ratio);
    }
    return decay ? b + 2.0 * uR * A + fabs(b) * 1e-15 : b + 2.0 * uR * fabs(b);
  }
  // euclidean / manhattan: similarity in (0, 1]
  if (decay) {
    // total = sim + bias*decay.  With bias >= 0 the outsider's similarity is at most min(1, s'), and the
    // similarity's error is relative to it (d(sim) <= d(dist)/(1+dist)^2 <= rel(dist) * sim).
    const double smax = (a.f.bias >= 0.0) ? fmin(1.0, fmax(s, 0.0) * (1.0 + uk) + 1e-300) : 1.0;
    double rel = (a.metric == HDB_EUCLIDEAN) ? 0.5 * (D * ua + 6.0 * uR + chain16) + 6.0 * uR
                                             : (D * ua + 4.0 * uR + chain16) + 3.0 * uR;
    double e = rel * smax + ua * qnorm;
    if (a.rdt == 0 && a.metric == HDB_EUCLIDEAN) {
      const double under = D * 5.9604644775390625e-8, dmin = 1.0 / smax - 1.0;
      e += (dmin > 1e-2) ? under / (2.0 * dmin) : sqrt(under);
    }
    return s + fabs(s) * uk + e;
  }
  if (!(s > 0.0)) return INFINITY;
  const double s_hi = s * (1.0 + uk);
  double d_lo = 1.0 / s_hi - 1.0 - 4.76837158203125e-7 - ua * qnorm;          // distance of the fast pass, lower bound
  if (d_lo < 0.0) d_lo = 0.0;
  double dc;
  if (a.metric == HDB_EUCLIDEAN) {
    double d2 = d_lo * d_lo * (1.0 - D * ua - 6.0 * uR - chain16);
    if (a.cand_count) {
      // batched pass: d^2 = |v|^2 + |q|^2 - 2 v.q on the tensor cores -> ABSOLUTE error from the cancellation
      const double A = (double)a.m.max_norm * qnorm;
      d2 -= 2.0 * A * (D * ua + (a.m.dtype == 1 ? 3.90625e-3 : 0.0)) + 8.0 * 1.1920928955078125e-7 * ((double)a.m.max_norm * a.m.max_norm + qnorm * qnorm);
      if (qcut) {
        // |q|^2 comes from the exact query, 2 v.q from the rounded one: 2 |v.(q^ - q)|
        if (a.m.dtype == 0 && !(qnorm < 6.0e4)) return INFINITY;
        d2 -= 2.0 * (A * qrel + qabs * sqrt(D) * (double)a.m.max_norm);
      }
    }
    if (a.rdt == 0) d2 -= D * 5.9604644775390625e-8;                          // float16 squares flushed below 2^-24
    dc = sqrt(d2 > 0.0 ? d2 : 0.0) * (1.0 - 3.0 * uR);
  } else {
    dc = d_lo * (1.0 - D * ua - 4.0 * uR - chain16);
    if (dc < 0.0) dc = 0.0;
  }
  return (1.0 / (1.0 + dc)) * (1.0 + 3.0 * uR);
}

}  // namespace hdb
