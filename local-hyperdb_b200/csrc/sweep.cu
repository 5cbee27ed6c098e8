// Fused score + select pass over one row shard (1, 2, 4 or 8 queries per launch): the replacement of
//   metric_func(vectors, q)            hyperdb/ranking_algorithm.py:168 (np.dot / norm / sum over N x D)
//   + recency_bias*exp(ts - max ts)    :179-186
//   argpartition / argsort             :199-200
// in ONE streaming read of the matrix.  HBM-bound: every row is read exactly once with 128-bit
// coalesced ld.global.nc loads (8 rows in flight per warp), the query lives in shared memory in the
// accumulate type, per-row scores are reduced with warp shuffles, and the epilogue (cosine norm,
// time decay, row mask/range, NaN -> -inf) feeds a per-warp candidate list guarded by a grid-wide
// threshold.  The N-length score vector is never written.  Output: per-CTA top-KP selection keys;
// csrc/finalize.cu merges them, re-scores the KP candidates in the reference's exact arithmetic and
// certifies the top-k (DESIGN.md "select then certify").
#include <cstdlib>

#include "sweep_common.cuh"
#include "sweep_hamming.h"

namespace hdb {

constexpr int kStagedDefaultLimit = 1024;

// the 18 (storage type, metric class, candidate class) instantiations of csrc/sweep_inst.cu
#define HDB_DECL_SWEEP(name) \
  int name##_kp32(const SweepParams& p, bool vec, int nq, int grid, size_t smem, cudaStream_t s); \
  int name##_kp128(const SweepParams& p, bool vec, int nq, int grid, size_t smem, cudaStream_t s)
HDB_DECL_SWEEP(sweep_f16_mc0); HDB_DECL_SWEEP(sweep_f16_mc1); HDB_DECL_SWEEP(sweep_f16_mc2);
HDB_DECL_SWEEP(sweep_f32_mc0); HDB_DECL_SWEEP(sweep_f32_mc1); HDB_DECL_SWEEP(sweep_f32_mc2);
HDB_DECL_SWEEP(sweep_f64_mc0); HDB_DECL_SWEEP(sweep_f64_mc1); HDB_DECL_SWEEP(sweep_f64_mc2);
#undef HDB_DECL_SWEEP
// the six instantiations of csrc/sweep_staged_inst.cu (TMA-staged row tiles, one query per pass)
#define HDB_DECL_STAGED(name) int name(const SweepParams& p, int kp, int grid, cudaStream_t s)
HDB_DECL_STAGED(staged_f16_mc0); HDB_DECL_STAGED(staged_f16_mc1); HDB_DECL_STAGED(staged_f16_mc2);
HDB_DECL_STAGED(staged_f32_mc0); HDB_DECL_STAGED(staged_f32_mc1); HDB_DECL_STAGED(staged_f32_mc2);
#undef HDB_DECL_STAGED

// Which single-query passes take the TMA-staged kernel: rows of at most `limit` bytes whose subset is tile-dense.
// HDB_SWEEP_STAGED (A/B switch): 0 = never, N > 0 = rows of at most N bytes.  Default: see DESIGN.md section 4.1b.
static int64_t staged_row_limit() {
  static const int64_t limit = [] {
    const char* e = getenv("HDB_SWEEP_STAGED");
    return e ? (int64_t)atoll(e) : (int64_t)kStagedDefaultLimit;
  }();
  return limit;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int sweep_grid_size(int device) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return sms * 2;                       // __launch_bounds__(256, 2): two CTAs resident per SM
}

// shared memory of a float sweep: NQ candidate lists per warp, NQ thresholds, the query tile
static size_t float_sweep_smem(const MatrixView& m, int kp, int nq, bool vec) {
  const size_t acc = m.dtype == 2 ? 8 : 4;
  const size_t lists = (size_t)kSweepWarps * nq * (kp <= 32 ? 128 : 2 * kp) * 8 + (size_t)((nq + 1) & ~1) * 8;
  size_t cols = (size_t)m.d;
  if (vec) {
    const size_t per = 16 / dtype_size(m.dtype);
    const size_t nvec = (size_t)m.d / per;
    cols = ((nvec + 31) / 32) * 32 * per;                  // padded to whole warp steps
  }
  return lists + cols * nq * acc;
}

static bool float_rows_vectorisable(const MatrixView& m) {
  return ((m.d * dtype_size(m.dtype)) % 16 == 0) && ((reinterpret_cast<uintptr_t>(m.rows) & 15) == 0);
}

// Largest number of queries one pass can take for this shape (1, 2, 4 or 8): bounded by shared memory (two CTAs per
// SM), by the accumulator registers (fp64: 4) and by the wide candidate class (k' = 128: 4).
int sweep_max_group(const MatrixView& m, int metric, int kp) {
  if (metric == HDB_HAMMING || metric == HDB_JACCARD) return hamming_sweep_max_group(m, kp);
  const bool vec = float_rows_vectorisable(m);
  int best = 1;
  for (int nq = 2; nq <= 8; nq *= 2) {
    if (nq == 8 && (kp > 32 || m.dtype == 2)) break;
    if (float_sweep_smem(m, kp, nq, vec) > 100 * 1024) break;
    best = nq;
  }
  return best;
}

int launch_sweep(const MatrixView& m, int metric, int rdt, const void* qa, const uint32_t* qbits, const double* qaux, const RowFilter& f,
                 int kp, const SweepOut& out, int nq, cudaStream_t s) {
  if (nq != 1 && nq != 2 && nq != 4 && nq != 8) return fail("sweep: a pass takes 1, 2, 4 or 8 queries");
  if (nq > sweep_max_group(m, metric, kp)) return fail("sweep: query group too large for this shape");
  if (m.n >= (int64_t(1) << 32)) return fail("sweep: more than 2^32 rows per shard");
  if (kp != 32 && kp != 128) return fail("sweep: unsupported candidate class");
  if (metric == HDB_HAMMING || metric == HDB_JACCARD) return launch_hamming_sweep(m, metric, qbits, f, kp, out, nq, s);
  SweepParams p;
  p.rows = reinterpret_cast<const char*>(m.rows);
  p.n = m.n; p.d = m.d;
  p.row_bytes = m.d * dtype_size(m.dtype);
  const bool vec = (p.row_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(m.rows) & 15) == 0);
  p.nvec = vec ? (int)(p.row_bytes / 16) : (int)m.d;
  p.qa = qa;
  p.inv_norms = (metric == HDB_COSINE) ? m.inv_norms : (metric == HDB_PEARSON ? m.pscale : nullptr);
  p.row_means = (metric == HDB_PEARSON) ? m.pmean : nullptr;
  p.qaux = (metric == HDB_PEARSON) ? qaux : nullptr;
  if (metric == HDB_PEARSON && (!m.pscale || !m.pmean || !qaux)) return fail("sweep: pearson columns missing");
  p.f = f; p.cand = out.cand; p.tau = out.tau; p.metric = metric;
  // result dtype float16 <=> rows and queries are float16: the prepared query values are float16 numbers
  static const bool no_fhfma = getenv("HDB_NO_FHFMA") != nullptr;          // A/B switch
  p.q_half = (m.dtype == 0 && rdt == 0 && !no_fhfma) ? 1 : 0;
  p.cand_stride = (int64_t)out.grid * kp;
  const int mc = (metric == HDB_DOT || metric == HDB_COSINE || metric == HDB_PEARSON) ? 0 : (metric == HDB_EUCLIDEAN ? 1 : 2);
  const size_t smem = float_sweep_smem(m, kp, nq, vec);
  if (smem > 200 * 1024) return fail("sweep: dimension too large for the fused pass");
  if (nq == 1 && m.dtype != 2 && vec && f.tile_dense && p.row_bytes <= staged_row_limit() && p.row_bytes <= 4096 &&
      (reinterpret_cast<uintptr_t>(m.rows) & 15) == 0 && m.n >= 4096) {
    typedef int (*SFn)(const SweepParams&, int, int, cudaStream_t);
    static const SFn staged[2][3] = {{staged_f16_mc0, staged_f16_mc1, staged_f16_mc2}, {staged_f32_mc0, staged_f32_mc1, staged_f32_mc2}};
    return staged[m.dtype][mc](p, kp, out.grid, s);
  }
  typedef int (*Fn)(const SweepParams&, bool, int, int, size_t, cudaStream_t);
#define HDB_PAIR(name) {name##_kp32, name##_kp128}
  static const Fn table[3][3][2] = {{HDB_PAIR(sweep_f16_mc0), HDB_PAIR(sweep_f16_mc1), HDB_PAIR(sweep_f16_mc2)},
                                    {HDB_PAIR(sweep_f32_mc0), HDB_PAIR(sweep_f32_mc1), HDB_PAIR(sweep_f32_mc2)},
                                    {HDB_PAIR(sweep_f64_mc0), HDB_PAIR(sweep_f64_mc1), HDB_PAIR(sweep_f64_mc2)}};
#undef HDB_PAIR
  return table[m.dtype][mc][kp <= 32 ? 0 : 1](p, vec, nq, out.grid, smem, s);
}

}  // namespace hdb
