// Row-wise passes at the HBM rate: everything that must touch EVERY stored row in the reference's exact arithmetic.
//
//   row_stats        get_norm_vector's norm per row (hyperdb/ranking_algorithm.py:8-21), NaN scan (:150), certificate statistics
//   pearson_stats    np.mean / np.std per row (:91,:94)
//   normalize_rows   get_norm_vector itself
//   full scores      the metric functions themselves (:24-61, :78-113; hdb_scores) and the exact path behind the fused select
//
// Round 1 ran these one THREAD per row (every load instruction of a warp touched 32 different rows: 2-4 % of the HBM
// rate).  Here a row is read once with coalesced 128-bit loads:
//   * order-dependent PAIRWISE sums (norm, mean, std, euclidean, manhattan, pearson): one WARP per row, the row's
//     terms staged in shared memory, NumPy's pairwise tree evaluated from a host-built plan (canonical.cuh, PwPlan);
//   * sequential CHAINS (np.dot's HALF_dot / the exact-dot definitions for fp32 / fp64): a warp takes 32 consecutive
//     rows, stages 512-byte column chunks of all 32 rows in shared memory (coalesced), then every lane continues the
//     chain of its own row out of shared memory (conflict-free 128-bit reads, odd pitch).
// Bit-identical to canonical_similarity (same operations in the same order); rows too long for shared memory or for the
// plan keep the thread-per-row kernels of finalize.cu / ingest.cu.
#include "canonical.cuh"
#include "hdb_internal.h"
#include "../../include/hyperdb_b200.h"

namespace hdb {

constexpr int kRwThreads = 256;
constexpr int kRwWarps = kRwThreads / 32;
constexpr int kChunkBytes = 512;                 // per row and chunk: 32 lanes x 16 bytes
constexpr int kTilePitch = kChunkBytes + 16;     // 33 sixteen-byte units: lane-per-row reads are conflict-free
constexpr int kChainWarps = 4;

template <int DT> struct ElemOf;
template <> struct ElemOf<0> { using T = __half; };
template <> struct ElemOf<1> { using T = float; };
template <> struct ElemOf<2> { using T = double; };

template <typename C> __device__ __forceinline__ C widen(__half x) { return (C)__half2float(x); }
template <typename C> __device__ __forceinline__ C widen(float x) { return (C)x; }
template <typename C> __device__ __forceinline__ C widen(double x) { return (C)x; }

__device__ __forceinline__ bool rw_kept(const RowFilter& f, int64_t row) {
  if (row < f.lo || row >= f.hi) return false;
  if (f.mask && !((f.mask[row >> 5] >> (row & 31)) & 1u)) return false;
  return true;
}

// fn(j0, e, cnt): elements j0 .. j0+cnt-1 of the row (cnt = 16/sizeof(T) on the vector path, else 1); four 128-bit loads
// in flight per lane.
template <typename T, bool VEC, typename Fn>
__device__ __forceinline__ void warp_row_foreach(const T* row, int d, int lane, Fn fn) {
  if (VEC) {
    constexpr int PER = 16 / (int)sizeof(T);
    const int nvec = d / PER;
    const uint4* p = reinterpret_cast<const uint4*>(row);
    for (int c0 = lane; c0 < nvec; c0 += 128) {
      uint4 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + 32 * u;
        raw[u] = c < nvec ? ld_stream16(p + c) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + 32 * u;
        if (c < nvec) fn(c * PER, reinterpret_cast<const T*>(&raw[u]), PER);
      }
    }
  } else {
    for (int j = lane; j < d; j += 32) { const T v = row[j]; fn(j, &v, 1); }
  }
}

// the row's values widened to the carrier C into the warp's shared-memory buffer
template <typename T, bool VEC, typename C>
__device__ __forceinline__ void stage_row(const T* row, int d, int lane, C* vals) {
  warp_row_foreach<T, VEC>(row, d, lane, [&](int j0, const T* e, int cnt) {
    for (int i = 0; i < cnt; ++i) vals[j0 + i] = widen<C>(e[i]);
  });
}

struct RowwiseSmem {     // layout of the dynamic shared memory of the warp-per-row kernels
  int warps;             // rows in flight per CTA
  size_t q_bytes;        // CTA-wide query (0 when the kernel has none)
  size_t per_warp;       // row buffer + plan result slots
  size_t total;
};
static RowwiseSmem rowwise_smem(int64_t d, size_t elem, bool with_query) {
  RowwiseSmem s;
  s.q_bytes = with_query ? (((size_t)d * elem + 15) & ~size_t(15)) : 0;
  s.per_warp = (((size_t)d * elem + 15) & ~size_t(15)) + 2 * kPlanLeaves * elem;
  const size_t budget = 96 * 1024;
  s.warps = 0;
  if (s.q_bytes + s.per_warp <= budget) {
    s.warps = (int)((budget - s.q_bytes) / s.per_warp);
    if (s.warps > kRwWarps) s.warps = kRwWarps;
  }
  s.total = s.q_bytes + (size_t)s.warps * s.per_warp;
  return s;
}

static int rowwise_grid(int64_t n, int warps) {
  int64_t blocks = (n + warps - 1) / warps;
  const int64_t cap = 148 * 8;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

// ---------------------------------------------------------------------------------------------
// row_stats: canonical norm, inverse norm, ||v||^2, NaN flag, max ||v||, max ||v|| / canonical norm
// ---------------------------------------------------------------------------------------------
template <int DT, bool VEC>
__global__ void __launch_bounds__(kRwThreads) row_stats_warp_kernel(const void* rows, int64_t n, int d, void* norms, void* inv_norms,
                                                                    float* sqnorms, float* stats, int* nan_flag, PwPlan plan, int warps,
                                                                    int per_warp) {
  using A = Arith<DT>;
  using C = typename A::C;
  using T = typename ElemOf<DT>::T;
  extern __shared__ __align__(16) unsigned char rw_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= warps) return;
  C* vals = reinterpret_cast<C*>(rw_smem + (size_t)warp * per_warp);
  C* res = vals + ((per_warp / sizeof(C)) - 2 * kPlanLeaves);
  float my_norm = 0.f, my_ratio = 0.f;
  bool bad = false;
  for (int64_t row = (int64_t)blockIdx.x * warps + warp; row < n; row += (int64_t)gridDim.x * warps) {
    const T* rp = reinterpret_cast<const T*>(rows) + row * d;
    double tsq = 0.0;
    warp_row_foreach<T, VEC>(rp, d, lane, [&](int j0, const T* e, int cnt) {
      for (int i = 0; i < cnt; ++i) {
        const C v = widen<C>(e[i]);
        bad |= (v != v);
        tsq = fma((double)v, (double)v, tsq);
        vals[j0 + i] = v;
      }
    });
    __syncwarp();
    C cn = A::sqrt(pairwise_sum_plan<DT>([&](int i) { const C v = vals[i]; return A::mul(v, v); }, plan, lane, res));
    if (cn == C(0)) cn = C(1);                          // ranking_algorithm.py:14-15
    tsq = warp_sum(tsq);
    if (lane == 0) {
      if (DT == 2) {
        reinterpret_cast<double*>(norms)[row] = (double)cn;
        reinterpret_cast<double*>(inv_norms)[row] = 1.0 / (double)cn;
      } else {
        reinterpret_cast<float*>(norms)[row] = (float)cn;
        reinterpret_cast<float*>(inv_norms)[row] = 1.0f / (float)cn;
      }
      if (sqnorms) sqnorms[row] = (float)tsq;           // ||v||^2 for the norm-expansion form of batched euclidean
    }
    const double tn = sqrt(tsq);
    my_norm = fmaxf(my_norm, (float)fmin(tn * (1.0 + 1e-6), 3.0e38));
    const double ratio = tn / (double)cn;
    my_ratio = fmaxf(my_ratio, (ratio == ratio) ? (float)fmin(ratio * (1.0 + 1e-6), 3.0e38) : 3.0e38f);
  }
  const unsigned any_bad = __ballot_sync(kFull, bad);
  if (lane == 0) {
    atomicMax(reinterpret_cast<int*>(stats), __float_as_int(my_norm));        // non-negative floats order like their bits
    atomicMax(reinterpret_cast<int*>(stats) + 1, __float_as_int(my_ratio));
    if (any_bad) atomicOr(nan_flag, 1);
  }
}

template <int DT>
static int launch_row_stats_dt(const MatrixView& m, void* norms, void* inv_norms, float* sqnorms, float* d_stats, int* d_nan,
                               const PwPlan& plan, const RowwiseSmem& sm, bool vec, cudaStream_t s) {
  auto kern = vec ? row_stats_warp_kernel<DT, true> : row_stats_warp_kernel<DT, false>;
  if (sm.total > 48 * 1024) HDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm.total));
  kern<<<rowwise_grid(m.n, sm.warps), kRwThreads, sm.total, s>>>(m.rows, m.n, (int)m.d, norms, inv_norms, sqnorms, d_stats, d_nan, plan,
                                                                 sm.warps, (int)sm.per_warp);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

static bool rows_vectorisable(const MatrixView& m) {
  return ((m.d * dtype_size(m.dtype)) % 16 == 0) && ((reinterpret_cast<uintptr_t>(m.rows) & 15) == 0);
}

// 1 = handled here; 0 = the caller keeps the thread-per-row kernel
int launch_row_stats_warp(const MatrixView& m, void* norms, void* inv_norms, float* sqnorms, float* d_stats, int* d_nan, cudaStream_t s,
                          int* handled) {
  *handled = 0;
  PwPlan plan;
  if (m.d > (1 << 20) || !pw_plan_build(plan, (int)m.d)) return 0;
  const RowwiseSmem sm = rowwise_smem(m.d, m.dtype == 2 ? 8 : 4, false);
  if (sm.warps < 1) return 0;
  *handled = 1;
  const bool vec = rows_vectorisable(m);
  if (m.dtype == 0) return launch_row_stats_dt<0>(m, norms, inv_norms, sqnorms, d_stats, d_nan, plan, sm, vec, s);
  if (m.dtype == 1) return launch_row_stats_dt<1>(m, norms, inv_norms, sqnorms, d_stats, d_nan, plan, sm, vec, s);
  return launch_row_stats_dt<2>(m, norms, inv_norms, sqnorms, d_stats, d_nan, plan, sm, vec, s);
}

// ---------------------------------------------------------------------------------------------
// pearson_stats: np.mean / np.std per row in NumPy's arithmetic, 1/(std*d), certificate statistics
// ---------------------------------------------------------------------------------------------
template <int DT, bool VEC>
__global__ void __launch_bounds__(kRwThreads) pearson_stats_warp_kernel(const void* rows, int64_t n, int d, void* pmean, void* pstd,
                                                                        void* pscale, float* stats, PwPlan plan, int warps, int per_warp) {
  using A = Arith<DT>;
  using C = typename A::C;
  using T = typename ElemOf<DT>::T;
  extern __shared__ __align__(16) unsigned char rw_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= warps) return;
  C* vals = reinterpret_cast<C*>(rw_smem + (size_t)warp * per_warp);
  C* res = vals + ((per_warp / sizeof(C)) - 2 * kPlanLeaves);
  float my_ratio = 0.f, my_cratio = 0.f, my_negstd = -3.0e38f;
  for (int64_t row = (int64_t)blockIdx.x * warps + warp; row < n; row += (int64_t)gridDim.x * warps) {
    const T* rp = reinterpret_cast<const T*>(rows) + row * d;
    stage_row<T, VEC>(rp, d, lane, vals);
    __syncwarp();
    C mean, sd;
    mean_std_core<DT>([&](int j) { return vals[j]; }, d, false,
                      [&](auto term) { return pairwise_sum_plan<1>(term, plan, lane, reinterpret_cast<float*>(res)); },
                      [&](auto term) { return pairwise_sum_plan<DT>(term, plan, lane, res); }, &mean, &sd);
    double true_sq = 0.0, cen_sq = 0.0;
    for (int j = lane; j < d; j += 32) {
      const double v = (double)vals[j];
      true_sq = fma(v, v, true_sq);
      cen_sq = fma(v - (double)mean, v - (double)mean, cen_sq);
    }
    true_sq = warp_sum(true_sq);
    cen_sq = warp_sum(cen_sq);
    const double scale = (sd == C(0)) ? __longlong_as_double(0x7ff8000000000000ll) : 1.0 / ((double)sd * (double)d);
    if (lane == 0) {
      if (DT == 2) {
        reinterpret_cast<double*>(pmean)[row] = (double)mean;
        reinterpret_cast<double*>(pstd)[row] = (double)sd;
        reinterpret_cast<double*>(pscale)[row] = scale;
      } else {
        reinterpret_cast<float*>(pmean)[row] = (float)mean;
        reinterpret_cast<float*>(pstd)[row] = (float)sd;
        reinterpret_cast<float*>(pscale)[row] = (float)scale;
      }
    }
    if (sd != C(0) && sd == sd) {
      const double ratio = sqrt(true_sq) / ((double)sd * sqrt((double)d));
      my_ratio = fmaxf(my_ratio, (ratio == ratio) ? (float)fmin(ratio * (1.0 + 1e-6), 3.0e38) : 3.0e38f);
      const double cratio = sqrt(cen_sq) / ((double)sd * sqrt((double)d));
      my_cratio = fmaxf(my_cratio, (cratio == cratio) ? (float)fmin(cratio * (1.0 + 1e-6), 3.0e38) : 3.0e38f);
      my_negstd = fmaxf(my_negstd, -(float)((double)sd * (1.0 - 1e-6)));
    }
    __syncwarp();
  }
  if (lane == 0) {
    atomicMax(reinterpret_cast<int*>(stats), __float_as_int(my_ratio));                 // non-negative: bit order = value order
    atomicMin(reinterpret_cast<unsigned*>(stats) + 1, __float_as_uint(my_negstd));      // non-positive: larger value = smaller bits
    atomicMax(reinterpret_cast<int*>(stats) + 2, __float_as_int(my_cratio));
  }
}

template <int DT>
static int launch_pearson_stats_dt(const MatrixView& m, void* pmean, void* pstd, void* pscale, float* d_stats, const PwPlan& plan,
                                   const RowwiseSmem& sm, bool vec, cudaStream_t s) {
  auto kern = vec ? pearson_stats_warp_kernel<DT, true> : pearson_stats_warp_kernel<DT, false>;
  if (sm.total > 48 * 1024) HDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm.total));
  kern<<<rowwise_grid(m.n, sm.warps), kRwThreads, sm.total, s>>>(m.rows, m.n, (int)m.d, pmean, pstd, pscale, d_stats, plan, sm.warps,
                                                                 (int)sm.per_warp);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

int launch_pearson_stats_warp(const MatrixView& m, void* pmean, void* pstd, void* pscale, float* d_stats, cudaStream_t s, int* handled) {
  *handled = 0;
  PwPlan plan;
  if (m.d > (1 << 20) || !pw_plan_build(plan, (int)m.d)) return 0;
  const RowwiseSmem sm = rowwise_smem(m.d, m.dtype == 2 ? 8 : 4, false);
  if (sm.warps < 1) return 0;
  *handled = 1;
  const bool vec = rows_vectorisable(m);
  if (m.dtype == 0) return launch_pearson_stats_dt<0>(m, pmean, pstd, pscale, d_stats, plan, sm, vec, s);
  if (m.dtype == 1) return launch_pearson_stats_dt<1>(m, pmean, pstd, pscale, d_stats, plan, sm, vec, s);
  return launch_pearson_stats_dt<2>(m, pmean, pstd, pscale, d_stats, plan, sm, vec, s);
}

// ---------------------------------------------------------------------------------------------
// get_norm_vector as a function (hdb_normalize_rows)
// ---------------------------------------------------------------------------------------------
template <int DT, bool VEC>
__global__ void __launch_bounds__(kRwThreads) normalize_rows_warp_kernel(const void* src, void* dst, int64_t n, int d, PwPlan plan, int warps,
                                                                         int per_warp) {
  using A = Arith<DT>;
  using C = typename A::C;
  using T = typename ElemOf<DT>::T;
  extern __shared__ __align__(16) unsigned char rw_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= warps) return;
  C* vals = reinterpret_cast<C*>(rw_smem + (size_t)warp * per_warp);
  C* res = vals + ((per_warp / sizeof(C)) - 2 * kPlanLeaves);
  for (int64_t row = (int64_t)blockIdx.x * warps + warp; row < n; row += (int64_t)gridDim.x * warps) {
    const T* rp = reinterpret_cast<const T*>(src) + row * d;
    stage_row<T, VEC>(rp, d, lane, vals);
    __syncwarp();
    C cn = A::sqrt(pairwise_sum_plan<DT>([&](int i) { const C v = vals[i]; return A::mul(v, v); }, plan, lane, res));
    if (cn == C(0)) cn = C(1);
    T* out = reinterpret_cast<T*>(dst) + row * d;
    for (int j = lane; j < d; j += 32) {
      const C u = A::div(vals[j], cn);
      if (DT == 0) reinterpret_cast<__half*>(out)[j] = __float2half_rn((float)u);
      else if (DT == 1) reinterpret_cast<float*>(out)[j] = (float)u;
      else reinterpret_cast<double*>(out)[j] = (double)u;
    }
    __syncwarp();
  }
}

int launch_normalize_rows_warp(int dtype, int64_t n, int64_t d, const void* src, void* dst, cudaStream_t s, int* handled) {
  *handled = 0;
  PwPlan plan;
  if (d > (1 << 20) || !pw_plan_build(plan, (int)d)) return 0;
  const RowwiseSmem sm = rowwise_smem(d, dtype == 2 ? 8 : 4, false);
  if (sm.warps < 1) return 0;
  *handled = 1;
  const bool vec = ((d * dtype_size(dtype)) % 16 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  const int grid = rowwise_grid(n, sm.warps);
#define HDB_NORM_LAUNCH(DT)                                                                                                     \
  do {                                                                                                                          \
    auto kern = vec ? normalize_rows_warp_kernel<DT, true> : normalize_rows_warp_kernel<DT, false>;                             \
    if (sm.total > 48 * 1024) HDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm.total)); \
    kern<<<grid, kRwThreads, sm.total, s>>>(src, dst, n, (int)d, plan, sm.warps, (int)sm.per_warp);                            \
  } while (0)
  if (dtype == 0) HDB_NORM_LAUNCH(0);
  else if (dtype == 1) HDB_NORM_LAUNCH(1);
  else HDB_NORM_LAUNCH(2);
#undef HDB_NORM_LAUNCH
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Full similarity vector, pairwise metrics (euclidean, manhattan, pearson): warp per row
// ---------------------------------------------------------------------------------------------
struct ScoreOut {
  double* totals;        // exact path: float64 total per row (dropped rows = -NaN), or nullptr
  void* typed;           // hdb_scores: the metric function's own output dtype, or nullptr
  int distance;          // euclidean: emit the distance itself (get_similarity_score=False, ranking_algorithm.py:49-52)
};

template <int RDT>
__device__ __forceinline__ void write_score(const ScoreOut& o, const RowFilter& f, int metric, int64_t row, double sim) {
  if (o.totals) {
    o.totals[row] = rw_kept(f, row) ? total_score(sim, f.decay, f.bias, row) : __longlong_as_double(-1ll);   // -NaN sorts last
    return;
  }
  if (metric == HDB_PEARSON) reinterpret_cast<double*>(o.typed)[row] = sim;        // np.zeros(N) receives the quotients
  else if (RDT == 0) reinterpret_cast<__half*>(o.typed)[row] = __float2half_rn((float)sim);
  else if (RDT == 1) reinterpret_cast<float*>(o.typed)[row] = (float)sim;
  else reinterpret_cast<double*>(o.typed)[row] = sim;
}

template <int SDT, int RDT, bool VEC>
__global__ void __launch_bounds__(kRwThreads) scores_pairwise_kernel(MatrixView m, RowFilter f, int metric, const double* qc, const double* qaux,
                                                                     ScoreOut o, PwPlan plan, int warps, int per_warp, int q_bytes) {
  using AR = Arith<RDT>;
  using AS = Arith<SDT>;
  using C = typename AR::C;
  using CS = typename AS::C;
  using T = typename ElemOf<SDT>::T;
  extern __shared__ __align__(16) unsigned char rw_smem[];
  C* s_q = reinterpret_cast<C*>(rw_smem);
  const int d = (int)m.d;
  for (int j = threadIdx.x; j < d; j += kRwThreads) s_q[j] = AR::from_double(qc[j]);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= warps) return;
  const double qstd = qaux ? qaux[0] : 1.0;            // pearson: np.std(query)
  C* terms = reinterpret_cast<C*>(rw_smem + q_bytes + (size_t)warp * per_warp);
  C* res = terms + ((per_warp / sizeof(C)) - 2 * kPlanLeaves);
  for (int64_t row = (int64_t)blockIdx.x * warps + warp; row < m.n; row += (int64_t)gridDim.x * warps) {
    const T* rp = reinterpret_cast<const T*>(m.rows) + row * d;
    CS mean = CS(0);
    if (metric == HDB_PEARSON) mean = SDT == 2 ? (CS) reinterpret_cast<const double*>(m.pmean)[row] : (CS) reinterpret_cast<const float*>(m.pmean)[row];
    warp_row_foreach<T, VEC>(rp, d, lane, [&](int j0, const T* e, int cnt) {
      for (int i = 0; i < cnt; ++i) {
        const CS vs = widen<CS>(e[i]);
        C term;
        if (metric == HDB_PEARSON) term = AR::mul((C)AS::sub(vs, mean), s_q[j0 + i]);      // (v - mean) in S, product in R
        else {
          const C df = AR::sub((C)vs, s_q[j0 + i]);
          term = metric == HDB_EUCLIDEAN ? AR::mul(df, df) : C(fabs(df));
        }
        terms[j0 + i] = term;
      }
    });
    __syncwarp();
    C sum = pairwise_sum_plan<RDT>([&](int i) { return terms[i]; }, plan, lane, res);
    if (lane == 0) {
      double sim;
      if (metric == HDB_PEARSON) {
        const double sd = SDT == 2 ? reinterpret_cast<const double*>(m.pstd)[row] : (double)reinterpret_cast<const float*>(m.pstd)[row];
        sim = pearson_quotient<RDT>(sum, sd, qstd, d);
      } else {
        if (metric == HDB_EUCLIDEAN) sum = AR::sqrt(sum);
        sim = o.distance ? (double)sum : (double)AR::div(C(1), AR::add(C(1), sum));
      }
      write_score<RDT>(o, f, metric, row, sim);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// Full similarity vector, chain metrics (dot, cosine): a warp owns 32 consecutive rows, lane = row
// ---------------------------------------------------------------------------------------------
template <int SDT, int RDT, bool VEC>
__global__ void __launch_bounds__(32 * kChainWarps) scores_chain_kernel(MatrixView m, RowFilter f, int metric, const double* qc, ScoreOut o,
                                                                        int q_bytes) {
  using AR = Arith<RDT>;
  using AS = Arith<SDT>;
  using C = typename AR::C;
  using CS = typename AS::C;
  using T = typename ElemOf<SDT>::T;
  constexpr int PER = 16 / (int)sizeof(T);
  constexpr int K = kChunkBytes / (int)sizeof(T);          // elements per chunk
  extern __shared__ __align__(16) unsigned char rw_smem[];
  C* s_q = reinterpret_cast<C*>(rw_smem);
  const int d = (int)m.d;
  for (int j = threadIdx.x; j < d; j += 32 * kChainWarps) s_q[j] = AR::from_double(qc[j]);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* tile = rw_smem + q_bytes + (size_t)warp * 32 * kTilePitch;
  const int64_t row_bytes = (int64_t)d * sizeof(T);
  const int64_t ntiles = (m.n + 31) / 32;
  const char* base = reinterpret_cast<const char*>(m.rows);
  for (int64_t t = (int64_t)blockIdx.x * kChainWarps + warp; t < ntiles; t += (int64_t)gridDim.x * kChainWarps) {
    const int64_t row0 = t * 32, row = row0 + lane;
    const int rows_here = (int)((m.n - row0) < 32 ? (m.n - row0) : 32);
    CS nrm = CS(1);
    if (metric == HDB_COSINE && row < m.n)
      nrm = SDT == 2 ? (CS) reinterpret_cast<const double*>(m.norms)[row] : (CS) reinterpret_cast<const float*>(m.norms)[row];
    const bool any_div = __any_sync(kFull, nrm != CS(1));            // x / 1 == x: skip the divisions when every norm is 1
    float acc32 = 0.f;
    double acc64 = 0.0, comp = 0.0;
    for (int k0 = 0; k0 < d; k0 += K) {
      // ---- load: piece `lane` of the chunk of each of the 32 rows (coalesced), eight rows in flight
      if (VEC) {
        const int64_t off = (int64_t)k0 * sizeof(T) + lane * 16;
        const bool col_ok = off < row_bytes;
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 8) {
          uint4 raw[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int r = r0 + u;
            raw[u] = (col_ok && r < rows_here) ? ld_stream16(base + (row0 + r) * row_bytes + off) : make_uint4(0, 0, 0, 0);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) *reinterpret_cast<uint4*>(tile + (r0 + u) * kTilePitch + lane * 16) = raw[u];
        }
      } else {
        for (int r = 0; r < rows_here; ++r) {
          const T* rp = reinterpret_cast<const T*>(base + (row0 + r) * row_bytes);
          T* dst = reinterpret_cast<T*>(tile + r * kTilePitch);
          for (int e = lane; e < K && k0 + e < d; e += 32) dst[e] = rp[k0 + e];
        }
      }
      __syncwarp();
      // ---- chain: this lane's row, in column order
      const int nval = (d - k0) < K ? (d - k0) : K;
      const unsigned char* mine = tile + lane * kTilePitch;
      for (int e0 = 0; e0 < nval; e0 += PER) {
        const uint4 raw = *reinterpret_cast<const uint4*>(mine + e0 * sizeof(T));
        const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          if (e0 + i < nval) {
            const CS vs = widen<CS>(e[i]);
            const C term = (metric == HDB_COSINE && any_div) ? (C)AS::div(vs, nrm) : (C)vs;
            const C q = s_q[k0 + e0 + i];
            if (RDT == 0) {
              acc32 = __fmaf_rn((float)term, (float)q, acc32);                   // HALF_dot: sequential float32 chain
            } else if (RDT == 1) {
              acc64 = __fma_rn((double)term, (double)q, acc64);                  // exact products, float64 chain
            } else {                                                             // Ogita-Rump-Oishi Dot2
              const double a = (double)term, b = (double)q;
              const double p = __dmul_rn(a, b);
              const double er = __fma_rn(a, b, -p);
              const double tt = __dadd_rn(acc64, p);
              const double z = __dsub_rn(tt, acc64);
              const double err = __dadd_rn(__dsub_rn(acc64, __dsub_rn(tt, z)), __dsub_rn(p, z));
              comp = __dadd_rn(comp, __dadd_rn(er, err));
              acc64 = tt;
            }
          }
        }
      }
      __syncwarp();
    }
    if (row < m.n) {
      double sim;
      if (RDT == 0) sim = (double)Arith<0>::rnd(acc32);
      else if (RDT == 1) sim = (double)(float)acc64;
      else sim = __dadd_rn(acc64, comp);
      write_score<RDT>(o, f, metric, row, sim);
    }
  }
}

template <int SDT, int RDT>
static int launch_scores_combo(const MatrixView& m, const RowFilter& f, int metric, const double* qc, const double* qaux, const ScoreOut& o,
                               bool vec, cudaStream_t s, int* handled) {
  const size_t elem = RDT == 2 ? 8 : 4;
  if (metric == HDB_DOT || metric == HDB_COSINE) {
    const size_t q_bytes = ((size_t)m.d * elem + 15) & ~size_t(15);
    const size_t smem = q_bytes + (size_t)kChainWarps * 32 * kTilePitch;
    if (smem > 200 * 1024) return 0;
    *handled = 1;
    auto kern = vec ? scores_chain_kernel<SDT, RDT, true> : scores_chain_kernel<SDT, RDT, false>;
    if (smem > 48 * 1024) HDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = ((m.n + 31) / 32 + kChainWarps - 1) / kChainWarps;
    if (blocks > 148 * 6) blocks = 148 * 6;
    kern<<<(unsigned)blocks, 32 * kChainWarps, smem, s>>>(m, f, metric, qc, o, (int)q_bytes);
    HDB_LAUNCHED();
    HDB_CUDA(cudaGetLastError());
    return 0;
  }
  PwPlan plan;
  if (!pw_plan_build(plan, (int)m.d)) return 0;
  const RowwiseSmem sm = rowwise_smem(m.d, elem, true);
  if (sm.warps < 1) return 0;
  *handled = 1;
  auto kern = vec ? scores_pairwise_kernel<SDT, RDT, true> : scores_pairwise_kernel<SDT, RDT, false>;
  if (sm.total > 48 * 1024) HDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm.total));
  kern<<<rowwise_grid(m.n, sm.warps), kRwThreads, sm.total, s>>>(m, f, metric, qc, metric == HDB_PEARSON ? qaux : nullptr, o, plan, sm.warps,
                                                                 (int)sm.per_warp, (int)sm.q_bytes);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

// Full score vector of every row in the reference's arithmetic.  totals != nullptr: float64 totals with the row
// subset and the time decay applied (exact path); else `typed` receives the metric function's own output.
// *handled = 0: this shape is not covered here (packed-bit metrics, very long rows) -- the caller keeps its kernel.
int launch_scores_rowwise(const MatrixView& m, const RowFilter& f, int metric, int rdt, const double* qc, const double* qaux,
                          double* totals, void* typed, int distance, cudaStream_t s, int* handled) {
  *handled = 0;
  if (m.n == 0) { *handled = 1; return 0; }
  if (metric == HDB_HAMMING || metric == HDB_JACCARD) return 0;
  if (m.d > (1 << 20)) return 0;
  ScoreOut o;
  o.totals = totals; o.typed = typed; o.distance = (metric == HDB_EUCLIDEAN) ? distance : 0;
  const bool vec = rows_vectorisable(m);
  switch (m.dtype * 3 + rdt) {
    case 0: return launch_scores_combo<0, 0>(m, f, metric, qc, qaux, o, vec, s, handled);
    case 1: return launch_scores_combo<0, 1>(m, f, metric, qc, qaux, o, vec, s, handled);
    case 2: return launch_scores_combo<0, 2>(m, f, metric, qc, qaux, o, vec, s, handled);
    case 4: return launch_scores_combo<1, 1>(m, f, metric, qc, qaux, o, vec, s, handled);
    case 5: return launch_scores_combo<1, 2>(m, f, metric, qc, qaux, o, vec, s, handled);
    case 8: return launch_scores_combo<2, 2>(m, f, metric, qc, qaux, o, vec, s, handled);
    default: return 0;
  }
}

}  // namespace hdb
