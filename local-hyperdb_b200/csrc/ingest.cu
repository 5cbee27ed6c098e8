// Ingest-time and query-preparation kernels (one-off / tiny; not on the streaming path).
//
//  row_stats     get_norm_vector's per-row norm (hyperdb/ranking_algorithm.py:8-21) computed ONCE with
//                NumPy's exact arithmetic, instead of once per cosine query over the whole matrix;
//                the NaN scan of :150; the two statistics of the certification bound.
//  pack_bits     check_and_binarize_vectors (:116-126): bit = (x > 0), 32 columns per word.
//  kept_ts_max / decay / stage1    the time-decay column of :179-183 and hyperdb.py:1334-1346.
//  prep_query    canonical query (unit query for cosine, normalised in the QUERY's dtype, :38),
//                accumulate-type copy for the sweep, sign bits, ||q||, NaN flag.
#include <cub/device/device_scan.cuh>

#include "canonical.cuh"
#include <cstring>
#include "hdb_internal.h"
#include "../../include/hyperdb_b200.h"

namespace hdb {

__device__ __forceinline__ unsigned long long order_f64(double x) {
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
double decode_ordered_double(unsigned long long o) {
  unsigned long long b = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
  double d;
  memcpy(&d, &b, 8);
  return d;
}

// ---------------------------------------------------------------------------------------------
template <int T>
__global__ void row_stats_kernel(const void* rows, int64_t n, int64_t d, void* norms, void* inv_norms, float* sqnorms,
                                 float* stats, int* nan_flag) {
  using A = Arith<T>;
  using C = typename A::C;
  int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  float my_norm = 0.f, my_ratio = 0.f;
  bool bad = false;
  if (row < n) {
    const char* base = reinterpret_cast<const char*>(rows) + row * d * dtype_size(T);
    C cn = canonical_norm<T>(base, d);
    double true_sq = 0.0;
    for (int64_t j = 0; j < d; ++j) {
      double v = load_as_double(base, T, j);
      bad |= (v != v);
      true_sq += v * v;
    }
    if (cn == C(0)) cn = C(1);                         // ranking_algorithm.py:14-15
    double tn = sqrt(true_sq);
    if (T == 2) {
      reinterpret_cast<double*>(norms)[row] = (double)cn;
      reinterpret_cast<double*>(inv_norms)[row] = 1.0 / (double)cn;
    } else {
      reinterpret_cast<float*>(norms)[row] = (float)cn;
      reinterpret_cast<float*>(inv_norms)[row] = 1.0f / (float)cn;
    }
    if (sqnorms) sqnorms[row] = (float)true_sq;          // ||v||^2 for the norm-expansion form of batched euclidean
    my_norm = (float)fmin(tn * (1.0 + 1e-6), 3.0e38);
    double ratio = tn / (double)cn;
    my_ratio = (ratio == ratio) ? (float)fmin(ratio * (1.0 + 1e-6), 3.0e38) : 3.0e38f;
  }
  // non-negative floats order like their bit patterns
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    my_norm = fmaxf(my_norm, __shfl_xor_sync(kFull, my_norm, o));
    my_ratio = fmaxf(my_ratio, __shfl_xor_sync(kFull, my_ratio, o));
  }
  unsigned any_bad = __ballot_sync(kFull, bad);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(reinterpret_cast<int*>(stats), __float_as_int(my_norm));
    atomicMax(reinterpret_cast<int*>(stats) + 1, __float_as_int(my_ratio));
    if (any_bad) atomicOr(nan_flag, 1);
  }
}

int launch_row_stats(const MatrixView& m, void* norms, void* inv_norms, float* sqnorms, float* d_stats, int* d_nan, cudaStream_t s) {
  if (m.n == 0) return 0;
  int handled = 0;
  HDB_TRY(launch_row_stats_warp(m, norms, inv_norms, sqnorms, d_stats, d_nan, s, &handled));      // warp per row, HBM rate
  if (handled) return 0;
  int threads = 128;
  int64_t blocks = (m.n + threads - 1) / threads;
  if (blocks > 0x7fffffff) return fail("row_stats: too many rows");
  if (m.dtype == 0) row_stats_kernel<0><<<(unsigned)blocks, threads, 0, s>>>(m.rows, m.n, m.d, norms, inv_norms, sqnorms, d_stats, d_nan);
  else if (m.dtype == 1) row_stats_kernel<1><<<(unsigned)blocks, threads, 0, s>>>(m.rows, m.n, m.d, norms, inv_norms, sqnorms, d_stats, d_nan);
  else row_stats_kernel<2><<<(unsigned)blocks, threads, 0, s>>>(m.rows, m.n, m.d, norms, inv_norms, sqnorms, d_stats, d_nan);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// pearson_correlation's per-row statistics (hyperdb/ranking_algorithm.py:91,94: np.mean / np.std along axis 1), computed
// once per matrix in NumPy's arithmetic instead of once per query; plus the sweep's scale 1/(std*d) and the two
// statistics of the certificate.  Thread per row (one-off pass).
template <int T>
__global__ void pearson_stats_kernel(const void* rows, int64_t n, int64_t d, void* pmean, void* pstd, void* pscale, float* stats) {
  using A = Arith<T>;
  using C = typename A::C;
  int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  float my_ratio = 0.f, my_cratio = 0.f, my_negstd = -3.0e38f;
  if (row < n) {
    const char* base = reinterpret_cast<const char*>(rows) + row * d * dtype_size(T);
    C mean, sd;
    canonical_mean_std<T>(base, d, false, &mean, &sd);
    double true_sq = 0.0, cen_sq = 0.0;
    for (int64_t j = 0; j < d; ++j) {
      const double v = load_as_double(base, T, j);
      true_sq += v * v;
      cen_sq += (v - (double)mean) * (v - (double)mean);
    }
    const double scale = (sd == C(0)) ? __longlong_as_double(0x7ff8000000000000ll) : 1.0 / ((double)sd * (double)d);
    if (T == 2) {
      reinterpret_cast<double*>(pmean)[row] = (double)mean;
      reinterpret_cast<double*>(pstd)[row] = (double)sd;
      reinterpret_cast<double*>(pscale)[row] = scale;
    } else {
      reinterpret_cast<float*>(pmean)[row] = (float)mean;
      reinterpret_cast<float*>(pstd)[row] = (float)sd;
      reinterpret_cast<float*>(pscale)[row] = (float)scale;
    }
    if (sd != C(0) && sd == sd) {
      const double ratio = sqrt(true_sq) / ((double)sd * sqrt((double)d));
      my_ratio = (ratio == ratio) ? (float)fmin(ratio * (1.0 + 1e-6), 3.0e38) : 3.0e38f;
      const double cratio = sqrt(cen_sq) / ((double)sd * sqrt((double)d));
      my_cratio = (cratio == cratio) ? (float)fmin(cratio * (1.0 + 1e-6), 3.0e38) : 3.0e38f;
      my_negstd = -(float)((double)sd * (1.0 - 1e-6));
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    my_ratio = fmaxf(my_ratio, __shfl_xor_sync(kFull, my_ratio, o));
    my_cratio = fmaxf(my_cratio, __shfl_xor_sync(kFull, my_cratio, o));
    my_negstd = fmaxf(my_negstd, __shfl_xor_sync(kFull, my_negstd, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(reinterpret_cast<int*>(stats), __float_as_int(my_ratio));                 // non-negative: bit order = value order
    atomicMin(reinterpret_cast<unsigned*>(stats) + 1, __float_as_uint(my_negstd));      // non-positive: larger value = smaller bits
    atomicMax(reinterpret_cast<int*>(stats) + 2, __float_as_int(my_cratio));
  }
}

// d_stats[0], [2] start at the running maxima (0 at first), d_stats[1] at the bits of the running -min std (0xffffffff at first)
int launch_pearson_stats(const MatrixView& m, void* pmean, void* pstd, void* pscale, float* d_stats, cudaStream_t s) {
  if (m.n == 0) return 0;
  int handled = 0;
  HDB_TRY(launch_pearson_stats_warp(m, pmean, pstd, pscale, d_stats, s, &handled));
  if (handled) return 0;
  int threads = 128;
  int64_t blocks = (m.n + threads - 1) / threads;
  if (blocks > 0x7fffffff) return fail("pearson_stats: too many rows");
  if (m.dtype == 0) pearson_stats_kernel<0><<<(unsigned)blocks, threads, 0, s>>>(m.rows, m.n, m.d, pmean, pstd, pscale, d_stats);
  else if (m.dtype == 1) pearson_stats_kernel<1><<<(unsigned)blocks, threads, 0, s>>>(m.rows, m.n, m.d, pmean, pstd, pscale, d_stats);
  else pearson_stats_kernel<2><<<(unsigned)blocks, threads, 0, s>>>(m.rows, m.n, m.d, pmean, pstd, pscale, d_stats);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// One warp per row: coalesced 32-element reads, ballot -> one word.
__global__ void pack_bits_kernel(const void* rows, int dtype, int64_t n, int64_t d, uint32_t* bits, int words) {
  int lane = threadIdx.x & 31;
  int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp; row < n; row += nwarps) {
    const char* base = reinterpret_cast<const char*>(rows) + row * d * dtype_size(dtype);
    for (int w = 0; w < words; ++w) {
      int64_t j = (int64_t)w * 32 + lane;
      bool on = (j < d) && (load_as_double(base, dtype, j) > 0.0);
      unsigned word = __ballot_sync(kFull, on);
      if (lane == 0) bits[row * words + w] = word;
    }
  }
}

int launch_pack_bits(const MatrixView& m, uint32_t* bits, int words, cudaStream_t s) {
  if (m.n == 0) return 0;
  int threads = 256;
  int64_t blocks = (m.n * 32 + threads - 1) / threads;
  if (blocks > 148 * 32) blocks = 148 * 32;
  pack_bits_kernel<<<(unsigned)blocks, threads, 0, s>>>(m.rows, m.dtype, m.n, m.d, bits, words);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool row_kept(const RowFilter& f, int64_t row) {
  if (row < f.lo || row >= f.hi) return false;
  if (f.mask && !((f.mask[row >> 5] >> (row & 31)) & 1u)) return false;
  return true;
}

__global__ void kept_ts_max_kernel(const double* ts, RowFilter f, int64_t n, unsigned long long* max_bits,
                                   unsigned long long* count, unsigned long long* windows) {
  // a warp visits whole 32-row windows (one mask word): kept rows, their newest timestamp, and the number of windows that
  // keep at least one row (how tile-dense the subset is: the staged sweep loads whole tiles)
  unsigned long long best = 0, cnt = 0, win = 0;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t base = warp0 * 32; base < n; base += nwarps * 32) {
    const int64_t i = base + lane;
    const bool kept = i < n && row_kept(f, i);
    if (kept) {
      ++cnt;
      if (ts) { unsigned long long o = order_f64(ts[i]); best = o > best ? o : best; }
    }
    if (__any_sync(kFull, kept) && lane == 0) ++win;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    unsigned long long b2 = __shfl_xor_sync(kFull, best, o);
    best = b2 > best ? b2 : best;
    cnt += __shfl_xor_sync(kFull, cnt, o);
  }
  if (lane == 0) {
    if (best) atomicMax(max_bits, best);
    if (cnt) atomicAdd(count, cnt);
    if (win && windows) atomicAdd(windows, win);
  }
}

int launch_kept_ts_max(const double* ts, const RowFilter& f, int64_t n, unsigned long long* d_max_bits,
                       unsigned long long* d_count, unsigned long long* d_windows, cudaStream_t s) {
  HDB_CUDA(cudaMemsetAsync(d_max_bits, 0, 8, s));
  HDB_CUDA(cudaMemsetAsync(d_count, 0, 8, s));
  if (d_windows) HDB_CUDA(cudaMemsetAsync(d_windows, 0, 8, s));
  if (n == 0) return 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  kept_ts_max_kernel<<<(unsigned)blocks, 256, 0, s>>>(ts, f, n, d_max_bits, d_count, d_windows);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

__global__ void decay_kernel(const double* ts, double* decay, int64_t n, double ts_max) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    decay[i] = exp(__dadd_rn(-ts_max, ts[i]));                // np.exp(-np.max(ts) + ts), ranking_algorithm.py:183
}
__global__ void stage1_kernel(double* ts, int64_t n, double bias1, double ts_max) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    ts[i] = __dmul_rn(bias1, exp(__dadd_rn(-ts_max, ts[i])));  // hyperdb/hyperdb.py:1344
}
int launch_decay(const double* ts, double* decay, int64_t n, double ts_max, cudaStream_t s) {
  if (n == 0) return 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  decay_kernel<<<(unsigned)blocks, 256, 0, s>>>(ts, decay, n, ts_max);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}
int launch_stage1(double* ts, int64_t n, double bias1, double ts_max, cudaStream_t s) {
  if (n == 0) return 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  stage1_kernel<<<(unsigned)blocks, 256, 0, s>>>(ts, n, bias1, ts_max);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// One CTA per query.  The query's canonical norm (np.linalg.norm in the query's own dtype) is split like
// the candidate re-scoring: squares element-wise by all threads into shared memory, NumPy's pairwise
// reduction by one thread out of shared memory.
template <int QDT>
__device__ double query_unit_norm(const void* q, int64_t d, void* s_sq, int tid, int nthreads, bool staged) {
  using A = Arith<QDT>;
  using C = typename A::C;
  __shared__ double s_result;
  C* sq = reinterpret_cast<C*>(s_sq);
  if (staged) {
    for (int64_t j = tid; j < d; j += nthreads) {
      C v = A::from_double(load_as_double(q, QDT, j));
      sq[j] = A::mul(v, v);
    }
    __syncthreads();
  }
  __shared__ PwScratch s_pw;
  if (staged) {
    if (tid < 32) {
      C n = A::sqrt(pairwise_sum_warp<QDT>([&](int i) { return sq[i]; }, (int)d, tid, &s_pw));
      if (n == 0) n = 1;
      if (tid == 0) s_result = (double)n;
    }
  } else if (tid == 0) {
    C n = canonical_norm<QDT>(q, d);
    if (n == 0) n = 1;
    s_result = (double)n;
  }
  __syncthreads();
  return s_result;
}

// np.mean(q) / np.std(q) of the query in ITS dtype (hyperdb/ranking_algorithm.py:90,93; q is 1-D, so np.mean takes the
// scalar branch).  Staged: values and squared deviations in shared memory, pairwise sums by warp 0.
template <int QDT>
__device__ void query_mean_std(const void* q, int64_t d, void* s_buf, int tid, int nthreads, bool staged, double* mean_out,
                               double* std_out) {
  using A = Arith<QDT>;
  using C = typename A::C;
  __shared__ double s_ms[2];
  __shared__ PwScratch s_pw2;
  if (staged) {
    C* x = reinterpret_cast<C*>(s_buf);
    for (int64_t j = tid; j < d; j += nthreads) x[j] = A::from_double(load_as_double(q, QDT, j));
    __syncthreads();
    if (tid < 32) {
      C mean, sd;
      mean_std_core<QDT>([&](int j) { return x[j]; }, (int)d, true,
                         [&](auto term) { return pairwise_sum_warp<1>(term, (int)d, tid, &s_pw2); },
                         [&](auto term) { return pairwise_sum_warp<QDT>(term, (int)d, tid, &s_pw2); }, &mean, &sd);
      if (tid == 0) { s_ms[0] = (double)mean; s_ms[1] = (double)sd; }
    }
  } else if (tid == 0) {
    C mean, sd;
    canonical_mean_std<QDT>(q, d, true, &mean, &sd);
    s_ms[0] = (double)mean; s_ms[1] = (double)sd;
  }
  __syncthreads();
  *mean_out = s_ms[0];
  *std_out = s_ms[1];
}

__global__ void prep_query_kernel(const void* queries, int qdt, int64_t d, int metric, int sdt, int words,
                                  QueryBuffers qb, int stage, unsigned long long* tau) {
  extern __shared__ __align__(16) unsigned char s_query[];
  __shared__ double s_red[32];
  __shared__ int s_nan;
  const int64_t b = blockIdx.x;
  const char* q = reinterpret_cast<const char*>(queries) + b * d * dtype_size(qdt);
  if (threadIdx.x == 0) {
    s_nan = 0;
    if (tau) tau[b] = 0;               // running threshold of this query's sweep starts below every key
  }
  double nrm = 1.0;
  if (metric == HDB_COSINE) {
    if (qdt == 0) nrm = query_unit_norm<0>(q, d, s_query, threadIdx.x, blockDim.x, stage);
    else if (qdt == 1) nrm = query_unit_norm<1>(q, d, s_query, threadIdx.x, blockDim.x, stage);
    else nrm = query_unit_norm<2>(q, d, s_query, threadIdx.x, blockDim.x, stage);
  }
  double qmean = 0.0, qstd = 1.0;
  if (metric == HDB_PEARSON) {
    if (qdt == 0) query_mean_std<0>(q, d, s_query, threadIdx.x, blockDim.x, stage, &qmean, &qstd);
    else if (qdt == 1) query_mean_std<1>(q, d, s_query, threadIdx.x, blockDim.x, stage, &qmean, &qstd);
    else query_mean_std<2>(q, d, s_query, threadIdx.x, blockDim.x, stage, &qmean, &qstd);
  }
  __syncthreads();
  const bool acc_f64 = (sdt == 2);
  double sq = 0.0, sb = 0.0;
  bool bad = false;
  for (int64_t j = threadIdx.x; j < d; j += blockDim.x) {
    double v = load_as_double(q, qdt, j);
    bad |= (v != v);
    double c = (metric == HDB_COSINE) ? unit_elem(v, nrm, qdt) : (metric == HDB_PEARSON ? sub_in(v, qmean, qdt) : v);
    qb.qc[b * d + j] = c;
    if (acc_f64) reinterpret_cast<double*>(qb.qa)[b * d + j] = c;
    else reinterpret_cast<float*>(qb.qa)[b * d + j] = (float)c;
    sq += c * c;
    sb += acc_f64 ? c : (double)(float)c;          // the sweep's own view of sum_j (q_j - mean)
  }
  if (qb.qbits) {       // sign bits: one coalesced 32-element read + ballot per word
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int w = threadIdx.x >> 5; w < words; w += nwarps) {
      const int64_t j = (int64_t)w * 32 + lane;
      const bool on = j < d && load_as_double(q, qdt, j) > 0.0;
      const unsigned word = __ballot_sync(kFull, on);
      if (lane == 0) qb.qbits[b * words + w] = word;
    }
  }
  sq = warp_sum(sq);
  sb = warp_sum(sb);
  if (__any_sync(kFull, bad) && (threadIdx.x & 31) == 0) atomicOr(&s_nan, 1);
  if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = sq; s_red[16 + (threadIdx.x >> 5)] = sb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0, tot_b = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { tot += s_red[w]; tot_b += s_red[16 + w]; }
    // pearson: ||q - mean||_2 / (std_q sqrt(d)) -- the query's factor of the certificate's magnitude bound
    qb.qnorm[b] = (metric == HDB_PEARSON) ? sqrt(tot) / (qstd * sqrt((double)d)) : sqrt(tot);
    qb.qflags[b] = s_nan ? HDB_FLAG_QUERY_NAN : 0u;
    if (qb.qaux) { qb.qaux[2 * b] = qstd; qb.qaux[2 * b + 1] = tot_b; }
  }
}

int launch_prep_query(const void* q, int q_dtype, int64_t nq, int64_t d, int metric, int sdt, int words,
                      const QueryBuffers& qb, unsigned long long* tau, cudaStream_t s) {
  if (nq == 0) return 0;
  const size_t bytes = (size_t)d * (q_dtype == 2 ? 8 : 4);          // squares (cosine) / values (pearson) in the carrier type
  const int stage = bytes <= 40 * 1024;
  prep_query_kernel<<<(unsigned)nq, 128, stage ? bytes : 0, s>>>(q, q_dtype, d, metric, sdt, words, qb, stage, tau);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Row order: inv = ord^-1, with a permutation check (every target hit exactly once, every value in range).
// ---------------------------------------------------------------------------------------------
__global__ void invert_order_kernel(const uint32_t* ord, uint32_t* inv, int64_t n, int* bad, int phase) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (phase == 0) inv[i] = 0xffffffffu;
    else if (phase == 1) {
      const uint32_t o = ord[i];
      if ((int64_t)o >= n) *bad = 1;
      else inv[o] = (uint32_t)i;
    } else if (inv[i] == 0xffffffffu || ord[inv[i]] != (uint32_t)i) *bad = 1;     // a value was missing (another one repeated)
  }
}
int launch_invert_order(const uint32_t* ord, uint32_t* inv, int64_t n, int* d_bad, cudaStream_t s) {
  HDB_CUDA(cudaMemsetAsync(d_bad, 0, 4, s));
  if (n == 0) return 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  for (int phase = 0; phase < 3; ++phase) {
    invert_order_kernel<<<(unsigned)blocks, 256, 0, s>>>(ord, inv, n, d_bad, phase);
    HDB_LAUNCHED();
  }
  HDB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// 128-bit digest of a query's VALUES (the key of HyperDB's query cache, hyperdb/hyperdb.py:1368-1379: the reference
// keys its LRU on tuple(query.tolist()), i.e. on the float64 values -- value-equal float16/32/64 queries share an
// entry, -0.0 == 0.0).  h0 = sum_j mix(bits_j + (j+1) * C0), h1 = sum_j mix((bits_j ^ C2) + (j+1) * C1) (mod 2^64)
// with bits_j the float64 bit pattern (-0.0 -> +0.0) and mix = splitmix64's finaliser: position-keyed, order-free, so
// one CTA per query reduces it with shuffles.  hyperdb_b200/hyperdb.py::query_digest_host is the NumPy statement.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long digest_mix(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
constexpr unsigned long long kDigestC0 = 0x9e3779b97f4a7c15ull, kDigestC1 = 0xd1b54a32d192ed03ull, kDigestC2 = 0xa0761d6478bd642full;

__global__ void query_digest_kernel(const void* queries, int qdt, int64_t d, unsigned long long* out) {
  __shared__ unsigned long long s_h[2][8];
  const int64_t b = blockIdx.x;
  const char* q = reinterpret_cast<const char*>(queries) + b * d * dtype_size(qdt);
  unsigned long long h0 = 0, h1 = 0;
  for (int64_t j = threadIdx.x; j < d; j += blockDim.x) {
    double v = load_as_double(q, qdt, j);
    if (v == 0.0) v = 0.0;                                       // -0.0 and +0.0 are the same tuple element
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    h0 += digest_mix(bits + (unsigned long long)(j + 1) * kDigestC0);
    h1 += digest_mix((bits ^ kDigestC2) + (unsigned long long)(j + 1) * kDigestC1);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { h0 += __shfl_xor_sync(kFull, h0, o); h1 += __shfl_xor_sync(kFull, h1, o); }
  if ((threadIdx.x & 31) == 0) { s_h[0][threadIdx.x >> 5] = h0; s_h[1][threadIdx.x >> 5] = h1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t0 = 0, t1 = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t0 += s_h[0][w]; t1 += s_h[1][w]; }
    out[2 * b] = t0; out[2 * b + 1] = t1;
  }
}

int launch_query_digest(const void* q, int q_dtype, int64_t nq, int64_t d, unsigned long long* out, cudaStream_t s) {
  if (nq == 0) return 0;
  query_digest_kernel<<<(unsigned)nq, 256, 0, s>>>(q, q_dtype, d, out);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

void query_digest_host(const void* q, int q_dtype, int64_t nq, int64_t d, unsigned long long* out) {
  for (int64_t b = 0; b < nq; ++b) {
    unsigned long long h0 = 0, h1 = 0;
    for (int64_t j = 0; j < d; ++j) {
      const int64_t i = b * d + j;
      double v = q_dtype == 0 ? (double)__half2float(reinterpret_cast<const __half*>(q)[i])
                              : (q_dtype == 1 ? (double)reinterpret_cast<const float*>(q)[i] : reinterpret_cast<const double*>(q)[i]);
      if (v == 0.0) v = 0.0;
      unsigned long long bits;
      memcpy(&bits, &v, 8);
      h0 += digest_mix(bits + (unsigned long long)(j + 1) * kDigestC0);
      h1 += digest_mix((bits ^ kDigestC2) + (unsigned long long)(j + 1) * kDigestC1);
    }
    out[2 * b] = h0; out[2 * b + 1] = h1;
  }
}

// ---------------------------------------------------------------------------------------------
// Row removal (hyperdb/hyperdb.py:718-728: np.vstack / boolean-mask copy of the whole matrix per remove_document):
// stable in-place compaction of the matrix and of every per-row column.
//   mark_removed:  keep[i] = 0 for every listed row          (keep[] starts as all ones)
//   exclusive scan of keep[] -> destination of every kept row; build_src_index inverts it: src[dest] = row
//   gather_rows:   bounce[j - j0] = column[src[j]] for one chunk of destinations [j0, j0+cnt); the caller then copies
//                  the bounce buffer to column[j0 ...].  src[j] >= j, so a chunk only overwrites rows that later
//                  chunks never read.
__global__ void mark_removed_kernel(uint32_t* keep, int64_t n, const int64_t* rows, int64_t count, int* err) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = rows[i];
    if (r < 0 || r >= n) atomicOr(err, 1);
    else keep[r] = 0u;                                   // duplicates write the same value
  }
}
__global__ void fill_u32_kernel(uint32_t* p, int64_t n, uint32_t v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void build_src_index_kernel(const uint32_t* keep, const uint32_t* pos, int64_t n, uint32_t* src, unsigned long long* n_new) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (keep[i]) src[pos[i]] = (uint32_t)i;
    if (i == n - 1) *n_new = (unsigned long long)pos[i] + keep[i];
  }
}
// One warp per destination row; VEC = 16-byte pieces, otherwise 2-byte pieces (every stored dtype is >= 2 bytes wide).
template <bool VEC>
__global__ void gather_rows_kernel(const char* col, char* bounce, int64_t row_bytes, const uint32_t* src, int64_t j0, int64_t cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t j = warp; j < cnt; j += nwarps) {
    const char* from = col + (int64_t)src[j0 + j] * row_bytes;
    char* to = bounce + j * row_bytes;
    if (VEC) {
      for (int64_t c = lane; c < row_bytes / 16; c += 32) reinterpret_cast<uint4*>(to)[c] = reinterpret_cast<const uint4*>(from)[c];
    } else {
      for (int64_t c = lane; c < row_bytes / 2; c += 32) reinterpret_cast<uint16_t*>(to)[c] = reinterpret_cast<const uint16_t*>(from)[c];
    }
  }
}

int launch_plan_removal(int64_t n, const int64_t* d_rows, int64_t count, uint32_t* keep, uint32_t* pos, uint32_t* src,
                        unsigned long long* d_n_new, int* d_err, void** scan_tmp, size_t* scan_tmp_bytes, cudaStream_t s) {
  if (n == 0) return 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  fill_u32_kernel<<<(unsigned)blocks, 256, 0, s>>>(keep, n, 1u);
  HDB_LAUNCHED();
  if (count > 0) {
    int64_t b2 = (count + 255) / 256;
    if (b2 > 148 * 16) b2 = 148 * 16;
    mark_removed_kernel<<<(unsigned)b2, 256, 0, s>>>(keep, n, d_rows, count, d_err);
    HDB_LAUNCHED();
  }
  size_t need = 0;
  HDB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, keep, pos, (int)n, s));
  if (*scan_tmp_bytes < need) {
    if (*scan_tmp) HDB_CUDA(cudaFree(*scan_tmp));
    *scan_tmp = nullptr; *scan_tmp_bytes = 0;
    HDB_CUDA(cudaMalloc(scan_tmp, need));
    *scan_tmp_bytes = need;
  }
  HDB_CUDA(cub::DeviceScan::ExclusiveSum(*scan_tmp, need, keep, pos, (int)n, s));
  HDB_LAUNCHED();
  build_src_index_kernel<<<(unsigned)blocks, 256, 0, s>>>(keep, pos, n, src, d_n_new);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

// Compacts one per-row column in place: column[j] = column[src[j]] for j in [0, n_new), chunk by chunk through `bounce`.
int launch_compact_column(void* column, int64_t row_bytes, int64_t n_new, const uint32_t* src, void* bounce, size_t bounce_bytes,
                          cudaStream_t s) {
  if (!column || n_new == 0 || row_bytes == 0) return 0;
  int64_t chunk = (int64_t)(bounce_bytes / (size_t)row_bytes);
  if (chunk < 1) return fail("compact: bounce buffer smaller than one row");
  const bool vec = (row_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(column) & 15) == 0) && ((reinterpret_cast<uintptr_t>(bounce) & 15) == 0);
  for (int64_t j0 = 0; j0 < n_new; j0 += chunk) {
    const int64_t cnt = n_new - j0 < chunk ? n_new - j0 : chunk;
    int64_t blocks = (cnt * 32 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (vec) gather_rows_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const char*>(column), reinterpret_cast<char*>(bounce), row_bytes, src, j0, cnt);
    else gather_rows_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const char*>(column), reinterpret_cast<char*>(bounce), row_bytes, src, j0, cnt);
    HDB_LAUNCHED();
    HDB_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(column) + j0 * row_bytes, bounce, (size_t)(cnt * row_bytes), cudaMemcpyDeviceToDevice, s));
  }
  HDB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// get_norm_vector as a function (hdb_normalize_rows): thread per row.
template <int T>
__global__ void normalize_rows_kernel(const void* src, void* dst, int64_t n, int64_t d) {
  using C = typename Arith<T>::C;
  int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= n) return;
  const char* base = reinterpret_cast<const char*>(src) + row * d * dtype_size(T);
  C cn = canonical_norm<T>(base, d);
  if (cn == C(0)) cn = C(1);
  for (int64_t j = 0; j < d; ++j) {
    double u = unit_elem(load_as_double(base, T, j), (double)cn, T);
    if (T == 0) reinterpret_cast<__half*>(dst)[row * d + j] = __float2half_rn((float)u);
    else if (T == 1) reinterpret_cast<float*>(dst)[row * d + j] = (float)u;
    else reinterpret_cast<double*>(dst)[row * d + j] = u;
  }
}
int launch_normalize_rows(int dtype, int64_t n, int64_t d, const void* src, void* dst, cudaStream_t s) {
  if (n == 0) return 0;
  int handled = 0;
  HDB_TRY(launch_normalize_rows_warp(dtype, n, d, src, dst, s, &handled));
  if (handled) return 0;
  int64_t blocks = (n + 127) / 128;
  if (dtype == 0) normalize_rows_kernel<0><<<(unsigned)blocks, 128, 0, s>>>(src, dst, n, d);
  else if (dtype == 1) normalize_rows_kernel<1><<<(unsigned)blocks, 128, 0, s>>>(src, dst, n, d);
  else normalize_rows_kernel<2><<<(unsigned)blocks, 128, 0, s>>>(src, dst, n, d);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hdb
