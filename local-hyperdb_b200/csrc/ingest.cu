// Ingest-time and query-preparation kernels (one-off / tiny; not on the streaming path).
//
//  row_stats     get_norm_vector's per-row norm (hyperdb/ranking_algorithm.py:8-21) computed ONCE with
//                NumPy's exact arithmetic, instead of once per cosine query over the whole matrix;
//                the NaN scan of :150; the two statistics of the certification bound.
//  pack_bits     check_and_binarize_vectors (:116-126): bit = (x > 0), 32 columns per word.
//  kept_ts_max / decay / stage1    the time-decay column of :179-183 and hyperdb.py:1334-1346.
//  prep_query    canonical query (unit query for cosine, normalised in the QUERY's dtype, :38),
//                accumulate-type copy for the sweep, sign bits, ||q||, NaN flag.
#include "canonical.cuh"
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
                                   unsigned long long* count) {
  unsigned long long best = 0, cnt = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (row_kept(f, i)) {
      ++cnt;
      if (ts) { unsigned long long o = order_f64(ts[i]); best = o > best ? o : best; }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    unsigned long long b2 = __shfl_xor_sync(kFull, best, o);
    best = b2 > best ? b2 : best;
    cnt += __shfl_xor_sync(kFull, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (best) atomicMax(max_bits, best);
    if (cnt) atomicAdd(count, cnt);
  }
}

int launch_kept_ts_max(const double* ts, const RowFilter& f, int64_t n, unsigned long long* d_max_bits,
                       unsigned long long* d_count, cudaStream_t s) {
  HDB_CUDA(cudaMemsetAsync(d_max_bits, 0, 8, s));
  HDB_CUDA(cudaMemsetAsync(d_count, 0, 8, s));
  if (n == 0) return 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  kept_ts_max_kernel<<<(unsigned)blocks, 256, 0, s>>>(ts, f, n, d_max_bits, d_count);
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
  __syncthreads();
  const bool acc_f64 = (sdt == 2);
  double sq = 0.0;
  bool bad = false;
  for (int64_t j = threadIdx.x; j < d; j += blockDim.x) {
    double v = load_as_double(q, qdt, j);
    bad |= (v != v);
    double c = (metric == HDB_COSINE) ? unit_elem(v, nrm, qdt) : v;
    qb.qc[b * d + j] = c;
    if (acc_f64) reinterpret_cast<double*>(qb.qa)[b * d + j] = c;
    else reinterpret_cast<float*>(qb.qa)[b * d + j] = (float)c;
    sq += c * c;
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
  if (__any_sync(kFull, bad) && (threadIdx.x & 31) == 0) atomicOr(&s_nan, 1);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tot += s_red[w];
    qb.qnorm[b] = sqrt(tot);
    qb.qflags[b] = s_nan ? HDB_FLAG_QUERY_NAN : 0u;
  }
}

int launch_prep_query(const void* q, int q_dtype, int64_t nq, int64_t d, int metric, int sdt, int words,
                      const QueryBuffers& qb, unsigned long long* tau, cudaStream_t s) {
  if (nq == 0) return 0;
  const size_t bytes = (size_t)d * (q_dtype == 2 ? 8 : 4);          // squares in the carrier type
  const int stage = bytes <= 40 * 1024;
  prep_query_kernel<<<(unsigned)nq, 128, stage ? bytes : 0, s>>>(q, q_dtype, d, metric, sdt, words, qb, stage, tau);
  HDB_LAUNCHED();
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
  int64_t blocks = (n + 127) / 128;
  if (dtype == 0) normalize_rows_kernel<0><<<(unsigned)blocks, 128, 0, s>>>(src, dst, n, d);
  else if (dtype == 1) normalize_rows_kernel<1><<<(unsigned)blocks, 128, 0, s>>>(src, dst, n, d);
  else normalize_rows_kernel<2><<<(unsigned)blocks, 128, 0, s>>>(src, dst, n, d);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hdb
