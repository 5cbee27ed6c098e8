// Internal host-side declarations shared by the translation units of libhyperdb_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <atomic>
#include <string>

namespace hdb {

extern thread_local std::string g_error;
extern std::atomic<int64_t> g_launches;     // instrumentation only; handles may live on different host threads
int fail(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

#define HDB_CUDA(x)                                             \
  do {                                                          \
    cudaError_t e__ = (x);                                      \
    if (e__ != cudaSuccess) return ::hdb::cuda_fail(e__, #x);   \
  } while (0)
#define HDB_TRY(x)            \
  do {                        \
    int rc__ = (x);           \
    if (rc__ != 0) return rc__; \
  } while (0)
#define HDB_LAUNCHED() (::hdb::g_launches.fetch_add(1, std::memory_order_relaxed))

constexpr int kMaxKP = 128;            // largest candidate-list class of the fused pass
constexpr int kSweepThreads = 256;     // 8 warps per CTA
constexpr int kSweepWarps = kSweepThreads / 32;

// Row subset + decay, as consumed by every scoring kernel.
struct RowFilter {
  const uint32_t* mask;   // 1 bit per local row, or nullptr
  int64_t lo, hi;         // kept range [lo, hi)
  const double* decay;    // exp(ts - max ts) per row, or nullptr
  double bias;            // recency_bias
  // Row order (hdb_matrix_set_row_order): the rows are STORED in another order than the caller numbers them (e.g.
  // clustered by a metadata key so that a filter keeps a few contiguous runs).  ord[p] = the caller's local index of
  // physical row p, inv = its inverse.  Selection keys, ties (lower index first) and reported ids use ord; mask, range,
  // timestamps and every other per-row column are in PHYSICAL order.  nullptr = identity.
  const uint32_t* ord = nullptr;
  const uint32_t* inv = nullptr;
  // host-side hint: 1 = the subset keeps (nearly) whole 32-row windows -- no mask, or a clustered store -- so kernels that
  // load whole tiles waste nothing; 0 = a sparse mask (rows are skipped one by one)
  int tile_dense = 1;
};

// ---- peer-memory exchange (exchange.cu): what a pushing kernel needs to know
constexpr int kXSlots = 4;                // ring of message slots per (reader, source)
enum { kCtrPushStep = 0, kCtrWaitStep = 1, kCtrPushDone = 2, kCtrWaitDone = 3, kCtrCount = 4 };
struct PushTarget {
  char* const* peer;        // device array [world]: every rank's exchange buffer as mapped here (nullptr = no exchange)
  int world, rank;
  int64_t max_words;        // 8-byte words per (slot, source) message
  size_t data_bytes;        // offset of the arrival flags [kXSlots][world]
  size_t consumed_off;      // offset of the consumed counters [world]
  unsigned long long* ctr;  // [kCtr*] device counters of this rank
  const char* local;        // this rank's own buffer
  int* error;
  int64_t nq, k;            // message geometry: scores nq*k (f64) | ids nq*k (i64) | counts nq (i64) | flags nq (u32)
};

struct QueryBuffers {     // per-batch device buffers written by prep_query
  void* qa;               // [B][d] query in the sweep's accumulate type (float, or double for f64 storage)
  double* qc;             // [B][d] canonical query values (dtype R widened to double)
  uint32_t* qbits;        // [B][words] sign bits
  double* qnorm;          // [B] ||qc||_2
  uint32_t* qflags;       // [B] HDB_FLAG_QUERY_NAN
  double* qaux;           // [B][2] pearson: np.std(q), sum_j (q_j - np.mean(q))
};

struct MatrixView {
  const void* rows;
  int dtype;
  int64_t n, d;
  int64_t row_offset;
  const void* norms;      // canonical norms: float (f16/f32 storage) or double (f64), zero -> 1
  const void* inv_norms;  // 1/norm in the accumulate type
  const float* sqnorms;   // ||v||^2 (float) for f16/f32 storage, or nullptr
  const uint32_t* bits;   // packed sign bits or nullptr
  int words;              // words per packed row
  float max_norm;         // max_i ||v_i||_2
  float max_ratio;        // max_i ||v_i||_2 / canonical norm_i
  // pearson columns (built on the first pearson query): np.mean / np.std per row in the storage dtype's carrier
  // (float for f16/f32, double for f64), and 1/(std*d) in the sweep's accumulate type (NaN for a constant row)
  const void* pmean;
  const void* pstd;
  const void* pscale;
  float max_pratio;       // max_i ||v_i||_2 / (std_i * sqrt(d)) over non-constant rows (certificate: the sweep's error scale)
  float max_cratio;       // max_i ||v_i - mean_i||_2 / (std_i * sqrt(d)), ~1 (certificate: the reference's error scale)
  float min_pstd;         // min_i std_i over non-constant rows
};

// ---- ingest.cu
int launch_row_stats(const MatrixView& m, void* norms, void* inv_norms, float* sqnorms, float* d_stats /*[2] max_norm,max_ratio*/,
                     int* d_nan, cudaStream_t s);
int launch_pack_bits(const MatrixView& m, uint32_t* bits, int words, cudaStream_t s);
int launch_pearson_stats(const MatrixView& m, void* pmean, void* pstd, void* pscale, float* d_stats /*[3] max_pratio, -min_pstd (ordered), max_cratio*/,
                         cudaStream_t s);
int launch_plan_removal(int64_t n, const int64_t* d_rows, int64_t count, uint32_t* keep /*[n]*/, uint32_t* pos /*[n]*/, uint32_t* src /*[n]*/,
                        unsigned long long* d_n_new, int* d_err, void** scan_tmp, size_t* scan_tmp_bytes, cudaStream_t s);
int launch_compact_column(void* column, int64_t row_bytes, int64_t n_new, const uint32_t* src, void* bounce, size_t bounce_bytes,
                          cudaStream_t s);
int launch_kept_ts_max(const double* ts, const RowFilter& f, int64_t n, unsigned long long* d_max_bits,
                       unsigned long long* d_count, unsigned long long* d_windows /*may be null*/, cudaStream_t s);
int launch_decay(const double* ts, double* decay, int64_t n, double ts_max, cudaStream_t s);
int launch_stage1(double* ts, int64_t n, double bias1, double ts_max, cudaStream_t s);
int launch_prep_query(const void* q, int q_dtype, int64_t nq, int64_t d, int metric, int sdt, int words,
                      const QueryBuffers& qb, unsigned long long* tau /*zeroed per query, or nullptr*/, cudaStream_t s);
int launch_invert_order(const uint32_t* ord, uint32_t* inv, int64_t n, int* d_bad, cudaStream_t s);
// 128-bit value digest of each query (HyperDB's cache key): device kernel and its host twin
int launch_query_digest(const void* q, int q_dtype, int64_t nq, int64_t d, unsigned long long* out, cudaStream_t s);
void query_digest_host(const void* q, int q_dtype, int64_t nq, int64_t d, unsigned long long* out);
int launch_normalize_rows(int dtype, int64_t n, int64_t d, const void* src, void* dst, cudaStream_t s);
double decode_ordered_double(unsigned long long bits);

// ---- sweep.cu : fused score + select pass, 1 / 2 / 4 / 8 queries per launch (one read of the matrix)
struct SweepOut {
  uint64_t* cand;          // [nq][grid][KP] per-CTA candidate keys (descending) of the first query of the pass
  unsigned long long* tau; // [nq] global running thresholds (must be zeroed before the launch)
  int grid;
};
int sweep_grid_size(int device);
// largest query group (1, 2, 4 or 8) one pass can take for this shape / metric / candidate class
int sweep_max_group(const MatrixView& m, int metric, int kp);
// qa / qbits / qaux point at the FIRST query of the group; consecutive queries are d accumulate-type elements, `words`
// sign-bit words and 2 doubles apart (the layout prep_query writes)
int launch_sweep(const MatrixView& m, int metric, int rdt /*result dtype = max(storage, query)*/, const void* qa, const uint32_t* qbits,
                 const double* qaux /*pearson*/, const RowFilter& f, int kp, const SweepOut& out, int nq, cudaStream_t s);

// ---- finalize.cu : merge + canonical re-score + certification; exact full-vector path
struct FinalizeArgs {
  MatrixView m;
  RowFilter f;
  int metric, rdt;               // result dtype R
  int kp, k;                     // candidate class, requested k (already clamped to >= 0)
  int64_t n_kept;
  int grid;                      // CTAs of the sweep
  const uint64_t* cand;          // [B][grid][kp]
  const unsigned long long* tau; // [B]
  QueryBuffers qb;
  int64_t* out_idx; double* out_score; int64_t* out_count; uint32_t* out_flags;   // device
  int* uncertified;              // device counter: queries that need the exact path
  int smem_bytes;                // dynamic shared memory of the finalize kernel (set by launch_finalize)
  // tensor-core batched path: per-query candidate buffers instead of per-CTA lists
  const unsigned* cand_count;    // [B] appended keys per query (may exceed cand_stride = overflow) or nullptr
  int64_t cand_stride;           // keys per query buffer
  const float* tau0;             // [B] select threshold of the batched pass (rows below it were never appended)
  int tau0_negd2;                // tau0 is -distance^2 (batched euclidean): convert to a similarity before use
  uint32_t extra_flags;          // OR-ed into out_flags (HDB_FLAG_TENSOR)
  PushTarget push;               // push.peer != nullptr: also store the results into every rank's exchange slot (fused push)
};
int launch_finalize(const FinalizeArgs& a, int64_t nq, cudaStream_t s);
int launch_full_scores(const MatrixView& m, const RowFilter& f, int metric, int rdt, const double* qc,
                       const uint32_t* qbits, const double* qaux, double* totals /*[n], masked rows = -NaN*/, cudaStream_t s);
int launch_scores_out(const MatrixView& m, int metric, int rdt, const double* qc, const uint32_t* qbits, const double* qaux,
                      void* out /*dtype R, uint64 (hamming) or float64 (jaccard, pearson)*/, int distance /*euclidean: the distance itself*/,
                      cudaStream_t s);
int exact_topk(int device, const double* totals, const uint32_t* inv, int64_t n, int64_t row_offset, int64_t k, int64_t n_kept,
               int64_t* out_idx, double* out_score, int64_t* out_count, void** scratch, size_t* scratch_bytes,
               cudaStream_t s);
int launch_merge_topk(int64_t n_lists, int64_t nq, int64_t k, int64_t ls_rec, int64_t ls_cnt, const double* scores, const int64_t* ids,
                      const int64_t* counts, int64_t* out_idx, double* out_score, int64_t* out_count, cudaStream_t s);

// ---- rowwise.cu : the same passes at the HBM rate (warp per row / lane-per-row chains); *handled = 0 -> not covered, use the kernels above
int launch_row_stats_warp(const MatrixView& m, void* norms, void* inv_norms, float* sqnorms, float* d_stats, int* d_nan, cudaStream_t s,
                          int* handled);
int launch_pearson_stats_warp(const MatrixView& m, void* pmean, void* pstd, void* pscale, float* d_stats, cudaStream_t s, int* handled);
int launch_normalize_rows_warp(int dtype, int64_t n, int64_t d, const void* src, void* dst, cudaStream_t s, int* handled);
int launch_scores_rowwise(const MatrixView& m, const RowFilter& f, int metric, int rdt, const double* qc, const double* qaux,
                          double* totals, void* typed, int distance, cudaStream_t s, int* handled);

// ---- batched_tc.cu : tcgen05 batched contraction + threshold select
struct TcWorkspace {
  __half* q16;             // [nq][d] fp16 queries (B operand)
  float* dense;            // [nq][sample_tiles*128] sample totals
  float* tau0;             // [nq]
  uint64_t* cand;          // [nq][cap]
  unsigned* cand_count;    // [nq]
  float* qsq;              // [nq] ||q||^2 (euclidean) or 0
  int cap;
  int64_t sample_tiles;
  void* rec;               // [n_sm][rec_cap] 16-byte CTA-private candidate records
  unsigned* rec_count;     // [n_sm]
  unsigned rec_cap;
  int force_single;        // testing: keep the single-CTA contraction even for wide batches
};
int batched_tc_supported(const MatrixView& m, int metric, int q_dtype, int64_t nq, bool has_decay);
int launch_batched_tc(const MatrixView& m, int metric, const RowFilter& f, const float* qa, const double* qnorm, int64_t nq, int kp, int device,
                      const TcWorkspace& ws, cudaStream_t s);

}  // namespace hdb

// ---- exchange.cu (host side, used by api.cu)
struct hdb_exchange;
namespace hdb {
int exchange_push_target(hdb_exchange* x, int64_t nq, int64_t k, PushTarget* out);
int64_t exchange_max_words(const hdb_exchange* x);
int exchange_device(const hdb_exchange* x);
cudaStream_t exchange_stream(const hdb_exchange* x);
int exchange_launch_push(hdb_exchange* x, cudaStream_t s, const void* mine, int64_t words);
int exchange_launch_wait_merge(hdb_exchange* x, cudaStream_t s, int64_t nq, int64_t k, int64_t* out_idx, double* out_score,
                               int64_t* out_count, uint32_t* out_flags);
}  // namespace hdb
