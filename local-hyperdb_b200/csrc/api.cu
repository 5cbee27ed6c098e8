// C ABI of libhyperdb_b200.so (declared in include/hyperdb_b200.h): handle lifetime, workspaces,
// path selection (fused sweep -> certify, tensor-core batched contraction, exact full-vector path).
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "hdb_common.cuh"
#include "hdb_internal.h"
#include "../../include/hyperdb_b200.h"

namespace hdb {
thread_local std::string g_error;
std::atomic<int64_t> g_launches{0};
int fail(const std::string& msg) { g_error = msg; return 1; }
int cuda_fail(cudaError_t e, const char* what) {
  g_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " at " + what;
  return 2;
}
}  // namespace hdb

using namespace hdb;

constexpr int64_t kChunk = 64;     // queries per fused batch chunk (bounds the candidate workspace)

struct hdb_matrix {
  int device = 0, dtype = 0;
  int64_t n = 0, d = 0, row_offset = 0;
  int64_t cap = 0;                  // row capacity of `rows` (when owned) and of every per-row column (>= n)
  void* rows = nullptr;
  bool owns_rows = false;
  void* norms = nullptr;
  void* inv_norms = nullptr;
  float* sqnorms = nullptr;
  uint32_t* bits = nullptr;
  int words = 0;
  float max_norm = 0.f, max_ratio = 1.f;
  void* pmean = nullptr; void* pstd = nullptr; void* pscale = nullptr;    // pearson columns, built on first use
  float max_pratio = 0.f, max_cratio = 0.f, min_pstd = 0.f;
  bool finalized = false;
  uint32_t* mask = nullptr;
  uint32_t* ord = nullptr; uint32_t* inv = nullptr;     // row order (hdb_matrix_set_row_order), both [n] or both null
  int64_t lo = 0, hi = 0, n_kept = 0;
  int64_t kept_windows = 0;         // 32-row windows with at least one kept row
  double* ts = nullptr;
  double* decay = nullptr;
  bool decay_valid = false;
  cudaStream_t stream = nullptr;
  int path_mode = 0;
  int max_group = 0;                // testing / A-B: cap on the queries per sweep pass (0 = as many as the shape allows)
  int grid = 0;
  // workspaces
  int64_t ws_q = 0;                 // query capacity of the buffers below
  void* q_raw = nullptr;            // staged raw queries (f64 worst case)
  QueryBuffers qb{};
  uint64_t* cand = nullptr;         // [kChunk][grid][kMaxKP]
  unsigned long long* tau = nullptr;
  int* uncertified = nullptr;
  // host-output staging: one device block [idx | score | count | flags] mirrored by one pinned host block
  char* o_block = nullptr; char* h_block = nullptr; size_t o_block_bytes = 0;
  int64_t* o_idx = nullptr; double* o_score = nullptr; int64_t* o_count = nullptr; uint32_t* o_flags = nullptr;
  double* totals = nullptr;
  void* sort_scratch = nullptr; size_t sort_scratch_bytes = 0;
  void* scan_tmp = nullptr; size_t scan_tmp_bytes = 0;      // row removal
  unsigned long long* misc = nullptr;   // [2] ordered max bits, count
  unsigned long long* digest = nullptr; int64_t digest_cap = 0;   // hdb_query_digest: 2 words per query
  float* stats = nullptr; int* nan_flag = nullptr;
  // query pipelining (hdb_matrix_set_post_stream): the per-query workspaces exist twice so that the certify step of
  // query i (post stream) can overlap the sweep of query i+1 (main stream)
  struct QuerySlot {
    int64_t ws_q = 0; void* q_raw = nullptr; QueryBuffers qb{}; uint64_t* cand = nullptr; unsigned long long* tau = nullptr;
    cudaEvent_t done = nullptr; bool pending = false;
  } slots[2];
  int cur_slot = 0;
  cudaStream_t post_stream = nullptr;
  cudaStream_t pre_stream = nullptr;     // pipelined mode: query preparation runs here, ahead of the main stream
  cudaEvent_t ev_select = nullptr, ev_prep = nullptr, ev_qready = nullptr;
  // sweep overlap: the sweeps of odd workspace slots run on a second internal stream, so the CTAs of query i+1 fill the
  // SMs that query i's persistent CTAs vacate during its tail (launch gap, ramp, straggler warps, per-CTA merge)
  cudaStream_t alt_stream = nullptr;
  cudaEvent_t ev_alt = nullptr;
  bool alt_pending = false;
  int overlap = 1;
  // tensor-core batched path workspace
  TcWorkspace tc{};
  int64_t tc_nq = 0;
  // per-launch event pairs around the dominant kernel (hdb_profile_*)
  std::vector<cudaEvent_t> prof_ev;
  size_t prof_used = 0;
  // row-sharded path: device-output queries also push their results to every rank of this exchange
  hdb_exchange* xchg = nullptr;
  // asynchronous host API (hdb_query_submit / hdb_query_collect): a ring of tickets, each with its device result
  // blocks, a pinned host mirror and a completion event
  struct Ticket {
    bool busy = false;
    int64_t nq = 0, k = 0;
    int world = 1;
    char* d_mine = nullptr; size_t mine_bytes = 0;     // local packed result [scores | ids | counts | flags]
    char* d_res = nullptr; char* h_res = nullptr; size_t res_bytes = 0;   // [idx | score | count | flags world*nq]
    cudaEvent_t ev = nullptr;
  };
  static constexpr int kTickets = 4;
  Ticket tickets[kTickets];
  int64_t next_ticket = 0;
  // last query (for hdb_time_last_query)
  struct { bool valid = false; int metric = 0, rdt = 0, kp = 0; int64_t nq = 0, k = 0; double bias = 0; } last;
};

static void slot_store(hdb_matrix* m) {
  hdb_matrix::QuerySlot& q = m->slots[m->cur_slot];
  q.ws_q = m->ws_q; q.q_raw = m->q_raw; q.qb = m->qb; q.cand = m->cand; q.tau = m->tau;
}
static void slot_load(hdb_matrix* m, int s) {
  m->cur_slot = s;
  const hdb_matrix::QuerySlot& q = m->slots[s];
  m->ws_q = q.ws_q; m->q_raw = q.q_raw; m->qb = q.qb; m->cand = q.cand; m->tau = q.tau;
}

static MatrixView view_of(const hdb_matrix* m) {
  MatrixView v;
  v.rows = m->rows; v.dtype = m->dtype; v.n = m->n; v.d = m->d; v.row_offset = m->row_offset;
  v.norms = m->norms; v.inv_norms = m->inv_norms; v.sqnorms = m->sqnorms; v.bits = m->bits; v.words = m->words;
  v.max_norm = m->max_norm; v.max_ratio = m->max_ratio;
  v.pmean = m->pmean; v.pstd = m->pstd; v.pscale = m->pscale; v.max_pratio = m->max_pratio; v.max_cratio = m->max_cratio; v.min_pstd = m->min_pstd;
  return v;
}
static int quiesce(hdb_matrix* m);
static RowFilter filter_of(const hdb_matrix* m, double bias, bool use_decay) {
  RowFilter f;
  f.mask = m->mask; f.lo = m->lo; f.hi = m->hi;
  f.decay = use_decay ? m->decay : nullptr;
  f.bias = bias;
  f.ord = m->ord; f.inv = m->inv;
  // at least 3 of 4 rows of the windows that are visited at all are kept
  f.tile_dense = (m->mask == nullptr) || (m->kept_windows > 0 && m->n_kept * 4 >= m->kept_windows * 32 * 3);
  return f;
}

template <typename T>
static int dev_alloc(T** p, size_t count) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  if (count == 0) count = 1;
  HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
  return 0;
}

static size_t cap_rows(const hdb_matrix* m) {
  const int64_t c = m->cap > m->n ? m->cap : m->n;
  return (size_t)(c > 0 ? c : 1);
}

// order the handle's stream after the last sweep that was launched on the alternate stream
static int join_alt(hdb_matrix* m) {
  if (m->alt_pending) {
    HDB_CUDA(cudaStreamWaitEvent(m->stream, m->ev_alt, 0));
    m->alt_pending = false;
  }
  return 0;
}

static int refresh_kept(hdb_matrix* m) {
  HDB_TRY(join_alt(m));
  RowFilter f = filter_of(m, 0.0, false);
  HDB_TRY(launch_kept_ts_max(nullptr, f, m->n, m->misc, m->misc + 1, m->misc + 2, m->stream));
  unsigned long long host[3];
  HDB_CUDA(cudaMemcpyAsync(host, m->misc, 24, cudaMemcpyDeviceToHost, m->stream));
  HDB_CUDA(cudaStreamSynchronize(m->stream));
  m->n_kept = (int64_t)host[1];
  m->kept_windows = (int64_t)host[2];
  return 0;
}

extern "C" {

const char* hdb_last_error(void) { return g_error.c_str(); }
int hdb_version(void) { return 100; }
int hdb_device_count(int* count) {
  HDB_CUDA(cudaGetDeviceCount(count));
  return 0;
}
int64_t hdb_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

int hdb_matrix_create(int device, int dtype, int64_t n_rows, int64_t dim, int64_t row_offset, hdb_matrix** out) {
  if (!out) return fail("hdb_matrix_create: out is NULL");
  if (dtype < 0 || dtype > 2) return fail("hdb_matrix_create: dtype must be HDB_F16/F32/F64");
  if (n_rows < 0 || dim <= 0) return fail("hdb_matrix_create: need n_rows >= 0 and dim > 0");
  int ndev = 0;
  HDB_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail("hdb_matrix_create: no such CUDA device");
  HDB_CUDA(cudaSetDevice(device));
  hdb_matrix* m = new hdb_matrix();
  m->device = device; m->dtype = dtype; m->n = n_rows; m->d = dim; m->row_offset = row_offset;
  m->cap = n_rows;
  m->lo = 0; m->hi = n_rows; m->n_kept = n_rows;
  m->grid = sweep_grid_size(device);
  int rc = dev_alloc(&m->misc, 4);
  if (!rc) rc = dev_alloc(&m->stats, 4);
  if (!rc) rc = dev_alloc(&m->nan_flag, 1);
  if (!rc) rc = dev_alloc(&m->uncertified, 1);
  if (rc) { delete m; return rc; }
  *out = m;
  return 0;
}

int hdb_matrix_destroy(hdb_matrix* m) {
  if (!m) return 0;
  cudaSetDevice(m->device);
  cudaStreamSynchronize(m->stream);
  // pipelined queries may still be running on the internal / post streams: nothing is freed under them
  if (m->pre_stream) cudaStreamSynchronize(m->pre_stream);
  if (m->alt_stream) cudaStreamSynchronize(m->alt_stream);
  if (m->post_stream) cudaStreamSynchronize(m->post_stream);
  if (m->owns_rows) cudaFree(m->rows);
  if (m->ord) cudaFree(m->ord);
  if (m->inv) cudaFree(m->inv);
  slot_store(m);
  {
    hdb_matrix::QuerySlot& other = m->slots[m->cur_slot ^ 1];
    void* extra[] = {other.q_raw, other.qb.qa, other.qb.qc, other.qb.qbits, other.qb.qnorm, other.qb.qflags, other.qb.qaux, other.cand, other.tau};
    for (void* p : extra) if (p) cudaFree(p);
    for (auto& q : m->slots) if (q.done) cudaEventDestroy(q.done);
    if (m->ev_select) cudaEventDestroy(m->ev_select);
    if (m->ev_prep) cudaEventDestroy(m->ev_prep);
    if (m->ev_qready) cudaEventDestroy(m->ev_qready);
    if (m->pre_stream) cudaStreamDestroy(m->pre_stream);
    if (m->alt_stream) { cudaStreamSynchronize(m->alt_stream); cudaStreamDestroy(m->alt_stream); }
    if (m->ev_alt) cudaEventDestroy(m->ev_alt);
  }
  void* ptrs[] = {m->norms, m->inv_norms, m->sqnorms, m->bits, m->pmean, m->pstd, m->pscale, m->mask, m->ts, m->decay, m->q_raw, m->qb.qa, m->qb.qc, m->qb.qbits,
                  m->qb.qnorm, m->qb.qflags, m->qb.qaux, m->cand, m->tau, m->uncertified, m->o_block,
                  m->totals, m->sort_scratch, m->scan_tmp, m->misc, m->stats, m->nan_flag, m->tc.q16, m->tc.dense, m->tc.tau0, m->tc.cand,
                  m->tc.cand_count, m->tc.rec, m->tc.rec_count, m->tc.qsq};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (m->h_block) cudaFreeHost(m->h_block);
  if (m->digest) cudaFree(m->digest);
  for (auto& t : m->tickets) {
    if (t.d_mine) cudaFree(t.d_mine);
    if (t.d_res) cudaFree(t.d_res);
    if (t.h_res) cudaFreeHost(t.h_res);
    if (t.ev) cudaEventDestroy(t.ev);
  }
  for (cudaEvent_t e : m->prof_ev) cudaEventDestroy(e);
  delete m;
  return 0;
}

int hdb_matrix_set_stream(hdb_matrix* m, void* cuda_stream) {
  if (!m) return fail("null handle");
  m->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  return 0;
}
int hdb_matrix_set_post_stream(hdb_matrix* m, void* cuda_stream) {
  if (!m) return fail("null handle");
  HDB_CUDA(cudaSetDevice(m->device));
  HDB_CUDA(cudaStreamSynchronize(m->stream));
  if (m->post_stream) HDB_CUDA(cudaStreamSynchronize(m->post_stream));
  if (m->alt_stream) HDB_CUDA(cudaStreamSynchronize(m->alt_stream));
  m->alt_pending = false;
  m->post_stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  if (m->post_stream) {
    if (!m->alt_stream) HDB_CUDA(cudaStreamCreateWithFlags(&m->alt_stream, cudaStreamNonBlocking));
    if (!m->ev_alt) HDB_CUDA(cudaEventCreateWithFlags(&m->ev_alt, cudaEventDisableTiming));
    if (!m->ev_select) HDB_CUDA(cudaEventCreateWithFlags(&m->ev_select, cudaEventDisableTiming));
    if (!m->ev_prep) HDB_CUDA(cudaEventCreateWithFlags(&m->ev_prep, cudaEventDisableTiming));
    if (!m->ev_qready) HDB_CUDA(cudaEventCreateWithFlags(&m->ev_qready, cudaEventDisableTiming));
    if (!m->pre_stream) {
      // The small kernels must not queue behind the NEXT query's sweep CTAs when the current sweep's CTAs leave the SMs:
      // query preparation runs on a high-priority stream (the caller should give the post stream a high priority too)
      int least = 0, greatest = 0;
      if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess ||
          cudaStreamCreateWithPriority(&m->pre_stream, cudaStreamNonBlocking, greatest) != cudaSuccess) {
        cudaGetLastError();
        HDB_CUDA(cudaStreamCreateWithFlags(&m->pre_stream, cudaStreamNonBlocking));
      }
    }
    for (auto& q : m->slots)
      if (!q.done) HDB_CUDA(cudaEventCreateWithFlags(&q.done, cudaEventDisableTiming));
    m->grid = sweep_grid_size(m->device) - 2;     // leave room on one SM for the certify / exchange / merge kernels
  } else {
    m->grid = sweep_grid_size(m->device);
  }
  // the candidate buffers are sized by the grid: drop them so that they are re-allocated
  slot_store(m);
  for (auto& q : m->slots) { if (q.cand) { cudaFree(q.cand); q.cand = nullptr; } q.pending = false; }
  slot_load(m, m->cur_slot);
  return 0;
}

int hdb_matrix_set_sweep_overlap(hdb_matrix* m, int on) {
  if (!m) return fail("null handle");
  HDB_CUDA(cudaSetDevice(m->device));
  HDB_TRY(join_alt(m));
  m->overlap = on ? 1 : 0;
  return 0;
}

int hdb_matrix_set_path(hdb_matrix* m, int mode) {
  if (!m) return fail("null handle");
  if (mode < 0 || mode > 4) return fail("hdb_matrix_set_path: mode must be 0..4");
  m->path_mode = mode;
  return 0;
}
int hdb_matrix_set_max_group(hdb_matrix* m, int max_queries_per_pass) {
  if (!m) return fail("null handle");
  if (max_queries_per_pass < 0) return fail("hdb_matrix_set_max_group: negative");
  m->max_group = max_queries_per_pass;
  return 0;
}
int hdb_matrix_info(const hdb_matrix* m, int* dtype, int64_t* n_rows, int64_t* dim, int64_t* row_offset, int64_t* n_kept) {
  if (!m) return fail("null handle");
  if (dtype) *dtype = m->dtype;
  if (n_rows) *n_rows = m->n;
  if (dim) *dim = m->d;
  if (row_offset) *row_offset = m->row_offset;
  if (n_kept) *n_kept = m->n_kept;
  return 0;
}

int hdb_matrix_upload(hdb_matrix* m, int64_t row_start, int64_t n_rows, const void* src, int src_space) {
  if (!m) return fail("null handle");
  if (row_start < 0 || n_rows < 0 || row_start + n_rows > m->n) return fail("hdb_matrix_upload: row range outside the shard");
  HDB_CUDA(cudaSetDevice(m->device));
  const size_t row_bytes = (size_t)m->d * dtype_size(m->dtype);
  if (!m->rows) {
    HDB_CUDA(cudaMalloc(&m->rows, cap_rows(m) * row_bytes));
    m->owns_rows = true;
  } else if (!m->owns_rows) {
    return fail("hdb_matrix_upload: the shard uses adopted memory");
  }
  if (n_rows == 0) return 0;
  if (!src) return fail("hdb_matrix_upload: src is NULL");
  HDB_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(m->rows) + (size_t)row_start * row_bytes, src, (size_t)n_rows * row_bytes,
                           src_space == HDB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, m->stream));
  HDB_CUDA(cudaStreamSynchronize(m->stream));
  m->finalized = false;
  return 0;
}

int hdb_matrix_adopt(hdb_matrix* m, void* device_rows) {
  if (!m) return fail("null handle");
  if (m->rows && m->owns_rows) return fail("hdb_matrix_adopt: the shard already owns storage");
  if (!device_rows && m->n > 0) return fail("hdb_matrix_adopt: NULL rows");
  m->rows = device_rows;
  m->owns_rows = false;
  m->finalized = false;
  return 0;
}

int hdb_matrix_finalize(hdb_matrix* m) {
  if (!m) return fail("null handle");
  HDB_CUDA(cudaSetDevice(m->device));
  if (!m->rows && m->n > 0) return fail("hdb_matrix_finalize: nothing uploaded");
  const size_t nsz = (m->dtype == 2) ? 8 : 4;
  if (m->norms) { cudaFree(m->norms); m->norms = nullptr; }
  if (m->inv_norms) { cudaFree(m->inv_norms); m->inv_norms = nullptr; }
  if (m->bits) { cudaFree(m->bits); m->bits = nullptr; }
  if (m->sqnorms) { cudaFree(m->sqnorms); m->sqnorms = nullptr; }
  for (void** p : {&m->pmean, &m->pstd, &m->pscale}) if (*p) { cudaFree(*p); *p = nullptr; }
  if (m->dtype != 2) HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->sqnorms), cap_rows(m) * 4));
  HDB_CUDA(cudaMalloc(&m->norms, cap_rows(m) * nsz));
  HDB_CUDA(cudaMalloc(&m->inv_norms, cap_rows(m) * nsz));
  HDB_CUDA(cudaMemsetAsync(m->stats, 0, 8, m->stream));
  HDB_CUDA(cudaMemsetAsync(m->nan_flag, 0, 4, m->stream));
  MatrixView v = view_of(m);
  HDB_TRY(launch_row_stats(v, m->norms, m->inv_norms, m->sqnorms, m->stats, m->nan_flag, m->stream));
  float hs[2] = {0.f, 1.f};
  int hnan = 0;
  HDB_CUDA(cudaMemcpyAsync(hs, m->stats, 8, cudaMemcpyDeviceToHost, m->stream));
  HDB_CUDA(cudaMemcpyAsync(&hnan, m->nan_flag, 4, cudaMemcpyDeviceToHost, m->stream));
  HDB_CUDA(cudaStreamSynchronize(m->stream));
  if (hnan) return fail("Vectors and query_vector should not contain NaN values.");
  m->max_norm = hs[0];
  m->max_ratio = hs[1] > 0.f ? hs[1] : 1.f;
  m->words = (int)(((m->d + 31) / 32 + 3) / 4 * 4);
  m->finalized = true;
  HDB_TRY(refresh_kept(m));
  return 0;
}

static int ensure_bits(hdb_matrix* m) {
  if (m->bits) return 0;
  HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->bits), cap_rows(m) * m->words * 4));
  return launch_pack_bits(view_of(m), m->bits, m->words, m->stream);
}

// np.mean / np.std columns and certificate statistics of rows [r0, r0+cnt); the statistics are merged into the running ones
static int pearson_stats_rows(hdb_matrix* m, int64_t r0, int64_t cnt) {
  if (cnt == 0) return 0;
  const size_t nsz = (m->dtype == 2) ? 8 : 4;
  uint32_t init[4] = {0u, 0xffffffffu, 0u, 0u};
  memcpy(&init[0], &m->max_pratio, 4);
  memcpy(&init[2], &m->max_cratio, 4);
  if (m->min_pstd > 0.f) { const float neg = -m->min_pstd; memcpy(&init[1], &neg, 4); }
  HDB_CUDA(cudaMemcpyAsync(m->stats, init, 16, cudaMemcpyHostToDevice, m->stream));
  MatrixView v = view_of(m);
  v.rows = reinterpret_cast<const char*>(m->rows) + (size_t)r0 * m->d * dtype_size(m->dtype);
  v.n = cnt;
  HDB_TRY(launch_pearson_stats(v, reinterpret_cast<char*>(m->pmean) + r0 * nsz, reinterpret_cast<char*>(m->pstd) + r0 * nsz,
                               reinterpret_cast<char*>(m->pscale) + r0 * nsz, m->stats, m->stream));
  uint32_t got[4];
  HDB_CUDA(cudaMemcpyAsync(got, m->stats, 16, cudaMemcpyDeviceToHost, m->stream));
  HDB_CUDA(cudaStreamSynchronize(m->stream));
  float negstd;
  memcpy(&m->max_pratio, &got[0], 4);
  memcpy(&m->max_cratio, &got[2], 4);
  memcpy(&negstd, &got[1], 4);
  m->min_pstd = (got[1] == 0xffffffffu || negstd < -1.0e38f) ? 0.f : -negstd;     // 0: every row is constant
  return 0;
}

// pearson_correlation's np.mean / np.std per row (ranking_algorithm.py:91,94), once per matrix instead of once per query
static int ensure_pearson(hdb_matrix* m) {
  if (m->pmean) return 0;
  const size_t nsz = (m->dtype == 2) ? 8 : 4;
  const size_t cnt = cap_rows(m);
  HDB_CUDA(cudaMalloc(&m->pmean, cnt * nsz));
  HDB_CUDA(cudaMalloc(&m->pstd, cnt * nsz));
  HDB_CUDA(cudaMalloc(&m->pscale, cnt * nsz));
  m->max_pratio = 0.f; m->max_cratio = 0.f; m->min_pstd = 0.f;
  return pearson_stats_rows(m, 0, m->n);
}

int hdb_matrix_set_mask(hdb_matrix* m, const uint32_t* bits, int src_space) {
  if (!m) return fail("null handle");
  HDB_CUDA(cudaSetDevice(m->device));
  HDB_TRY(join_alt(m));
  if (!bits) {
    if (m->mask) { cudaStreamSynchronize(m->stream); cudaFree(m->mask); m->mask = nullptr; }
  } else {
    const size_t words = (size_t)((m->n + 31) / 32);
    if (!m->mask) HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->mask), ((cap_rows(m) + 31) / 32) * 4));
    HDB_CUDA(cudaMemcpyAsync(m->mask, bits, words * 4, src_space == HDB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                             m->stream));
  }
  m->decay_valid = false;
  return refresh_kept(m);
}

int hdb_matrix_set_row_order(hdb_matrix* m, const uint32_t* order, int src_space) {
  if (!m) return fail("null handle");
  if (!m->finalized) return fail("hdb_matrix_set_row_order: call hdb_matrix_finalize first");
  HDB_CUDA(cudaSetDevice(m->device));
  HDB_TRY(quiesce(m));
  if (!order) {
    if (m->ord) { cudaFree(m->ord); m->ord = nullptr; }
    if (m->inv) { cudaFree(m->inv); m->inv = nullptr; }
    return 0;
  }
  if (m->n >= (int64_t(1) << 32)) return fail("hdb_matrix_set_row_order: more than 2^32 rows per shard");
  if (!m->ord) HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->ord), (size_t)cap_rows(m) * 4));
  if (!m->inv) HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->inv), (size_t)cap_rows(m) * 4));
  HDB_CUDA(cudaMemcpyAsync(m->ord, order, (size_t)m->n * 4, src_space == HDB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                           m->stream));
  int bad = 0;
  int rc = launch_invert_order(m->ord, m->inv, m->n, m->nan_flag, m->stream);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(&bad, m->nan_flag, 4, cudaMemcpyDeviceToHost, m->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "hdb_matrix_set_row_order");
  }
  if (!rc && bad) rc = fail("hdb_matrix_set_row_order: `order` is not a permutation of 0 .. n_rows-1");
  if (rc) {
    cudaFree(m->ord); cudaFree(m->inv);
    m->ord = nullptr; m->inv = nullptr;
  }
  return rc;
}

int hdb_matrix_set_range(hdb_matrix* m, int64_t lo, int64_t hi) {
  if (!m) return fail("null handle");
  if (lo < 0) lo = 0;
  if (hi > m->n) hi = m->n;
  if (hi < lo) hi = lo;
  m->lo = lo; m->hi = hi;
  m->decay_valid = false;
  HDB_CUDA(cudaSetDevice(m->device));
  return refresh_kept(m);
}

int hdb_matrix_set_timestamps(hdb_matrix* m, const double* ts, int src_space) {
  if (!m) return fail("null handle");
  HDB_CUDA(cudaSetDevice(m->device));
  HDB_TRY(join_alt(m));
  m->decay_valid = false;
  if (!ts) {
    if (m->ts) { cudaStreamSynchronize(m->stream); cudaFree(m->ts); m->ts = nullptr; }
    if (m->decay) { cudaFree(m->decay); m->decay = nullptr; }
    return 0;
  }
  if (!m->ts) HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->ts), cap_rows(m) * 8));
  if (!m->decay) HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->decay), cap_rows(m) * 8));
  HDB_CUDA(cudaMemcpyAsync(m->ts, ts, (size_t)m->n * 8, src_space == HDB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                           m->stream));
  HDB_CUDA(cudaStreamSynchronize(m->stream));
  return 0;
}

int hdb_matrix_kept_ts_max(hdb_matrix* m, double* ts_max, int64_t* n_kept) {
  if (!m) return fail("null handle");
  if (!m->ts) return fail("hdb_matrix_kept_ts_max: no timestamps set");
  HDB_CUDA(cudaSetDevice(m->device));
  HDB_TRY(join_alt(m));
  RowFilter f = filter_of(m, 0.0, false);
  HDB_TRY(launch_kept_ts_max(m->ts, f, m->n, m->misc, m->misc + 1, m->misc + 2, m->stream));
  unsigned long long host[3];
  HDB_CUDA(cudaMemcpyAsync(host, m->misc, 24, cudaMemcpyDeviceToHost, m->stream));
  HDB_CUDA(cudaStreamSynchronize(m->stream));
  if (ts_max) *ts_max = host[1] ? decode_ordered_double(host[0]) : -INFINITY;
  if (n_kept) *n_kept = (int64_t)host[1];
  m->n_kept = (int64_t)host[1];
  m->kept_windows = (int64_t)host[2];
  return 0;
}

int hdb_matrix_set_decay_reference(hdb_matrix* m, double ts_max) {
  if (!m) return fail("null handle");
  if (!m->ts) return fail("hdb_matrix_set_decay_reference: no timestamps set");
  HDB_CUDA(cudaSetDevice(m->device));
  HDB_TRY(join_alt(m));
  if (!m->decay) HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->decay), cap_rows(m) * 8));     // dropped by a mutation
  HDB_TRY(launch_decay(m->ts, m->decay, m->n, ts_max, m->stream));
  m->decay_valid = true;
  return 0;
}

int hdb_matrix_stage1_recency(hdb_matrix* m, double bias1, double ts_max) {
  if (!m) return fail("null handle");
  if (!m->ts) return fail("hdb_matrix_stage1_recency: no timestamps set");
  HDB_CUDA(cudaSetDevice(m->device));
  HDB_TRY(join_alt(m));
  m->decay_valid = false;
  return launch_stage1(m->ts, m->n, bias1, ts_max, m->stream);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Mutation of a resident shard: the reference re-materialises the whole matrix on every
// add (np.concatenate, hyperdb/hyperdb.py:504-509) and every remove_document (np.vstack / mask copy, :718-728).
// ---------------------------------------------------------------------------------------------
static int quiesce(hdb_matrix* m) {
  HDB_CUDA(cudaSetDevice(m->device));
  HDB_CUDA(cudaStreamSynchronize(m->stream));
  if (m->post_stream) HDB_CUDA(cudaStreamSynchronize(m->post_stream));
  if (m->pre_stream) HDB_CUDA(cudaStreamSynchronize(m->pre_stream));
  if (m->alt_stream) HDB_CUDA(cudaStreamSynchronize(m->alt_stream));
  m->alt_pending = false;
  for (auto& q : m->slots) q.pending = false;
  return 0;
}

// the per-query row subset / decay state does not survive a change of the row set
static void drop_row_state(hdb_matrix* m) {
  if (m->mask) { cudaFree(m->mask); m->mask = nullptr; }
  if (m->decay) { cudaFree(m->decay); m->decay = nullptr; }
  m->decay_valid = false;
  if (m->totals) { cudaFree(m->totals); m->totals = nullptr; }
}

static int grow_column(void** p, size_t row_bytes, int64_t n_used, int64_t new_cap, cudaStream_t s) {
  if (!*p) return 0;
  void* fresh = nullptr;
  HDB_CUDA(cudaMalloc(&fresh, (size_t)(new_cap > 0 ? new_cap : 1) * row_bytes));
  if (n_used > 0) HDB_CUDA(cudaMemcpyAsync(fresh, *p, (size_t)n_used * row_bytes, cudaMemcpyDeviceToDevice, s));
  HDB_CUDA(cudaStreamSynchronize(s));
  cudaFree(*p);
  *p = fresh;
  return 0;
}

extern "C" int hdb_matrix_reserve(hdb_matrix* m, int64_t capacity_rows) {
  if (!m) return fail("null handle");
  if (m->rows && !m->owns_rows) return fail("hdb_matrix_reserve: the shard uses adopted memory");
  if (capacity_rows <= m->cap) return 0;
  HDB_TRY(quiesce(m));
  const size_t row_bytes = (size_t)m->d * dtype_size(m->dtype);
  const size_t nsz = (m->dtype == 2) ? 8 : 4;
  if (!m->rows) {
    HDB_CUDA(cudaMalloc(&m->rows, (size_t)capacity_rows * row_bytes));
    m->owns_rows = true;
  } else {
    HDB_TRY(grow_column(&m->rows, row_bytes, m->n, capacity_rows, m->stream));
  }
  HDB_TRY(grow_column(&m->norms, nsz, m->n, capacity_rows, m->stream));
  HDB_TRY(grow_column(&m->inv_norms, nsz, m->n, capacity_rows, m->stream));
  HDB_TRY(grow_column(reinterpret_cast<void**>(&m->sqnorms), 4, m->n, capacity_rows, m->stream));
  HDB_TRY(grow_column(reinterpret_cast<void**>(&m->bits), (size_t)m->words * 4, m->n, capacity_rows, m->stream));
  HDB_TRY(grow_column(&m->pmean, nsz, m->n, capacity_rows, m->stream));
  HDB_TRY(grow_column(&m->pstd, nsz, m->n, capacity_rows, m->stream));
  HDB_TRY(grow_column(&m->pscale, nsz, m->n, capacity_rows, m->stream));
  HDB_TRY(grow_column(reinterpret_cast<void**>(&m->ts), 8, m->n, capacity_rows, m->stream));
  HDB_TRY(grow_column(reinterpret_cast<void**>(&m->decay), 8, m->n, capacity_rows, m->stream));
  if (m->mask) {          // one bit per row, whole words
    uint32_t* fresh = nullptr;
    HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&fresh), (size_t)((capacity_rows + 31) / 32) * 4));
    HDB_CUDA(cudaMemcpy(fresh, m->mask, (size_t)((m->n + 31) / 32) * 4, cudaMemcpyDeviceToDevice));
    cudaFree(m->mask);
    m->mask = fresh;
  }
  if (m->totals) { cudaFree(m->totals); m->totals = nullptr; }
  m->cap = capacity_rows;
  return 0;
}

extern "C" int hdb_matrix_append(hdb_matrix* m, int64_t n_rows, const void* src, int src_space) {
  if (!m) return fail("null handle");
  if (!m->finalized) return fail("hdb_matrix_append: call hdb_matrix_finalize first");
  if (m->rows && !m->owns_rows) return fail("hdb_matrix_append: the shard uses adopted memory");
  if (m->ord) return fail("hdb_matrix_append: a row order is set (rebuild the shard, or clear the order first)");
  if (n_rows < 0) return fail("hdb_matrix_append: negative row count");
  if (n_rows == 0) return 0;
  if (!src) return fail("hdb_matrix_append: src is NULL");
  if (m->n + n_rows >= (int64_t(1) << 32)) return fail("hdb_matrix_append: more than 2^32 rows per shard");
  HDB_TRY(quiesce(m));
  if (m->n + n_rows > m->cap || !m->rows) {
    int64_t want = m->cap + m->cap / 2;                       // geometric growth: amortised O(1) copies per appended row
    if (want < m->n + n_rows) want = m->n + n_rows;
    HDB_TRY(hdb_matrix_reserve(m, want));
  }
  const size_t row_bytes = (size_t)m->d * dtype_size(m->dtype);
  const size_t nsz = (m->dtype == 2) ? 8 : 4;
  const int64_t r0 = m->n;
  HDB_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(m->rows) + (size_t)r0 * row_bytes, src, (size_t)n_rows * row_bytes,
                           src_space == HDB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, m->stream));
  // ingest statistics of the new rows only, merged into the running maxima
  float hs[2] = {m->max_norm, m->max_ratio};
  HDB_CUDA(cudaMemcpyAsync(m->stats, hs, 8, cudaMemcpyHostToDevice, m->stream));
  HDB_CUDA(cudaMemsetAsync(m->nan_flag, 0, 4, m->stream));
  MatrixView sub = view_of(m);
  sub.rows = reinterpret_cast<const char*>(m->rows) + (size_t)r0 * row_bytes;
  sub.n = n_rows;
  HDB_TRY(launch_row_stats(sub, reinterpret_cast<char*>(m->norms) + r0 * nsz, reinterpret_cast<char*>(m->inv_norms) + r0 * nsz,
                           m->sqnorms ? m->sqnorms + r0 : nullptr, m->stats, m->nan_flag, m->stream));
  int hnan = 0;
  HDB_CUDA(cudaMemcpyAsync(hs, m->stats, 8, cudaMemcpyDeviceToHost, m->stream));
  HDB_CUDA(cudaMemcpyAsync(&hnan, m->nan_flag, 4, cudaMemcpyDeviceToHost, m->stream));
  HDB_CUDA(cudaStreamSynchronize(m->stream));
  if (hnan) return fail("Vectors and query_vector should not contain NaN values.");       // the shard is unchanged
  m->max_norm = hs[0];
  m->max_ratio = hs[1] > 0.f ? hs[1] : 1.f;
  if (m->bits) HDB_TRY(launch_pack_bits(sub, m->bits + (size_t)r0 * m->words, m->words, m->stream));
  if (m->pmean) HDB_TRY(pearson_stats_rows(m, r0, n_rows));
  if (m->ts) { cudaFree(m->ts); m->ts = nullptr; }          // timestamps of the new rows are unknown: set them again
  drop_row_state(m);
  m->n += n_rows;
  m->lo = 0; m->hi = m->n;
  return refresh_kept(m);
}

extern "C" int hdb_matrix_remove_rows(hdb_matrix* m, const int64_t* local_rows, int64_t count, int src_space) {
  if (!m) return fail("null handle");
  if (!m->finalized) return fail("hdb_matrix_remove_rows: call hdb_matrix_finalize first");
  if (m->rows && !m->owns_rows) return fail("hdb_matrix_remove_rows: the shard uses adopted memory");
  if (m->ord) return fail("hdb_matrix_remove_rows: a row order is set (rebuild the shard, or clear the order first)");
  if (count < 0) return fail("hdb_matrix_remove_rows: negative count");
  if (count == 0 || m->n == 0) return count == 0 ? 0 : fail("hdb_matrix_remove_rows: row index out of range");
  if (m->n > 0x7fffffff) return fail("hdb_matrix_remove_rows: more than 2^31 rows per shard");
  if (!local_rows) return fail("hdb_matrix_remove_rows: NULL rows");
  HDB_TRY(quiesce(m));
  const int64_t n = m->n;
  int64_t* d_rows = nullptr;
  uint32_t* plan = nullptr;            // keep | pos | src
  void* bounce = nullptr;
  int rc = 0;
  auto cleanup = [&]() { if (d_rows) cudaFree(d_rows); if (plan) cudaFree(plan); if (bounce) cudaFree(bounce); };
  const int64_t* rows_dev = local_rows;
  if (src_space != HDB_DEVICE) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&d_rows), (size_t)count * 8);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_rows, local_rows, (size_t)count * 8, cudaMemcpyHostToDevice, m->stream);
    if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "remove_rows staging"); }
    rows_dev = d_rows;
  }
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&plan), (size_t)n * 12);
  if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "remove_rows plan"); }
  cudaMemsetAsync(m->nan_flag, 0, 4, m->stream);
  rc = launch_plan_removal(n, rows_dev, count, plan, plan + n, plan + 2 * n, m->misc, m->nan_flag, &m->scan_tmp, &m->scan_tmp_bytes, m->stream);
  unsigned long long n_new = 0;
  int bad = 0;
  if (!rc) {
    e = cudaMemcpyAsync(&n_new, m->misc, 8, cudaMemcpyDeviceToHost, m->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, m->nan_flag, 4, cudaMemcpyDeviceToHost, m->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "remove_rows plan readback");
  }
  if (!rc && bad) rc = fail("hdb_matrix_remove_rows: row index out of range");
  if (!rc && (int64_t)n_new < n) {
    const size_t row_bytes = (size_t)m->d * dtype_size(m->dtype);
    const size_t nsz = (m->dtype == 2) ? 8 : 4;
    size_t bounce_bytes = (size_t)64 << 20;
    if (bounce_bytes < row_bytes) bounce_bytes = row_bytes;
    e = cudaMalloc(&bounce, bounce_bytes);
    if (e != cudaSuccess) rc = cuda_fail(e, "remove_rows bounce");
    const uint32_t* src = plan + 2 * n;
    struct Col { void* p; size_t rb; } cols[] = {{m->rows, row_bytes}, {m->norms, nsz}, {m->inv_norms, nsz}, {m->sqnorms, 4},
                                                  {m->bits, (size_t)m->words * 4}, {m->pmean, nsz}, {m->pstd, nsz}, {m->pscale, nsz}, {m->ts, 8}};
    for (const Col& c : cols)
      if (!rc) rc = launch_compact_column(c.p, (int64_t)c.rb, (int64_t)n_new, src, bounce, bounce_bytes, m->stream);
    if (!rc) { e = cudaStreamSynchronize(m->stream); if (e != cudaSuccess) rc = cuda_fail(e, "remove_rows compaction"); }
    if (!rc) {
      drop_row_state(m);                 // max_norm / max_ratio / pearson statistics stay valid as bounds over a subset
      m->n = (int64_t)n_new;
      m->lo = 0; m->hi = m->n;
    }
  }
  cleanup();
  if (rc) return rc;
  return refresh_kept(m);
}

extern "C" int hdb_matrix_set_row_offset(hdb_matrix* m, int64_t row_offset) {
  if (!m) return fail("null handle");
  if (row_offset < 0) return fail("hdb_matrix_set_row_offset: negative offset");
  m->row_offset = row_offset;
  return 0;
}

extern "C" {

// ---------------------------------------------------------------------------------------------
static int ensure_workspace(hdb_matrix* m, int64_t nq, int64_t k) {
  const int64_t cq = nq < kChunk ? (nq < 1 ? 1 : nq) : kChunk;       // fused chunk capacity
  if (m->ws_q < nq) {
    HDB_TRY(dev_alloc(reinterpret_cast<char**>(&m->q_raw), (size_t)nq * m->d * 8));
    HDB_TRY(dev_alloc(reinterpret_cast<char**>(&m->qb.qa), (size_t)nq * m->d * 8));
    HDB_TRY(dev_alloc(&m->qb.qc, (size_t)nq * m->d));
    HDB_TRY(dev_alloc(&m->qb.qbits, (size_t)nq * m->words));
    HDB_TRY(dev_alloc(&m->qb.qnorm, (size_t)nq));
    HDB_TRY(dev_alloc(&m->qb.qflags, (size_t)nq));
    HDB_TRY(dev_alloc(&m->qb.qaux, (size_t)nq * 2));
    HDB_TRY(dev_alloc(&m->tau, (size_t)nq));
    m->ws_q = nq;
  }
  if (!m->cand) HDB_TRY(dev_alloc(&m->cand, (size_t)kChunk * m->grid * kMaxKP));
  (void)cq;
  const size_t kk = (size_t)(k > 0 ? k : 1);
  const size_t need = (size_t)nq * kk * 16 + (size_t)nq * 8 + (size_t)nq * 4;
  if (m->o_block_bytes < need) {
    if (m->o_block) cudaFree(m->o_block);
    if (m->h_block) cudaFreeHost(m->h_block);
    m->o_block = nullptr; m->h_block = nullptr; m->o_block_bytes = 0;
    HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->o_block), need));
    HDB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&m->h_block), need));
    m->o_block_bytes = need;
  }
  m->o_idx = reinterpret_cast<int64_t*>(m->o_block);
  m->o_score = reinterpret_cast<double*>(m->o_block + (size_t)nq * kk * 8);
  m->o_count = reinterpret_cast<int64_t*>(m->o_block + (size_t)nq * kk * 16);
  m->o_flags = reinterpret_cast<uint32_t*>(m->o_block + (size_t)nq * kk * 16 + (size_t)nq * 8);
  if (!m->totals) HDB_TRY(dev_alloc(&m->totals, cap_rows(m)));
  return 0;
}

static int pick_kp(const hdb_matrix* m, int64_t k) {
  if (m->path_mode == 1) return 0;
  if (m->path_mode == 4 && k <= 100) return 128;          // repair mode: the wide candidate class on the streaming sweep
  if (k <= 16) return 32;
  if (k <= 100) return 128;
  return 0;
}

// Enqueue the fused path for queries [b0, b0+cnt) of the prepared batch; results go to (idx, score, count, flags).
static int run_fused(hdb_matrix* m, int metric, int rdt, int kp, int64_t b0, int64_t cnt, int64_t k, const RowFilter& f,
                     int64_t* idx, double* score, int64_t* count, uint32_t* flags, cudaStream_t fin_stream = nullptr,
                     cudaStream_t sweep_stream = nullptr, const PushTarget* push = nullptr) {
  MatrixView v = view_of(m);
  cudaStream_t sw = sweep_stream ? sweep_stream : m->stream;
  if (!(fin_stream && fin_stream != m->stream)) HDB_CUDA(cudaMemsetAsync(m->tau + b0, 0, (size_t)cnt * 8, sw));   // pipelined: prep zeroed it
  const int elt = (m->dtype == 2) ? 8 : 4;
  // one pass of the matrix serves up to `gmax` queries (multi-query sweep): a batch of B queries costs ~B / gmax reads
  const int gmax = m->max_group > 0 ? (m->max_group < sweep_max_group(v, metric, kp) ? m->max_group : sweep_max_group(v, metric, kp))
                                    : sweep_max_group(v, metric, kp);
  for (int64_t i = 0; i < cnt;) {
    int g = 1;
    while (g * 2 <= gmax && i + g * 2 <= cnt) g *= 2;
    SweepOut so;
    so.cand = m->cand + (size_t)i * m->grid * kp;
    so.tau = m->tau + b0 + i;
    so.grid = m->grid;
    const void* qa = reinterpret_cast<const char*>(m->qb.qa) + (size_t)(b0 + i) * m->d * elt;
    const uint32_t* qbits = m->qb.qbits + (size_t)(b0 + i) * m->words;
    const bool prof = m->prof_used + 2 <= m->prof_ev.size();
    if (prof) HDB_CUDA(cudaEventRecord(m->prof_ev[m->prof_used], sw));
    HDB_TRY(launch_sweep(v, metric, rdt, qa, qbits, m->qb.qaux + 2 * (b0 + i), f, kp, so, g, sw));
    if (prof) { HDB_CUDA(cudaEventRecord(m->prof_ev[m->prof_used + 1], sw)); m->prof_used += 2; }
    i += g;
  }
  FinalizeArgs a;
  a.m = v; a.f = f; a.metric = metric; a.rdt = rdt; a.kp = kp; a.k = (int)k; a.n_kept = m->n_kept; a.grid = m->grid;
  a.cand = m->cand; a.tau = m->tau + b0;
  a.qb = m->qb;
  a.qb.qc += b0 * m->d; a.qb.qbits += b0 * m->words; a.qb.qnorm += b0; a.qb.qflags += b0; a.qb.qaux += 2 * b0;
  a.out_idx = idx + b0 * k; a.out_score = score + b0 * k; a.out_count = count + b0; a.out_flags = flags ? flags + b0 : nullptr;
  a.uncertified = m->uncertified;
  a.cand_count = nullptr; a.cand_stride = 0; a.tau0 = nullptr; a.extra_flags = 0; a.tau0_negd2 = 0;
  a.push = PushTarget{};
  if (push) a.push = *push;          // row-sharded: the certify kernel stores its results into every rank's exchange slot
  if (fin_stream && fin_stream != m->stream) {
    // pipelined: the certify step runs on the post stream, ordered after the sweeps by an event
    HDB_CUDA(cudaEventRecord(m->ev_select, sw));
    HDB_CUDA(cudaStreamWaitEvent(fin_stream, m->ev_select, 0));
    return launch_finalize(a, cnt, fin_stream);
  }
  return launch_finalize(a, cnt, sw);
}

constexpr int64_t kTcChunk = 4096;       // queries per tensor-core batch (bounds the sample / candidate workspaces)

static int ensure_tc_workspace(hdb_matrix* m, int64_t nq, int kp) {
  const int64_t all_tiles = (m->n + 127) / 128;
  int64_t sample_tiles = all_tiles / 8;
  if (sample_tiles > 512) sample_tiles = 512;
  if (sample_tiles < 1) sample_tiles = 1;
  int64_t cap = 1024;
  const int64_t expect = 2 * m->n * kp / (sample_tiles * 128);
  while (cap < expect && cap < 65536) cap <<= 1;
  // CTA-private record buffers: expected records = n * nq * kp / sample_rows over all SMs; 3x slack
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
  double expect_rec = (double)m->n * (double)nq * kp / (double)(sample_tiles * 128) / sms;
  int64_t rec_cap = (int64_t)(3.0 * expect_rec) + 8192;
  if (rec_cap > (int64_t)1 << 21) rec_cap = (int64_t)1 << 21;
  if (const char* e = getenv("HDB_TC_REC_CAP")) {                // tests: force the record buffers to overflow
    const int64_t forced = atoll(e);
    if (forced >= 16) rec_cap = forced;
  }
  if (m->tc_nq >= nq && m->tc.cap >= cap && m->tc.sample_tiles == sample_tiles && m->tc.rec_cap >= (unsigned)rec_cap && !getenv("HDB_TC_REC_CAP")) return 0;
  void* ptrs[] = {m->tc.q16, m->tc.dense, m->tc.tau0, m->tc.cand, m->tc.cand_count, m->tc.rec, m->tc.rec_count, m->tc.qsq};
  for (void* p : ptrs) if (p) cudaFree(p);
  m->tc = TcWorkspace{};
  m->tc_nq = 0;
  HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->tc.q16), (size_t)nq * m->d * 2));
  HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->tc.dense), (size_t)nq * sample_tiles * 128 * 4));
  HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->tc.tau0), (size_t)nq * 4));
  HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->tc.cand), (size_t)nq * cap * 8));
  HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->tc.cand_count), (size_t)nq * 4));
  HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->tc.qsq), (size_t)nq * 4));
  HDB_CUDA(cudaMalloc(&m->tc.rec, (size_t)sms * rec_cap * 16));
  HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->tc.rec_count), (size_t)sms * 4));
  m->tc.rec_cap = (unsigned)rec_cap;
  m->tc.cap = (int)cap;
  m->tc.sample_tiles = sample_tiles;
  m->tc_nq = nq;
  return 0;
}

// Tensor-core batched path for queries [b0, b0+cnt): sample -> thresholds -> select, then the same finalize.
static int run_tensor(hdb_matrix* m, int metric, int rdt, int kp, int64_t b0, int64_t cnt, int64_t k, const RowFilter& f,
                      int64_t* idx, double* score, int64_t* count, uint32_t* flags) {
  MatrixView v = view_of(m);
  HDB_TRY(ensure_tc_workspace(m, cnt < kTcChunk ? cnt : kTcChunk, kp));
  const float* qa = reinterpret_cast<const float*>(m->qb.qa) + (size_t)b0 * m->d;
  const bool prof = m->prof_used + 2 <= m->prof_ev.size();
  if (prof) HDB_CUDA(cudaEventRecord(m->prof_ev[m->prof_used], m->stream));
  HDB_TRY(launch_batched_tc(v, metric, f, qa, m->qb.qnorm + b0, cnt, kp, m->device, m->tc, m->stream));
  if (prof) { HDB_CUDA(cudaEventRecord(m->prof_ev[m->prof_used + 1], m->stream)); m->prof_used += 2; }
  FinalizeArgs a;
  a.m = v; a.f = f; a.metric = metric; a.rdt = rdt; a.kp = kp; a.k = (int)k; a.n_kept = m->n_kept; a.grid = 0;
  a.cand = m->tc.cand; a.tau = nullptr;
  a.qb = m->qb;
  a.qb.qc += b0 * m->d; a.qb.qbits += b0 * m->words; a.qb.qnorm += b0; a.qb.qflags += b0; a.qb.qaux += 2 * b0;
  a.out_idx = idx + b0 * k; a.out_score = score + b0 * k; a.out_count = count + b0; a.out_flags = flags ? flags + b0 : nullptr;
  a.uncertified = m->uncertified;
  a.cand_count = m->tc.cand_count; a.cand_stride = m->tc.cap; a.tau0 = m->tc.tau0; a.extra_flags = HDB_FLAG_TENSOR;
  a.push = PushTarget{};
  a.tau0_negd2 = (metric == HDB_EUCLIDEAN);
  return launch_finalize(a, cnt, m->stream);
}

static int run_exact(hdb_matrix* m, int metric, int rdt, int64_t b, int64_t k, const RowFilter& f, int64_t* idx, double* score,
                     int64_t* count) {
  MatrixView v = view_of(m);
  HDB_TRY(launch_full_scores(v, f, metric, rdt, m->qb.qc + b * m->d, m->qb.qbits + b * m->words, m->qb.qaux + 2 * b, m->totals, m->stream));
  return exact_topk(m->device, m->totals, m->inv, m->n, m->row_offset, k, m->n_kept, idx + b * k, score + b * k, count + b,
                    &m->sort_scratch, &m->sort_scratch_bytes, m->stream);
}

// Batched pearson on the tensor cores (batched_tc.cu) can be switched off for an A/B against the multi-query sweep:
// HDB_TC_PEARSON=0.  The other tensor-path metrics have no switch (--path / hdb_matrix_set_path_mode covers them).
static bool tc_metric_enabled(int metric) {
  if (metric != HDB_PEARSON) return true;
  static const bool on = [] { const char* e = getenv("HDB_TC_PEARSON"); return !(e && e[0] == '0'); }();
  return on;
}

int hdb_query(hdb_matrix* m, int metric, const void* queries, int q_dtype, int q_space, int64_t nq, int64_t top_k,
              double recency_bias, int64_t* out_idx, double* out_score, int64_t* out_count, uint32_t* out_flags,
              int out_space) {
  if (!m) return fail("null handle");
  if (!m->finalized) return fail("hdb_query: call hdb_matrix_finalize first");
  if (metric < 0 || metric > 6) return fail("Unknown metric");
  if (q_dtype < 0 || q_dtype > 2) return fail("hdb_query: q_dtype must be HDB_F16/F32/F64");
  if (nq < 0) return fail("hdb_query: negative query count");
  if (nq == 0) return 0;
  if (!queries || !out_count) return fail("hdb_query: NULL argument");
  HDB_CUDA(cudaSetDevice(m->device));
  const int64_t k = top_k > 0 ? top_k : 0;
  if (k > 0 && (!out_idx || !out_score)) return fail("hdb_query: NULL output");
  const bool use_decay = m->ts != nullptr && m->decay_valid;
  if (m->ts && !m->decay_valid && recency_bias != 0.0)
    return fail("hdb_query: timestamps set but hdb_matrix_set_decay_reference was not called");
  const bool pipelined = m->post_stream != nullptr && out_space == HDB_DEVICE;
  if (pipelined) {
    // rotate to the other workspace slot; its previous user (two queries ago) must have finished certifying
    slot_store(m);
    slot_load(m, m->cur_slot ^ 1);
    hdb_matrix::QuerySlot& qs = m->slots[m->cur_slot];
    if (qs.pending) HDB_CUDA(cudaStreamWaitEvent(m->pre_stream, qs.done, 0));
  }
  if (!pipelined) HDB_TRY(join_alt(m));
  if (!pipelined && m->post_stream) {
    // a host-output call while pipelining is on: order it after every certify step still in flight on the post stream
    for (auto& qs : m->slots)
      if (qs.pending) { HDB_CUDA(cudaStreamWaitEvent(m->stream, qs.done, 0)); }
  }
  HDB_TRY(ensure_workspace(m, nq, k));
  const int rdt = m->dtype > q_dtype ? m->dtype : q_dtype;
  if (metric == HDB_HAMMING || metric == HDB_JACCARD) HDB_TRY(ensure_bits(m));
  if (metric == HDB_PEARSON) HDB_TRY(ensure_pearson(m));

  const void* q_dev = queries;
  if (q_space == HDB_HOST) {
    HDB_CUDA(cudaMemcpyAsync(m->q_raw, queries, (size_t)nq * m->d * dtype_size(q_dtype), cudaMemcpyHostToDevice,
                             pipelined ? m->pre_stream : m->stream));
    q_dev = m->q_raw;
  }
  if (pipelined) {
    // The preparation of THIS query runs on the internal pre stream, ahead of the sweeps still queued on the main
    // stream.  A device-resident query may have been produced by work enqueued on the handle's stream just before this
    // call (an embedding kernel, a copy): the pre stream is ordered after that work by an event.  With overlapping
    // sweeps the main stream's last sweep is two queries old, so the wait costs nothing.
    if (q_space == HDB_DEVICE) {
      HDB_CUDA(cudaEventRecord(m->ev_qready, m->stream));
      HDB_CUDA(cudaStreamWaitEvent(m->pre_stream, m->ev_qready, 0));
    }
    HDB_TRY(launch_prep_query(q_dev, q_dtype, nq, m->d, metric, m->dtype, m->words, m->qb, m->tau, m->pre_stream));
    HDB_CUDA(cudaEventRecord(m->ev_prep, m->pre_stream));
    HDB_CUDA(cudaStreamWaitEvent(m->stream, m->ev_prep, 0));
  } else {
    HDB_TRY(launch_prep_query(q_dev, q_dtype, nq, m->d, metric, m->dtype, m->words, m->qb, nullptr, m->stream));
  }
  const RowFilter f = filter_of(m, recency_bias, use_decay);
  const bool dev_out = (out_space == HDB_DEVICE);
  int64_t* idx = dev_out ? out_idx : m->o_idx;
  double* score = dev_out ? out_score : m->o_score;
  int64_t* count = dev_out ? out_count : m->o_count;
  uint32_t* flags = dev_out ? out_flags : m->o_flags;

  // row-sharded path: the local result also goes to every rank of the attached exchange
  PushTarget push{};
  bool want_push = false, fused_push = false;
  const int64_t msg_words = 2 * nq * k + nq + (nq + 1) / 2;
  if (m->xchg && dev_out) {
    if (!out_flags || reinterpret_cast<char*>(out_idx) != reinterpret_cast<char*>(out_score) + (size_t)nq * k * 8 ||
        reinterpret_cast<char*>(out_count) != reinterpret_cast<char*>(out_idx) + (size_t)nq * k * 8 ||
        reinterpret_cast<char*>(out_flags) != reinterpret_cast<char*>(out_count) + (size_t)nq * 8)
      return fail("hdb_query: with an exchange attached the device outputs must form one packed block [scores | ids | counts | flags]");
    if (msg_words > exchange_max_words(m->xchg)) return fail("hdb_query: the batch does not fit the attached exchange buffer");
    HDB_TRY(exchange_push_target(m->xchg, nq, k, &push));
    want_push = true;
  }

  int kp = pick_kp(m, k);
  const size_t qsm = (size_t)m->d * (m->dtype == 2 ? 8 : 4);
  if (kp && qsm + 40000 > 200 * 1024) kp = 0;                     // query does not fit next to the lists
  if (kp == 0 && m->path_mode == 2) return fail("hdb_query: fused path forced but not applicable");
  m->last.valid = true; m->last.metric = metric; m->last.rdt = rdt; m->last.kp = kp; m->last.nq = nq; m->last.k = k;
  m->last.bias = recency_bias;

  if (k == 0) {
    HDB_CUDA(cudaMemsetAsync(count, 0, (size_t)nq * 8, m->stream));
    if (flags) HDB_CUDA(cudaMemcpyAsync(flags, m->qb.qflags, (size_t)nq * 4, cudaMemcpyDeviceToDevice, m->stream));
  } else if (kp && m->path_mode != 2 && m->path_mode != 4 && !m->ord && tc_metric_enabled(metric) &&
             batched_tc_supported(view_of(m), metric, q_dtype, nq, use_decay)) {
    if (m->dtype == 1) kp = 128;         // tf32 select: wider error band, so certify a wider candidate list
    // kind::f16 batches of hundreds of queries: with 32 candidates about one query in 24 000 fails its certificate (the
    // non-IEEE accumulation band against the gap between the 10th and the 32nd score), i.e. most 4096-query batches carry
    // a repair; 128 candidates certify them all for 7 % of the device time and the same end-to-end rate without the
    // repair's variance (HDB_TC_WIDE=0: the narrow class, A/B)
    static const bool tc_wide = [] { const char* e = getenv("HDB_TC_WIDE"); return !(e && e[0] == '0'); }();
    if (tc_wide && nq >= 256 && k <= 100) kp = 128;
    m->last.kp = kp;
    HDB_CUDA(cudaMemsetAsync(m->uncertified, 0, 4, m->stream));
    for (int64_t b0 = 0; b0 < nq; b0 += kTcChunk) {
      const int64_t cnt = nq - b0 < kTcChunk ? nq - b0 : kTcChunk;
      HDB_TRY(run_tensor(m, metric, rdt, kp, b0, cnt, k, f, idx, score, count, flags));
    }
  } else if (kp) {
    HDB_CUDA(cudaMemsetAsync(m->uncertified, 0, 4, m->stream));
    bool on_post = false;
    for (int64_t b0 = 0; b0 < nq; b0 += kChunk) {
      const int64_t cnt = nq - b0 < kChunk ? nq - b0 : kChunk;
      // only a single chunk can hand its certify step to the post stream (the candidate buffer is reused per chunk)
      const bool hand_over = pipelined && nq <= kChunk;
      cudaStream_t sw = nullptr;
      if (hand_over && m->overlap && m->alt_stream && (m->cur_slot & 1)) {
        sw = m->alt_stream;                                  // odd slots sweep on the alternate stream (see hdb_matrix::alt_stream)
        HDB_CUDA(cudaStreamWaitEvent(sw, m->ev_prep, 0));
      }
      fused_push = want_push && nq <= kChunk;              // one certify launch covers the batch: it pushes by itself
      HDB_TRY(run_fused(m, metric, rdt, kp, b0, cnt, k, f, idx, score, count, flags, hand_over ? m->post_stream : nullptr, sw,
                        fused_push ? &push : nullptr));
      if (sw) { HDB_CUDA(cudaEventRecord(m->ev_alt, sw)); m->alt_pending = true; }
      on_post = hand_over;
    }
    if (pipelined && on_post) {
      if (want_push && !fused_push) HDB_TRY(exchange_launch_push(m->xchg, m->post_stream, out_score, msg_words));
      hdb_matrix::QuerySlot& qs = m->slots[m->cur_slot];
      HDB_CUDA(cudaEventRecord(qs.done, m->post_stream));
      qs.pending = true;
      slot_store(m);
      return 0;
    }
  } else {
    for (int64_t b = 0; b < nq; ++b) HDB_TRY(run_exact(m, metric, rdt, b, k, f, idx, score, count));
    if (flags) HDB_CUDA(cudaMemcpyAsync(flags, m->qb.qflags, (size_t)nq * 4, cudaMemcpyDeviceToDevice, m->stream));
  }
  if (pipelined) {
    // paths that ran entirely on the main stream: results must still become visible in post-stream order
    HDB_CUDA(cudaEventRecord(m->ev_select, m->stream));
    HDB_CUDA(cudaStreamWaitEvent(m->post_stream, m->ev_select, 0));
    slot_store(m);
  }
  if (want_push && !fused_push)         // tensor-core batches, the exact path, k = 0: push the packed block
    HDB_TRY(exchange_launch_push(m->xchg, pipelined ? m->post_stream : m->stream, out_score, msg_words));
  if (dev_out) return 0;

  // host outputs: ONE device->pinned-host copy of [idx | score | count | flags], one synchronisation;
  // queries the certificate rejected are repaired with the exact path and copied again
  const size_t kk = (size_t)(k > 0 ? k : 1);
  const size_t blk = (size_t)nq * kk * 16 + (size_t)nq * 8 + (size_t)nq * 4;
  const int64_t* h_idx = reinterpret_cast<const int64_t*>(m->h_block);
  const double* h_score = reinterpret_cast<const double*>(m->h_block + (size_t)nq * kk * 8);
  const int64_t* h_count = reinterpret_cast<const int64_t*>(m->h_block + (size_t)nq * kk * 16);
  uint32_t* h_flags = reinterpret_cast<uint32_t*>(m->h_block + (size_t)nq * kk * 16 + (size_t)nq * 8);
  HDB_CUDA(cudaMemcpyAsync(m->h_block, m->o_block, blk, cudaMemcpyDeviceToHost, m->stream));
  HDB_CUDA(cudaStreamSynchronize(m->stream));
  bool any_nan = false, repaired = false;
  std::vector<uint32_t> fixed;
  for (int64_t b = 0; b < nq; ++b) {
    if (h_flags[b] & HDB_FLAG_QUERY_NAN) any_nan = true;
    if (h_flags[b] & HDB_FLAG_UNCERTIFIED) {
      if (m->path_mode == 2) return fail("hdb_query: fused path forced but the certificate failed");
      if (fixed.empty()) fixed.assign(h_flags, h_flags + nq);
      bool done = false;
      if (((h_flags[b] & HDB_FLAG_TENSOR) || m->last.kp < 128) && k <= 100) {
        // a query the batched pass, or the narrow candidate class, could not certify: first retry with the streaming
        // sweep and the wide candidate class (IEEE fp32 accumulation, 128 candidates), only then pay for the exact path
        HDB_CUDA(cudaMemsetAsync(m->uncertified, 0, 4, m->stream));
        HDB_TRY(run_fused(m, metric, rdt, 128, b, 1, k, f, idx, score, count, flags));
        int bad = 0;
        HDB_CUDA(cudaMemcpyAsync(&bad, m->uncertified, 4, cudaMemcpyDeviceToHost, m->stream));
        HDB_CUDA(cudaStreamSynchronize(m->stream));
        done = (bad == 0);
        if (done) fixed[b] = (fixed[b] & ~(HDB_FLAG_UNCERTIFIED | HDB_FLAG_TENSOR));
      }
      if (!done) {
        HDB_TRY(run_exact(m, metric, rdt, b, k, f, idx, score, count));
        fixed[b] = (fixed[b] & ~HDB_FLAG_UNCERTIFIED) | HDB_FLAG_FALLBACK;
      }
      repaired = true;
    }
  }
  if (any_nan) return fail("Vectors and query_vector should not contain NaN values.");
  if (repaired) {
    HDB_CUDA(cudaMemcpyAsync(m->h_block, m->o_block, blk, cudaMemcpyDeviceToHost, m->stream));
    HDB_CUDA(cudaStreamSynchronize(m->stream));
    memcpy(h_flags, fixed.data(), (size_t)nq * 4);
  }
  if (k > 0) {
    memcpy(out_idx, h_idx, (size_t)nq * k * 8);
    memcpy(out_score, h_score, (size_t)nq * k * 8);
  }
  memcpy(out_count, h_count, (size_t)nq * 8);
  if (out_flags) memcpy(out_flags, h_flags, (size_t)nq * 4);
  return 0;
}

int hdb_matrix_attach_exchange(hdb_matrix* m, hdb_exchange* x) {
  if (!m) return fail("null handle");
  if (x && exchange_device(x) != m->device) return fail("hdb_matrix_attach_exchange: the exchange lives on another device");
  HDB_TRY(quiesce(m));
  m->xchg = x;
  return 0;
}

// ---- asynchronous host API --------------------------------------------------------------------------------------
static int ticket_reserve(hdb_matrix::Ticket& t, int64_t nq, int64_t k, int world) {
  const size_t kk = (size_t)(k > 0 ? k : 0);
  const size_t mine = ((size_t)nq * kk * 16 + (size_t)nq * 8 + (size_t)((nq + 1) / 2) * 8);
  const size_t res = (size_t)nq * kk * 16 + (size_t)nq * 8 + (((size_t)world * nq * 4 + 7) & ~size_t(7));
  if (t.mine_bytes < mine) {
    if (t.d_mine) cudaFree(t.d_mine);
    t.d_mine = nullptr; t.mine_bytes = 0;
    HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&t.d_mine), mine));
    t.mine_bytes = mine;
  }
  const size_t need = res > mine ? res : mine;
  if (t.res_bytes < need) {
    if (t.d_res) cudaFree(t.d_res);
    if (t.h_res) cudaFreeHost(t.h_res);
    t.d_res = nullptr; t.h_res = nullptr; t.res_bytes = 0;
    HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&t.d_res), need));
    HDB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&t.h_res), need));
    t.res_bytes = need;
  }
  if (!t.ev) HDB_CUDA(cudaEventCreateWithFlags(&t.ev, cudaEventDisableTiming));
  t.nq = nq; t.k = (int64_t)kk; t.world = world;
  return 0;
}

int hdb_query_submit(hdb_matrix* m, int metric, const void* queries, int q_dtype, int q_space, int64_t nq, int64_t top_k,
                     double recency_bias, int world, int64_t* ticket) {
  if (!m || !ticket) return fail("hdb_query_submit: NULL argument");
  if (nq < 1) return fail("hdb_query_submit: need at least one query");
  if (world < 1 || (world > 1 && !m->xchg)) return fail("hdb_query_submit: world > 1 needs an attached exchange");
  HDB_CUDA(cudaSetDevice(m->device));
  hdb_matrix::Ticket& t = m->tickets[m->next_ticket % hdb_matrix::kTickets];
  if (t.busy) return fail("hdb_query_submit: too many queries in flight, collect the oldest ticket first");
  const int64_t k = top_k > 0 ? top_k : 0;
  HDB_TRY(ticket_reserve(t, nq, k, m->xchg ? world : 1));
  double* sc = reinterpret_cast<double*>(t.d_mine);
  int64_t* idx = reinterpret_cast<int64_t*>(t.d_mine + (size_t)nq * k * 8);
  int64_t* cnt = reinterpret_cast<int64_t*>(t.d_mine + (size_t)nq * k * 16);
  uint32_t* flg = reinterpret_cast<uint32_t*>(t.d_mine + (size_t)nq * k * 16 + (size_t)nq * 8);
  HDB_TRY(hdb_query(m, metric, queries, q_dtype, q_space, nq, top_k, recency_bias, idx, sc, cnt, flg, HDB_DEVICE));
  if (m->xchg) {
    // the wait + merge kernel is ordered after the pushes through the arrival flags, not through a stream
    cudaStream_t xs = exchange_stream(m->xchg);
    int64_t* r_idx = reinterpret_cast<int64_t*>(t.d_res);
    double* r_sc = reinterpret_cast<double*>(t.d_res + (size_t)nq * k * 8);
    int64_t* r_cnt = reinterpret_cast<int64_t*>(t.d_res + (size_t)nq * k * 16);
    uint32_t* r_flg = reinterpret_cast<uint32_t*>(t.d_res + (size_t)nq * k * 16 + (size_t)nq * 8);
    HDB_TRY(exchange_launch_wait_merge(m->xchg, xs, nq, k, r_idx, r_sc, r_cnt, r_flg));
    const size_t res = (size_t)nq * k * 16 + (size_t)nq * 8 + (((size_t)t.world * nq * 4 + 7) & ~size_t(7));
    HDB_CUDA(cudaMemcpyAsync(t.h_res, t.d_res, res, cudaMemcpyDeviceToHost, xs));
    HDB_CUDA(cudaEventRecord(t.ev, xs));
  } else {
    // pipelined mode (a post stream is set): results become valid in post-stream order; else on the handle's stream
    cudaStream_t rs = m->post_stream ? m->post_stream : m->stream;
    const size_t mine = (size_t)nq * k * 16 + (size_t)nq * 8 + (size_t)((nq + 1) / 2) * 8;
    HDB_CUDA(cudaMemcpyAsync(t.h_res, t.d_mine, mine, cudaMemcpyDeviceToHost, rs));
    HDB_CUDA(cudaEventRecord(t.ev, rs));
  }
  t.busy = true;
  *ticket = m->next_ticket++;
  return 0;
}

int hdb_query_collect(hdb_matrix* m, int64_t ticket, int64_t* out_idx, double* out_score, int64_t* out_count, uint32_t* out_flags) {
  if (!m || !out_count) return fail("hdb_query_collect: NULL argument");
  if (ticket < 0 || ticket >= m->next_ticket || ticket + hdb_matrix::kTickets < m->next_ticket) return fail("hdb_query_collect: unknown ticket");
  hdb_matrix::Ticket& t = m->tickets[ticket % hdb_matrix::kTickets];
  if (!t.busy) return fail("hdb_query_collect: the ticket was already collected");
  HDB_CUDA(cudaSetDevice(m->device));
  HDB_CUDA(cudaEventSynchronize(t.ev));
  t.busy = false;
  const size_t nk = (size_t)t.nq * t.k;
  if (m->xchg) {          // [idx | score | count | flags world*nq]
    if (nk) { memcpy(out_idx, t.h_res, nk * 8); memcpy(out_score, t.h_res + nk * 8, nk * 8); }
    memcpy(out_count, t.h_res + nk * 16, (size_t)t.nq * 8);
    if (out_flags) memcpy(out_flags, t.h_res + nk * 16 + (size_t)t.nq * 8, (size_t)t.world * t.nq * 4);
  } else {                // the local packed block [scores | ids | counts | flags]
    if (nk) { memcpy(out_score, t.h_res, nk * 8); memcpy(out_idx, t.h_res + nk * 8, nk * 8); }
    memcpy(out_count, t.h_res + nk * 16, (size_t)t.nq * 8);
    if (out_flags) memcpy(out_flags, t.h_res + nk * 16 + (size_t)t.nq * 8, (size_t)t.nq * 4);
  }
  return 0;
}

int hdb_profile_enable(hdb_matrix* m, int max_pairs) {
  if (!m) return fail("null handle");
  HDB_CUDA(cudaSetDevice(m->device));
  for (cudaEvent_t e : m->prof_ev) cudaEventDestroy(e);
  m->prof_ev.clear();
  m->prof_used = 0;
  for (int i = 0; i < 2 * max_pairs; ++i) {
    cudaEvent_t e;
    HDB_CUDA(cudaEventCreate(&e));
    m->prof_ev.push_back(e);
  }
  return 0;
}

int hdb_profile_read(hdb_matrix* m, int* n_launches, float* total_ms) {
  if (!m || !n_launches || !total_ms) return fail("null argument");
  HDB_CUDA(cudaSetDevice(m->device));
  HDB_TRY(join_alt(m));
  HDB_CUDA(cudaStreamSynchronize(m->stream));
  float sum = 0.f;
  for (size_t i = 0; i + 1 < m->prof_used; i += 2) {
    float ms = 0.f;
    HDB_CUDA(cudaEventElapsedTime(&ms, m->prof_ev[i], m->prof_ev[i + 1]));
    sum += ms;
  }
  *n_launches = (int)(m->prof_used / 2);
  *total_ms = sum;
  m->prof_used = 0;
  return 0;
}

int hdb_time_last_query(hdb_matrix* m, int what, int iters, float* ms_per_iter) {
  if (!m || !ms_per_iter) return fail("null argument");
  if (!m->last.valid || m->last.kp == 0 || m->last.k == 0) return fail("hdb_time_last_query: no fused query to replay");
  if (iters < 1) iters = 1;
  HDB_CUDA(cudaSetDevice(m->device));
  HDB_TRY(join_alt(m));
  const bool use_decay = m->ts != nullptr && m->decay_valid;
  const RowFilter f = filter_of(m, m->last.bias, use_decay);
  cudaEvent_t e0, e1;
  HDB_CUDA(cudaEventCreate(&e0));
  HDB_CUDA(cudaEventCreate(&e1));
  const int64_t nq = m->last.nq < kChunk ? m->last.nq : kChunk;
  MatrixView v = view_of(m);
  const int elt = (m->dtype == 2) ? 8 : 4;
  HDB_CUDA(cudaEventRecord(e0, m->stream));
  for (int it = 0; it < iters; ++it) {
    if (what == 1) {
      HDB_CUDA(cudaMemsetAsync(m->tau, 0, 8, m->stream));
      SweepOut so; so.cand = m->cand; so.tau = m->tau; so.grid = m->grid;
      (void)elt;
      HDB_TRY(launch_sweep(v, m->last.metric, m->last.rdt, m->qb.qa, m->qb.qbits, m->qb.qaux, f, m->last.kp, so, 1, m->stream));
    } else {
      HDB_TRY(run_fused(m, m->last.metric, m->last.rdt, m->last.kp, 0, nq, m->last.k, f, m->o_idx, m->o_score, m->o_count,
                        m->o_flags));
    }
  }
  HDB_CUDA(cudaEventRecord(e1, m->stream));
  HDB_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  HDB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_per_iter = ms / iters;
  return 0;
}

int hdb_scores(hdb_matrix* m, int metric, const void* query, int q_dtype, int q_space, void* out, int out_space,
               int* out_dtype) {
  return hdb_scores_ex(m, metric, query, q_dtype, q_space, out, out_space, out_dtype, 0);
}

int hdb_scores_ex(hdb_matrix* m, int metric, const void* query, int q_dtype, int q_space, void* out, int out_space,
                  int* out_dtype, int flags) {
  if (!m) return fail("null handle");
  if (!m->finalized) return fail("hdb_scores: call hdb_matrix_finalize first");
  if (metric < 0 || metric > 6) return fail("Unknown metric");
  if (q_dtype < 0 || q_dtype > 2) return fail("hdb_scores: q_dtype must be HDB_F16/F32/F64");
  if (!query || !out) return fail("hdb_scores: NULL argument");
  if (m->ord) return fail("hdb_scores: a row order is set (the full score vector is defined on the caller's row order)");
  HDB_CUDA(cudaSetDevice(m->device));
  HDB_TRY(join_alt(m));
  HDB_TRY(ensure_workspace(m, 1, 1));
  const int rdt = m->dtype > q_dtype ? m->dtype : q_dtype;
  if (metric == HDB_HAMMING || metric == HDB_JACCARD) HDB_TRY(ensure_bits(m));
  if (metric == HDB_PEARSON) HDB_TRY(ensure_pearson(m));
  const void* q_dev = query;
  if (q_space == HDB_HOST) {
    HDB_CUDA(cudaMemcpyAsync(m->q_raw, query, (size_t)m->d * dtype_size(q_dtype), cudaMemcpyHostToDevice, m->stream));
    q_dev = m->q_raw;
  }
  HDB_TRY(launch_prep_query(q_dev, q_dtype, 1, m->d, metric, m->dtype, m->words, m->qb, nullptr, m->stream));
  const bool f64_out = (metric == HDB_JACCARD || metric == HDB_PEARSON);
  const size_t esz = (metric == HDB_HAMMING || f64_out) ? 8 : (size_t)dtype_size(rdt);
  if (out_dtype) *out_dtype = (metric == HDB_HAMMING) ? 3 : (f64_out ? 2 : rdt);
  void* dst = out;
  void* tmp = nullptr;
  if (out_space == HDB_HOST) {
    HDB_CUDA(cudaMalloc(&tmp, (size_t)(m->n ? m->n : 1) * esz));
    dst = tmp;
  }
  int rc = launch_scores_out(view_of(m), metric, rdt, m->qb.qc, m->qb.qbits, m->qb.qaux, dst, (flags & HDB_SCORES_DISTANCE) ? 1 : 0, m->stream);
  if (!rc && out_space == HDB_HOST) {
    cudaError_t e = cudaMemcpyAsync(out, tmp, (size_t)m->n * esz, cudaMemcpyDeviceToHost, m->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "hdb_scores copy");
    uint32_t qf = 0;
    if (!rc) {
      e = cudaMemcpy(&qf, m->qb.qflags, 4, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) rc = cuda_fail(e, "hdb_scores flags");
    }
    if (!rc && (qf & HDB_FLAG_QUERY_NAN)) rc = fail("Vectors and query_vector should not contain NaN values.");
  }
  if (tmp) cudaFree(tmp);
  return rc;
}

int hdb_query_digest_host(const void* queries, int q_dtype, int64_t n_queries, int64_t dim, uint64_t* digest_out) {
  if (!queries || !digest_out) return fail("hdb_query_digest_host: NULL argument");
  if (q_dtype < 0 || q_dtype > 2) return fail("hdb_query_digest_host: bad dtype");
  if (n_queries < 0 || dim <= 0) return fail("hdb_query_digest_host: bad shape");
  query_digest_host(queries, q_dtype, n_queries, dim, reinterpret_cast<unsigned long long*>(digest_out));
  return 0;
}

int hdb_query_digest(hdb_matrix* m, const void* queries, int q_dtype, int q_space, int64_t n_queries, uint64_t* digest_out) {
  if (!m) return fail("null handle");
  if (!queries || !digest_out) return fail("hdb_query_digest: NULL argument");
  if (q_dtype < 0 || q_dtype > 2) return fail("hdb_query_digest: bad dtype");
  if (n_queries <= 0) return 0;
  if (q_space == HDB_HOST) {
    query_digest_host(queries, q_dtype, n_queries, m->d, reinterpret_cast<unsigned long long*>(digest_out));
    return 0;
  }
  HDB_CUDA(cudaSetDevice(m->device));
  if (m->digest_cap < n_queries) {
    if (m->digest) cudaFree(m->digest);
    m->digest = nullptr; m->digest_cap = 0;
    HDB_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->digest), (size_t)n_queries * 16));
    m->digest_cap = n_queries;
  }
  HDB_TRY(launch_query_digest(queries, q_dtype, n_queries, m->d, m->digest, m->stream));
  HDB_CUDA(cudaMemcpyAsync(digest_out, m->digest, (size_t)n_queries * 16, cudaMemcpyDeviceToHost, m->stream));
  HDB_CUDA(cudaStreamSynchronize(m->stream));
  return 0;
}

int hdb_normalize_rows(int device, int dtype, int64_t n_rows, int64_t dim, const void* src, int src_space, void* dst,
                       int dst_space) {
  if (dtype < 0 || dtype > 2) return fail("hdb_normalize_rows: bad dtype");
  if (n_rows < 0 || dim <= 0) return fail("hdb_normalize_rows: bad shape");
  if (n_rows == 0) return 0;
  HDB_CUDA(cudaSetDevice(device));
  const size_t bytes = (size_t)n_rows * dim * dtype_size(dtype);
  void *dsrc = nullptr, *ddst = nullptr;
  const void* s = src;
  void* d = dst;
  int rc = 0;
  if (src_space == HDB_HOST) {
    HDB_CUDA(cudaMalloc(&dsrc, bytes));
    cudaError_t e = cudaMemcpy(dsrc, src, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(dsrc); return cuda_fail(e, "normalize upload"); }
    s = dsrc;
  }
  if (dst_space == HDB_HOST) {
    cudaError_t e = cudaMalloc(&ddst, bytes);
    if (e != cudaSuccess) { if (dsrc) cudaFree(dsrc); return cuda_fail(e, "normalize alloc"); }
    d = ddst;
  }
  rc = launch_normalize_rows(dtype, n_rows, dim, s, d, nullptr);
  if (!rc && dst_space == HDB_HOST) {
    cudaError_t e = cudaMemcpy(dst, ddst, bytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cuda_fail(e, "normalize download");
  }
  if (dsrc) cudaFree(dsrc);
  if (ddst) cudaFree(ddst);
  return rc;
}

int hdb_merge_topk(int device, void* cuda_stream, int64_t n_lists, int64_t nq, int64_t k, int64_t list_stride, const double* scores,
                   const int64_t* ids, const int64_t* counts, int in_space, int64_t* out_idx, double* out_score,
                   int64_t* out_count, int out_space) {
  if (n_lists <= 0 || nq < 0 || k < 0) return fail("hdb_merge_topk: bad sizes");
  if (nq == 0) return 0;
  HDB_CUDA(cudaSetDevice(device));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
  if (list_stride != 0 && in_space != HDB_DEVICE) return fail("hdb_merge_topk: strided input must be device memory");
  const int64_t ls_rec = list_stride ? list_stride : nq * k;
  const int64_t ls_cnt = list_stride ? list_stride : nq;
  const size_t rec = (size_t)n_lists * nq * (k ? k : 1);
  std::vector<void*> tmp;
  auto stage_in = [&](const void* p, size_t bytes, const void** out) -> int {
    if (in_space == HDB_DEVICE) { *out = p; return 0; }
    void* d = nullptr;
    HDB_CUDA(cudaMalloc(&d, bytes ? bytes : 8));
    tmp.push_back(d);
    HDB_CUDA(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, s));
    *out = d;
    return 0;
  };
  const void *ds = nullptr, *di = nullptr, *dc = nullptr;
  int rc = stage_in(scores, rec * 8, &ds);
  if (!rc) rc = stage_in(ids, rec * 8, &di);
  if (!rc) rc = stage_in(counts, (size_t)n_lists * nq * 8, &dc);
  int64_t* oi = out_idx; double* os = out_score; int64_t* oc = out_count;
  if (!rc && out_space == HDB_HOST) {
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, (size_t)nq * (k ? k : 1) * 16 + (size_t)nq * 8);
    if (e != cudaSuccess) rc = cuda_fail(e, "merge alloc");
    else {
      tmp.push_back(d);
      oi = reinterpret_cast<int64_t*>(d);
      os = reinterpret_cast<double*>(oi + nq * (k ? k : 1));
      oc = reinterpret_cast<int64_t*>(os + nq * (k ? k : 1));
    }
  }
  if (!rc) rc = launch_merge_topk(n_lists, nq, k, ls_rec, ls_cnt, (const double*)ds, (const int64_t*)di, (const int64_t*)dc, oi, os, oc, s);
  if (!rc && out_space == HDB_HOST) {
    cudaError_t e = cudaSuccess;
    if (k) e = cudaMemcpyAsync(out_idx, oi, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && k) e = cudaMemcpyAsync(out_score, os, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_count, oc, (size_t)nq * 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) rc = cuda_fail(e, "merge download");
  } else if (!tmp.empty()) {
    cudaStreamSynchronize(s);
  }
  for (void* p : tmp) cudaFree(p);
  return rc;
}

}  // extern "C"
