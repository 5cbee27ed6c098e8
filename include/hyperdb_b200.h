/*
 * hyperdb_b200.h -- C ABI of the B200-native brute-force ranking engine for local-hyperDB.
 *
 * The reference has no FFI: its boundary for this path is the Python module
 * hyperdb/ranking_algorithm.py (imported at hyperdb/hyperdb.py:13) and its single call site
 * hyperdb/hyperdb.py:1556-1558.  Each entry point below names the reference lines it replaces.
 * The Python host (local-hyperdb_b200/hyperdb_b200/) binds these with ctypes; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; hdb_last_error() returns a
 *     thread-local message (mapped to ValueError / RuntimeError by the host).
 *   - pointers are plain host or device addresses; `*_space` says which (HDB_HOST / HDB_DEVICE).
 *   - all work is enqueued on the handle's stream (default: the legacy default stream, which is
 *     also PyTorch's default current stream).  Calls with host outputs synchronise that stream
 *     before returning; calls with device outputs do not synchronise.
 *   - one caller thread per handle.  There is no CPU fallback: without a CUDA device every
 *     compute entry point fails.
 */
#ifndef HYPERDB_B200_H
#define HYPERDB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hdb_matrix hdb_matrix; /* opaque: one row shard of the stored matrix on one GPU */
typedef struct hdb_exchange hdb_exchange; /* opaque: this rank's end of the peer-memory candidate exchange */

enum { HDB_F16 = 0, HDB_F32 = 1, HDB_F64 = 2 };   /* HyperDB fp_precision, hyperdb/hyperdb.py:65-66,80 */
enum { HDB_HOST = 0, HDB_DEVICE = 1 };
enum {                                              /* metric dispatch, hyperdb/ranking_algorithm.py:155-163 */
  HDB_DOT = 0,        /* dot_product        :24-30  */
  HDB_COSINE = 1,     /* cosine_similarity  :32-42  */
  HDB_EUCLIDEAN = 2,  /* euclidean_metric   :44-52  */
  HDB_MANHATTAN = 3,  /* manhattan_distance :54-61  */
  HDB_HAMMING = 4,    /* hamming_distance   :128-147 */
  HDB_JACCARD = 5,    /* jaccard_similarity :63-76 (same packed sign bits as hamming) */
  HDB_PEARSON = 6     /* pearson_correlation :78-113 (per-row np.mean / np.std columns, built on first use) */
};
/* bits of the per-query flags word written by hdb_query */
enum {
  HDB_FLAG_FALLBACK = 1,   /* the fused select did not certify; the exact full-vector path produced the result */
  HDB_FLAG_QUERY_NAN = 2,  /* the query holds a NaN (reference raises ValueError, ranking_algorithm.py:150-151) */
  HDB_FLAG_TENSOR = 4,     /* candidates came from the tcgen05 batched contraction */
  HDB_FLAG_UNCERTIFIED = 8, /* device-output mode only: the certificate failed and nothing was recomputed;
                               the caller must repeat the query with host outputs or path mode 1 */
  HDB_FLAG_EXCHANGE_ERROR = 16 /* sharded path: a rank did not deliver its candidates in time; the step holds no results
                                  (count 0) and the exchange must be rebuilt */
};

const char* hdb_last_error(void);
int hdb_version(void);
int hdb_device_count(int* count);

/* ---- storage: replaces `self.vectors` (hyperdb/hyperdb.py:127-135, :911) ------------------ */
/* A shard of n_rows x dim elements of `dtype`, row-major, on CUDA device `device`.
 * row_offset = global row id of the shard's first row (0 on a single GPU). */
int hdb_matrix_create(int device, int dtype, int64_t n_rows, int64_t dim, int64_t row_offset,
                      hdb_matrix** out);
int hdb_matrix_destroy(hdb_matrix* m);
/* Copy rows [row_start, row_start+n_rows) from host or device memory (same dtype, row-major). */
int hdb_matrix_upload(hdb_matrix* m, int64_t row_start, int64_t n_rows, const void* src, int src_space);
/* Use caller-owned device memory (n_rows x dim, row-major, 16-byte aligned) without copying. */
int hdb_matrix_adopt(hdb_matrix* m, void* device_rows);
/* Ingest pass: per-row norms exactly as get_norm_vector computes them (ranking_algorithm.py:8-21),
 * NaN scan (replaces the per-query scan of :150), error-bound statistics.  Must be called after the
 * last upload and before the first query.  Fails if the matrix holds a NaN. */
int hdb_matrix_finalize(hdb_matrix* m);
/* Run the handle's work on this cudaStream_t (NULL = legacy default stream). */
int hdb_matrix_set_stream(hdb_matrix* m, void* cuda_stream);
/* Query pipelining for device-output calls.  With a post stream set, hdb_query(..., HDB_DEVICE) enqueues the query
 * preparation on an internal high-priority stream, the streaming sweep on the handle's stream (or its internal twin,
 * see hdb_matrix_set_sweep_overlap) and the certify step on `post_stream`, so that the certify (and whatever the
 * caller enqueues after it on the post stream: the candidate exchange, the merge) of query i overlaps the sweep of
 * query i+1.  Ordering contract: a device-resident query is read only after all work enqueued on the HANDLE'S stream
 * before the call (the internal stream waits for an event recorded there); a query produced on any other stream must
 * be complete, or that stream must be joined into the handle's stream, before hdb_query is called.  Results are valid
 * in POST-stream order.  NULL switches pipelining off.
 * Give the post stream a HIGH priority (cudaStreamCreateWithPriority): its small kernels then take the SMs a finishing
 * sweep vacates before the next query's pending sweep CTAs do. */
int hdb_matrix_set_post_stream(hdb_matrix* m, void* post_stream);
/* Pipelined mode only: let the sweeps of consecutive queries overlap (default on).  Odd-numbered queries sweep on an
 * internal second stream, so the next query's CTAs fill the SMs the current one vacates during its tail.  Results are
 * still ordered on the post stream. */
int hdb_matrix_set_sweep_overlap(hdb_matrix* m, int on);
int hdb_matrix_info(const hdb_matrix* m, int* dtype, int64_t* n_rows, int64_t* dim, int64_t* row_offset,
                    int64_t* n_kept);

/* ---- mutation of a resident shard: replaces np.concatenate on add (hyperdb/hyperdb.py:504-509) and the
 *      np.vstack / boolean-mask copies of remove_document (:718-728).  Owning shards only (not adopted memory).
 *      The row mask, the kept range, the decay column and (append only) the timestamps are reset: set them again. */
/* Make room for `capacity_rows` rows in the matrix and every per-row column (amortises appends). */
int hdb_matrix_reserve(hdb_matrix* m, int64_t capacity_rows);
/* Append n_rows rows (same dtype, row-major, host or device memory) after the last row.  Norms, sign bits and pearson
 * statistics are computed for the NEW rows only.  Fails, leaving the shard unchanged, if the new rows hold a NaN. */
int hdb_matrix_append(hdb_matrix* m, int64_t n_rows, const void* src, int src_space);
/* Remove the listed LOCAL rows (any order, duplicates allowed); the remaining rows keep their relative order
 * (stable in-place compaction on the device), so local row i of the result is the i-th surviving row. */
int hdb_matrix_remove_rows(hdb_matrix* m, const int64_t* local_rows, int64_t count, int src_space);
/* Global row id of the shard's first row (changes when an earlier shard grows or shrinks). */
int hdb_matrix_set_row_offset(hdb_matrix* m, int64_t row_offset);

/* ---- row subset: replaces the filters' output (hyperdb/hyperdb.py:1119-1134, :1218-1308) ---- */
/* Keep only rows whose bit is set (bit i of word i/32, LSB first, local row ids); NULL keeps all. */
int hdb_matrix_set_mask(hdb_matrix* m, const uint32_t* bits, int src_space);
/* Row order (replaces nothing in the reference, which filters Python lists: hyperdb/hyperdb.py:1218-1257 builds the kept
 * rows per query): the shard's rows are STORED in another order than the caller numbers them -- typically clustered by a
 * metadata key at ingest, so that a metadata filter keeps a few contiguous runs of rows and the masked sweep streams them
 * at the full HBM rate instead of skipping every other 768-byte row.  order[p] = the caller's local index of physical row
 * p (a permutation of 0 .. n_rows-1; host or device).  Reported ids are row_offset + order[p] and ties resolve on them
 * (lower index first), exactly as if the rows were stored in the caller's order; mask bits, the kept range, timestamps
 * and uploads stay in PHYSICAL order.  NULL clears.  With an order set the shard cannot be mutated, hdb_scores is
 * refused and batches run on the streaming sweeps (no tensor-core path). */
int hdb_matrix_set_row_order(hdb_matrix* m, const uint32_t* order, int src_space);
/* Keep only local rows in [lo, hi) (apply_skip_doc keeps one contiguous range); (0, n_rows) keeps all. */
int hdb_matrix_set_range(hdb_matrix* m, int64_t lo, int64_t hi);

/* ---- time decay: recency_bias * exp(ts - max ts), ranking_algorithm.py:179-183 --------------- */
/* Per-row timestamps (float64, local rows); NULL removes them. */
int hdb_matrix_set_timestamps(hdb_matrix* m, const double* ts, int src_space);
/* max of the timestamps over KEPT rows of this shard (hyperdb/hyperdb.py:1334-1344) and their count;
 * the host all-reduces (MAX) these over shards. */
int hdb_matrix_kept_ts_max(hdb_matrix* m, double* ts_max, int64_t* n_kept);
/* Build the decay column exp(ts - ts_max) with the GLOBAL maximum. */
int hdb_matrix_set_decay_reference(hdb_matrix* m, double ts_max);
/* Stage 1 of the HyperDB.query composition: replace ts by bias1*exp(ts - ts_max)
 * in place on the device (what _handle_timestamps returns, hyperdb/hyperdb.py:1344-1346). */
int hdb_matrix_stage1_recency(hdb_matrix* m, double bias1, double ts_max);

/* ---- the hot path: hyperDB_ranking_algorithm_sort, ranking_algorithm.py:149-204 -------------- */
/* n_queries independent queries (row-major n_queries x dim, dtype q_dtype) against the shard.
 * For query b: out_count[b] = min(top_k, kept rows); out_idx[b*top_k + j] = GLOBAL row id and
 * out_score[b*top_k + j] = float64 score of rank j, ordered by (score desc, row id asc).
 * Scores are the reference's arithmetic (see DESIGN.md "canonical scores"); NaN similarities (jaccard 0/0, pearson of a
 * constant row or query) rank last as -inf (ranking_algorithm.py:174).  metric is one of HDB_DOT .. HDB_PEARSON.
 * out_flags may be NULL.  All outputs live in `out_space`.  With device outputs nothing is synchronised and a query
 * whose certificate failed only carries HDB_FLAG_UNCERTIFIED; with host outputs it is repaired before the call returns. */
int hdb_query(hdb_matrix* m, int metric, const void* queries, int q_dtype, int q_space,
              int64_t n_queries, int64_t top_k, double recency_bias,
              int64_t* out_idx, double* out_score, int64_t* out_count, uint32_t* out_flags,
              int out_space);

/* ---- asynchronous host API: the same query without a synchronisation per call --------------------------------------
 * hdb_query_submit enqueues the query (host or device `queries`), the copy of the result block to pinned host memory and
 * a completion event, and returns a ticket; up to 4 tickets may be in flight, so that (with a post stream set, see
 * hdb_matrix_set_post_stream) the host-to-device copy, the sweep, the certify step and the device-to-host copy of
 * consecutive queries overlap.  hdb_query_collect waits for one ticket and
 * writes out_idx / out_score [n_queries x top_k], out_count [n_queries] and out_flags.  Nothing is repaired here: a query
 * whose flags carry HDB_FLAG_UNCERTIFIED must be repeated through hdb_query with host outputs (the Python host does).
 * world = 1: results of this shard, out_flags [n_queries].  world > 1 (an exchange is attached, see
 * hdb_matrix_attach_exchange): every rank submits the same batch; the ticket completes with the MERGED top-k of all
 * shards and out_flags [world][n_queries] holds every shard's flags. */
int hdb_query_submit(hdb_matrix* m, int metric, const void* queries, int q_dtype, int q_space, int64_t n_queries,
                     int64_t top_k, double recency_bias, int world, int64_t* ticket);
int hdb_query_collect(hdb_matrix* m, int64_t ticket, int64_t* out_idx, double* out_score, int64_t* out_count,
                      uint32_t* out_flags);

/* ---- full similarity vector: the metric functions themselves, ranking_algorithm.py:24-113,:128-147 */
/* out holds n_rows values of the NumPy result dtype promote(matrix dtype, q_dtype)
 * (uint64 for HDB_HAMMING, float64 for HDB_JACCARD and HDB_PEARSON, NaN where the reference returns NaN),
 * reported through *out_dtype (HDB_F16/F32/F64; 3 = uint64). */
int hdb_scores(hdb_matrix* m, int metric, const void* query, int q_dtype, int q_space,
               void* out, int out_space, int* out_dtype);
/* The same with option bits.  HDB_SCORES_DISTANCE: for HDB_EUCLIDEAN write the distance np.linalg.norm(v - q) itself
 * instead of 1/(1+d) -- euclidean_metric(..., get_similarity_score=False), ranking_algorithm.py:49-52 (ignored for
 * the other metrics). */
enum { HDB_SCORES_DISTANCE = 1 };
int hdb_scores_ex(hdb_matrix* m, int metric, const void* query, int q_dtype, int q_space,
                  void* out, int out_space, int* out_dtype, int flags);
/* L2-normalised copy of `n_rows` x `dim` values (get_norm_vector, ranking_algorithm.py:8-21). */
int hdb_normalize_rows(int device, int dtype, int64_t n_rows, int64_t dim, const void* src, int src_space,
                       void* dst, int dst_space);

/* ---- multi-GPU: final merge of the all-gathered per-shard candidates --- */
/* n_lists candidate lists (one per shard) for nq queries: list l has scores at scores + l*list_stride
 * laid out [query][k], ids at ids + l*list_stride, counts at counts + l*list_stride laid out [query]
 * (strides in 8-byte elements; list_stride = 0 means dense arrays [list][query][k] / [list][query]).
 * The strided form lets one all-gathered buffer [list][scores | ids | counts ...] be merged in place.
 * Writes the merged top-k per query ordered by (score desc, id asc). */
int hdb_merge_topk(int device, void* cuda_stream, int64_t n_lists, int64_t n_queries, int64_t k, int64_t list_stride,
                   const double* scores, const int64_t* ids, const int64_t* counts, int in_space,
                   int64_t* out_idx, double* out_score, int64_t* out_count, int out_space);

/* ---- multi-GPU: the same exchange + merge over NVLink peer memory, without a collective library call -------------
 * One process per GPU.  Every rank creates an exchange (a small device buffer of 4 slots x world x max_words 8-byte
 * words plus arrival flags and per-reader consumed counters), publishes its CUDA IPC handle (hdb_exchange_handle_bytes() bytes) to the other ranks by
 * any means (the host uses torch.distributed once, at set-up) and connects to all of them.  hdb_exchange_step then
 * enqueues on `cuda_stream`: peer stores of this rank's packed result block
 *     [scores nq*k (f64) | ids nq*k (i64) | counts nq (i64) | flags nq (u32, padded to 8 bytes)]      (`words` words)
 * into every rank's buffer, a bounded wait until all `world` blocks of this step have arrived, and the merge of the
 * world x k candidates per query by (score desc, global id asc).  out_flags receives [world][nq] per-shard flags.
 * Every rank must call hdb_exchange_step the same number of times with the same (nq, k). */
int hdb_exchange_create(int device, int world, int rank, int64_t max_words, hdb_exchange** out);
/* Fused form: after attaching, every device-output hdb_query on `m` (outputs laid out as ONE packed block
 * [scores | ids | counts | flags]) also delivers its results to every rank of `x` -- stored by the certify kernel itself
 * where the batch ends in one certify launch, by a push kernel otherwise -- and the caller only enqueues the second half
 * with hdb_exchange_collect_async (or uses hdb_query_submit / hdb_query_collect).  NULL detaches. */
int hdb_matrix_attach_exchange(hdb_matrix* m, hdb_exchange* x);
/* wait + merge of the current step on the exchange's own high-priority stream (hdb_exchange_stream). */
int hdb_exchange_collect_async(hdb_exchange* x, int64_t n_queries, int64_t k, int64_t* out_idx, double* out_score,
                               int64_t* out_count, uint32_t* out_flags);
int hdb_exchange_stream(hdb_exchange* x, void** cuda_stream);
/* Run wait + merge on a CALLER-OWNED stream instead (it must outlive the exchange; give it the highest priority).  For
 * hosts whose allocator tracks stream uses (torch's record_stream): the exchange's internal stream dies with
 * hdb_exchange_destroy, a recorded use of it must not.  Call before the first step. */
int hdb_exchange_set_stream(hdb_exchange* x, void* cuda_stream);
int hdb_exchange_destroy(hdb_exchange* x);
int hdb_exchange_handle_bytes(void);
int hdb_exchange_local_handle(hdb_exchange* x, void* handle_out);
/* all_handles: world consecutive IPC handles, indexed by rank (this rank's own entry is ignored). */
int hdb_exchange_connect(hdb_exchange* x, const void* all_handles);
/* Same-process form (several shards of one process, tests): peers given as device pointers from hdb_exchange_local_buffer. */
int hdb_exchange_connect_pointers(hdb_exchange* x, void* const* peer_buffers);
int hdb_exchange_local_buffer(hdb_exchange* x, void** buffer);
int hdb_exchange_step(hdb_exchange* x, void* cuda_stream, const void* mine, int64_t words, int64_t n_queries, int64_t k,
                      int64_t* out_idx, double* out_score, int64_t* out_count, uint32_t* out_flags);
/* The two halves of a step (a process that drives several ranks must enqueue every rank's push before any wait:
 * a waiting kernel may sit in front of another stream's push in the same hardware queue). */
int hdb_exchange_push(hdb_exchange* x, void* cuda_stream, const void* mine, int64_t words);
int hdb_exchange_wait_merge(hdb_exchange* x, void* cuda_stream, int64_t n_queries, int64_t k, int64_t* out_idx, double* out_score,
                            int64_t* out_count, uint32_t* out_flags);
/* 1 if a wait gave up after 10 s because a peer never delivered (synchronises the device).  The step that timed out
 * carries HDB_FLAG_UNCERTIFIED | HDB_FLAG_EXCHANGE_ERROR in every out_flags entry and count 0; the error is sticky and
 * the ranks' step counters may disagree afterwards: destroy and rebuild the exchange. */
int hdb_exchange_error(hdb_exchange* x, int* error);

/* ---- query-side front end ------------------------------------------ */
/* 128-bit digest of each query's VALUES (as float64, -0.0 == +0.0): the key of HyperDB's query cache, replacing
 * `tuple(query_input.tolist())` (hyperdb/hyperdb.py:1368-1379) -- a CUDA-tensor query is hashed where it lives (one tiny
 * kernel, 16 bytes per query back to the host), a host query on the host with the same function.  digest_out: HOST,
 * 2 words per query.  Device queries are read on the handle's stream. */
int hdb_query_digest(hdb_matrix* m, const void* queries, int q_dtype, int q_space, int64_t n_queries, uint64_t* digest_out);
/* The same digest of HOST queries without a handle (no device is touched): n_queries rows of `dim` values. */
int hdb_query_digest_host(const void* queries, int q_dtype, int64_t n_queries, int64_t dim, uint64_t* digest_out);

/* ---- instrumentation --------------------------------------------------------------------------- */
/* Kernel launches issued by this library since the last reset (bench.py's gpu_launches). */
int64_t hdb_launch_count(int reset);
/* Time `iters` back-to-back repetitions of the last hdb_query configuration with CUDA events on the
 * handle's stream (device-resident inputs/outputs); returns the mean milliseconds per repetition.
 * what: 0 = the whole device-side query (prep + select + certify), 1 = the dominant kernel only
 * (the streaming sweep, or the batched contraction). */
int hdb_time_last_query(hdb_matrix* m, int what, int iters, float* ms_per_iter);
/* Per-launch timing of the dominant kernel INSIDE normal hdb_query calls: after
 * hdb_profile_enable(m, max_pairs) every launch of the streaming sweep (or batched contraction) is
 * bracketed by a CUDA event pair on the handle's stream (up to max_pairs; 0 disables);
 * hdb_profile_read synchronises, returns the number of recorded launches and the sum of their
 * durations in milliseconds, and resets the recorder. */
int hdb_profile_enable(hdb_matrix* m, int max_pairs);
int hdb_profile_read(hdb_matrix* m, int* n_launches, float* total_ms);
/* Force a path: 0 = automatic, 1 = always the exact full-vector path, 2 = fused sweep only (fail instead of falling
 * back), 3 = tensor-core batched path when eligible, 4 = streaming sweep with the WIDE candidate class (128 candidates;
 * the first repair step for a query the automatic path could not certify). */
int hdb_matrix_set_path(hdb_matrix* m, int mode);
/* The reference ranks one query per call (hyperdb/ranking_algorithm.py:149-204 is entered once per query, the metric
 * functions :24-147 read the whole matrix each time).  Here
 * small batches on the streaming path share ONE read of the matrix between up to 8 queries (multi-query sweep; the
 * group size follows from the shape: 8 for fp16/fp32 rows with top_k <= 16, 4 for fp64 rows, top_k <= 100 and the
 * bit-packed metrics).  This caps the group for A/B measurements and tests: 1 = one pass per query, 0 = no cap. */
int hdb_matrix_set_max_group(hdb_matrix* m, int max_queries_per_pass);

#ifdef __cplusplus
}
#endif
#endif /* HYPERDB_B200_H */
