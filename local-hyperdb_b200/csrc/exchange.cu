// Candidate exchange of the row-sharded path over NVLink / NVSwitch PEER MEMORY (the one exchange
// step per query -- every rank's k (score, global id) records to every rank -- followed by the G*k -> k merge).
//
// Every rank owns a small exchange buffer that its peers map through CUDA IPC:
//     [kXSlots][world][max_words] message words | arrival flags [kXSlots][world] | consumed counters [world]
// One step is at most TWO kernels and no library call:
//   push        the certify kernel itself (finalize.cu, PushTarget) stores the rank's result straight into EVERY rank's
//               slot with peer stores and its last CTA publishes a per-(slot, source) sequence flag with a system-scope
//               release -- no kernel boundary between certify and exchange.  Paths that do not end in one certify
//               launch (tensor-core batches, the exact path, k = 0) use exchange_push_kernel on the packed result block.
//   wait+merge  ONE kernel, one CTA per query, on the exchange's OWN high-priority stream: spins (system-scope acquire,
//               bounded by a timeout) until all G flags of the slot carry this step's number, merges the G lists that
//               now sit in LOCAL memory by (score desc, global id asc), hands out the per-shard certificate flags, and
//               its last CTA advances the step counter and tells every peer "slot consumed".
// Flow control is explicit: a rank stores step s into slot s % kXSlots of rank g only once g's consumed counter says
// that step s - kXSlots was merged there, so ranks are decoupled by up to kXSlots steps instead of advancing in
// lockstep behind one in-order stream.  Both step counters live in device memory: nothing depends on a host value.
#include <cstring>
#include <vector>

#include "hdb_common.cuh"
#include "hdb_exchange.cuh"
#include "hdb_internal.h"
#include "../../include/hyperdb_b200.h"

using namespace hdb;

namespace hdb {

// Unfused push: one CTA per destination rank copies the packed block [scores | ids | counts | flags] of this rank.
__global__ void __launch_bounds__(512) exchange_push_kernel(PushTarget t, const unsigned long long* mine, int64_t words) {
  __shared__ unsigned long long s_step;
  if (threadIdx.x == 0) s_step = t.ctr[kCtrPushStep];
  __syncthreads();
  const unsigned long long step = s_step;
  const int g = blockIdx.x;
  if (threadIdx.x == 0) push_wait_consumed(t, g, step);
  __syncthreads();
  unsigned long long* dst = push_slot(t, g, step);
  if ((reinterpret_cast<uintptr_t>(mine) & 15) == 0) {      // max_words is even: every slot starts on a 16-byte boundary
    const uint4* s4 = reinterpret_cast<const uint4*>(mine);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (int64_t i = threadIdx.x; i < words / 2; i += blockDim.x) d4[i] = s4[i];
    if ((words & 1) && threadIdx.x == 0) dst[words - 1] = mine[words - 1];
  } else {
    for (int64_t i = threadIdx.x; i < words; i += blockDim.x) dst[i] = mine[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) push_publish(t, step, (unsigned)gridDim.x);
}

// wait + merge, one CTA per query
__global__ void __launch_bounds__(256) exchange_wait_merge_kernel(PushTarget t, int64_t nq, int64_t k, int64_t* out_idx, double* out_score,
                                                                  int64_t* out_count, uint32_t* out_flags /*[world][nq]*/) {
  __shared__ unsigned long long s_step;
  __shared__ int s_timeout;
  if (threadIdx.x == 0) { s_step = t.ctr[kCtrWaitStep]; s_timeout = 0; }
  __syncthreads();
  const unsigned long long step = s_step;
  const int slot = (int)(step % kXSlots);
  const int world = t.world;
  const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(t.local + t.data_bytes) + slot * world;
  if ((int)threadIdx.x < world) {
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flags + threadIdx.x) < step + 1) {
      if (global_ns() - t0 > kXTimeoutNs) { s_timeout = 1; break; }                // a peer died: do not hang the GPU
      __nanosleep(100);
    }
  }
  __syncthreads();
  const bool timed_out = s_timeout != 0;
  const int64_t mw = t.max_words;
  const unsigned long long* base = reinterpret_cast<const unsigned long long*>(t.local) + (int64_t)slot * world * mw;
  const double* scores = reinterpret_cast<const double*>(base);
  const int64_t* ids = reinterpret_cast<const int64_t*>(base) + nq * k;
  const int64_t* counts = reinterpret_cast<const int64_t*>(base) + 2 * nq * k;
  const int64_t b = blockIdx.x;
  const int64_t total = (int64_t)world * k;
  int64_t have = 0;
  if (!timed_out)
    for (int l = 0; l < world; ++l) have += counts[l * mw + b];
  const int64_t kk = have < k ? have : k;
  for (int64_t i = threadIdx.x; i < k; i += blockDim.x) {
    out_idx[b * k + i] = -1;
    out_score[b * k + i] = -INFINITY;
  }
  __syncthreads();
  if (!timed_out) {
    for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
      const int64_t l = e / k, j = e % k;
      if (j >= counts[l * mw + b]) continue;
      const double sc = scores[l * mw + b * k + j];
      const int64_t r = ids[l * mw + b * k + j];
      int64_t rank = 0;
      for (int l2 = 0; l2 < world && rank < kk; ++l2) {
        const int64_t c2 = counts[l2 * mw + b];
        const double* s2 = scores + l2 * mw + b * k;
        const int64_t* i2 = ids + l2 * mw + b * k;
        for (int64_t j2 = 0; j2 < c2; ++j2) {
          const double t2 = s2[j2];
          if (t2 > sc || (t2 == sc && i2[j2] < r)) ++rank;      // (score desc, global id asc): the lower-index tie rule across shards
          else if (t2 < sc) break;                              // each list is sorted descending
        }
      }
      if (rank < kk) { out_idx[b * k + rank] = r; out_score[b * k + rank] = sc; }
    }
  }
  if (out_flags && (int)threadIdx.x < world) {
    const int g = threadIdx.x;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(base + (int64_t)g * mw + 2 * nq * k + nq);
    // a rank that never delivered: the step is unusable on every query (stale slot contents must not pass for results)
    out_flags[(int64_t)g * nq + b] = timed_out ? (HDB_FLAG_UNCERTIFIED | HDB_FLAG_EXCHANGE_ERROR) : src[b];
  }
  if (threadIdx.x == 0) {
    out_count[b] = kk;
    if (timed_out) *t.error = 1;
  }
  // ---- last CTA: advance the step counter and release the slot to every peer
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long old = atomicAdd(t.ctr + kCtrWaitDone, 1ull);
    if (old == (unsigned long long)gridDim.x - 1) {
      t.ctr[kCtrWaitDone] = 0;
      __threadfence_system();
      for (int g = 0; g < world; ++g) {
        unsigned long long* consumed = reinterpret_cast<unsigned long long*>(t.peer[g] + t.consumed_off) + t.rank;
        st_release_sys(consumed, step + 1);
      }
      t.ctr[kCtrWaitStep] = step + 1;
    }
  }
}

}  // namespace hdb

// ---------------------------------------------------------------------------------------------
struct hdb_exchange {
  int device = 0, world = 1, rank = 0;
  int64_t max_words = 0;                 // capacity of one rank's message in 8-byte words
  char* local = nullptr;
  size_t data_bytes = 0, consumed_off = 0, bytes = 0;
  std::vector<char*> peer;               // peer[g]: rank g's buffer as mapped into this process (peer[rank] == local)
  std::vector<char> ipc_opened;
  char** d_peer = nullptr;               // device copy of peer[]
  unsigned long long* d_ctr = nullptr;   // [kCtr*]: steps pushed, steps merged, CTA arrival counters
  int* d_error = nullptr;                // 1 = a wait timed out (sticky: rebuild the exchange)
  cudaStream_t xs = nullptr;             // wait + merge run here
  bool xs_owned = true;                  // false: a caller-owned stream (hdb_exchange_set_stream)
  // A push KERNEL (large batches: tensor-core path, chunked sweeps) must have run before this rank's waiters start: a
  // batch of thousands of queries means thousands of spinning wait CTAs on the high-priority stream, enough to keep the
  // push kernel off the SMs for good -- every rank then waits for every other rank's push (seen at 8 GPUs, B = 4096).
  cudaEvent_t ev_push = nullptr;
  bool push_pending = false;
  bool connected = false;
};

namespace hdb {
int exchange_push_target(hdb_exchange* x, int64_t nq, int64_t k, PushTarget* out) {
  if (!x || !x->connected) return fail("exchange: not connected");
  PushTarget t;
  t.peer = x->d_peer; t.world = x->world; t.rank = x->rank; t.max_words = x->max_words;
  t.data_bytes = x->data_bytes; t.consumed_off = x->consumed_off; t.ctr = x->d_ctr; t.local = x->local; t.error = x->d_error;
  t.nq = nq; t.k = k;
  *out = t;
  return 0;
}
int64_t exchange_max_words(const hdb_exchange* x) { return x ? x->max_words : 0; }
int exchange_device(const hdb_exchange* x) { return x ? x->device : -1; }
cudaStream_t exchange_stream(const hdb_exchange* x) { return x ? x->xs : nullptr; }

int exchange_launch_push(hdb_exchange* x, cudaStream_t s, const void* mine, int64_t words) {
  PushTarget t;
  HDB_TRY(exchange_push_target(x, 0, 0, &t));
  if (words > x->max_words || words < 1) return fail("hdb_exchange_push: message does not fit the exchange buffer");
  exchange_push_kernel<<<x->world, 512, 0, s>>>(t, reinterpret_cast<const unsigned long long*>(mine), words);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  if (!x->ev_push) HDB_CUDA(cudaEventCreateWithFlags(&x->ev_push, cudaEventDisableTiming));
  HDB_CUDA(cudaEventRecord(x->ev_push, s));
  x->push_pending = true;
  return 0;
}
int exchange_launch_wait_merge(hdb_exchange* x, cudaStream_t s, int64_t nq, int64_t k, int64_t* out_idx, double* out_score,
                               int64_t* out_count, uint32_t* out_flags) {
  PushTarget t;
  HDB_TRY(exchange_push_target(x, nq, k, &t));
  if (nq <= 0 || k < 0 || 2 * nq * k + nq + (nq + 1) / 2 > x->max_words) return fail("hdb_exchange_wait_merge: bad sizes");
  if (x->push_pending) {                 // the waiters start after this rank's own push kernel (no-op on the push's stream)
    HDB_CUDA(cudaStreamWaitEvent(s, x->ev_push, 0));
    x->push_pending = false;
  }
  exchange_wait_merge_kernel<<<(unsigned)nq, 256, 0, s>>>(t, nq, k, out_idx, out_score, out_count, out_flags);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace hdb

extern "C" {

int hdb_exchange_create(int device, int world, int rank, int64_t max_words, hdb_exchange** out) {
  if (!out) return fail("hdb_exchange_create: out is NULL");
  if (world < 1 || world > 64 || rank < 0 || rank >= world || max_words < 1) return fail("hdb_exchange_create: bad arguments");
  HDB_CUDA(cudaSetDevice(device));
  hdb_exchange* x = new hdb_exchange();
  max_words += max_words & 1;           // even: 16-byte aligned slots (every rank rounds the same way)
  x->device = device; x->world = world; x->rank = rank; x->max_words = max_words;
  x->data_bytes = (size_t)kXSlots * world * max_words * 8;
  x->consumed_off = x->data_bytes + (size_t)kXSlots * world * 8;
  x->bytes = x->consumed_off + (size_t)world * 8;
  x->peer.assign(world, nullptr);
  x->ipc_opened.assign(world, 0);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&x->local), x->bytes);
  if (e == cudaSuccess) e = cudaMemset(x->local, 0, x->bytes);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&x->d_peer), sizeof(char*) * world);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&x->d_ctr), 8 * kCtrCount);
  if (e == cudaSuccess) e = cudaMemset(x->d_ctr, 0, 8 * kCtrCount);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&x->d_error), 4);
  if (e == cudaSuccess) e = cudaMemset(x->d_error, 0, 4);
  if (e == cudaSuccess) {
    // wait + merge must get an SM the moment the flags arrive, ahead of pending sweep CTAs: highest priority
    int least = 0, greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess ||
        cudaStreamCreateWithPriority(&x->xs, cudaStreamNonBlocking, greatest) != cudaSuccess) {
      cudaGetLastError();
      e = cudaStreamCreateWithFlags(&x->xs, cudaStreamNonBlocking);
    }
  }
  if (e != cudaSuccess) { int rc = cuda_fail(e, "hdb_exchange_create"); delete x; return rc; }
  x->peer[rank] = x->local;
  *out = x;
  return 0;
}

int hdb_exchange_destroy(hdb_exchange* x) {
  if (!x) return 0;
  cudaSetDevice(x->device);
  cudaDeviceSynchronize();
  for (int g = 0; g < x->world; ++g)
    if (x->ipc_opened[g] && x->peer[g]) cudaIpcCloseMemHandle(x->peer[g]);
  if (x->xs && x->xs_owned) cudaStreamDestroy(x->xs);
  if (x->ev_push) cudaEventDestroy(x->ev_push);
  if (x->local) cudaFree(x->local);
  if (x->d_peer) cudaFree(x->d_peer);
  if (x->d_ctr) cudaFree(x->d_ctr);
  if (x->d_error) cudaFree(x->d_error);
  delete x;
  return 0;
}

int hdb_exchange_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int hdb_exchange_local_handle(hdb_exchange* x, void* handle_out) {
  if (!x || !handle_out) return fail("hdb_exchange_local_handle: NULL argument");
  HDB_CUDA(cudaSetDevice(x->device));
  cudaIpcMemHandle_t h;
  HDB_CUDA(cudaIpcGetMemHandle(&h, x->local));
  memcpy(handle_out, &h, sizeof(h));
  return 0;
}

static int finish_connect(hdb_exchange* x) {
  for (int g = 0; g < x->world; ++g)
    if (!x->peer[g]) return fail("hdb_exchange_connect: a peer buffer is missing");
  HDB_CUDA(cudaMemcpy(x->d_peer, x->peer.data(), sizeof(char*) * x->world, cudaMemcpyHostToDevice));
  x->connected = true;
  return 0;
}

int hdb_exchange_connect(hdb_exchange* x, const void* all_handles) {
  if (!x || !all_handles) return fail("hdb_exchange_connect: NULL argument");
  HDB_CUDA(cudaSetDevice(x->device));
  const char* hs = reinterpret_cast<const char*>(all_handles);
  for (int g = 0; g < x->world; ++g) {
    if (g == x->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, hs + (size_t)g * sizeof(h), sizeof(h));
    void* p = nullptr;
    HDB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    x->peer[g] = reinterpret_cast<char*>(p);
    x->ipc_opened[g] = 1;
  }
  return finish_connect(x);
}

int hdb_exchange_connect_pointers(hdb_exchange* x, void* const* peer_buffers) {
  if (!x || !peer_buffers) return fail("hdb_exchange_connect_pointers: NULL argument");
  HDB_CUDA(cudaSetDevice(x->device));
  for (int g = 0; g < x->world; ++g)
    if (g != x->rank) x->peer[g] = reinterpret_cast<char*>(peer_buffers[g]);
  return finish_connect(x);
}

int hdb_exchange_local_buffer(hdb_exchange* x, void** buffer) {
  if (!x || !buffer) return fail("hdb_exchange_local_buffer: NULL argument");
  *buffer = x->local;
  return 0;
}

int hdb_exchange_stream(hdb_exchange* x, void** cuda_stream) {
  if (!x || !cuda_stream) return fail("hdb_exchange_stream: NULL argument");
  *cuda_stream = x->xs;
  return 0;
}

int hdb_exchange_set_stream(hdb_exchange* x, void* cuda_stream) {
  if (!x || !cuda_stream) return fail("hdb_exchange_set_stream: NULL argument");
  HDB_CUDA(cudaSetDevice(x->device));
  if (x->xs) {
    HDB_CUDA(cudaStreamSynchronize(x->xs));
    if (x->xs_owned) cudaStreamDestroy(x->xs);
  }
  x->xs = reinterpret_cast<cudaStream_t>(cuda_stream);
  x->xs_owned = false;
  return 0;
}

int hdb_exchange_push(hdb_exchange* x, void* cuda_stream, const void* mine, int64_t words) {
  if (!x || !mine) return fail("hdb_exchange_push: NULL argument");
  if (!x->connected) return fail("hdb_exchange_push: not connected");
  HDB_CUDA(cudaSetDevice(x->device));
  return exchange_launch_push(x, reinterpret_cast<cudaStream_t>(cuda_stream), mine, words);
}

int hdb_exchange_wait_merge(hdb_exchange* x, void* cuda_stream, int64_t nq, int64_t k, int64_t* out_idx, double* out_score,
                            int64_t* out_count, uint32_t* out_flags) {
  if (!x || !out_count) return fail("hdb_exchange_wait_merge: NULL argument");
  if (!x->connected) return fail("hdb_exchange_wait_merge: not connected");
  HDB_CUDA(cudaSetDevice(x->device));
  return exchange_launch_wait_merge(x, reinterpret_cast<cudaStream_t>(cuda_stream), nq, k, out_idx, out_score, out_count, out_flags);
}

int hdb_exchange_collect_async(hdb_exchange* x, int64_t nq, int64_t k, int64_t* out_idx, double* out_score, int64_t* out_count,
                               uint32_t* out_flags) {
  if (!x || !out_count) return fail("hdb_exchange_collect_async: NULL argument");
  if (!x->connected) return fail("hdb_exchange_collect_async: not connected");
  HDB_CUDA(cudaSetDevice(x->device));
  return exchange_launch_wait_merge(x, x->xs, nq, k, out_idx, out_score, out_count, out_flags);
}

int hdb_exchange_step(hdb_exchange* x, void* cuda_stream, const void* mine, int64_t words, int64_t nq, int64_t k, int64_t* out_idx,
                      double* out_score, int64_t* out_count, uint32_t* out_flags) {
  if (words < 2 * nq * k + nq) return fail("hdb_exchange_step: message shorter than its own records");
  HDB_TRY(hdb_exchange_push(x, cuda_stream, mine, words));
  return hdb_exchange_wait_merge(x, cuda_stream, nq, k, out_idx, out_score, out_count, out_flags);
}

int hdb_exchange_error(hdb_exchange* x, int* error) {
  if (!x || !error) return fail("hdb_exchange_error: NULL argument");
  HDB_CUDA(cudaSetDevice(x->device));
  HDB_CUDA(cudaMemcpy(error, x->d_error, 4, cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
