// Candidate exchange of the row-sharded path over NVLink / NVSwitch PEER MEMORY (the one exchange
// step per query -- every rank's k (score, global id) records to every rank -- followed by the G*k -> k merge).
//
// The NCCL form (torch.distributed.all_gather_into_tensor + hdb_merge_topk) costs ~17 us of host time per collective
// plus the Python around it, and a sharded step at 8 GPUs is a ~0.3 ms sweep: the step was HOST-bound (measured,
// profiles/r01b_bench/mg_diag_2gpu.log: 137 us of host work per step at 2 GPUs).  Here every rank owns a small exchange buffer
// that its peers map through CUDA IPC; one step is two tiny kernels on the post stream and no library call:
//   push : copies this rank's packed result block straight into EVERY rank's buffer with peer stores, then publishes
//          a per-(slot, source) sequence flag with a system-scope release;
//   wait : spins (system-scope acquire, bounded by a timeout) until all G flags of the slot carry this step's number,
//          copies the per-shard certificate flags out and advances the device-side step counter;
//   merge: the existing merge_topk kernel over the G lists that now sit in LOCAL memory.
// The step number lives in device memory, so the same three launches can be replayed from a CUDA graph.
// Slot reuse: a rank pushes step s+1 only after its own wait(s) returned, i.e. after every peer pushed step s, i.e.
// after every peer finished wait(s-1); with kSlots >= 3 a slot is never overwritten before its reader is done.
#include <vector>

#include "hdb_common.cuh"
#include "hdb_internal.h"
#include "../../include/hyperdb_b200.h"

using namespace hdb;

constexpr int kXSlots = 4;

struct hdb_exchange {
  int device = 0, world = 1, rank = 0;
  int64_t max_words = 0;                 // capacity of one rank's message in 8-byte words
  char* local = nullptr;                 // [kXSlots][world][max_words] words | flags [kXSlots][world] u64
  size_t data_bytes = 0, bytes = 0;
  std::vector<char*> peer;               // peer[g]: rank g's buffer as mapped into this process (peer[rank] == local)
  std::vector<char> ipc_opened;
  char** d_peer = nullptr;               // device copy of peer[]
  unsigned long long* d_step = nullptr;  // steps completed on this rank
  int* d_error = nullptr;                // 1 = a wait timed out
  bool connected = false;
};

namespace hdb {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// one CTA per destination rank
__global__ void __launch_bounds__(512) exchange_push_kernel(char* const* peer, int world, int rank, int64_t max_words, size_t data_bytes,
                                                            const unsigned long long* step_ptr, const unsigned long long* mine, int64_t words) {
  const unsigned long long step = *step_ptr;
  const int slot = (int)(step % kXSlots);
  char* dst_base = peer[blockIdx.x];
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(dst_base) + ((int64_t)slot * world + rank) * max_words;
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
  if (threadIdx.x == 0) {
    unsigned long long* flag = reinterpret_cast<unsigned long long*>(dst_base + data_bytes) + (slot * world + rank);
    st_release_sys(flag, step + 1);
  }
}

// single CTA: wait for the G pushes of this step, hand the per-shard flags to the caller, advance the step counter
__global__ void __launch_bounds__(128) exchange_wait_kernel(char* local, int world, int64_t max_words, size_t data_bytes,
                                                            unsigned long long* step_ptr, int* error, int64_t nq, int64_t k,
                                                            uint32_t* out_flags /*[world][nq]*/) {
  const unsigned long long step = *step_ptr;
  const int slot = (int)(step % kXSlots);
  const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(local + data_bytes) + slot * world;
  if (threadIdx.x < world) {
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flags + threadIdx.x) < step + 1) {
      if (global_ns() - t0 > 10ull * 1000 * 1000 * 1000) { *error = 1; break; }     // a peer died: do not hang the GPU
      __nanosleep(200);
    }
  }
  __syncthreads();
  if (out_flags) {
    for (int64_t i = threadIdx.x; i < world * nq; i += blockDim.x) {
      const int64_t g = i / nq, b = i % nq;
      const uint32_t* src = reinterpret_cast<const uint32_t*>(reinterpret_cast<const unsigned long long*>(local) +
                                                              ((int64_t)slot * world + g) * max_words + 2 * nq * k + nq);
      out_flags[i] = src[b];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *step_ptr = step + 1;
}

// the merge reads its lists relative to the CURRENT slot: a tiny kernel resolves the slot so the launch needs no host value
__global__ void exchange_merge_kernel(const char* local, int world, int64_t max_words, const unsigned long long* step_ptr, int64_t nq, int64_t k,
                                      int64_t* out_idx, double* out_score, int64_t* out_count) {
  // runs AFTER exchange_wait_kernel, which already advanced the counter: the slot just completed is step - 1
  const unsigned long long step = *step_ptr - 1;
  const int slot = (int)(step % kXSlots);
  const unsigned long long* base = reinterpret_cast<const unsigned long long*>(local) + (int64_t)slot * world * max_words;
  const double* scores = reinterpret_cast<const double*>(base);
  const int64_t* ids = reinterpret_cast<const int64_t*>(base) + nq * k;
  const int64_t* counts = reinterpret_cast<const int64_t*>(base) + 2 * nq * k;
  const int64_t b = blockIdx.x;
  const int64_t total = (int64_t)world * k;
  int64_t have = 0;
  for (int l = 0; l < world; ++l) have += counts[l * max_words + b];
  const int64_t kk = have < k ? have : k;
  for (int64_t i = threadIdx.x; i < k; i += blockDim.x) {
    out_idx[b * k + i] = -1;
    out_score[b * k + i] = -INFINITY;
  }
  __syncthreads();
  for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
    const int64_t l = e / k, j = e % k;
    if (j >= counts[l * max_words + b]) continue;
    const double t = scores[l * max_words + b * k + j];
    const int64_t r = ids[l * max_words + b * k + j];
    int64_t rank = 0;
    for (int l2 = 0; l2 < world && rank < kk; ++l2) {
      const int64_t c2 = counts[l2 * max_words + b];
      const double* s2 = scores + l2 * max_words + b * k;
      const int64_t* i2 = ids + l2 * max_words + b * k;
      for (int64_t j2 = 0; j2 < c2; ++j2) {
        const double t2 = s2[j2];
        if (t2 > t || (t2 == t && i2[j2] < r)) ++rank;      // (score desc, global id asc): the lower-index tie rule across shards
        else if (t2 < t) break;                            // each list is sorted descending
      }
    }
    if (rank < kk) { out_idx[b * k + rank] = r; out_score[b * k + rank] = t; }
  }
  if (threadIdx.x == 0) out_count[b] = kk;
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
  x->bytes = x->data_bytes + (size_t)kXSlots * world * 8;
  x->peer.assign(world, nullptr);
  x->ipc_opened.assign(world, 0);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&x->local), x->bytes);
  if (e == cudaSuccess) e = cudaMemset(x->local, 0, x->bytes);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&x->d_peer), sizeof(char*) * world);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&x->d_step), 8);
  if (e == cudaSuccess) e = cudaMemset(x->d_step, 0, 8);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&x->d_error), 4);
  if (e == cudaSuccess) e = cudaMemset(x->d_error, 0, 4);
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
  if (x->local) cudaFree(x->local);
  if (x->d_peer) cudaFree(x->d_peer);
  if (x->d_step) cudaFree(x->d_step);
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

int hdb_exchange_push(hdb_exchange* x, void* cuda_stream, const void* mine, int64_t words) {
  if (!x || !mine) return fail("hdb_exchange_push: NULL argument");
  if (!x->connected) return fail("hdb_exchange_push: not connected");
  if (words > x->max_words || words < 1) return fail("hdb_exchange_push: message does not fit the exchange buffer");
  HDB_CUDA(cudaSetDevice(x->device));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
  exchange_push_kernel<<<x->world, 512, 0, s>>>(x->d_peer, x->world, x->rank, x->max_words, x->data_bytes, x->d_step,
                                                reinterpret_cast<const unsigned long long*>(mine), words);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

int hdb_exchange_wait_merge(hdb_exchange* x, void* cuda_stream, int64_t nq, int64_t k, int64_t* out_idx, double* out_score,
                            int64_t* out_count, uint32_t* out_flags) {
  if (!x || !out_count) return fail("hdb_exchange_wait_merge: NULL argument");
  if (!x->connected) return fail("hdb_exchange_wait_merge: not connected");
  if (nq <= 0 || k < 0 || 2 * nq * k + nq > x->max_words) return fail("hdb_exchange_wait_merge: bad sizes");
  HDB_CUDA(cudaSetDevice(x->device));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
  exchange_wait_kernel<<<1, 128, 0, s>>>(x->local, x->world, x->max_words, x->data_bytes, x->d_step, x->d_error, nq, k, out_flags);
  HDB_LAUNCHED();
  exchange_merge_kernel<<<(unsigned)nq, 256, 0, s>>>(x->local, x->world, x->max_words, x->d_step, nq, k, out_idx, out_score, out_count);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
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
