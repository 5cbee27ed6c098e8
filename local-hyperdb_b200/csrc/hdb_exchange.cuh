// Device-side protocol of the peer-memory candidate exchange (csrc/exchange.cu), shared with the certify kernel
// (csrc/finalize.cu), which pushes its results straight into the peers' slots.
#pragma once
#include "hdb_common.cuh"
#include "hdb_internal.h"

namespace hdb {

constexpr unsigned long long kXTimeoutNs = 10ull * 1000 * 1000 * 1000;     // a peer died: give up instead of hanging the GPU

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

// rank g's slot for THIS rank's message of step `step`
__device__ __forceinline__ unsigned long long* push_slot(const PushTarget& t, int g, unsigned long long step) {
  const int slot = (int)(step % kXSlots);
  return reinterpret_cast<unsigned long long*>(t.peer[g]) + ((int64_t)slot * t.world + t.rank) * t.max_words;
}

// Flow control (one thread): rank g must have merged step - kXSlots before its slot is overwritten.  g writes its
// consumed counter into OUR buffer, so this spins on local memory; in steady state the condition already holds.
__device__ __forceinline__ void push_wait_consumed(const PushTarget& t, int g, unsigned long long step) {
  if (step < (unsigned long long)kXSlots) return;
  const unsigned long long* consumed = reinterpret_cast<const unsigned long long*>(t.local + t.consumed_off) + g;
  const unsigned long long need = step - kXSlots + 1;
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(consumed) < need) {
    if (global_ns() - t0 > kXTimeoutNs) { *t.error = 1; break; }
    __nanosleep(100);
  }
}

// Called by ONE thread of every CTA of a pushing kernel after its peer stores (+ __threadfence_system + barrier):
// the last of `n_ctas` publishes the arrival flag of this step in every rank's buffer and advances the push counter.
__device__ __forceinline__ void push_publish(const PushTarget& t, unsigned long long step, unsigned n_ctas) {
  const unsigned long long old = atomicAdd(t.ctr + kCtrPushDone, 1ull);
  if (old != (unsigned long long)n_ctas - 1) return;
  t.ctr[kCtrPushDone] = 0;
  __threadfence_system();                          // the other CTAs' (fenced) stores are ordered before the flags below
  const int slot = (int)(step % kXSlots);
  for (int g = 0; g < t.world; ++g) {
    unsigned long long* flag = reinterpret_cast<unsigned long long*>(t.peer[g] + t.data_bytes) + (slot * t.world + t.rank);
    st_release_sys(flag, step + 1);
  }
  t.ctr[kCtrPushStep] = step + 1;
}

}  // namespace hdb
