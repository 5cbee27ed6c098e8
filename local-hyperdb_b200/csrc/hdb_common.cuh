// Shared device/host helpers for the hyperdb_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math.h>

namespace hdb {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// Selection key of the fused pass: [ordered float32 score : 32][~row : 32].  Larger key = better
// rank under (score desc, row asc); keys are unique because rows are.  0 is below every real key.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t order_f32(float s) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(s);
#else
  union { float f; uint32_t u; } c; c.f = s; uint32_t b = c.u;
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float unorder_f32(uint32_t o) {
  uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
  if (score != score) score = -INFINITY;                       // NaN -> -inf, ranking_algorithm.py:174
  return (uint64_t(order_f32(score)) << 32) | uint64_t(~row);
}
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return unorder_f32(uint32_t(k >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t k) { return ~uint32_t(k); }

// 128-bit streaming load: read-only path, do not allocate in L1 (each byte is used once).
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// element -> double / float for any storage dtype (runtime switch; used off the hot path)
__device__ __forceinline__ double load_as_double(const void* base, int dtype, int64_t i) {
  if (dtype == 0) return (double)__half2float(reinterpret_cast<const __half*>(base)[i]);
  if (dtype == 1) return (double)reinterpret_cast<const float*>(base)[i];
  return reinterpret_cast<const double*>(base)[i];
}

__host__ __device__ __forceinline__ int dtype_size(int dtype) { return dtype == 0 ? 2 : (dtype == 1 ? 4 : 8); }
__host__ __device__ __forceinline__ double unit_roundoff(int dtype) {
  return dtype == 0 ? 4.8828125e-4 /*2^-11*/ : (dtype == 1 ? 5.9604644775390625e-8 /*2^-24*/
                                                            : 1.1102230246251565e-16 /*2^-53*/);
}

// bitonic sort (descending) of n = power-of-two u64 keys in shared memory by `nthreads` threads
// that all call this; `sync` must synchronise exactly those threads.
template <typename SyncFn>
__device__ __forceinline__ void bitonic_desc(uint64_t* a, int n, int tid, int nthreads, SyncFn sync) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (n >> 1); t += nthreads) {
        int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int hi = lo | j;
        bool desc = ((lo & k) == 0);
        uint64_t x = a[lo], y = a[hi];
        if ((x < y) == desc) { a[lo] = y; a[hi] = x; }
      }
      sync();
    }
  }
}

}  // namespace hdb
