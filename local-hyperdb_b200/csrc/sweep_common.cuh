// Shared pieces of the fused score + select sweeps (csrc/sweep.cu and its per-type translation units): storage traits,
// kernel parameters, the per-warp candidate lists and the per-CTA merge.
#pragma once
#include "hdb_common.cuh"
#include "hdb_internal.h"
#include "../../include/hyperdb_b200.h"

namespace hdb {

constexpr int kRows = 8;              // rows per warp per group

template <typename T> struct Store;
template <> struct Store<__half> { using Acc = float; static constexpr int kPerVec = 8; };
template <> struct Store<float>  { using Acc = float; static constexpr int kPerVec = 4; };
template <> struct Store<double> { using Acc = double; static constexpr int kPerVec = 2; };

struct SweepParams {
  const char* rows;
  int64_t n, d;
  int64_t row_bytes;
  int nvec;                 // 16-byte vectors per row (vector path) or elements per row (scalar path)
  const void* qa;           // query in the accumulate type (global)
  const void* inv_norms;    // accumulate type, or nullptr (not cosine / pearson); pearson: 1/(std*d)
  const void* row_means;    // pearson: np.mean per row (accumulate type), else nullptr
  const double* qaux;       // pearson: {np.std(q), sum_j (q_j - mean)} of this query, else nullptr
  RowFilter f;
  uint64_t* cand;
  // 1: float16 rows AND float16 queries on the dot / cosine / pearson class: the prepared query values are float16 numbers,
  // so the sweep keeps them as float16 in shared memory and multiplies with the mixed-precision FMA (see accum_hh)
  int q_half;
  unsigned long long* tau;  // [NQ] running thresholds of the queries of this pass
  int64_t cand_stride;      // keys between the candidate blocks of consecutive queries (grid * KP)
  int metric;
};

__device__ __forceinline__ float abs_of(float x) { return fabsf(x); }
__device__ __forceinline__ double abs_of(double x) { return fabs(x); }
__device__ __forceinline__ float sqrt_of(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_of(double x) { return sqrt(x); }

// MC: 0 = dot/cosine, 1 = squared L2, 2 = L1
template <int MC, typename Acc>
__device__ __forceinline__ void accum(Acc& a, Acc v, Acc q) {
  if (MC == 0) {
    a = fma(v, q, a);
  } else if (MC == 1) {
    Acc df = v - q;
    a = fma(df, df, a);
  } else {
    a += abs_of(v - q);
  }
}

template <int MC>
__device__ __forceinline__ void accum_vec(float& a, const uint4& raw, const float* q, __half) {
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 v = __half22float2(h[i]);
    accum<MC, float>(a, v.x, q[2 * i]);
    accum<MC, float>(a, v.y, q[2 * i + 1]);
  }
}
template <int MC>
__device__ __forceinline__ void accum_vec(float& a, const uint4& raw, const float* q, float) {
  const float* v = reinterpret_cast<const float*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) accum<MC, float>(a, v[i], q[i]);
}
template <int MC>
__device__ __forceinline__ void accum_vec(double& a, const uint4& raw, const double* q, double) {
  const double* v = reinterpret_cast<const double*>(&raw);
#pragma unroll
  for (int i = 0; i < 2; ++i) accum<MC, double>(a, v[i], q[i]);
}

// float16 x float16 + float32 -> float32 in ONE instruction (PTX fma.rn.f32.f16, SASS FHFMA with .H0/.H1 operand selectors):
// the product of two float16 numbers is exact in float32, so this is bit-for-bit fmaf(float(v), float(q), a) without the two
// conversions -- the instruction count of the float16 dot sweep halves.  8 elements of a stored vector against 8 query
// elements, in element order.
__device__ __forceinline__ void accum_hh(float& a, const uint4& v, const uint4& q) {
  asm("{\n\t.reg .b16 vl, vh, ql, qh;\n\t"
      "mov.b32 {vl, vh}, %1;\n\tmov.b32 {ql, qh}, %5;\n\tfma.rn.f32.f16 %0, vl, ql, %0;\n\tfma.rn.f32.f16 %0, vh, qh, %0;\n\t"
      "mov.b32 {vl, vh}, %2;\n\tmov.b32 {ql, qh}, %6;\n\tfma.rn.f32.f16 %0, vl, ql, %0;\n\tfma.rn.f32.f16 %0, vh, qh, %0;\n\t"
      "mov.b32 {vl, vh}, %3;\n\tmov.b32 {ql, qh}, %7;\n\tfma.rn.f32.f16 %0, vl, ql, %0;\n\tfma.rn.f32.f16 %0, vh, qh, %0;\n\t"
      "mov.b32 {vl, vh}, %4;\n\tmov.b32 {ql, qh}, %8;\n\tfma.rn.f32.f16 %0, vl, ql, %0;\n\tfma.rn.f32.f16 %0, vh, qh, %0;\n\t}"
      : "+f"(a) : "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(q.x), "r"(q.y), "r"(q.z), "r"(q.w));
}

// 8 per-lane partial sums -> every lane holds the full sum of row ((lane>>4)&1)*4 + ((lane>>3)&1)*2 + ((lane>>2)&1)
template <typename Acc>
__device__ __forceinline__ Acc reduce8(Acc (&acc)[kRows], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    Acc send = b4 ? acc[i] : acc[i + 4];
    Acc keep = b4 ? acc[i + 4] : acc[i];
    acc[i] = keep + __shfl_xor_sync(kFull, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    Acc send = b3 ? acc[i] : acc[i + 2];
    Acc keep = b3 ? acc[i + 2] : acc[i];
    acc[i] = keep + __shfl_xor_sync(kFull, send, 8);
  }
  {
    Acc send = b2 ? acc[0] : acc[1];
    Acc keep = b2 ? acc[1] : acc[0];
    acc[0] = keep + __shfl_xor_sync(kFull, send, 4);
  }
  acc[0] += __shfl_xor_sync(kFull, acc[0], 2);
  acc[0] += __shfl_xor_sync(kFull, acc[0], 1);
  return acc[0];
}

// ---------------------------------------------------------------------------------------------
// Per-warp candidate list in shared memory (CAP = 4*KP... see kCap), guarded by thresholds.
// ---------------------------------------------------------------------------------------------
template <int KP> struct ListCfg { static constexpr int kCap = (KP <= 32) ? 128 : 2 * KP; };

template <int KP>
struct WarpList {
  static constexpr int kCap = ListCfg<KP>::kCap;
  uint64_t* buf;       // this warp's kCap slots
  int cnt;
  uint64_t tau;        // keys <= tau cannot be in the global top-KP

  __device__ __forceinline__ void compact(int lane, unsigned long long* s_tau, unsigned long long* g_tau) {
    for (int i = cnt + lane; i < kCap; i += 32) buf[i] = 0;
    __syncwarp();
    bitonic_desc(buf, kCap, lane, 32, [] { __syncwarp(); });
    if (cnt >= KP) {
      cnt = KP;
      uint64_t mine = buf[KP - 1];
      if (mine > tau) {
        tau = mine;
        if (lane == 0) {
          atomicMax(s_tau, (unsigned long long)mine);
          atomicMax(g_tau, (unsigned long long)mine);
        }
      }
    }
  }

  // warp-collective: lanes with `pass` append their key
  __device__ __forceinline__ void push(bool pass, uint64_t key, int lane, unsigned long long* s_tau,
                                       unsigned long long* g_tau) {
    unsigned m = __ballot_sync(kFull, pass);
    if (m == 0) return;
    if (pass) buf[cnt + __popc(m & ((1u << lane) - 1u))] = key;
    cnt += __popc(m);
    __syncwarp();
    // first fill: establish a threshold as soon as KP entries exist; later: only when nearly full
    if (cnt > kCap - 32 || (tau == 0 && cnt >= KP)) compact(lane, s_tau, g_tau);
  }
};

// End of kernel: every warp's list is sorted; the CTA's top-KP of the kSweepWarps*KP head entries is found by rank
// counting (keys are unique): one pass of broadcast shared-memory reads instead of a 50-70 step bitonic sort.
template <int KP>
__device__ __forceinline__ void cta_merge_and_store(uint64_t* s_lists, WarpList<KP>& wl, int lane, int warp,
                                                    unsigned long long* s_tau, unsigned long long* g_tau,
                                                    uint64_t* cand_out) {
  constexpr int kCap = ListCfg<KP>::kCap;
  wl.compact(lane, s_tau, g_tau);                        // sorted descending, at most KP valid entries, zeros after
  for (int i = KP + lane; i < kCap; i += 32) wl.buf[i] = 0;
  for (int i = threadIdx.x; i < KP; i += kSweepThreads) cand_out[i] = 0;
  __syncthreads();
  constexpr int kTotal = kSweepWarps * KP;
  for (int e = threadIdx.x; e < kTotal; e += kSweepThreads) {
    const uint64_t mine = s_lists[(e / KP) * kCap + (e % KP)];
    if (mine == 0) continue;
    int rank = 0;
    for (int w = 0; w < kSweepWarps; ++w) {
      const uint64_t* lst = s_lists + w * kCap;
      // lists are sorted: stop at the first key that is not larger
      for (int j = 0; j < KP; ++j) {
        if (lst[j] > mine) ++rank; else break;
      }
      if (rank >= KP) break;
    }
    if (rank < KP) cand_out[rank] = mine;
  }
}

// Selection key of physical row `row`.  With a row order set the key carries the row's index in the CALLER's numbering
// (ties -> lower index there); that index is only loaded for rows whose score alone could still pass the threshold
// (make_key(score, 0) is the largest key the score can give), i.e. for a handful of rows per warp.  0 never passes.
__device__ __forceinline__ uint64_t ordered_key(const RowFilter& f, float score, uint32_t row, bool alive, uint64_t tau) {
  if (f.ord == nullptr) return make_key(score, row);
  if (!alive || make_key(score, 0u) <= tau) return 0;
  return make_key(score, f.ord[row]);
}

// keep bits of the 32-row window w (rows 32w .. 32w+31): mask word AND kept range AND row count
__device__ __forceinline__ uint32_t window_keep_bits(const RowFilter& f, int64_t w, int64_t n) {
  uint32_t bits = f.mask ? f.mask[w] : 0xffffffffu;
  const int64_t row0 = w * 32;
  const int64_t hi = f.hi < n ? f.hi : n;
  if (row0 < f.lo) { const int64_t s = f.lo - row0; bits = s >= 32 ? 0u : (bits & (0xffffffffu << s)); }
  if (row0 + 32 > hi) { const int64_t keep = hi - row0; bits = keep <= 0 ? 0u : (bits & (0xffffffffu >> (32 - keep))); }
  return bits;
}

}  // namespace hdb
