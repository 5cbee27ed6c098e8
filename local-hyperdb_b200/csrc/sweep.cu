// Fused score + select pass over one row shard (single query per launch): the replacement of
//   metric_func(vectors, q)            hyperdb/ranking_algorithm.py:168 (np.dot / norm / sum over N x D)
//   + recency_bias*exp(ts - max ts)    :179-186
//   argpartition / argsort             :199-200
// in ONE streaming read of the matrix.  HBM-bound: every row is read exactly once with 128-bit
// coalesced ld.global.nc loads (8 rows in flight per warp), the query lives in shared memory in the
// accumulate type, per-row scores are reduced with warp shuffles, and the epilogue (cosine norm,
// time decay, row mask/range, NaN -> -inf) feeds a per-warp candidate list guarded by a grid-wide
// threshold.  The N-length score vector is never written.  Output: per-CTA top-KP selection keys;
// csrc/finalize.cu merges them, re-scores the KP candidates in the reference's exact arithmetic and
// certifies the top-k (DESIGN.md "select then certify").
#include <cstdlib>

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
  unsigned long long* tau;
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

// keep bits of the 32-row window w (rows 32w .. 32w+31): mask word AND kept range AND row count
__device__ __forceinline__ uint32_t window_keep_bits(const RowFilter& f, int64_t w, int64_t n) {
  uint32_t bits = f.mask ? f.mask[w] : 0xffffffffu;
  const int64_t row0 = w * 32;
  const int64_t hi = f.hi < n ? f.hi : n;
  if (row0 < f.lo) { const int64_t s = f.lo - row0; bits = s >= 32 ? 0u : (bits & (0xffffffffu << s)); }
  if (row0 + 32 > hi) { const int64_t keep = hi - row0; bits = keep <= 0 ? 0u : (bits & (0xffffffffu >> (32 - keep))); }
  return bits;
}

// ---------------------------------------------------------------------------------------------
// float sweeps
// ---------------------------------------------------------------------------------------------
template <typename T, int MC, int KP, bool VEC>
__global__ void __launch_bounds__(kSweepThreads, 2) sweep_kernel(SweepParams p) {
  using Acc = typename Store<T>::Acc;
  constexpr int kPerVec = Store<T>::kPerVec;
  constexpr int kCap = ListCfg<KP>::kCap;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* s_lists = reinterpret_cast<uint64_t*>(smem_raw);                       // [warps][kCap]
  unsigned long long* s_tau = reinterpret_cast<unsigned long long*>(s_lists + kSweepWarps * kCap);
  Acc* s_q = reinterpret_cast<Acc*>(s_tau + 2);                                    // [d], 16-byte aligned

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) *s_tau = 0;
  for (int64_t j = threadIdx.x; j < p.d; j += kSweepThreads) s_q[j] = reinterpret_cast<const Acc*>(p.qa)[j];
  __syncthreads();

  WarpList<KP> wl;
  wl.buf = s_lists + warp * kCap;
  wl.cnt = 0;
  wl.tau = 0;

  // A warp walks windows of 32 consecutive rows (one mask word) and, inside a window, batches of up to 8
  // KEPT rows: dropped rows are never loaded and 8 independent 16-byte loads per lane stay in flight
  // whatever the mask density is.
  const int64_t nwin = (p.n + 31) / 32;
  const int64_t wstride = (int64_t)gridDim.x * kSweepWarps;
  const int my_row = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  const bool rep = (lane & 3) == 0;
  const Acc* inv = reinterpret_cast<const Acc*>(p.inv_norms);
  // pearson (MC == 0 only): sum_j (v_j - mean_v) b_j = v.b - mean_v * sum(b), then / (std_v * d) / std_q
  const Acc* pmeans = reinterpret_cast<const Acc*>(p.row_means);
  Acc q_sumb = Acc(0), q_scale = Acc(1);
  if (MC == 0 && p.qaux) {
    q_sumb = (Acc)p.qaux[1];
    q_scale = (p.qaux[0] == 0.0) ? (Acc)__longlong_as_double(0x7ff8000000000000ll) : (Acc)(1.0 / p.qaux[0]);
  }
  int since_refresh = 0;
  const int64_t g0 = (int64_t)blockIdx.x * kSweepWarps + warp;
  uint32_t next_bits = (g0 < nwin) ? window_keep_bits(p.f, g0, p.n) : 0u;

  for (int64_t g = g0; g < nwin; g += wstride) {
    uint32_t bits = next_bits;
    next_bits = (g + wstride < nwin) ? window_keep_bits(p.f, g + wstride, p.n) : 0u;    // prefetch the next mask word
    if (bits == 0) continue;
    // refresh the threshold from the CTA (cheap) and, now and then, from the grid
    {
      unsigned long long t = *reinterpret_cast<volatile unsigned long long*>(s_tau);
      if (++since_refresh >= 4) {
        since_refresh = 0;
        unsigned long long gt = *reinterpret_cast<volatile unsigned long long*>(p.tau);
        if (gt > t) { t = gt; if (lane == 0) atomicMax(s_tau, gt); }
      }
      if (t > wl.tau) wl.tau = t;
    }
    const int64_t row0 = g * 32;
    const char* base = p.rows + row0 * p.row_bytes;
    while (bits) {
      // the next (up to) 8 kept rows of the window: warp-uniform offsets
      uint32_t ro[kRows];                 // byte offsets inside the window (32 rows < 4 GB)
      unsigned keep = 0;
      int my_off = 0;
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const int pos = __ffs(bits) - 1;
        const bool ok = bits != 0;
        ro[r] = (uint32_t)(ok ? pos : 0) * (uint32_t)p.row_bytes;
        keep |= (ok ? 1u : 0u) << r;
        if (r == my_row) my_off = pos;
        bits &= bits - 1;
      }
      // per-row side inputs, issued before the streaming loop so their latency is hidden
      const bool mine_kept = (keep >> my_row) & 1u;
      const int64_t mrow = row0 + my_off;
      Acc my_inv = Acc(1), my_mean = Acc(0);
      double my_decay = 0.0;
      if (rep && mine_kept) {
        if (inv) my_inv = inv[mrow];
        if (MC == 0 && pmeans) my_mean = pmeans[mrow];
        if (p.f.decay) my_decay = p.f.decay[mrow];
      }

      Acc acc[kRows];
#pragma unroll
      for (int r = 0; r < kRows; ++r) acc[r] = Acc(0);

      if (VEC) {
#pragma unroll 2
        for (int c = lane; c < p.nvec; c += 32) {
          uint4 raw[kRows];
#pragma unroll
          for (int r = 0; r < kRows; ++r) {
            if ((keep >> r) & 1u) raw[r] = ld_stream16(base + (ro[r] + (uint32_t)c * 16u));
            else raw[r] = make_uint4(0, 0, 0, 0);
          }
          Acc q[kPerVec];
#pragma unroll
          for (int i = 0; i < kPerVec; ++i) q[i] = s_q[c * kPerVec + i];
#pragma unroll
          for (int r = 0; r < kRows; ++r) accum_vec<MC>(acc[r], raw[r], q, T());
        }
      } else {
        for (int c = lane; c < p.nvec; c += 32) {
          const Acc q = s_q[c];
#pragma unroll
          for (int r = 0; r < kRows; ++r) {
            if ((keep >> r) & 1u) accum<MC, Acc>(acc[r], (Acc) reinterpret_cast<const T*>(base + ro[r])[c], q);
          }
        }
      }

      Acc total = reduce8(acc, lane);
      // epilogue: similarity, decay, key
      float score;
      if (MC == 0) {
        if (pmeans) total = (total - my_mean * q_sumb) * q_scale;
        total = total * my_inv;
        if (p.f.decay) score = (float)((double)total + p.f.bias * my_decay);
        else score = (float)total;
      } else {
        Acc dist = (MC == 1) ? sqrt_of(total) : total;
        Acc sim = Acc(1) / (Acc(1) + dist);
        if (p.f.decay) score = (float)((double)sim + p.f.bias * my_decay);
        else score = (float)sim;
      }
      const uint64_t key = make_key(score, (uint32_t)mrow);
      wl.push(rep && mine_kept && key > wl.tau, key, lane, s_tau, p.tau);
    }
  }

  cta_merge_and_store<KP>(s_lists, wl, lane, warp, s_tau, p.tau, p.cand + (int64_t)blockIdx.x * KP);
}

// ---------------------------------------------------------------------------------------------
// Hamming sweep on the bit-packed matrix: words (multiple of 4) u32 per row = nvec 16-byte vectors.
// LPR lanes share a row (LPR = smallest power of two >= nvec, capped at 32); score = D - popcount(xor).
// ---------------------------------------------------------------------------------------------
struct HammingParams {
  const uint32_t* bits;
  const uint32_t* qbits;
  int64_t n, d;
  int nvec, lpr;
  int jaccard;               // 0: score = d - popcount(xor); 1: score = popcount(and) / popcount(or)
  RowFilter f;
  uint64_t* cand;
  unsigned long long* tau;
};

template <int KP>
__global__ void __launch_bounds__(kSweepThreads, 2) sweep_hamming_kernel(HammingParams p) {
  constexpr int kCap = ListCfg<KP>::kCap;
  constexpr int kPasses = 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* s_lists = reinterpret_cast<uint64_t*>(smem_raw);
  unsigned long long* s_tau = reinterpret_cast<unsigned long long*>(s_lists + kSweepWarps * kCap);
  uint4* s_q = reinterpret_cast<uint4*>(s_tau + 2);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) *s_tau = 0;
  for (int j = threadIdx.x; j < p.nvec; j += kSweepThreads) s_q[j] = reinterpret_cast<const uint4*>(p.qbits)[j];
  __syncthreads();

  WarpList<KP> wl;
  wl.buf = s_lists + warp * kCap;
  wl.cnt = 0;
  wl.tau = 0;

  const int lpr = p.lpr, rpp = 32 / lpr;                // rows per pass
  const int sub = lane % lpr, slot = lane / lpr;
  const int64_t rows_per_group = (int64_t)rpp * kPasses;
  const int64_t ngroups = (p.n + rows_per_group - 1) / rows_per_group;
  const int64_t wstride = (int64_t)gridDim.x * kSweepWarps;
  const int64_t hi = p.f.hi < p.n ? p.f.hi : p.n;
  int since_refresh = 0;
  const bool single_vec = p.nvec <= lpr;
  const uint4 my_q = (sub < p.nvec) ? s_q[sub] : make_uint4(0, 0, 0, 0);

  for (int64_t g = (int64_t)blockIdx.x * kSweepWarps + warp; g < ngroups; g += wstride) {
    {
      unsigned long long t = *reinterpret_cast<volatile unsigned long long*>(s_tau);
      if (++since_refresh >= 16) {
        since_refresh = 0;
        unsigned long long gt = *reinterpret_cast<volatile unsigned long long*>(p.tau);
        if (gt > t) { t = gt; if (lane == 0) atomicMax(s_tau, gt); }
      }
      if (t > wl.tau) wl.tau = t;
    }
    const int64_t row0 = g * rows_per_group;
    int diff[kPasses];
    bool kept[kPasses];
#pragma unroll
    for (int r = 0; r < kPasses; ++r) {
      const int64_t row = row0 + (int64_t)r * rpp + slot;
      bool k = row >= p.f.lo && row < hi;
      if (k && p.f.mask) k = (p.f.mask[row >> 5] >> (row & 31)) & 1u;
      kept[r] = k;
    }
    if (single_vec) {
      // one 16-byte vector per lane per row: issue all kPasses loads before the first popcount
      uint4 v[kPasses];
#pragma unroll
      for (int r = 0; r < kPasses; ++r) {
        const int64_t row = row0 + (int64_t)r * rpp + slot;
        v[r] = (kept[r] && sub < p.nvec) ? ld_stream16(reinterpret_cast<const uint4*>(p.bits) + row * p.nvec + sub)
                                         : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int r = 0; r < kPasses; ++r) {
        const uint4 q = (kept[r] && sub < p.nvec) ? my_q : make_uint4(0, 0, 0, 0);
        if (p.jaccard)
          diff[r] = (__popc(v[r].x & q.x) + __popc(v[r].y & q.y) + __popc(v[r].z & q.z) + __popc(v[r].w & q.w)) +
                    ((__popc(v[r].x | q.x) + __popc(v[r].y | q.y) + __popc(v[r].z | q.z) + __popc(v[r].w | q.w)) << 16);
        else
          diff[r] = __popc(v[r].x ^ q.x) + __popc(v[r].y ^ q.y) + __popc(v[r].z ^ q.z) + __popc(v[r].w ^ q.w);
      }
    } else {
#pragma unroll
      for (int r = 0; r < kPasses; ++r) {
        const int64_t row = row0 + (int64_t)r * rpp + slot;
        int dsum = 0;
        if (kept[r]) {
          const uint4* rowp = reinterpret_cast<const uint4*>(p.bits) + row * p.nvec;
          for (int c = sub; c < p.nvec; c += lpr) {
            uint4 v = ld_stream16(rowp + c);
            uint4 q = s_q[c];
            if (p.jaccard)      // low 16 bits: popcount(and), high 16 bits: popcount(or)  (d <= 32768 per the smem limit)
              dsum += (__popc(v.x & q.x) + __popc(v.y & q.y) + __popc(v.z & q.z) + __popc(v.w & q.w)) +
                      ((__popc(v.x | q.x) + __popc(v.y | q.y) + __popc(v.z | q.z) + __popc(v.w | q.w)) << 16);
            else
              dsum += __popc(v.x ^ q.x) + __popc(v.y ^ q.y) + __popc(v.z ^ q.z) + __popc(v.w ^ q.w);
          }
        }
        diff[r] = dsum;
      }
    }
#pragma unroll
    for (int r = 0; r < kPasses; ++r) {
      int dsum = diff[r];
      for (int o = lpr >> 1; o; o >>= 1) dsum += __shfl_xor_sync(kFull, dsum, o);
      const int64_t row = row0 + (int64_t)r * rpp + slot;
      double exact = p.jaccard ? (double)(dsum & 0xffff) / (double)(dsum >> 16) : (double)((int)p.d - dsum);
      if (p.f.decay && kept[r] && sub == 0) exact += p.f.bias * p.f.decay[row];
      const float score = (float)exact;
      const uint64_t key = make_key(score, (uint32_t)row);
      wl.push(sub == 0 && kept[r] && key > wl.tau, key, lane, s_tau, p.tau);
    }
  }
  cta_merge_and_store<KP>(s_lists, wl, lane, warp, s_tau, p.tau, p.cand + (int64_t)blockIdx.x * KP);
}

// Fast form for rows of at most 32 vectors (d <= 4096): LPR (compile time) lanes share a row, a warp takes one
// 32-row window (one mask word) per iteration, every lane issues min(LPR,8) independent 16-byte loads, and the
// per-lane popcounts are reduced with a TRANSPOSING butterfly so that every lane ends up owning the total
// of one distinct row: one ballot/push per round instead of one per row group.
template <int PPR, int LPR>
__device__ __forceinline__ int transpose_reduce(int (&v)[PPR], int sub) {
#pragma unroll
  for (int o = LPR / 2; o >= PPR && o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < PPR; ++i) v[i] += __shfl_xor_sync(kFull, v[i], o);
  }
  int n = PPR;
#pragma unroll
  for (int o = PPR / 2; o >= 1; o >>= 1) {
    const bool up = sub & o;
    const int half = n / 2;
#pragma unroll
    for (int i = 0; i < PPR / 2; ++i) {
      if (i < half) {
        const int send = up ? v[i] : v[i + half];
        const int keep = up ? v[i + half] : v[i];
        v[i] = keep + __shfl_xor_sync(kFull, send, o);
      }
    }
    n = half;
  }
  return v[0];                       // total of pass (sub & (PPR-1))
}

template <int KP, int LPR, bool JAC>
__global__ void __launch_bounds__(kSweepThreads, 2) sweep_hamming_lpr_kernel(HammingParams p) {
  constexpr int kCap = ListCfg<KP>::kCap;
  constexpr int PPR = LPR < 8 ? LPR : 8;       // passes per round = loads in flight per lane
  constexpr int ROUNDS = LPR / PPR;
  constexpr int RPP = 32 / LPR;                // rows per pass
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* s_lists = reinterpret_cast<uint64_t*>(smem_raw);
  unsigned long long* s_tau = reinterpret_cast<unsigned long long*>(s_lists + kSweepWarps * kCap);
  uint4* s_q = reinterpret_cast<uint4*>(s_tau + 2);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) *s_tau = 0;
  for (int j = threadIdx.x; j < p.nvec; j += kSweepThreads) s_q[j] = reinterpret_cast<const uint4*>(p.qbits)[j];
  __syncthreads();

  WarpList<KP> wl;
  wl.buf = s_lists + warp * kCap;
  wl.cnt = 0;
  wl.tau = 0;

  const int sub = lane % LPR, slot = lane / LPR;
  const bool lane_has = sub < p.nvec;
  const uint4 my_q = lane_has ? s_q[sub] : make_uint4(0, 0, 0, 0);
  const bool rep = sub < PPR;
  const int my_pass = sub & (PPR - 1);
  const int64_t nwin = (p.n + 31) / 32;
  const int64_t wstride = (int64_t)gridDim.x * kSweepWarps;
  const uint4* vbits = reinterpret_cast<const uint4*>(p.bits);
  int since_refresh = 0;
  const int64_t g0 = (int64_t)blockIdx.x * kSweepWarps + warp;
  uint32_t next_bits = (g0 < nwin) ? window_keep_bits(p.f, g0, p.n) : 0u;

  // the PPR loads of round `round` of window `win` (dropped rows / idle lanes: the query itself, i.e. xor = 0)
  auto load_round = [&](uint4 (&v)[PPR], int64_t win, uint32_t wbits, int round) {
#pragma unroll
    for (int r = 0; r < PPR; ++r) {
      const int loc = (round * PPR + r) * RPP + slot;
      const bool k = ((wbits >> loc) & 1u) && lane_has;
      v[r] = k ? ld_stream16(vbits + (win * 32 + loc) * p.nvec + sub) : my_q;
    }
  };
  // Software pipeline: round 0 of the NEXT window is in flight (registers) while the current window is counted,
  // reduced and pushed, so every warp keeps 32 rows of loads outstanding at all times.
  uint4 cur[PPR];
  if (g0 < nwin) load_round(cur, g0, next_bits, 0);

  for (int64_t g = g0; g < nwin; g += wstride) {
    const uint32_t bits = next_bits;
    next_bits = (g + wstride < nwin) ? window_keep_bits(p.f, g + wstride, p.n) : 0u;
    uint4 nxt[PPR];
    if (g + wstride < nwin) load_round(nxt, g + wstride, next_bits, 0);
    if (bits != 0) {
      {
        unsigned long long t = *reinterpret_cast<volatile unsigned long long*>(s_tau);
        if (++since_refresh >= 8) {
          since_refresh = 0;
          unsigned long long gt = *reinterpret_cast<volatile unsigned long long*>(p.tau);
          if (gt > t) { t = gt; if (lane == 0) atomicMax(s_tau, gt); }
        }
        if (t > wl.tau) wl.tau = t;
      }
      const int64_t row0 = g * 32;
#pragma unroll
      for (int round = 0; round < ROUNDS; ++round) {
        const int my_loc = (round * PPR + my_pass) * RPP + slot;        // the row this lane will own after the reduce
        const bool my_kept = (bits >> my_loc) & 1u;
        double my_decay = 0.0;
        if (p.f.decay && rep && my_kept) my_decay = p.f.decay[row0 + my_loc];
        uint4 v[PPR];
        if (round == 0) {
#pragma unroll
          for (int r = 0; r < PPR; ++r) v[r] = cur[r];
        } else {
          load_round(v, g, bits, round);
        }
        int cnt[PPR];
#pragma unroll
        for (int r = 0; r < PPR; ++r) {
          if (JAC)          // both popcounts packed in one int (16 bits each: d <= 4096 here), reduced together
            cnt[r] = (__popc(v[r].x & my_q.x) + __popc(v[r].y & my_q.y) + __popc(v[r].z & my_q.z) + __popc(v[r].w & my_q.w)) +
                     ((__popc(v[r].x | my_q.x) + __popc(v[r].y | my_q.y) + __popc(v[r].z | my_q.z) + __popc(v[r].w | my_q.w)) << 16);
          else
            cnt[r] = __popc(v[r].x ^ my_q.x) + __popc(v[r].y ^ my_q.y) + __popc(v[r].z ^ my_q.z) + __popc(v[r].w ^ my_q.w);
        }
        const int diff = transpose_reduce<PPR, LPR>(cnt, sub);
        double exact = JAC ? (double)(diff & 0xffff) / (double)(diff >> 16) : (double)((int)p.d - diff);
        if (p.f.decay) exact += p.f.bias * my_decay;
        const float score = (float)exact;
        const uint64_t key = make_key(score, (uint32_t)(row0 + my_loc));
        wl.push(rep && my_kept && key > wl.tau, key, lane, s_tau, p.tau);
      }
    }
#pragma unroll
    for (int r = 0; r < PPR; ++r) cur[r] = nxt[r];
  }
  cta_merge_and_store<KP>(s_lists, wl, lane, warp, s_tau, p.tau, p.cand + (int64_t)blockIdx.x * KP);
}

// ---------------------------------------------------------------------------------------------
// Staged form for unmasked shards with rows of at most 8 vectors (d <= 1024): the window of 32 consecutive rows is ONE
// contiguous block of HBM (32 * NVEC * 16 bytes); every lane copies NVEC coalesced 16-byte pieces of it straight into
// shared memory with cp.async (no registers, no L1), double-buffered per warp, and then OWNS one row: it reads its row
// back with NVEC conflict-free 128-bit shared loads (row pitch = an odd number of 16-byte units) and popcounts it
// against the query held in registers.  No shuffles, no transposing reduce: ~3x fewer instructions per row than the
// cooperative form above, which was issue-latency bound (0.49 IPC per scheduler, 4 warps) rather than HBM bound.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <int NVEC> struct StagedCfg {
  static constexpr int kPitch = NVEC | 1;                 // 16-byte units per staged row
  static constexpr int kStageU4 = 32 * kPitch;            // one window
  static constexpr size_t kBytes = (size_t)kSweepWarps * 2 * kStageU4 * 16;
};

template <int KP, int NVEC, bool JAC>
__global__ void __launch_bounds__(kSweepThreads, 2) sweep_hamming_staged_kernel(HammingParams p) {
  constexpr int kCap = ListCfg<KP>::kCap;
  constexpr int PITCH = StagedCfg<NVEC>::kPitch;
  constexpr int STAGE = StagedCfg<NVEC>::kStageU4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* s_lists = reinterpret_cast<uint64_t*>(smem_raw);
  unsigned long long* s_tau = reinterpret_cast<unsigned long long*>(s_lists + kSweepWarps * kCap);
  uint4* s_stage = reinterpret_cast<uint4*>(s_tau + 2);                             // [warps][2][STAGE]

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) *s_tau = 0;
  uint4 q[NVEC];
#pragma unroll
  for (int c = 0; c < NVEC; ++c) q[c] = reinterpret_cast<const uint4*>(p.qbits)[c];
  __syncthreads();

  WarpList<KP> wl;
  wl.buf = s_lists + warp * kCap;
  wl.cnt = 0;
  wl.tau = 0;

  uint4* my_stage = s_stage + (size_t)warp * 2 * STAGE;
  // piece i = j*32 + lane of a window is piece (i % NVEC) of row (i / NVEC): coalesced in HBM, scattered into the padded rows
  int dst_off[NVEC];
#pragma unroll
  for (int j = 0; j < NVEC; ++j) {
    const int i = j * 32 + lane;
    dst_off[j] = (i / NVEC) * PITCH + (i % NVEC);
  }
  const int64_t nwin = (p.n + 31) / 32;
  const int64_t wstride = (int64_t)gridDim.x * kSweepWarps;
  const uint4* vbits = reinterpret_cast<const uint4*>(p.bits);
  const int64_t total_u4 = p.n * NVEC;

  auto issue = [&](int buf, int64_t win) {
    const int64_t base = win * (32 * NVEC);
    uint4* dst = my_stage + buf * STAGE;
#pragma unroll
    for (int j = 0; j < NVEC; ++j) {
      const int64_t src = base + j * 32 + lane;
      if (src < total_u4) cp_async16(dst + dst_off[j], vbits + src);
    }
  };

  int since_refresh = 0;
  const int64_t g0 = (int64_t)blockIdx.x * kSweepWarps + warp;
  uint32_t bits = (g0 < nwin) ? window_keep_bits(p.f, g0, p.n) : 0u;
  if (bits) issue(0, g0);
  cp_async_commit();
  int buf = 0;

  for (int64_t g = g0; g < nwin; g += wstride) {
    const int64_t gn = g + wstride;
    const uint32_t next_bits = (gn < nwin) ? window_keep_bits(p.f, gn, p.n) : 0u;
    if (next_bits) issue(buf ^ 1, gn);                    // overlaps everything below
    cp_async_commit();
    cp_async_wait<1>();                                   // this lane's pieces of window g have landed ...
    __syncwarp();                                         // ... and so have everybody else's
    if (bits) {
      {
        unsigned long long t = *reinterpret_cast<volatile unsigned long long*>(s_tau);
        if (++since_refresh >= 8) {
          since_refresh = 0;
          unsigned long long gt = *reinterpret_cast<volatile unsigned long long*>(p.tau);
          if (gt > t) { t = gt; if (lane == 0) atomicMax(s_tau, gt); }
        }
        if (t > wl.tau) wl.tau = t;
      }
      const int64_t row = g * 32 + lane;
      const bool kept = (bits >> lane) & 1u;
      double my_decay = 0.0;
      if (p.f.decay && kept) my_decay = p.f.decay[row];
      const uint4* rowp = my_stage + buf * STAGE + lane * PITCH;
      int a = 0, b = 0;
#pragma unroll
      for (int c = 0; c < NVEC; ++c) {
        const uint4 v = rowp[c];
        if (JAC) {
          a += __popc(v.x & q[c].x) + __popc(v.y & q[c].y) + __popc(v.z & q[c].z) + __popc(v.w & q[c].w);
          b += __popc(v.x | q[c].x) + __popc(v.y | q[c].y) + __popc(v.z | q[c].z) + __popc(v.w | q[c].w);
        } else {
          a += __popc(v.x ^ q[c].x) + __popc(v.y ^ q[c].y) + __popc(v.z ^ q[c].z) + __popc(v.w ^ q[c].w);
        }
      }
      double exact = JAC ? (double)a / (double)b : (double)((int)p.d - a);
      if (p.f.decay) exact += p.f.bias * my_decay;
      const uint64_t key = make_key((float)exact, (uint32_t)row);
      wl.push(kept && key > wl.tau, key, lane, s_tau, p.tau);
    }
    __syncwarp();                                         // every lane is done with `buf` before the next issue refills it
    buf ^= 1;
    bits = next_bits;
  }
  cp_async_wait<0>();
  cta_merge_and_store<KP>(s_lists, wl, lane, warp, s_tau, p.tau, p.cand + (int64_t)blockIdx.x * KP);
}

template <int KP, int NVEC, bool JAC>
static int launch_hamming_staged3(const HammingParams& hp, int grid, cudaStream_t s) {
  auto kern = sweep_hamming_staged_kernel<KP, NVEC, JAC>;
  const size_t smem = (size_t)kSweepWarps * ListCfg<KP>::kCap * 8 + 16 + StagedCfg<NVEC>::kBytes;
  if (smem > 48 * 1024) HDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kSweepThreads, smem, s>>>(hp);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}
template <int KP, int NVEC>
static int launch_hamming_staged2(const HammingParams& hp, int grid, cudaStream_t s) {
  return hp.jaccard ? launch_hamming_staged3<KP, NVEC, true>(hp, grid, s) : launch_hamming_staged3<KP, NVEC, false>(hp, grid, s);
}
template <int KP>
static int launch_hamming_staged(const HammingParams& hp, int grid, cudaStream_t s) {
  switch (hp.nvec) {
    case 1: return launch_hamming_staged2<KP, 1>(hp, grid, s);
    case 2: return launch_hamming_staged2<KP, 2>(hp, grid, s);
    case 3: return launch_hamming_staged2<KP, 3>(hp, grid, s);
    case 4: return launch_hamming_staged2<KP, 4>(hp, grid, s);
    case 5: return launch_hamming_staged2<KP, 5>(hp, grid, s);
    case 6: return launch_hamming_staged2<KP, 6>(hp, grid, s);
    case 7: return launch_hamming_staged2<KP, 7>(hp, grid, s);
    default: return launch_hamming_staged2<KP, 8>(hp, grid, s);
  }
}

template <int KP, int LPR, bool JAC>
static int launch_hamming_lpr2(const HammingParams& hp, int grid, size_t smem, cudaStream_t s) {
  auto kern = sweep_hamming_lpr_kernel<KP, LPR, JAC>;
  if (smem > 48 * 1024) HDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kSweepThreads, smem, s>>>(hp);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}
template <int KP, int LPR>
static int launch_hamming_lpr(const HammingParams& hp, int grid, size_t smem, cudaStream_t s) {
  return hp.jaccard ? launch_hamming_lpr2<KP, LPR, true>(hp, grid, smem, s) : launch_hamming_lpr2<KP, LPR, false>(hp, grid, smem, s);
}
template <int KP>
static int launch_hamming_kp(const HammingParams& hp, int grid, size_t smem, cudaStream_t s) {
  switch (hp.lpr) {
    case 1: return launch_hamming_lpr<KP, 1>(hp, grid, smem, s);
    case 2: return launch_hamming_lpr<KP, 2>(hp, grid, smem, s);
    case 4: return launch_hamming_lpr<KP, 4>(hp, grid, smem, s);
    case 8: return launch_hamming_lpr<KP, 8>(hp, grid, smem, s);
    case 16: return launch_hamming_lpr<KP, 16>(hp, grid, smem, s);
    default: return launch_hamming_lpr<KP, 32>(hp, grid, smem, s);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// A/B testing: HDB_HAMMING_COOPERATIVE=1 keeps the register-cooperative hamming kernels for every shape
static bool force_cooperative_hamming() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HDB_HAMMING_COOPERATIVE"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
static size_t list_smem(int kp) { return (size_t)kSweepWarps * (kp <= 32 ? 128 : 2 * kp) * 8 + 16; }

int sweep_grid_size(int device) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return sms * 2;                       // __launch_bounds__(256, 2): two CTAs resident per SM
}

template <typename T, int MC, int KP, bool VEC>
static int launch_one(const SweepParams& p, int grid, size_t smem, cudaStream_t s) {
  auto kern = sweep_kernel<T, MC, KP, VEC>;
  if (smem > 48 * 1024) HDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kSweepThreads, smem, s>>>(p);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}
template <typename T, int MC, int KP>
static int launch_vec(const SweepParams& p, bool vec, int grid, size_t smem, cudaStream_t s) {
  return vec ? launch_one<T, MC, KP, true>(p, grid, smem, s) : launch_one<T, MC, KP, false>(p, grid, smem, s);
}
template <typename T, int MC>
static int launch_kp(const SweepParams& p, bool vec, int kp, int grid, size_t smem, cudaStream_t s) {
  return kp <= 32 ? launch_vec<T, MC, 32>(p, vec, grid, smem, s) : launch_vec<T, MC, 128>(p, vec, grid, smem, s);
}
template <typename T>
static int launch_mc(const SweepParams& p, bool vec, int mc, int kp, int grid, size_t smem, cudaStream_t s) {
  if (mc == 0) return launch_kp<T, 0>(p, vec, kp, grid, smem, s);
  if (mc == 1) return launch_kp<T, 1>(p, vec, kp, grid, smem, s);
  return launch_kp<T, 2>(p, vec, kp, grid, smem, s);
}

int launch_sweep(const MatrixView& m, int metric, const void* qa, const uint32_t* qbits, const double* qaux, const RowFilter& f, int kp,
                 const SweepOut& out, cudaStream_t s) {
  if (m.n >= (int64_t(1) << 32)) return fail("sweep: more than 2^32 rows per shard");
  if (kp != 32 && kp != 128) return fail("sweep: unsupported candidate class");
  if (metric == HDB_HAMMING || metric == HDB_JACCARD) {
    HammingParams hp;
    hp.jaccard = metric == HDB_JACCARD;
    hp.bits = m.bits; hp.qbits = qbits; hp.n = m.n; hp.d = m.d;
    hp.nvec = m.words / 4;
    int lpr = 1;
    while (lpr < hp.nvec && lpr < 32) lpr <<= 1;
    hp.lpr = lpr;
    hp.f = f; hp.cand = out.cand; hp.tau = out.tau;
    size_t smem = list_smem(kp) + (size_t)hp.nvec * 16;
    if (smem > 200 * 1024) return fail("sweep: dimension too large for the fused hamming pass");
    if (hp.nvec <= 8 && !f.mask && !force_cooperative_hamming())       // contiguous windows: cp.async-staged, lane-per-row form
      return kp <= 32 ? launch_hamming_staged<32>(hp, out.grid, s) : launch_hamming_staged<128>(hp, out.grid, s);
    if (hp.nvec <= 32) return kp <= 32 ? launch_hamming_kp<32>(hp, out.grid, smem, s) : launch_hamming_kp<128>(hp, out.grid, smem, s);
    if (kp <= 32) {
      if (smem > 48 * 1024) HDB_CUDA(cudaFuncSetAttribute(sweep_hamming_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      sweep_hamming_kernel<32><<<out.grid, kSweepThreads, smem, s>>>(hp);
    } else {
      if (smem > 48 * 1024) HDB_CUDA(cudaFuncSetAttribute(sweep_hamming_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      sweep_hamming_kernel<128><<<out.grid, kSweepThreads, smem, s>>>(hp);
    }
    HDB_LAUNCHED();
    HDB_CUDA(cudaGetLastError());
    return 0;
  }
  SweepParams p;
  p.rows = reinterpret_cast<const char*>(m.rows);
  p.n = m.n; p.d = m.d;
  p.row_bytes = m.d * dtype_size(m.dtype);
  const bool vec = (p.row_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(m.rows) & 15) == 0);
  p.nvec = vec ? (int)(p.row_bytes / 16) : (int)m.d;
  p.qa = qa;
  p.inv_norms = (metric == HDB_COSINE) ? m.inv_norms : (metric == HDB_PEARSON ? m.pscale : nullptr);
  p.row_means = (metric == HDB_PEARSON) ? m.pmean : nullptr;
  p.qaux = (metric == HDB_PEARSON) ? qaux : nullptr;
  if (metric == HDB_PEARSON && (!m.pscale || !m.pmean || !qaux)) return fail("sweep: pearson columns missing");
  p.f = f; p.cand = out.cand; p.tau = out.tau; p.metric = metric;
  const int mc = (metric == HDB_DOT || metric == HDB_COSINE || metric == HDB_PEARSON) ? 0 : (metric == HDB_EUCLIDEAN ? 1 : 2);
  size_t smem = list_smem(kp) + (size_t)m.d * (m.dtype == 2 ? 8 : 4);
  if (smem > 200 * 1024) return fail("sweep: dimension too large for the fused pass");
  if (m.dtype == 0) return launch_mc<__half>(p, vec, mc, kp, out.grid, smem, s);
  if (m.dtype == 1) return launch_mc<float>(p, vec, mc, kp, out.grid, smem, s);
  return launch_mc<double>(p, vec, mc, kp, out.grid, smem, s);
}

}  // namespace hdb
