// The streaming sweep over float16 / float32 / float64 rows: NQ queries per pass.  Included by csrc/sweep_inst.cu, which
// the Makefile compiles once per (storage type, metric class) so that the instantiations build in parallel.
#pragma once
#include <type_traits>
#include "sweep_common.cuh"

namespace hdb {

// ---------------------------------------------------------------------------------------------
// float sweeps: NQ queries per pass (one read of the matrix, NQ score streams)
// ---------------------------------------------------------------------------------------------
// A warp keeps R rows in flight and NQ queries: R*NQ accumulators per lane (8x1, 8x2, 4x4, 4x8).  After the column
// loop a TRANSPOSING butterfly leaves every lane with the full sum of ONE (query, row) pair -- value index
// v = lane >> log2(32 / (R*NQ)), query = v / R, row = v % R -- so the epilogue, the key and the threshold test of all
// pairs run in parallel and each query feeds its own candidate list.
template <int NQ> struct MqCfg {
  static constexpr int kR = (NQ <= 2) ? 8 : 4;
  static constexpr int kV = kR * NQ;                       // 8, 16, 16, 32
  static constexpr int kRepShift = (kV == 8) ? 2 : (kV == 16 ? 1 : 0);
  static constexpr int kLanesPerQuery = 32 / NQ;
};

template <int V, typename Acc>
__device__ __forceinline__ Acc reduce_transpose(Acc (&a)[V], int lane) {
  int n = V;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    if (n > 1) {
      const int half = n >> 1;
      const bool up = lane & o;
#pragma unroll
      for (int i = 0; i < V / 2; ++i) {
        if (i < half) {
          const Acc send = up ? a[i] : a[i + half];
          const Acc keep = up ? a[i + half] : a[i];
          a[i] = keep + __shfl_xor_sync(kFull, send, o);
        }
      }
      n = half;
    } else {
      a[0] += __shfl_xor_sync(kFull, a[0], o);
    }
  }
  return a[0];
}

// NQ candidate lists per warp (list j at base + j * kCap); a lane belongs to the query it owns after the reduce.
template <int KP, int NQ>
struct WarpListMQ {
  static constexpr int kCap = ListCfg<KP>::kCap;
  uint64_t* base;      // this warp's NQ lists
  int my_j;            // the query this lane feeds
  int cnt;             // entries of list my_j (uniform over the lanes of one query)
  uint64_t tau;        // keys <= tau cannot be in the global top-KP of query my_j

  // warp-collective: sort list j (cj entries), keep the best KP, publish its KP-th key as a threshold
  __device__ __forceinline__ void compact(int j, int cj, int lane, unsigned long long* s_tau, unsigned long long* g_tau) {
    uint64_t* buf = base + j * kCap;
    for (int i = cj + lane; i < kCap; i += 32) buf[i] = 0;
    __syncwarp();
    bitonic_desc(buf, kCap, lane, 32, [] { __syncwarp(); });
    if (cj >= KP) {
      const uint64_t thr = buf[KP - 1];
      if (lane == 0) {
        atomicMax(s_tau + j, (unsigned long long)thr);
        atomicMax(g_tau + j, (unsigned long long)thr);
      }
      if (my_j == j) { cnt = KP; if (thr > tau) tau = thr; }
    }
  }

  // warp-collective: lanes with `pass` append their key to the list of their query
  __device__ __forceinline__ void push(bool pass, uint64_t key, int lane, unsigned long long* s_tau, unsigned long long* g_tau) {
    if (__ballot_sync(kFull, pass) == 0) return;
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      const unsigned m = __ballot_sync(kFull, pass && my_j == j);
      if (m != 0 && my_j == j) {
        if (pass) base[j * kCap + cnt + __popc(m & ((1u << lane) - 1u))] = key;
        cnt += __popc(m);
      }
    }
    __syncwarp();
    // first fill: establish a threshold as soon as KP entries exist; later: only when nearly full
    unsigned need = __ballot_sync(kFull, cnt > kCap - 32 || (tau == 0 && cnt >= KP));
    while (need) {
      const int leader = __ffs(need) - 1;
      const int j = __shfl_sync(kFull, my_j, leader);
      const int cj = __shfl_sync(kFull, cnt, leader);
      compact(j, cj, lane, s_tau, g_tau);
      need &= ~__ballot_sync(kFull, my_j == j);
    }
  }
};

// End of kernel, per query: every warp's list is sorted; the CTA's top-KP of the kSweepWarps*KP head entries is found by
// rank counting (keys are unique): one pass of broadcast shared-memory reads instead of a 50-70 step bitonic sort.
template <int KP, int NQ>
__device__ __forceinline__ void cta_merge_and_store_mq(uint64_t* s_lists, WarpListMQ<KP, NQ>& wl, int lane, int warp,
                                                       unsigned long long* s_tau, unsigned long long* g_tau, uint64_t* cand_out,
                                                       int64_t cand_stride) {
  constexpr int kCap = ListCfg<KP>::kCap;
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    const int cj = __shfl_sync(kFull, wl.cnt, j * MqCfg<NQ>::kLanesPerQuery);
    wl.compact(j, cj, lane, s_tau, g_tau);                 // sorted descending, at most KP valid entries, zeros after
    for (int i = KP + lane; i < kCap; i += 32) wl.base[j * kCap + i] = 0;
  }
  for (int i = threadIdx.x; i < KP * NQ; i += kSweepThreads) cand_out[(i / KP) * cand_stride + (i % KP)] = 0;
  __syncthreads();
  constexpr int kTotal = kSweepWarps * KP;
  for (int j = 0; j < NQ; ++j) {
    for (int e = threadIdx.x; e < kTotal; e += kSweepThreads) {
      const uint64_t mine = s_lists[((e / KP) * NQ + j) * kCap + (e % KP)];
      if (mine == 0) continue;
      int rank = 0;
      for (int w = 0; w < kSweepWarps; ++w) {
        const uint64_t* lst = s_lists + (w * NQ + j) * kCap;
        for (int i = 0; i < KP; ++i) {                     // lists are sorted: stop at the first key that is not larger
          if (lst[i] > mine) ++rank; else break;
        }
        if (rank >= KP) break;
      }
      if (rank < KP) cand_out[j * cand_stride + rank] = mine;
    }
  }
}

// 4 consecutive elements (2 for double) of one stored vector against the same elements of one query
template <int MC>
__device__ __forceinline__ void accum_piece(float& a, const uint4& raw, int e4, const float4& q, __half) {
  const __half2* h = reinterpret_cast<const __half2*>(&raw) + 2 * e4;
  const float2 v0 = __half22float2(h[0]), v1 = __half22float2(h[1]);
  accum<MC, float>(a, v0.x, q.x);
  accum<MC, float>(a, v0.y, q.y);
  accum<MC, float>(a, v1.x, q.z);
  accum<MC, float>(a, v1.y, q.w);
}
template <int MC>
__device__ __forceinline__ void accum_piece(float& a, const uint4& raw, int, const float4& q, float) {
  const float* v = reinterpret_cast<const float*>(&raw);
  accum<MC, float>(a, v[0], q.x);
  accum<MC, float>(a, v[1], q.y);
  accum<MC, float>(a, v[2], q.z);
  accum<MC, float>(a, v[3], q.w);
}
template <int MC>
__device__ __forceinline__ void accum_piece(double& a, const uint4& raw, int, const double2& q, double) {
  const double* v = reinterpret_cast<const double*>(&raw);
  accum<MC, double>(a, v[0], q.x);
  accum<MC, double>(a, v[1], q.y);
}
template <typename Acc> struct PieceOf { using P = float4; static constexpr int kElems = 4; };
template <> struct PieceOf<double> { using P = double2; static constexpr int kElems = 2; };

template <typename T, int MC, int KP, bool VEC, int NQ>
__global__ void __launch_bounds__(kSweepThreads, 2) sweep_kernel(SweepParams p) {
  using Acc = typename Store<T>::Acc;
  using Piece = typename PieceOf<Acc>::P;
  constexpr int kPerVec = Store<T>::kPerVec;
  constexpr int kCap = ListCfg<KP>::kCap;
  constexpr int R = MqCfg<NQ>::kR;
  constexpr int V = MqCfg<NQ>::kV;
  constexpr int kE4 = kPerVec / PieceOf<Acc>::kElems;        // pieces per stored vector and query: 2 (half), 1 (float, double)
  constexpr int NP = kE4 * NQ;                               // pieces a lane needs per column step
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* s_lists = reinterpret_cast<uint64_t*>(smem_raw);                       // [warps][NQ][kCap]
  unsigned long long* s_tau = reinterpret_cast<unsigned long long*>(s_lists + kSweepWarps * NQ * kCap);   // [NQ], padded to 16 B
  Acc* s_q = reinterpret_cast<Acc*>(s_tau + ((NQ + 1) & ~1));

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < NQ) s_tau[threadIdx.x] = 0;
  // Query tile.  VEC with several queries: piece-major so that the 32 lanes of a warp read 32 CONSECUTIVE 16-byte pieces (conflict-free
  // LDS.128): piece p = e4 * NQ + j of column step `st` for lane l sits at ((st * NP + p) * 32 + l).  Scalar path: [NQ][d].
  constexpr bool kHalfDot = std::is_same<T, __half>::value && MC == 0 && VEC;
  const bool qh = kHalfDot && p.q_half != 0;                 // float16 query tile + mixed-precision FMA (accum_hh)
  if (qh) {
    // [step][query][lane] pieces of 8 float16 values (one stored vector's worth): conflict-free LDS.128, one per 8 elements
    __half* s_qh = reinterpret_cast<__half*>(s_q);
    const int steps = (p.nvec + 31) / 32;
    for (int i = threadIdx.x; i < steps * NQ * 32 * 8; i += kSweepThreads) {
      const int t = i % 8, l = (i / 8) % 32, j = (i / 256) % NQ, st = i / (256 * NQ);
      const int c = st * 32 + l;
      s_qh[i] = (c < p.nvec) ? __float2half_rn((float)reinterpret_cast<const Acc*>(p.qa)[(int64_t)j * p.d + (int64_t)c * 8 + t]) : __float2half_rn(0.f);
    }
  } else if (VEC && NQ > 1) {
    const int steps = (p.nvec + 31) / 32;
    constexpr int kPE = PieceOf<Acc>::kElems;
    for (int i = threadIdx.x; i < steps * NP * 32 * kPE; i += kSweepThreads) {
      const int t = i % kPE, l = (i / kPE) % 32, pp = (i / (kPE * 32)) % NP, st = i / (kPE * 32 * NP);
      const int c = st * 32 + l, e4 = pp / NQ, j = pp % NQ;
      const int64_t col = (int64_t)c * kPerVec + e4 * kPE + t;
      s_q[i] = (c < p.nvec) ? reinterpret_cast<const Acc*>(p.qa)[(int64_t)j * p.d + col] : Acc(0);
    }
  } else {
    for (int64_t i = threadIdx.x; i < (int64_t)NQ * p.d; i += kSweepThreads) s_q[i] = reinterpret_cast<const Acc*>(p.qa)[i];
  }
  __syncthreads();

  const int my_v = lane >> MqCfg<NQ>::kRepShift;             // the (query, row) pair this lane owns after the reduce
  const int my_j = my_v / R, my_row = my_v % R;
  const bool rep = (lane & ((1 << MqCfg<NQ>::kRepShift) - 1)) == 0;
  WarpListMQ<KP, NQ> wl;
  wl.base = s_lists + (size_t)warp * NQ * kCap;
  wl.my_j = my_j;
  wl.cnt = 0;
  wl.tau = 0;

  // A warp walks windows of 32 consecutive rows (one mask word) and, inside a window, batches of up to R
  // KEPT rows: dropped rows are never loaded and R independent 16-byte loads per lane stay in flight
  // whatever the mask density is.
  const int64_t nwin = (p.n + 31) / 32;
  const int64_t wstride = (int64_t)gridDim.x * kSweepWarps;
  const Acc* inv = reinterpret_cast<const Acc*>(p.inv_norms);
  // pearson (MC == 0 only): sum_j (v_j - mean_v) b_j = v.b - mean_v * sum(b), then / (std_v * d) / std_q
  const Acc* pmeans = reinterpret_cast<const Acc*>(p.row_means);
  Acc q_sumb = Acc(0), q_scale = Acc(1);
  if (MC == 0 && p.qaux) {
    q_sumb = (Acc)p.qaux[2 * my_j + 1];
    q_scale = (p.qaux[2 * my_j] == 0.0) ? (Acc)__longlong_as_double(0x7ff8000000000000ll) : (Acc)(1.0 / p.qaux[2 * my_j]);
  }
  int since_refresh = 0;
  const int64_t g0 = (int64_t)blockIdx.x * kSweepWarps + warp;
  uint32_t next_bits = (g0 < nwin) ? window_keep_bits(p.f, g0, p.n) : 0u;
  const Piece* s_q4 = reinterpret_cast<const Piece*>(s_q);

  for (int64_t g = g0; g < nwin; g += wstride) {
    uint32_t bits = next_bits;
    next_bits = (g + wstride < nwin) ? window_keep_bits(p.f, g + wstride, p.n) : 0u;    // prefetch the next mask word
    if (bits == 0) continue;
    // refresh the threshold from the CTA (cheap) and, now and then, from the grid
    {
      unsigned long long t = *reinterpret_cast<volatile unsigned long long*>(s_tau + my_j);
      if (++since_refresh >= 4) {
        since_refresh = 0;
        unsigned long long gt = *reinterpret_cast<volatile unsigned long long*>(p.tau + my_j);
        if (gt > t) { t = gt; if ((lane & (MqCfg<NQ>::kLanesPerQuery - 1)) == 0) atomicMax(s_tau + my_j, gt); }
      }
      if (t > wl.tau) wl.tau = t;
    }
    const int64_t row0 = g * 32;
    const char* base = p.rows + row0 * p.row_bytes;
    while (bits) {
      // the next (up to) R kept rows of the window: warp-uniform offsets
      uint32_t ro[R];                     // byte offsets inside the window (32 rows < 4 GB)
      unsigned keep = 0;
      int my_off = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
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

      Acc acc[V];                          // acc[j * R + r]
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = Acc(0);

      if (VEC && NQ >= 4) {
        // NQ * R * kPerVec accumulate operations per lane and column step: the step is long enough to hide one HBM round
        // trip, so the rows of step st + 1 are requested before step st is consumed.  Two register buffers used in turn
        // (no copies), one 64-bit pointer per row advanced once per pair of steps (the loads use immediate offsets), full
        // steps without any predicate: a row slot that holds no kept row reads the window's first row (valid memory, its
        // sums are never pushed).  Only the last, partial step (d not a multiple of 32 vectors) zero-fills by lane.
        const char* rp[R];
#pragma unroll
        for (int r = 0; r < R; ++r) rp[r] = base + ro[r] + (uint32_t)lane * 16u;
        const int nsteps = (p.nvec + 31) >> 5, full = p.nvec >> 5;
        auto load = [&](uint4 (&dst)[R], int step, int byte_off) {
          if (step < full) {
#pragma unroll
            for (int r = 0; r < R; ++r) dst[r] = ld_stream16(rp[r] + byte_off);
          } else {
            const bool ok = step * 32 + lane < p.nvec;
#pragma unroll
            for (int r = 0; r < R; ++r) dst[r] = ok ? ld_stream16(rp[r] + byte_off) : make_uint4(0, 0, 0, 0);
          }
        };
        auto compute = [&](const uint4 (&src)[R], int step) {
          if (qh) {
            const uint4* qp = reinterpret_cast<const uint4*>(s_q) + (size_t)step * NQ * 32 + lane;
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
              const uint4 q = qp[j * 32];
#pragma unroll
              for (int r = 0; r < R; ++r) {
                if constexpr (kHalfDot) accum_hh(acc[j * R + r], src[r], q);
              }
            }
            return;
          }
          const Piece* qp = s_q4 + (size_t)step * NP * 32 + lane;
#pragma unroll
          for (int e4 = 0; e4 < kE4; ++e4) {
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
              const Piece q = qp[(e4 * NQ + j) * 32];
#pragma unroll
              for (int r = 0; r < R; ++r) accum_piece<MC>(acc[j * R + r], src[r], e4, q, T());
            }
          }
        };
        uint4 b0[R], b1[R];
        load(b0, 0, 0);
        int st = 0;
#pragma unroll 1
        for (; st + 2 <= nsteps; st += 2) {
          load(b1, st + 1, 512);
          compute(b0, st);
          if (st + 2 < nsteps) load(b0, st + 2, 1024);
          compute(b1, st + 1);
#pragma unroll
          for (int r = 0; r < R; ++r) rp[r] += 1024;
        }
        if (st < nsteps) compute(b0, st);
      } else if (VEC && NQ == 1 && qh) {
        // one float16 query against float16 rows: 8 mixed-precision FMAs per 16-byte vector, no conversions
        const uint4* s_qv = reinterpret_cast<const uint4*>(s_q);
#pragma unroll 2
        for (int c = lane; c < p.nvec; c += 32) {
          uint4 raw[R];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if ((keep >> r) & 1u) raw[r] = ld_stream16(base + (ro[r] + (uint32_t)c * 16u));
            else raw[r] = make_uint4(0, 0, 0, 0);
          }
          const uint4 q = s_qv[c];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if constexpr (kHalfDot) accum_hh(acc[r], raw[r], q);
          }
        }
      } else if (VEC && NQ == 1) {
        // one query: plain [d] query tile, two column steps (16 independent 16-byte loads per lane) in flight
#pragma unroll 2
        for (int c = lane; c < p.nvec; c += 32) {
          uint4 raw[R];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if ((keep >> r) & 1u) raw[r] = ld_stream16(base + (ro[r] + (uint32_t)c * 16u));
            else raw[r] = make_uint4(0, 0, 0, 0);
          }
          Acc q[kPerVec];
#pragma unroll
          for (int i = 0; i < kPerVec; ++i) q[i] = s_q[c * kPerVec + i];
#pragma unroll
          for (int r = 0; r < R; ++r) accum_vec<MC>(acc[r], raw[r], q, T());
        }
      } else if (VEC) {
        int st = 0;
#pragma unroll 1
        for (int c = lane; c < p.nvec; c += 32, ++st) {
          uint4 raw[R];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if ((keep >> r) & 1u) raw[r] = ld_stream16(base + (ro[r] + (uint32_t)c * 16u));
            else raw[r] = make_uint4(0, 0, 0, 0);
          }
          if (qh) {
            const uint4* qv = reinterpret_cast<const uint4*>(s_q) + (size_t)st * NQ * 32 + lane;
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
              const uint4 q = qv[j * 32];
#pragma unroll
              for (int r = 0; r < R; ++r) {
                if constexpr (kHalfDot) accum_hh(acc[j * R + r], raw[r], q);
              }
            }
            continue;
          }
          const Piece* qp = s_q4 + (size_t)st * NP * 32 + lane;
#pragma unroll
          for (int e4 = 0; e4 < kE4; ++e4) {
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
              const Piece q = qp[(e4 * NQ + j) * 32];
#pragma unroll
              for (int r = 0; r < R; ++r) accum_piece<MC>(acc[j * R + r], raw[r], e4, q, T());
            }
          }
        }
      } else {
        for (int c = lane; c < p.nvec; c += 32) {
          Acc v[R];
#pragma unroll
          for (int r = 0; r < R; ++r) v[r] = ((keep >> r) & 1u) ? (Acc) reinterpret_cast<const T*>(base + ro[r])[c] : Acc(0);
#pragma unroll
          for (int j = 0; j < NQ; ++j) {
            const Acc q = s_q[(int64_t)j * p.d + c];
#pragma unroll
            for (int r = 0; r < R; ++r) accum<MC, Acc>(acc[j * R + r], v[r], ((keep >> r) & 1u) ? q : Acc(0));
          }
        }
      }

      Acc total = reduce_transpose<V, Acc>(acc, lane);
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
      const uint64_t key = ordered_key(p.f, score, (uint32_t)mrow, rep && mine_kept, wl.tau);
      wl.push(rep && mine_kept && key > wl.tau, key, lane, s_tau, p.tau);
    }
  }

  cta_merge_and_store_mq<KP, NQ>(s_lists, wl, lane, warp, s_tau, p.tau, p.cand + (int64_t)blockIdx.x * KP, p.cand_stride);
}

template <typename T, int MC, int KP, bool VEC, int NQ>
static int launch_one(const SweepParams& p, int grid, size_t smem, cudaStream_t s) {
  auto kern = sweep_kernel<T, MC, KP, VEC, NQ>;
  if (smem > 48 * 1024) HDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kSweepThreads, smem, s>>>(p);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}
template <typename T, int MC, int KP, bool VEC>
static int launch_nq(const SweepParams& p, int nq, int grid, size_t smem, cudaStream_t s) {
  if (nq == 1) return launch_one<T, MC, KP, VEC, 1>(p, grid, smem, s);
  if (nq == 2) return launch_one<T, MC, KP, VEC, 2>(p, grid, smem, s);
  if (nq == 4) return launch_one<T, MC, KP, VEC, 4>(p, grid, smem, s);
  if constexpr (KP <= 32 && sizeof(typename Store<T>::Acc) == 4) {
    if (nq == 8) return launch_one<T, MC, KP, VEC, 8>(p, grid, smem, s);
  }
  return fail("sweep: unsupported query group");
}
template <typename T, int MC, int KP>
static int launch_vec(const SweepParams& p, bool vec, int nq, int grid, size_t smem, cudaStream_t s) {
  return vec ? launch_nq<T, MC, KP, true>(p, nq, grid, smem, s) : launch_nq<T, MC, KP, false>(p, nq, grid, smem, s);
}
}  // namespace hdb
