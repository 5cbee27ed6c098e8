#pragma once
// Bit-packed popcount sweeps (hamming_distance, jaccard_similarity: hyperdb/ranking_algorithm.py:128-147, :63-76) with the
// same fused select as the float sweeps; up to 4 queries per pass.
#include <cstdlib>

#include "sweep_common.cuh"

#ifndef HDB_HAM_JAC
#define HDB_HAM_JAC 0      // the metric of this translation unit (csrc/sweep_hamming_inst.cu): 0 hamming, 1 jaccard
#endif

namespace hdb {

// ---------------------------------------------------------------------------------------------
// Hamming sweep on the bit-packed matrix: words (multiple of 4) u32 per row = nvec 16-byte vectors.
// LPR lanes share a row (LPR = smallest power of two >= nvec, capped at 32); score = D - popcount(xor).
// ---------------------------------------------------------------------------------------------
struct HammingParams {
  const uint32_t* bits;
  const uint32_t* qbits;     // [NQ][words] sign bits of the queries of this pass
  int64_t n, d;
  int nvec, lpr;
  int jaccard;               // 0: score = d - popcount(xor); 1: score = popcount(and) / popcount(or)
  RowFilter f;
  uint64_t* cand;
  unsigned long long* tau;   // [NQ]
  int64_t cand_stride;       // keys between the candidate blocks of consecutive queries (grid * KP)
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
      const uint64_t key = ordered_key(p.f, score, (uint32_t)row, sub == 0 && kept[r], wl.tau);
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

template <int KP, int LPR, bool JAC, int NQ>
__global__ void __launch_bounds__(kSweepThreads, 2) sweep_hamming_lpr_kernel(HammingParams p) {
  constexpr int kCap = ListCfg<KP>::kCap;
  constexpr int PPR = LPR < 8 ? LPR : 8;       // passes per round = loads in flight per lane
  constexpr int ROUNDS = LPR / PPR;
  constexpr int RPP = 32 / LPR;                // rows per pass
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* s_lists = reinterpret_cast<uint64_t*>(smem_raw);                        // [NQ][warps][kCap]
  unsigned long long* s_tau = reinterpret_cast<unsigned long long*>(s_lists + NQ * kSweepWarps * kCap);
  uint4* s_q = reinterpret_cast<uint4*>(s_tau + ((NQ + 1) & ~1));                   // [NQ][nvec]

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < NQ) s_tau[threadIdx.x] = 0;
  for (int j = threadIdx.x; j < NQ * p.nvec; j += kSweepThreads) s_q[j] = reinterpret_cast<const uint4*>(p.qbits)[j];
  __syncthreads();

  WarpList<KP> wl[NQ];
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    wl[j].buf = s_lists + ((size_t)j * kSweepWarps + warp) * kCap;
    wl[j].cnt = 0;
    wl[j].tau = 0;
  }

  const int sub = lane % LPR, slot = lane / LPR;
  const bool lane_has = sub < p.nvec;
  uint4 my_q[NQ];
#pragma unroll
  for (int j = 0; j < NQ; ++j) my_q[j] = lane_has ? s_q[j * p.nvec + sub] : make_uint4(0, 0, 0, 0);
  const bool rep = sub < PPR;
  const int my_pass = sub & (PPR - 1);
  const int64_t nwin = (p.n + 31) / 32;
  const int64_t wstride = (int64_t)gridDim.x * kSweepWarps;
  const uint4* vbits = reinterpret_cast<const uint4*>(p.bits);
  int since_refresh = 0;
  const int64_t g0 = (int64_t)blockIdx.x * kSweepWarps + warp;
  uint32_t next_bits = (g0 < nwin) ? window_keep_bits(p.f, g0, p.n) : 0u;

  // the PPR loads of round `round` of window `win` (dropped rows / idle lanes: zeros, never pushed)
  auto load_round = [&](uint4 (&v)[PPR], int64_t win, uint32_t wbits, int round) {
#pragma unroll
    for (int r = 0; r < PPR; ++r) {
      const int loc = (round * PPR + r) * RPP + slot;
      const bool k = ((wbits >> loc) & 1u) && lane_has;
      v[r] = k ? ld_stream16(vbits + (win * 32 + loc) * p.nvec + sub) : make_uint4(0, 0, 0, 0);
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
        const bool grid_too = ++since_refresh >= 8;
        if (grid_too) since_refresh = 0;
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
          unsigned long long t = *reinterpret_cast<volatile unsigned long long*>(s_tau + j);
          if (grid_too) {
            unsigned long long gt = *reinterpret_cast<volatile unsigned long long*>(p.tau + j);
            if (gt > t) { t = gt; if (lane == 0) atomicMax(s_tau + j, gt); }
          }
          if (t > wl[j].tau) wl[j].tau = t;
        }
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
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
          const uint4 q = my_q[j];
          int cnt[PPR];
#pragma unroll
          for (int r = 0; r < PPR; ++r) {
            if (JAC)          // both popcounts packed in one int (16 bits each: d <= 4096 here), reduced together
              cnt[r] = (__popc(v[r].x & q.x) + __popc(v[r].y & q.y) + __popc(v[r].z & q.z) + __popc(v[r].w & q.w)) +
                       ((__popc(v[r].x | q.x) + __popc(v[r].y | q.y) + __popc(v[r].z | q.z) + __popc(v[r].w | q.w)) << 16);
            else
              cnt[r] = __popc(v[r].x ^ q.x) + __popc(v[r].y ^ q.y) + __popc(v[r].z ^ q.z) + __popc(v[r].w ^ q.w);
          }
          const int diff = transpose_reduce<PPR, LPR>(cnt, sub);
          // dropped rows were loaded as zeros: their counts are the query's own and are never pushed
          double exact = JAC ? (double)(diff & 0xffff) / (double)(diff >> 16) : (double)((int)p.d - diff);
          if (p.f.decay) exact += p.f.bias * my_decay;
          const float score = (float)exact;
          const uint64_t key = ordered_key(p.f, score, (uint32_t)(row0 + my_loc), rep && my_kept, wl[j].tau);
          wl[j].push(rep && my_kept && key > wl[j].tau, key, lane, s_tau + j, p.tau + j);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < PPR; ++r) cur[r] = nxt[r];
  }
#pragma unroll
  for (int j = 0; j < NQ; ++j)
    cta_merge_and_store<KP>(s_lists + (size_t)j * kSweepWarps * kCap, wl[j], lane, warp, s_tau + j, p.tau + j,
                            p.cand + j * p.cand_stride + (int64_t)blockIdx.x * KP);
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

template <int KP, int NVEC, bool JAC, int NQ>
__global__ void __launch_bounds__(kSweepThreads, 2) sweep_hamming_staged_kernel(HammingParams p) {
  constexpr int kCap = ListCfg<KP>::kCap;
  constexpr int PITCH = StagedCfg<NVEC>::kPitch;
  constexpr int STAGE = StagedCfg<NVEC>::kStageU4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* s_lists = reinterpret_cast<uint64_t*>(smem_raw);                        // [NQ][warps][kCap]
  unsigned long long* s_tau = reinterpret_cast<unsigned long long*>(s_lists + NQ * kSweepWarps * kCap);
  uint4* s_stage = reinterpret_cast<uint4*>(s_tau + ((NQ + 1) & ~1));               // [warps][2][STAGE]
  uint4* s_qb = s_stage + (size_t)kSweepWarps * 2 * STAGE;                          // [NQ][NVEC] (NQ > 1: broadcast reads)

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < NQ) s_tau[threadIdx.x] = 0;
  uint4 q[NVEC];                 // NQ == 1: the query lives in registers
#pragma unroll
  for (int c = 0; c < NVEC; ++c) q[c] = reinterpret_cast<const uint4*>(p.qbits)[c];
  if (NQ > 1)
    for (int i = threadIdx.x; i < NQ * NVEC; i += kSweepThreads) s_qb[i] = reinterpret_cast<const uint4*>(p.qbits)[i];
  __syncthreads();

  WarpList<KP> wl[NQ];
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    wl[j].buf = s_lists + ((size_t)j * kSweepWarps + warp) * kCap;
    wl[j].cnt = 0;
    wl[j].tau = 0;
  }

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
        const bool grid_too = ++since_refresh >= 8;
        if (grid_too) since_refresh = 0;
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
          unsigned long long t = *reinterpret_cast<volatile unsigned long long*>(s_tau + j);
          if (grid_too) {
            unsigned long long gt = *reinterpret_cast<volatile unsigned long long*>(p.tau + j);
            if (gt > t) { t = gt; if (lane == 0) atomicMax(s_tau + j, gt); }
          }
          if (t > wl[j].tau) wl[j].tau = t;
        }
      }
      const int64_t row = g * 32 + lane;
      const bool kept = (bits >> lane) & 1u;
      double my_decay = 0.0;
      if (p.f.decay && kept) my_decay = p.f.decay[row];
      const uint4* rowp = my_stage + buf * STAGE + lane * PITCH;
      uint4 v[NVEC];
#pragma unroll
      for (int c = 0; c < NVEC; ++c) v[c] = rowp[c];
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        int a = 0, b = 0;
#pragma unroll
        for (int c = 0; c < NVEC; ++c) {
          const uint4 qq = (NQ == 1) ? q[c] : s_qb[j * NVEC + c];
          if (JAC) {
            a += __popc(v[c].x & qq.x) + __popc(v[c].y & qq.y) + __popc(v[c].z & qq.z) + __popc(v[c].w & qq.w);
            b += __popc(v[c].x | qq.x) + __popc(v[c].y | qq.y) + __popc(v[c].z | qq.z) + __popc(v[c].w | qq.w);
          } else {
            a += __popc(v[c].x ^ qq.x) + __popc(v[c].y ^ qq.y) + __popc(v[c].z ^ qq.z) + __popc(v[c].w ^ qq.w);
          }
        }
        double exact = JAC ? (double)a / (double)b : (double)((int)p.d - a);
        if (p.f.decay) exact += p.f.bias * my_decay;
        const uint64_t key = ordered_key(p.f, (float)exact, (uint32_t)row, kept, wl[j].tau);
        wl[j].push(kept && key > wl[j].tau, key, lane, s_tau + j, p.tau + j);
      }
    }
    __syncwarp();                                         // every lane is done with `buf` before the next issue refills it
    buf ^= 1;
    bits = next_bits;
  }
  cp_async_wait<0>();
#pragma unroll
  for (int j = 0; j < NQ; ++j)
    cta_merge_and_store<KP>(s_lists + (size_t)j * kSweepWarps * kCap, wl[j], lane, warp, s_tau + j, p.tau + j,
                            p.cand + j * p.cand_stride + (int64_t)blockIdx.x * KP);
}

static size_t hamming_list_smem(int kp, int nq) { return (size_t)kSweepWarps * nq * (kp <= 32 ? 128 : 2 * kp) * 8 + (size_t)((nq + 1) & ~1) * 8; }

template <int KP, int NVEC, bool JAC, int NQ>
static int launch_hamming_staged4(const HammingParams& hp, int grid, cudaStream_t s) {
  auto kern = sweep_hamming_staged_kernel<KP, NVEC, JAC, NQ>;
  const size_t smem = hamming_list_smem(KP, NQ) + StagedCfg<NVEC>::kBytes + (size_t)NQ * NVEC * 16;
  if (smem > 48 * 1024) HDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kSweepThreads, smem, s>>>(hp);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}
template <int KP, int NVEC, bool JAC>
static int launch_hamming_staged3(const HammingParams& hp, int nq, int grid, cudaStream_t s) {
  if (nq == 1) return launch_hamming_staged4<KP, NVEC, JAC, 1>(hp, grid, s);
  if (nq == 2) return launch_hamming_staged4<KP, NVEC, JAC, 2>(hp, grid, s);
  if constexpr (KP <= 32) { if (nq == 4) return launch_hamming_staged4<KP, NVEC, JAC, 4>(hp, grid, s); }
  return fail("sweep: unsupported hamming query group");
}
template <int KP, int NVEC>
static int launch_hamming_staged2(const HammingParams& hp, int nq, int grid, cudaStream_t s) {
  return launch_hamming_staged3<KP, NVEC, HDB_HAM_JAC != 0>(hp, nq, grid, s);
}
template <int KP>
static int launch_hamming_staged(const HammingParams& hp, int nq, int grid, cudaStream_t s) {
  switch (hp.nvec) {
    case 1: return launch_hamming_staged2<KP, 1>(hp, nq, grid, s);
    case 2: return launch_hamming_staged2<KP, 2>(hp, nq, grid, s);
    case 3: return launch_hamming_staged2<KP, 3>(hp, nq, grid, s);
    case 4: return launch_hamming_staged2<KP, 4>(hp, nq, grid, s);
    case 5: return launch_hamming_staged2<KP, 5>(hp, nq, grid, s);
    case 6: return launch_hamming_staged2<KP, 6>(hp, nq, grid, s);
    case 7: return launch_hamming_staged2<KP, 7>(hp, nq, grid, s);
    default: return launch_hamming_staged2<KP, 8>(hp, nq, grid, s);
  }
}

template <int KP, int LPR, bool JAC, int NQ>
static int launch_hamming_lpr3(const HammingParams& hp, int grid, cudaStream_t s) {
  auto kern = sweep_hamming_lpr_kernel<KP, LPR, JAC, NQ>;
  const size_t smem = hamming_list_smem(KP, NQ) + (size_t)NQ * hp.nvec * 16;
  if (smem > 48 * 1024) HDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kSweepThreads, smem, s>>>(hp);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}
template <int KP, int LPR, bool JAC>
static int launch_hamming_lpr2(const HammingParams& hp, int nq, int grid, cudaStream_t s) {
  if (nq == 1) return launch_hamming_lpr3<KP, LPR, JAC, 1>(hp, grid, s);
  if (nq == 2) return launch_hamming_lpr3<KP, LPR, JAC, 2>(hp, grid, s);
  if constexpr (KP <= 32) { if (nq == 4) return launch_hamming_lpr3<KP, LPR, JAC, 4>(hp, grid, s); }
  return fail("sweep: unsupported hamming query group");
}
template <int KP, int LPR>
static int launch_hamming_lpr(const HammingParams& hp, int nq, int grid, cudaStream_t s) {
  return launch_hamming_lpr2<KP, LPR, HDB_HAM_JAC != 0>(hp, nq, grid, s);
}
template <int KP>
static int launch_hamming_kp(const HammingParams& hp, int nq, int grid, cudaStream_t s) {
  switch (hp.lpr) {
    case 1: return launch_hamming_lpr<KP, 1>(hp, nq, grid, s);
    case 2: return launch_hamming_lpr<KP, 2>(hp, nq, grid, s);
    case 4: return launch_hamming_lpr<KP, 4>(hp, nq, grid, s);
    case 8: return launch_hamming_lpr<KP, 8>(hp, nq, grid, s);
    case 16: return launch_hamming_lpr<KP, 16>(hp, nq, grid, s);
    default: return launch_hamming_lpr<KP, 32>(hp, nq, grid, s);
  }
}

}  // namespace hdb
