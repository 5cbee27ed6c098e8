// The streaming sweep with TMA-staged row tiles (one query per pass, float16 / float32 rows of at most 4 KB).
//
// csrc/sweep_float.cuh keeps the row stream in registers: a warp requests 8 rows, waits one HBM round trip, reduces, selects
// and only then requests the next 8 -- a dependent chain per batch.  That is the HBM rate for rows of a few KB, but a row
// RATE limit (~6 G rows/s per GPU) for short rows: 100M x 384 fp16 (768-byte rows, BASELINE.json config C4) streams at
// 4.8 TB/s.  Here the row stream is decoupled from the arithmetic:
//   * a tile = TR = 32 / 16 / 8 consecutive rows = ONE contiguous piece of HBM of at most 32 KB;
//   * a producer warp (one elected lane) moves tiles into a ring of shared-memory stages with cp.async.bulk (the TMA engine,
//     no registers, no address arithmetic per 16 bytes) and mbarrier transaction counts;
//   * the 8 consumer warps take TR / 8 rows of every tile each, read them with conflict-free LDS.128 against the query tile
//     (piece-major in shared memory), reduce with the transposing butterfly and feed the same per-warp candidate lists /
//     grid-wide threshold as the register sweep.  A stage is released as soon as its rows are in registers' sums, before
//     the select step.
// Row subsets: a tile whose rows are all dropped is neither loaded nor visited; inside a loaded tile dropped rows are simply
// not pushed.  The host picks this kernel when no mask is set or the mask is tile-dense (clustered storage,
// hdb_matrix_set_row_order); sparse random masks stay on the register sweep, which never loads a dropped row.
#pragma once
#include "sweep_float.cuh"

namespace hdb {

constexpr int kStagedThreads = kSweepThreads + 32;      // 8 consumer warps + 1 producer warp
constexpr int kStagedMaxStages = 4;

__device__ __forceinline__ uint32_t st_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void st_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(st_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void st_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = st_smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
// one contiguous piece of global memory -> shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void st_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(st_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(st_smem_u32(bar)) : "memory");
}

struct StagedGeom {
  int stages;          // ring depth (2 .. kStagedMaxStages)
  int stage_bytes;     // TR * row_bytes rounded up to 128
};

// keep bits of tile t (TR rows starting at row t * TR), bit i = row t * TR + i
template <int TR>
__device__ __forceinline__ uint32_t tile_keep_bits(const RowFilter& f, int64_t t, int64_t n) {
  const int64_t row0 = t * TR;
  const uint32_t w = window_keep_bits(f, row0 >> 5, n);
  if (TR == 32) return w;
  return (w >> (uint32_t)(row0 & 31)) & ((1u << TR) - 1u);
}

template <typename T, int MC, int KP, int RPW>
__global__ void __launch_bounds__(kStagedThreads, 2) sweep_staged_kernel(SweepParams p, StagedGeom geo) {
  using Acc = typename Store<T>::Acc;
  using Piece = typename PieceOf<Acc>::P;
  constexpr int kPerVec = Store<T>::kPerVec;
  constexpr int kCap = ListCfg<KP>::kCap;
  constexpr int TR = 8 * RPW;
  constexpr int kE4 = kPerVec / PieceOf<Acc>::kElems;        // query pieces per stored vector: 2 (half), 1 (float)
  constexpr int kRepShift = RPW == 4 ? 3 : (RPW == 2 ? 4 : 5);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* s_lists = reinterpret_cast<uint64_t*>(smem_raw);                                    // [warps][kCap]
  unsigned long long* s_tau = reinterpret_cast<unsigned long long*>(s_lists + kSweepWarps * kCap);
  uint64_t* s_full = reinterpret_cast<uint64_t*>(s_tau + 2);                                    // [kStagedMaxStages]
  uint64_t* s_empty = s_full + kStagedMaxStages;
  Piece* s_q4 = reinterpret_cast<Piece*>(s_empty + kStagedMaxStages);                           // piece-major query tile
  const int steps = (p.nvec + 31) / 32;
  unsigned char* s_stage = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(s_q4 + (size_t)steps * kE4 * 32) + 127) & ~uintptr_t(127));

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    *s_tau = 0;
    for (int s = 0; s < geo.stages; ++s) { st_mbar_init(&s_full[s], 1); st_mbar_init(&s_empty[s], kSweepWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr bool kHalfDot = std::is_same<T, __half>::value && MC == 0;
  const bool qh = kHalfDot && p.q_half != 0;                 // float16 query tile + mixed-precision FMA (accum_hh)
  if (qh) {
    __half* s_qh = reinterpret_cast<__half*>(s_q4);          // [step][lane] pieces of 8 float16 values
    for (int i = threadIdx.x; i < steps * 32 * 8; i += kStagedThreads) {
      const int c = i / 8;
      s_qh[i] = (c < p.nvec) ? __float2half_rn((float)reinterpret_cast<const Acc*>(p.qa)[i]) : __float2half_rn(0.f);
    }
  } else {
    constexpr int kPE = PieceOf<Acc>::kElems;
    Acc* s_q = reinterpret_cast<Acc*>(s_q4);
    for (int i = threadIdx.x; i < steps * kE4 * 32 * kPE; i += kStagedThreads) {
      const int t = i % kPE, l = (i / kPE) % 32, e4 = (i / (kPE * 32)) % kE4, st = i / (kPE * 32 * kE4);
      const int c = st * 32 + l;
      const int64_t col = (int64_t)c * kPerVec + e4 * kPE + t;
      s_q[i] = (c < p.nvec) ? reinterpret_cast<const Acc*>(p.qa)[col] : Acc(0);
    }
  }
  __syncthreads();

  const int64_t ntiles = (p.n + TR - 1) / TR;
  const uint32_t tile_bytes = (uint32_t)(TR * p.row_bytes);
  // no mask, the whole shard kept: every tile is visited and only the last one can be partial
  const bool simple = p.f.mask == nullptr && p.f.lo <= 0 && p.f.hi >= p.n;
  const uint32_t all_rows = TR == 32 ? 0xffffffffu : ((1u << TR) - 1u);
  if (warp == kSweepWarps) {
    // ---- producer: the warp fetches the keep bits of 32 of this CTA's tiles at once (one per lane: with a mask every word
    //      is a global load, and a chain of one load per tile would pace the ring), then one lane issues the non-empty ones
    int s = 0;
    uint32_t wait_parity = 1;              // a fresh mbarrier passes a wait on parity 1: the first round finds every slot free
    for (int64_t t0 = blockIdx.x; t0 < ntiles; t0 += 32 * (int64_t)gridDim.x) {
      const int64_t tl = t0 + (int64_t)lane * gridDim.x;
      uint32_t mybits = 0;
      if (tl < ntiles) mybits = (simple && tl != ntiles - 1) ? all_rows : tile_keep_bits<TR>(p.f, tl, p.n);
      unsigned nonempty = __ballot_sync(kFull, mybits != 0);
      if (lane == 0) {
        while (nonempty) {
          const int j = __ffs(nonempty) - 1;
          nonempty &= nonempty - 1;
          const int64_t t = t0 + (int64_t)j * gridDim.x;
          st_mbar_wait(&s_empty[s], wait_parity);
          const int64_t row0 = t * TR;
          const uint32_t bytes = (t == ntiles - 1) ? (uint32_t)((p.n - row0) * p.row_bytes) : tile_bytes;
          st_mbar_expect_tx(&s_full[s], bytes);
          st_bulk_load(s_stage + (size_t)s * geo.stage_bytes, p.rows + row0 * p.row_bytes, bytes, &s_full[s]);
          if (++s == geo.stages) { s = 0; wait_parity ^= 1u; }
        }
      }
      __syncwarp();
    }
  } else {
    // ---- consumers
    WarpList<KP> wl;
    wl.buf = s_lists + warp * kCap;
    wl.cnt = 0;
    wl.tau = 0;
    const int my_v = lane >> kRepShift;                        // the row (of this warp's RPW) this lane owns after the reduce
    const bool rep = (lane & ((1 << kRepShift) - 1)) == 0;
    const Acc* inv = reinterpret_cast<const Acc*>(p.inv_norms);
    const Acc* pmeans = reinterpret_cast<const Acc*>(p.row_means);
    Acc q_sumb = Acc(0), q_scale = Acc(1);
    if (MC == 0 && p.qaux) {
      q_sumb = (Acc)p.qaux[1];
      q_scale = (p.qaux[0] == 0.0) ? (Acc)__longlong_as_double(0x7ff8000000000000ll) : (Acc)(1.0 / p.qaux[0]);
    }
    int s = 0, since_refresh = 0;
    uint32_t parity = 0;
    const uint32_t my_rows_off = (uint32_t)(warp * RPW) * (uint32_t)p.row_bytes;
    const int my_local = warp * RPW + my_v;
    // the keep bits of 32 tiles at once (one per lane), then the non-empty ones in order: an empty tile costs nothing
    for (int64_t t0 = blockIdx.x; t0 < ntiles; t0 += 32 * (int64_t)gridDim.x) {
      const int64_t tl = t0 + (int64_t)lane * gridDim.x;
      uint32_t mybits = 0;
      if (tl < ntiles) mybits = (simple && tl != ntiles - 1) ? all_rows : tile_keep_bits<TR>(p.f, tl, p.n);
      unsigned nonempty = __ballot_sync(kFull, mybits != 0);
      while (nonempty) {
      const int jt = __ffs(nonempty) - 1;
      nonempty &= nonempty - 1;
      const uint32_t bits = __shfl_sync(kFull, mybits, jt);
      const int64_t t = t0 + (int64_t)jt * gridDim.x;
      const int cur = s;
      const uint32_t cur_parity = parity;
      if (++s == geo.stages) { s = 0; parity ^= 1u; }
      {
        unsigned long long tt = *reinterpret_cast<volatile unsigned long long*>(s_tau);
        if (++since_refresh >= 4) {
          since_refresh = 0;
          unsigned long long gt = *reinterpret_cast<volatile unsigned long long*>(p.tau);
          if (gt > tt) { tt = gt; if (lane == 0) atomicMax(s_tau, gt); }
        }
        if (tt > wl.tau) wl.tau = tt;
      }
      // per-row side inputs of the row this lane will own, requested before the tile is waited for
      const bool mine_kept = (bits >> my_local) & 1u;
      const int64_t mrow = t * TR + my_local;
      Acc my_inv = Acc(1), my_mean = Acc(0);
      double my_decay = 0.0;
      if (rep && mine_kept) {
        if (inv) my_inv = inv[mrow];
        if (MC == 0 && pmeans) my_mean = pmeans[mrow];
        if (p.f.decay) my_decay = p.f.decay[mrow];
      }
      st_mbar_wait(&s_full[cur], cur_parity);
      const unsigned char* rows = s_stage + (uint32_t)cur * (uint32_t)geo.stage_bytes + my_rows_off + (uint32_t)lane * 16u;
      Acc acc[RPW];
#pragma unroll
      for (int r = 0; r < RPW; ++r) acc[r] = Acc(0);
      int st = 0;
      if (qh) {
        const uint4* s_qv = reinterpret_cast<const uint4*>(s_q4);
#pragma unroll 2
        for (int c = lane; c < p.nvec; c += 32, ++st) {
          const uint4 q = s_qv[c];
#pragma unroll
          for (int r = 0; r < RPW; ++r) {
            const uint4 raw = *reinterpret_cast<const uint4*>(rows + (uint32_t)r * (uint32_t)p.row_bytes + (uint32_t)st * 512u);
            if constexpr (kHalfDot) accum_hh(acc[r], raw, q);
          }
        }
      } else
#pragma unroll 2
      for (int c = lane; c < p.nvec; c += 32, ++st) {
        Piece q[kE4];
#pragma unroll
        for (int e4 = 0; e4 < kE4; ++e4) q[e4] = s_q4[(st * kE4 + e4) * 32 + lane];
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
          const uint4 raw = *reinterpret_cast<const uint4*>(rows + (uint32_t)r * (uint32_t)p.row_bytes + (uint32_t)st * 512u);
#pragma unroll
          for (int e4 = 0; e4 < kE4; ++e4) accum_piece<MC>(acc[r], raw, e4, q[e4], T());
        }
      }
      __syncwarp();
      if (lane == 0) st_mbar_arrive(&s_empty[cur]);               // the stage may be refilled while this warp selects
      Acc total = reduce_transpose<RPW, Acc>(acc, lane);
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
    // every warp's list sorted, at most KP valid entries, zeros after
    wl.compact(lane, s_tau, p.tau);
    for (int i = KP + lane; i < kCap; i += 32) wl.buf[i] = 0;
  }
  uint64_t* cand_out = p.cand + (int64_t)blockIdx.x * KP;
  for (int i = threadIdx.x; i < KP; i += kStagedThreads) cand_out[i] = 0;
  __syncthreads();
  // the CTA's top-KP of the 8 * KP head entries by rank counting (keys are unique)
  constexpr int kTotal = kSweepWarps * KP;
  for (int e = threadIdx.x; e < kTotal; e += kStagedThreads) {
    const uint64_t mine = s_lists[(e / KP) * kCap + (e % KP)];
    if (mine == 0) continue;
    int rank = 0;
    for (int w = 0; w < kSweepWarps; ++w) {
      const uint64_t* lst = s_lists + w * kCap;
      for (int j = 0; j < KP; ++j) {
        if (lst[j] > mine) ++rank; else break;
      }
      if (rank >= KP) break;
    }
    if (rank < KP) cand_out[rank] = mine;
  }
}

// rows per tile for a row of `row_bytes` bytes (tiles of at most 32 KB), 0 = not eligible
inline int staged_tile_rows(int64_t row_bytes) {
  if (row_bytes <= 0 || row_bytes % 16 != 0) return 0;
  if (row_bytes <= 1024) return 32;
  if (row_bytes <= 2048) return 16;
  if (row_bytes <= 4096) return 8;
  return 0;
}

template <typename T, int MC, int KP, int RPW>
static int launch_staged_one(const SweepParams& p, int grid, cudaStream_t s) {
  using Acc = typename Store<T>::Acc;
  constexpr int kCap = ListCfg<KP>::kCap;
  constexpr int kE4 = Store<T>::kPerVec / PieceOf<Acc>::kElems;
  const int steps = (p.nvec + 31) / 32;
  const size_t fixed = (size_t)kSweepWarps * kCap * 8 + 16 + 2 * kStagedMaxStages * 8 + (size_t)steps * kE4 * 32 * 16 + 128;
  StagedGeom geo;
  geo.stage_bytes = (int)(((int64_t)8 * RPW * p.row_bytes + 127) & ~int64_t(127));
  const size_t budget = 110 * 1024;                                  // two CTAs per SM
  if (fixed + 2 * (size_t)geo.stage_bytes > budget) return fail("staged sweep: the tile ring does not fit");
  geo.stages = (int)((budget - fixed) / geo.stage_bytes);
  if (geo.stages > kStagedMaxStages) geo.stages = kStagedMaxStages;
  const size_t smem = fixed + (size_t)geo.stages * geo.stage_bytes;
  auto kern = sweep_staged_kernel<T, MC, KP, RPW>;
  HDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kStagedThreads, smem, s>>>(p, geo);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

template <typename T, int MC>
static int launch_staged(const SweepParams& p, int kp, int grid, cudaStream_t s) {
  const int tr = staged_tile_rows(p.row_bytes);
  if (kp <= 32) {
    if (tr == 32) return launch_staged_one<T, MC, 32, 4>(p, grid, s);
    if (tr == 16) return launch_staged_one<T, MC, 32, 2>(p, grid, s);
    if (tr == 8) return launch_staged_one<T, MC, 32, 1>(p, grid, s);
  } else {
    if (tr == 32) return launch_staged_one<T, MC, 128, 4>(p, grid, s);
    if (tr == 16) return launch_staged_one<T, MC, 128, 2>(p, grid, s);
    if (tr == 8) return launch_staged_one<T, MC, 128, 1>(p, grid, s);
  }
  return fail("staged sweep: row size not eligible");
}

}  // namespace hdb
