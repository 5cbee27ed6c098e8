// Batched dot / cosine as a tcgen05 tensor-core contraction with a fused threshold-select epilogue.
//
// Replaces B independent calls of np.dot(vectors, q) (hyperdb/ranking_algorithm.py:29,41; the reference has no
// batched API) by S[row, b] = sum_d V[row, d] * Q[b, d] on the 5th-generation tensor cores:
//   * operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a 4-stage shared-memory ring,
//   * tcgen05.mma (M = 128 rows x N = BN queries x K = 16) issued by one elected thread, fp32 accumulators in
//     TMEM, double buffered (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1,
//   * epilogue warps read TMEM with tcgen05.ld (thread = row, registers = queries), apply the cosine inverse
//     norm / time decay / keep mask, and either
//       - DENSE mode: write the scores of a strided row SAMPLE query-major (coalesced) so that a per-query
//         threshold tau0[b] = KP-th best of the sample can be selected, or
//       - SELECT mode: append (score, row) keys with score >= tau0[b] to a per-query candidate buffer.
//     The N x B score matrix is never written.  finalize.cu then re-scores the best KP candidates of every query
//     in the reference's exact arithmetic and certifies the top-k exactly as for the streaming sweep.
// Tile order: query tile fastest, so the CTAs running at the same time share the V row tile through L2 and the
// matrix is read from HBM once; Q stays L2 resident.
#include <cuda.h>
#include <cstdlib>

#include "hdb_common.cuh"
#include "hdb_internal.h"
#include "../../include/hyperdb_b200.h"

namespace hdb {

constexpr int kTcEpiWarps = 16;          // four epilogue warps per TMEM lane quarter, each takes a quarter of the columns
constexpr int kTcThreads = 64 + 32 * kTcEpiWarps;   // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, rest: epilogue
constexpr int kTcStages = 4;
constexpr int kTileM = 128;
constexpr int kTileKBytes = 128;         // one 128-byte swizzle span of K per stage row

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <bool TF32>
__device__ __forceinline__ void umma(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if (TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
  }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (8-row groups 1024 bytes apart)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);      // start address
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
// kind::f16 / kind::tf32 instruction descriptor: fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool tf32) {
  return (1u << 4) | ((tf32 ? 2u : 0u) << 7) | ((tf32 ? 2u : 0u) << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
struct TcParamsBase {
  int64_t n, d;                 // shard rows, dim
  int64_t nq;                   // queries in this batch (padded to BN by TMA zero fill)
  int64_t n_tiles_m, n_tiles_q; // row tiles visited (all, or the sample), query tiles
  int64_t sample_stride;        // DENSE mode: visited row tile i is matrix row tile i * sample_stride
  const char* rows;             // matrix base (for the L2 prefetch)
  const float* inv_norms;       // cosine or nullptr
  const float* sqnorms;         // euclidean (norm expansion) or nullptr
  const float* qsq;             // [nq] ||q||^2 for euclidean, else nullptr
  RowFilter f;
};
// ---------------------------------------------------------------------------------------------
// SELECT epilogue of one accumulator tile for one warp: columns [c_begin, c_end) of this warp's 32 TMEM lanes.
// Branch-free screening (FFMA + predicate-OR per element), the next chunk's tcgen05.ld in flight while the
// current one is screened, and the append path only for lanes that actually hold a candidate.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Candidates found by the epilogue go to a CTA-private record buffer: the slot comes from a SHARED-memory counter
// (tens of cycles) and the 16-byte record {key, query} is a fire-and-forget global store, so no global atomic
// round trip ever sits on the epilogue's critical path.  bucket_records_kernel distributes the records to the
// per-query candidate lists afterwards, where the atomics are throughput- not latency-bound.
struct SelectSink {
  uint4* rec;              // this CTA's records
  unsigned* s_count;       // shared-memory counter
  unsigned cap;
  const float* qsq;        // euclidean: the screened value is |q|^2 - d^2; the key holds the similarity 1/(1+d)
};

__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return v;
}

// Screen 32 columns of this warp's 32 rows.  Fast path: FFMA + compare per element into a per-lane bit mask (no
// branches).  Slow path (some lane found a candidate): warp-uniform loop over the union of the masks, the column is
// re-read from TMEM with a 1-column tcgen05.ld (registers cannot be indexed dynamically), the owners append.
__device__ __forceinline__ void screen_chunk(const uint32_t (&r)[32], uint32_t taddr_chunk, const float* tau, float inv, float dec,
                                             int64_t b0, uint32_t row, const SelectSink& sink) {
  const float4* tq = reinterpret_cast<const float4*>(tau);
  uint32_t mask = 0;
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 tv = tq[j4];
    mask |= (fmaf(__uint_as_float(r[j4 * 4 + 0]), inv, dec) >= tv.x ? 1u : 0u) << (j4 * 4 + 0);
    mask |= (fmaf(__uint_as_float(r[j4 * 4 + 1]), inv, dec) >= tv.y ? 1u : 0u) << (j4 * 4 + 1);
    mask |= (fmaf(__uint_as_float(r[j4 * 4 + 2]), inv, dec) >= tv.z ? 1u : 0u) << (j4 * 4 + 2);
    mask |= (fmaf(__uint_as_float(r[j4 * 4 + 3]), inv, dec) >= tv.w ? 1u : 0u) << (j4 * 4 + 3);
  }
  uint32_t wmask = __reduce_or_sync(kFull, mask);
  while (wmask) {
    const int j = __ffs(wmask) - 1;
    wmask &= wmask - 1;
    float s = fmaf(__uint_as_float(tmem_ld1(taddr_chunk + (uint32_t)j)), inv, dec);
    if ((mask >> j) & 1u) {
      if (sink.qsq) {
        const float d2 = sink.qsq[b0 + j] - s;
        s = 1.f / (1.f + sqrtf(d2 > 0.f ? d2 : 0.f));
      }
      const unsigned pos = atomicAdd(sink.s_count, 1u);
      const uint64_t key = make_key(s, row);
      if (pos < sink.cap) sink.rec[pos] = make_uint4((uint32_t)key, (uint32_t)(key >> 32), (uint32_t)(b0 + j), 0u);
    }
  }
}

// tau_tile: the BN thresholds of this query tile (shared memory); qbase = first query of the tile
__device__ __forceinline__ void select_epilogue_tile(uint32_t taddr, int c_begin, int c_end, const float* tau_tile, float inv,
                                                     float dec, int64_t qbase, uint32_t row, const SelectSink& sink, int debug) {
#pragma unroll 1
  for (int c0 = c_begin; c0 < c_end; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(taddr + (uint32_t)c0, r);
    if (debug != 1) screen_chunk(r, taddr + (uint32_t)c0, tau_tile + c0, inv, dec, qbase + c0, row, sink);
  }
}

// per-row side inputs of the epilogue; a dropped row gets dec = NaN so that every comparison is false
__device__ __forceinline__ void row_inputs(const TcParamsBase& p, int64_t row, float& inv, float& dec) {
  bool kept = row < p.n && row >= p.f.lo && row < p.f.hi;
  if (kept && p.f.mask) kept = (p.f.mask[row >> 5] >> (row & 31)) & 1u;
  inv = 0.f;
  dec = __int_as_float(0x7fc00000);
  if (kept) {
    if (p.sqnorms) {                 // euclidean: screen -d^2 + |q|^2 = 2 v.q - |v|^2  (|q|^2 is folded into the threshold)
      inv = 2.f;
      dec = -p.sqnorms[row];
    } else {
      inv = p.inv_norms ? p.inv_norms[row] : 1.f;
      dec = p.f.decay ? (float)(p.f.bias * p.f.decay[row]) : 0.f;
    }
  }
}

struct TcParams : TcParamsBase {
  // DENSE mode
  float* dense;                 // [nq][n_tiles_m * 128] totals of the sample (dropped rows = -inf)
  // SELECT mode
  const float* tau0;            // [nq]
  uint64_t* cand;               // [nq][cap]
  unsigned* cand_count;         // [nq]
  int cap;
  uint4* rec;                   // [gridDim.x][rec_cap] CTA-private candidate records
  unsigned* rec_count;          // [gridDim.x]
  unsigned rec_cap;
  int debug;                    // HDB_TC_DEBUG: 1 = epilogue drains TMEM but skips the compare/append (timing experiments)
  int prefetch_dist;            // row tiles between the L2 bulk prefetch and its use (single-CTA form)
};

template <int BN, bool TF32, bool DENSE>
__global__ void __launch_bounds__(kTcThreads, 1)
batched_tc_kernel(const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_q, TcParams p) {
  constexpr int kABytes = kTileM * kTileKBytes;          // 16 KB
  constexpr int kBBytes = BN * kTileKBytes;              // up to 32 KB
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kKPerStage = kTileKBytes / (TF32 ? 4 : 2);     // elements of K per stage
  constexpr uint32_t kIdesc = make_idesc(kTileM, BN, TF32);
  constexpr int kTmemCols = 2 * BN;                      // two accumulator buffers (power of two >= 32)

  extern __shared__ __align__(1024) unsigned char tc_smem[];
  // the 128-byte swizzle atoms must start on 1024-byte boundaries
  unsigned char* stage_base = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(stage_base + kTcStages * kStageBytes);
  uint64_t* empty = full + kTcStages;
  uint64_t* acc_full = empty + kTcStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  unsigned* s_rec_count = tmem_slot + 1;
  float* s_tau = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(full) + 128);   // [n_tiles_q * BN], 16-byte aligned
  if (threadIdx.x == 0) *s_rec_count = 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = (int)((p.d + kKPerStage - 1) / kKPerStage);
  if (!DENSE) {      // every threshold of the batch lives in shared memory for the whole kernel (<= 16 KB)
    for (int64_t i = threadIdx.x; i < p.n_tiles_q * BN; i += kTcThreads)
      s_tau[i] = i < p.nq ? p.tau0[i] + (p.qsq ? p.qsq[i] : 0.f) : INFINITY;
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], kTcEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int64_t row_bytes = p.d * (TF32 ? 4 : 2);
      // Tile order: a CTA keeps one ROW tile for all query tiles before moving on, so the row tile comes from HBM
      // once and is re-read from L2 by the same CTA within microseconds (no reliance on CTAs staying in lockstep).
      for (int64_t sq = 0;; ++sq) {
        const int64_t mv = (int64_t)blockIdx.x + (sq / p.n_tiles_q) * gridDim.x;       // visited row-tile index
        if (mv >= p.n_tiles_m) break;
        const int64_t mt = mv * p.sample_stride;
        const int64_t qt = sq % p.n_tiles_q;
        if (qt == 0) {
          // The TMA boxes below gather 128-byte pieces of 128 rows (poor DRAM locality when they miss L2).  A row
          // tile itself is ONE contiguous region: stream the tile `prefetch_dist` visits ahead into L2 now with a
          // bulk prefetch (on the first visit also the ones in between).
          for (int ahead = (sq == 0 ? 1 : p.prefetch_dist); ahead <= p.prefetch_dist; ++ahead) {
            const int64_t mvn = mv + (int64_t)ahead * gridDim.x;
            if (mvn >= p.n_tiles_m) break;
            const int64_t mtn = mvn * p.sample_stride;
            const int64_t rows_n = (p.n - mtn * kTileM) < kTileM ? (p.n - mtn * kTileM) : kTileM;
            const int64_t region = (rows_n * row_bytes) & ~int64_t(15);
            if (region > 0)
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.rows + mtn * kTileM * row_bytes), "r"((uint32_t)region) : "memory");
          }
        }
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* sa = stage_base + stage * kStageBytes;
          mbar_expect_tx(&full[stage], kStageBytes);
          tma_load_2d(sa, &map_v, &full[stage], kb * kKPerStage, (int)(mt * kTileM));
          tma_load_2d(sa + kABytes, &map_q, &full[stage], kb * kKPerStage, (int)(qt * BN));
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int64_t sq = 0;; ++sq) {
        if ((int64_t)blockIdx.x + (sq / p.n_tiles_q) * gridDim.x >= p.n_tiles_m) break;
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_c = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * kStageBytes);
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kTileKBytes / 32; ++k)      // 32 bytes of K per MMA (16 halves / 8 tf32)
            umma<TF32>(tmem_c, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), kIdesc, (kb | k) ? 1u : 0u);
          tcgen05_commit(&empty[stage]);                   // frees the stage when these MMAs retire
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit(&acc_full[acc]);                    // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue: 8 warps; TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =====
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;                      // which quarter of the columns
    int acc = 0;
    uint32_t acc_phase = 0;
    const SelectSink sink{p.rec + (size_t)blockIdx.x * p.rec_cap, s_rec_count, p.rec_cap, p.qsq};
    float inv = 0.f, dec = 0.f;
    if ((int64_t)blockIdx.x < p.n_tiles_m)
      row_inputs(p, blockIdx.x * p.sample_stride * kTileM + quarter * 32 + lane, inv, dec);
    for (int64_t sq = 0;; ++sq) {
      const int64_t mt_visit = (int64_t)blockIdx.x + (sq / p.n_tiles_q) * gridDim.x;
      if (mt_visit >= p.n_tiles_m) break;
      const int64_t mt = mt_visit * p.sample_stride;
      const int64_t qt = sq % p.n_tiles_q;
      const int64_t row = mt * kTileM + quarter * 32 + lane;
      // side inputs of the NEXT row tile: their global-load latency hides behind this tile
      float ninv = inv, ndec = dec;
      if (qt == p.n_tiles_q - 1 && mt_visit + gridDim.x < p.n_tiles_m)
        row_inputs(p, (mt_visit + gridDim.x) * p.sample_stride * kTileM + quarter * 32 + lane, ninv, ndec);
      const int64_t valid = (p.nq - qt * BN) < BN ? (p.nq - qt * BN) : BN;       // real queries in this tile
      constexpr int kColsPerWarp = BN / 4 < 32 ? 32 : BN / 4;                     // 32-column chunks; BN = 64: two warps idle
      const int c_begin = part * kColsPerWarp;
      const int c_end = (int)(valid < (part + 1) * kColsPerWarp ? valid : (part + 1) * kColsPerWarp);
      mbar_wait(&acc_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
      if (DENSE) {
        const int64_t ld = p.n_tiles_m * kTileM;
        const int64_t col = mt_visit * kTileM + quarter * 32 + lane;
#pragma unroll 1
        for (int c0 = c_begin; c0 < c_end; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + (uint32_t)c0, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int64_t b = qt * BN + c0 + j;
            if (b < p.nq) {
              const float s = fmaf(__uint_as_float(r[j]), inv, dec) - (p.qsq ? p.qsq[b] : 0.f);     // euclidean: -d^2
              p.dense[b * ld + col] = (s == s) ? s : -INFINITY;
            }
          }
        }
      } else {
        select_epilogue_tile(taddr, c_begin, c_end, s_tau + qt * BN, inv, dec, qt * BN, (uint32_t)row, sink, p.debug);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      inv = ninv;
      dec = ndec;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (!DENSE && threadIdx.x == 0) p.rec_count[blockIdx.x] = *s_rec_count;      // may exceed rec_cap: overflow, seen by the bucket pass
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair form of the SELECT pass (cta_group::2): the two CTAs of a cluster contract one 256-row x 256-query
// tile.  Each CTA stages its own 128 rows of V and its own 128 queries (32 KB per stage instead of 48 KB, six
// stages instead of four, a third less L2 traffic per flop); the leader CTA issues tcgen05.mma.cta_group::2,
// the accumulator is split by rows over the two CTAs' TMEM and each CTA's epilogue warps drain their half.
// Barriers: every TMA of both CTAs completes on the LEADER's full barrier; tcgen05.commit multicasts the
// empty / accumulator-full arrivals to both CTAs; the peer's epilogue releases the accumulator remotely.
// ---------------------------------------------------------------------------------------------
constexpr int kTc2Stages = 6;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <bool TF32>
__device__ __forceinline__ void umma_pair(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if (TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
  }
}

template <bool TF32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
batched_tc_pair_kernel(const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_q, TcParams p) {
  constexpr int BN = 256;                                // queries per pair tile (128 staged by each CTA)
  constexpr int kABytes = kTileM * kTileKBytes;          // 16 KB: this CTA's 128 rows
  constexpr int kBBytes = (BN / 2) * kTileKBytes;        // 16 KB: this CTA's 128 queries
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kKPerStage = kTileKBytes / (TF32 ? 4 : 2);
  constexpr uint32_t kIdesc = make_idesc(2 * kTileM, BN, TF32);
  constexpr int kTmemCols = 2 * BN;

  extern __shared__ __align__(1024) unsigned char tc_smem[];
  unsigned char* stage_base = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(stage_base + kTc2Stages * kStageBytes);
  uint64_t* empty = full + kTc2Stages;
  uint64_t* acc_full = empty + kTc2Stages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  unsigned* s_rec_count = tmem_slot + 1;
  float* s_tau = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(full) + 256);      // [n_tiles_q * BN]
  if (threadIdx.x == 0) *s_rec_count = 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t i = threadIdx.x; i < p.n_tiles_q * BN; i += kTcThreads)
    s_tau[i] = i < p.nq ? p.tau0[i] + (p.qsq ? p.qsq[i] : 0.f) : INFINITY;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const int kblocks = (int)((p.d + kKPerStage - 1) / kKPerStage);
  const int64_t tiles_m2 = (p.n_tiles_m + 1) / 2;        // 256-row tiles
  const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTc2Stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 2 * kTcEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs; all transactions complete on the leader's full barrier) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int64_t row_bytes = p.d * (TF32 ? 4 : 2);
      for (int64_t sq = 0;; ++sq) {
        const int64_t m2 = pair + (sq / p.n_tiles_q) * npairs;    // 256-row tile of the pair: kept for all query tiles
        if (m2 >= tiles_m2) break;
        const int64_t mt = m2 * 2 + rank;                         // this CTA's 128-row tile
        const int64_t qt = sq % p.n_tiles_q;
        if (qt == 0 && m2 + npairs < tiles_m2 && p.debug != 2) {  // L2 bulk prefetch of this CTA's NEXT row tile
          const int64_t mtn = (m2 + npairs) * 2 + rank;
          const int64_t rows_n = (p.n - mtn * kTileM) < kTileM ? (p.n - mtn * kTileM) : kTileM;
          const int64_t region = rows_n > 0 ? (rows_n * row_bytes) & ~int64_t(15) : 0;
          if (region > 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.rows + mtn * kTileM * row_bytes), "r"((uint32_t)region) : "memory");
        }
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* sa = stage_base + stage * kStageBytes;
          const uint32_t leader_full = map_to_cta(smem_u32(&full[stage]), 0);
          if (leader) mbar_expect_tx(&full[stage], 2 * kStageBytes);
          tma_load_2d_pair(sa, &map_v, leader_full, kb * kKPerStage, (int)(mt * kTileM));
          tma_load_2d_pair(sa + kABytes, &map_q, leader_full, kb * kKPerStage, (int)(qt * BN + rank * (BN / 2)));
          if (++stage == kTc2Stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (leader && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int64_t sq = 0;; ++sq) {
        if (pair + (sq / p.n_tiles_q) * npairs >= tiles_m2) break;
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_c = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * kStageBytes);
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kTileKBytes / 32; ++k)
            umma_pair<TF32>(tmem_c, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), kIdesc, (kb | k) ? 1u : 0u);
          tcgen05_commit_pair(&empty[stage]);
          if (++stage == kTc2Stages) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit_pair(&acc_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue: this CTA's 128 rows x 256 queries =====
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    const SelectSink sink{p.rec + (size_t)blockIdx.x * p.rec_cap, s_rec_count, p.rec_cap, p.qsq};
    const uint32_t leader_acc_empty0 = map_to_cta(smem_u32(&acc_empty[0]), 0);
    float inv = 0.f, dec = 0.f;
    if (pair < tiles_m2) row_inputs(p, (pair * 2 + rank) * kTileM + quarter * 32 + lane, inv, dec);
    for (int64_t sq = 0;; ++sq) {
      const int64_t m2 = pair + (sq / p.n_tiles_q) * npairs;
      if (m2 >= tiles_m2) break;
      const int64_t mt = m2 * 2 + rank;
      const int64_t qt = sq % p.n_tiles_q;
      const int64_t row = mt * kTileM + quarter * 32 + lane;
      float ninv = inv, ndec = dec;
      if (qt == p.n_tiles_q - 1 && m2 + npairs < tiles_m2)
        row_inputs(p, ((m2 + npairs) * 2 + rank) * kTileM + quarter * 32 + lane, ninv, ndec);
      const int64_t valid = (p.nq - qt * BN) < BN ? (p.nq - qt * BN) : BN;
      const int c_begin = part * (BN / 4);
      const int c_end = (int)(valid < (part + 1) * (BN / 4) ? valid : (part + 1) * (BN / 4));
      mbar_wait(&acc_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
      select_epilogue_tile(taddr, c_begin, c_end, s_tau + qt * BN, inv, dec, qt * BN, (uint32_t)row, sink, p.debug);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(leader_acc_empty0 + (uint32_t)(acc * 8));       // the leader's MMA thread waits on it
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      inv = ninv;
      dec = ndec;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                                      // the peer may still read this CTA's smem / signal its barriers
  if (threadIdx.x == 0) p.rec_count[blockIdx.x] = *s_rec_count;
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-private records -> per-query candidate lists.  One CTA per source CTA; a source buffer that overflowed
// poisons every query by setting the TOP BIT of its counter with an atomic OR: the other CTAs' concurrent atomicAdds
// cannot clear it (a plain store of UINT_MAX could be wrapped back to a small count by a later add), so finalize
// sees count > capacity, reports the query uncertified and the host retries.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bucket_records_kernel(const uint4* rec, const unsigned* rec_count, unsigned rec_cap,
                                                             uint64_t* cand, unsigned* cand_count, int cap, int64_t nq) {
  const unsigned n = rec_count[blockIdx.x];
  if (n > rec_cap) {
    for (int64_t b = threadIdx.x; b < nq; b += 256) atomicOr(&cand_count[b], 0x80000000u);
    return;
  }
  const uint4* src = rec + (size_t)blockIdx.x * rec_cap;
  for (unsigned i = threadIdx.x; i < n; i += 256) {
    const uint4 r = src[i];
    const unsigned pos = atomicAdd(&cand_count[r.z], 1u);
    if (pos < (unsigned)cap) cand[(size_t)r.z * cap + pos] = ((uint64_t)r.y << 32) | r.x;
  }
}

// ---------------------------------------------------------------------------------------------
// tau0[b] = KP-th largest of the dense sample row of query b (8-bit radix select on the ordered bits);
// fewer than KP finite values -> -inf (everything passes).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_threshold_kernel(const float* dense, int64_t ld, int kp, float* tau0) {
  __shared__ unsigned hist[256];
  __shared__ unsigned sel[2];
  const float* v = dense + (int64_t)blockIdx.x * ld;
  uint32_t prefix = 0, pmask = 0;
  unsigned need = (unsigned)kp;
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t i = threadIdx.x; i < ld; i += 256) {
      const uint32_t o = order_f32(v[i]);
      if ((o & pmask) == prefix) atomicAdd(&hist[(o >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned cum = 0;
      int chosen = 0;
      for (int bin = 255; bin >= 0; --bin) {
        if (cum + hist[bin] >= need) { chosen = bin; break; }
        if (bin > 0) cum += hist[bin];
      }
      sel[0] = (unsigned)chosen;
      sel[1] = need - cum;
    }
    __syncthreads();
    prefix |= sel[0] << shift;
    pmask |= 0xffu << shift;
    need = sel[1];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float t = unorder_f32(prefix);
    // hist[] of the last pass counts the sample values EQUAL to the threshold.  A degenerate query (all zeros for dot / cosine, a
    // constant one for pearson) ties on every row: its select pass would append the whole matrix, overflow the CTA-private record
    // buffers and, through the poison of bucket_records_kernel, send every query of the batch to the repair path.  Give such a
    // query an unreachable threshold instead: nothing is appended, it alone comes back uncertified and is repaired.
    if (hist[sel[0]] > 4u * (unsigned)kp && t > -INFINITY) t = INFINITY;
    tau0[blockIdx.x] = t;
  }
}

__global__ void square_norms_kernel(const double* qnorm, float* qsq, int64_t nq) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < nq) qsq[i] = (float)(qnorm[i] * qnorm[i]);
}

// fp16 copy of the prepared (accumulate-type) queries: the tensor-core B operand
__global__ void queries_to_half_kernel(const float* qa, __half* q16, int64_t count) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    q16[i] = __float2half_rn(qa[i]);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int make_map(CUtensorMap* map, const void* base, bool tf32, int64_t rows, int64_t d, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail("cuTensorMapEncodeTiled is not available from the driver");
  const int esz = tf32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)d * esz};
  cuuint32_t box[2] = {(cuuint32_t)(kTileKBytes / esz), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
  return 0;
}

template <int BN, bool TF32, bool DENSE>
static int launch_tc(const CUtensorMap& mv, const CUtensorMap& mq, const TcParams& p, int grid, cudaStream_t s) {
  auto kern = batched_tc_kernel<BN, TF32, DENSE>;
  const size_t smem = (size_t)kTcStages * (kTileM + BN) * kTileKBytes + 128 + (DENSE ? 0 : (size_t)p.n_tiles_q * BN * 4) + 16 + 1024;
  HDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kTcThreads, smem, s>>>(mv, mq, p);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

int batched_tc_supported(const MatrixView& m, int metric, int q_dtype, int64_t nq, bool has_decay) {
  if (metric != HDB_DOT && metric != HDB_COSINE && metric != HDB_EUCLIDEAN && metric != HDB_PEARSON) return 0;
  if (metric == HDB_EUCLIDEAN && (has_decay || !m.sqnorms)) return 0;   // 1/(1+d) + decay is not monotone in -d^2
  // pearson: the contraction runs on the centred query b = q - mean(q) and the epilogue scales by 1/(std_v d); the query's own
  // 1/std_q is a positive factor per query, left to the certify step -- which a per-row decay term would not commute with
  if (metric == HDB_PEARSON && (has_decay || !m.pscale)) return 0;
  if (m.dtype == 2) return 0;                                  // no fp64 tensor path
  // The B operand has the storage precision: exact only for a query no wider than the store.  HDB_TC_MIXED=1 lets wider queries
  // (fp32 / fp64 queries over an fp16 store, fp64 over fp32) through as well: the operand is then the ROUNDED canonical query and the
  // certificate carries the rounding (certificate.cuh, `qcut`; proven against the oracle on the CPU, tests/test_emul_canonical.py).
  // Off by default: that configuration has not run on hardware yet.
  static const bool mixed = [] { const char* e = getenv("HDB_TC_MIXED"); return e && e[0] == '1'; }();
  if (q_dtype > m.dtype && !mixed) return 0;
  if ((m.d * dtype_size(m.dtype)) % 16 != 0 || (reinterpret_cast<uintptr_t>(m.rows) & 15)) return 0;
  if (m.n < 65536 || nq < 2) return 0;                         // one query, or a tiny shard (sample = n / 8 rows < 8192): the streaming sweep
  if (m.n >= (int64_t(1) << 31)) return 0;
  return 1;
}

template <bool TF32>
static int launch_tc_bn(int BN, bool dense, const CUtensorMap& mv, const CUtensorMap& mq, const TcParams& p, int grid, cudaStream_t s) {
  if (dense) {
    if (BN == 64) return launch_tc<64, TF32, true>(mv, mq, p, grid, s);
    if (BN == 128) return launch_tc<128, TF32, true>(mv, mq, p, grid, s);
    return launch_tc<256, TF32, true>(mv, mq, p, grid, s);
  }
  if (BN == 64) return launch_tc<64, TF32, false>(mv, mq, p, grid, s);
  if (BN == 128) return launch_tc<128, TF32, false>(mv, mq, p, grid, s);
  return launch_tc<256, TF32, false>(mv, mq, p, grid, s);
}

// Runs sample -> thresholds -> select for queries [0, nq) of the prepared batch.  Candidate keys land in
// ws.cand[b][0..min(count, cap)).
int launch_batched_tc(const MatrixView& m, int metric, const RowFilter& f, const float* qa, const double* qnorm, int64_t nq, int kp, int device,
                      const TcWorkspace& ws, cudaStream_t s) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const bool tf32 = (m.dtype == 1);      // fp32 storage: kind::tf32 reads the fp32 operands directly (10-bit mantissa)
  const void* qop = qa;
  if (!tf32) {
    const int64_t count = nq * m.d;
    queries_to_half_kernel<<<(unsigned)((count + 255) / 256 > 4096 ? 4096 : (count + 255) / 256), 256, 0, s>>>(qa, ws.q16, count);
    HDB_LAUNCHED();
    qop = ws.q16;
  }
  CUtensorMap mv, mq;
  HDB_TRY(make_map(&mv, m.rows, tf32, m.n, m.d, kTileM));
  const int BN = nq <= 64 ? 64 : (nq <= 128 ? 128 : 256);
  HDB_TRY(make_map(&mq, qop, tf32, nq, m.d, BN));
  TcParams p;
  p.n = m.n; p.d = m.d; p.nq = nq;
  p.n_tiles_q = (nq + BN - 1) / BN;
  p.rows = reinterpret_cast<const char*>(m.rows);
  p.debug = getenv("HDB_TC_DEBUG") ? atoi(getenv("HDB_TC_DEBUG")) : 0;
  p.prefetch_dist = getenv("HDB_TC_PREFETCH_DIST") ? atoi(getenv("HDB_TC_PREFETCH_DIST")) : 1;
  if (p.prefetch_dist < 1) p.prefetch_dist = 1;
  p.inv_norms = metric == HDB_COSINE ? reinterpret_cast<const float*>(m.inv_norms)
              : metric == HDB_PEARSON ? reinterpret_cast<const float*>(m.pscale) : nullptr;     // NaN for a constant row: never appended
  p.sqnorms = nullptr; p.qsq = nullptr;
  if (metric == HDB_EUCLIDEAN) {
    square_norms_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, s>>>(qnorm, ws.qsq, nq);
    HDB_LAUNCHED();
    p.sqnorms = m.sqnorms; p.qsq = ws.qsq;
  }
  p.f = f;
  const int64_t all_tiles = (m.n + kTileM - 1) / kTileM;
  // ---- dense sample
  p.n_tiles_m = ws.sample_tiles;
  p.sample_stride = all_tiles / ws.sample_tiles;
  p.dense = ws.dense; p.tau0 = nullptr; p.cand = nullptr; p.cand_count = nullptr; p.cap = 0;
  p.rec = nullptr; p.rec_count = nullptr; p.rec_cap = 0;
  HDB_TRY(tf32 ? launch_tc_bn<true>(BN, true, mv, mq, p, sms, s) : launch_tc_bn<false>(BN, true, mv, mq, p, sms, s));
  sample_threshold_kernel<<<(unsigned)nq, 256, 0, s>>>(ws.dense, ws.sample_tiles * kTileM, kp, ws.tau0);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  // ---- select over all rows
  HDB_CUDA(cudaMemsetAsync(ws.cand_count, 0, (size_t)nq * 4, s));
  p.n_tiles_m = all_tiles;
  p.sample_stride = 1;
  p.dense = nullptr; p.tau0 = ws.tau0; p.cand = ws.cand; p.cand_count = ws.cand_count; p.cap = ws.cap;
  p.rec = reinterpret_cast<uint4*>(ws.rec); p.rec_count = ws.rec_count; p.rec_cap = ws.rec_cap;
  auto bucket = [&]() -> int {
    bucket_records_kernel<<<sms, 256, 0, s>>>(reinterpret_cast<const uint4*>(ws.rec), ws.rec_count, ws.rec_cap, ws.cand, ws.cand_count,
                                               ws.cap, nq);
    HDB_LAUNCHED();
    HDB_CUDA(cudaGetLastError());
    return 0;
  };
  if (BN == 256 && (sms % 2) == 0 && !ws.force_single) {
    // CTA-pair contraction: each CTA's query box is 128 rows
    CUtensorMap mq2;
    HDB_TRY(make_map(&mq2, qop, tf32, nq, m.d, 128));
    const size_t smem = (size_t)kTc2Stages * (kTileM + 128) * kTileKBytes + 256 + (size_t)p.n_tiles_q * 256 * 4 + 16 + 1024;
    if (tf32) {
      HDB_CUDA(cudaFuncSetAttribute(batched_tc_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      batched_tc_pair_kernel<true><<<sms, kTcThreads, smem, s>>>(mv, mq2, p);
    } else {
      HDB_CUDA(cudaFuncSetAttribute(batched_tc_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      batched_tc_pair_kernel<false><<<sms, kTcThreads, smem, s>>>(mv, mq2, p);
    }
    HDB_LAUNCHED();
    HDB_CUDA(cudaGetLastError());
    return bucket();
  }
  HDB_TRY(tf32 ? launch_tc_bn<true>(BN, false, mv, mq, p, sms, s) : launch_tc_bn<false>(BN, false, mv, mq, p, sms, s));
  return bucket();
}

}  // namespace hdb
