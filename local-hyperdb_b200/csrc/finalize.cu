// Second half of the hot path: merge the per-CTA candidate keys of the sweep, re-score the KP
// candidates in the reference's exact arithmetic (canonical.cuh), order them by
// (float64 score desc, row asc) as hyperdb/ranking_algorithm.py:194-204 would (ties -> lower index,
// north_star), and CERTIFY that no row outside the candidate list can belong to the top-k.
// Plus: the exact full-vector path (every row canonical -> stable radix sort) that serves
//   * the metric functions themselves (hdb_scores; ranking_algorithm.py:24-61,:128-147),
//   * queries the certificate rejects, top_k beyond the fused classes, and tests,
// and the multi-GPU final merge of all-gathered candidates.
#include <cub/device/device_radix_sort.cuh>

#include "canonical.cuh"
#include "certificate.cuh"
#include "hdb_exchange.cuh"
#include "hdb_internal.h"
#include "../../include/hyperdb_b200.h"

namespace hdb {

constexpr int kSurvCap = 2048;      // candidate keys >= the final threshold that one CTA can merge
constexpr int kFinThreads = 512;
constexpr uint32_t kFlagUncertified = 8u;

// row order (RowFilter::ord / inv): a key carries the caller's row index; the data live at the physical row
__device__ __forceinline__ uint32_t phys_row(const RowFilter& f, uint32_t idx) { return f.inv ? f.inv[idx] : idx; }

template <typename CS> __device__ __forceinline__ CS to_carrier(__half x) { return (CS)__half2float(x); }
template <typename CS> __device__ __forceinline__ CS to_carrier(float x) { return (CS)x; }
template <typename CS> __device__ __forceinline__ CS to_carrier(double x) { return (CS)x; }

// 16 strided elements of one row, all loads issued before any use (hides the global-load latency)
template <typename T>
__device__ __forceinline__ void load_batch(const void* src, int64_t j0, int64_t d, double (&v)[16]) {
  const T* p = reinterpret_cast<const T*>(src);
  T raw[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int64_t j = j0 + 32 * u;
    raw[u] = p[j < d ? j : j0];
  }
#pragma unroll
  for (int u = 0; u < 16; ++u) v[u] = (double)(float)raw[u];
}
template <>
__device__ __forceinline__ void load_batch<double>(const void* src, int64_t j0, int64_t d, double (&v)[16]) {
  const double* p = reinterpret_cast<const double*>(src);
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int64_t j = j0 + 32 * u;
    v[u] = p[j < d ? j : j0];
  }
}

// Typed fast path of the canonical re-scoring: T = storage type, SDT/RDT = storage / result dtype ids
// (SDT <= RDT).  Same arithmetic as canonical_term / canonical_reduce, without runtime dtype switches or
// double round-trips; v / norm is skipped when the norm is exactly 1 (x / 1 == x in IEEE arithmetic).
template <typename T, int SDT, int RDT>
__device__ void rescore_smem(const FinalizeArgs& a, int64_t b, const uint64_t* surv, int m, double* c_tot,
                             uint32_t* c_row, unsigned char* smem, PwScratch* s_pw) {
  using AR = Arith<RDT>;
  using AS = Arith<SDT>;
  using C = typename AR::C;
  using CS = typename AS::C;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d = (int)a.m.d;
  const int pitch = d | 1;                                                // odd: conflict-free column walks
  C* s_q = reinterpret_cast<C*>(smem);
  C* s_terms = reinterpret_cast<C*>(smem + (((size_t)d * sizeof(C) + 15) & ~size_t(15)));
  const int64_t avail = (int64_t)a.smem_bytes - (reinterpret_cast<unsigned char*>(s_terms) - smem);
  int batch = (int)(avail / ((int64_t)pitch * sizeof(C)));
  if (batch > m) batch = m;
  const double* qc = a.qb.qc + b * a.m.d;
  for (int j = tid; j < d; j += kFinThreads) s_q[j] = AR::from_double(qc[j]);
  const T* grows = reinterpret_cast<const T*>(a.m.rows);
  const int metric = a.metric;
  for (int base = 0; base < m; base += batch) {
    const int nb = (m - base) < batch ? (m - base) : batch;
    __syncthreads();
    // ---- phase A
    for (int r = warp; r < nb; r += kFinThreads / 32) {
      const uint32_t row = phys_row(a.f, key_row(surv[base + r]));
      const T* src = grows + (int64_t)row * d;
      CS nrm = CS(1);
      if (metric == HDB_COSINE) nrm = SDT == 2 ? (CS) reinterpret_cast<const double*>(a.m.norms)[row]
                                                : (CS) reinterpret_cast<const float*>(a.m.norms)[row];
      if (metric == HDB_PEARSON) nrm = SDT == 2 ? (CS) reinterpret_cast<const double*>(a.m.pmean)[row]
                                                 : (CS) reinterpret_cast<const float*>(a.m.pmean)[row];
      C* dst = s_terms + (size_t)r * pitch;
      for (int j0 = lane; j0 < d; j0 += 32 * 8) {
        T raw[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int j = j0 + 32 * u; raw[u] = src[j < d ? j : j0]; }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + 32 * u;
          if (j < d) {
            const CS vs = to_carrier<CS>(raw[u]);
            C term;
            if (metric == 0) term = (C)vs;
            else if (metric == 1) term = (nrm == CS(1)) ? (C)vs : (C)AS::div(vs, nrm);
            else if (metric == HDB_PEARSON) term = AR::mul((C)AS::sub(vs, nrm), s_q[j]);     // (v - mean) in S, product in R
            else {
              const C df = AR::sub((C)vs, s_q[j]);
              term = metric == 2 ? AR::mul(df, df) : C(fabs(df));
            }
            dst[j] = term;
          }
        }
      }
    }
    __syncthreads();
    // ---- phase B
    if (metric <= 1) {
      const int t = tid - base;
      if (t >= 0 && t < nb) {
        const uint32_t idx = key_row(surv[tid]);
        const uint32_t row = phys_row(a.f, idx);
        const C* terms = s_terms + (size_t)t * pitch;
        const double sim = canonical_reduce<RDT>(metric, [&](int j) { return terms[j]; }, [&](int j) { return s_q[j]; }, d);
        c_tot[tid] = total_score(sim, a.f.decay, a.f.bias, row);
        c_row[tid] = idx;
      }
    } else {
      for (int r = warp; r < nb; r += kFinThreads / 32) {
        const uint32_t idx = key_row(surv[base + r]);
        const uint32_t row = phys_row(a.f, idx);
        const C* terms = s_terms + (size_t)r * pitch;
        C dist = pairwise_sum_warp<RDT>([&](int j) { return terms[j]; }, d, lane, &s_pw[warp]);
        double sim;
        if (metric == HDB_PEARSON) {
          const double sd = SDT == 2 ? reinterpret_cast<const double*>(a.m.pstd)[row] : (double)reinterpret_cast<const float*>(a.m.pstd)[row];
          sim = pearson_quotient<RDT>(dist, sd, a.qb.qaux[2 * b], d);
        } else {
          if (metric == 2) dist = AR::sqrt(dist);
          sim = (double)AR::div(C(1), AR::add(C(1), dist));
        }
        if (lane == 0) {
          c_tot[base + r] = total_score(sim, a.f.decay, a.f.bias, row);
          c_row[base + r] = idx;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kFinThreads) finalize_kernel(FinalizeArgs a) {
  __shared__ uint64_t surv[kSurvCap];
  __shared__ double c_tot[kMaxKP];
  __shared__ uint32_t c_row[kMaxKP];
  __shared__ double o_tot[kMaxKP];
  __shared__ uint32_t o_row[kMaxKP];
  __shared__ int s_count;
  __shared__ unsigned s_hist[256];
  __shared__ PwScratch s_pw[kFinThreads / 32];
  __shared__ int s_sel[3];
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  const int kk = (int)((int64_t)a.k < a.n_kept ? (int64_t)a.k : a.n_kept);
  const uint32_t qflag = a.qb.qflags[b];
  if (tid == 0) s_count = 0;
  __syncthreads();
  const unsigned long long tau = a.cand_count ? 0ull : a.tau[b];
  const uint64_t* cand = a.cand_count ? a.cand + b * a.cand_stride : a.cand + b * (int64_t)a.grid * a.kp;
  bool overflow = false;
  int total = a.grid * a.kp;
  if (a.cand_count) {
    const unsigned appended = a.cand_count[b];
    overflow = appended > (unsigned)a.cand_stride;          // some qualifying rows were dropped: cannot certify
    total = (int)(overflow ? (unsigned)a.cand_stride : appended);
  }
  // Tighten the threshold before gathering: every per-CTA list is sorted, so the KP-th largest of the lists' HEADS
  // (KP distinct keys) is a lower bound of the KP-th key overall -- far better than the sweep's running threshold
  // when the shard is small and each warp saw only a few hundred rows.  Keeps the gather below on its fast path.
  __shared__ uint64_t s_heads[1024];
  __shared__ unsigned long long s_tau_star;
  if (tid == 0) s_tau_star = 0;
  if (!a.cand_count && a.grid >= a.kp && a.grid <= 1024) {
    for (int i = tid; i < a.grid; i += kFinThreads) s_heads[i] = cand[(int64_t)i * a.kp];
    __syncthreads();
    for (int i = tid; i < a.grid; i += kFinThreads) {
      const uint64_t mine = s_heads[i];
      int rank = 0;
      for (int u = 0; u < a.grid; ++u) rank += s_heads[u] > mine;
      if (rank == a.kp - 1 && mine != 0) s_tau_star = mine;          // keys are unique: exactly one thread matches
    }
  }
  __syncthreads();
  const unsigned long long tau_eff = tau > s_tau_star ? tau : (unsigned long long)s_tau_star;
  for (int i0 = tid; i0 < total; i0 += kFinThreads * 8) {       // 8 loads in flight per thread
    uint64_t key[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * kFinThreads;
      key[u] = i < total ? cand[i] : 0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (key[u] != 0 && key[u] >= tau_eff) {
        int pos = atomicAdd(&s_count, 1);
        if (pos < kSurvCap) surv[pos] = key[u];
      }
    }
  }
  __syncthreads();
  int found = s_count;
  if (found > kSurvCap) {
    // No useful threshold was established (small shard: every warp saw fewer than KP rows).  Radix-select
    // the KP-th largest key over all per-CTA candidates, 8 bits per pass, until the survivors fit.
    unsigned long long prefix = 0, pmask = 0;
    int need = a.kp;
    for (int shift = 56; shift >= 0; shift -= 8) {
      for (int i = tid; i < 256; i += kFinThreads) s_hist[i] = 0;
      __syncthreads();
      for (int i = tid; i < total; i += kFinThreads) {
        const uint64_t key = cand[i];
        if (key != 0 && (key & pmask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        int cum = 0, chosen = 0;
        for (int bin = 255; bin >= 0; --bin) {
          const int c = (int)s_hist[bin];
          if (cum + c >= need) { chosen = bin; break; }
          if (bin > 0) cum += c;
        }
        need -= cum;
        s_sel[0] = chosen;
        s_sel[1] = need;
        s_sel[2] = ((a.kp - need) + (int)s_hist[chosen] <= kSurvCap) ? 1 : 0;
      }
      __syncthreads();
      prefix |= (unsigned long long)s_sel[0] << shift;
      pmask |= 0xffull << shift;
      need = s_sel[1];
      const int done = s_sel[2];
      __syncthreads();
      if (done) break;
    }
    if (tid == 0) s_count = 0;
    __syncthreads();
    for (int i = tid; i < total; i += kFinThreads) {
      const uint64_t key = cand[i];
      if (key != 0 && key >= prefix) {
        const int pos = atomicAdd(&s_count, 1);
        if (pos < kSurvCap) surv[pos] = key;
      }
    }
    __syncthreads();
    found = s_count < kSurvCap ? s_count : kSurvCap;
  }
  // order the survivors: rank by counting when they are few (one barrier), bitonic sort otherwise
  if (found <= kFinThreads) {
    uint64_t mine = 0;
    int rank = 0;
    if (tid < found) {
      mine = surv[tid];
      for (int j = 0; j < found; ++j) rank += surv[j] > mine;
    }
    __syncthreads();
    if (tid < found) surv[rank] = mine;
    __syncthreads();
  } else {
    int pow2 = 512;
    while (pow2 < found) pow2 <<= 1;
    for (int i = found + tid; i < pow2; i += kFinThreads) surv[i] = 0;
    __syncthreads();
    bitonic_desc(surv, pow2, tid, kFinThreads, [] { __syncthreads(); });
  }
  const int m = found < a.kp ? found : a.kp;          // candidates to certify

  // Canonical re-scoring in two phases: (A) all threads compute the per-column terms of the candidate rows
  // (coalesced global reads, element-wise reference arithmetic) into shared memory; (B) the reference's
  // order-dependent reduction out of shared memory (thread per candidate for the dot chains, warp per
  // candidate for NumPy's pairwise sums).
  extern __shared__ __align__(16) unsigned char fin_smem[];
  if (a.smem_bytes > 0) {
    const int combo = a.m.dtype * 3 + a.rdt;
    switch (combo) {
      case 0: rescore_smem<__half, 0, 0>(a, b, surv, m, c_tot, c_row, fin_smem, s_pw); break;
      case 1: rescore_smem<__half, 0, 1>(a, b, surv, m, c_tot, c_row, fin_smem, s_pw); break;
      case 2: rescore_smem<__half, 0, 2>(a, b, surv, m, c_tot, c_row, fin_smem, s_pw); break;
      case 4: rescore_smem<float, 1, 1>(a, b, surv, m, c_tot, c_row, fin_smem, s_pw); break;
      case 5: rescore_smem<float, 1, 2>(a, b, surv, m, c_tot, c_row, fin_smem, s_pw); break;
      default: rescore_smem<double, 2, 2>(a, b, surv, m, c_tot, c_row, fin_smem, s_pw); break;
    }
  } else if (tid < m) {
    // hamming, or rows too long for shared memory: straight from global memory
    CanonArgs ca;
    ca.sdt = a.m.dtype; ca.d = a.m.d; ca.words = a.m.words;
    ca.qbits = a.qb.qbits ? a.qb.qbits + b * a.m.words : nullptr;
    ca.metric = a.metric;
    ca.qc = a.qb.qc + b * a.m.d;
    ca.qstd = a.qb.qaux ? a.qb.qaux[2 * b] : 1.0;
    ca.distance = 0;
    const uint32_t idx = key_row(surv[tid]);
    const uint32_t row = phys_row(a.f, idx);
    const uint32_t* bitrow = a.m.bits ? a.m.bits + (int64_t)row * a.m.words : nullptr;
    double nrm = 1.0, aux2 = 0.0;
    if (a.metric == HDB_COSINE)
      nrm = a.m.dtype == 2 ? reinterpret_cast<const double*>(a.m.norms)[row] : (double)reinterpret_cast<const float*>(a.m.norms)[row];
    if (a.metric == HDB_PEARSON) {
      nrm = a.m.dtype == 2 ? reinterpret_cast<const double*>(a.m.pmean)[row] : (double)reinterpret_cast<const float*>(a.m.pmean)[row];
      aux2 = a.m.dtype == 2 ? reinterpret_cast<const double*>(a.m.pstd)[row] : (double)reinterpret_cast<const float*>(a.m.pstd)[row];
    }
    const char* rowp = reinterpret_cast<const char*>(a.m.rows) + (int64_t)row * a.m.d * dtype_size(a.m.dtype);
    double sim;
    if (a.metric == HDB_HAMMING || a.metric == HDB_JACCARD) {
      // packed rows are whole 16-byte vectors: independent 128-bit loads instead of a word-by-word chain
      const uint4* vr = reinterpret_cast<const uint4*>(bitrow);
      const uint4* vq = reinterpret_cast<const uint4*>(ca.qbits);
      int diff = 0, inter = 0, uni = 0;
#pragma unroll 4
      for (int v = 0; v < a.m.words / 4; ++v) {
        const uint4 x = vr[v], q = vq[v];
        diff += __popc(x.x ^ q.x) + __popc(x.y ^ q.y) + __popc(x.z ^ q.z) + __popc(x.w ^ q.w);
        inter += __popc(x.x & q.x) + __popc(x.y & q.y) + __popc(x.z & q.z) + __popc(x.w & q.w);
        uni += __popc(x.x | q.x) + __popc(x.y | q.y) + __popc(x.z | q.z) + __popc(x.w | q.w);
      }
      sim = a.metric == HDB_HAMMING ? (double)((int)a.m.d - diff) : __ddiv_rn((double)inter, (double)uni);
    } else {
      sim = canonical_similarity_rt(ca, a.rdt, rowp, bitrow, nrm, aux2);
    }
    c_tot[tid] = total_score(sim, a.f.decay, a.f.bias, row);
    c_row[tid] = idx;
  }
  __syncthreads();
  if (tid < m) {
    const double t = c_tot[tid];
    const uint32_t r = c_row[tid];
    int rank = 0;
    for (int j = 0; j < m; ++j) {
      const double tj = c_tot[j];
      rank += (tj > t) || (tj == t && c_row[j] < r);
    }
    o_tot[rank] = t;
    o_row[rank] = r;
  }
  __syncthreads();
  bool certified = true;
  const bool exact_keys = (a.metric == HDB_HAMMING) && (a.f.decay == nullptr);
  if (kk > 0 && a.n_kept > (int64_t)m && !exact_keys) {
    if (m < kk) certified = false;
    else {
      // rows outside the candidates: key <= the KP-th key, or (batched pass, fewer than KP appended) below tau0
      double s_edge = (double)key_score(surv[m - 1]);
      if (m < a.kp && a.tau0) {
        const double t0 = (double)a.tau0[b];
        s_edge = a.tau0_negd2 ? 1.0 / (1.0 + sqrt(t0 < 0.0 ? -t0 : 0.0)) : t0;
      }
      const double qstd = a.qb.qaux ? a.qb.qaux[2 * b] : 1.0;
      double qsumb = 0.0;
      if (a.cand_count && a.metric == HDB_PEARSON) {
        // the batched pass screens v.b / (std_v d): its keys and tau0 still lack the query's 1 / std_q (> 0: same order per query)
        s_edge = s_edge / qstd;
        qsumb = a.qb.qaux[2 * b + 1];
      }
      const double bound = outsider_bound(s_edge, a, a.qb.qnorm[b], qstd, qsumb);
      certified = o_tot[kk - 1] > bound;
    }
  }
  if (m < kk || overflow) certified = false;
  const uint32_t flags_out = qflag | a.extra_flags | (certified ? 0u : kFlagUncertified);
  for (int i = tid; i < a.k; i += kFinThreads) {
    const bool have = i < kk && i < m;
    a.out_idx[b * a.k + i] = have ? (int64_t)o_row[i] + a.m.row_offset : -1;
    a.out_score[b * a.k + i] = have ? o_tot[i] : -INFINITY;
  }
  if (tid == 0) {
    a.out_count[b] = kk;
    if (a.out_flags) a.out_flags[b] = flags_out;
    if (!certified) atomicAdd(a.uncertified, 1);
  }
  // ---- fused push (row-sharded path): this query's records go straight into EVERY rank's exchange slot over NVLink
  //      peer memory; the last CTA of the launch publishes the arrival flags.  No kernel boundary, no packed copy.
  if (a.push.peer) {
    __shared__ unsigned long long s_push_step;
    if (tid == 0) s_push_step = a.push.ctr[kCtrPushStep];
    __syncthreads();
    const unsigned long long step = s_push_step;
    if (tid < a.push.world) push_wait_consumed(a.push, tid, step);
    __syncthreads();
    const int64_t nq = a.push.nq, k = a.k;
    for (int g = 0; g < a.push.world; ++g) {
      unsigned long long* dst = push_slot(a.push, g, step);
      for (int i = tid; i < k; i += kFinThreads) {
        const bool have = i < kk && i < m;
        reinterpret_cast<double*>(dst)[b * k + i] = have ? o_tot[i] : -INFINITY;
        reinterpret_cast<int64_t*>(dst)[nq * k + b * k + i] = have ? (int64_t)o_row[i] + a.m.row_offset : -1;
      }
      if (tid == 0) {
        reinterpret_cast<int64_t*>(dst)[2 * nq * k + b] = kk;
        reinterpret_cast<uint32_t*>(dst + 2 * nq * k + nq)[b] = flags_out;
      }
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) push_publish(a.push, step, gridDim.x);
  }
}

int launch_finalize(const FinalizeArgs& a_in, int64_t nq, cudaStream_t s) {
  if (nq == 0) return 0;
  FinalizeArgs a = a_in;
  // query (8 B/dim) + the terms of as many candidate rows as fit in ~160 KB; 0 = read global memory directly
  const int64_t row_pitch = (a.m.d | 1) * (a.rdt == 2 ? 8 : 4);
  const int64_t qbytes = (a.m.d * (a.rdt == 2 ? 8 : 4) + 15) & ~int64_t(15);
  int64_t want = qbytes + row_pitch * a.kp;
  if (want > 160 * 1024) want = 160 * 1024;
  if (want < qbytes + row_pitch || a.metric == HDB_HAMMING || a.metric == HDB_JACCARD) want = 0;
  a.smem_bytes = (int)want;
  // the attribute belongs to a device's context (one process may hold shards on several) and handles may live on
  // different host threads: a per-device atomic flag; setting it twice is harmless
  static std::atomic<bool> attr_set[64];
  int dev = 0;
  HDB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
    HDB_CUDA(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    if (dev >= 0 && dev < 64) attr_set[dev].store(true, std::memory_order_release);
  }
  finalize_kernel<<<(unsigned)nq, kFinThreads, (size_t)a.smem_bytes, s>>>(a);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// exact full-vector path
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool kept_row(const RowFilter& f, int64_t row) {
  if (row < f.lo || row >= f.hi) return false;
  if (f.mask && !((f.mask[row >> 5] >> (row & 31)) & 1u)) return false;
  return true;
}

__device__ __forceinline__ CanonArgs canon_args(const MatrixView& m, int metric, const double* qc, const uint32_t* qbits,
                                                const double* qaux) {
  CanonArgs ca;
  ca.sdt = m.dtype; ca.d = m.d; ca.qc = qc;
  ca.qbits = qbits; ca.words = m.words; ca.metric = metric;
  ca.qstd = qaux ? qaux[0] : 1.0;
  ca.distance = 0;
  return ca;
}

__device__ __forceinline__ double row_canonical(const MatrixView& m, CanonArgs& ca, int rdt, int64_t row) {
  double nrm = 1.0;
  if (ca.metric == HDB_COSINE)
    nrm = m.dtype == 2 ? reinterpret_cast<const double*>(m.norms)[row] : (double)reinterpret_cast<const float*>(m.norms)[row];
  double aux2 = 0.0;
  if (ca.metric == HDB_PEARSON) {
    nrm = m.dtype == 2 ? reinterpret_cast<const double*>(m.pmean)[row] : (double)reinterpret_cast<const float*>(m.pmean)[row];
    aux2 = m.dtype == 2 ? reinterpret_cast<const double*>(m.pstd)[row] : (double)reinterpret_cast<const float*>(m.pstd)[row];
  }
  const void* rowp = reinterpret_cast<const char*>(m.rows) + row * m.d * dtype_size(m.dtype);
  const uint32_t* bitrow = m.bits ? m.bits + row * (int64_t)m.words : nullptr;
  return canonical_similarity_rt(ca, rdt, rowp, bitrow, nrm, aux2);
}

__global__ void full_scores_kernel(MatrixView m, RowFilter f, int metric, int rdt, const double* qc,
                                   const uint32_t* qbits, const double* qaux, double* totals) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= m.n) return;
  if (!kept_row(f, row)) { totals[row] = __longlong_as_double(-1ll); return; }   // -NaN sorts last (descending)
  CanonArgs ca = canon_args(m, metric, qc, qbits, qaux);
  totals[row] = total_score(row_canonical(m, ca, rdt, row), f.decay, f.bias, row);
}

int launch_full_scores(const MatrixView& m, const RowFilter& f, int metric, int rdt, const double* qc,
                       const uint32_t* qbits, const double* qaux, double* totals, cudaStream_t s) {
  if (m.n == 0) return 0;
  int handled = 0;
  HDB_TRY(launch_scores_rowwise(m, f, metric, rdt, qc, qaux, totals, nullptr, 0, s, &handled));     // csrc/rowwise.cu: HBM-rate forms
  if (handled) return 0;
  int64_t blocks = (m.n + 127) / 128;
  full_scores_kernel<<<(unsigned)blocks, 128, 0, s>>>(m, f, metric, rdt, qc, qbits, qaux, totals);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

// the metric function's own output: dtype R (uint64 for hamming)
__global__ void scores_out_kernel(MatrixView m, int metric, int rdt, const double* qc, const uint32_t* qbits, const double* qaux, void* out,
                                  int distance) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= m.n) return;
  CanonArgs ca = canon_args(m, metric, qc, qbits, qaux);
  ca.distance = distance;
  const double v = row_canonical(m, ca, rdt, row);
  if (metric == HDB_HAMMING) reinterpret_cast<unsigned long long*>(out)[row] = (unsigned long long)(long long)v;
  else if (metric == HDB_JACCARD || metric == HDB_PEARSON) reinterpret_cast<double*>(out)[row] = v;      // np.zeros(N) receives the quotients
  else if (rdt == 0) reinterpret_cast<__half*>(out)[row] = __float2half_rn((float)v);
  else if (rdt == 1) reinterpret_cast<float*>(out)[row] = (float)v;
  else reinterpret_cast<double*>(out)[row] = v;
}

int launch_scores_out(const MatrixView& m, int metric, int rdt, const double* qc, const uint32_t* qbits, const double* qaux, void* out,
                      int distance, cudaStream_t s) {
  if (m.n == 0) return 0;
  if (metric != HDB_EUCLIDEAN) distance = 0;
  int handled = 0;
  RowFilter none;
  none.mask = nullptr; none.lo = 0; none.hi = m.n; none.decay = nullptr; none.bias = 0.0;
  HDB_TRY(launch_scores_rowwise(m, none, metric, rdt, qc, qaux, nullptr, out, distance, s, &handled));
  if (handled) return 0;
  int64_t blocks = (m.n + 127) / 128;
  scores_out_kernel<<<(unsigned)blocks, 128, 0, s>>>(m, metric, rdt, qc, qbits, qaux, out, distance);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

__global__ void iota_kernel(int64_t* v, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = i;
}
__global__ void take_topk_kernel(const double* keys, const int64_t* vals, int64_t k, int64_t kk, int64_t row_offset,
                                 int64_t* out_idx, double* out_score, int64_t* out_count) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < k; i += (int64_t)gridDim.x * blockDim.x) {
    out_idx[i] = i < kk ? vals[i] + row_offset : -1;
    out_score[i] = i < kk ? keys[i] : -INFINITY;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *out_count = kk;
}

// Stable descending radix sort of (total, row): equal totals keep ascending row order.
__global__ void gather_totals_kernel(const double* totals, const uint32_t* inv, int64_t n, double* out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = totals[inv[i]];
}

// inv (row order set): the totals arrive in PHYSICAL row order and are first brought into the caller's order, so that the
// stable sort resolves ties on the caller's row index and the sorted values ARE the reported ids.
int exact_topk(int device, const double* totals, const uint32_t* inv, int64_t n, int64_t row_offset, int64_t k, int64_t n_kept,
               int64_t* out_idx, double* out_score, int64_t* out_count, void** scratch, size_t* scratch_bytes,
               cudaStream_t s) {
  (void)device;
  if (n > 0x7fffffff) return fail("exact path: more than 2^31 rows per shard");
  const int64_t kk = k < n_kept ? k : n_kept;
  size_t temp = 0;
  HDB_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, temp, (const double*)nullptr, (double*)nullptr,
                                                     (const int64_t*)nullptr, (int64_t*)nullptr, (int)n, 0, 64, s));
  const size_t arr = ((size_t)n * 8 + 255) & ~size_t(255);
  const size_t need = (inv ? 4 : 3) * arr + temp + 256;
  if (*scratch_bytes < need) {
    if (*scratch) HDB_CUDA(cudaFree(*scratch));
    *scratch = nullptr; *scratch_bytes = 0;
    HDB_CUDA(cudaMalloc(scratch, need));
    *scratch_bytes = need;
  }
  char* base = reinterpret_cast<char*>(*scratch);
  double* keys_out = reinterpret_cast<double*>(base);
  int64_t* vals_in = reinterpret_cast<int64_t*>(base + arr);
  int64_t* vals_out = reinterpret_cast<int64_t*>(base + 2 * arr);
  void* tmp = base + (inv ? 4 : 3) * arr;
  if (n > 0) {
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (inv) {
      double* ordered = reinterpret_cast<double*>(base + 3 * arr);
      gather_totals_kernel<<<(unsigned)blocks, 256, 0, s>>>(totals, inv, n, ordered);
      HDB_LAUNCHED();
      totals = ordered;
    }
    iota_kernel<<<(unsigned)blocks, 256, 0, s>>>(vals_in, n);
    HDB_LAUNCHED();
    HDB_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp, temp, totals, keys_out, vals_in, vals_out, (int)n, 0, 64, s));
    HDB_LAUNCHED();
  }
  if (k > 0) {
    int64_t blocks = (k + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    take_topk_kernel<<<(unsigned)blocks, 256, 0, s>>>(keys_out, vals_out, k, kk, row_offset, out_idx, out_score, out_count);
  } else {
    take_topk_kernel<<<1, 32, 0, s>>>(keys_out, vals_out, 0, 0, row_offset, out_idx, out_score, out_count);
  }
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU final merge: n_lists x k records per query -> top-k by (score desc, id asc).
// One CTA per query; rank by counting (n_lists*k is at most a few thousand records).
// ---------------------------------------------------------------------------------------------
__global__ void merge_topk_kernel(int64_t n_lists, int64_t nq, int64_t k, int64_t ls_rec, int64_t ls_cnt, const double* scores, const int64_t* ids,
                                  const int64_t* counts, int64_t* out_idx, double* out_score, int64_t* out_count) {
  const int64_t b = blockIdx.x;
  const int64_t total = n_lists * k;
  int64_t have = 0;
  for (int64_t l = 0; l < n_lists; ++l) have += counts[l * ls_cnt + b];
  const int64_t kk = have < k ? have : k;
  for (int64_t i = threadIdx.x; i < k; i += blockDim.x) {
    out_idx[b * k + i] = -1;
    out_score[b * k + i] = -INFINITY;
  }
  __syncthreads();
  for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
    const int64_t l = e / k, j = e % k;
    if (j >= counts[l * ls_cnt + b]) continue;
    const double t = scores[l * ls_rec + b * k + j];
    const int64_t r = ids[l * ls_rec + b * k + j];
    int64_t rank = 0;
    for (int64_t l2 = 0; l2 < n_lists && rank < kk; ++l2) {
      const int64_t c2 = counts[l2 * ls_cnt + b];
      const double* s2 = scores + l2 * ls_rec + b * k;
      const int64_t* i2 = ids + l2 * ls_rec + b * k;
      for (int64_t j2 = 0; j2 < c2; ++j2) {
        const double t2 = s2[j2];
        if (t2 > t || (t2 == t && i2[j2] < r)) ++rank;
        else if (t2 < t) break;                    // each list is sorted descending
      }
    }
    if (rank < kk) { out_idx[b * k + rank] = r; out_score[b * k + rank] = t; }
  }
  if (threadIdx.x == 0) out_count[b] = kk;
}

int launch_merge_topk(int64_t n_lists, int64_t nq, int64_t k, int64_t ls_rec, int64_t ls_cnt, const double* scores, const int64_t* ids,
                      const int64_t* counts, int64_t* out_idx, double* out_score, int64_t* out_count, cudaStream_t s) {
  if (nq == 0) return 0;
  merge_topk_kernel<<<(unsigned)nq, 256, 0, s>>>(n_lists, nq, k, ls_rec, ls_cnt, scores, ids, counts, out_idx, out_score, out_count);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hdb
