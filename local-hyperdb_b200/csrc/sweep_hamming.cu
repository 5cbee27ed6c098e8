// Host side of the bit-packed popcount sweeps (hamming_distance, jaccard_similarity): picks the kernel form and calls the
// (candidate class, metric) instantiation of csrc/sweep_hamming_inst.cu -- four translation units that build in parallel.
#include <cstdlib>

#include "sweep_hamming.cuh"
#include "sweep_hamming.h"

namespace hdb {

#define HDB_DECL_HAM(name) int name(const HammingParams& hp, int nq, int grid, int form, size_t smem, cudaStream_t s)
HDB_DECL_HAM(hamming_kp32_ham); HDB_DECL_HAM(hamming_kp32_jac); HDB_DECL_HAM(hamming_kp128_ham); HDB_DECL_HAM(hamming_kp128_jac);
#undef HDB_DECL_HAM

// packed-bit sweeps: up to 4 queries per pass (2 with the wide candidate class); rows beyond 32 vectors: 1
static int hamming_max_group(const MatrixView& m, int kp) {
  if (m.words / 4 > 32) return 1;
  return kp <= 32 ? 4 : 2;
}

// A/B testing: HDB_HAMMING_COOPERATIVE=1 keeps the register-cooperative hamming kernels for every shape
static bool force_cooperative_hamming() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HDB_HAMMING_COOPERATIVE"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
static size_t list_smem(int kp) { return (size_t)kSweepWarps * (kp <= 32 ? 128 : 2 * kp) * 8 + 16; }

int hamming_sweep_max_group(const MatrixView& m, int kp) { return hamming_max_group(m, kp); }

int launch_hamming_sweep(const MatrixView& m, int metric, const uint32_t* qbits, const RowFilter& f, int kp, const SweepOut& out, int nq,
                         cudaStream_t s) {
  HammingParams hp;
  hp.jaccard = metric == HDB_JACCARD;
  hp.bits = m.bits; hp.qbits = qbits; hp.n = m.n; hp.d = m.d;
  hp.nvec = m.words / 4;
  int lpr = 1;
  while (lpr < hp.nvec && lpr < 32) lpr <<= 1;
  hp.lpr = lpr;
  hp.f = f; hp.cand = out.cand; hp.tau = out.tau;
  hp.cand_stride = (int64_t)out.grid * kp;
  size_t smem = list_smem(kp) + (size_t)hp.nvec * 16;
  if (smem > 200 * 1024) return fail("sweep: dimension too large for the fused hamming pass");
  typedef int (*Fn)(const HammingParams&, int, int, int, size_t, cudaStream_t);
  static const Fn table[2][2] = {{hamming_kp32_ham, hamming_kp32_jac}, {hamming_kp128_ham, hamming_kp128_jac}};
  const Fn fn = table[kp <= 32 ? 0 : 1][hp.jaccard ? 1 : 0];
  // form 0: contiguous windows, cp.async-staged, lane-per-row; 1: LPR lanes per row (masked shards, wider rows);
  // 2: rows beyond 4096 bits (one query per pass)
  if (hp.nvec <= 8 && !f.mask && !force_cooperative_hamming()) return fn(hp, nq, out.grid, 0, smem, s);
  if (hp.nvec <= 32) return fn(hp, nq, out.grid, 1, smem, s);
  if (nq != 1) return fail("sweep: rows beyond 4096 bits take one query per pass");
  return fn(hp, nq, out.grid, 2, smem, s);
}

}  // namespace hdb
