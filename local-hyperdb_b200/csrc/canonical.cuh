// Canonical scores: the reference's arithmetic, operation by operation (device side).
//
// The reference (hyperdb/ranking_algorithm.py) delegates to NumPy; the NumPy loops on the path are
// restated in oracle/canonical.py (the spec, checked bit for bit against the real reference) and
// mirrored here:
//   HALF_dot              -> sequential float32 chain, one final round to float16      (:29, :41)
//   pairwise_sum          -> 8 interleaved accumulators per <=128-element block, recursive halving
//                            (np.linalg.norm :9,:49 and np.sum :59)
//   float16 ufuncs        -> the float32 operation then a round to float16
//   float32/float64 np.dot-> OpenBLAS order is unknowable; canonical = exact dot rounded once
//                            (fp64 accumulation for fp32, compensated Dot2 for fp64)
// Values of the result dtype R are carried in `C` (float for f16/f32, double for f64); every
// operation is an explicit IEEE round-to-nearest intrinsic so nvcc cannot contract it into an FMA.
#pragma once
#include "hdb_common.cuh"

namespace hdb {

template <int RDT> struct Arith;
template <> struct Arith<0> {                     // float16 result dtype
  using C = float;
  static __device__ __forceinline__ C rnd(float x) { return __half2float(__float2half_rn(x)); }
  static __device__ __forceinline__ C sub(C a, C b) { return rnd(__fsub_rn(a, b)); }
  static __device__ __forceinline__ C add(C a, C b) { return rnd(__fadd_rn(a, b)); }
  static __device__ __forceinline__ C mul(C a, C b) { return rnd(__fmul_rn(a, b)); }
  static __device__ __forceinline__ C div(C a, C b) { return rnd(__fdiv_rn(a, b)); }
  static __device__ __forceinline__ C sqrt(C a) { return rnd(__fsqrt_rn(a)); }
  static __device__ __forceinline__ C acc_add(C a, C b) { return __fadd_rn(a, b); }   // float32 accumulator
  static __device__ __forceinline__ C acc_done(C a) { return rnd(a); }
  static __device__ __forceinline__ C from_double(double x) { return (float)x; }
};
template <> struct Arith<1> {                     // float32
  using C = float;
  static __device__ __forceinline__ C sub(C a, C b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ C add(C a, C b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ C mul(C a, C b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ C div(C a, C b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ C sqrt(C a) { return __fsqrt_rn(a); }
  static __device__ __forceinline__ C acc_add(C a, C b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ C acc_done(C a) { return a; }
  static __device__ __forceinline__ C from_double(double x) { return (float)x; }
};
template <> struct Arith<2> {                     // float64
  using C = double;
  static __device__ __forceinline__ C sub(C a, C b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ C add(C a, C b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ C mul(C a, C b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ C div(C a, C b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ C sqrt(C a) { return __dsqrt_rn(a); }
  static __device__ __forceinline__ C acc_add(C a, C b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ C acc_done(C a) { return a; }
  static __device__ __forceinline__ C from_double(double x) { return x; }
};

// NumPy's pairwise summation of term(0..n-1); `Term` returns accumulator-typed values.
template <int RDT, typename Term>
__device__ typename Arith<RDT>::C pairwise_sum(Term term, int n) {
  using A = Arith<RDT>;
  using C = typename A::C;
  auto leaf = [&](int off, int m) -> C {
    if (m < 8) {
      C acc = 0;
      for (int i = 0; i < m; ++i) acc = A::acc_add(acc, term(off + i));
      return acc;
    }
    C r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = term(off + j);
    int full = m - (m % 8), i = 8;
    for (; i < full; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = A::acc_add(r[j], term(off + i + j));
    }
    C acc = A::acc_add(A::acc_add(A::acc_add(r[0], r[1]), A::acc_add(r[2], r[3])),
                       A::acc_add(A::acc_add(r[4], r[5]), A::acc_add(r[6], r[7])));
    for (; i < m; ++i) acc = A::acc_add(acc, term(off + i));
    return acc;
  };
  if (n <= 128) return A::acc_done(leaf(0, n));
  struct Frame { int off, n, state; C left; };
  Frame st[30];
  int sp = 0;
  st[0] = Frame{0, n, 0, C(0)};
  C ret = 0;
  while (sp >= 0) {
    Frame& f = st[sp];
    if (f.n <= 128) { ret = leaf(f.off, f.n); --sp; continue; }
    int half = f.n / 2;
    half -= half % 8;
    if (f.state == 0) { f.state = 1; st[sp + 1] = Frame{f.off, half, 0, C(0)}; ++sp; continue; }
    if (f.state == 1) { f.left = ret; f.state = 2; st[sp + 1] = Frame{f.off + half, f.n - half, 0, C(0)}; ++sp; continue; }
    ret = A::acc_add(f.left, ret);
    --sp;
  }
  return A::acc_done(ret);
}

// np.linalg.norm of one row/vector in dtype DT: sqrt(add.reduce(x*x)); zero -> caller's business.
template <int DT>
__device__ typename Arith<DT>::C canonical_norm(const void* x, int64_t d) {
  using A = Arith<DT>;
  using C = typename A::C;
  auto term = [&](int i) -> C { C v = A::from_double(load_as_double(x, DT, i)); return A::mul(v, v); };
  return A::sqrt(pairwise_sum<DT>(term, (int)d));
}

// v / norm in the STORAGE dtype's arithmetic (get_norm_vector runs in the array's own dtype).
__device__ __forceinline__ double unit_elem(double v, double norm, int dt) {
  if (dt == 0) return (double)Arith<0>::div((float)v, (float)norm);
  if (dt == 1) return (double)Arith<1>::div((float)v, (float)norm);
  return Arith<2>::div(v, norm);
}

// Exact dot product rounded once to R (canonical for OpenBLAS-backed np.dot); HALF_dot for f16.
template <int RDT, typename GetV, typename GetQ>
__device__ typename Arith<RDT>::C canonical_dot(GetV gv, GetQ gq, int d) {
  if (RDT == 0) {
    float acc = 0.f;
    for (int j = 0; j < d; ++j) acc = __fmaf_rn((float)gv(j), (float)gq(j), acc);   // products exact in fp32
    return Arith<RDT>::from_double((double)Arith<0>::rnd(acc));
  } else if (RDT == 1) {
    double acc = 0.0;
    for (int j = 0; j < d; ++j) acc = __fma_rn((double)gv(j), (double)gq(j), acc);  // products exact in fp64
    return Arith<RDT>::from_double((double)(float)acc);
  } else {
    double s = 0.0, c = 0.0;                                                        // Ogita-Rump-Oishi Dot2
    for (int j = 0; j < d; ++j) {
      double a = (double)gv(j), b = (double)gq(j);
      double p = __dmul_rn(a, b);
      double e = __fma_rn(a, b, -p);
      double t = __dadd_rn(s, p);
      double z = __dsub_rn(t, s);
      double err = __dadd_rn(__dsub_rn(s, __dsub_rn(t, z)), __dsub_rn(p, z));
      c = __dadd_rn(c, __dadd_rn(e, err));
      s = t;
    }
    return Arith<RDT>::from_double(__dadd_rn(s, c));
  }
}

struct CanonArgs {
  const void* rows;      // storage, row-major
  int sdt;               // storage dtype
  int64_t d;
  const double* qc;      // canonical query (values of dtype R, or unit query for cosine), length d
  const uint32_t* bits;  // packed sign bits (hamming) or nullptr
  const uint32_t* qbits;
  int words;             // 32-bit words per packed row
  int metric;
};

// Similarity of one row in the reference's arithmetic, returned as the R-typed value widened to double
// (hamming: the integer D - popcount).
template <int RDT>
__device__ double canonical_similarity(const CanonArgs& a, int64_t row, double nrm) {
  using A = Arith<RDT>;
  using C = typename A::C;
  const int d = (int)a.d;
  if (a.metric == 4) {
    const uint32_t* r = a.bits + row * (int64_t)a.words;
    int diff = 0;
    for (int w = 0; w < a.words; ++w) diff += __popc(r[w] ^ a.qbits[w]);
    return (double)(d - diff);
  }
  const char* base = reinterpret_cast<const char*>(a.rows) + row * a.d * dtype_size(a.sdt);
  auto gq = [&](int j) -> C { return A::from_double(a.qc[j]); };
  if (a.metric == 0) {
    auto gv = [&](int j) -> C { return A::from_double(load_as_double(base, a.sdt, j)); };
    return (double)canonical_dot<RDT>(gv, gq, d);
  }
  if (a.metric == 1) {
    auto gv = [&](int j) -> C { return A::from_double(unit_elem(load_as_double(base, a.sdt, j), nrm, a.sdt)); };
    return (double)canonical_dot<RDT>(gv, gq, d);
  }
  if (a.metric == 2) {
    auto term = [&](int j) -> C {
      C df = A::sub(A::from_double(load_as_double(base, a.sdt, j)), gq(j));
      return A::mul(df, df);
    };
    C dist = A::sqrt(pairwise_sum<RDT>(term, d));
    return (double)A::div(C(1), A::add(C(1), dist));
  }
  auto term = [&](int j) -> C { return fabs(A::sub(A::from_double(load_as_double(base, a.sdt, j)), gq(j))); };
  C dist = pairwise_sum<RDT>(term, d);
  return (double)A::div(C(1), A::add(C(1), dist));
}

// nrm: the row's canonical norm (zero already replaced by 1); only read for cosine
__device__ __forceinline__ double canonical_similarity_rt(const CanonArgs& a, int rdt, int64_t row, double nrm) {
  if (rdt == 0) return canonical_similarity<0>(a, row, nrm);
  if (rdt == 1) return canonical_similarity<1>(a, row, nrm);
  return canonical_similarity<2>(a, row, nrm);
}

// ranking_algorithm.py:171-186: float64 score, NaN -> -inf, + recency_bias * exp(ts - max ts)
__device__ __forceinline__ double total_score(double sim, const double* decay, double bias, int64_t row) {
  if (sim != sim) sim = -INFINITY;
  return decay ? __dadd_rn(sim, __dmul_rn(bias, decay[row])) : sim;
}

}  // namespace hdb
