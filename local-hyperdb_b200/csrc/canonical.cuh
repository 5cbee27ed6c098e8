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

// Warp-cooperative form of the same summation, bit-identical to pairwise_sum: the recursion's leaves
// (<=128 elements) are independent, and inside a leaf the 8 interleaved accumulators are independent
// chains, so 8 lanes own one leaf (4 leaves per warp at a time); the 8 accumulators are combined by a
// 3-step butterfly, which is exactly ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) because IEEE addition is
// commutative.  Leaf results are then combined in the recursion's post-order by every lane redundantly.
// `term(i)` must be readable by any lane; all 32 lanes must call.  scratch: kPwMaxLeaves int2 + C slots.
constexpr int kPwMaxLeaves = 64;
struct PwScratch { int2 leaf[kPwMaxLeaves]; double res[kPwMaxLeaves]; };

template <int RDT, typename Term>
__device__ typename Arith<RDT>::C pairwise_sum_warp(Term term, int n, int lane, PwScratch* sc) {
  using A = Arith<RDT>;
  using C = typename A::C;
  // enumerate the leaves in order (every lane walks the same recursion; lane 0 records them)
  int nleaves = 0;
  {
    int2 st[30];
    int sp = 0;
    st[0] = make_int2(0, n);
    while (sp >= 0) {
      const int2 f = st[sp--];
      if (f.y <= 128) {
        if (nleaves < kPwMaxLeaves && lane == 0) sc->leaf[nleaves] = f;
        ++nleaves;
        continue;
      }
      int half = f.y / 2;
      half -= half % 8;
      st[++sp] = make_int2(f.x + half, f.y - half);     // right pushed first: left is popped first
      st[++sp] = make_int2(f.x, half);
    }
  }
  if (nleaves > kPwMaxLeaves) return pairwise_sum<RDT>(term, n);      // very long rows: sequential form
  __syncwarp();
  const int sub = lane & 7, grp = lane >> 3;
  for (int base = 0; base < nleaves; base += 4) {
    const int li = base + grp;
    const bool have = li < nleaves;
    const int2 f = have ? sc->leaf[li] : make_int2(0, 0);
    C acc = 0;
    const int m = f.y, full = m - (m % 8);
    if (have && m >= 8) {
      acc = term(f.x + sub);
      for (int i = 8; i < full; i += 8) acc = A::acc_add(acc, term(f.x + i + sub));
    }
    acc = A::acc_add(acc, __shfl_xor_sync(kFull, acc, 1));
    acc = A::acc_add(acc, __shfl_xor_sync(kFull, acc, 2));
    acc = A::acc_add(acc, __shfl_xor_sync(kFull, acc, 4));
    if (have && sub == 0) {
      if (m < 8) {
        acc = 0;
        for (int i = 0; i < m; ++i) acc = A::acc_add(acc, term(f.x + i));
      } else {
        for (int i = full; i < m; ++i) acc = A::acc_add(acc, term(f.x + i));
      }
      sc->res[li] = (double)acc;
    }
  }
  __syncwarp();
  // combine the leaf results in post-order (same stack machine as pairwise_sum, leaves looked up)
  struct Frame { int off, n, state; C left; };
  Frame st[30];
  int sp = 0, next_leaf = 0;
  st[0] = Frame{0, n, 0, C(0)};
  C ret = 0;
  while (sp >= 0) {
    Frame& f = st[sp];
    if (f.n <= 128) { ret = A::from_double(sc->res[next_leaf++]); --sp; continue; }
    int half = f.n / 2;
    half -= half % 8;
    if (f.state == 0) { f.state = 1; st[sp + 1] = Frame{f.off, half, 0, C(0)}; ++sp; continue; }
    if (f.state == 1) { f.left = ret; f.state = 2; st[sp + 1] = Frame{f.off + half, f.n - half, 0, C(0)}; ++sp; continue; }
    ret = A::acc_add(f.left, ret);
    --sp;
  }
  __syncwarp();
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
#pragma unroll 8
    for (int j = 0; j < d; ++j) acc = __fmaf_rn((float)gv(j), (float)gq(j), acc);   // products exact in fp32
    return Arith<RDT>::from_double((double)Arith<0>::rnd(acc));
  } else if (RDT == 1) {
    double acc = 0.0;
#pragma unroll 8
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
  int sdt;               // storage dtype
  int64_t d;
  const double* qc;      // canonical query (values of dtype R, or unit query for cosine), length d
  const uint32_t* qbits; // packed sign bits of the query (hamming) or nullptr
  int words;             // 32-bit words per packed row
  int metric;
};

// Element-wise half of a canonical score (parallelisable): the per-column term the reference feeds into
// its reduction.  v = stored element, q = canonical query element (both widened to double).
template <int RDT>
__device__ __forceinline__ typename Arith<RDT>::C canonical_term(int metric, double v, double q, double nrm, int sdt) {
  using A = Arith<RDT>;
  using C = typename A::C;
  if (metric == 0) return A::from_double(v);
  if (metric == 1) return A::from_double(unit_elem(v, nrm, sdt));
  C df = A::sub(A::from_double(v), A::from_double(q));
  return metric == 2 ? A::mul(df, df) : C(fabs(df));
}

// Order-dependent half: the reference's reduction over the terms (sequential by nature).
template <int RDT, typename GetTerm, typename GetQ>
__device__ double canonical_reduce(int metric, GetTerm term, GetQ gq, int d) {
  using A = Arith<RDT>;
  using C = typename A::C;
  if (metric <= 1) return (double)canonical_dot<RDT>(term, gq, d);
  C dist = pairwise_sum<RDT>(term, d);
  if (metric == 2) dist = A::sqrt(dist);
  return (double)A::div(C(1), A::add(C(1), dist));
}

// Similarity of one row in the reference's arithmetic, returned as the R-typed value widened to double
// (hamming: the integer D - popcount).  `rowp` points at the row's d stored elements, `bitrow` at its
// packed sign bits; nrm is the row's canonical norm (cosine only).
template <int RDT>
__device__ double canonical_similarity(const CanonArgs& a, const void* rowp, const uint32_t* bitrow, double nrm) {
  using A = Arith<RDT>;
  using C = typename A::C;
  const int d = (int)a.d;
  if (a.metric == 4) {
    int diff = 0;
    for (int w = 0; w < a.words; ++w) diff += __popc(bitrow[w] ^ a.qbits[w]);
    return (double)(d - diff);
  }
  if (a.metric == 5) {          // jaccard: uint64 / uint64 -> float64 true division; 0/0 = NaN (ranking_algorithm.py:75)
    int inter = 0, uni = 0;
    for (int w = 0; w < a.words; ++w) { inter += __popc(bitrow[w] & a.qbits[w]); uni += __popc(bitrow[w] | a.qbits[w]); }
    return __ddiv_rn((double)inter, (double)uni);
  }
  auto gq = [&](int j) -> C { return A::from_double(a.qc[j]); };
  auto term = [&](int j) -> C { return canonical_term<RDT>(a.metric, load_as_double(rowp, a.sdt, j), a.qc[j], nrm, a.sdt); };
  return canonical_reduce<RDT>(a.metric, term, gq, d);
}

__device__ __forceinline__ double canonical_similarity_rt(const CanonArgs& a, int rdt, const void* rowp,
                                                          const uint32_t* bitrow, double nrm) {
  if (rdt == 0) return canonical_similarity<0>(a, rowp, bitrow, nrm);
  if (rdt == 1) return canonical_similarity<1>(a, rowp, bitrow, nrm);
  return canonical_similarity<2>(a, rowp, bitrow, nrm);
}

// ranking_algorithm.py:171-186: float64 score, NaN -> -inf, + recency_bias * exp(ts - max ts)
__device__ __forceinline__ double total_score(double sim, const double* decay, double bias, int64_t row) {
  if (sim != sim) sim = -INFINITY;
  return decay ? __dadd_rn(sim, __dmul_rn(bias, decay[row])) : sim;
}

}  // namespace hdb
