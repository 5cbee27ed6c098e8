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
  static __device__ __forceinline__ C rnd_c(C x) { return rnd(x); }
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
  static __device__ __forceinline__ C rnd_c(C x) { return x; }
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
  static __device__ __forceinline__ C rnd_c(C x) { return x; }
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

// ---------------------------------------------------------------------------------------------
// Pairwise PLAN: the shape of NumPy's pairwise recursion depends only on n, so the row-wise kernels (csrc/rowwise.cu:
// one warp per row, every row has the same length) get it precomputed once on the host instead of walking the
// recursion per row: the leaves (offset, length <= 128) in order, and the post-order additions as a flat program
// over result slots (leaf i -> slot i, addition t -> slot kPlanLeaves + t).
// ---------------------------------------------------------------------------------------------
constexpr int kPlanLeaves = 64;
struct PwPlan {
  int n, nleaves, nadds, root;
  int leaf_off[kPlanLeaves];
  short leaf_len[kPlanLeaves];
  unsigned char add_dst[kPlanLeaves], add_a[kPlanLeaves], add_b[kPlanLeaves];
};

// false: more than kPlanLeaves leaves (rows longer than ~4-8 k elements): the caller keeps the per-row recursion
__host__ __device__ inline bool pw_plan_build(PwPlan& p, int n) {
  p.n = n; p.nleaves = 0; p.nadds = 0; p.root = 0;
  struct Frame { int off, n, state, left; };
  Frame st[32];
  int sp = 0, ret = 0;
  st[0] = Frame{0, n, 0, 0};
  while (sp >= 0) {
    Frame& f = st[sp];
    if (f.n <= 128) {
      if (p.nleaves >= kPlanLeaves) return false;
      p.leaf_off[p.nleaves] = f.off;
      p.leaf_len[p.nleaves] = (short)f.n;
      ret = p.nleaves++;
      --sp;
      continue;
    }
    int half = f.n / 2;
    half -= half % 8;
    if (f.state == 0) { f.state = 1; st[sp + 1] = Frame{f.off, half, 0, 0}; ++sp; continue; }
    if (f.state == 1) { f.left = ret; f.state = 2; st[sp + 1] = Frame{f.off + half, f.n - half, 0, 0}; ++sp; continue; }
    const int dst = kPlanLeaves + p.nadds;
    p.add_dst[p.nadds] = (unsigned char)dst;
    p.add_a[p.nadds] = (unsigned char)f.left;
    p.add_b[p.nadds] = (unsigned char)ret;
    ++p.nadds;
    ret = dst;
    --sp;
  }
  p.root = ret;
  return true;
}

// One leaf exactly as pairwise_sum's: 8 interleaved accumulators, ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail.
template <int RDT, typename Term>
__device__ inline typename Arith<RDT>::C pw_leaf_seq(Term term, int off, int m) {
  using A = Arith<RDT>;
  using C = typename A::C;
  if (m < 8) {
    C acc = 0;
    for (int i = 0; i < m; ++i) acc = A::acc_add(acc, term(off + i));
    return acc;
  }
  C r[8];
  for (int j = 0; j < 8; ++j) r[j] = term(off + j);
  const int full = m - (m % 8);
  int i = 8;
  for (; i < full; i += 8)
    for (int j = 0; j < 8; ++j) r[j] = A::acc_add(r[j], term(off + i + j));
  C acc = A::acc_add(A::acc_add(A::acc_add(r[0], r[1]), A::acc_add(r[2], r[3])),
                     A::acc_add(A::acc_add(r[4], r[5]), A::acc_add(r[6], r[7])));
  for (; i < m; ++i) acc = A::acc_add(acc, term(off + i));
  return acc;
}

// Sequential evaluation of a plan (the host-side tests check it against np.add.reduce; the warp form below does the
// same additions in the same order).
template <int RDT, typename Term>
__device__ inline typename Arith<RDT>::C pairwise_sum_plan_seq(Term term, const PwPlan& plan) {
  using A = Arith<RDT>;
  using C = typename A::C;
  C res[2 * kPlanLeaves];
  for (int l = 0; l < plan.nleaves; ++l) res[l] = pw_leaf_seq<RDT>(term, plan.leaf_off[l], plan.leaf_len[l]);
  for (int t = 0; t < plan.nadds; ++t) res[plan.add_dst[t]] = A::acc_add(res[plan.add_a[t]], res[plan.add_b[t]]);
  return A::acc_done(res[plan.root]);
}

#ifdef __CUDACC__
// Warp form: 8 lanes own one leaf (the 8 interleaved accumulators), 4 leaves per round, 3-step butterfly; lane 0 then
// runs the addition program over the per-warp result slots `res` (2*kPlanLeaves values of C in shared memory).
// term(i) must be readable by every lane; all 32 lanes call and all return the sum.
template <int RDT, typename Term>
__device__ __forceinline__ typename Arith<RDT>::C pairwise_sum_plan(Term term, const PwPlan& plan, int lane, typename Arith<RDT>::C* res) {
  using A = Arith<RDT>;
  using C = typename A::C;
  const int sub = lane & 7, grp = lane >> 3;
  for (int base = 0; base < plan.nleaves; base += 4) {
    const int li = base + grp;
    const bool have = li < plan.nleaves;
    const int off = have ? plan.leaf_off[li] : 0;
    const int m = have ? plan.leaf_len[li] : 0, full = m - (m % 8);
    C acc = 0;
    if (m >= 8) {
      acc = term(off + sub);
      for (int i = 8; i < full; i += 8) acc = A::acc_add(acc, term(off + i + sub));
    }
    acc = A::acc_add(acc, __shfl_xor_sync(kFull, acc, 1));
    acc = A::acc_add(acc, __shfl_xor_sync(kFull, acc, 2));
    acc = A::acc_add(acc, __shfl_xor_sync(kFull, acc, 4));
    if (have && sub == 0) {
      if (m < 8) {
        acc = 0;
        for (int i = 0; i < m; ++i) acc = A::acc_add(acc, term(off + i));
      } else {
        for (int i = full; i < m; ++i) acc = A::acc_add(acc, term(off + i));
      }
      res[li] = acc;
    }
  }
  __syncwarp();
  if (lane == 0)
    for (int t = 0; t < plan.nadds; ++t) res[plan.add_dst[t]] = A::acc_add(res[plan.add_a[t]], res[plan.add_b[t]]);
  __syncwarp();
  const C out = A::acc_done(res[plan.root]);
  __syncwarp();                                  // `res` may be rewritten by the caller's next sum
  return out;
}
#endif

// np.linalg.norm of one row/vector in dtype DT: sqrt(add.reduce(x*x)); zero -> caller's business.
template <int DT>
__device__ typename Arith<DT>::C canonical_norm(const void* x, int64_t d) {
  using A = Arith<DT>;
  using C = typename A::C;
  auto term = [&](int i) -> C { C v = A::from_double(load_as_double(x, DT, i)); return A::mul(v, v); };
  return A::sqrt(pairwise_sum<DT>(term, (int)d));
}

// float64 -> dtype DT in ONE rounding (a float64 ufunc result cast to the output array's dtype), carried in C
template <int DT>
__device__ __forceinline__ typename Arith<DT>::C round_to(double x) {
  if (DT == 0) return (typename Arith<DT>::C)__half2float(__double2half(x));
  if (DT == 1) return (typename Arith<DT>::C)(float)x;
  return (typename Arith<DT>::C)x;
}

// np.mean / np.std of d values of dtype DT (numpy._core._methods._mean / _var / _std), `val(j)` = element j in C:
//   mean: pairwise sum (float16: in float32, unrounded), / d formed in float64 (the count is a NumPy integer scalar),
//         cast to DT -- through float32 first for a float16 ARRAY result, directly for a 1-D input (`scalar`);
//   std : arrmean = DT(sum_DT / d), dev = (x - arrmean)^2 in DT, sqrt(DT(sum_DT(dev) / d)).
// hyperdb/ranking_algorithm.py:90-94.  `sum` abstracts the pairwise summation (sequential or warp-cooperative).
template <int DT, typename Val, typename Sum32, typename SumDT>
__device__ void mean_std_core(Val val, int d, bool scalar, Sum32 sum_f32, SumDT sum_dt, typename Arith<DT>::C* mean,
                              typename Arith<DT>::C* stdv) {
  using A = Arith<DT>;
  using C = typename A::C;
  const double dd = (double)d;
  C sum_rounded;
  if (DT == 0) {
    const float s32 = (float)sum_f32([&](int j) { return (float)val(j); });
    const double qd = (double)s32 / dd;
    *mean = scalar ? round_to<DT>(qd) : round_to<DT>((double)(float)qd);
    sum_rounded = A::rnd_c((C)s32);
  } else {
    sum_rounded = sum_dt([&](int j) { return val(j); });
    *mean = round_to<DT>((double)sum_rounded / dd);
  }
  const C arrmean = round_to<DT>((double)sum_rounded / dd);
  const C ssq = sum_dt([&](int j) { const C df = A::sub(val(j), arrmean); return A::mul(df, df); });
  *stdv = A::sqrt(round_to<DT>((double)ssq / dd));
}

template <int DT>
__device__ void canonical_mean_std(const void* x, int64_t d, bool scalar, typename Arith<DT>::C* mean, typename Arith<DT>::C* stdv) {
  using C = typename Arith<DT>::C;
  auto val = [&](int j) -> C { return Arith<DT>::from_double(load_as_double(x, DT, j)); };
  mean_std_core<DT>(val, (int)d, scalar,
                    [&](auto term) { return pairwise_sum<1>(term, (int)d); },
                    [&](auto term) { return pairwise_sum<DT>(term, (int)d); }, mean, stdv);
}

// x - y in dtype `dt`'s arithmetic (runtime switch)
__device__ __forceinline__ double sub_in(double x, double y, int dt) {
  if (dt == 0) return (double)Arith<0>::sub((float)x, (float)y);
  if (dt == 1) return (double)Arith<1>::sub((float)x, (float)y);
  return Arith<2>::sub(x, y);
}

// pearson_correlation's last lines (ranking_algorithm.py:100-111): denominator = std_v * std_q * d in R; quotient only where
// the denominator is non-zero (np.zeros elsewhere); NaN whenever either side is constant.
template <int RDT>
__device__ __forceinline__ double pearson_quotient(typename Arith<RDT>::C num, double std_v, double std_q, int d) {
  using A = Arith<RDT>;
  using C = typename A::C;
  if (std_v == 0.0 || std_q == 0.0) return __longlong_as_double(0x7ff8000000000000ll);
  const C den = A::mul(A::mul(A::from_double(std_v), A::from_double(std_q)), round_to<RDT>((double)d));
  return den != C(0) ? (double)A::div(num, den) : 0.0;
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
  double qstd;           // pearson: np.std(query) (query dtype, widened)
  int distance;          // euclidean: return the distance itself (get_similarity_score=False, ranking_algorithm.py:49-52)
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
__device__ double canonical_reduce(int metric, GetTerm term, GetQ gq, int d, bool distance = false) {
  using A = Arith<RDT>;
  using C = typename A::C;
  if (metric <= 1) return (double)canonical_dot<RDT>(term, gq, d);
  C dist = pairwise_sum<RDT>(term, d);
  if (metric == 2) dist = A::sqrt(dist);
  if (distance) return (double)dist;
  return (double)A::div(C(1), A::add(C(1), dist));
}

// Similarity of one row in the reference's arithmetic, returned as the R-typed value widened to double
// (hamming: the integer D - popcount).  `rowp` points at the row's d stored elements, `bitrow` at its
// packed sign bits; nrm is the row's canonical norm (cosine only).
template <int RDT>
__device__ double canonical_similarity(const CanonArgs& a, const void* rowp, const uint32_t* bitrow, double nrm, double aux2 = 0.0) {
  using A = Arith<RDT>;
  using C = typename A::C;
  const int d = (int)a.d;
  if (a.metric == 6) {          // pearson: nrm = np.mean(row), aux2 = np.std(row); qc = q - np.mean(q) in the query's dtype
    auto pterm = [&](int j) -> C {
      return A::mul(A::from_double(sub_in(load_as_double(rowp, a.sdt, j), nrm, a.sdt)), A::from_double(a.qc[j]));
    };
    return pearson_quotient<RDT>(pairwise_sum<RDT>(pterm, d), aux2, a.qstd, d);
  }
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
  return canonical_reduce<RDT>(a.metric, term, gq, d, a.metric == 2 && a.distance != 0);
}

__device__ __forceinline__ double canonical_similarity_rt(const CanonArgs& a, int rdt, const void* rowp,
                                                          const uint32_t* bitrow, double nrm, double aux2 = 0.0) {
  if (rdt == 0) return canonical_similarity<0>(a, rowp, bitrow, nrm, aux2);
  if (rdt == 1) return canonical_similarity<1>(a, rowp, bitrow, nrm, aux2);
  return canonical_similarity<2>(a, rowp, bitrow, nrm, aux2);
}

// ranking_algorithm.py:171-186: float64 score, NaN -> -inf, + recency_bias * exp(ts - max ts)
__device__ __forceinline__ double total_score(double sim, const double* decay, double bias, int64_t row) {
  if (sim != sim) sim = -INFINITY;
  return decay ? __dadd_rn(sim, __dmul_rn(bias, decay[row])) : sim;
}

}  // namespace hdb
