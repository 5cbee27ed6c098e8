// Host build of the PRODUCT's canonical arithmetic (local-hyperdb_b200/csrc/canonical.cuh) for CPU-side tests:
// the CUDA round-to-nearest intrinsics are mapped to plain IEEE operations (compiled with -ffp-contract=off, so no
// FMA contraction), cuda_fp16.h supplies the host half conversions, and the same templates that the certify / exact
// kernels instantiate are exported through a small C interface.  TEST INFRASTRUCTURE: lets tests/test_emul_canonical.py
// compare the device source with NumPy on thousands of random inputs without a GPU.
#include <cmath>
#include <cstdint>
#include <cstring>

#include <cuda_runtime.h>
#include <cuda_fp16.h>

static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline float __fmaf_rn(float a, float b, float c) { return std::fma(a, b, c); }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline unsigned __shfl_xor_sync(unsigned, unsigned v, int) { return v; }        // warp forms are not instantiated here
static inline float __shfl_xor_sync(unsigned, float v, int) { return v; }
static inline double __shfl_xor_sync(unsigned, double v, int) { return v; }
static inline unsigned long long __shfl_xor_sync(unsigned, unsigned long long v, int) { return v; }
static inline void __syncwarp() {}
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }

#include "../../local-hyperdb_b200/csrc/canonical.cuh"
#include "../../local-hyperdb_b200/csrc/certificate.cuh"

using namespace hdb;

extern "C" {

// NumPy's pairwise sum of n values of dtype dt (0 f16, 1 f32, 2 f64), result widened to double
double emul_pairwise_sum(int dt, const void* x, int n) {
  if (dt == 0) return (double)pairwise_sum<0>([&](int i) { return (float)load_as_double(x, 0, i); }, n);
  if (dt == 1) return (double)pairwise_sum<1>([&](int i) { return (float)load_as_double(x, 1, i); }, n);
  return pairwise_sum<2>([&](int i) { return load_as_double(x, 2, i); }, n);
}

// the same sum through the host-built pairwise PLAN the row-wise kernels use (csrc/rowwise.cu); -1 leaves = no plan
double emul_pairwise_sum_plan(int dt, const void* x, int n, int* nleaves) {
  PwPlan plan;
  if (!pw_plan_build(plan, n)) { *nleaves = -1; return 0.0; }
  *nleaves = plan.nleaves;
  if (dt == 0) return (double)pairwise_sum_plan_seq<0>([&](int i) { return (float)load_as_double(x, 0, i); }, plan);
  if (dt == 1) return (double)pairwise_sum_plan_seq<1>([&](int i) { return (float)load_as_double(x, 1, i); }, plan);
  return pairwise_sum_plan_seq<2>([&](int i) { return load_as_double(x, 2, i); }, plan);
}

double emul_norm(int dt, const void* x, int64_t d) {
  if (dt == 0) return (double)canonical_norm<0>(x, d);
  if (dt == 1) return (double)canonical_norm<1>(x, d);
  return canonical_norm<2>(x, d);
}

void emul_mean_std(int dt, const void* x, int64_t d, int scalar, double* mean, double* stdv) {
  if (dt == 0) { float m, s; canonical_mean_std<0>(x, d, scalar != 0, &m, &s); *mean = m; *stdv = s; }
  else if (dt == 1) { float m, s; canonical_mean_std<1>(x, d, scalar != 0, &m, &s); *mean = m; *stdv = s; }
  else canonical_mean_std<2>(x, d, scalar != 0, mean, stdv);
}

// similarity of one stored row in the reference's arithmetic (what the exact path and the certify step compute).
// qc: the prepared canonical query as prep_query stores it (float64 carrier): the query itself, the unit query (cosine) or
// q - mean(q) (pearson); nrm / aux2: the row's canonical norm (cosine) or mean / std (pearson); qstd: np.std(q) (pearson)
double emul_similarity(int rdt, int sdt, int metric, const void* row, const double* qc, int64_t d, double nrm, double aux2, double qstd,
                       const uint32_t* bitrow, const uint32_t* qbits, int words) {
  CanonArgs a;
  a.sdt = sdt; a.d = d; a.qc = qc; a.qbits = qbits; a.words = words; a.metric = metric; a.qstd = qstd; a.distance = 0;
  return canonical_similarity_rt(a, rdt, row, bitrow, nrm, aux2);
}

// csrc/certificate.cuh::outsider_bound with the statistics a shard would carry
double emul_outsider_bound(double s, int metric, int rdt, int sdt, int64_t d, double max_norm, double max_ratio, double max_pratio,
                           double max_cratio, double min_pstd, int has_decay, double bias, int tensor_path, double qnorm, double qstd) {
  FinalizeArgs a;
  std::memset(&a, 0, sizeof(a));
  static double dummy_decay = 0.0;
  static unsigned dummy_count = 0;
  a.m.d = d; a.m.dtype = sdt;
  a.m.max_norm = (float)max_norm; a.m.max_ratio = (float)max_ratio;
  a.m.max_pratio = (float)max_pratio; a.m.max_cratio = (float)max_cratio; a.m.min_pstd = (float)min_pstd;
  a.rdt = rdt; a.metric = metric;
  a.f.decay = has_decay ? &dummy_decay : nullptr;
  a.f.bias = bias;
  a.cand_count = tensor_path ? &dummy_count : nullptr;
  return outsider_bound(s, a, qnorm, qstd);
}

// the batched tensor pass's certificate: the same statistics plus sum_j (q_j - mean_q) (pearson: the mean correction that pass
// leaves out); s is the key score already divided by std_q, as finalize.cu passes it
double emul_outsider_bound_tc(double s, int metric, int rdt, int sdt, int64_t d, double max_norm, double max_ratio, double max_pratio,
                              double max_cratio, double min_pstd, double qnorm, double qstd, double qsumb) {
  FinalizeArgs a;
  std::memset(&a, 0, sizeof(a));
  static unsigned dummy_count = 0;
  a.m.d = d; a.m.dtype = sdt;
  a.m.max_norm = (float)max_norm; a.m.max_ratio = (float)max_ratio;
  a.m.max_pratio = (float)max_pratio; a.m.max_cratio = (float)max_cratio; a.m.min_pstd = (float)min_pstd;
  a.rdt = rdt; a.metric = metric;
  a.cand_count = &dummy_count;
  return outsider_bound(s, a, qnorm, qstd, qsumb);
}

double emul_unit_elem(double v, double norm, int dt) { return unit_elem(v, norm, dt); }
double emul_sub_in(double x, double y, int dt) { return sub_in(x, y, dt); }

}  // extern "C"
