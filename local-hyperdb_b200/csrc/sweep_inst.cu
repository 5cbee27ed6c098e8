// One (storage type, metric class) instantiation of the float sweep (csrc/sweep_float.cuh); compiled nine times by the
// Makefile with -DHDB_SWEEP_T=<__half|float|double> -DHDB_SWEEP_MC=<0|1|2> -DHDB_SWEEP_FN=<symbol>.
#include "sweep_float.cuh"

namespace hdb {
int HDB_SWEEP_FN(const SweepParams& p, bool vec, int kp, int nq, int grid, size_t smem, cudaStream_t s) {
  return launch_kp<HDB_SWEEP_T, HDB_SWEEP_MC>(p, vec, kp, nq, grid, smem, s);
}
}  // namespace hdb
