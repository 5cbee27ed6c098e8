// One (storage type, metric class, candidate class) instantiation of the float sweep (csrc/sweep_float.cuh); compiled 18
// times by the Makefile with -DHDB_SWEEP_T=<__half|float|double> -DHDB_SWEEP_MC=<0|1|2> -DHDB_SWEEP_KP=<32|128>
// -DHDB_SWEEP_FN=<symbol> so that the builds run in parallel.
#include "sweep_float.cuh"

namespace hdb {
int HDB_SWEEP_FN(const SweepParams& p, bool vec, int nq, int grid, size_t smem, cudaStream_t s) {
  return launch_vec<HDB_SWEEP_T, HDB_SWEEP_MC, HDB_SWEEP_KP>(p, vec, nq, grid, smem, s);
}
}  // namespace hdb
