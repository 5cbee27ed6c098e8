// One (storage type, metric class) instantiation of the TMA-staged sweep (csrc/sweep_staged.cuh); compiled six times by
// the Makefile with -DHDB_SWEEP_T=<__half|float> -DHDB_SWEEP_MC=<0|1|2> -DHDB_SWEEP_FN=<symbol>.
#include "sweep_staged.cuh"

namespace hdb {
int HDB_SWEEP_FN(const SweepParams& p, int kp, int grid, cudaStream_t s) {
  return launch_staged<HDB_SWEEP_T, HDB_SWEEP_MC>(p, kp, grid, s);
}
}  // namespace hdb
