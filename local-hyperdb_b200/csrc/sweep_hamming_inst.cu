// One (candidate class, metric) instantiation of the bit-packed sweeps (csrc/sweep_hamming.cuh); compiled four times by the
// Makefile with -DHDB_HAM_KP=<32|128> -DHDB_HAM_JAC=<0|1> -DHDB_HAM_FN=<symbol>.
#include "sweep_hamming.cuh"

namespace hdb {
int HDB_HAM_FN(const HammingParams& hp, int nq, int grid, int form, size_t smem, cudaStream_t s) {
  if (form == 0) return launch_hamming_staged<HDB_HAM_KP>(hp, nq, grid, s);
  if (form == 1) return launch_hamming_kp<HDB_HAM_KP>(hp, nq, grid, s);
  auto kern = sweep_hamming_kernel<HDB_HAM_KP>;
  if (smem > 48 * 1024) HDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kSweepThreads, smem, s>>>(hp);
  HDB_LAUNCHED();
  HDB_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace hdb
