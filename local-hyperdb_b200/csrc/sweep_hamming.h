// Host entry points of csrc/sweep_hamming.cu.
#pragma once
#include "hdb_internal.h"

namespace hdb {
int hamming_sweep_max_group(const MatrixView& m, int kp);
int launch_hamming_sweep(const MatrixView& m, int metric, const uint32_t* qbits, const RowFilter& f, int kp, const SweepOut& out, int nq,
                         cudaStream_t s);
}  // namespace hdb
