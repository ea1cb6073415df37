// Dslash kernels, precision = double, gauge reconstruct = 18 (see tmq_dslash_inst.cuh)
#include "tmq_dslash_inst.cuh"
namespace tmq {
cudaError_t launch_dslash_d18(int epi, bool multi, const DslashArgs<double> &A, cudaStream_t st) {
  return launch_dslash_t<double, 18>(epi, multi, A, st);
}
}  // namespace tmq
