// Dslash kernels, precision = double, gauge reconstruct = 12 (see tmq_dslash_inst.cuh)
#include "tmq_dslash_inst.cuh"
namespace tmq {
cudaError_t launch_dslash_d12(int epi, bool multi, const DslashArgs<double> &A, cudaStream_t st) {
  return launch_dslash_t<double, 12>(epi, multi, A, st);
}
}  // namespace tmq
