// Dslash kernels, precision = float, gauge reconstruct = 18 (see tmq_dslash_inst.cuh)
#include "tmq_dslash_inst.cuh"
namespace tmq {
cudaError_t launch_dslash_s18(int epi, bool multi, const DslashArgs<float> &A, cudaStream_t st) {
  return launch_dslash_t<float, 18>(epi, multi, A, st);
}
}  // namespace tmq
