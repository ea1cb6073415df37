// Dslash kernels, precision = float, gauge reconstruct = 12 (see tmq_dslash_inst.cuh)
#include "tmq_dslash_inst.cuh"
namespace tmq {
cudaError_t launch_dslash_s12(int epi, bool multi, const DslashArgs<float> &A, cudaStream_t st) {
  return launch_dslash_t<float, 12>(epi, multi, A, st);
}
}  // namespace tmq
