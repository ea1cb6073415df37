// Dslash kernels, precision = float, gauge reconstruct = 8 (see tmq_dslash_inst.cuh, tmq_site.cuh: reconstruct_from8)
#include "tmq_dslash_inst.cuh"
namespace tmq {
cudaError_t launch_dslash_s8(int epi, bool multi, const DslashArgs<float> &A, cudaStream_t st) {
  return launch_dslash_t<float, 8>(epi, multi, A, st);
}
}  // namespace tmq
