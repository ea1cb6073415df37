// tmq_dslash_inst.cuh -- the Dslash kernel and its launcher, instantiated once per (precision, recon)
// translation unit (dslash_d12.cu, ...).  One thread per output parity site; a CTA owns a compact 4-d
// block of sites (Enum in tmq_types.h).  HBM-bound: no tensor cores.
#pragma once
#include <cuda_runtime.h>
#include "tmq_site.cuh"
#include "tmq_reduce.cuh"

namespace tmq {

#ifndef TMQ_DSLASH_BLOCK
#define TMQ_DSLASH_BLOCK 128
#endif

template <typename F> struct MinBlocks { static constexpr int v = 3; };
template <> struct MinBlocks<float> { static constexpr int v = 6; };

template <typename F, int RECON, int EPI, bool MULTI>
__global__ void __launch_bounds__(TMQ_DSLASH_BLOCK, MinBlocks<F>::v)
dslash_kernel(const __grid_constant__ DslashArgs<F> A) {
  const uint32_t e = blockIdx.x * TMQ_DSLASH_BLOCK + threadIdx.x;
  F alpha = 0;
  if (EpiTraits<EPI>::RED == 2) alpha = (F)(A.scal[A.alpha_num] / A.scal[A.alpha_den]);
  double red[1] = {0.0};
  if (e < (uint32_t)A.en.nsites) red[0] = dslash_site<F, RECON, EPI, MULTI>(A, e, alpha);
  if (EpiTraits<EPI>::RED != 0) block_reduce_finalize<1>(red, A.partials, A.ticket, A.scal, A.red_slot, A.red_accum != 0);
}

template <typename F, int RECON, int EPI>
static cudaError_t launch_epi(bool multi, const DslashArgs<F> &A, cudaStream_t st) {
  const int grid = (A.en.nsites + TMQ_DSLASH_BLOCK - 1) / TMQ_DSLASH_BLOCK;
  if (grid == 0) return cudaSuccess;
  if (multi) dslash_kernel<F, RECON, EPI, true><<<grid, TMQ_DSLASH_BLOCK, 0, st>>>(A);
  else       dslash_kernel<F, RECON, EPI, false><<<grid, TMQ_DSLASH_BLOCK, 0, st>>>(A);
  return cudaGetLastError();
}

template <typename F, int RECON>
static cudaError_t launch_dslash_t(int epi, bool multi, const DslashArgs<F> &A, cudaStream_t st) {
  switch (epi) {
    case EPI_PLAIN:    return launch_epi<F, RECON, EPI_PLAIN>(multi, A, st);
    case EPI_TW:       return launch_epi<F, RECON, EPI_TW>(multi, A, st);
    case EPI_TW_XPAY:  return launch_epi<F, RECON, EPI_TW_XPAY>(multi, A, st);
    case EPI_XPAY:     return launch_epi<F, RECON, EPI_XPAY>(multi, A, st);
    case EPI_XPAY_TW3: return launch_epi<F, RECON, EPI_XPAY_TW3>(multi, A, st);
    case EPI_MDAGM2:   return launch_epi<F, RECON, EPI_MDAGM2>(multi, A, st);
    case EPI_TWX_XPAY: return launch_epi<F, RECON, EPI_TWX_XPAY>(multi, A, st);
    case EPI_CG4:      return launch_epi<F, RECON, EPI_CG4>(multi, A, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace tmq
