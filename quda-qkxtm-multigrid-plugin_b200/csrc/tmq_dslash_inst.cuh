// tmq_dslash_inst.cuh -- the Dslash kernel and its launcher, instantiated once per (precision, recon)
// translation unit (dslash_d12.cu, ...).  One thread per output parity site; a CTA owns a compact 4-d
// block of sites (Enum in tmq_types.h).  HBM-bound: no tensor cores.
#pragma once
#include <cuda_runtime.h>
#include "tmq_site.cuh"
#include "tmq_pack.cuh"
#include "tmq_reduce.cuh"

namespace tmq {

#ifndef TMQ_DSLASH_BLOCK
#define TMQ_DSLASH_BLOCK 128
#endif

// CTAs per SM.  fp64: every epilogue fits the 128-register budget of 4 resident CTAs (16 warps) with a 12-36 byte spill;
// the fourth CTA hides the recon-12 / SU(3) fp64 latency and is worth 4-10% (A/B: profiles/r02_variant_minblocks.log,
// r03).  What used to push the x-term epilogues to 250-400 bytes of spill was the optional L2-prefetch code path
// (measured: no gain), now removed; the epilogue itself runs in three (vector j, j+3) pairs to stay small.
#ifndef TMQ_MINBLOCKS_D
#define TMQ_MINBLOCKS_D 4
#endif
#ifndef TMQ_MINBLOCKS_D_LIGHT
#define TMQ_MINBLOCKS_D_LIGHT 4
#endif
template <typename F, int EPI, int RECON> struct MinBlocks { static constexpr int v = (EPI == EPI_PLAIN || EPI == EPI_TW) ? TMQ_MINBLOCKS_D_LIGHT : TMQ_MINBLOCKS_D; };
// fp32: 7 CTAs/SM (72 registers, 4-24 byte spills) beat 6 by 4-5% with recon-12 and lose 1% with recon-18; 8 CTAs (64
// registers, ~100 byte spills) lose everywhere (A/B: profiles/r14_variant_fp32.log)
#ifndef TMQ_MINBLOCKS_S
#define TMQ_MINBLOCKS_S 7
#endif
template <int EPI, int RECON> struct MinBlocks<float, EPI, RECON> { static constexpr int v = RECON == 12 ? TMQ_MINBLOCKS_S : 6; };

// Boundary CTAs of a fused sharded launch: wait until every neighbour has published this application's
// sequence number (its faces have landed in our ghost buffers over NVLink).  The wait is bounded by WALL-CLOCK time
// (%globaltimer; TMQ_OPT_HALO_TIMEOUT_MS, default 120 s: ordinary rank skew -- file I/O on one rank, a slow first NCCL init --
// must not trip it, only a dead neighbour).  On a timeout the device error scalar is raised and the CTA returns false: its
// sites are NOT computed from stale ghosts (the caller still takes part in the grid reduction), and the host fails the call.
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool halo_wait(const HaloWait &hw) {
  __shared__ int timed_out;
  if (threadIdx.x == 0) timed_out = 0;
  __syncthreads();
  if ((int)threadIdx.x < hw.n) {
    const unsigned int *f = hw.flag[threadIdx.x];
    unsigned int v = 0;
    unsigned long long t0 = 0;
    int spins = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (hw.exact ? (v == hw.seq) : ((int)(v - hw.seq) >= 0)) break;
      if (++spins == 64) t0 = global_ns();                 // the clock is only read once the wait is not instantaneous
      if (spins > 64 && (spins & 255) == 0 && global_ns() - t0 > hw.timeout_ns) { *((volatile double *)hw.err) = 1.0; timed_out = 1; break; }
      __nanosleep(spins < 64 ? 100 : 500);
    }
  }
  __syncthreads();
  return timed_out == 0;
}

// twisted-clover variants: the site matrices add 36 vector loads per application and a second spinor in the x-term
// epilogues; one CTA less per SM keeps them out of local memory
template <typename F> struct MinBlocksClover { static constexpr int v = 3; };
template <> struct MinBlocksClover<float> { static constexpr int v = 5; };

template <typename F, int RECON, int EPI, bool MULTI, bool CLOVER = false>
__global__ void __launch_bounds__(TMQ_DSLASH_BLOCK, (CLOVER ? MinBlocksClover<F>::v : MinBlocks<F, EPI, RECON>::v))
dslash_kernel(const __grid_constant__ DslashArgs<F> A) {
  if (cg_iteration_is_stale(A.scal, A.cg_iter)) return;      // block-uniform
  uint32_t blk = blockIdx.x;
  const Enum *en = &A.en;
  bool boundary = MULTI && A.all_boundary;
  bool ghosts_ok = true;
  if (MULTI && blk >= (uint32_t)A.npre) {
    // fused sharded launch: the first npre interior CTAs overlap the halo transfer; the boundary CTAs come next
    // (they wait on the arrival flags, normally already set) and the remaining interior CTAs keep the SMs busy
    // behind them, so the cold boundary work never forms the tail of the launch
    const uint32_t nb = (uint32_t)(A.nblk[1] + A.nblk[2]);
    if (blk < (uint32_t)A.npre + nb) {
      blk -= (uint32_t)A.npre;
      if (blk < (uint32_t)A.nblk[1]) en = &A.en_b[0];
      else { blk -= (uint32_t)A.nblk[1]; en = &A.en_b[1]; }
      boundary = true;
      if (A.hw.n > 0) ghosts_ok = halo_wait(A.hw);
    } else {
      blk -= nb;
    }
  }
  const uint32_t e = blk * TMQ_DSLASH_BLOCK + threadIdx.x;
  F alpha = 0;
  if (EpiTraits<EPI>::RED == 2) alpha = (F)(A.scal[A.alpha_num] / A.scal[A.alpha_den]);
  double red[1] = {0.0};
  if (e < (uint32_t)en->nsites && ghosts_ok) {
    // interior sites never touch a ghost zone: they run the branch-free single-GPU body (the ghost-aware body
    // costs ~8% because its conditional loads cannot be hoisted); only boundary CTAs pay for it
    if (MULTI && boundary) red[0] = dslash_site<F, RECON, EPI, true, CLOVER>(A, *en, e, alpha);
    else                   red[0] = dslash_site<F, RECON, EPI, false, CLOVER>(A, *en, e, alpha);
  }
  if constexpr (MULTI && EPI != EPI_CG4) {
    // fused halo exchange: pack this launch's output faces for the next application into the neighbours' arenas (block-uniform branch)
    if (boundary && A.pk_on) {
      if (e < (uint32_t)en->nsites && ghosts_ok) pack_out<F, RECON>(A, *en, e);
      if (A.pk_on == 1) {                  // peer stores (halo mode 3); mode 4 packs into local send buffers and the copy engines publish
        __threadfence_system();            // this thread's peer stores are ordered before the ticket
        __syncthreads();
        if (threadIdx.x == 0) {
          const unsigned int nb = (unsigned int)(A.nblk[1] + A.nblk[2]);
          const unsigned int t = atomicInc(A.pk.ticket, nb - 1);
          if (t == nb - 1) {
            __threadfence_system();
            for (int s = 0; s < A.pk.nslot; s++)
              for (int d = 0; d < 2; d++)
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(A.pk.flag[s][d]), "r"(A.pk.seq) : "memory");
          }
        }
      }
    }
  }
  if (EpiTraits<EPI>::RED != 0)
    block_reduce_finalize<1>(red, A.partials, A.ticket, A.scal, A.red_slot, A.red_accum != 0, (EPI == EPI_CG4 && A.cg_local_stop) ? A.cg_iter : 0);
}

template <typename F, int RECON, int EPI>
static cudaError_t launch_epi(bool multi, const DslashArgs<F> &A, cudaStream_t st) {
  // callers fill nblk[] (segment sizes in CTAs); a plain launch has nblk = {ceil(nsites/128), 0, 0}
  const int grid = A.nblk[0] + A.nblk[1] + A.nblk[2];
  if (grid == 0) return cudaSuccess;
  // twisted-clover: only the epilogues that apply a site matrix have a clover variant
  constexpr bool HAS_SITE_OP = EpiTraits<EPI>::TW1 || EpiTraits<EPI>::TW3 || EpiTraits<EPI>::TWX;
  if (A.cl_inv != nullptr && HAS_SITE_OP) {
    if constexpr (HAS_SITE_OP) {
      if (multi) dslash_kernel<F, RECON, EPI, true, true><<<grid, TMQ_DSLASH_BLOCK, 0, st>>>(A);
      else       dslash_kernel<F, RECON, EPI, false, true><<<grid, TMQ_DSLASH_BLOCK, 0, st>>>(A);
    }
    return cudaGetLastError();
  }
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
    case EPI_CHEB:     return launch_epi<F, RECON, EPI_CHEB>(multi, A, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace tmq
