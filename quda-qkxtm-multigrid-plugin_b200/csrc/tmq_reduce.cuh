// tmq_reduce.cuh -- deterministic grid reductions: warp-shuffle tree -> shared memory -> one partial per
// CTA -> the last CTA to finish (atomic ticket) sums the partials in a fixed order and publishes the
// result in the device scalar block.  The summation order depends only on the launch shape, so results
// are bit-reproducible run to run (needed for the CG iteration-count gate).
#pragma once
#include "tmq_types.h"

namespace tmq {

// device scalar block (doubles).  The CG keeps its recurrence scalars here so that no kernel waits on the host.
enum {
  SC_R2_0 = 0, SC_R2_1 = 1,   // |r|^2, double-buffered by iteration parity
  SC_PAP = 2,                 // <p, A p> = |M p|^2
  SC_B2 = 3,
  SC_T0 = 4, SC_T1 = 5, SC_T2 = 6, SC_T3 = 7,   // generic reduction results (norm2, dots)
  SC_ONE = 8, SC_ZERO = 9,
  SC_ERR = 10,                // raised by a Dslash kernel whose halo wait timed out
  SC_BARRIER = 11,            // scratch of tmq_barrier's all-reduce
  SC_STOP = 12,               // CG: tol^2 |b|^2, the device-side copy of the stopping threshold
  SC_DONE = 13,               // CG: 0 while iterating, else the (1-based) iteration whose residual met SC_STOP: launches of later iterations exit at once
  SC_COUNT = 16
};

#if defined(__CUDACC__)
// launches of a CG iteration that the host enqueued before it knew that an earlier iteration had converged
__device__ __forceinline__ bool cg_iteration_is_stale(const double *scal, int cg_iter) {
  if (cg_iter <= 0) return false;
  const double d = *((const volatile double *)(scal + SC_DONE));
  return d != 0.0 && (double)cg_iter > d;
}
template <int N>
__device__ __forceinline__ void block_reduce_finalize(double (&v)[N], double *partials, unsigned int *ticket,
                                                      double *scal, int slot0, bool accum = false, int cg_iter_done = 0) {
  __shared__ double sm[32 * N];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int j = 0; j < N; j++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_down_sync(0xffffffffu, v[j], o);
    if (lane == 0) sm[warp * N + j] = v[j];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int j = 0; j < N; j++) {
      double s = (lane < nwarp) ? sm[lane * N + j] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if (lane == 0) partials[(size_t)blockIdx.x * N + j] = s;
    }
    if (lane == 0) {
      __threadfence();
      unsigned int t = atomicInc(ticket, gridDim.x - 1);   // wraps back to 0 on the last CTA
      is_last = (t == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double acc[N];
#pragma unroll
  for (int j = 0; j < N; j++) acc[j] = 0.0;
  for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
    for (int j = 0; j < N; j++) acc[j] += __ldcg(&partials[(size_t)b * N + j]);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < N; j++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_down_sync(0xffffffffu, acc[j], o);
    if (lane == 0) sm[warp * N + j] = acc[j];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int j = 0; j < N; j++) {
      double s = (lane < nwarp) ? sm[lane * N + j] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if (lane == 0) {
        const double tot = accum ? scal[slot0 + j] + s : s;
        scal[slot0 + j] = tot;
        // the CG's stopping test on the device (single rank: this IS the global sum): later iterations that the host has
        // already enqueued find SC_DONE set and exit
        if (j == 0 && cg_iter_done > 0 && tot <= scal[SC_STOP] && scal[SC_DONE] == 0.0) scal[SC_DONE] = (double)cg_iter_done;
      }
    }
  }
}
#endif

}  // namespace tmq
