// tmq_halo.cu -- ghost-zone packing for the sharded Dslash.
//
// For a partitioned dimension d (z or t) and an application with output parity p (input parity q = 1-p):
//   * slice x_d = 0 of `in` goes BACKWARD to rank-1, which needs it for its forward hop
//       (1 - s g_d) U_d(x) in(x+d):   we send h = (1 - s g_d) in      (link is receiver-local)
//   * slice x_d = L-1 goes FORWARD to rank+1, which needs it for its backward hop
//       (1 + s g_d) U_d(x-d)^dag in(x-d): the link lives here, so we send u = U_d^dag (1 + s g_d) in
// 12 reals per face site either way (half spinor), stored as vec[3][face].  The face index of a site is
// its (xh, y, other) lexicographic index inside the slice -- identical on sender and receiver because a
// hop along d keeps x, hence xh.  This replaces upstream dslash_pack.cu + face_buffer.cpp and the
// plug-in's own host-staged ghost exchange (lib/qudaQKXTM_Vector.cpp:172-382).
#include "tmq_internal.h"
#include "tmq_site.cuh"
#include "tmq_pack.cuh"

namespace tmq {

// NCCL path: pack into local send buffers.  blockIdx.y = 0: backward-going face (slice 0), 1: forward-going (L-1)
template <typename F, int RECON, int MU>
__global__ void __launch_bounds__(128) halo_pack_kernel(const __grid_constant__ DslashArgs<F> A, VecT<F> *send_bwd,
                                                        VecT<F> *send_fwd) {
  if (cg_iteration_is_stale(A.scal, A.cg_iter)) return;
  const int f = blockIdx.x * 128 + threadIdx.x;
  if (f >= A.g.face[MU]) return;
  const bool fwd = blockIdx.y == 1;
  pack_site<F, RECON, MU>(A, f, fwd, fwd ? send_fwd : send_bwd);
}

// Peer-memory path: ONE launch packs every partitioned dimension and stores the faces straight into the
// neighbours' ghost buffers through NVLink-mapped pointers (no staging copy, no NCCL kernel); the last CTA to
// finish publishes the application's sequence number in the neighbours' arrival flags (release, system scope).
// blockIdx.y = direction, blockIdx.z = partitioned-dimension slot.
template <typename F, int RECON>
__global__ void __launch_bounds__(128) halo_pack_p2p_kernel(const __grid_constant__ DslashArgs<F> A,
                                                            const __grid_constant__ PackDst<F> D) {
  if (cg_iteration_is_stale(A.scal, A.cg_iter)) return;
  const int slot = blockIdx.z;
  const int mu = D.dim[slot];
  const bool fwd = blockIdx.y == 1;
  const int f = blockIdx.x * 128 + threadIdx.x;
  // fused mode, after faces that were sent ahead had to be discarded: the neighbours' buffers may only be overwritten once the
  // neighbours have published the discarded sequence number (A.hw), i.e. are done reading what this launch overwrites
  if (A.hw.n > 0) {
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    if ((int)threadIdx.x < A.hw.n) {
      unsigned int v = 0;
      unsigned long long t0 = 0, now = 0;
      int spins = 0;
      for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(A.hw.flag[threadIdx.x]) : "memory");
        if ((int)(v - A.hw.seq) >= 0) break;
        if (++spins == 64) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        if (spins > 64 && (spins & 255) == 0) {
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          if (now - t0 > A.hw.timeout_ns) { *((volatile double *)A.hw.err) = 1.0; bad = 1; break; }
        }
        __nanosleep(200);
      }
    }
    __syncthreads();
    if (bad) return;
  }
  if (f < A.g.face[mu]) {
    if (mu == 3) pack_site<F, RECON, 3>(A, f, fwd, D.dst[slot][fwd ? 1 : 0]);
    else         pack_site<F, RECON, 2>(A, f, fwd, D.dst[slot][fwd ? 1 : 0]);
  }
  __threadfence_system();            // this thread's peer stores are ordered before the ticket
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int t = atomicInc(D.ticket, total - 1);
    if (t == total - 1) {
      __threadfence_system();
      for (int s = 0; s < D.nslot; s++)
        for (int d = 0; d < 2; d++)
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(D.flag[s][d]), "r"(D.seq) : "memory");
    }
  }
}

template <typename F, int RECON>
static cudaError_t pack_t(const DslashArgs<F> &A, int dim, void *sb, void *sf, cudaStream_t st) {
  dim3 grid((A.g.face[dim] + 127) / 128, 2);
  if (dim == 3) halo_pack_kernel<F, RECON, 3><<<grid, 128, 0, st>>>(A, (VecT<F> *)sb, (VecT<F> *)sf);
  else if (dim == 2) halo_pack_kernel<F, RECON, 2><<<grid, 128, 0, st>>>(A, (VecT<F> *)sb, (VecT<F> *)sf);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t halo_pack(int prec, int recon, const DslashArgs<double> *Ad, const DslashArgs<float> *As, int dim,
                      void *send_bwd, void *send_fwd, cudaStream_t st) {
  if (prec == 8)
    return recon == 8 ? pack_t<double, 8>(*Ad, dim, send_bwd, send_fwd, st)
                      : (recon == 12 ? pack_t<double, 12>(*Ad, dim, send_bwd, send_fwd, st) : pack_t<double, 18>(*Ad, dim, send_bwd, send_fwd, st));
  return recon == 8 ? pack_t<float, 8>(*As, dim, send_bwd, send_fwd, st)
                    : (recon == 12 ? pack_t<float, 12>(*As, dim, send_bwd, send_fwd, st) : pack_t<float, 18>(*As, dim, send_bwd, send_fwd, st));
}

template <typename F> static cudaError_t pack_p2p_t(int recon, const DslashArgs<F> &A, const PackDst<F> &D, cudaStream_t st) {
  int maxface = 0;
  for (int s = 0; s < D.nslot; s++) maxface = A.g.face[D.dim[s]] > maxface ? A.g.face[D.dim[s]] : maxface;
  dim3 grid((maxface + 127) / 128, 2, D.nslot);
  if (recon == 8)       halo_pack_p2p_kernel<F, 8><<<grid, 128, 0, st>>>(A, D);
  else if (recon == 12) halo_pack_p2p_kernel<F, 12><<<grid, 128, 0, st>>>(A, D);
  else                  halo_pack_p2p_kernel<F, 18><<<grid, 128, 0, st>>>(A, D);
  return cudaGetLastError();
}
cudaError_t halo_pack_p2p(int recon, const DslashArgs<double> &A, const PackDst<double> &D, cudaStream_t st) { return pack_p2p_t<double>(recon, A, D, st); }
cudaError_t halo_pack_p2p(int recon, const DslashArgs<float> &A, const PackDst<float> &D, cudaStream_t st) { return pack_p2p_t<float>(recon, A, D, st); }


// ---- CG update fused with the halo exchange of the next iteration (fused halo mode) -----------------------------------------------
// x += alpha p ; p = r + beta p, one thread per site (18 vector loads, 12 stores: the same bytes as the streaming update kernel); the
// threads on a partitioned boundary pack the new p -- the input of the next iteration's first Dslash launch -- straight into the
// neighbours' arenas, the last CTA publishes the arrival flags.  A.parity = parity of p's sites, A.pk / A.pk_dsign as in the Dslash launch.
template <typename F, int RECON>
__global__ void __launch_bounds__(128) cg_update_pack_kernel(VecT<F> *x, VecT<F> *p, const VecT<F> *__restrict__ r, const double *scal, int an, int ad,
                                                             int bn, int bd, const __grid_constant__ DslashArgs<F> A) {
  if (cg_iteration_is_stale(scal, A.cg_iter)) return;
  const Geom &g = A.g;
  const int idx = blockIdx.x * 128 + threadIdx.x;
  if (idx < g.Vh) {
    const F alpha = (F)(scal[an] / scal[ad]), beta = (F)(scal[bn] / scal[bd]);
    Spinor<F> ps;
#pragma unroll
    for (int j = 0; j < 6; j++) {
      const size_t o = (size_t)j * g.Vh + idx;
      VecT<F> xv = x[o], pv = p[o];
      const VecT<F> rv = r[o];
      xv.a += alpha * pv.a; xv.b += alpha * pv.b; xv.c += alpha * pv.c; xv.d += alpha * pv.d;
      pv.a = rv.a + beta * pv.a; pv.b = rv.b + beta * pv.b; pv.c = rv.c + beta * pv.c; pv.d = rv.d + beta * pv.d;
      x[o] = xv; p[o] = pv;
      unpack_vec(ps, j, pv);
    }
    const int plane = g.X[1] * g.Xh, vol3 = plane * g.X[2];
    const int t = idx / vol3, rem = idx - t * vol3, z = rem / plane, rem2 = rem - z * plane, y = rem2 / g.Xh, xh = rem2 - y * g.Xh;
    if (on_packed_boundary(A, z, t)) pack_spinor<F, RECON>(A, ps, idx, xh, y, z, t);
  }
  if (A.pk_on == 2) return;              // halo mode 4: the faces went into this rank's own send buffers; the copy engines publish
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicInc(A.pk.ticket, gridDim.x - 1);
    if (t == gridDim.x - 1) {
      __threadfence_system();
      for (int s = 0; s < A.pk.nslot; s++)
        for (int d = 0; d < 2; d++)
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(A.pk.flag[s][d]), "r"(A.pk.seq) : "memory");
    }
  }
}
template <typename F>
static cudaError_t cg_update_pack_t(int recon, void *x, void *p, const void *r, const double *scal, int an, int ad, int bn, int bd, const DslashArgs<F> &A, cudaStream_t st) {
  const int grid = (A.g.Vh + 127) / 128;
  if (recon == 8)       cg_update_pack_kernel<F, 8><<<grid, 128, 0, st>>>((VecT<F> *)x, (VecT<F> *)p, (const VecT<F> *)r, scal, an, ad, bn, bd, A);
  else if (recon == 12) cg_update_pack_kernel<F, 12><<<grid, 128, 0, st>>>((VecT<F> *)x, (VecT<F> *)p, (const VecT<F> *)r, scal, an, ad, bn, bd, A);
  else                  cg_update_pack_kernel<F, 18><<<grid, 128, 0, st>>>((VecT<F> *)x, (VecT<F> *)p, (const VecT<F> *)r, scal, an, ad, bn, bd, A);
  return cudaGetLastError();
}
cudaError_t cg_update_pack(int recon, void *x, void *p, const void *r, const double *scal, int an, int ad, int bn, int bd, const DslashArgs<double> &A, cudaStream_t st) {
  return cg_update_pack_t<double>(recon, x, p, r, scal, an, ad, bn, bd, A, st);
}
cudaError_t cg_update_pack(int recon, void *x, void *p, const void *r, const double *scal, int an, int ad, int bn, int bd, const DslashArgs<float> &A, cudaStream_t st) {
  return cg_update_pack_t<float>(recon, x, p, r, scal, an, ad, bn, bd, A, st);
}

// ---- scalar all-reduce over peer memory -------------------------------------------------------------------------
__global__ void __launch_bounds__(32) p2p_allreduce_kernel(const __grid_constant__ P2PRed R) {
  if (cg_iteration_is_stale(R.scal, R.cg_iter)) return;
  const int t = threadIdx.x;
  const int buf = (int)(R.seq & 1u);
  if (t < R.nranks) {
    double *dst = R.mbox[t] + (size_t)buf * 4 * TMQ_MAX_RANKS;
    for (int j = 0; j < R.n; j++) dst[j * TMQ_MAX_RANKS + R.rank] = R.scal[R.slot + j];
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(R.mflag[t] + buf * TMQ_MAX_RANKS + R.rank), "r"(R.seq) : "memory");
    // wait for rank t's contribution to arrive in our own mailbox
    const unsigned int *f = R.mflag[R.rank] + buf * TMQ_MAX_RANKS + t;
    unsigned int v = 0;
    unsigned long long t0 = 0, now = 0;
    int spins = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if ((int)(v - R.seq) >= 0) break;
      if (++spins == 64) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      if (spins > 64 && (spins & 255) == 0) {              // wall-clock bound (see halo_wait in tmq_dslash_inst.cuh)
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > R.timeout_ns) { *((volatile double *)R.err) = 1.0; break; }
      }
      __nanosleep(spins < 64 ? 50 : 500);
    }
  }
  __syncwarp();
  if (t < R.n) {
    const volatile double *src = R.mbox[R.rank] + (size_t)buf * 4 * TMQ_MAX_RANKS + t * TMQ_MAX_RANKS;
    double s = 0.0;
    for (int r = 0; r < R.nranks; r++) s += src[r];      // fixed rank order: identical bits on every rank
    R.scal[R.slot + t] = s;
    // the CG's stopping test on the global |r|^2: every rank has the same bits, so every rank takes the same decision
    if (t == 0 && R.cg_stop && R.cg_iter > 0 && s <= R.scal[SC_STOP] && R.scal[SC_DONE] == 0.0) R.scal[SC_DONE] = (double)R.cg_iter;
  }
}
cudaError_t p2p_allreduce(const P2PRed &R, cudaStream_t st) {
  p2p_allreduce_kernel<<<1, 32, 0, st>>>(R);
  return cudaGetLastError();
}

}  // namespace tmq
