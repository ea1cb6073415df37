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

namespace tmq {

template <typename F> __device__ __forceinline__ void store_half(VecT<F> *base, int f, int fstride, const Half<F> &h) {
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const int k0 = 2 * j, k1 = 2 * j + 1;
    VecT<F> v;
    v.a = h.h[k0 / 3][k0 % 3][0]; v.b = h.h[k0 / 3][k0 % 3][1];
    v.c = h.h[k1 / 3][k1 % 3][0]; v.d = h.h[k1 / 3][k1 % 3][1];
    base[(size_t)j * fstride + f] = v;
  }
}

template <typename F, int MU> __device__ __forceinline__ void project_any(Half<F> &h, const Spinor<F> &p, F sg) {
  if constexpr (MU < 3) project<F, MU>(h, p, sg);
  else {
    const int o = sg > (F)0 ? 2 : 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      h.h[0][c][0] = 2 * (o ? p.v[2][c][0] : p.v[0][c][0]); h.h[0][c][1] = 2 * (o ? p.v[2][c][1] : p.v[0][c][1]);
      h.h[1][c][0] = 2 * (o ? p.v[3][c][0] : p.v[1][c][0]); h.h[1][c][1] = 2 * (o ? p.v[3][c][1] : p.v[1][c][1]);
    }
  }
}

// blockIdx.y = 0: backward-going face (slice 0), 1: forward-going face (slice L-1)
template <typename F, int RECON, int MU>
__global__ void __launch_bounds__(128) halo_pack_kernel(const __grid_constant__ DslashArgs<F> A, VecT<F> *send_bwd,
                                                        VecT<F> *send_fwd) {
  const Geom &g = A.g;
  const int face = g.face[MU];
  const int f = blockIdx.x * 128 + threadIdx.x;
  if (f >= face) return;
  const int q = 1 - A.parity;
  const bool fwd = blockIdx.y == 1;
  const int slice = fwd ? g.X[MU] - 1 : 0;
  int idx;
  if (MU == 3) idx = slice * face + f;
  else {   // MU == 2: f = (t*Y + y)*Xh + xh
    const int plane = g.X[1] * g.Xh;
    const int t = f / plane, rem = f - t * plane;
    idx = (t * g.X[2] + slice) * plane + rem;
  }
  Spinor<F> p;
  load_spinor(p, A.in, idx, g.Vh);
  Half<F> h;
  if (!fwd) {
    project_any<F, MU>(h, p, A.dsign);          // receiver's forward hop: 1 - s g
    store_half(send_bwd, f, face, h);
  } else {
    project_any<F, MU>(h, p, -A.dsign);         // receiver's backward hop: 1 + s g
    Link<F> L;
    const F s12 = (MU == 3 && g.tb_last) ? (F)g.tb_sign : (F)1;
    load_link<F, RECON>(L, A.gauge, q, MU, idx, g.Vh, s12);
    Half<F> u;
    su3_apply<F, true>(u, L, h);
    store_half(send_fwd, f, face, u);
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
  if (prec == 8) return recon == 12 ? pack_t<double, 12>(*Ad, dim, send_bwd, send_fwd, st) : pack_t<double, 18>(*Ad, dim, send_bwd, send_fwd, st);
  return recon == 12 ? pack_t<float, 12>(*As, dim, send_bwd, send_fwd, st) : pack_t<float, 18>(*As, dim, send_bwd, send_fwd, st);
}

}  // namespace tmq
