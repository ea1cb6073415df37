// tmq_pack.cuh -- ghost-face packing shared by the stand-alone pack kernels (tmq_halo.cu) and by the fused Dslash launch, whose
// boundary CTAs pack the faces of their OWN OUTPUT for the next application and store them straight into the neighbours' ghost
// arenas over NVLink (tmq_dslash_inst.cuh): compute and halo exchange in one kernel, no pack launch, no copy-engine transfer.
#pragma once
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

// one face site: project (and for the forward-going face multiply by U^dag) and store into `dst`
template <typename F, int RECON, int MU>
__device__ __forceinline__ void pack_site(const DslashArgs<F> &A, int f, bool fwd, VecT<F> *dst) {
  const Geom &g = A.g;
  const int face = g.face[MU];
  const int q = 1 - A.parity;
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
    store_half(dst, f, face, h);
  } else {
    project_any<F, MU>(h, p, -A.dsign);         // receiver's backward hop: 1 + s g
    Link<F> L;
    const F s12 = (MU == 3 && g.tb_last) ? (F)g.tb_sign : (F)1;
    load_link<F, RECON>(L, A.gauge, q, MU, idx, g.Vh, s12);
    Half<F> u;
    su3_apply<F, true>(u, L, h);
    store_half(dst, f, face, u);
  }
}


// Fused path: a thread that holds the complete spinor p of site (xh, y, z, t; cb index idx) of the field the NEXT application reads (site
// parity A.parity, projector sign A.pk_dsign) packs it as a face site of every partitioned dimension it lies on the boundary of.  Slice 0
// goes backward ((1 - s g) psi), slice L-1 goes forward (U^dag (1 + s g) psi, the link lives here).
template <typename F, int RECON>
__device__ __forceinline__ void pack_spinor(const DslashArgs<F> &A, const Spinor<F> &p, int idx, int xh, int y, int z, int t) {
  const Geom &g = A.g;
#pragma unroll
  for (int s = 0; s < 2; s++) {
    if (s >= A.pk.nslot) break;
    const int mu = A.pk.dim[s];
    const int cm = mu == 3 ? t : z, L = g.X[mu], face = g.face[mu];
    const int f = mu == 3 ? idx - cm * face : (t * g.X[1] + y) * g.Xh + xh;
    if (cm == 0) {
      Half<F> h;
      if (mu == 3) project_any<F, 3>(h, p, A.pk_dsign); else project_any<F, 2>(h, p, A.pk_dsign);
      store_half(A.pk.dst[s][0], f, face, h);
    }
    if (cm == L - 1) {
      Half<F> h, u;
      if (mu == 3) project_any<F, 3>(h, p, -A.pk_dsign); else project_any<F, 2>(h, p, -A.pk_dsign);
      Link<F> Lk;
      const F s12 = (mu == 3 && g.tb_last) ? (F)g.tb_sign : (F)1;
      load_link<F, RECON>(Lk, A.gauge, A.parity, mu, idx, g.Vh, s12);
      su3_apply<F, true>(u, Lk, h);
      store_half(A.pk.dst[s][1], f, face, u);
    }
  }
}
template <typename F> __device__ __forceinline__ bool on_packed_boundary(const DslashArgs<F> &A, int z, int t) {
  bool any = false;
#pragma unroll
  for (int s = 0; s < 2; s++) {
    if (s >= A.pk.nslot) break;
    const int mu = A.pk.dim[s], cm = mu == 3 ? t : z;
    any = any || cm == 0 || cm == A.g.X[mu] - 1;
  }
  return any;
}
// the thread that has just stored output site `e` of a Dslash launch: the spinor is re-read from A.out (the same thread stored it a few
// instructions ago, so the loads are served by the store queue / L2 and cost no HBM traffic)
template <typename F, int RECON>
__device__ __forceinline__ void pack_out(const DslashArgs<F> &A, const Enum &en, uint32_t e) {
  const Geom &g = A.g;
  const SiteCoord c = decode_site(g, en, A.parity, e);
  if (!on_packed_boundary(A, c.z, c.t)) return;
  Spinor<F> p;
#pragma unroll
  for (int j = 0; j < 6; j++) unpack_vec(p, j, A.out[(size_t)j * g.Vh + c.idx]);
  pack_spinor<F, RECON>(A, p, c.idx, c.xh, c.y, c.z, c.t);
}

}  // namespace tmq
