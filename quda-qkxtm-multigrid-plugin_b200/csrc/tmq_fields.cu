// tmq_fields.cu -- layout converters between the reference's host/device orders and the native layout,
// plus the plaquette sanity check and the QKXTM container kernels.  One-off or per-RHS work, all
// HBM-streaming.
//   gauge  : host QDP even-odd [parity][cb][3][3][2] double (qkxtm/QKXTM_util.cpp:840-857)  -> native
//   spinor : QKXTM device SoA d[(s*3+c)*V + x_lex] complex (lib/qudaQKXTM_Vector.cpp:72-81) <-> native,
//            replacing uploadToCuda_core.h / downloadFromCuda_core.h / scaleVector_core.h
//   spinor : host even-odd AoS [cb][s][c][2] double (upstream host reference order)          <-> native
#include "tmq_internal.h"
#include "tmq_site.cuh"

namespace tmq {

constexpr int FB = 256;

// ---- gauge ---------------------------------------------------------------------------------------------
template <typename F, int RECON>
__global__ void gauge_reorder_kernel(void *dst, const double *__restrict__ src, int mu, int Vh) {
  // src: [2][Vh][18] doubles for direction mu
  const int i = blockIdx.x * FB + threadIdx.x;
  if (i >= 2 * Vh) return;
  const int parity = i / Vh, idx = i - parity * Vh;
  const double *s = src + (size_t)i * 18;
  if (RECON == 8) {
    // (U01, U02), (U10, tan(arg U00 / 4), tan(arg U20 / 4)); the phases are taken in double whatever F is
    VecT<F> *b = (VecT<F> *)dst + (size_t)((parity * 4 + mu) * 2) * (size_t)Vh + idx;
    VecT<F> v; v.a = (F)s[2]; v.b = (F)s[3]; v.c = (F)s[4]; v.d = (F)s[5];
    b[0] = v;
    v.a = (F)s[6]; v.b = (F)s[7]; v.c = (F)tan(0.25 * atan2(s[1], s[0])); v.d = (F)tan(0.25 * atan2(s[13], s[12]));
    b[(size_t)Vh] = v;
  } else if (RECON == 12) {
    VecT<F> *b = (VecT<F> *)dst + (size_t)((parity * 4 + mu) * 3) * (size_t)Vh + idx;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      VecT<F> v; v.a = (F)s[4 * j]; v.b = (F)s[4 * j + 1]; v.c = (F)s[4 * j + 2]; v.d = (F)s[4 * j + 3];
      b[(size_t)j * Vh] = v;
    }
  } else {
    CplxT<F> *b = (CplxT<F> *)dst + (size_t)((parity * 4 + mu) * 9) * (size_t)Vh + idx;
#pragma unroll
    for (int k = 0; k < 9; k++) { CplxT<F> c; c.re = (F)s[2 * k]; c.im = (F)s[2 * k + 1]; b[(size_t)k * Vh] = c; }
  }
}

cudaError_t gauge_reorder(int prec, int recon, void *dst, const double *src_mu, int mu, int Vh, cudaStream_t st) {
  const int grid = (2 * Vh + FB - 1) / FB;
  if (prec == 8) {
    if (recon == 8)       gauge_reorder_kernel<double, 8><<<grid, FB, 0, st>>>(dst, src_mu, mu, Vh);
    else if (recon == 12) gauge_reorder_kernel<double, 12><<<grid, FB, 0, st>>>(dst, src_mu, mu, Vh);
    else                  gauge_reorder_kernel<double, 18><<<grid, FB, 0, st>>>(dst, src_mu, mu, Vh);
  } else {
    if (recon == 8)       gauge_reorder_kernel<float, 8><<<grid, FB, 0, st>>>(dst, src_mu, mu, Vh);
    else if (recon == 12) gauge_reorder_kernel<float, 12><<<grid, FB, 0, st>>>(dst, src_mu, mu, Vh);
    else                  gauge_reorder_kernel<float, 18><<<grid, FB, 0, st>>>(dst, src_mu, mu, Vh);
  }
  return cudaGetLastError();
}

// ---- spinor: QKXTM SoA <-> native -----------------------------------------------------------------------
// One thread per pair of lexicographic sites (2*sid, 2*sid+1): one of them is even, the other odd
// (same decomposition as uploadToCuda_core.h:7-20), so each thread moves 32 contiguous bytes per
// component on the QKXTM side and one cb site per parity on the native side.
template <typename F, typename Q>
__global__ void from_qkxtm_kernel(VecT<F> *even, VecT<F> *odd, const CplxT<Q> *__restrict__ qk, Geom g) {
  const int sid = blockIdx.x * FB + threadIdx.x;
  if (sid >= g.Vh) return;
  const int row = sid / g.Xh;
  const int y = row % g.X[1], z = (row / g.X[1]) % g.X[2], t = row / (g.X[1] * g.X[2]);
  const int oddFirst = (y + z + t) & 1;   // parity of lexicographic site 2*sid
  const size_t V = (size_t)2 * g.Vh;
  VecT<F> *dst0 = oddFirst ? odd : even, *dst1 = oddFirst ? even : odd;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    CplxT<Q> a0 = qk[(size_t)(2 * j) * V + 2 * (size_t)sid], a1 = qk[(size_t)(2 * j) * V + 2 * (size_t)sid + 1];
    CplxT<Q> b0 = qk[(size_t)(2 * j + 1) * V + 2 * (size_t)sid], b1 = qk[(size_t)(2 * j + 1) * V + 2 * (size_t)sid + 1];
    if (dst0) { VecT<F> v; v.a = (F)a0.re; v.b = (F)a0.im; v.c = (F)b0.re; v.d = (F)b0.im; dst0[(size_t)j * g.Vh + sid] = v; }
    if (dst1) { VecT<F> v; v.a = (F)a1.re; v.b = (F)a1.im; v.c = (F)b1.re; v.d = (F)b1.im; dst1[(size_t)j * g.Vh + sid] = v; }
  }
}

template <typename F, typename Q>
__global__ void to_qkxtm_kernel(CplxT<Q> *qk, const VecT<F> *__restrict__ even, const VecT<F> *__restrict__ odd,
                                F scale, Geom g) {
  const int sid = blockIdx.x * FB + threadIdx.x;
  if (sid >= g.Vh) return;
  const int row = sid / g.Xh;
  const int y = row % g.X[1], z = (row / g.X[1]) % g.X[2], t = row / (g.X[1] * g.X[2]);
  const int oddFirst = (y + z + t) & 1;
  const size_t V = (size_t)2 * g.Vh;
  const VecT<F> *src0 = oddFirst ? odd : even, *src1 = oddFirst ? even : odd;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    VecT<F> v0 = {0, 0, 0, 0}, v1 = {0, 0, 0, 0};   // absent parity is zero-filled (downloadFromCuda_core.h)
    if (src0) v0 = src0[(size_t)j * g.Vh + sid];
    if (src1) v1 = src1[(size_t)j * g.Vh + sid];
    CplxT<Q> c;
    c.re = (Q)(scale * v0.a); c.im = (Q)(scale * v0.b); qk[(size_t)(2 * j) * V + 2 * (size_t)sid] = c;
    c.re = (Q)(scale * v1.a); c.im = (Q)(scale * v1.b); qk[(size_t)(2 * j) * V + 2 * (size_t)sid + 1] = c;
    c.re = (Q)(scale * v0.c); c.im = (Q)(scale * v0.d); qk[(size_t)(2 * j + 1) * V + 2 * (size_t)sid] = c;
    c.re = (Q)(scale * v1.c); c.im = (Q)(scale * v1.d); qk[(size_t)(2 * j + 1) * V + 2 * (size_t)sid + 1] = c;
  }
}

cudaError_t spinor_from_qkxtm(int prec, void *even, void *odd, const void *qk, int qprec, const Geom &g, cudaStream_t st) {
  const int grid = (g.Vh + FB - 1) / FB;
  if (prec == 8 && qprec == 8) from_qkxtm_kernel<double, double><<<grid, FB, 0, st>>>((VecT<double> *)even, (VecT<double> *)odd, (const CplxT<double> *)qk, g);
  else if (prec == 8 && qprec == 4) from_qkxtm_kernel<double, float><<<grid, FB, 0, st>>>((VecT<double> *)even, (VecT<double> *)odd, (const CplxT<float> *)qk, g);
  else if (prec == 4 && qprec == 8) from_qkxtm_kernel<float, double><<<grid, FB, 0, st>>>((VecT<float> *)even, (VecT<float> *)odd, (const CplxT<double> *)qk, g);
  else from_qkxtm_kernel<float, float><<<grid, FB, 0, st>>>((VecT<float> *)even, (VecT<float> *)odd, (const CplxT<float> *)qk, g);
  return cudaGetLastError();
}
cudaError_t spinor_to_qkxtm(void *qk, int qprec, int prec, const void *even, const void *odd, double scale,
                            const Geom &g, cudaStream_t st) {
  const int grid = (g.Vh + FB - 1) / FB;
  if (prec == 8 && qprec == 8) to_qkxtm_kernel<double, double><<<grid, FB, 0, st>>>((CplxT<double> *)qk, (const VecT<double> *)even, (const VecT<double> *)odd, scale, g);
  else if (prec == 8 && qprec == 4) to_qkxtm_kernel<double, float><<<grid, FB, 0, st>>>((CplxT<float> *)qk, (const VecT<double> *)even, (const VecT<double> *)odd, scale, g);
  else if (prec == 4 && qprec == 8) to_qkxtm_kernel<float, double><<<grid, FB, 0, st>>>((CplxT<double> *)qk, (const VecT<float> *)even, (const VecT<float> *)odd, (float)scale, g);
  else to_qkxtm_kernel<float, float><<<grid, FB, 0, st>>>((CplxT<float> *)qk, (const VecT<float> *)even, (const VecT<float> *)odd, (float)scale, g);
  return cudaGetLastError();
}

// ---- spinor: host even-odd AoS (already copied to the device, double) <-> native -----------------------------
template <typename F> __global__ void from_aos_kernel(VecT<F> *dst, const double *__restrict__ aos, int Vh) {
  const int i = blockIdx.x * FB + threadIdx.x;
  if (i >= Vh) return;
  const double *s = aos + (size_t)i * 24;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    VecT<F> v; v.a = (F)s[4 * j]; v.b = (F)s[4 * j + 1]; v.c = (F)s[4 * j + 2]; v.d = (F)s[4 * j + 3];
    dst[(size_t)j * Vh + i] = v;
  }
}
template <typename F> __global__ void to_aos_kernel(double *aos, const VecT<F> *__restrict__ src, int Vh) {
  const int i = blockIdx.x * FB + threadIdx.x;
  if (i >= Vh) return;
  double *s = aos + (size_t)i * 24;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    VecT<F> v = src[(size_t)j * Vh + i];
    s[4 * j] = v.a; s[4 * j + 1] = v.b; s[4 * j + 2] = v.c; s[4 * j + 3] = v.d;
  }
}
// plug-in host order [x_lex][spin][colour][re,im] (lib/qudaQKXTM_Vector.cpp:72-81 BEFORE packVector) <-> native, both parities at once:
// thread sid owns the lexicographic sites 2 sid and 2 sid + 1 (384 contiguous bytes), one of each parity (uploadToCuda_core.h:7-20)
template <typename F> __global__ void from_aos_lex_kernel(VecT<F> *even, VecT<F> *odd, const double *__restrict__ aos, Geom g) {
  const int sid = blockIdx.x * FB + threadIdx.x;
  if (sid >= g.Vh) return;
  const int row = sid / g.Xh;
  const int y = row % g.X[1], z = (row / g.X[1]) % g.X[2], t = row / (g.X[1] * g.X[2]);
  const int oddFirst = (y + z + t) & 1;
  VecT<F> *dst0 = oddFirst ? odd : even, *dst1 = oddFirst ? even : odd;
  const double *s = aos + (size_t)sid * 48;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    VecT<F> v; v.a = (F)s[4 * j]; v.b = (F)s[4 * j + 1]; v.c = (F)s[4 * j + 2]; v.d = (F)s[4 * j + 3];
    dst0[(size_t)j * g.Vh + sid] = v;
    VecT<F> w; w.a = (F)s[24 + 4 * j]; w.b = (F)s[24 + 4 * j + 1]; w.c = (F)s[24 + 4 * j + 2]; w.d = (F)s[24 + 4 * j + 3];
    dst1[(size_t)j * g.Vh + sid] = w;
  }
}
template <typename F> __global__ void to_aos_lex_kernel(double *aos, const VecT<F> *__restrict__ even, const VecT<F> *__restrict__ odd, double scale, Geom g) {
  const int sid = blockIdx.x * FB + threadIdx.x;
  if (sid >= g.Vh) return;
  const int row = sid / g.Xh;
  const int y = row % g.X[1], z = (row / g.X[1]) % g.X[2], t = row / (g.X[1] * g.X[2]);
  const int oddFirst = (y + z + t) & 1;
  const VecT<F> *src0 = oddFirst ? odd : even, *src1 = oddFirst ? even : odd;
  double *s = aos + (size_t)sid * 48;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const VecT<F> v = src0[(size_t)j * g.Vh + sid], w = src1[(size_t)j * g.Vh + sid];
    s[4 * j] = scale * v.a; s[4 * j + 1] = scale * v.b; s[4 * j + 2] = scale * v.c; s[4 * j + 3] = scale * v.d;
    s[24 + 4 * j] = scale * w.a; s[24 + 4 * j + 1] = scale * w.b; s[24 + 4 * j + 2] = scale * w.c; s[24 + 4 * j + 3] = scale * w.d;
  }
}
cudaError_t spinor_from_host_lex(int prec, void *even, void *odd, const double *d_aos, const Geom &g, cudaStream_t st) {
  const int grid = (g.Vh + FB - 1) / FB;
  if (prec == 8) from_aos_lex_kernel<double><<<grid, FB, 0, st>>>((VecT<double> *)even, (VecT<double> *)odd, d_aos, g);
  else from_aos_lex_kernel<float><<<grid, FB, 0, st>>>((VecT<float> *)even, (VecT<float> *)odd, d_aos, g);
  return cudaGetLastError();
}
cudaError_t spinor_to_host_lex(double *d_aos, int prec, const void *even, const void *odd, double scale, const Geom &g, cudaStream_t st) {
  const int grid = (g.Vh + FB - 1) / FB;
  if (prec == 8) to_aos_lex_kernel<double><<<grid, FB, 0, st>>>(d_aos, (const VecT<double> *)even, (const VecT<double> *)odd, scale, g);
  else to_aos_lex_kernel<float><<<grid, FB, 0, st>>>(d_aos, (const VecT<float> *)even, (const VecT<float> *)odd, scale, g);
  return cudaGetLastError();
}

cudaError_t spinor_from_host_eo(int prec, void *dst, const double *d_aos, int Vh, cudaStream_t st) {
  const int grid = (Vh + FB - 1) / FB;
  if (prec == 8) from_aos_kernel<double><<<grid, FB, 0, st>>>((VecT<double> *)dst, d_aos, Vh);
  else from_aos_kernel<float><<<grid, FB, 0, st>>>((VecT<float> *)dst, d_aos, Vh);
  return cudaGetLastError();
}
cudaError_t spinor_to_host_eo(double *d_aos, int prec, const void *src, int Vh, cudaStream_t st) {
  const int grid = (Vh + FB - 1) / FB;
  if (prec == 8) to_aos_kernel<double><<<grid, FB, 0, st>>>(d_aos, (const VecT<double> *)src, Vh);
  else to_aos_kernel<float><<<grid, FB, 0, st>>>(d_aos, (const VecT<float> *)src, Vh);
  return cudaGetLastError();
}

// ---- plaquette on the native fp64 gauge (single rank; sanity check only) -------------------------------------
__device__ __forceinline__ void mm(double (&c)[3][3][2], const double (&a)[3][3][2], const double (&b)[3][3][2]) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double re = 0, im = 0;
#pragma unroll
      for (int k = 0; k < 3; k++) { TMQ_CMAC(re, im, a[i][k][0], a[i][k][1], b[k][j][0], b[k][j][1]); }
      c[i][j][0] = re; c[i][j][1] = im;
    }
}
template <int RECON>
__global__ void __launch_bounds__(128) plaquette_kernel(const void *gauge, Geom g, BlasRed r) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  double red[1] = {0.0};
  if (i < 2 * g.Vh) {
    const int parity = i / g.Vh, idx = i - parity * g.Vh;
    int xh = idx % g.Xh, row = idx / g.Xh;
    int c[4]; c[1] = row % g.X[1]; c[2] = (row / g.X[1]) % g.X[2]; c[3] = row / (g.X[1] * g.X[2]);
    c[0] = 2 * xh + ((c[1] + c[2] + c[3] + parity) & 1);
    auto cbidx = [&](const int (&x)[4]) { return (((x[3] * g.X[2] + x[2]) * g.X[1] + x[1]) * g.X[0] + x[0]) >> 1; };
    auto sign_at = [&](int mu, const int (&x)[4]) { return (mu == 3 && x[3] == g.X[3] - 1) ? (double)g.tb_sign : 1.0; };
    for (int mu = 0; mu < 4; mu++)
      for (int nu = mu + 1; nu < 4; nu++) {
        int xm[4] = {c[0], c[1], c[2], c[3]}, xn[4] = {c[0], c[1], c[2], c[3]};
        xm[mu] = (xm[mu] + 1) % g.X[mu]; xn[nu] = (xn[nu] + 1) % g.X[nu];
        Link<double> A, B, C, D;
        load_link<double, RECON>(A, gauge, parity, mu, idx, g.Vh, sign_at(mu, c));
        load_link<double, RECON>(B, gauge, 1 - parity, nu, cbidx(xm), g.Vh, sign_at(nu, xm));
        load_link<double, RECON>(C, gauge, parity, nu, idx, g.Vh, sign_at(nu, c));
        load_link<double, RECON>(D, gauge, 1 - parity, mu, cbidx(xn), g.Vh, sign_at(mu, xn));
        double ab[3][3][2], cd[3][3][2];
        mm(ab, A.u, B.u); mm(cd, C.u, D.u);
        // Re tr (ab cd^dag) = sum_ij Re(ab_ij conj(cd_ij))
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int b = 0; b < 3; b++) red[0] += ab[a][b][0] * cd[a][b][0] + ab[a][b][1] * cd[a][b][1];
      }
  }
  block_reduce_finalize<1>(red, r.partials, r.ticket, r.scal, r.slot);
}
cudaError_t plaquette_launch(int recon, const void *gauge_d, const Geom &g, const BlasRed &r, cudaStream_t st) {
  const int grid = (2 * g.Vh + 127) / 128;
  if (recon == 8) plaquette_kernel<8><<<grid, 128, 0, st>>>(gauge_d, g, r);
  else if (recon == 12) plaquette_kernel<12><<<grid, 128, 0, st>>>(gauge_d, g, r);
  else plaquette_kernel<18><<<grid, 128, 0, st>>>(gauge_d, g, r);
  return cudaGetLastError();
}

// ---- the containers' ghost zones (lib/qudaQKXTM_Field.cpp:116-125, lib/qudaQKXTM_kernels.cu:160-170) --------------------------------
// A container's device array is [ncomp][V] complex followed, for every partitioned dimension d in ascending order, by the "plus" ghost
// (the forward neighbour's slice 0) and the "minus" ghost (the backward neighbour's slice L-1), each [ncomp][surface3D[d]]; the face
// site index is the lexicographic index of the three other coordinates (LEXIC_ZYX for t, LEXIC_TYX for z, include/qudaQKXTM_utils.h:25-29).
QkGhost qk_ghost_layout(const Geom &g) {
  QkGhost G;
  const size_t V = (size_t)2 * g.Vh;
  size_t last = V;
  for (int d = 0; d < 4; d++) {
    G.surf[d] = g.part[d] ? V / g.X[d] : 0;
    G.plus[d] = G.minus[d] = 0;
    if (g.part[d]) { G.plus[d] = last; G.minus[d] = last + G.surf[d]; last += 2 * G.surf[d]; }
  }
  G.total_sites = last;
  return G;
}
// gather the slices 0 (lo) and L-1 (hi) of dimension d, all components: out[comp][face]
template <typename Q>
__global__ void qk_face_gather_kernel(CplxT<Q> *lo, CplxT<Q> *hi, const CplxT<Q> *__restrict__ d, Geom g, int dim, int ncomp) {
  const size_t V = (size_t)2 * g.Vh, surf = V / g.X[dim];
  const size_t i = (size_t)blockIdx.x * FB + threadIdx.x;
  if (i >= surf * ncomp) return;
  const size_t comp = i / surf, f = i - comp * surf;
  size_t x0, x1;
  if (dim == 3) { x0 = f; x1 = (size_t)(g.X[3] - 1) * surf + f; }
  else {   // dim == 2: f = t * (X Y) + (y X + x)
    const size_t plane = (size_t)g.X[0] * g.X[1], t = f / plane, rem = f - t * plane;
    x0 = (t * g.X[2]) * plane + rem; x1 = (t * g.X[2] + g.X[2] - 1) * plane + rem;
  }
  lo[i] = d[comp * V + x0];
  hi[i] = d[comp * V + x1];
}
cudaError_t qkxtm_face_gather(void *lo, void *hi, const void *d, int prec, const Geom &g, int dim, int ncomp, cudaStream_t st) {
  const size_t n = ((size_t)2 * g.Vh / g.X[dim]) * ncomp;
  const int grid = (int)((n + FB - 1) / FB);
  if (prec == 8) qk_face_gather_kernel<double><<<grid, FB, 0, st>>>((CplxT<double> *)lo, (CplxT<double> *)hi, (const CplxT<double> *)d, g, dim, ncomp);
  else qk_face_gather_kernel<float><<<grid, FB, 0, st>>>((CplxT<float> *)lo, (CplxT<float> *)hi, (const CplxT<float> *)d, g, dim, ncomp);
  return cudaGetLastError();
}

// ---- plaquette on the QKXTM gauge layout d[((dir*3+c1)*3+c2)*V + x_lex] (lib/qudaQKXTM_Gauge.cpp:73-89):
//      QKXTM_Gauge::calculatePlaq (lib/qudaQKXTM_Gauge.cpp:376-386, lib/code_pieces/plaquette_core.h).  On a partitioned dimension the
//      forward neighbour across the boundary is read from the container's "plus" ghost zone (plaquette_core.h:30-50).
template <typename Q>
__global__ void __launch_bounds__(128) qk_plaquette_kernel(const CplxT<Q> *__restrict__ gq, Geom g, QkGhost gh, BlasRed r) {
  const size_t V = (size_t)2 * g.Vh;
  const size_t i = (size_t)blockIdx.x * 128 + threadIdx.x;
  double red[1] = {0.0};
  if (i < V) {
    int c[4];
    c[0] = (int)(i % g.X[0]); c[1] = (int)((i / g.X[0]) % g.X[1]);
    c[2] = (int)((i / ((size_t)g.X[0] * g.X[1])) % g.X[2]); c[3] = (int)(i / ((size_t)g.X[0] * g.X[1] * g.X[2]));
    auto lex = [&](const int (&x)[4]) { return (size_t)x[0] + (size_t)g.X[0] * (x[1] + (size_t)g.X[1] * (x[2] + (size_t)g.X[2] * x[3])); };
    auto ld = [&](double (&u)[3][3][2], int mu, size_t x) {
#pragma unroll
      for (int k = 0; k < 9; k++) { CplxT<Q> z = gq[((size_t)mu * 9 + k) * V + x]; u[k / 3][k % 3][0] = (double)z.re; u[k / 3][k % 3][1] = (double)z.im; }
    };
    // link mu at the site one step forward in direction d from c (ghost zone if that crosses a partitioned boundary)
    auto ld_fwd = [&](double (&u)[3][3][2], int mu, int d) {
      if (g.part[d] && c[d] == g.X[d] - 1) {
        size_t f;
        if (d == 3) f = (size_t)c[0] + (size_t)g.X[0] * (c[1] + (size_t)g.X[1] * c[2]);
        else f = (size_t)c[0] + (size_t)g.X[0] * (c[1] + (size_t)g.X[1] * c[3]);      // d == 2 (only z and t are ever partitioned)
        const size_t base = gh.plus[d] * 36 + (size_t)mu * 9 * gh.surf[d] + f;
#pragma unroll
        for (int k = 0; k < 9; k++) { CplxT<Q> z = gq[base + (size_t)k * gh.surf[d]]; u[k / 3][k % 3][0] = (double)z.re; u[k / 3][k % 3][1] = (double)z.im; }
      } else {
        int xn[4] = {c[0], c[1], c[2], c[3]};
        xn[d] = (xn[d] + 1) % g.X[d];
        ld(u, mu, lex(xn));
      }
    };
    for (int mu = 0; mu < 4; mu++)
      for (int nu = mu + 1; nu < 4; nu++) {
        double A[3][3][2], B[3][3][2], C[3][3][2], D[3][3][2], ab[3][3][2], cd[3][3][2];
        ld(A, mu, i); ld_fwd(B, nu, mu); ld(C, nu, i); ld_fwd(D, mu, nu);
        mm(ab, A, B); mm(cd, C, D);
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int b = 0; b < 3; b++) red[0] += ab[a][b][0] * cd[a][b][0] + ab[a][b][1] * cd[a][b][1];
      }
  }
  block_reduce_finalize<1>(red, r.partials, r.ticket, r.scal, r.slot);
}
cudaError_t qkxtm_plaquette(const void *gq, int prec, const Geom &g, const BlasRed &r, cudaStream_t st) {
  const int grid = (2 * g.Vh + 127) / 128;
  const QkGhost gh = qk_ghost_layout(g);
  if (prec == 8) qk_plaquette_kernel<double><<<grid, 128, 0, st>>>((const CplxT<double> *)gq, g, gh, r);
  else qk_plaquette_kernel<float><<<grid, 128, 0, st>>>((const CplxT<float> *)gq, g, gh, r);
  return cudaGetLastError();
}

// ---- QKXTM container kernels on the QKXTM layout ---------------------------------------------------------------
template <typename Q> __global__ void qk_scale_kernel(CplxT<Q> *d, Q a, size_t n) {
  for (size_t i = (size_t)blockIdx.x * FB + threadIdx.x; i < n; i += (size_t)gridDim.x * FB) {
    CplxT<Q> c = d[i]; c.re *= a; c.im *= a; d[i] = c;
  }
}
template <typename D, typename S> __global__ void qk_cast_kernel(CplxT<D> *dst, const CplxT<S> *src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * FB + threadIdx.x; i < n; i += (size_t)gridDim.x * FB) {
    CplxT<S> c = src[i]; CplxT<D> o; o.re = (D)c.re; o.im = (D)c.im; dst[i] = o;
  }
}
// gamma5 in the UKQCD basis = swap spin 0<->2, 1<->3 (apply_gamma5_vector_core.h:1-16)
template <typename Q> __global__ void qk_gamma5_kernel(CplxT<Q> *d, size_t V) {
  for (size_t i = (size_t)blockIdx.x * FB + threadIdx.x; i < 6 * V; i += (size_t)gridDim.x * FB) {
    CplxT<Q> a = d[i], b = d[i + 6 * V]; d[i] = b; d[i + 6 * V] = a;
  }
}
static int sgrid(size_t n) { size_t b = (n + FB - 1) / FB; size_t cap = (size_t)blas_grid(); return (int)(b < cap ? (b ? b : 1) : cap); }
cudaError_t qkxtm_scale(void *d, int prec, double a, size_t n, cudaStream_t st) {
  if (prec == 8) qk_scale_kernel<double><<<sgrid(n), FB, 0, st>>>((CplxT<double> *)d, a, n);
  else qk_scale_kernel<float><<<sgrid(n), FB, 0, st>>>((CplxT<float> *)d, (float)a, n);
  return cudaGetLastError();
}
cudaError_t qkxtm_cast(void *dst, int dprec, const void *src, int sprec, size_t n, cudaStream_t st) {
  if (dprec == sprec) return cudaMemcpyAsync(dst, src, n * 2 * dprec, cudaMemcpyDeviceToDevice, st);
  if (dprec == 8) qk_cast_kernel<double, float><<<sgrid(n), FB, 0, st>>>((CplxT<double> *)dst, (const CplxT<float> *)src, n);
  else qk_cast_kernel<float, double><<<sgrid(n), FB, 0, st>>>((CplxT<float> *)dst, (const CplxT<double> *)src, n);
  return cudaGetLastError();
}
cudaError_t qkxtm_gamma5(void *d, int prec, int V, cudaStream_t st) {
  const size_t n = (size_t)6 * V;
  if (prec == 8) qk_gamma5_kernel<double><<<sgrid(n), FB, 0, st>>>((CplxT<double> *)d, (size_t)V);
  else qk_gamma5_kernel<float><<<sgrid(n), FB, 0, st>>>((CplxT<float> *)d, (size_t)V);
  return cudaGetLastError();
}

}  // namespace tmq
