// tmq_site.cuh -- the per-site body of the twisted-mass Dslash family.
//
// One function, dslash_site<F, RECON, EPI, MULTI>(), computes for one output parity site
//     h(x) = sum_mu [ (1 - s g_mu) U_mu(x) in(x+mu) + (1 + s g_mu) U_mu(x-mu)^dag in(x-mu) ]      (s = +1: D, -1: D^dag)
// in the UKQCD basis and applies the fused epilogue (tmq_types.h).  Hop convention:
// reference lib/code_pieces/fixSinkContractions_noether_core.h:117-137; gamma matrices:
// lib/code_pieces/gammas_tm_base.h:21-32,148-171; gamma5 = spin swap (apply_gamma5_vector_core.h);
// checkerboard geometry: qkxtm/QKXTM_util.cpp:405-470; recon-12 third row and its boundary sign:
// qkxtm/QKXTM_util.cpp:281-295.  Written from the maths, not from QUDA's dslash_core.
//
// The body is __host__ __device__ so that the CPU test-suite can run the identical index / projector
// / epilogue code on host arrays (tests/hostemu, test infrastructure only) before GPU time is spent.
// The product never calls it on the host.
#pragma once
#include "tmq_types.h"

namespace tmq {

// ---- loads / stores with cache policy ---------------------------------------------------------------
// spinor neighbours: read-only path, allocate in L1 (each input site is used by 8 outputs, most of them
// in the same CTA tile).  gauge links: used exactly once per application -> no L1 allocation, L2
// evict-first so they do not push the spinor working set out of L2.
TMQ_HD VecT<double> ld_spinor(const VecT<double> *p) {
#if defined(__CUDA_ARCH__)
  VecT<double> r;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
  return r;
#else
  return *p;
#endif
}
TMQ_HD VecT<float> ld_spinor(const VecT<float> *p) {
#if defined(__CUDA_ARCH__)
  VecT<float> r;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.a), "=f"(r.b), "=f"(r.c), "=f"(r.d) : "l"(p));
  return r;
#else
  return *p;
#endif
}
TMQ_HD VecT<double> ld_stream(const VecT<double> *p) {
#if defined(__CUDA_ARCH__)
  VecT<double> r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
  return r;
#else
  return *p;
#endif
}
TMQ_HD VecT<float> ld_stream(const VecT<float> *p) {
#if defined(__CUDA_ARCH__)
  VecT<float> r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.a), "=f"(r.b), "=f"(r.c), "=f"(r.d) : "l"(p));
  return r;
#else
  return *p;
#endif
}
TMQ_HD CplxT<double> ld_stream(const CplxT<double> *p) {
#if defined(__CUDA_ARCH__)
  CplxT<double> r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.re), "=d"(r.im) : "l"(p));
  return r;
#else
  return *p;
#endif
}
TMQ_HD CplxT<float> ld_stream(const CplxT<float> *p) {
#if defined(__CUDA_ARCH__)
  CplxT<float> r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.re), "=f"(r.im) : "l"(p));
  return r;
#else
  return *p;
#endif
}

// ---- site enumeration -----------------------------------------------------------------------------------
struct SiteCoord {
  int xh, y, z, t;   // local coordinates (xh = x/2)
  int xodd;          // x & 1
  int idx;           // cb-lexicographic index inside the parity block
};

TMQ_HD SiteCoord decode_site(const Geom &g, const Enum &en, int parity, uint32_t e) {
  SiteCoord c;
  uint32_t r = fd_div(e, en.dXh);
  c.xh = (int)(e - r * en.dXh.d);
  uint32_t q = fd_div(r, en.dTy); int ly = (int)(r - q * en.dTy.d); r = q;
  q = fd_div(r, en.dTz); int lz = (int)(r - q * en.dTz.d); r = q;
  q = fd_div(r, en.dTt); int lt = (int)(r - q * en.dTt.d); r = q;
  q = fd_div(r, en.dNy); int ty = (int)(r - q * en.dNy.d); r = q;
  q = fd_div(r, en.dNz); int tz = (int)(r - q * en.dNz.d); int tt = (int)q;
  c.y = en.lo[0] + (ty * (int)en.dTy.d + ly) * en.step[0];
  c.z = en.lo[1] + (tz * (int)en.dTz.d + lz) * en.step[1];
  c.t = en.lo[2] + (tt * (int)en.dTt.d + lt) * en.step[2];
  c.xodd = (c.y + c.z + c.t + parity) & 1;
  c.idx = ((c.t * g.X[2] + c.z) * g.X[1] + c.y) * g.Xh + c.xh;
  return c;
}

// ---- small complex helpers on explicit re/im scalars ---------------------------------------------------
// acc += a*b
#define TMQ_CMAC(accr, acci, ar, ai, br, bi) \
  { accr += (ar) * (br); accr -= (ai) * (bi); acci += (ar) * (bi); acci += (ai) * (br); }
// acc += conj(a)*b
#define TMQ_CMAC_CONJ(accr, acci, ar, ai, br, bi) \
  { accr += (ar) * (br); accr += (ai) * (bi); acci += (ar) * (bi); acci -= (ai) * (br); }

template <typename F> struct Link { F u[3][3][2]; };

// rows 0,1 stored; row 2 = conj(row0 x row1) * sign  (qkxtm/QKXTM_util.cpp:281-295, u0 = t_boundary)
template <typename F> TMQ_HD void reconstruct_row2(Link<F> &L, F sign) {
#define U(r, c, p) L.u[r][c][p]
  F ar, ai;
  // U20 = conj(U01 U12 - U02 U11)
  ar = U(0, 1, 0) * U(1, 2, 0) - U(0, 1, 1) * U(1, 2, 1) - U(0, 2, 0) * U(1, 1, 0) + U(0, 2, 1) * U(1, 1, 1);
  ai = U(0, 1, 0) * U(1, 2, 1) + U(0, 1, 1) * U(1, 2, 0) - U(0, 2, 0) * U(1, 1, 1) - U(0, 2, 1) * U(1, 1, 0);
  U(2, 0, 0) = sign * ar; U(2, 0, 1) = -sign * ai;
  // U21 = conj(U02 U10 - U00 U12)
  ar = U(0, 2, 0) * U(1, 0, 0) - U(0, 2, 1) * U(1, 0, 1) - U(0, 0, 0) * U(1, 2, 0) + U(0, 0, 1) * U(1, 2, 1);
  ai = U(0, 2, 0) * U(1, 0, 1) + U(0, 2, 1) * U(1, 0, 0) - U(0, 0, 0) * U(1, 2, 1) - U(0, 0, 1) * U(1, 2, 0);
  U(2, 1, 0) = sign * ar; U(2, 1, 1) = -sign * ai;
  // U22 = conj(U00 U11 - U01 U10)
  ar = U(0, 0, 0) * U(1, 1, 0) - U(0, 0, 1) * U(1, 1, 1) - U(0, 1, 0) * U(1, 0, 0) + U(0, 1, 1) * U(1, 0, 1);
  ai = U(0, 0, 0) * U(1, 1, 1) + U(0, 0, 1) * U(1, 1, 0) - U(0, 1, 0) * U(1, 0, 1) - U(0, 1, 1) * U(1, 0, 0);
  U(2, 2, 0) = sign * ar; U(2, 2, 1) = -sign * ai;
#undef U
}

// 8-real format (tools/recon8_study.py): U01, U02, U10 and t00 = tan(arg U00 / 4), t20 = tan(arg U20 / 4) are stored; for a unitary
// link (times the boundary sign)  |U00|^2 = 1 - |U01|^2 - |U02|^2,  |U20|^2 = |U01|^2 + |U02|^2 - |U10|^2,  the phases come back by
// the rational half-angle formulas (1 + t^2 in [1, 2]: no singular point, no trigonometric function), U11 and U12 solve
//   conj(U01) U11 + conj(U02) U12 = -U10 conj(U00)   (rows 0 and 1 orthogonal)
//   -U02 U11 + U01 U12 = sign conj(U20)              (row 2 = sign conj(row0 x row1), column 0)
// whose determinant is |U01|^2 + |U02|^2 (links with U01 = U02 = 0, e.g. a unit field, cannot be stored: refused at load time).
// reciprocal and reciprocal square root for the unpacking: a single-precision hardware seed and two Newton steps (the IEEE division
// and square root of fp64 made the 8-real kernels compute-bound: 1.02 ms against 0.96 ms for the 12-real K1 at 48^3x96,
// profiles/r25_sweep_recon8_exact_sqrt_div.jsonl); arguments lie in [1e-6, 8], well inside the single-precision range
TMQ_HD double fast_rcp(double x) {
#if defined(__CUDA_ARCH__)
  double r = (double)__frcp_rn((float)x);
  r = r * (2.0 - x * r);
  return r * (2.0 - x * r);
#else
  return 1.0 / x;
#endif
}
TMQ_HD float fast_rcp(float x) {
#if defined(__CUDA_ARCH__)
  return __frcp_rn(x);
#else
  return 1.0f / x;
#endif
}
// sqrt(x) for x >= 0 as x * rsqrt(x); 0 for x below the smallest normal single (rounding residue of an exactly vanishing entry)
TMQ_HD double fast_sqrt_nn(double x) {
#if defined(__CUDA_ARCH__)
  if (!(x > 1e-36)) return 0.0;
  double r = (double)rsqrtf((float)x);
  r = r * (1.5 - 0.5 * x * r * r);
  r = r * (1.5 - 0.5 * x * r * r);
  return x * r;
#else
  return x > 0.0 ? sqrt(x) : 0.0;
#endif
}
TMQ_HD float fast_sqrt_nn(float x) {
#if defined(__CUDA_ARCH__)
  return x > 1e-36f ? x * rsqrtf(x) : 0.0f;
#else
  return x > 0.0f ? sqrtf(x) : 0.0f;
#endif
}

template <typename F> TMQ_HD void reconstruct_from8(Link<F> &L, F t00, F t20, F sign) {
#define U(r, c, p) L.u[r][c][p]
  const F rs = U(0, 1, 0) * U(0, 1, 0) + U(0, 1, 1) * U(0, 1, 1) + U(0, 2, 0) * U(0, 2, 0) + U(0, 2, 1) * U(0, 2, 1);
  const F b2 = U(1, 0, 0) * U(1, 0, 0) + U(1, 0, 1) * U(1, 0, 1);
  const F m00 = fast_sqrt_nn((F)1 - rs), m20 = fast_sqrt_nn(rs - b2);
  // one reciprocal serves the three divisions: 1/(1 + t00^2), 1/(1 + t20^2), 1/rs
  const F e00 = (F)1 + t00 * t00, e20 = (F)1 + t20 * t20;
  const F q = fast_rcp(e00 * e20 * rs);
  const F inv = q * e00 * e20;
  {
    const F d = q * e20 * rs, c2 = ((F)1 - t00 * t00) * d, s2 = (F)2 * t00 * d;
    U(0, 0, 0) = m00 * (c2 * c2 - s2 * s2); U(0, 0, 1) = m00 * ((F)2 * s2 * c2);
  }
  {
    const F d = q * e00 * rs, c2 = ((F)1 - t20 * t20) * d, s2 = (F)2 * t20 * d;
    U(2, 0, 0) = m20 * (c2 * c2 - s2 * s2); U(2, 0, 1) = m20 * ((F)2 * s2 * c2);
  }
  // rhs1 = -U10 conj(U00), rhs2 = sign conj(U20)
  const F r1r = -(U(1, 0, 0) * U(0, 0, 0) + U(1, 0, 1) * U(0, 0, 1)), r1i = -(U(1, 0, 1) * U(0, 0, 0) - U(1, 0, 0) * U(0, 0, 1));
  const F r2r = sign * U(2, 0, 0), r2i = -sign * U(2, 0, 1);
  // U11 = (rhs1 U01 - conj(U02) rhs2) / rs
  U(1, 1, 0) = inv * (r1r * U(0, 1, 0) - r1i * U(0, 1, 1) - (U(0, 2, 0) * r2r + U(0, 2, 1) * r2i));
  U(1, 1, 1) = inv * (r1r * U(0, 1, 1) + r1i * U(0, 1, 0) - (U(0, 2, 0) * r2i - U(0, 2, 1) * r2r));
  // U12 = (conj(U01) rhs2 + U02 rhs1) / rs
  U(1, 2, 0) = inv * (U(0, 1, 0) * r2r + U(0, 1, 1) * r2i + U(0, 2, 0) * r1r - U(0, 2, 1) * r1i);
  U(1, 2, 1) = inv * (U(0, 1, 0) * r2i - U(0, 1, 1) * r2r + U(0, 2, 0) * r1i + U(0, 2, 1) * r1r);
  F ar, ai;
  // U21 = sign conj(U02 U10 - U00 U12),  U22 = sign conj(U00 U11 - U01 U10)   (as reconstruct_row2)
  ar = U(0, 2, 0) * U(1, 0, 0) - U(0, 2, 1) * U(1, 0, 1) - U(0, 0, 0) * U(1, 2, 0) + U(0, 0, 1) * U(1, 2, 1);
  ai = U(0, 2, 0) * U(1, 0, 1) + U(0, 2, 1) * U(1, 0, 0) - U(0, 0, 0) * U(1, 2, 1) - U(0, 0, 1) * U(1, 2, 0);
  U(2, 1, 0) = sign * ar; U(2, 1, 1) = -sign * ai;
  ar = U(0, 0, 0) * U(1, 1, 0) - U(0, 0, 1) * U(1, 1, 1) - U(0, 1, 0) * U(1, 0, 0) + U(0, 1, 1) * U(1, 0, 1);
  ai = U(0, 0, 0) * U(1, 1, 1) + U(0, 0, 1) * U(1, 1, 0) - U(0, 1, 0) * U(1, 0, 1) - U(0, 1, 1) * U(1, 0, 0);
  U(2, 2, 0) = sign * ar; U(2, 2, 1) = -sign * ai;
#undef U
}

template <typename F, int RECON>
TMQ_HD void load_link(Link<F> &L, const void *gauge, int parity, int mu, int idx, int stride, F sign12) {
  if (RECON == 8) {
    const VecT<F> *b = (const VecT<F> *)gauge + (size_t)((parity * 4 + mu) * 2) * (size_t)stride + idx;
    VecT<F> v0 = ld_stream(b), v1 = ld_stream(b + stride);
    L.u[0][1][0] = v0.a; L.u[0][1][1] = v0.b; L.u[0][2][0] = v0.c; L.u[0][2][1] = v0.d;
    L.u[1][0][0] = v1.a; L.u[1][0][1] = v1.b;
    reconstruct_from8(L, v1.c, v1.d, sign12);
  } else if (RECON == 12) {
    const VecT<F> *b = (const VecT<F> *)gauge + (size_t)((parity * 4 + mu) * 3) * (size_t)stride + idx;
    VecT<F> v0 = ld_stream(b), v1 = ld_stream(b + stride), v2 = ld_stream(b + 2 * (size_t)stride);
    L.u[0][0][0] = v0.a; L.u[0][0][1] = v0.b; L.u[0][1][0] = v0.c; L.u[0][1][1] = v0.d;
    L.u[0][2][0] = v1.a; L.u[0][2][1] = v1.b; L.u[1][0][0] = v1.c; L.u[1][0][1] = v1.d;
    L.u[1][1][0] = v2.a; L.u[1][1][1] = v2.b; L.u[1][2][0] = v2.c; L.u[1][2][1] = v2.d;
    reconstruct_row2(L, sign12);
  } else {
    const CplxT<F> *b = (const CplxT<F> *)gauge + (size_t)((parity * 4 + mu) * 9) * (size_t)stride + idx;
#pragma unroll
    for (int k = 0; k < 9; k++) {
      CplxT<F> c = ld_stream(b + (size_t)k * stride);
      L.u[k / 3][k % 3][0] = c.re; L.u[k / 3][k % 3][1] = c.im;
    }
  }
}

// half spinor: 2 spin x 3 colour complex
template <typename F> struct Half { F h[2][3][2]; };

// uh = U h (forward) or U^dag h (backward)
template <typename F, bool DAG> TMQ_HD void su3_apply(Half<F> &o, const Link<F> &L, const Half<F> &h) {
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int a = 0; a < 3; a++) {
      F re = 0, im = 0;
#pragma unroll
      for (int b = 0; b < 3; b++) {
        if (!DAG) { TMQ_CMAC(re, im, L.u[a][b][0], L.u[a][b][1], h.h[s][b][0], h.h[s][b][1]); }
        else      { TMQ_CMAC_CONJ(re, im, L.u[b][a][0], L.u[b][a][1], h.h[s][b][0], h.h[s][b][1]); }
      }
      o.h[s][a][0] = re; o.h[s][a][1] = im;
    }
}

// full spinor in registers: o[spin][colour][re/im]
template <typename F> struct Spinor { F v[4][3][2]; };

template <typename F> TMQ_HD void unpack_vec(Spinor<F> &p, int j, const VecT<F> &v) {
  // vector j holds complex k = 2j, 2j+1 with k = 3*spin + colour
  const int k0 = 2 * j, k1 = 2 * j + 1;
  p.v[k0 / 3][k0 % 3][0] = v.a; p.v[k0 / 3][k0 % 3][1] = v.b;
  p.v[k1 / 3][k1 % 3][0] = v.c; p.v[k1 / 3][k1 % 3][1] = v.d;
}
template <typename F> TMQ_HD VecT<F> pack_vec(const Spinor<F> &p, int j) {
  const int k0 = 2 * j, k1 = 2 * j + 1;
  VecT<F> v;
  v.a = p.v[k0 / 3][k0 % 3][0]; v.b = p.v[k0 / 3][k0 % 3][1];
  v.c = p.v[k1 / 3][k1 % 3][0]; v.d = p.v[k1 / 3][k1 % 3][1];
  return v;
}

template <typename F> TMQ_HD void load_spinor(Spinor<F> &p, const VecT<F> *base, int idx, int stride) {
#pragma unroll
  for (int j = 0; j < 6; j++) unpack_vec(p, j, ld_spinor(base + (size_t)j * stride + idx));
}
// only spins {2*half, 2*half+1}: vectors 3*half .. 3*half+2
template <typename F> TMQ_HD void load_spinor_half(Spinor<F> &p, const VecT<F> *base, int idx, int stride, int half) {
#pragma unroll
  for (int j = 0; j < 3; j++) {
    VecT<F> v = ld_spinor(base + (size_t)(3 * half + j) * stride + idx);
    // unpack with compile-time register indices for both cases
    if (half) unpack_vec(p, 3 + j, v); else unpack_vec(p, j, v);
  }
}
// ghost faces may have been written by a peer GPU during this launch (peer-memory halo path): coherent loads,
// ordered after the acquire of the arrival flag
TMQ_HD VecT<double> ld_ghost(const VecT<double> *p) {
#if defined(__CUDA_ARCH__)
  VecT<double> r;
  asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p) : "memory");
  return r;
#else
  return *p;
#endif
}
TMQ_HD VecT<float> ld_ghost(const VecT<float> *p) {
#if defined(__CUDA_ARCH__)
  VecT<float> r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.a), "=f"(r.b), "=f"(r.c), "=f"(r.d) : "l"(p) : "memory");
  return r;
#else
  return *p;
#endif
}
template <typename F> TMQ_HD void load_half_ghost(Half<F> &h, const VecT<F> *base, int fidx, int fstride) {
#pragma unroll
  for (int j = 0; j < 3; j++) {
    VecT<F> v = ld_ghost(base + (size_t)j * fstride + fidx);
    const int k0 = 2 * j, k1 = 2 * j + 1;
    h.h[k0 / 3][k0 % 3][0] = v.a; h.h[k0 / 3][k0 % 3][1] = v.b;
    h.h[k1 / 3][k1 % 3][0] = v.c; h.h[k1 / 3][k1 % 3][1] = v.d;
  }
}

// ---- spin projection / reconstruction in the UKQCD basis, P = 1 - sg * gamma_mu --------------------
//   mu=0 (x): h0 = p0 - sg i p3, h1 = p1 - sg i p2 ; rows 2,3 = sg i h1, sg i h0
//   mu=1 (y): h0 = p0 - sg p3,   h1 = p1 + sg p2   ; rows 2,3 = sg h1,  -sg h0
//   mu=2 (z): h0 = p0 - sg i p2, h1 = p1 + sg i p3 ; rows 2,3 = sg i h0, -sg i h1
//   mu=3 (t): sg=+1: h = 2 p2, 2 p3 -> rows 2,3 ; sg=-1: h = 2 p0, 2 p1 -> rows 0,1
template <typename F, int MU> TMQ_HD void project(Half<F> &h, const Spinor<F> &p, F sg) {
#pragma unroll
  for (int c = 0; c < 3; c++) {
    if constexpr (MU == 0) {
      h.h[0][c][0] = p.v[0][c][0] + sg * p.v[3][c][1]; h.h[0][c][1] = p.v[0][c][1] - sg * p.v[3][c][0];
      h.h[1][c][0] = p.v[1][c][0] + sg * p.v[2][c][1]; h.h[1][c][1] = p.v[1][c][1] - sg * p.v[2][c][0];
    } else if constexpr (MU == 1) {
      h.h[0][c][0] = p.v[0][c][0] - sg * p.v[3][c][0]; h.h[0][c][1] = p.v[0][c][1] - sg * p.v[3][c][1];
      h.h[1][c][0] = p.v[1][c][0] + sg * p.v[2][c][0]; h.h[1][c][1] = p.v[1][c][1] + sg * p.v[2][c][1];
    } else if constexpr (MU == 2) {
      h.h[0][c][0] = p.v[0][c][0] + sg * p.v[2][c][1]; h.h[0][c][1] = p.v[0][c][1] - sg * p.v[2][c][0];
      h.h[1][c][0] = p.v[1][c][0] - sg * p.v[3][c][1]; h.h[1][c][1] = p.v[1][c][1] + sg * p.v[3][c][0];
    }
  }
}
template <typename F, int MU> TMQ_HD void accumulate(Spinor<F> &o, const Half<F> &u, F sg) {
#pragma unroll
  for (int c = 0; c < 3; c++) {
    o.v[0][c][0] += u.h[0][c][0]; o.v[0][c][1] += u.h[0][c][1];
    o.v[1][c][0] += u.h[1][c][0]; o.v[1][c][1] += u.h[1][c][1];
    if constexpr (MU == 0) {        // row2 += sg i u1 ; row3 += sg i u0
      o.v[2][c][0] -= sg * u.h[1][c][1]; o.v[2][c][1] += sg * u.h[1][c][0];
      o.v[3][c][0] -= sg * u.h[0][c][1]; o.v[3][c][1] += sg * u.h[0][c][0];
    } else if constexpr (MU == 1) { // row2 += sg u1 ; row3 -= sg u0
      o.v[2][c][0] += sg * u.h[1][c][0]; o.v[2][c][1] += sg * u.h[1][c][1];
      o.v[3][c][0] -= sg * u.h[0][c][0]; o.v[3][c][1] -= sg * u.h[0][c][1];
    } else if constexpr (MU == 2) { // row2 += sg i u0 ; row3 -= sg i u1
      o.v[2][c][0] -= sg * u.h[0][c][1]; o.v[2][c][1] += sg * u.h[0][c][0];
      o.v[3][c][0] += sg * u.h[1][c][1]; o.v[3][c][1] -= sg * u.h[1][c][0];
    }
  }
}

// one of the 8 hop terms.  FWD: (1 - s g_mu) U_mu(x) in(x+mu) ; !FWD: (1 + s g_mu) U_mu(x-mu)^dag in(x-mu)
// nidx = neighbour cb index (other parity); cross = the neighbour lies across a partitioned boundary
template <typename F, int RECON, int MU, bool FWD, bool MULTI>
TMQ_HD void hop_term(Spinor<F> &o, const DslashArgs<F> &A, const SiteCoord &c, int nidx, bool cross, int fidx,
                     F sign12) {
  const int stride = A.g.Vh;
  const F sg = FWD ? A.dsign : -A.dsign;
  Half<F> h, u;
  if constexpr (MU < 3) {
    if (MULTI && cross) {
      // ghost faces hold projected half spinors; the one arriving from the backward neighbour is
      // already multiplied by U^dag on the sender (the link lives there)
      if (FWD) {
        load_half_ghost(h, A.ghost[MU][1], fidx, A.g.face[MU]);
        Link<F> L; load_link<F, RECON>(L, A.gauge, A.parity, MU, c.idx, stride, (F)1);
        su3_apply<F, false>(u, L, h);
      } else {
        load_half_ghost(u, A.ghost[MU][0], fidx, A.g.face[MU]);
      }
    } else {
      Spinor<F> p; load_spinor(p, A.in, nidx, stride);
      project<F, MU>(h, p, sg);
      Link<F> L;
      if (FWD) { load_link<F, RECON>(L, A.gauge, A.parity, MU, c.idx, stride, (F)1); su3_apply<F, false>(u, L, h); }
      else     { load_link<F, RECON>(L, A.gauge, 1 - A.parity, MU, nidx, stride, (F)1); su3_apply<F, true>(u, L, h); }
    }
    accumulate<F, MU>(o, u, sg);
  } else {
    // t direction: P = diag(1-sg,1-sg,1+sg,1+sg): only two spin components are read
    const bool lower = sg > (F)0;   // sg=+1 -> spins 2,3
    if (MULTI && cross) {
      if (FWD) {
        load_half_ghost(h, A.ghost[3][1], fidx, A.g.face[3]);
        Link<F> L; load_link<F, RECON>(L, A.gauge, A.parity, 3, c.idx, stride, sign12);
        su3_apply<F, false>(u, L, h);
      } else {
        load_half_ghost(u, A.ghost[3][0], fidx, A.g.face[3]);
      }
    } else {
      Spinor<F> p;
      if (lower) {
        load_spinor_half(p, A.in, nidx, stride, 1);
#pragma unroll
        for (int cc = 0; cc < 3; cc++) {
          h.h[0][cc][0] = 2 * p.v[2][cc][0]; h.h[0][cc][1] = 2 * p.v[2][cc][1];
          h.h[1][cc][0] = 2 * p.v[3][cc][0]; h.h[1][cc][1] = 2 * p.v[3][cc][1];
        }
      } else {
        load_spinor_half(p, A.in, nidx, stride, 0);
#pragma unroll
        for (int cc = 0; cc < 3; cc++) {
          h.h[0][cc][0] = 2 * p.v[0][cc][0]; h.h[0][cc][1] = 2 * p.v[0][cc][1];
          h.h[1][cc][0] = 2 * p.v[1][cc][0]; h.h[1][cc][1] = 2 * p.v[1][cc][1];
        }
      }
      Link<F> L;
      if (FWD) { load_link<F, RECON>(L, A.gauge, A.parity, 3, c.idx, stride, sign12); su3_apply<F, false>(u, L, h); }
      else     { load_link<F, RECON>(L, A.gauge, 1 - A.parity, 3, nidx, stride, sign12); su3_apply<F, true>(u, L, h); }
    }
    if (lower) {
#pragma unroll
      for (int cc = 0; cc < 3; cc++) {
        o.v[2][cc][0] += u.h[0][cc][0]; o.v[2][cc][1] += u.h[0][cc][1];
        o.v[3][cc][0] += u.h[1][cc][0]; o.v[3][cc][1] += u.h[1][cc][1];
      }
    } else {
#pragma unroll
      for (int cc = 0; cc < 3; cc++) {
        o.v[0][cc][0] += u.h[0][cc][0]; o.v[0][cc][1] += u.h[0][cc][1];
        o.v[1][cc][0] += u.h[1][cc][0]; o.v[1][cc][1] += u.h[1][cc][1];
      }
    }
  }
}

// y = c (x + i a g5 x), g5 = spin swap 0<->2, 1<->3
template <typename F> TMQ_HD void twist(Spinor<F> &y, const Spinor<F> &x, F c, F a) {
#pragma unroll
  for (int s = 0; s < 4; s++)
#pragma unroll
    for (int cc = 0; cc < 3; cc++) {
      y.v[s][cc][0] = c * (x.v[s][cc][0] - a * x.v[s ^ 2][cc][1]);
      y.v[s][cc][1] = c * (x.v[s][cc][1] + a * x.v[s ^ 2][cc][0]);
    }
}

// o <- M o (or M^dag o) with M = diag(M+, M-) in the chiral basis chi^(+-)_s = psi_s +- psi_{s+2} (s = 0,1): gamma5 is the
// spin swap in the UKQCD basis (apply_gamma5_vector_core.h), so its eigenvectors are the sums / differences of the upper
// and lower spin pairs and sigma_munu -- hence the clover term -- is block diagonal there.  M: vec[36][stride].
template <typename F> TMQ_HD void clover_mul(Spinor<F> &o, const VecT<F> *M, int idx, int stride, bool dag) {
  F chi[2][6][2], y[2][6][2];
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int r = 0; r < 2; r++) {
        chi[0][s * 3 + c][r] = o.v[s][c][r] + o.v[s + 2][c][r];
        chi[1][s * 3 + c][r] = o.v[s][c][r] - o.v[s + 2][c][r];
        y[0][s * 3 + c][r] = 0; y[1][s * 3 + c][r] = 0;
      }
#pragma unroll
  for (int b = 0; b < 2; b++)
#pragma unroll
    for (int v = 0; v < 18; v++) {
      const VecT<F> m = M[(size_t)(b * 18 + v) * stride + idx];
      const int k0 = 2 * v, k1 = 2 * v + 1;
      const int r0 = k0 / 6, c0 = k0 % 6, r1 = k1 / 6, c1 = k1 % 6;
      if (!dag) {
        TMQ_CMAC(y[b][r0][0], y[b][r0][1], m.a, m.b, chi[b][c0][0], chi[b][c0][1]);
        TMQ_CMAC(y[b][r1][0], y[b][r1][1], m.c, m.d, chi[b][c1][0], chi[b][c1][1]);
      } else {
        TMQ_CMAC_CONJ(y[b][c0][0], y[b][c0][1], m.a, m.b, chi[b][r0][0], chi[b][r0][1]);
        TMQ_CMAC_CONJ(y[b][c1][0], y[b][c1][1], m.c, m.d, chi[b][r1][0], chi[b][r1][1]);
      }
    }
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int r = 0; r < 2; r++) {
        o.v[s][c][r] = (F)0.5 * (y[0][s * 3 + c][r] + y[1][s * 3 + c][r]);
        o.v[s + 2][c][r] = (F)0.5 * (y[0][s * 3 + c][r] - y[1][s * 3 + c][r]);
      }
}

// the same twist on one (vector j, vector j+3) pair: u' = c (u + i a l), l' = c (l + i a u)
template <typename F> TMQ_HD void twist_pair(VecT<F> &u, VecT<F> &l, F c, F a) {
  VecT<F> ou, ol;
  ou.a = c * (u.a - a * l.b); ou.b = c * (u.b + a * l.a); ou.c = c * (u.c - a * l.d); ou.d = c * (u.d + a * l.c);
  ol.a = c * (l.a - a * u.b); ol.b = c * (l.b + a * u.a); ol.c = c * (l.c - a * u.d); ol.d = c * (l.d + a * u.c);
  u = ou; l = ol;
}
template <typename F> TMQ_HD double norm2_vec(const VecT<F> &v) {
  return (double)v.a * (double)v.a + (double)v.b * (double)v.b + (double)v.c * (double)v.c + (double)v.d * (double)v.d;
}

template <int EPI> struct EpiTraits {
  static constexpr bool TW1 = (EPI == EPI_TW || EPI == EPI_TW_XPAY || EPI == EPI_MDAGM2);
  static constexpr bool XTERM = (EPI >= EPI_TW_XPAY);
  static constexpr bool TWX = (EPI == EPI_TWX_XPAY || EPI == EPI_CG4 || EPI == EPI_CHEB);
  static constexpr bool CHEB = (EPI == EPI_CHEB);
  static constexpr bool TW3 = (EPI == EPI_XPAY_TW3 || EPI == EPI_MDAGM2);
  static constexpr int RED = (EPI == EPI_MDAGM2) ? 1 : (EPI == EPI_CG4 ? 2 : 0);
};

// Computes one output site; returns this site's contribution to the fused reduction (0 if none).
template <typename F, int RECON, int EPI, bool MULTI, bool CLOVER = false>
TMQ_HD double dslash_site(const DslashArgs<F> &A, const Enum &en, uint32_t e, F alpha) {
  typedef EpiTraits<EPI> T;
  const Geom &g = A.g;
  const SiteCoord c = decode_site(g, en, A.parity, e);
  const int Xh = g.Xh, stride = g.Vh;
  Spinor<F> o;
#pragma unroll
  for (int s = 0; s < 4; s++)
#pragma unroll
    for (int cc = 0; cc < 3; cc++) { o.v[s][cc][0] = 0; o.v[s][cc][1] = 0; }

  // recon-12 boundary sign: the links U_t(T-1) carry the anti-periodic -1 (QKXTM_util.cpp:698-705),
  // so their reconstructed third row needs the same factor (QKXTM_util.cpp:292-294)
  const F s12_f = (g.tb_last && c.t == g.X[3] - 1) ? (F)g.tb_sign : (F)1;
  // backward t link U_t(x-t): a boundary link iff this site sits at global t = 0 (unused on the ghost
  // path, where the sender applied its own U^dag)
  const F s12_b = (g.tb_first && c.t == 0) ? (F)g.tb_sign : (F)1;

  // +x / -x : xh' = xh + xodd (fwd), xh - (1 - xodd) (bwd), periodic or ghost
  {
    int nx = c.xh + c.xodd; bool cross = false; int fidx = 0;
    if (nx == Xh) { nx = 0; if (MULTI && g.part[0]) { cross = true; fidx = (c.t * g.X[2] + c.z) * g.X[1] + c.y; fidx >>= 1; } }
    hop_term<F, RECON, 0, true, MULTI>(o, A, c, c.idx - c.xh + nx, cross, fidx, (F)1);
    nx = c.xh - (1 - c.xodd); cross = false;
    if (nx < 0) { nx = Xh - 1; if (MULTI && g.part[0]) { cross = true; fidx = ((c.t * g.X[2] + c.z) * g.X[1] + c.y) >> 1; } }
    hop_term<F, RECON, 0, false, MULTI>(o, A, c, c.idx - c.xh + nx, cross, fidx, (F)1);
  }
  // +-y
  {
    bool cross = false; int fidx = 0;
    int n = c.idx + Xh;
    if (c.y == g.X[1] - 1) { n = c.idx - (g.X[1] - 1) * Xh; if (MULTI && g.part[1]) { cross = true; fidx = (c.t * g.X[2] + c.z) * Xh + c.xh; } }
    hop_term<F, RECON, 1, true, MULTI>(o, A, c, n, cross, fidx, (F)1);
    cross = false; n = c.idx - Xh;
    if (c.y == 0) { n = c.idx + (g.X[1] - 1) * Xh; if (MULTI && g.part[1]) { cross = true; fidx = (c.t * g.X[2] + c.z) * Xh + c.xh; } }
    hop_term<F, RECON, 1, false, MULTI>(o, A, c, n, cross, fidx, (F)1);
  }
  // +-z
  {
    const int sz = g.X[1] * Xh;
    bool cross = false; int fidx = 0;
    int n = c.idx + sz;
    if (c.z == g.X[2] - 1) { n = c.idx - (g.X[2] - 1) * sz; if (MULTI && g.part[2]) { cross = true; fidx = (c.t * g.X[1] + c.y) * Xh + c.xh; } }
    hop_term<F, RECON, 2, true, MULTI>(o, A, c, n, cross, fidx, (F)1);
    cross = false; n = c.idx - sz;
    if (c.z == 0) { n = c.idx + (g.X[2] - 1) * sz; if (MULTI && g.part[2]) { cross = true; fidx = (c.t * g.X[1] + c.y) * Xh + c.xh; } }
    hop_term<F, RECON, 2, false, MULTI>(o, A, c, n, cross, fidx, (F)1);
  }
  // +-t
  {
    const int st = g.X[2] * g.X[1] * Xh;
    bool cross = false; int fidx = 0;
    int n = c.idx + st;
    if (c.t == g.X[3] - 1) { n = c.idx - (g.X[3] - 1) * st; if (MULTI && g.part[3]) { cross = true; fidx = c.idx - (g.X[3] - 1) * st; } }
    hop_term<F, RECON, 3, true, MULTI>(o, A, c, n, cross, fidx, s12_f);
    cross = false; n = c.idx - st;
    if (c.t == 0) { n = c.idx + (g.X[3] - 1) * st; if (MULTI && g.part[3]) { cross = true; fidx = c.idx; } }
    hop_term<F, RECON, 3, false, MULTI>(o, A, c, n, cross, fidx, s12_b);
  }

  // ---- twisted-clover epilogue: the site-constant twist matrices become the site's 6x6 chiral blocks ----
  if constexpr (CLOVER) {
    double redc = 0.0;
    if (T::TW1) clover_mul(o, A.cl_inv, c.idx, stride, A.cl_dag1 != 0);
    if (T::XTERM) {
      Spinor<F> x;
#pragma unroll
      for (int j = 0; j < 6; j++) unpack_vec(x, j, A.x[(size_t)j * stride + c.idx]);
      if (T::TWX && !A.cl_plain_x) {
        // (C + i ax g5) x
        Spinor<F> cx = x;
        clover_mul(cx, A.cl_c, c.idx, stride, false);
#pragma unroll
        for (int s = 0; s < 4; s++)
#pragma unroll
          for (int cc = 0; cc < 3; cc++) {
            cx.v[s][cc][0] -= A.e.ax * x.v[s ^ 2][cc][1];
            cx.v[s][cc][1] += A.e.ax * x.v[s ^ 2][cc][0];
          }
        x = cx;
      }
#pragma unroll
      for (int s = 0; s < 4; s++)
#pragma unroll
        for (int cc = 0; cc < 3; cc++) {
          o.v[s][cc][0] = x.v[s][cc][0] + A.e.k * o.v[s][cc][0];
          o.v[s][cc][1] = x.v[s][cc][1] + A.e.k * o.v[s][cc][1];
        }
    }
    if (T::RED == 1) {
#pragma unroll
      for (int j = 0; j < 6; j++) {
        const VecT<F> yv = pack_vec(o, j);
        redc += norm2_vec(yv);
        if (A.out2) A.out2[(size_t)j * stride + c.idx] = yv;
      }
    }
    if (T::TW3) clover_mul(o, A.cl_inv, c.idx, stride, A.cl_dag3 != 0);
#pragma unroll
    for (int j = 0; j < 6; j++) {
      VecT<F> zv = pack_vec(o, j);
      if (T::RED == 2) {
        VecT<F> rv = A.r[(size_t)j * stride + c.idx];
        rv.a -= alpha * zv.a; rv.b -= alpha * zv.b; rv.c -= alpha * zv.c; rv.d -= alpha * zv.d;
        redc += norm2_vec(rv);
        A.r[(size_t)j * stride + c.idx] = rv;
      } else if (T::CHEB) {
        const VecT<F> yv = A.y[(size_t)j * stride + c.idx];
        zv.a = A.e.d1 * zv.a + A.e.d2 * yv.a; zv.b = A.e.d1 * zv.b + A.e.d2 * yv.b;
        zv.c = A.e.d1 * zv.c + A.e.d2 * yv.c; zv.d = A.e.d1 * zv.d + A.e.d2 * yv.d;
        if (A.e.d3 != (F)0) {
          const VecT<F> rv = A.r[(size_t)j * stride + c.idx];
          zv.a += A.e.d3 * rv.a; zv.b += A.e.d3 * rv.b; zv.c += A.e.d3 * rv.c; zv.d += A.e.d3 * rv.d;
        }
        A.out[(size_t)j * stride + c.idx] = zv;
      } else {
        A.out[(size_t)j * stride + c.idx] = zv;
      }
    }
    return redc;
  }

  // ---- epilogue ----
  // Processed in three (vector j, vector j+3) pairs -- gamma5 pairs spin s with s^2, i.e. vector j with j+3 -- so that
  // only 16 extra values (one pair of x / y / r vectors) are live at a time instead of a whole second spinor; the
  // compiler fences keep the three pairs from being batched again.
  double red = 0.0;
#pragma unroll
  for (int jp = 0; jp < 3; jp++) {
    VecT<F> ou = pack_vec(o, jp), ol = pack_vec(o, jp + 3);
    if (T::TW1) twist_pair(ou, ol, A.e.c1, A.e.a1);
    if (T::XTERM) {
      VecT<F> xu = A.x[(size_t)jp * stride + c.idx], xl = A.x[(size_t)(jp + 3) * stride + c.idx];
      if (T::TWX) twist_pair(xu, xl, A.e.cx, A.e.ax);
      ou.a = xu.a + A.e.k * ou.a; ou.b = xu.b + A.e.k * ou.b; ou.c = xu.c + A.e.k * ou.c; ou.d = xu.d + A.e.k * ou.d;
      ol.a = xl.a + A.e.k * ol.a; ol.b = xl.b + A.e.k * ol.b; ol.c = xl.c + A.e.k * ol.c; ol.d = xl.d + A.e.k * ol.d;
    }
    if (T::RED == 1) red += norm2_vec(ou) + norm2_vec(ol);
    if (T::TW3) twist_pair(ou, ol, A.e.c3, A.e.a3);
    if (T::RED == 2) {
      // r <- r - alpha z ; |r|^2
      VecT<F> ru = A.r[(size_t)jp * stride + c.idx], rl = A.r[(size_t)(jp + 3) * stride + c.idx];
      ru.a -= alpha * ou.a; ru.b -= alpha * ou.b; ru.c -= alpha * ou.c; ru.d -= alpha * ou.d;
      rl.a -= alpha * ol.a; rl.b -= alpha * ol.b; rl.c -= alpha * ol.c; rl.d -= alpha * ol.d;
      red += norm2_vec(ru) + norm2_vec(rl);
      A.r[(size_t)jp * stride + c.idx] = ru; A.r[(size_t)(jp + 3) * stride + c.idx] = rl;
    } else if (T::CHEB) {
      // three-term recurrence of the Chebyshev-accelerated operator (reference lib/qudaQKXTM_Deflation.cpp:1040-1056):
      // out = d1 (M^dag M y) + d2 y + d3 r, with z = M^dag M y formed above; out may alias r (site-local)
      const VecT<F> yu = A.y[(size_t)jp * stride + c.idx], yl = A.y[(size_t)(jp + 3) * stride + c.idx];
      ou.a = A.e.d1 * ou.a + A.e.d2 * yu.a; ou.b = A.e.d1 * ou.b + A.e.d2 * yu.b;
      ou.c = A.e.d1 * ou.c + A.e.d2 * yu.c; ou.d = A.e.d1 * ou.d + A.e.d2 * yu.d;
      ol.a = A.e.d1 * ol.a + A.e.d2 * yl.a; ol.b = A.e.d1 * ol.b + A.e.d2 * yl.b;
      ol.c = A.e.d1 * ol.c + A.e.d2 * yl.c; ol.d = A.e.d1 * ol.d + A.e.d2 * yl.d;
      if (A.e.d3 != (F)0) {
        const VecT<F> ru = A.r[(size_t)jp * stride + c.idx], rl = A.r[(size_t)(jp + 3) * stride + c.idx];
        ou.a += A.e.d3 * ru.a; ou.b += A.e.d3 * ru.b; ou.c += A.e.d3 * ru.c; ou.d += A.e.d3 * ru.d;
        ol.a += A.e.d3 * rl.a; ol.b += A.e.d3 * rl.b; ol.c += A.e.d3 * rl.c; ol.d += A.e.d3 * rl.d;
      }
      A.out[(size_t)jp * stride + c.idx] = ou; A.out[(size_t)(jp + 3) * stride + c.idx] = ol;
    } else {
      A.out[(size_t)jp * stride + c.idx] = ou; A.out[(size_t)(jp + 3) * stride + c.idx] = ol;
    }
#if defined(__CUDA_ARCH__)
    if (T::XTERM) asm volatile("" ::: "memory");
#endif
  }
  return red;
}

}  // namespace tmq
