// tmq_clover.cu -- the twisted-clover variant of the path (reference call sites
// lib/qudaQKXTM_interface.cpp:1750-1756, qkxtm/MG_Bench.cpp:243-251,605-608: loadCloverQuda(NULL, NULL, &inv_param) with
// clover_coeff = csw * kappa, i.e. upstream builds the clover field on the device from the resident gauge field).
//
// Operator (kappa normalisation): M_full = A - kappa D,  A = C + i a g5,  a = 2 kappa mu,
//     C(x) = 1 + i (csw kappa) sum_{mu<nu} sigma_munu F_munu(x),   sigma_munu = (i/2)[g_mu, g_nu],
//     F_munu = (Q_munu - Q_munu^dag) / 8,  Q_munu = sum of the four plaquette leaves in the mu-nu plane around x
// (the Sheikholeslami-Wohlert term D_W + csw (i/4) sigma_munu F_munu divided by 4 + m0 = 1/(2 kappa); written from the
// definition, upstream's clover code is not vendored -- "parity unpinned" like the hop).
//
// Storage (B200-first, not QUDA's packed order): in the chiral basis chi^(+-)_s = psi_s +- psi_{s+2} (gamma5 = spin swap in
// the UKQCD basis, apply_gamma5_vector_core.h) sigma_munu is block diagonal, so C = diag(C+, C-) with two 6x6 blocks per
// site and A^-1 = diag((C+ + i a)^-1, (C- - i a)^-1); A^-dag is the conjugate transpose of the same numbers.  Both are kept
// as full complex 6x6 blocks in SoA form vec[2 parity][36][Vh] (one 32-byte vector = 2 complex; coalesced like the spinors);
// the Dslash epilogues rotate to the chiral basis in registers (tmq_site.cuh: clover_mul).
#include <math.h>
#include <string.h>
#include <vector>
#include "../../include/tmq.h"
#include "tmq_internal.h"
#include "tmq_site.cuh"

namespace tmq {

struct SigmaConst {
  // chiral blocks of sigma_munu, plane order (0,1)(0,2)(0,3)(1,2)(1,3)(2,3): s[plane][block][row][col][re,im]
  double s[6][2][2][2][2];
};

// UKQCD gamma matrices (reference lib/code_pieces/gammas_tm_base.h:21-32), sigma_munu = (i/2)[g_mu, g_nu] and their
// projection on the eigenspaces of gamma5 = spin swap
static void build_sigma(SigmaConst &S) {
  typedef double C2[2];
  double g[4][4][4][2];
  memset(g, 0, sizeof(g));
  auto set = [&](int mu, int r, int c, double re, double im) { g[mu][r][c][0] = re; g[mu][r][c][1] = im; };
  set(0, 0, 3, 0, 1); set(0, 1, 2, 0, 1); set(0, 2, 1, 0, -1); set(0, 3, 0, 0, -1);
  set(1, 0, 3, 1, 0); set(1, 1, 2, -1, 0); set(1, 2, 1, -1, 0); set(1, 3, 0, 1, 0);
  set(2, 0, 2, 0, 1); set(2, 1, 3, 0, -1); set(2, 2, 0, 0, -1); set(2, 3, 1, 0, 1);
  set(3, 0, 0, 1, 0); set(3, 1, 1, 1, 0); set(3, 2, 2, -1, 0); set(3, 3, 3, -1, 0);
  (void)sizeof(C2);
  // T: chi = T psi, rows (e0+e2, e1+e3, e0-e2, e1-e3)/sqrt(2)
  const double h = 1.0 / sqrt(2.0);
  double T[4][4] = {{h, 0, h, 0}, {0, h, 0, h}, {h, 0, -h, 0}, {0, h, 0, -h}};
  int plane = 0;
  for (int mu = 0; mu < 4; mu++)
    for (int nu = mu + 1; nu < 4; nu++, plane++) {
      double sg[4][4][2];
      for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
          double re = 0, im = 0;   // [g_mu, g_nu]
          for (int k = 0; k < 4; k++) {
            re += g[mu][r][k][0] * g[nu][k][c][0] - g[mu][r][k][1] * g[nu][k][c][1];
            im += g[mu][r][k][0] * g[nu][k][c][1] + g[mu][r][k][1] * g[nu][k][c][0];
            re -= g[nu][r][k][0] * g[mu][k][c][0] - g[nu][r][k][1] * g[mu][k][c][1];
            im -= g[nu][r][k][0] * g[mu][k][c][1] + g[nu][r][k][1] * g[mu][k][c][0];
          }
          sg[r][c][0] = -0.5 * im;   // (i/2) (re + i im)
          sg[r][c][1] = 0.5 * re;
        }
      // T sg T^T
      double t[4][4][2];
      for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
          double re = 0, im = 0;
          for (int k = 0; k < 4; k++)
            for (int l = 0; l < 4; l++) { re += T[r][k] * sg[k][l][0] * T[c][l]; im += T[r][k] * sg[k][l][1] * T[c][l]; }
          t[r][c][0] = re; t[r][c][1] = im;
        }
      for (int b = 0; b < 2; b++)
        for (int r = 0; r < 2; r++)
          for (int c = 0; c < 2; c++) { S.s[plane][b][r][c][0] = t[2 * b + r][2 * b + c][0]; S.s[plane][b][r][c][1] = t[2 * b + r][2 * b + c][1]; }
    }
}

struct M3 { double u[3][3][2]; };
__device__ __forceinline__ void m3_mul(M3 &c, const M3 &a, const M3 &b) {        // c = a b
  M3 t;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double re = 0, im = 0;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        re += a.u[i][k][0] * b.u[k][j][0] - a.u[i][k][1] * b.u[k][j][1];
        im += a.u[i][k][0] * b.u[k][j][1] + a.u[i][k][1] * b.u[k][j][0];
      }
      t.u[i][j][0] = re; t.u[i][j][1] = im;
    }
  c = t;
}
__device__ __forceinline__ void m3_dag(M3 &c, const M3 &a) {
  M3 t;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) { t.u[i][j][0] = a.u[j][i][0]; t.u[i][j][1] = -a.u[j][i][1]; }
  c = t;
}

struct Coord4 { int x[4]; };

// gauge field with a one-site halo in the partitioned dimensions (z, t): the clover leaves reach x +- mu +- nu, i.e. across
// rank boundaries and their corners.  Plain fp64 links, [site_ext][mu][3][3][re,im], site_ext lexicographic over
// (x, y, z + hz, t + ht); built once per tmq_clover_load and freed afterwards.
struct ExtGauge {
  const double *d;      // nullptr: single rank without forced partition -> periodic wrap on the native field
  int h[4];             // halo width per dimension (0 | 1)
  int E[4];             // extended extents
};
__device__ __forceinline__ size_t ext_site(const ExtGauge &x, const int c[4]) {
  return (((size_t)(c[3] + x.h[3]) * x.E[2] + (c[2] + x.h[2])) * x.E[1] + c[1]) * x.E[0] + c[0];
}
template <int RECON>
__device__ __forceinline__ void link_at(M3 &m, const void *gauge, const ExtGauge &ext, const Geom &g, Coord4 c, int mu) {
  // periodic wrap on the local lattice where the dimension is not partitioned; halo sites otherwise
#pragma unroll
  for (int d = 0; d < 4; d++) {
    if (ext.d != nullptr && ext.h[d]) continue;
    if (c.x[d] < 0) c.x[d] += g.X[d];
    if (c.x[d] >= g.X[d]) c.x[d] -= g.X[d];
  }
  if (ext.d != nullptr) {
    const double *p = ext.d + (ext_site(ext, c.x) * 4 + mu) * 18;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) { m.u[i][j][0] = p[(i * 3 + j) * 2]; m.u[i][j][1] = p[(i * 3 + j) * 2 + 1]; }
    return;
  }
  const int parity = (c.x[0] + c.x[1] + c.x[2] + c.x[3]) & 1;
  const int idx = ((c.x[3] * g.X[2] + c.x[2]) * g.X[1] + c.x[1]) * g.Xh + (c.x[0] >> 1);
  const double s12 = (mu == 3 && g.tb_last && c.x[3] == g.X[3] - 1) ? (double)g.tb_sign : 1.0;
  Link<double> L;
  load_link<double, RECON>(L, gauge, parity, mu, idx, g.Vh, s12);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) { m.u[i][j][0] = L.u[i][j][0]; m.u[i][j][1] = L.u[i][j][1]; }
}
__device__ __forceinline__ Coord4 shifted(Coord4 c, int mu, int d) { c.x[mu] += d; return c; }

// C(x) for one site: thread = (parity, cb index)
template <int RECON>
__global__ void __launch_bounds__(128) clover_compute_kernel(VecT<double> *C, const void *gauge, ExtGauge ext, Geom g, SigmaConst S, double coeff) {
  const int e = blockIdx.x * 128 + threadIdx.x;
  if (e >= 2 * g.Vh) return;
  const int parity = e / g.Vh, idx = e - parity * g.Vh;
  Coord4 c;
  {
    int r = idx;
    const int xh = r % g.Xh; r /= g.Xh;
    c.x[1] = r % g.X[1]; r /= g.X[1];
    c.x[2] = r % g.X[2]; c.x[3] = r / g.X[2];
    c.x[0] = 2 * xh + ((c.x[1] + c.x[2] + c.x[3] + parity) & 1);
  }
  double blk[2][6][6][2];
#pragma unroll
  for (int b = 0; b < 2; b++)
    for (int r = 0; r < 6; r++)
      for (int cc = 0; cc < 6; cc++) { blk[b][r][cc][0] = (r == cc) ? 1.0 : 0.0; blk[b][r][cc][1] = 0.0; }
  int plane = 0;
  for (int mu = 0; mu < 4; mu++)
    for (int nu = mu + 1; nu < 4; nu++, plane++) {
      M3 Q, a, b2, t;
      // leaf 1: U_mu(x) U_nu(x+mu) U_mu(x+nu)^dag U_nu(x)^dag
      link_at<RECON>(a, gauge, ext, g, c, mu); link_at<RECON>(b2, gauge, ext, g, shifted(c, mu, 1), nu); m3_mul(Q, a, b2);
      link_at<RECON>(a, gauge, ext, g, shifted(c, nu, 1), mu); m3_dag(a, a); m3_mul(Q, Q, a);
      link_at<RECON>(a, gauge, ext, g, c, nu); m3_dag(a, a); m3_mul(Q, Q, a);
      // leaf 2: U_nu(x) U_mu(x-mu+nu)^dag U_nu(x-mu)^dag U_mu(x-mu)
      link_at<RECON>(t, gauge, ext, g, c, nu);
      link_at<RECON>(a, gauge, ext, g, shifted(shifted(c, mu, -1), nu, 1), mu); m3_dag(a, a); m3_mul(t, t, a);
      link_at<RECON>(a, gauge, ext, g, shifted(c, mu, -1), nu); m3_dag(a, a); m3_mul(t, t, a);
      link_at<RECON>(a, gauge, ext, g, shifted(c, mu, -1), mu); m3_mul(t, t, a);
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) { Q.u[i][j][0] += t.u[i][j][0]; Q.u[i][j][1] += t.u[i][j][1]; }
      // leaf 3: U_mu(x-mu)^dag U_nu(x-mu-nu)^dag U_mu(x-mu-nu) U_nu(x-nu)
      link_at<RECON>(t, gauge, ext, g, shifted(c, mu, -1), mu); m3_dag(t, t);
      link_at<RECON>(a, gauge, ext, g, shifted(shifted(c, mu, -1), nu, -1), nu); m3_dag(a, a); m3_mul(t, t, a);
      link_at<RECON>(a, gauge, ext, g, shifted(shifted(c, mu, -1), nu, -1), mu); m3_mul(t, t, a);
      link_at<RECON>(a, gauge, ext, g, shifted(c, nu, -1), nu); m3_mul(t, t, a);
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) { Q.u[i][j][0] += t.u[i][j][0]; Q.u[i][j][1] += t.u[i][j][1]; }
      // leaf 4: U_nu(x-nu)^dag U_mu(x-nu) U_nu(x+mu-nu) U_mu(x)^dag
      link_at<RECON>(t, gauge, ext, g, shifted(c, nu, -1), nu); m3_dag(t, t);
      link_at<RECON>(a, gauge, ext, g, shifted(c, nu, -1), mu); m3_mul(t, t, a);
      link_at<RECON>(a, gauge, ext, g, shifted(shifted(c, mu, 1), nu, -1), nu); m3_mul(t, t, a);
      link_at<RECON>(a, gauge, ext, g, c, mu); m3_dag(a, a); m3_mul(t, t, a);
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) { Q.u[i][j][0] += t.u[i][j][0]; Q.u[i][j][1] += t.u[i][j][1]; }
      // F = (Q - Q^dag)/8 ; block_b += i coeff sigma_b (x) F
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
          const double fr = 0.125 * (Q.u[i][j][0] - Q.u[j][i][0]), fi = 0.125 * (Q.u[i][j][1] + Q.u[j][i][1]);
          for (int bb = 0; bb < 2; bb++)
            for (int s = 0; s < 2; s++)
              for (int sp = 0; sp < 2; sp++) {
                const double sr = S.s[plane][bb][s][sp][0], si = S.s[plane][bb][s][sp][1];
                // i coeff (sr + i si)(fr + i fi)
                const double pr = sr * fr - si * fi, pi = sr * fi + si * fr;
                blk[bb][s * 3 + i][sp * 3 + j][0] -= coeff * pi;
                blk[bb][s * 3 + i][sp * 3 + j][1] += coeff * pr;
              }
        }
    }
  VecT<double> *out = C + (size_t)parity * 36 * g.Vh;
  for (int b = 0; b < 2; b++)
    for (int v = 0; v < 18; v++) {
      const int k0 = 2 * v, k1 = 2 * v + 1;
      VecT<double> o;
      o.a = blk[b][k0 / 6][k0 % 6][0]; o.b = blk[b][k0 / 6][k0 % 6][1];
      o.c = blk[b][k1 / 6][k1 % 6][0]; o.d = blk[b][k1 / 6][k1 % 6][1];
      out[(size_t)(b * 18 + v) * g.Vh + idx] = o;
    }
}

// Ainv block = (C_b + i a sgn_b)^-1, sgn = +1 for the gamma5 = +1 block, -1 for the other; thread = (parity, site, block)
__global__ void __launch_bounds__(128) clover_invert_kernel(VecT<double> *Ainv, const VecT<double> *C, int Vh, double a, double *err) {
  const long long e = (long long)blockIdx.x * 128 + threadIdx.x;
  if (e >= (long long)4 * Vh) return;
  const int b = (int)(e / ((long long)2 * Vh));
  const long long rest = e - (long long)b * 2 * Vh;
  const int parity = (int)(rest / Vh), idx = (int)(rest - (long long)parity * Vh);
  const VecT<double> *src = C + (size_t)parity * 36 * Vh + (size_t)b * 18 * Vh;
  VecT<double> *dst = Ainv + (size_t)parity * 36 * Vh + (size_t)b * 18 * Vh;
  double m[6][12][2];   // [A | 1]
  for (int v = 0; v < 18; v++) {
    const VecT<double> x = src[(size_t)v * Vh + idx];
    const int k0 = 2 * v, k1 = 2 * v + 1;
    m[k0 / 6][k0 % 6][0] = x.a; m[k0 / 6][k0 % 6][1] = x.b;
    m[k1 / 6][k1 % 6][0] = x.c; m[k1 / 6][k1 % 6][1] = x.d;
  }
  const double tw = b == 0 ? a : -a;
  for (int r = 0; r < 6; r++) {
    m[r][r][1] += tw;
    for (int cc = 0; cc < 6; cc++) { m[r][6 + cc][0] = (r == cc) ? 1.0 : 0.0; m[r][6 + cc][1] = 0.0; }
  }
  // Gauss-Jordan with partial pivoting
  for (int p = 0; p < 6; p++) {
    int piv = p;
    double best = m[p][p][0] * m[p][p][0] + m[p][p][1] * m[p][p][1];
    for (int r = p + 1; r < 6; r++) {
      const double v2 = m[r][p][0] * m[r][p][0] + m[r][p][1] * m[r][p][1];
      if (v2 > best) { best = v2; piv = r; }
    }
    if (!(best > 0.0)) { *err = 2.0; return; }      // singular clover block
    if (piv != p)
      for (int cc = 0; cc < 12; cc++) {
        const double tr = m[p][cc][0], ti = m[p][cc][1];
        m[p][cc][0] = m[piv][cc][0]; m[p][cc][1] = m[piv][cc][1];
        m[piv][cc][0] = tr; m[piv][cc][1] = ti;
      }
    const double ir = m[p][p][0] / best, ii = -m[p][p][1] / best;    // 1 / pivot
    for (int cc = 0; cc < 12; cc++) {
      const double xr = m[p][cc][0], xi = m[p][cc][1];
      m[p][cc][0] = xr * ir - xi * ii; m[p][cc][1] = xr * ii + xi * ir;
    }
    for (int r = 0; r < 6; r++) {
      if (r == p) continue;
      const double fr = m[r][p][0], fi = m[r][p][1];
      for (int cc = 0; cc < 12; cc++) {
        m[r][cc][0] -= fr * m[p][cc][0] - fi * m[p][cc][1];
        m[r][cc][1] -= fr * m[p][cc][1] + fi * m[p][cc][0];
      }
    }
  }
  for (int v = 0; v < 18; v++) {
    const int k0 = 2 * v, k1 = 2 * v + 1;
    VecT<double> o;
    o.a = m[k0 / 6][6 + k0 % 6][0]; o.b = m[k0 / 6][6 + k0 % 6][1];
    o.c = m[k1 / 6][6 + k1 % 6][0]; o.d = m[k1 / 6][6 + k1 % 6][1];
    dst[(size_t)v * Vh + idx] = o;
  }
}

// ---- building the extended gauge field --------------------------------------------------------------------------------
template <int RECON>
__global__ void __launch_bounds__(128) ext_fill_kernel(double *E, ExtGauge ext, const void *gauge, Geom g) {
  const int e = blockIdx.x * 128 + threadIdx.x;
  if (e >= 2 * g.Vh) return;
  const int parity = e / g.Vh, idx = e - parity * g.Vh;
  int c[4];
  {
    int r = idx;
    const int xh = r % g.Xh; r /= g.Xh;
    c[1] = r % g.X[1]; r /= g.X[1];
    c[2] = r % g.X[2]; c[3] = r / g.X[2];
    c[0] = 2 * xh + ((c[1] + c[2] + c[3] + parity) & 1);
  }
  double *dst = E + ext_site(ext, c) * 72;
  for (int mu = 0; mu < 4; mu++) {
    const double s12 = (mu == 3 && g.tb_last && c[3] == g.X[3] - 1) ? (double)g.tb_sign : 1.0;
    Link<double> L;
    load_link<double, RECON>(L, gauge, parity, mu, idx, g.Vh, s12);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) { dst[mu * 18 + (i * 3 + j) * 2] = L.u[i][j][0]; dst[mu * 18 + (i * 3 + j) * 2 + 1] = L.u[i][j][1]; }
  }
}
// z faces are strided in the extended array: gather slices z = 0 and z = Z-1 (interior t) / scatter into z = -1 and z = Z
__global__ void ext_zface_kernel(double *E, ExtGauge ext, Geom g, double *buf_lo, double *buf_hi, int scatter) {
  const size_t n = (size_t)g.X[0] * g.X[1] * g.X[3] * 72;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % 72);
    size_t r = i / 72;
    int c[4];
    c[0] = (int)(r % g.X[0]); r /= g.X[0];
    c[1] = (int)(r % g.X[1]); c[3] = (int)(r / g.X[1]);
    if (!scatter) {
      c[2] = 0; buf_lo[i] = E[ext_site(ext, c) * 72 + k];
      c[2] = g.X[2] - 1; buf_hi[i] = E[ext_site(ext, c) * 72 + k];
    } else {
      c[2] = -1; E[ext_site(ext, c) * 72 + k] = buf_lo[i];        // from the backward neighbour's z = Z-1
      c[2] = g.X[2]; E[ext_site(ext, c) * 72 + k] = buf_hi[i];    // from the forward neighbour's z = 0
    }
  }
}

// site-local application on one parity block: out = M in / M^dag in, plus i a g5 in when `a` != 0 (A = C + i a g5)
template <typename F>
__global__ void __launch_bounds__(128) clover_apply_kernel(VecT<F> *out, const VecT<F> *in, const VecT<F> *M, int Vh, int dag, F a) {
  const int idx = blockIdx.x * 128 + threadIdx.x;
  if (idx >= Vh) return;
  Spinor<F> x, y;
#pragma unroll
  for (int j = 0; j < 6; j++) unpack_vec(x, j, in[(size_t)j * Vh + idx]);
  y = x;
  clover_mul(y, M, idx, Vh, dag != 0);
  if (a != (F)0) {
#pragma unroll
    for (int s = 0; s < 4; s++)
#pragma unroll
      for (int c = 0; c < 3; c++) { y.v[s][c][0] -= a * x.v[s ^ 2][c][1]; y.v[s][c][1] += a * x.v[s ^ 2][c][0]; }
  }
#pragma unroll
  for (int j = 0; j < 6; j++) out[(size_t)j * Vh + idx] = pack_vec(y, j);
}

cudaError_t clover_apply(int prec, void *out, const void *in, const void *M, int Vh, int dag, double a, cudaStream_t st) {
  const int grid = (Vh + 127) / 128;
  if (prec == 8) clover_apply_kernel<double><<<grid, 128, 0, st>>>((VecT<double> *)out, (const VecT<double> *)in, (const VecT<double> *)M, Vh, dag, a);
  else clover_apply_kernel<float><<<grid, 128, 0, st>>>((VecT<float> *)out, (const VecT<float> *)in, (const VecT<float> *)M, Vh, dag, (float)a);
  return cudaGetLastError();
}

// (re)build A^-1 for the current kappa, mu from the resident C
int clover_update_inverse(tmq_ctx *c) {
  if (!c->clover_on) return 0;
  const double a = 2.0 * c->kappa * c->mu;
  if (c->clov_inv_valid && c->clov_inv_a == a) return 0;
  const long long n = (long long)4 * c->g.Vh;
  clover_invert_kernel<<<(unsigned int)((n + 127) / 128), 128, 0, c->stream>>>((VecT<double> *)c->clov_inv_d.d, (const VecT<double> *)c->clov_c_d.d, c->g.Vh, a,
                                                                             c->scal + SC_ERR);
  TMQ_CUDA(cudaGetLastError()); c->launches++;
  TMQ_CUDA(blas_copy(c->clov_inv_s.d, 4, c->clov_inv_d.d, 8, (size_t)72 * c->g.Vh, c->stream)); c->launches++;
  double e = 0;
  TMQ_TRY(fetch_scal(c, SC_ERR, 1, &e));
  if (e != 0.0) {
    const double zero = 0.0;
    cudaMemcpyAsync(c->scal + SC_ERR, &zero, sizeof(double), cudaMemcpyHostToDevice, c->stream);
    cudaStreamSynchronize(c->stream);
    set_error("singular clover block: C + i a gamma5 cannot be inverted");
    return 1;
  }
  c->clov_inv_a = a; c->clov_inv_valid = true;
  return 0;
}

}  // namespace tmq

using namespace tmq;

extern "C" {

int tmq_clover_load(tmq_ctx *c, double clover_coeff) {
  TMQ_REQUIRE(c, "null context");
  TMQ_REQUIRE(c->gauge_d.d != nullptr, "no gauge field loaded (tmq_gauge_load): the clover field is built from it");
  TMQ_CUDA(cudaSetDevice(c->device));
  const size_t nv = (size_t)72 * c->g.Vh;     // vectors per field: 2 parities x 36
  GaugeStore *st[4] = {&c->clov_c_d, &c->clov_inv_d, &c->clov_c_s, &c->clov_inv_s};
  for (int i = 0; i < 4; i++) {
    const size_t bytes = nv * (i < 2 ? 32 : 16);
    if (st[i]->bytes != bytes) {
      if (st[i]->d) { TMQ_CUDA(cudaStreamSynchronize(c->stream)); TMQ_CUDA(cudaFree(st[i]->d)); st[i]->d = nullptr; }
      TMQ_CUDA(cudaMalloc(&st[i]->d, bytes));
      st[i]->bytes = bytes;
    }
  }
  SigmaConst S;
  build_sigma(S);
  const unsigned int grid = (unsigned int)((2 * (size_t)c->g.Vh + 127) / 128);
  // sharded lattice: the leaves need the neighbours' links (and the z-t corners): build a gauge field with a one-site halo
  // in the partitioned dimensions -- z faces first (packed), then whole t slices of the extended array, which carry the z
  // halo rows along and so fill the corners
  ExtGauge ext;
  memset(&ext, 0, sizeof(ext));
  double *E = nullptr, *zbuf = nullptr;
  if (c->multi) {
    const Geom &g = c->g;
    for (int d = 0; d < 4; d++) { ext.h[d] = g.part[d] ? 1 : 0; ext.E[d] = g.X[d] + 2 * ext.h[d]; }
    const size_t next = (size_t)ext.E[0] * ext.E[1] * ext.E[2] * ext.E[3];
    TMQ_CUDA(cudaMalloc((void **)&E, next * 72 * sizeof(double)));
    TMQ_CUDA(cudaMemsetAsync(E, 0, next * 72 * sizeof(double), c->stream));
    if (c->recon == 8) ext_fill_kernel<8><<<grid, 128, 0, c->stream>>>(E, ext, c->gauge_d.d, g);
    else if (c->recon == 12) ext_fill_kernel<12><<<grid, 128, 0, c->stream>>>(E, ext, c->gauge_d.d, g);
    else ext_fill_kernel<18><<<grid, 128, 0, c->stream>>>(E, ext, c->gauge_d.d, g);
    TMQ_CUDA(cudaGetLastError()); c->launches++;
    if (g.part[2]) {
      const size_t nz = (size_t)g.X[0] * g.X[1] * g.X[3] * 72;
      TMQ_CUDA(cudaMalloc((void **)&zbuf, 4 * nz * sizeof(double)));
      double *s_lo = zbuf, *s_hi = zbuf + nz, *r_lo = zbuf + 2 * nz, *r_hi = zbuf + 3 * nz;
      ext_zface_kernel<<<blas_grid(), 256, 0, c->stream>>>(E, ext, g, s_lo, s_hi, 0);
      TMQ_CUDA(cudaGetLastError()); c->launches++;
      // my z = 0 slice goes to the backward neighbour (its z = Z halo), my z = Z-1 slice to the forward neighbour (its z = -1 halo)
      TMQ_TRY(comm_sendrecv_dim(c, 2, s_lo, s_hi, /*from fwd: its z = 0*/ r_hi, /*from bwd: its z = Z-1*/ r_lo, nz * sizeof(double), c->stream));
      ext_zface_kernel<<<blas_grid(), 256, 0, c->stream>>>(E, ext, g, r_lo, r_hi, 1);
      TMQ_CUDA(cudaGetLastError()); c->launches++;
    }
    if (g.part[3]) {
      const size_t slice = (size_t)ext.E[0] * ext.E[1] * ext.E[2] * 72;       // doubles per extended t slice
      double *t_first = E + slice * 1, *t_last = E + slice * (size_t)g.X[3];     // interior t = 0 and t = T-1 (ext index t + 1)
      double *halo_lo = E, *halo_hi = E + slice * (size_t)(g.X[3] + 1);
      TMQ_TRY(comm_sendrecv_dim(c, 3, t_first, t_last, /*from fwd: its t = 0*/ halo_hi, /*from bwd: its t = T-1*/ halo_lo,
                                slice * sizeof(double), c->stream));
    }
    ext.d = E;
  }
  if (c->recon == 8) clover_compute_kernel<8><<<grid, 128, 0, c->stream>>>((VecT<double> *)c->clov_c_d.d, c->gauge_d.d, ext, c->g, S, clover_coeff);
  else if (c->recon == 12) clover_compute_kernel<12><<<grid, 128, 0, c->stream>>>((VecT<double> *)c->clov_c_d.d, c->gauge_d.d, ext, c->g, S, clover_coeff);
  else clover_compute_kernel<18><<<grid, 128, 0, c->stream>>>((VecT<double> *)c->clov_c_d.d, c->gauge_d.d, ext, c->g, S, clover_coeff);
  TMQ_CUDA(cudaGetLastError()); c->launches++;
  if (E) { TMQ_CUDA(cudaStreamSynchronize(c->stream)); cudaFree(E); if (zbuf) cudaFree(zbuf); }
  TMQ_CUDA(blas_copy(c->clov_c_s.d, 4, c->clov_c_d.d, 8, nv, c->stream)); c->launches++;
  c->clover_on = true; c->clover_coeff = clover_coeff; c->clov_inv_valid = false;
  if (c->op_set) TMQ_TRY(clover_update_inverse(c));
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int tmq_clover_free(tmq_ctx *c) {
  if (!c) return 0;
  GaugeStore *st[4] = {&c->clov_c_d, &c->clov_inv_d, &c->clov_c_s, &c->clov_inv_s};
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (int i = 0; i < 4; i++) { if (st[i]->d) cudaFree(st[i]->d); st[i]->d = nullptr; st[i]->bytes = 0; }
  c->clover_on = false; c->clov_inv_valid = false;
  return 0;
}

}  // extern "C"
