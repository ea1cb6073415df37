/*
 * tm_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see tm_oracle.h for scope and for the
 * "parity unpinned" statement).  Plain C99 + OpenMP.  Every function cites the reference
 * file:line whose convention it restates; bodies marked [upstream-shape] follow the published
 * structure of QUDA's host reference (tests/wilson_dslash_reference.cpp, tests/dslash_util.h,
 * tests/blas_reference.cpp), which is not in /root/reference.
 */
#include "tm_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SS 24 /* spinorSiteSize, include/QKXTM_util.h:7 */
#define GS 18 /* gaugeSiteSize,  include/QKXTM_util.h:6 */

static int Z[4];
static int V, Vh;

/* projector table P[2mu + s][row][col][re,im]; 2mu = 1-gamma_mu, 2mu+1 = 1+gamma_mu */
static double PROJ[8][4][4][2];
static double G5[4][4][2];
static double GAM[4][4][4][2];   /* gamma_mu as set by orc_set_gamma (sigma_munu of the clover term) */

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
/* launchers such as torchrun export OMP_NUM_THREADS=1 to every worker: the timed CPU baseline sets its thread count itself */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---- geometry: qkxtm/QKXTM_util.cpp:94-128 (setDims), :418-442, :455-470, :191-197 ---------- */
void orc_set_lattice(const int X[4]) {
  V = 1;
  for (int d = 0; d < 4; d++) { Z[d] = X[d]; V *= X[d]; }
  Vh = V / 2;
}
int orc_volume(void) { return V; }

int orc_full_lattice_index(int i, int oddBit) {
  int X1h = Z[0] / 2;
  int za = i / X1h;
  int zb = za / Z[1];
  int x2 = za - zb * Z[1];
  int x4 = zb / Z[2];
  int x3 = zb - x4 * Z[2];
  int x1odd = (x2 + x3 + x4 + oddBit) & 1;
  return 2 * i + x1odd;
}

int orc_neighbor_index(int i, int oddBit, int dx4, int dx3, int dx2, int dx1) {
  int Y = orc_full_lattice_index(i, oddBit);
  int x4 = Y / (Z[2] * Z[1] * Z[0]);
  int x3 = (Y / (Z[1] * Z[0])) % Z[2];
  int x2 = (Y / Z[0]) % Z[1];
  int x1 = Y % Z[0];
  x4 = (x4 + dx4 + Z[3]) % Z[3];
  x3 = (x3 + dx3 + Z[2]) % Z[2];
  x2 = (x2 + dx2 + Z[1]) % Z[1];
  x1 = (x1 + dx1 + Z[0]) % Z[0];
  return (x4 * (Z[2] * Z[1] * Z[0]) + x3 * (Z[1] * Z[0]) + x2 * Z[0] + x1) / 2;
}

int orc_odd_bit(int Y) {
  int x4 = Y / (Z[2] * Z[1] * Z[0]);
  int x3 = (Y / (Z[1] * Z[0])) % Z[2];
  int x2 = (Y / Z[0]) % Z[1];
  int x1 = Y % Z[0];
  return (x4 + x3 + x2 + x1) % 2;
}

/* ---- gamma basis: caller supplies gamma_mu (UKQCD values: gammas_tm_base.h:21-32) ------------ */
static void cmat_mul(double c[4][4][2], double a[4][4][2], double b[4][4][2]) {
  double t[4][4][2];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double re = 0, im = 0;
      for (int k = 0; k < 4; k++) {
        re += a[i][k][0] * b[k][j][0] - a[i][k][1] * b[k][j][1];
        im += a[i][k][0] * b[k][j][1] + a[i][k][1] * b[k][j][0];
      }
      t[i][j][0] = re; t[i][j][1] = im;
    }
  memcpy(c, t, sizeof(t));
}

void orc_set_gamma(const double g[4][4][4][2]) {
  double gm[4][4][4][2];
  memcpy(gm, g, sizeof(gm));
  memcpy(GAM, g, sizeof(GAM));
  for (int mu = 0; mu < 4; mu++)
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) {
        double id = (i == j) ? 1.0 : 0.0;
        PROJ[2 * mu][i][j][0] = id - gm[mu][i][j][0];
        PROJ[2 * mu][i][j][1] = -gm[mu][i][j][1];
        PROJ[2 * mu + 1][i][j][0] = id + gm[mu][i][j][0];
        PROJ[2 * mu + 1][i][j][1] = gm[mu][i][j][1];
      }
  /* gamma5 = gamma_x gamma_y gamma_z gamma_t; in UKQCD this is the spin swap of
   * apply_gamma5_vector_core.h:1-16 (checked in tests) */
  cmat_mul(G5, gm[0], gm[1]);
  cmat_mul(G5, G5, gm[2]);
  cmat_mul(G5, G5, gm[3]);
}

void orc_get_gamma5(double g5[4][4][2]) { memcpy(g5, G5, sizeof(G5)); }

/* ---- gauge helpers ---------------------------------------------------------------------------- */

/* qkxtm/QKXTM_util.cpp:698-705: multiply U_t on the last time slice (both parities) by sign */
void orc_apply_t_boundary(double *gauge[4], int sign) {
  for (int j = (Z[0] / 2) * Z[1] * Z[2] * (Z[3] - 1); j < Vh; j++)
    for (int i = 0; i < GS; i++) {
      gauge[3][j * GS + i] *= sign;
      gauge[3][(Vh + j) * GS + i] *= sign;
    }
}

/* qkxtm/QKXTM_util.cpp:240-245 */
static void acc_conj_prod(double *a, const double *b, const double *c, int sign) {
  a[0] += sign * (b[0] * c[0] - b[1] * c[1]);
  a[1] -= sign * (b[0] * c[1] + b[1] * c[0]);
}

/* qkxtm/QKXTM_util.cpp:281-295: w = (u x v)^* scaled by u0 (u0 carries the T-boundary sign) */
void orc_su3_reconstruct12(double *mat, double u0) {
  double *u = mat, *v = mat + 6, *w = mat + 12;
  for (int n = 0; n < 6; n++) w[n] = 0.0;
  acc_conj_prod(w + 0, u + 2, v + 4, +1);
  acc_conj_prod(w + 0, u + 4, v + 2, -1);
  acc_conj_prod(w + 2, u + 4, v + 0, +1);
  acc_conj_prod(w + 2, u + 0, v + 4, -1);
  acc_conj_prod(w + 4, u + 0, v + 2, +1);
  acc_conj_prod(w + 4, u + 2, v + 0, -1);
  for (int n = 0; n < 6; n++) w[n] *= u0;
}

/* c = a b (3x3 complex row-major) */
static void su3_mm(double *c, const double *a, const double *b) {
  double t[18];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double re = 0, im = 0;
      for (int k = 0; k < 3; k++) {
        const double *x = a + (i * 3 + k) * 2, *y = b + (k * 3 + j) * 2;
        re += x[0] * y[0] - x[1] * y[1];
        im += x[0] * y[1] + x[1] * y[0];
      }
      t[(i * 3 + j) * 2] = re; t[(i * 3 + j) * 2 + 1] = im;
    }
  memcpy(c, t, sizeof(t));
}
/* c = a b^dag */
static void su3_mmd(double *c, const double *a, const double *b) {
  double t[18];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double re = 0, im = 0;
      for (int k = 0; k < 3; k++) {
        const double *x = a + (i * 3 + k) * 2, *y = b + (j * 3 + k) * 2;
        re += x[0] * y[0] + x[1] * y[1];
        im += -x[0] * y[1] + x[1] * y[0];
      }
      t[(i * 3 + j) * 2] = re; t[(i * 3 + j) * 2 + 1] = im;
    }
  memcpy(c, t, sizeof(t));
}

static const double *link_at(double *gauge[4], int mu, int Y) {
  /* QDP even-odd: gauge[mu] = [even Vh | odd Vh] x 18 (qkxtm/QKXTM_util.cpp:840-857) */
  int odd = orc_odd_bit(Y);
  return gauge[mu] + ((long)odd * Vh + Y / 2) * GS;
}

static int shift_lex(int Y, int mu, int d) {
  int x[4];
  x[0] = Y % Z[0]; x[1] = (Y / Z[0]) % Z[1]; x[2] = (Y / (Z[0] * Z[1])) % Z[2]; x[3] = Y / (Z[0] * Z[1] * Z[2]);
  x[mu] = (x[mu] + d + Z[mu]) % Z[mu];
  return ((x[3] * Z[2] + x[2]) * Z[1] + x[1]) * Z[0] + x[0];
}

/* Sum over sites and the 6 planes of Re tr[U_mu(x) U_nu(x+mu) U_mu(x+nu)^dag U_nu(x)^dag],
 * normalised by V*3*6 (lib/code_pieces/plaquette_core.h, lib/qudaQKXTM_kernels.cu:957). */
double orc_plaquette(double *gauge[4]) {
  double sum = 0.0;
#pragma omp parallel for reduction(+ : sum)
  for (int Y = 0; Y < V; Y++) {
    for (int mu = 0; mu < 4; mu++)
      for (int nu = mu + 1; nu < 4; nu++) {
        double a[18], b[18];
        su3_mm(a, link_at(gauge, mu, Y), link_at(gauge, nu, shift_lex(Y, mu, 1)));
        su3_mm(b, link_at(gauge, nu, Y), link_at(gauge, mu, shift_lex(Y, nu, 1)));
        double p[18];
        su3_mmd(p, a, b);
        sum += p[0] + p[8] + p[16];
      }
  }
  return sum / ((double)V * 3 * 6);
}

/* ---- hop [upstream-shape: dslashReference in tests/wilson_dslash_reference.cpp] --------------- */

/* out[s] = sum_s' P[s][s'] in[s'] for each colour (multiplySpinorByDiracProjector) */
static void project(double *out, int proj, const double *in) {
  for (int s = 0; s < 4; s++)
    for (int c = 0; c < 3; c++) {
      double re = 0, im = 0;
      for (int t = 0; t < 4; t++) {
        double pr = PROJ[proj][s][t][0], pi = PROJ[proj][s][t][1];
        double xr = in[(t * 3 + c) * 2], xi = in[(t * 3 + c) * 2 + 1];
        re += pr * xr - pi * xi;
        im += pr * xi + pi * xr;
      }
      out[(s * 3 + c) * 2] = re; out[(s * 3 + c) * 2 + 1] = im;
    }
}

/* y_a = sum_b U_ab x_b : index order of lib/code_pieces/core_def.h:498-531 (su3Mul) */
static void su3_mul(double *y, const double *U, const double *x) {
  for (int a = 0; a < 3; a++) {
    double re = 0, im = 0;
    for (int b = 0; b < 3; b++) {
      double ur = U[(a * 3 + b) * 2], ui = U[(a * 3 + b) * 2 + 1];
      re += ur * x[b * 2] - ui * x[b * 2 + 1];
      im += ur * x[b * 2 + 1] + ui * x[b * 2];
    }
    y[a * 2] = re; y[a * 2 + 1] = im;
  }
}
/* y_a = sum_b conj(U_ba) x_b (su3Tmul) */
static void su3_tmul(double *y, const double *U, const double *x) {
  for (int a = 0; a < 3; a++) {
    double re = 0, im = 0;
    for (int b = 0; b < 3; b++) {
      double ur = U[(b * 3 + a) * 2], ui = -U[(b * 3 + a) * 2 + 1];
      re += ur * x[b * 2] - ui * x[b * 2 + 1];
      im += ur * x[b * 2 + 1] + ui * x[b * 2];
    }
    y[a * 2] = re; y[a * 2 + 1] = im;
  }
}

static const int DX[8][4] = {
  /* {dx4,dx3,dx2,dx1} for dir 0..7 = +x,-x,+y,-y,+z,-z,+t,-t */
  {0, 0, 0, 1}, {0, 0, 0, -1}, {0, 0, 1, 0}, {0, 0, -1, 0},
  {0, 1, 0, 0}, {0, -1, 0, 0}, {1, 0, 0, 0}, {-1, 0, 0, 0}};

/* (D in)(x) = sum_mu [ (1-g_mu) U_mu(x) in(x+mu) + (1+g_mu) U_mu(x-mu)^dag in(x-mu) ]
 * (sign convention witnessed by fixSinkContractions_noether_core.h:117-137); dagger swaps the
 * two projectors.  res lives on parity oddBit, in on the other parity. */
void orc_dslash(double *res, double *gauge[4], const double *in, int oddBit, int daggerBit) {
#pragma omp parallel for
  for (int i = 0; i < Vh; i++) {
    double acc[SS];
    for (int k = 0; k < SS; k++) acc[k] = 0.0;
    for (int dir = 0; dir < 8; dir++) {
      int mu = dir / 2, back = dir & 1;
      int nb = orc_neighbor_index(i, oddBit, DX[dir][0], DX[dir][1], DX[dir][2], DX[dir][3]);
      /* gaugeLink: forward -> U_mu(x) in this parity's block; backward -> U_mu(x-mu) in the
       * other parity's block at the neighbour's index */
      const double *U = back ? gauge[mu] + ((long)(1 - oddBit) * Vh + nb) * GS
                             : gauge[mu] + ((long)oddBit * Vh + i) * GS;
      const double *psi = in + (long)nb * SS;
      int proj = 2 * mu + ((dir + daggerBit) & 1);
      double h[SS], uh[SS];
      project(h, proj, psi);
      for (int s = 0; s < 4; s++) {
        if (back) su3_tmul(uh + s * 6, U, h + s * 6);
        else      su3_mul(uh + s * 6, U, h + s * 6);
      }
      for (int k = 0; k < SS; k++) acc[k] += uh[k];
    }
    for (int k = 0; k < SS; k++) res[(long)i * SS + k] = acc[k];
  }
}

/* [upstream-shape: twistGamma5].  A = 1 + i a gamma5, a = 2 kappa mu (sign of mu carries the
 * flavour: lib/qudaQKXTM_interface.cpp:504,515).  INVERSE: a -> -a, b = 1/(1+a^2); dagger: a -> -a */
void orc_twist_gamma5(double *out, const double *in, int daggerBit, double kappa, double mu,
                      int inverse, int nsites) {
  double a = 2.0 * kappa * mu, b = 1.0;
  if (inverse) { b = 1.0 / (1.0 + a * a); a = -a; }
  if (daggerBit) a = -a;
#pragma omp parallel for
  for (int i = 0; i < nsites; i++) {
    const double *x = in + (long)i * SS;
    double t[SS];
    for (int s = 0; s < 4; s++)
      for (int c = 0; c < 3; c++) {
        /* (gamma5 x)_s */
        double gr = 0, gi = 0;
        for (int u = 0; u < 4; u++) {
          double pr = G5[s][u][0], pi = G5[s][u][1];
          double xr = x[(u * 3 + c) * 2], xi = x[(u * 3 + c) * 2 + 1];
          gr += pr * xr - pi * xi;
          gi += pr * xi + pi * xr;
        }
        /* x + i a g5 x */
        t[(s * 3 + c) * 2]     = b * (x[(s * 3 + c) * 2] - a * gi);
        t[(s * 3 + c) * 2 + 1] = b * (x[(s * 3 + c) * 2 + 1] + a * gr);
      }
    for (int k = 0; k < SS; k++) out[(long)i * SS + k] = t[k];
  }
}

/* ---- twisted-clover: A = C + i a gamma5 with the site-dependent clover matrix
 * C(x) = 1 + i coeff sum_{mu<nu} sigma_munu (x) F_munu(x), sigma_munu = (i/2)[gamma_mu, gamma_nu],
 * F_munu = (Q_munu - Q_munu^dag)/8, Q_munu = sum of the four plaquette leaves around x (Sheikholeslami-Wohlert term
 * D_W + csw (i/4) sigma_munu F_munu divided by 4 + m0; coeff = csw kappa = inv_param.clover_coeff of
 * qkxtm/MG_Bench.cpp:243-251).  The oracle keeps C as a dense 12x12 complex matrix per site, in whatever gamma basis is
 * set, [parity][cb][12][12][re,im]; no chiral-basis shortcut, inverse by Gaussian elimination. --------------------- */
static const double *CLOV = NULL;

void orc_set_clover(const double *clov) { CLOV = clov; }

void orc_clover_compute(double *clov, double *gauge[4], double coeff) {
  /* sigma_munu for the 6 planes */
  double sig[6][4][4][2];
  int plane = 0;
  for (int mu = 0; mu < 4; mu++)
    for (int nu = mu + 1; nu < 4; nu++, plane++) {
      double ab[4][4][2], ba[4][4][2];
      cmat_mul(ab, GAM[mu], GAM[nu]);
      cmat_mul(ba, GAM[nu], GAM[mu]);
      for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
          double re = ab[i][j][0] - ba[i][j][0], im = ab[i][j][1] - ba[i][j][1];
          sig[plane][i][j][0] = -0.5 * im;   /* (i/2)(re + i im) */
          sig[plane][i][j][1] = 0.5 * re;
        }
    }
#pragma omp parallel for
  for (int Y = 0; Y < V; Y++) {
    double *C = clov + ((long)orc_odd_bit(Y) * Vh + Y / 2) * 288;
    for (int k = 0; k < 288; k++) C[k] = 0.0;
    for (int k = 0; k < 12; k++) C[(k * 12 + k) * 2] = 1.0;
    int pl = 0;
    for (int mu = 0; mu < 4; mu++)
      for (int nu = mu + 1; nu < 4; nu++, pl++) {
        int xpm = shift_lex(Y, mu, 1), xpn = shift_lex(Y, nu, 1), xmm = shift_lex(Y, mu, -1), xmn = shift_lex(Y, nu, -1);
        int xmm_pn = shift_lex(xmm, nu, 1), xmm_mn = shift_lex(xmm, nu, -1), xpm_mn = shift_lex(xpm, nu, -1);
        double Q[18], a[18], b[18], t[18], one[18] = {1, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 1, 0};
        /* leaf 1: U_mu(x) U_nu(x+mu) U_mu(x+nu)^dag U_nu(x)^dag */
        su3_mm(a, link_at(gauge, mu, Y), link_at(gauge, nu, xpm));
        su3_mmd(a, a, link_at(gauge, mu, xpn));
        su3_mmd(Q, a, link_at(gauge, nu, Y));
        /* leaf 2: U_nu(x) U_mu(x-mu+nu)^dag U_nu(x-mu)^dag U_mu(x-mu) */
        su3_mmd(a, link_at(gauge, nu, Y), link_at(gauge, mu, xmm_pn));
        su3_mmd(a, a, link_at(gauge, nu, xmm));
        su3_mm(t, a, link_at(gauge, mu, xmm));
        for (int k = 0; k < 18; k++) Q[k] += t[k];
        /* leaf 3: U_mu(x-mu)^dag U_nu(x-mu-nu)^dag U_mu(x-mu-nu) U_nu(x-nu) */
        su3_mmd(a, one, link_at(gauge, mu, xmm));
        su3_mmd(a, a, link_at(gauge, nu, xmm_mn));
        su3_mm(a, a, link_at(gauge, mu, xmm_mn));
        su3_mm(t, a, link_at(gauge, nu, xmn));
        for (int k = 0; k < 18; k++) Q[k] += t[k];
        /* leaf 4: U_nu(x-nu)^dag U_mu(x-nu) U_nu(x+mu-nu) U_mu(x)^dag */
        su3_mmd(a, one, link_at(gauge, nu, xmn));
        su3_mm(a, a, link_at(gauge, mu, xmn));
        su3_mm(a, a, link_at(gauge, nu, xpm_mn));
        su3_mmd(t, a, link_at(gauge, mu, Y));
        for (int k = 0; k < 18; k++) Q[k] += t[k];
        /* F = (Q - Q^dag)/8 */
        for (int i = 0; i < 3; i++)
          for (int j = 0; j < 3; j++) {
            b[(i * 3 + j) * 2] = 0.125 * (Q[(i * 3 + j) * 2] - Q[(j * 3 + i) * 2]);
            b[(i * 3 + j) * 2 + 1] = 0.125 * (Q[(i * 3 + j) * 2 + 1] + Q[(j * 3 + i) * 2 + 1]);
          }
        /* C += i coeff sigma (x) F */
        for (int s = 0; s < 4; s++)
          for (int sp = 0; sp < 4; sp++) {
            double sr = sig[pl][s][sp][0], si = sig[pl][s][sp][1];
            if (sr == 0.0 && si == 0.0) continue;
            for (int i = 0; i < 3; i++)
              for (int j = 0; j < 3; j++) {
                double fr = b[(i * 3 + j) * 2], fi = b[(i * 3 + j) * 2 + 1];
                double pr = sr * fr - si * fi, pi = sr * fi + si * fr;
                double *e = C + ((s * 3 + i) * 12 + (sp * 3 + j)) * 2;
                e[0] -= coeff * pi;
                e[1] += coeff * pr;
              }
          }
      }
  }
}

/* out = (C + i a g5) in, its inverse, and their daggers on `nsites` sites of the parity block `parity`
 * (with CLOV == NULL: the plain twist).  C is hermitian, so the dagger only flips the sign of a. */
static void site_A(double *out, const double *in, int daggerBit, double kappa, double mu, int inverse, int parity, int nsites) {
  if (!CLOV) { orc_twist_gamma5(out, in, daggerBit, kappa, mu, inverse, nsites); return; }
  double a = 2.0 * kappa * mu;
  if (daggerBit) a = -a;
#pragma omp parallel for
  for (int i = 0; i < nsites; i++) {
    const double *C = CLOV + ((long)parity * Vh + i) * 288;
    double M[12][13][2];
    for (int r = 0; r < 12; r++) {
      for (int c = 0; c < 12; c++) {
        /* i a g5 : g5[s][s'] delta_cc' */
        int s = r / 3, sp = c / 3;
        double gr = (r % 3 == c % 3) ? G5[s][sp][0] : 0.0, gi = (r % 3 == c % 3) ? G5[s][sp][1] : 0.0;
        M[r][c][0] = C[(r * 12 + c) * 2] - a * gi;
        M[r][c][1] = C[(r * 12 + c) * 2 + 1] + a * gr;
      }
      M[r][12][0] = in[(long)i * SS + 2 * r]; M[r][12][1] = in[(long)i * SS + 2 * r + 1];
    }
    double y[12][2];
    if (!inverse) {
      for (int r = 0; r < 12; r++) {
        double re = 0, im = 0;
        for (int c = 0; c < 12; c++) {
          re += M[r][c][0] * M[c][12][0] - M[r][c][1] * M[c][12][1];
          im += M[r][c][0] * M[c][12][1] + M[r][c][1] * M[c][12][0];
        }
        y[r][0] = re; y[r][1] = im;
      }
    } else {
      /* Gaussian elimination with partial pivoting on [M | in] */
      for (int p = 0; p < 12; p++) {
        int piv = p; double best = M[p][p][0] * M[p][p][0] + M[p][p][1] * M[p][p][1];
        for (int r = p + 1; r < 12; r++) {
          double v2 = M[r][p][0] * M[r][p][0] + M[r][p][1] * M[r][p][1];
          if (v2 > best) { best = v2; piv = r; }
        }
        if (piv != p)
          for (int c = 0; c < 13; c++) {
            double tr = M[p][c][0], ti = M[p][c][1];
            M[p][c][0] = M[piv][c][0]; M[p][c][1] = M[piv][c][1]; M[piv][c][0] = tr; M[piv][c][1] = ti;
          }
        double ir = M[p][p][0] / best, ii = -M[p][p][1] / best;
        for (int c = p; c < 13; c++) {
          double xr = M[p][c][0], xi = M[p][c][1];
          M[p][c][0] = xr * ir - xi * ii; M[p][c][1] = xr * ii + xi * ir;
        }
        for (int r = 0; r < 12; r++) {
          if (r == p) continue;
          double fr = M[r][p][0], fi = M[r][p][1];
          for (int c = p; c < 13; c++) {
            M[r][c][0] -= fr * M[p][c][0] - fi * M[p][c][1];
            M[r][c][1] -= fr * M[p][c][1] + fi * M[p][c][0];
          }
        }
      }
      for (int r = 0; r < 12; r++) { y[r][0] = M[r][12][0]; y[r][1] = M[r][12][1]; }
    }
    for (int r = 0; r < 12; r++) { out[(long)i * SS + 2 * r] = y[r][0]; out[(long)i * SS + 2 * r + 1] = y[r][1]; }
  }
}

/* public form of the site operator (tests) */
void orc_site_A(double *out, const double *in, int daggerBit, double kappa, double mu, int inverse, int parity) {
  site_A(out, in, daggerBit, kappa, mu, inverse, parity, Vh);
}

/* [upstream-shape: tm_dslash] */
void orc_tm_dslash(double *res, double *gauge[4], const double *in, double kappa, double mu,
                   int oddBit, int daggerBit) {
  if (daggerBit) {
    double *tmp = (double *)malloc((size_t)Vh * SS * sizeof(double));
    site_A(tmp, in, daggerBit, kappa, mu, 1, 1 - oddBit, Vh);
    orc_dslash(res, gauge, tmp, oddBit, daggerBit);
    free(tmp);
  } else {
    orc_dslash(res, gauge, in, oddBit, daggerBit);
    site_A(res, res, daggerBit, kappa, mu, 1, oddBit, Vh);
  }
}

/* [upstream-shape: tm_matpc].  matpc: 0 even-even, 1 odd-odd, 2 even-even-asym, 3 odd-odd-asym
 * (qkxtm/Calc_Loops.cpp:443-450,712-713). */
void orc_tm_matpc(double *out, double *gauge[4], const double *in, double kappa, double mu,
                  int matpc, int daggerBit) {
  size_t n = (size_t)Vh * SS;
  double *tmp = (double *)malloc(n * sizeof(double));
  int p = matpc & 1;      /* parity the operator acts on */
  int asym = matpc >= 2;
  double k2 = -kappa * kappa;
  if (!asym) {
    orc_tm_dslash(tmp, gauge, in, kappa, mu, 1 - p, daggerBit);
    orc_tm_dslash(out, gauge, tmp, kappa, mu, p, daggerBit);
    orc_xpay(in, k2, out, (long)n);
  } else {
    /* M = A - k^2 D A^-1 D ;  M^dag = A^dag - k^2 D^dag A^-dag D^dag */
    orc_dslash(tmp, gauge, in, 1 - p, daggerBit);
    site_A(tmp, tmp, daggerBit, kappa, mu, 1, 1 - p, Vh);
    orc_dslash(out, gauge, tmp, p, daggerBit);
    site_A(tmp, in, daggerBit, kappa, mu, 0, p, Vh);
    orc_xpay(tmp, k2, out, (long)n);
  }
  free(tmp);
}

void orc_tm_mdagm(double *out, double *gauge[4], const double *in, double kappa, double mu, int matpc) {
  size_t n = (size_t)Vh * SS;
  double *tmp = (double *)malloc(n * sizeof(double));
  orc_tm_matpc(tmp, gauge, in, kappa, mu, matpc, 0);
  orc_tm_matpc(out, gauge, tmp, kappa, mu, matpc, 1);
  free(tmp);
}

/* [upstream-shape: tm_mat].  Full field = [even Vh | odd Vh].  out = A in - kappa D in. */
void orc_tm_mat(double *out, double *gauge[4], const double *in, double kappa, double mu, int daggerBit) {
  size_t n = (size_t)Vh * SS;
  const double *inE = in, *inO = in + n;
  double *outE = out, *outO = out + n;
  double *tmp = (double *)malloc(2 * n * sizeof(double));
  orc_dslash(outO, gauge, inE, 1, daggerBit);
  orc_dslash(outE, gauge, inO, 0, daggerBit);
  site_A(tmp, inE, daggerBit, kappa, mu, 0, 0, Vh);
  site_A(tmp + n, inO, daggerBit, kappa, mu, 0, 1, Vh);
  orc_xpay(tmp, -kappa, out, 2 * (long)n);
  free(tmp);
}

/* Even-odd preparation for M_full = A - kappa D, MAT solution:
 *   sym  : src_p = A^-1 (b_p + kappa D_{p,q} A^-1 b_q)
 *   asym : src_p =       b_p + kappa D_{p,q} A^-1 b_q      with q = 1 - p */
void orc_prepare(double *src, double *gauge[4], const double *b, double kappa, double mu, int matpc) {
  size_t n = (size_t)Vh * SS;
  int p = matpc & 1, asym = matpc >= 2;
  const double *bp = b + (size_t)p * n, *bq = b + (size_t)(1 - p) * n;
  double *tmp = (double *)malloc(n * sizeof(double));
  site_A(tmp, bq, 0, kappa, mu, 1, 1 - p, Vh);
  orc_dslash(src, gauge, tmp, p, 0);
  orc_xpay(bp, kappa, src, (long)n);
  if (!asym) site_A(src, src, 0, kappa, mu, 1, p, Vh);
  free(tmp);
}

/* x_q = A^-1 (b_q + kappa D_{q,p} x_p); x_p is already in place in x. */
void orc_reconstruct(double *x, double *gauge[4], const double *b, double kappa, double mu, int matpc) {
  size_t n = (size_t)Vh * SS;
  int p = matpc & 1;
  const double *bq = b + (size_t)(1 - p) * n;
  double *xp = x + (size_t)p * n, *xq = x + (size_t)(1 - p) * n;
  orc_dslash(xq, gauge, xp, 1 - p, 0);
  orc_xpay(bq, kappa, xq, (long)n);
  site_A(xq, xq, 0, kappa, mu, 1, 1 - p, Vh);
}

/* ---- blas [upstream-shape: tests/blas_reference.cpp] ------------------------------------------ */
void orc_ax(double a, double *x, long n) {
#pragma omp parallel for
  for (long i = 0; i < n; i++) x[i] *= a;
}
void orc_axpy(double a, const double *x, double *y, long n) {
#pragma omp parallel for
  for (long i = 0; i < n; i++) y[i] += a * x[i];
}
void orc_xpay(const double *x, double a, double *y, long n) {
#pragma omp parallel for
  for (long i = 0; i < n; i++) y[i] = x[i] + a * y[i];
}
void orc_mxpy(const double *x, double *y, long n) {
#pragma omp parallel for
  for (long i = 0; i < n; i++) y[i] -= x[i];
}
double orc_norm2(const double *x, long n) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s)
  for (long i = 0; i < n; i++) s += x[i] * x[i];
  return s;
}
double orc_redot(const double *x, const double *y, long n) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s)
  for (long i = 0; i < n; i++) s += x[i] * y[i];
  return s;
}
void orc_cdot(const double *x, const double *y, long n, double out[2]) {
  double re = 0.0, im = 0.0;
#pragma omp parallel for reduction(+ : re, im)
  for (long i = 0; i < n / 2; i++) {
    re += x[2 * i] * y[2 * i] + x[2 * i + 1] * y[2 * i + 1];
    im += x[2 * i] * y[2 * i + 1] - x[2 * i + 1] * y[2 * i];
  }
  out[0] = re; out[1] = im;
}

/* ---- CG on M^dag M.  The reference has no host CG: this is the plain CG
 * one writes over the host matpc + host blas, the comparison invert_test's residual check
 * implies.  Call sequence: lib/qudaQKXTM_interface.cpp:2031-2037. ------------------------------ */
int orc_cg_mdagm(double *x, double *gauge[4], const double *b, double kappa, double mu, int matpc,
                 double tol, int maxiter, int pr_beta, double *true_res, double *r2_hist) {
  long n = (long)Vh * SS;
  double *r = (double *)malloc(n * sizeof(double));
  double *p = (double *)malloc(n * sizeof(double));
  double *Ap = (double *)malloc(n * sizeof(double));
  double *rold = pr_beta ? (double *)malloc(n * sizeof(double)) : NULL;
  double b2 = orc_norm2(b, n);
  /* x0 = 0 */
  memset(x, 0, n * sizeof(double));
  memcpy(r, b, n * sizeof(double));
  memcpy(p, b, n * sizeof(double));
  double r2 = b2, stop = tol * tol * b2;
  int k = 0;
  if (r2_hist) r2_hist[0] = r2;
  while (r2 > stop && k < maxiter) {
    orc_tm_mdagm(Ap, gauge, p, kappa, mu, matpc);
    double pAp = orc_redot(p, Ap, n);
    double alpha = r2 / pAp;
    if (pr_beta) memcpy(rold, r, n * sizeof(double));
    orc_axpy(-alpha, Ap, r, n);
    double r2n = orc_norm2(r, n);
    double beta;
    if (pr_beta) {
      double sigma = r2n - orc_redot(r, rold, n);
      beta = (sigma > 0 ? sigma : r2n) / r2;
    } else beta = r2n / r2;
    orc_axpy(alpha, p, x, n);
    orc_xpay(r, beta, p, n);
    r2 = r2n;
    k++;
    if (r2_hist) r2_hist[k] = r2;
  }
  if (true_res) {
    orc_tm_mdagm(Ap, gauge, x, kappa, mu, matpc);
    memcpy(r, b, n * sizeof(double));
    orc_mxpy(Ap, r, n);
    *true_res = sqrt(orc_norm2(r, n) / b2);
  }
  free(r); free(p); free(Ap);
  if (rold) free(rold);
  return k;
}
