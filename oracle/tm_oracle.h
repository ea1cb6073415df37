/*
 * tm_oracle.h -- CPU oracle for the even-odd twisted-mass Wilson Dslash + CG on M^dag M.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product library (libtmq.so)
 * never links, loads or calls anything in oracle/.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in upstream QUDA (lattice/quda, branch
 * feature/multigrid, ~v0.8.0 -> 0.9.0-dev, mid 2017; no version pin -- reference CMakeLists.txt:18,
 * README:4,50,138), which is not vendored in /root/reference and cannot be built here.  The
 * reference ships no test source, golden vector or known-answer value for this path
 * (reference tests/CMakeLists.txt only compiles upstream tests).  This file therefore restates
 * the published algorithm (upstream tests/wilson_dslash_reference.cpp + tests/dslash_util.h +
 * tests/blas_reference.cpp, by structure) on top of the conventions that ARE pinned in-tree:
 *
 *   checkerboard / neighbour indexing ....... qkxtm/QKXTM_util.cpp:405-470
 *   QDP even-odd gauge order, row-major 3x3 .. qkxtm/QKXTM_util.cpp:840-857, lib/qudaQKXTM_Gauge.cpp:73-89
 *   anti-periodic T folded into U_t(T-1) ..... qkxtm/QKXTM_util.cpp:682-705
 *   recon-12 third row + boundary sign ....... qkxtm/QKXTM_util.cpp:281-295
 *   UKQCD gamma matrices, 1 -+ gamma_mu ...... lib/code_pieces/gammas_tm_base.h:21-32,148-171
 *   gamma5 = spin swap 0<->2, 1<->3 .......... lib/code_pieces/apply_gamma5_vector_core.h:1-16
 *   hop sign convention ...................... lib/code_pieces/fixSinkContractions_noether_core.h:117-137
 *   host spinor order [x][s][c][ri] .......... lib/qudaQKXTM_Vector.cpp:72-81
 *   CG call sequence on M^dag M .............. lib/qudaQKXTM_interface.cpp:2020-2041
 *
 * and is pinned by algebraic identities instead of golden vectors (tests/test_oracle_*.py):
 * Clifford algebra, <x,My> = <M^dag x,y>, gamma5-hermiticity, free-field plane-wave eigenvalue,
 * Schur-complement identity, gauge covariance, an independent dense numpy restatement on the
 * full (non-checkerboarded) lattice, and a DeGrand-Rossi <-> UKQCD change of basis.
 *
 * Field shapes (as in the upstream host reference):
 *   parity spinor : double[Vh][4 spin][3 colour][2]     (cb index = lexicographic index / 2)
 *   gauge         : 4 arrays gauge[mu], each double[even Vh | odd Vh][3][3][2], row-major
 */
#ifndef TM_ORACLE_H
#define TM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* geometry ------------------------------------------------------------------------------------ */
void orc_set_lattice(const int X[4]);                 /* sets the globals Z[], V, Vh              */
int  orc_volume(void);
int  orc_full_lattice_index(int i, int oddBit);       /* QKXTM_util.cpp:418-442                   */
int  orc_neighbor_index(int i, int oddBit, int dx4, int dx3, int dx2, int dx1); /* :455-470       */
int  orc_odd_bit(int Y);                              /* :191-197                                 */

/* gamma basis: g[mu][row][col][re,im], mu = 0..3 = x,y,z,t.  Builds the projector table
 * P[2mu] = 1 - gamma_mu, P[2mu+1] = 1 + gamma_mu and gamma5 = g0 g1 g2 g3.                      */
void orc_set_gamma(const double g[4][4][4][2]);
void orc_get_gamma5(double g5[4][4][2]);

/* gauge helpers ------------------------------------------------------------------------------- */
void orc_apply_t_boundary(double *gauge[4], int sign);                  /* :698-705 (sign = -1)   */
void orc_su3_reconstruct12(double *mat18, double u0);                   /* :281-295               */
double orc_plaquette(double *gauge[4]);     /* sum Re tr P /(V*3*6), lib/qudaQKXTM_kernels.cu:957  */

/* operator ------------------------------------------------------------------------------------ */
/* res = D_{p,p'} in : spin-projected 8-direction hop, no 1/2, no kappa.  oddBit = parity of res. */
void orc_dslash(double *res, double *gauge[4], const double *in, int oddBit, int daggerBit);
/* out = b (in + i a gamma5 in), a = +-2 kappa mu, b = 1 (direct) or 1/(1+a^2) (inverse).
 * inverse flips the sign of a, dagger flips it again.                                            */
void orc_twist_gamma5(double *out, const double *in, int daggerBit, double kappa, double mu,
                      int inverse, int nsites);
/* out = A^-1 D in (or, with dagger, D^dag A^-dag in -- twist BEFORE the hop)                      */
void orc_tm_dslash(double *res, double *gauge[4], const double *in, double kappa, double mu,
                   int oddBit, int daggerBit);
/* symmetric:   out = in - kappa^2 A^-1 D A^-1 D in          (matpc 0 = even-even, 1 = odd-odd)
 * asymmetric:  out = A in - kappa^2 D A^-1 D in             (matpc 2 = ee-asym,   3 = oo-asym)   */
void orc_tm_matpc(double *out, double *gauge[4], const double *in, double kappa, double mu,
                  int matpc, int daggerBit);
void orc_tm_mdagm(double *out, double *gauge[4], const double *in, double kappa, double mu, int matpc);
/* full operator on [even Vh | odd Vh]: out = A in - kappa D in (kappa normalisation)              */
void orc_tm_mat(double *out, double *gauge[4], const double *in, double kappa, double mu, int daggerBit);
/* even-odd source preparation / solution reconstruction for the symmetric pc operator, MAT
 * solution type (interface.cpp:2020,2040).  b, x are full fields [even|odd].                      */
void orc_prepare(double *src, double *gauge[4], const double *b, double kappa, double mu, int matpc);
void orc_reconstruct(double *x, double *gauge[4], const double *b, double kappa, double mu, int matpc);

/* twisted-clover: dense 12x12 site matrices C(x), [parity][cb][12][12][re,im] = 288 doubles per site, built from the
 * gauge field with coeff = csw * kappa (see tm_oracle.c).  orc_set_clover(clov) switches EVERY operator above from the
 * constant twist A = 1 + i a g5 to A = C + i a g5 (NULL switches back); orc_site_A applies A, A^-1 and their daggers. */
void orc_clover_compute(double *clov, double *gauge[4], double coeff);
void orc_set_clover(const double *clov);
void orc_site_A(double *out, const double *in, int daggerBit, double kappa, double mu, int inverse, int parity);

/* blas (upstream tests/blas_reference.cpp shapes) ---------------------------------------------- */
void   orc_ax(double a, double *x, long n);
void   orc_axpy(double a, const double *x, double *y, long n);       /* y += a x                  */
void   orc_xpay(const double *x, double a, double *y, long n);       /* y = x + a y               */
void   orc_mxpy(const double *x, double *y, long n);                 /* y -= x                    */
double orc_norm2(const double *x, long n);
double orc_redot(const double *x, const double *y, long n);
void   orc_cdot(const double *x, const double *y, long n, double out[2]); /* sum conj(x) y        */

/* plain CG on M_pc^dag M_pc.  Returns iterations.  Stops when r2 < tol^2 |b|^2.
 * pr_beta = 0: beta = r2_new/r2_old; 1: beta = <r_new, r_new - r_old>/r2_old (upstream's choice)  */
int orc_cg_mdagm(double *x, double *gauge[4], const double *b, double kappa, double mu, int matpc,
                 double tol, int maxiter, int pr_beta, double *true_res, double *r2_hist);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
