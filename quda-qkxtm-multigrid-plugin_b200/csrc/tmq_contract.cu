// tmq_contract.cu -- what the plug-in does with the solved columns right after the hot path: the propagator container
// kernels and the meson two-point contraction of calcMG_threepTwop_EvenOdd (reference lib/qudaQKXTM_interface.cpp:1190-1223):
//   copyPropagator / absorbVectorToDevice (+ the time-slice variants)  lib/qudaQKXTM_Vector.cpp:464-512, Propagator.cpp:90-106,511-550
//   conjugate (vector, propagator)                                      lib/code_pieces/conjugate_{vector,propagator}_core.h
//   apply_gamma5 (propagator)                                           lib/code_pieces/apply_gamma5_propagator_core.h
//   rotateToPhysicalBase_device(sign)                                   lib/code_pieces/rotateToPhysicalBase_core.h
//   contractMesons (position and momentum space)                        lib/code_pieces/contractMesons_core{,_PosSpace}.h,
//                                                                       lib/qudaQKXTM_kernels.cu:1127-1225, Contraction.cpp:1606-1648
// All on the QKXTM device layouts: vector d[(s*3+c)*V + x], propagator d[((mu*4+nu)*9 + c1*3+c2)*V + x] (complex, x lexicographic).
//
// Meson contraction.  The reference's tables c_mesons_indices / c_mesons_values (lib/qudaQKXTM_kernels.cu:77-78) are the
// non-zero entries of      C_G(x) = s_G  tr[ G S(x) G^dag g5 S(x)^dag g5 ],     s_G = +1 for G in {g5, 1, g5 g_mu}, -1 for g_mu
// (channel order pseudoscalar, scalar, g5g1..g5g4, g1..g4, lib/qudaQKXTM_interface.cpp:305-314), evaluated for S = prop1 and for
// S = prop2.  Here the tensors are DERIVED from the UKQCD gamma matrices at first use (nothing is tabulated): every G is a
// monomial matrix whose permutation is an XOR of the spin index, a -> a ^ m_G, and g5 is a -> a ^ 2, so with phases
// G[a][a^m] = phi_a
//     C_G = s_G sum_{colour pairs} sum_{a,g} phi_a conj(phi_g) S[a^m][g^m] conj(S[a^2][g^2]) ,      phi_a conj(phi_g) = +-1.
// The 16 products S[a^m][g^m] conj(S[a^2][g^2]) are shared by the channels with the same m (4 groups).
//
// The reference then projects every time slice on the momentum list inside the same kernel with one 64-thread shared-memory
// tree per momentum (cost  V * Nmoms) and finishes the sum over blocks on the host, one launch + one blocking D2H per time
// slice.  Here: one site kernel over the whole local lattice (HBM-bound: 2 x 144 complex read, 20 complex written per site),
// then a SEPARABLE Fourier sum -- over x for the distinct p_x, over y for the distinct (p_x, p_y), over z for the momenta --
// so the projection costs about one more read of the site values instead of Nmoms reads.  All sums are in fp64 and in a fixed
// order (warp shuffle trees), the result is identical from run to run.  On a sharded lattice every rank fills its own time
// slices of the GLOBAL-T result and one all-reduce both sums the z ranks and gathers the t ranks (the reference's MPI_Reduce
// over GK_spaceComm + MPI_Gather over GK_timeComm, Contraction.cpp:1638, 98-111).
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>
#include <algorithm>
#include <complex>
#include <map>
#include <vector>
#include "../../include/tmq.h"
#include "tmq_internal.h"

namespace tmq {

constexpr int CT_BLOCK = 128;

// ---- site-local container kernels --------------------------------------------------------------------------------------
template <typename F> __global__ void __launch_bounds__(256) qk_conj_kernel(CplxT<F> *__restrict__ d, size_t n) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) d[i].im = -d[i].im;
}

// gamma5 on the sink spin index of a propagator: rows mu <-> mu + 2 (UKQCD gamma5 is the spin swap)
template <typename F> __global__ void __launch_bounds__(256) qk_gamma5_prop_kernel(CplxT<F> *__restrict__ d, size_t V) {
  const size_t e = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= V * 72) return;
  const size_t pq = e / V, x = e - pq * V;           // pq = mu * 36 + (nu, c1, c2), mu in {0, 1}
  const CplxT<F> lo = d[pq * V + x], hi = d[(pq + 72) * V + x];
  d[pq * V + x] = hi;
  d[(pq + 72) * V + x] = lo;
}

// twisted -> physical basis: P <- 1/2 (1 + i s g5) P (1 + i s g5), per colour pair
template <typename F> __global__ void __launch_bounds__(CT_BLOCK) qk_rotate_kernel(CplxT<F> *__restrict__ d, size_t V, F s) {
  const size_t e = (size_t)blockIdx.x * CT_BLOCK + threadIdx.x;
  if (e >= V * 9) return;
  const size_t cc = e / V, x = e - cc * V;
  F P[16][2];
#pragma unroll
  for (int k = 0; k < 16; k++) { const CplxT<F> v = d[((size_t)k * 9 + cc) * V + x]; P[k][0] = v.re; P[k][1] = v.im; }
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int g = 0; g < 4; g++) {
      const int k = a * 4 + g, ka = (a ^ 2) * 4 + g, kg = a * 4 + (g ^ 2), kag = (a ^ 2) * 4 + (g ^ 2);
      CplxT<F> o;      // i s z = (-s Im z, s Re z)
      o.re = (F)0.5 * (P[k][0] - s * P[ka][1] - s * P[kg][1] - P[kag][0]);
      o.im = (F)0.5 * (P[k][1] + s * P[ka][0] + s * P[kg][0] - P[kag][1]);
      d[((size_t)k * 9 + cc) * V + x] = o;
    }
}

// ---- meson contraction: site kernel ------------------------------------------------------------------------------------
template <typename F> struct MesonSigns { F w[10][16]; };   // s_G phi_a conj(phi_g), index a*4+g; kernel parameter (constant bank)
// XOR masks of the ten channels in the UKQCD basis (checked against the gamma matrices by meson_signs())
#define TMQ_MESON_XORS {2, 0, 1, 1, 0, 2, 3, 3, 2, 0}

// the channels i0, i1 (, i2) share the XOR mask M: acc[i] += w_i[a][g] * S[a^M][g^M] conj(S[a^2][g^2]) over the 16 (a, g)
template <int M, typename F>
__device__ __forceinline__ void meson_group(const F (&S)[16][2], const MesonSigns<F> &W, F (&acc)[10][2], int i0, int i1, int i2) {
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int g = 0; g < 4; g++) {
      const int k1 = (a ^ M) * 4 + (g ^ M), k2 = (a ^ 2) * 4 + (g ^ 2), q = a * 4 + g;
      const F tr = S[k1][0] * S[k2][0] + S[k1][1] * S[k2][1];      // S1 conj(S2)
      const F ti = S[k1][1] * S[k2][0] - S[k1][0] * S[k2][1];
      acc[i0][0] += W.w[i0][q] * tr; acc[i0][1] += W.w[i0][q] * ti;
      acc[i1][0] += W.w[i1][q] * tr; acc[i1][1] += W.w[i1][q] * ti;
      if (i2 >= 0) { acc[i2][0] += W.w[i2][q] * tr; acc[i2][1] += W.w[i2][q] * ti; }
    }
}

// fp32 propagators: T[M][a*4+g] += S[a^M][g^M] conj(S[a^2][g^2]) for the four XOR masks (64 complex accumulators, 4 FFMA each);
// the channel sums are taken once, after the nine colour pairs, from the summed products
template <int M>
__device__ __forceinline__ void meson_products(const float (&S)[16][2], float (&T)[16][2]) {
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int g = 0; g < 4; g++) {
      const int k1 = (a ^ M) * 4 + (g ^ M), k2 = (a ^ 2) * 4 + (g ^ 2), q = a * 4 + g;
      T[q][0] += S[k1][0] * S[k2][0]; T[q][0] += S[k1][1] * S[k2][1];
      T[q][1] += S[k1][1] * S[k2][0]; T[q][1] -= S[k1][0] * S[k2][1];
    }
}
__device__ __forceinline__ void meson_channels(const float (&T)[16][2], const MesonSigns<float> &W, float (&c)[10][2], int i0, int i1, int i2) {
#pragma unroll
  for (int q = 0; q < 16; q++) {
    c[i0][0] += W.w[i0][q] * T[q][0]; c[i0][1] += W.w[i0][q] * T[q][1];
    c[i1][0] += W.w[i1][q] * T[q][0]; c[i1][1] += W.w[i1][q] * T[q][1];
    if (i2 >= 0) { c[i2][0] += W.w[i2][q] * T[q][0]; c[i2][1] += W.w[i2][q] * T[q][1]; }
  }
}

// site values csite[(iu*10 + ip) * V + x] = C_ip(x) of propagator iu = blockIdx.y (complex double).  fp64 propagators: everything
// in fp64.  fp32 propagators: the per-site sum in fp32 (as the reference, which does the whole lattice sum in float); everything
// after the site kernel is fp64.
template <typename F>
__global__ void __launch_bounds__(CT_BLOCK, (sizeof(F) == 4 ? 3 : 2)) meson_site_kernel(CplxT<double> *__restrict__ csite, const CplxT<F> *__restrict__ prop1,
                                                             const CplxT<F> *__restrict__ prop2, size_t V, MesonSigns<F> W) {
  const size_t x = (size_t)blockIdx.x * CT_BLOCK + threadIdx.x;
  if (x >= V) return;
  const int iu = blockIdx.y;
  const CplxT<F> *__restrict__ prop = iu ? prop2 : prop1;
  F c[10][2];
#pragma unroll
  for (int ip = 0; ip < 10; ip++) { c[ip][0] = 0; c[ip][1] = 0; }
  if constexpr (sizeof(F) == 8) {
#pragma unroll 1
    for (int cc = 0; cc < 9; cc++) {
      F S[16][2];
#pragma unroll
      for (int k = 0; k < 16; k++) { const CplxT<F> v = prop[((size_t)k * 9 + cc) * V + x]; S[k][0] = v.re; S[k][1] = v.im; }
      meson_group<0>(S, W, c, 1, 4, 9);
      meson_group<1>(S, W, c, 2, 3, -1);
      meson_group<2>(S, W, c, 0, 5, 8);
      meson_group<3>(S, W, c, 6, 7, -1);
    }
  } else {
    float T0[16][2], T1[16][2], T2[16][2], T3[16][2];
#pragma unroll
    for (int q = 0; q < 16; q++) { T0[q][0] = T0[q][1] = T1[q][0] = T1[q][1] = T2[q][0] = T2[q][1] = T3[q][0] = T3[q][1] = 0.f; }
#pragma unroll 1
    for (int cc = 0; cc < 9; cc++) {
      float S[16][2];
#pragma unroll
      for (int k = 0; k < 16; k++) { const CplxT<F> v = prop[((size_t)k * 9 + cc) * V + x]; S[k][0] = v.re; S[k][1] = v.im; }
      meson_products<0>(S, T0); meson_products<1>(S, T1); meson_products<2>(S, T2); meson_products<3>(S, T3);
    }
    meson_channels(T0, W, c, 1, 4, 9);
    meson_channels(T1, W, c, 2, 3, -1);
    meson_channels(T2, W, c, 0, 5, 8);
    meson_channels(T3, W, c, 6, 7, -1);
  }
#pragma unroll
  for (int ip = 0; ip < 10; ip++) {
    CplxT<double> o; o.re = (double)c[ip][0]; o.im = (double)c[ip][1];
    csite[((size_t)iu * 10 + ip) * V + x] = o;
  }
}


// ---- baryon two-point contraction -------------------------------------------------------------------------------------------------
// The reference's ten channels (lib/code_pieces/contractBaryons_core.h; nucl_nucl, nucl_roper, roper_nucl, roper_roper,
// deltapp_deltamm_11/22/33, deltap_deltaz_11/22/33, lib/qudaQKXTM_interface.cpp:294-303) all have the form
//   C[g][g'] = sum Gs[a,b] conj(Gr)[a',b'] Xs[g,d] Xr[g',d'] eps_{ijk} eps_{i'j'k'} sum_terms coef P1[a,s1]^{i c1} P2[b,s2]^{j c2} P3[d,s3]^{k c3}
// where (s1,s2,s3) / (c1,c2,c3) is a permutation of the source spins (a',b',d') / colours (i',j',k'): nucleon J = eps (u^T C g5 d) u,
// "roper" J = eps (u^T C d) g5 u, Delta J_k = eps (u^T C g_k u) u; its index / value tables (lib/qudaQKXTM_kernels.cu:79-88) are the
// non-zero entries of Gs x conj(Gr) x Xs x Xr (checked numerically, oracle/oracle.py: baryon_channels).  All four matrices are
// monomial (one entry per row), so they are kept as a permutation and a phase per row, derived from the gamma matrices on the host.
// The reference evaluates every term as a 16 x 36 x 16-fold sum of triple products per (g, g'); here a term whose third line ends on
// the open source index is a scalar (16 products) times a spin matrix, and the others are two 4x4 matrix products per colour
// combination:  M[d][a] = sum_A conj(Gr)_A P3[d][y(A)] Q[row(a)][x(A)],  R[d][D] += sum_a M[d][a] Gs_a W[row'(a)][D].
struct Mono { int perm[4]; float re[4], im[4]; };          // row r: column perm[r], value (re, im)
struct BaryonTerm { double coef; int prop[3]; int slot[3]; };   // prop: 0 = the channel's own propagator, 1 = the other; slot 0,1,2 = (a', b', d')
struct BaryonChannel { Mono gs, grc; int xs_inv[4], xr_inv[4]; float xs_re[4], xs_im[4], xr_re[4], xr_im[4]; int term0, nterm; };
struct BaryonTables { BaryonChannel ch[10]; BaryonTerm term[16]; };
__constant__ BaryonTables c_baryon;

constexpr int BY_SITES = 32;      // sites per CTA = lanes of a warp; 4 warps share the 20 (channel, propagator) tasks
constexpr int BY_THREADS = 128;

template <typename F> struct Cx { F re, im; };
template <typename F> __device__ __forceinline__ Cx<F> cmul(Cx<F> a, Cx<F> b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
template <typename F> __device__ __forceinline__ void cmac(Cx<F> &acc, Cx<F> a, Cx<F> b) {
  acc.re += a.re * b.re; acc.re -= a.im * b.im; acc.im += a.re * b.im; acc.im += a.im * b.re;
}

// site values csite[((iu*10 + ip)*16 + g*4 + g') * nsites + x] for the nsites sites starting at site0 (complex double)
template <typename F>
__global__ void __launch_bounds__(BY_THREADS) baryon_site_kernel(CplxT<double> *__restrict__ csite, const CplxT<F> *__restrict__ prop1,
                                                                const CplxT<F> *__restrict__ prop2, size_t V, size_t site0, size_t nsites) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CplxT<F> *P = (CplxT<F> *)smem_raw;                       // [2 prop][144 comp][BY_SITES]
  const size_t base = (size_t)blockIdx.x * BY_SITES;
  for (int idx = threadIdx.x; idx < 2 * 144 * BY_SITES; idx += BY_THREADS) {
    const int s = idx % BY_SITES, k = (idx / BY_SITES) % 144, pr = idx / (BY_SITES * 144);
    CplxT<F> v; v.re = 0; v.im = 0;
    if (base + s < nsites) v = (pr ? prop2 : prop1)[(size_t)k * V + site0 + base + s];
    P[idx] = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool live = base + lane < nsites;
  // P element (spin row mu, spin column nu, colour row c1, colour column c2) of propagator pr at this lane's site
  auto ld = [&](int pr, int mu, int nu, int c1, int c2) -> Cx<F> {
    const CplxT<F> v = P[((pr * 144) + (mu * 4 + nu) * 9 + c1 * 3 + c2) * BY_SITES + lane];
    return {v.re, v.im};
  };
  for (int task = warp; task < 20; task += 4) {
    const int ip = task >> 1, iu = task & 1;
    const BaryonChannel &ch = c_baryon.ch[ip];
    Cx<F> R[4][4];
#pragma unroll
    for (int d = 0; d < 4; d++)
#pragma unroll
      for (int D = 0; D < 4; D++) R[d][D] = {0, 0};
    for (int it = 0; it < ch.nterm; it++) {
      const BaryonTerm &tm = c_baryon.term[ch.term0 + it];
      const int p1 = tm.prop[0] ^ iu, p2 = tm.prop[1] ^ iu, p3 = tm.prop[2] ^ iu;
      const int s1 = tm.slot[0], s2 = tm.slot[1], s3 = tm.slot[2];
#pragma unroll 1
      for (int cc = 0; cc < 36; cc++) {
        const int e1 = cc / 6, e2 = cc - e1 * 6;
        // the six permutations of (0,1,2): even ones first
        int ci[3], cs[3];
        ci[0] = e1 < 3 ? e1 : (e1 - 3); ci[1] = e1 < 3 ? (e1 + 1) % 3 : (e1 - 3 + 2) % 3; ci[2] = 3 - ci[0] - ci[1];
        cs[0] = e2 < 3 ? e2 : (e2 - 3); cs[1] = e2 < 3 ? (e2 + 1) % 3 : (e2 - 3 + 2) % 3; cs[2] = 3 - cs[0] - cs[1];
        const F w = (F)(((e1 < 3) == (e2 < 3)) ? tm.coef : -tm.coef);
        const int k1 = s1 == 0 ? cs[0] : (s1 == 1 ? cs[1] : cs[2]), k2 = s2 == 0 ? cs[0] : (s2 == 1 ? cs[1] : cs[2]),
                  k3 = s3 == 0 ? cs[0] : (s3 == 1 ? cs[1] : cs[2]);
        if (s3 == 2) {
          // third line ends on the open index: scalar * P3
          Cx<F> sc = {0, 0};
#pragma unroll
          for (int a = 0; a < 4; a++)
#pragma unroll
            for (int A = 0; A < 4; A++) {
              const int b = ch.gs.perm[a], B = ch.grc.perm[A];
              const Cx<F> ph = cmul<F>({(F)ch.gs.re[a], (F)ch.gs.im[a]}, {(F)ch.grc.re[A], (F)ch.grc.im[A]});
              const Cx<F> u = ld(p1, a, s1 == 0 ? A : B, ci[0], k1), v = ld(p2, b, s2 == 0 ? A : B, ci[1], k2);
              cmac(sc, cmul(ph, u), v);
            }
          sc.re *= w; sc.im *= w;
#pragma unroll
          for (int d = 0; d < 4; d++)
#pragma unroll
            for (int D = 0; D < 4; D++) cmac(R[d][D], sc, ld(p3, d, D, ci[2], k3));
        } else {
          // line `wl` (1 or 2) ends on the open index; the other one (`ol`) and line 3 are tied over the source spin A
          const bool w1 = s1 == 2;                                   // true: line 1 is the open one
          const int po = w1 ? p2 : p1, pw = w1 ? p1 : p2, so = w1 ? s2 : s1, ko = w1 ? k2 : k1, kw = w1 ? k1 : k2;
          const int cio = w1 ? ci[1] : ci[0], ciw = w1 ? ci[0] : ci[1];
          Cx<F> Q[4][4], Wm[4][4];
#pragma unroll
          for (int a = 0; a < 4; a++) {
            const int ro = w1 ? ch.gs.perm[a] : a, rw = w1 ? a : ch.gs.perm[a];
            const Cx<F> phs = {(F)(w * ch.gs.re[a]), (F)(w * ch.gs.im[a])};
#pragma unroll
            for (int A = 0; A < 4; A++) {
              const int B = ch.grc.perm[A];
              Q[a][A] = cmul<F>({(F)ch.grc.re[A], (F)ch.grc.im[A]}, ld(po, ro, so == 0 ? A : B, cio, ko));
              Wm[a][A] = cmul(phs, ld(pw, rw, A, ciw, kw));          // second index of Wm is the open source spin D
            }
          }
#pragma unroll
          for (int d = 0; d < 4; d++) {
            Cx<F> P3r[4], Md[4];
#pragma unroll
            for (int A = 0; A < 4; A++) P3r[A] = ld(p3, d, s3 == 0 ? A : ch.grc.perm[A], ci[2], k3);
#pragma unroll
            for (int a = 0; a < 4; a++) {
              Md[a] = {0, 0};
#pragma unroll
              for (int A = 0; A < 4; A++) cmac(Md[a], P3r[A], Q[a][A]);
            }
#pragma unroll
            for (int D = 0; D < 4; D++)
#pragma unroll
              for (int a = 0; a < 4; a++) cmac(R[d][D], Md[a], Wm[a][D]);
          }
        }
      }
    }
    // out[g][g'] = sum Xs[g][d] Xr[g'][D] R[d][D]: the monomial X matrices only move and rephase the entries
    if (live) {
#pragma unroll
      for (int d = 0; d < 4; d++)
#pragma unroll
        for (int D = 0; D < 4; D++) {
          const Cx<F> ph = cmul<F>({(F)ch.xs_re[d], (F)ch.xs_im[d]}, {(F)ch.xr_re[D], (F)ch.xr_im[D]});
          const Cx<F> o = cmul(ph, R[d][D]);
          CplxT<double> v; v.re = (double)o.re; v.im = (double)o.im;
          csite[((size_t)((iu * 10 + ip) * 16 + ch.xs_inv[d] * 4 + ch.xr_inv[D])) * nsites + base + lane] = v;
        }
    }
  }
}

// ---- fixed-sink three-point function: sequential sources and the ultra-local insertion ----------------------------------------------
// (lib/code_pieces/seqSourceFixSinkPart{1,2}_core.h + projectors_tm_base.h, fixSinkContractions_local_core.h + gammas_tm_base.h;
//  restated as formulas in oracle/oracle.py: seq_source_part1/2, projector_tm, operator_tm, fixsink_local_site, each checked against
//  the reference's kernel body to rounding.)  With G = C g5, P the twisted-basis projector, (nu_f, c2_f) the fixed source spin / colour:
//   part 2:  S[n][w] = - eps_{uvw} eps_{UV c2_f} G[m,n] G[nu_f,k] P[b,a] ( T[m,b]^{uU} T[a,k]^{vV} + T[m,k]^{uU} T[a,b]^{vV} )
//   part 1:  S[n][w] = - eps eps' G[m,g] G[j,k] P[b,a] T2[g,j]^{uU} ( d_{mn} d_{b nu_f} T1[a,k] + d_{mn} d_{k nu_f} T1[a,b]
//                                                                    + d_{an} d_{b nu_f} T1[m,k] + d_{an} d_{k nu_f} T1[m,b] )^{vV}
// One thread per spatial site of ONE time slice, called 12 times per sink: negligible next to the 12 solves that follow.
struct SeqSrcTables { int gperm[4], ginv[4]; double gre[4], gim[4]; double pre[4][4], pim[4][4]; int nu_f, c2_f; };

template <typename F>
__global__ void __launch_bounds__(128) seq_source_kernel(CplxT<F> *__restrict__ out, size_t V, size_t V3, int timeslice, const CplxT<F> *__restrict__ T1,
                                                        const CplxT<F> *__restrict__ T2, SeqSrcTables t, int part) {
  const size_t sid = (size_t)blockIdx.x * 128 + threadIdx.x;
  if (sid >= V3) return;
  auto ld = [&](const CplxT<F> *T, int mu, int nu, int c1, int c2) -> Cx<double> {
    const CplxT<F> v = T[((size_t)(mu * 4 + nu) * 9 + c1 * 3 + c2) * V3 + sid];
    return {(double)v.re, (double)v.im};
  };
  auto G = [&](int row) -> Cx<double> { return {t.gre[row], t.gim[row]}; };          // G[row][gperm[row]]
  auto Pj = [&](int b, int a) -> Cx<double> { return {t.pre[b][a], t.pim[b][a]}; };
  Cx<double> S[4][3];
#pragma unroll
  for (int n = 0; n < 4; n++)
#pragma unroll
    for (int w = 0; w < 3; w++) S[n][w] = {0, 0};
  const int kf = t.gperm[t.nu_f];               // G[nu_f][kf] != 0
  const int jf = t.ginv[t.nu_f];                // G[jf][nu_f] != 0
  for (int e1 = 0; e1 < 6; e1++) {
    const int u = e1 < 3 ? e1 : e1 - 3, v = e1 < 3 ? (e1 + 1) % 3 : (e1 - 3 + 2) % 3, w = 3 - u - v;
    for (int e2 = 0; e2 < 2; e2++) {
      const int U = e2 == 0 ? (t.c2_f + 1) % 3 : (t.c2_f + 2) % 3, Vc = 3 - U - t.c2_f;
      const double sg = -(((e1 < 3) == (e2 == 0)) ? 1.0 : -1.0);        // - sgn sgn'
#pragma unroll
      for (int n = 0; n < 4; n++) {
        Cx<double> acc = {0, 0};
        if (part == 2) {
          const int m = t.ginv[n];                                       // G[m][n]
          const Cx<double> gg = cmul(G(m), G(t.nu_f));
          for (int b = 0; b < 4; b++)
            for (int a = 0; a < 4; a++) {
              const Cx<double> p = Pj(b, a);
              if (p.re == 0.0 && p.im == 0.0) continue;
              Cx<double> tt = cmul(ld(T1, m, b, u, U), ld(T1, a, kf, v, Vc));
              cmac(tt, ld(T1, m, kf, u, U), ld(T1, a, b, v, Vc));
              cmac(acc, cmul(gg, p), tt);
            }
        } else {
          const int gn = t.gperm[n];                                     // m = n: G[n][gn]
          for (int j = 0; j < 4; j++) {
            const int k = t.gperm[j];
            // (m = n, b = nu_f): sum_a P[nu_f][a] T2[gn][j] T1[a][k]
            const Cx<double> c0 = cmul(cmul(G(n), G(j)), ld(T2, gn, j, u, U));
            for (int a = 0; a < 4; a++) cmac(acc, cmul(c0, Pj(t.nu_f, a)), ld(T1, a, k, v, Vc));
          }
          {  // (m = n, k = nu_f -> j = jf): sum_{a,b} P[b][a] T2[gn][jf] T1[a][b]
            const Cx<double> c0 = cmul(cmul(G(n), G(jf)), ld(T2, gn, jf, u, U));
            for (int b = 0; b < 4; b++)
              for (int a = 0; a < 4; a++) cmac(acc, cmul(c0, Pj(b, a)), ld(T1, a, b, v, Vc));
          }
          for (int m = 0; m < 4; m++) {
            const int g = t.gperm[m];
            // (a = n, b = nu_f): sum_j P[nu_f][n] T2[g][j] T1[m][k(j)]
            for (int j = 0; j < 4; j++)
              cmac(acc, cmul(cmul(cmul(G(m), G(j)), Pj(t.nu_f, n)), ld(T2, g, j, u, U)), ld(T1, m, t.gperm[j], v, Vc));
            // (a = n, k = nu_f -> j = jf): sum_b P[b][n] T2[g][jf] T1[m][b]
            const Cx<double> c1 = cmul(cmul(G(m), G(jf)), ld(T2, g, jf, u, U));
            for (int b = 0; b < 4; b++) cmac(acc, cmul(c1, Pj(b, n)), ld(T1, m, b, v, Vc));
          }
        }
        // the colour index w is a run-time value: add into the matching accumulator without indexing the register array
#pragma unroll
        for (int ww = 0; ww < 3; ww++)
          if (ww == w) { S[n][ww].re += sg * acc.re; S[n][ww].im += sg * acc.im; }
      }
    }
  }
#pragma unroll
  for (int n = 0; n < 4; n++)
#pragma unroll
    for (int w = 0; w < 3; w++) {
      CplxT<F> o; o.re = (F)S[n][w].re; o.im = (F)S[n][w].im;
      out[(size_t)(n * 3 + w) * V + (size_t)timeslice * V3 + sid] = o;
    }
}

// ultra-local insertion: C_iop(x) = sum_{n,r} Gamma_iop[n][r] sum_{m,b,a} F[r][m]^{ba}(x) S[n][m]^{ba}(x), 16 operators
struct OpTables { double re[16][16], im[16][16]; };     // [iop][n*4 + r]
__constant__ OpTables c_ops;
template <typename F>
__global__ void __launch_bounds__(CT_BLOCK) fixsink_local_site_kernel(CplxT<double> *__restrict__ csite, const CplxT<F> *__restrict__ fwd,
                                                                     const CplxT<F> *__restrict__ seq, size_t V) {
  const size_t x = (size_t)blockIdx.x * CT_BLOCK + threadIdx.x;
  if (x >= V) return;
  Cx<F> M[4][4];                                           // [n][r]
#pragma unroll
  for (int n = 0; n < 4; n++)
#pragma unroll
    for (int r = 0; r < 4; r++) M[n][r] = {0, 0};
#pragma unroll 1
  for (int mc = 0; mc < 36; mc++) {                        // (m, b, a)
    const int m = mc / 9, ba = mc - m * 9;
    Cx<F> Fr[4], Sn[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const CplxT<F> f = fwd[((size_t)(r * 4 + m) * 9 + ba) * V + x], q = seq[((size_t)(r * 4 + m) * 9 + ba) * V + x];
      Fr[r] = {f.re, f.im}; Sn[r] = {q.re, q.im};
    }
#pragma unroll
    for (int n = 0; n < 4; n++)
#pragma unroll
      for (int r = 0; r < 4; r++) cmac(M[n][r], Fr[r], Sn[n]);
  }
  for (int iop = 0; iop < 16; iop++) {
    Cx<double> acc = {0, 0};
#pragma unroll
    for (int n = 0; n < 4; n++)
#pragma unroll
      for (int r = 0; r < 4; r++) cmac<double>(acc, {c_ops.re[iop][n * 4 + r], c_ops.im[iop][n * 4 + r]}, {(double)M[n][r].re, (double)M[n][r].im});
    CplxT<double> o; o.re = acc.re; o.im = acc.im;
    csite[(size_t)iop * V + x] = o;
  }
}

// conserved-current (Noether) and one-derivative insertions (lib/code_pieces/fixSinkContractions_{noether,oneD}_core.h; formulas in
// oracle.fixsink_derivative_site, pinned to the reference's kernel bodies).  Per site and direction d the four hop blocks
//   Af = S(x) U_d(x) F(x+d),  Ab = S(x) U_d(x-d)^dag F(x-d),  Bf = S(x+d) U_d(x)^dag F(x),  Bb = S(x-d) U_d(x-d) F(x)
// (4x4 spin matrices; colours and the source spin summed) only enter as P = Ab + Bf and Q = Af + Bb:
//   noether[d] = 1/4 ( tr[(1+g_d)^T P] - tr[(1-g_d)^T Q] ),     oneD[iop][d] = 1/4 tr[Gamma_iop^T (Q - P)].
// One thread per (site, direction); periodic neighbours on this rank (the entry point refuses a split lattice).
struct DerivTables { double pg_re[4][16], pg_im[4][16], mg_re[4][16], mg_im[4][16]; };     // (1 + g_d)[k*4+l], (1 - g_d)[k*4+l]
__constant__ DerivTables c_deriv;

// M[k][l] += sum_{p, a, b, c} S[k,p]^{ab} V^{ac} F[l,p]^{cb},  V = U (DAG = false) or U^dag (DAG = true); S, F, U point at the sites to use
template <typename F, bool DAG>
__device__ __forceinline__ void hop_block(Cx<F> (&M)[4][4], const CplxT<F> *__restrict__ S, const CplxT<F> *__restrict__ Fw, const CplxT<F> *__restrict__ U,
                                          size_t V) {
  Cx<F> u[3][3];                                            // V^{ac}
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const CplxT<F> g = DAG ? U[(size_t)(c * 3 + a) * V] : U[(size_t)(a * 3 + c) * V];
      u[a][c] = {g.re, DAG ? -g.im : g.im};
    }
#pragma unroll 1
  for (int pb = 0; pb < 12; pb++) {
    const int p = pb / 3, b = pb - p * 3;
    Cx<F> W[4][3];                                          // W[l][a] = sum_c V^{ac} F[l,p]^{cb}
#pragma unroll
    for (int l = 0; l < 4; l++) {
      Cx<F> f[3];
#pragma unroll
      for (int c = 0; c < 3; c++) { const CplxT<F> v = Fw[((size_t)(l * 4 + p) * 9 + c * 3 + b) * V]; f[c] = {v.re, v.im}; }
#pragma unroll
      for (int a = 0; a < 3; a++) {
        W[l][a] = {0, 0};
#pragma unroll
        for (int c = 0; c < 3; c++) cmac(W[l][a], u[a][c], f[c]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      Cx<F> sv[3];
#pragma unroll
      for (int a = 0; a < 3; a++) { const CplxT<F> v = S[((size_t)(k * 4 + p) * 9 + a * 3 + b) * V]; sv[a] = {v.re, v.im}; }
#pragma unroll
      for (int l = 0; l < 4; l++)
#pragma unroll
        for (int a = 0; a < 3; a++) cmac(M[k][l], sv[a], W[l][a]);
    }
  }
}

// csite[ch][nsites]: ch = d (noether), 4 + d*16 + iop (oneD); sites site0 .. site0 + nsites - 1 of the local lattice
template <typename F>
__global__ void __launch_bounds__(CT_BLOCK) fixsink_deriv_site_kernel(CplxT<double> *__restrict__ csite, const CplxT<F> *__restrict__ fwd,
                                                                     const CplxT<F> *__restrict__ seq, const CplxT<F> *__restrict__ gauge, size_t V,
                                                                     size_t site0, size_t nsites, int X0, int X1, int X2, int X3) {
  const size_t i = (size_t)blockIdx.x * CT_BLOCK + threadIdx.x;
  if (i >= nsites) return;
  const int d = blockIdx.y;
  const size_t x = site0 + i;
  const int L[4] = {X0, X1, X2, X3};
  const size_t str[4] = {1, (size_t)X0, (size_t)X0 * X1, (size_t)X0 * X1 * X2};
  const int cd_ = (int)((x / str[d]) % L[d]);
  const size_t xp = cd_ == L[d] - 1 ? x - (size_t)(L[d] - 1) * str[d] : x + str[d];
  const size_t xm = cd_ == 0 ? x + (size_t)(L[d] - 1) * str[d] : x - str[d];
  const CplxT<F> *Ud = gauge + (size_t)d * 9 * V;
  Cx<F> P[4][4], Q[4][4];
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int l = 0; l < 4; l++) { P[k][l] = {0, 0}; Q[k][l] = {0, 0}; }
  hop_block<F, false>(Q, seq + x, fwd + xp, Ud + x, V);      // Af
  hop_block<F, false>(Q, seq + xm, fwd + x, Ud + xm, V);     // Bb
  hop_block<F, true>(P, seq + x, fwd + xm, Ud + xm, V);      // Ab
  hop_block<F, true>(P, seq + xp, fwd + x, Ud + x, V);       // Bf
  Cx<double> n = {0, 0};
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int l = 0; l < 4; l++) {
      const int q = k * 4 + l;
      cmac<double>(n, {c_deriv.pg_re[d][q], c_deriv.pg_im[d][q]}, {(double)P[k][l].re, (double)P[k][l].im});
      cmac<double>(n, {-c_deriv.mg_re[d][q], -c_deriv.mg_im[d][q]}, {(double)Q[k][l].re, (double)Q[k][l].im});
    }
  CplxT<double> o; o.re = 0.25 * n.re; o.im = 0.25 * n.im;
  csite[(size_t)d * nsites + i] = o;
  for (int iop = 0; iop < 16; iop++) {
    Cx<double> acc = {0, 0};
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
      for (int l = 0; l < 4; l++)
        cmac<double>(acc, {c_ops.re[iop][k * 4 + l], c_ops.im[iop][k * 4 + l]}, {(double)Q[k][l].re - (double)P[k][l].re, (double)Q[k][l].im - (double)P[k][l].im});
    o.re = 0.25 * acc.re; o.im = 0.25 * acc.im;
    csite[(size_t)(4 + d * 16 + iop) * nsites + i] = o;
  }
}

// ---- one axis of the separable Fourier sum -----------------------------------------------------------------------------
// in [ch][parent][outer][L] (L fastest) -> out[ch][child][outer]:  out = sum_k tab[q][k] in[.., k] for every entry q of the parent's
// child list (child_start / child_out); one warp per input row, lanes stride the row, fixed-order shuffle tree.
struct DftStage {
  const CplxT<double> *in;
  CplxT<double> *out;
  const CplxT<double> *tab;      // [nlist][L]
  const int *child_start;        // [nparent + 1]
  const int *child_out;          // [nlist] -> output index
  int nparent, nout, L, nch;
  size_t outer;
};
constexpr int DFT_MAXK = 8;      // L <= 256
__global__ void __launch_bounds__(256) axis_dft_kernel(DftStage s) {
  const size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const size_t nrows = (size_t)s.nch * s.nparent * s.outer;
  if (row >= nrows) return;
  const size_t o = row % s.outer, cp = row / s.outer;
  const int p = (int)(cp % s.nparent), ch = (int)(cp / s.nparent);
  double vr[DFT_MAXK], vi[DFT_MAXK];
#pragma unroll
  for (int j = 0; j < DFT_MAXK; j++) {
    const int k = lane + 32 * j;
    if (k < s.L) { const CplxT<double> v = s.in[row * s.L + k]; vr[j] = v.re; vi[j] = v.im; } else { vr[j] = 0; vi[j] = 0; }
  }
  for (int q = s.child_start[p]; q < s.child_start[p + 1]; q++) {
    double ar = 0, ai = 0;
#pragma unroll
    for (int j = 0; j < DFT_MAXK; j++) {
      const int k = lane + 32 * j;
      if (k < s.L) {
        const CplxT<double> e = s.tab[(size_t)q * s.L + k];
        ar += e.re * vr[j] - e.im * vi[j];
        ai += e.re * vi[j] + e.im * vr[j];
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { ar += __shfl_down_sync(0xffffffffu, ar, off); ai += __shfl_down_sync(0xffffffffu, ai, off); }
    if (lane == 0) { CplxT<double> r; r.re = ar; r.im = ai; s.out[((size_t)ch * s.nout + s.child_out[q]) * s.outer + o] = r; }
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------------
typedef std::complex<double> cd;
struct Mat4 { cd m[4][4]; };
static Mat4 mul(const Mat4 &a, const Mat4 &b) {
  Mat4 r;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { cd s = 0; for (int k = 0; k < 4; k++) s += a.m[i][k] * b.m[k][j]; r.m[i][j] = s; }
  return r;
}
// UKQCD gamma matrices (lib/code_pieces/gammas_tm_base.h:21-32; gamma5 = g1 g2 g3 g4 = spin swap, apply_gamma5_vector_core.h)
static void ukqcd_gammas(Mat4 g[5]) {
  const cd I(0, 1);
  for (int k = 0; k < 5; k++) for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) g[k].m[i][j] = 0;
  g[0].m[0][3] = I; g[0].m[1][2] = I; g[0].m[2][1] = -I; g[0].m[3][0] = -I;
  g[1].m[0][3] = 1; g[1].m[1][2] = -1; g[1].m[2][1] = -1; g[1].m[3][0] = 1;
  g[2].m[0][2] = I; g[2].m[1][3] = -I; g[2].m[2][0] = -I; g[2].m[3][1] = I;
  g[3].m[0][0] = 1; g[3].m[1][1] = 1; g[3].m[2][2] = -1; g[3].m[3][3] = -1;
  g[4] = mul(mul(g[0], g[1]), mul(g[2], g[3]));
}
// signs s_G phi_a conj(phi_g) of the ten channels, derived from the gamma matrices; fails if a channel's permutation is not the
// XOR the site kernel was compiled for, or if a weight is not +-1
static int meson_signs(MesonSigns<double> *W) {
  Mat4 g[5];
  ukqcd_gammas(g);
  Mat4 one;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) one.m[i][j] = i == j ? 1.0 : 0.0;
  Mat4 G[10] = {g[4], one, mul(g[4], g[0]), mul(g[4], g[1]), mul(g[4], g[2]), mul(g[4], g[3]), g[0], g[1], g[2], g[3]};
  const int xors[10] = TMQ_MESON_XORS;
  for (int a = 0; a < 4; a++)
    for (int b = 0; b < 4; b++)
      if (std::abs(g[4].m[a][b] - ((a ^ 2) == b ? 1.0 : 0.0)) > 1e-15) { set_error("gamma5 is not the spin swap"); return 1; }
  for (int ip = 0; ip < 10; ip++) {
    const double sG = ip < 6 ? 1.0 : -1.0;
    cd phi[4];
    for (int a = 0; a < 4; a++) {
      for (int b = 0; b < 4; b++)
        if ((b == (a ^ xors[ip])) != (std::abs(G[ip].m[a][b]) > 0.5)) { set_error("meson channel %d: unexpected spin permutation", ip); return 1; }
      phi[a] = G[ip].m[a][a ^ xors[ip]];
    }
    for (int a = 0; a < 4; a++)
      for (int b = 0; b < 4; b++) {
        const cd w = sG * phi[a] * std::conj(phi[b]);
        if (std::abs(w.imag()) > 1e-15 || std::abs(std::abs(w.real()) - 1.0) > 1e-15) { set_error("meson channel %d: weight is not +-1", ip); return 1; }
        W->w[ip][a * 4 + b] = w.real();
      }
  }
  return 0;
}


// one entry per row? -> permutation + phases
static int to_mono(const Mat4 &m, Mono *o, const char *what) {
  for (int r = 0; r < 4; r++) {
    int n = 0;
    for (int cidx = 0; cidx < 4; cidx++)
      if (std::abs(m.m[r][cidx]) > 1e-12) { o->perm[r] = cidx; o->re[r] = (float)m.m[r][cidx].real(); o->im[r] = (float)m.m[r][cidx].imag(); n++; }
    if (n != 1) { set_error("baryon tables: %s is not a monomial matrix", what); return 1; }
  }
  return 0;
}
static Mat4 conj4(const Mat4 &a) { Mat4 r; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = std::conj(a.m[i][j]); return r; }
// the channel / term tables, derived from the UKQCD gamma matrices (C = g4 g2)
static int baryon_tables(BaryonTables *t) {
  Mat4 g[5];
  ukqcd_gammas(g);
  Mat4 one;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) one.m[i][j] = i == j ? 1.0 : 0.0;
  const Mat4 C = mul(g[3], g[1]), Cg5 = mul(C, g[4]);
  memset(t, 0, sizeof(*t));
  // terms: nucleon-type (2), Delta++ (6), Delta+ (8)  (contractBaryons_core.h:68-69, 360-366, 445-453)
  const BaryonTerm nucl[2] = {{+1., {0, 1, 0}, {0, 1, 2}}, {-1., {0, 1, 0}, {2, 1, 0}}};
  const BaryonTerm dpp[6] = {{+1., {0, 0, 0}, {1, 2, 0}}, {-1., {0, 0, 0}, {2, 1, 0}}, {+1., {0, 0, 0}, {2, 0, 1}},
                             {-1., {0, 0, 0}, {0, 2, 1}}, {-1., {0, 0, 0}, {1, 0, 2}}, {+1., {0, 0, 0}, {0, 1, 2}}};
  const double th = 1. / 3.;
  const BaryonTerm dp[8] = {{-4 * th, {0, 1, 0}, {2, 1, 0}}, {+2 * th, {0, 1, 0}, {1, 2, 0}}, {+2 * th, {0, 0, 1}, {2, 0, 1}},
                            {-2 * th, {0, 0, 1}, {0, 2, 1}}, {-2 * th, {0, 1, 0}, {0, 2, 1}}, {-1 * th, {0, 0, 1}, {1, 0, 2}},
                            {+1 * th, {0, 0, 1}, {0, 1, 2}}, {+4 * th, {0, 1, 0}, {0, 1, 2}}};
  memcpy(&t->term[0], nucl, sizeof(nucl)); memcpy(&t->term[2], dpp, sizeof(dpp)); memcpy(&t->term[8], dp, sizeof(dp));
  struct Def { Mat4 gs, gr, xs, xr; int term0, nterm; };
  std::vector<Def> defs = {{Cg5, Cg5, one, one, 0, 2}, {Cg5, C, one, g[4], 0, 2}, {C, Cg5, g[4], one, 0, 2}, {C, C, g[4], g[4], 0, 2}};
  for (int k = 0; k < 3; k++) defs.push_back({mul(C, g[k]), mul(C, g[k]), one, one, 2, 6});
  for (int k = 0; k < 3; k++) defs.push_back({mul(C, g[k]), mul(C, g[k]), one, one, 8, 8});
  for (int ip = 0; ip < 10; ip++) {
    BaryonChannel &c = t->ch[ip];
    Mono xs, xr;
    if (to_mono(defs[ip].gs, &c.gs, "Gs") || to_mono(conj4(defs[ip].gr), &c.grc, "conj(Gr)") || to_mono(defs[ip].xs, &xs, "Xs") || to_mono(defs[ip].xr, &xr, "Xr")) return 1;
    for (int gidx = 0; gidx < 4; gidx++) {      // Xs[g][d] != 0 for d = xs.perm[g]: entry d of R goes to row g
      c.xs_inv[xs.perm[gidx]] = gidx; c.xs_re[xs.perm[gidx]] = xs.re[gidx]; c.xs_im[xs.perm[gidx]] = xs.im[gidx];
      c.xr_inv[xr.perm[gidx]] = gidx; c.xr_re[xr.perm[gidx]] = xr.re[gidx]; c.xr_im[xr.perm[gidx]] = xr.im[gidx];
    }
    c.term0 = defs[ip].term0; c.nterm = defs[ip].nterm;
  }
  return 0;
}

// 1/2 (1 + i s g5) M (1 + i s g5): physical -> twisted basis at maximal twist
static Mat4 twist_rotate(const Mat4 &M, const Mat4 &g5, double s) {
  Mat4 R, out;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) R.m[i][j] = (i == j ? 1.0 : 0.0) + cd(0, s) * g5.m[i][j];
  out = mul(mul(R, M), R);
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out.m[i][j] *= 0.5;
  return out;
}
static Mat4 add4(const Mat4 &a, const Mat4 &b) { Mat4 r; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r; }
static Mat4 scale4(const Mat4 &a, cd z) { Mat4 r; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = z * a.m[i][j]; return r; }
static Mat4 eye4() { Mat4 r; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = i == j ? 1.0 : 0.0; return r; }
// projectors_tm_base.h as a formula: 1/4 (1 + g4) [ i g5 g_k | summed over k ] rotated with s = +1 (proton) / -1 (neutron);
// pid: 0 G4, 1 G5G123, 2..4 G5G1..G5G3 (WHICHPROJECTOR, include/qudaQKXTM_utils.h:129)
static Mat4 projector_tm(int pid, int particle) {
  Mat4 g[5];
  ukqcd_gammas(g);
  const Mat4 P0 = scale4(add4(eye4(), g[3]), 0.25);
  Mat4 Pk[3];
  for (int k = 0; k < 3; k++) Pk[k] = mul(P0, scale4(mul(g[4], g[k]), cd(0, 1)));
  Mat4 phys = pid == 0 ? P0 : (pid == 1 ? add4(add4(Pk[0], Pk[1]), Pk[2]) : Pk[pid - 2]);
  return twist_rotate(phys, g[4], particle == 0 ? +1.0 : -1.0);
}
// gammas_tm_base.h as a formula: 1, g1..g4, g5, g5 g1..g5 g4, -i sigma_{12,13,23,41,42,43} rotated with s = +1 for (proton, part 1) and
// (neutron, part 2), -1 otherwise
static void operator_tables(OpTables *t, int particle, int partflag) {
  Mat4 g[5];
  ukqcd_gammas(g);
  std::vector<Mat4> ops = {eye4(), g[0], g[1], g[2], g[3], g[4], mul(g[4], g[0]), mul(g[4], g[1]), mul(g[4], g[2]), mul(g[4], g[3])};
  const int pr[6][2] = {{0, 1}, {0, 2}, {1, 2}, {3, 0}, {3, 1}, {3, 2}};
  for (int q = 0; q < 6; q++) {
    const Mat4 ab = mul(g[pr[q][0]], g[pr[q][1]]), ba = mul(g[pr[q][1]], g[pr[q][0]]);
    ops.push_back(scale4(add4(ab, scale4(ba, -1.0)), cd(0, -0.5)));          // -i * 1/2 [g_a, g_b]
  }
  const double s = ((particle == 0) == (partflag == 1)) ? +1.0 : -1.0;
  for (int iop = 0; iop < 16; iop++) {
    const Mat4 G = twist_rotate(ops[iop], g[4], s);
    for (int n = 0; n < 4; n++) for (int r = 0; r < 4; r++) { t->re[iop][n * 4 + r] = G.m[n][r].real(); t->im[iop][n * 4 + r] = G.m[n][r].imag(); }
  }
}

// carve-out of the context's grow-only contraction work space (256-byte aligned pieces)
struct WsPlan {
  size_t total = 0;
  size_t add(size_t bytes) { const size_t off = total; total += (bytes + 255) & ~(size_t)255; return off; }
};
static int ensure_contract_ws(tmq_ctx *c, size_t bytes) {
  if (c->contract_ws_bytes >= bytes) return 0;
  if (c->contract_ws) { TMQ_CUDA(cudaStreamSynchronize(c->stream)); TMQ_CUDA(cudaFree(c->contract_ws)); c->contract_ws = nullptr; c->contract_ws_bytes = 0; }
  TMQ_CUDA(cudaMalloc(&c->contract_ws, bytes));
  c->contract_ws_bytes = bytes;
  return 0;
}
template <typename T> static cudaError_t upload(void *dst, const std::vector<T> &h, cudaStream_t st) {
  return cudaMemcpyAsync(dst, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st);
}

static cudaError_t run_stage(const DftStage &s, cudaStream_t st) {
  const size_t nrows = (size_t)s.nch * s.nparent * s.outer;
  axis_dft_kernel<<<(unsigned int)((nrows + 7) / 8), 256, 0, st>>>(s);
  return cudaGetLastError();
}

// exp(sgn 2 pi i q (k + off - k0) / Ltot) for k = 0..L-1 (sgn = -1: two-point functions, +1: the three-point insertion)
static void phase_row(std::vector<CplxT<double>> &tab, int q, int L, int off, int k0, int Ltot, int sgn = -1) {
  for (int k = 0; k < L; k++) {
    // reduce the integer numerator first: the phase is exact to the last bit of the argument
    long long num = ((long long)q * (k + off - k0)) % Ltot;
    const double ph = 2.0 * M_PI * (double)num / (double)Ltot;
    CplxT<double> e; e.re = cos(ph); e.im = sgn * sin(ph);
    tab.push_back(e);
  }
}


// ---- the separable momentum projection as a reusable plan ------------------------------------------------------------------------
// Built once per call from the momentum list; project() runs the three axis stages on the site values of nt consecutive local
// time slices, csite[ch][nt*Z*Y*X], and adds the result into this rank's slices of the caller's GLOBAL-T host array
// corr[t global][imom][ch][re,im].
struct MomProjector {
  tmq_ctx *c = nullptr;
  int nmoms = 0, npx = 0, npair = 0, X = 0, Y = 0, Z = 0, T = 0;
  std::vector<CplxT<double>> tab1, tab2, tab3;
  std::vector<int> start1, out1, start2, out2, start3, out3;
  // byte offsets into the work space (filled by plan())
  size_t o_w1 = 0, o_w2 = 0, o_w3 = 0, o_tab1 = 0, o_tab2 = 0, o_tab3 = 0, o_s1 = 0, o_s2 = 0, o_s3 = 0, o_o1 = 0, o_o2 = 0, o_o3 = 0;

  void build(tmq_ctx *ctx, const int *moms, int n, const int src_pos[3], int sgn = -1) {
    c = ctx; nmoms = n;
    X = c->g.X[0]; Y = c->g.X[1]; Z = c->g.X[2]; T = c->g.X[3];
    std::vector<int> px_list;
    std::map<int, int> px_idx;
    std::vector<std::pair<int, int>> pair_list;      // (ix, py), grouped by ix
    std::map<std::pair<int, int>, int> pair_idx;
    for (int m = 0; m < nmoms; m++) if (!px_idx.count(moms[3 * m])) { px_idx[moms[3 * m]] = (int)px_list.size(); px_list.push_back(moms[3 * m]); }
    for (int ix = 0; ix < (int)px_list.size(); ix++)
      for (int m = 0; m < nmoms; m++) {
        if (px_idx[moms[3 * m]] != ix) continue;
        const std::pair<int, int> key(ix, moms[3 * m + 1]);
        if (!pair_idx.count(key)) { pair_idx[key] = (int)pair_list.size(); pair_list.push_back(key); }
      }
    npx = (int)px_list.size(); npair = (int)pair_list.size();
    const int gX = X * c->grid[0], gY = Y * c->grid[1], gZ = Z * c->grid[2];
    // stage 1 (x): one parent, children = all p_x
    start1 = {0, npx}; out1.assign(npx, 0); start2.assign(npx + 1, 0); start3.assign(npair + 1, 0);
    for (int ix = 0; ix < npx; ix++) { out1[ix] = ix; phase_row(tab1, px_list[ix], X, c->coord[0] * X, src_pos[0], gX, sgn); }
    // stage 2 (y): parent ix -> its pairs (contiguous by construction)
    for (int ix = 0; ix < npx; ix++) {
      start2[ix] = (int)out2.size();
      for (int j = 0; j < npair; j++) if (pair_list[j].first == ix) { out2.push_back(j); phase_row(tab2, pair_list[j].second, Y, c->coord[1] * Y, src_pos[1], gY, sgn); }
    }
    start2[npx] = (int)out2.size();
    // stage 3 (z): parent pair -> the momenta with that (p_x, p_y), output index = the caller's momentum index
    for (int j = 0; j < npair; j++) {
      start3[j] = (int)out3.size();
      for (int m = 0; m < nmoms; m++)
        if (px_idx[moms[3 * m]] == pair_list[j].first && moms[3 * m + 1] == pair_list[j].second) { out3.push_back(m); phase_row(tab3, moms[3 * m + 2], Z, c->coord[2] * Z, src_pos[2], gZ, sgn); }
    }
    start3[npair] = (int)out3.size();
  }
  // reserve the stage buffers for nch channels x nt time slices and the tables
  void plan(WsPlan &p, int nch, int nt) {
    const size_t CB = sizeof(CplxT<double>);
    o_w1 = p.add((size_t)nch * npx * nt * Z * Y * CB); o_w2 = p.add((size_t)nch * npair * nt * Z * CB); o_w3 = p.add((size_t)nch * nmoms * nt * CB);
    o_tab1 = p.add(tab1.size() * CB); o_tab2 = p.add(tab2.size() * CB); o_tab3 = p.add(tab3.size() * CB);
    o_s1 = p.add(start1.size() * sizeof(int)); o_s2 = p.add(start2.size() * sizeof(int)); o_s3 = p.add(start3.size() * sizeof(int));
    o_o1 = p.add(out1.size() * sizeof(int)); o_o2 = p.add(out2.size() * sizeof(int)); o_o3 = p.add(out3.size() * sizeof(int));
  }
  int upload_tables(char *ws, cudaStream_t st) {
    TMQ_CUDA(upload(ws + o_tab1, tab1, st)); TMQ_CUDA(upload(ws + o_tab2, tab2, st)); TMQ_CUDA(upload(ws + o_tab3, tab3, st));
    TMQ_CUDA(upload(ws + o_s1, start1, st)); TMQ_CUDA(upload(ws + o_s2, start2, st)); TMQ_CUDA(upload(ws + o_s3, start3, st));
    TMQ_CUDA(upload(ws + o_o1, out1, st)); TMQ_CUDA(upload(ws + o_o2, out2, st)); TMQ_CUDA(upload(ws + o_o3, out3, st));
    return 0;
  }
  // csite: [nch][nt*Z*Y*X] site values of the local time slices t0 .. t0+nt-1; corr: [gT][nmoms][nch][2] host doubles
  int project(char *ws, const CplxT<double> *csite, int nch, int t0, int nt, double *corr, cudaStream_t st) {
    typedef const CplxT<double> *cptr;
    const size_t outer1 = (size_t)nt * Z * Y, outer2 = (size_t)nt * Z, outer3 = (size_t)nt;
    DftStage s1 = {csite, (CplxT<double> *)(ws + o_w1), (cptr)(ws + o_tab1), (const int *)(ws + o_s1), (const int *)(ws + o_o1), 1, npx, X, nch, outer1};
    DftStage s2 = {(cptr)(ws + o_w1), (CplxT<double> *)(ws + o_w2), (cptr)(ws + o_tab2), (const int *)(ws + o_s2), (const int *)(ws + o_o2), npx, npair, Y, nch, outer2};
    DftStage s3 = {(cptr)(ws + o_w2), (CplxT<double> *)(ws + o_w3), (cptr)(ws + o_tab3), (const int *)(ws + o_s3), (const int *)(ws + o_o3), npair, nmoms, Z, nch, outer3};
    TMQ_CUDA(run_stage(s1, st)); TMQ_CUDA(run_stage(s2, st)); TMQ_CUDA(run_stage(s3, st));
    c->launches += 3;
    std::vector<double> h((size_t)2 * nch * nmoms * nt);
    TMQ_CUDA(cudaMemcpyAsync(h.data(), ws + o_w3, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    TMQ_CUDA(cudaStreamSynchronize(st));
    const int t_off = c->coord[3] * T;
    for (int t = 0; t < nt; t++)
      for (int m = 0; m < nmoms; m++)
        for (int ch = 0; ch < nch; ch++) {
          const size_t src = (((size_t)ch * nmoms + m) * nt + t) * 2, dst = ((((size_t)(t0 + t + t_off)) * nmoms + m) * nch + ch) * 2;
          corr[dst] = h[src]; corr[dst + 1] = h[src + 1];
        }
    return 0;
  }
};
// sum over the z ranks and gather over the t ranks in one all-reduce of the zero-padded global-T host buffer
static int allreduce_host(tmq_ctx *c, char *ws, size_t off, double *corr, size_t n, cudaStream_t st) {
  if (c->nranks <= 1) return 0;
  double *g = (double *)(ws + off);
  TMQ_CUDA(cudaMemcpyAsync(g, corr, n * sizeof(double), cudaMemcpyHostToDevice, st));
  TMQ_REQUIRE(n > 4 && n < ((size_t)1 << 31), "internal: reduction length out of range");      // > 4 doubles: always the NCCL path of comm_allreduce
  TMQ_TRY(comm_allreduce(c, g, (int)n, st));
  TMQ_CUDA(cudaMemcpyAsync(corr, g, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  TMQ_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // namespace tmq

using namespace tmq;

extern "C" {

int tmq_qkxtm_conjugate(tmq_ctx *c, void *d, int prec, int ncomp) {
  TMQ_REQUIRE(c && d, "null argument");
  TMQ_REQUIRE(prec == 8 || prec == 4, "bad precision");
  TMQ_REQUIRE(ncomp > 0, "bad component count");
  TMQ_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)2 * c->g.Vh * ncomp;
  const unsigned int grid = (unsigned int)((n + 255) / 256);
  if (prec == 8) qk_conj_kernel<double><<<grid, 256, 0, c->stream>>>((CplxT<double> *)d, n);
  else qk_conj_kernel<float><<<grid, 256, 0, c->stream>>>((CplxT<float> *)d, n);
  TMQ_CUDA(cudaGetLastError());
  c->launches++;
  return 0;
}

int tmq_qkxtm_gamma5_prop(tmq_ctx *c, void *d_prop, int prec) {
  TMQ_REQUIRE(c && d_prop, "null argument");
  TMQ_REQUIRE(prec == 8 || prec == 4, "bad precision");
  TMQ_CUDA(cudaSetDevice(c->device));
  const size_t V = (size_t)2 * c->g.Vh, n = V * 72;
  const unsigned int grid = (unsigned int)((n + 255) / 256);
  if (prec == 8) qk_gamma5_prop_kernel<double><<<grid, 256, 0, c->stream>>>((CplxT<double> *)d_prop, V);
  else qk_gamma5_prop_kernel<float><<<grid, 256, 0, c->stream>>>((CplxT<float> *)d_prop, V);
  TMQ_CUDA(cudaGetLastError());
  c->launches++;
  return 0;
}

int tmq_qkxtm_rotate_physical(tmq_ctx *c, void *d_prop, int prec, int sign) {
  TMQ_REQUIRE(c && d_prop, "null argument");
  TMQ_REQUIRE(prec == 8 || prec == 4, "bad precision");
  TMQ_REQUIRE(sign == 1 || sign == -1, "The sign can be only +-1");      // lib/qudaQKXTM_Propagator.cpp:110
  TMQ_CUDA(cudaSetDevice(c->device));
  const size_t V = (size_t)2 * c->g.Vh, n = V * 9;
  const unsigned int grid = (unsigned int)((n + CT_BLOCK - 1) / CT_BLOCK);
  if (prec == 8) qk_rotate_kernel<double><<<grid, CT_BLOCK, 0, c->stream>>>((CplxT<double> *)d_prop, V, (double)sign);
  else qk_rotate_kernel<float><<<grid, CT_BLOCK, 0, c->stream>>>((CplxT<float> *)d_prop, V, (float)sign);
  TMQ_CUDA(cudaGetLastError());
  c->launches++;
  return 0;
}

int tmq_qkxtm_column_copy(tmq_ctx *c, void *d_prop, long long prop_sites, long long prop_site0, void *d_vec, long long vec_sites,
                          long long vec_site0, long long nsites, int prec, int nu, int c2, int to_prop) {
  TMQ_REQUIRE(c && d_prop && d_vec, "null argument");
  TMQ_REQUIRE(prec == 8 || prec == 4, "bad precision");
  TMQ_REQUIRE(nu >= 0 && nu < 4 && c2 >= 0 && c2 < 3, "bad column");
  TMQ_REQUIRE(nsites > 0 && prop_site0 >= 0 && vec_site0 >= 0 && prop_site0 + nsites <= prop_sites && vec_site0 + nsites <= vec_sites,
              "site range out of bounds");
  TMQ_CUDA(cudaSetDevice(c->device));
  const size_t cb = (size_t)2 * prec;
  for (int mu = 0; mu < 4; mu++)
    for (int c1 = 0; c1 < 3; c1++) {
      char *p = (char *)d_prop + (((size_t)(mu * 4 + nu) * 9 + c1 * 3 + c2) * prop_sites + prop_site0) * cb;
      char *v = (char *)d_vec + ((size_t)(mu * 3 + c1) * vec_sites + vec_site0) * cb;
      TMQ_CUDA(cudaMemcpyAsync(to_prop ? p : v, to_prop ? v : p, (size_t)nsites * cb, cudaMemcpyDeviceToDevice, c->stream));
    }
  return 0;
}

int tmq_qkxtm_contract_mesons(tmq_ctx *c, const void *d_prop1, const void *d_prop2, int prec, const int *moms, int nmoms,
                              const int src_pos[3], double *corr_mom, double *corr_pos) {
  TMQ_REQUIRE(c && d_prop1 && d_prop2, "null argument");
  TMQ_REQUIRE(prec == 8 || prec == 4, "bad precision");
  TMQ_REQUIRE(corr_mom || corr_pos, "no output requested");
  TMQ_REQUIRE(!corr_mom || (moms && nmoms > 0 && src_pos), "momentum-space output needs a momentum list and a source position");
  TMQ_CUDA(cudaSetDevice(c->device));
  static MesonSigns<double> W;
  static MesonSigns<float> Wf;
  static bool have_w = false;
  if (!have_w) {
    TMQ_TRY(meson_signs(&W));
    for (int ip = 0; ip < 10; ip++) for (int q = 0; q < 16; q++) Wf.w[ip][q] = (float)W.w[ip][q];
    have_w = true;
  }
  const int X = c->g.X[0], Y = c->g.X[1], Z = c->g.X[2], T = c->g.X[3];
  const size_t V = (size_t)2 * c->g.Vh;
  TMQ_REQUIRE(X <= 32 * DFT_MAXK && Y <= 32 * DFT_MAXK && Z <= 32 * DFT_MAXK, "spatial extent above %d not supported", 32 * DFT_MAXK);
  cudaStream_t st = c->stream;

  MomProjector mp;
  if (corr_mom) mp.build(c, moms, nmoms, src_pos);
  const int gT = T * c->grid[3];
  const size_t ntot = corr_mom ? (size_t)gT * nmoms * 40 : 0;
  WsPlan plan;
  const size_t o_csite = plan.add(V * 20 * sizeof(CplxT<double>));
  if (corr_mom) mp.plan(plan, 20, T);
  const size_t o_glob = plan.add(ntot * sizeof(double));
  TMQ_TRY(ensure_contract_ws(c, plan.total));
  char *ws = (char *)c->contract_ws;
  CplxT<double> *csite = (CplxT<double> *)(ws + o_csite);

  const dim3 grid((unsigned int)((V + CT_BLOCK - 1) / CT_BLOCK), 2);
  if (prec == 8) meson_site_kernel<double><<<grid, CT_BLOCK, 0, st>>>(csite, (const CplxT<double> *)d_prop1, (const CplxT<double> *)d_prop2, V, W);
  else meson_site_kernel<float><<<grid, CT_BLOCK, 0, st>>>(csite, (const CplxT<float> *)d_prop1, (const CplxT<float> *)d_prop2, V, Wf);
  TMQ_CUDA(cudaGetLastError());
  c->launches++;

  if (corr_pos) {
    // position space: [t][spatial site][iu][ip][re,im] like the reference's corr[2*sv + 2*SpVol*it + ri][pt][mes] (kernels.cu:1160-1165)
    std::vector<double> h(V * 40);
    TMQ_CUDA(cudaMemcpyAsync(h.data(), csite, V * 40 * sizeof(double), cudaMemcpyDeviceToHost, st));
    TMQ_CUDA(cudaStreamSynchronize(st));
    for (size_t x = 0; x < V; x++)
      for (int ch = 0; ch < 20; ch++) { corr_pos[(x * 20 + ch) * 2] = h[((size_t)ch * V + x) * 2]; corr_pos[(x * 20 + ch) * 2 + 1] = h[((size_t)ch * V + x) * 2 + 1]; }
  }
  if (!corr_mom) return 0;

  for (size_t i = 0; i < ntot; i++) corr_mom[i] = 0.0;
  TMQ_TRY(mp.upload_tables(ws, st));
  TMQ_TRY(mp.project(ws, csite, 20, 0, T, corr_mom, st));
  TMQ_TRY(allreduce_host(c, ws, o_glob, corr_mom, ntot, st));
  return 0;
}

int tmq_qkxtm_contract_baryons(tmq_ctx *c, const void *d_prop1, const void *d_prop2, int prec, const int *moms, int nmoms,
                               const int src_pos[3], double *corr_mom) {
  TMQ_REQUIRE(c && d_prop1 && d_prop2 && corr_mom && moms && src_pos, "null argument");
  TMQ_REQUIRE(prec == 8 || prec == 4, "bad precision");
  TMQ_REQUIRE(nmoms > 0, "empty momentum list");
  TMQ_CUDA(cudaSetDevice(c->device));
  // __constant__ memory is per device: upload on every call (2.6 KB) rather than tracking which devices have it
  static BaryonTables tables;
  static bool have_tables = false;
  if (!have_tables) { TMQ_TRY(baryon_tables(&tables)); have_tables = true; }
  TMQ_CUDA(cudaMemcpyToSymbolAsync(c_baryon, &tables, sizeof(tables), 0, cudaMemcpyHostToDevice, c->stream));
  const int X = c->g.X[0], Y = c->g.X[1], Z = c->g.X[2], T = c->g.X[3];
  const size_t V = (size_t)2 * c->g.Vh, V3 = (size_t)X * Y * Z;
  TMQ_REQUIRE(X <= 32 * DFT_MAXK && Y <= 32 * DFT_MAXK && Z <= 32 * DFT_MAXK, "spatial extent above %d not supported", 32 * DFT_MAXK);
  cudaStream_t st = c->stream;
  const int NCH = 320;                                    // 2 propagator assignments x 10 channels x 4 x 4 spin components
  // time slices per pass: the site values of a pass (320 complex doubles per site) are held to about 2 GiB
  int nt = (int)(((size_t)2 << 30) / (V3 * NCH * sizeof(CplxT<double>)));
  if (c->opt_contract_slices > 0 && c->opt_contract_slices < nt) nt = c->opt_contract_slices;
  if (nt < 1) nt = 1;
  if (nt > T) nt = T;
  MomProjector mp;
  mp.build(c, moms, nmoms, src_pos);
  const int gT = T * c->grid[3];
  const size_t ntot = (size_t)gT * nmoms * NCH * 2;
  WsPlan plan;
  const size_t o_csite = plan.add((size_t)nt * V3 * NCH * sizeof(CplxT<double>));
  mp.plan(plan, NCH, nt);
  const size_t o_glob = plan.add(ntot * sizeof(double));
  TMQ_TRY(ensure_contract_ws(c, plan.total));
  char *ws = (char *)c->contract_ws;
  CplxT<double> *csite = (CplxT<double> *)(ws + o_csite);
  TMQ_TRY(mp.upload_tables(ws, st));
  for (size_t i = 0; i < ntot; i++) corr_mom[i] = 0.0;
  const size_t smem = (size_t)2 * 144 * BY_SITES * 2 * prec;
  if (prec == 8) TMQ_CUDA(cudaFuncSetAttribute(baryon_site_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else TMQ_CUDA(cudaFuncSetAttribute(baryon_site_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int t0 = 0; t0 < T; t0 += nt) {
    const int n = t0 + nt <= T ? nt : T - t0;
    const size_t nsites = (size_t)n * V3, site0 = (size_t)t0 * V3;
    const unsigned int grid = (unsigned int)((nsites + BY_SITES - 1) / BY_SITES);
    if (prec == 8) baryon_site_kernel<double><<<grid, BY_THREADS, smem, st>>>(csite, (const CplxT<double> *)d_prop1, (const CplxT<double> *)d_prop2, V, site0, nsites);
    else baryon_site_kernel<float><<<grid, BY_THREADS, smem, st>>>(csite, (const CplxT<float> *)d_prop1, (const CplxT<float> *)d_prop2, V, site0, nsites);
    TMQ_CUDA(cudaGetLastError());
    c->launches++;
    TMQ_TRY(mp.project(ws, csite, NCH, t0, n, corr_mom, st));
  }
  TMQ_TRY(allreduce_host(c, ws, o_glob, corr_mom, ntot, st));
  return 0;
}

int tmq_qkxtm_seq_source(tmq_ctx *c, void *d_vec_out, int timeslice, const void *d_prop3d_1, const void *d_prop3d_2, int prec, int nu, int c2,
                         int pid, int particle, int part) {
  TMQ_REQUIRE(c && d_vec_out && d_prop3d_1, "null argument");
  TMQ_REQUIRE(prec == 8 || prec == 4, "bad precision");
  TMQ_REQUIRE(part == 1 || part == 2, "part must be 1 or 2");
  TMQ_REQUIRE(part == 2 || d_prop3d_2, "part 1 needs two 3-d propagators");
  TMQ_REQUIRE(nu >= 0 && nu < 4 && c2 >= 0 && c2 < 3, "bad source spin / colour");
  TMQ_REQUIRE(pid >= 0 && pid < 5 && (particle == 0 || particle == 1), "bad projector / particle");
  TMQ_REQUIRE(timeslice >= 0 && timeslice < c->g.X[3], "time slice outside the local lattice");
  TMQ_CUDA(cudaSetDevice(c->device));
  Mat4 g[5];
  ukqcd_gammas(g);
  Mono gm;
  TMQ_TRY(to_mono(mul(mul(g[3], g[1]), g[4]), &gm, "C g5"));
  SeqSrcTables t;
  for (int r = 0; r < 4; r++) { t.gperm[r] = gm.perm[r]; t.ginv[gm.perm[r]] = r; t.gre[r] = gm.re[r]; t.gim[r] = gm.im[r]; }
  const Mat4 P = projector_tm(pid, particle);
  for (int b = 0; b < 4; b++) for (int a = 0; a < 4; a++) { t.pre[b][a] = std::abs(P.m[b][a]) > 1e-3 ? P.m[b][a].real() : 0.0; t.pim[b][a] = std::abs(P.m[b][a]) > 1e-3 ? P.m[b][a].imag() : 0.0; }
  t.nu_f = nu; t.c2_f = c2;
  const size_t V = (size_t)2 * c->g.Vh, V3 = V / c->g.X[3];
  const unsigned int grid = (unsigned int)((V3 + 127) / 128);
  if (prec == 8) seq_source_kernel<double><<<grid, 128, 0, c->stream>>>((CplxT<double> *)d_vec_out, V, V3, timeslice, (const CplxT<double> *)d_prop3d_1, (const CplxT<double> *)d_prop3d_2, t, part);
  else seq_source_kernel<float><<<grid, 128, 0, c->stream>>>((CplxT<float> *)d_vec_out, V, V3, timeslice, (const CplxT<float> *)d_prop3d_1, (const CplxT<float> *)d_prop3d_2, t, part);
  TMQ_CUDA(cudaGetLastError());
  c->launches++;
  return 0;
}

int tmq_qkxtm_fixsink_local(tmq_ctx *c, const void *d_seq_prop, const void *d_fwd_prop, int prec, int particle, int partflag, const int *moms,
                            int nmoms, const int src_pos[3], double *corr_mom) {
  TMQ_REQUIRE(c && d_seq_prop && d_fwd_prop && moms && src_pos && corr_mom, "null argument");
  TMQ_REQUIRE(prec == 8 || prec == 4, "bad precision");
  TMQ_REQUIRE(nmoms > 0, "empty momentum list");
  TMQ_REQUIRE((particle == 0 || particle == 1) && (partflag == 1 || partflag == 2), "bad particle / part");
  TMQ_CUDA(cudaSetDevice(c->device));
  static OpTables ops;                                   // must outlive the asynchronous upload
  operator_tables(&ops, particle, partflag);
  cudaStream_t st = c->stream;
  TMQ_CUDA(cudaStreamSynchronize(st));                   // a previous call's upload of `ops` has completed
  TMQ_CUDA(cudaMemcpyToSymbolAsync(c_ops, &ops, sizeof(ops), 0, cudaMemcpyHostToDevice, st));
  const int X = c->g.X[0], Y = c->g.X[1], Z = c->g.X[2], T = c->g.X[3];
  const size_t V = (size_t)2 * c->g.Vh;
  TMQ_REQUIRE(X <= 32 * DFT_MAXK && Y <= 32 * DFT_MAXK && Z <= 32 * DFT_MAXK, "spatial extent above %d not supported", 32 * DFT_MAXK);
  MomProjector mp;
  mp.build(c, moms, nmoms, src_pos, +1);                 // exp(+i p x) (fixSinkContractions_local_core.h:52-56)
  const int gT = T * c->grid[3];
  const size_t ntot = (size_t)gT * nmoms * 16 * 2;
  WsPlan plan;
  const size_t o_csite = plan.add(V * 16 * sizeof(CplxT<double>));
  mp.plan(plan, 16, T);
  const size_t o_glob = plan.add(ntot * sizeof(double));
  TMQ_TRY(ensure_contract_ws(c, plan.total));
  char *ws = (char *)c->contract_ws;
  CplxT<double> *csite = (CplxT<double> *)(ws + o_csite);
  const unsigned int grid = (unsigned int)((V + CT_BLOCK - 1) / CT_BLOCK);
  if (prec == 8) fixsink_local_site_kernel<double><<<grid, CT_BLOCK, 0, st>>>(csite, (const CplxT<double> *)d_fwd_prop, (const CplxT<double> *)d_seq_prop, V);
  else fixsink_local_site_kernel<float><<<grid, CT_BLOCK, 0, st>>>(csite, (const CplxT<float> *)d_fwd_prop, (const CplxT<float> *)d_seq_prop, V);
  TMQ_CUDA(cudaGetLastError());
  c->launches++;
  for (size_t i = 0; i < ntot; i++) corr_mom[i] = 0.0;
  TMQ_TRY(mp.upload_tables(ws, st));
  TMQ_TRY(mp.project(ws, csite, 16, 0, T, corr_mom, st));
  TMQ_TRY(allreduce_host(c, ws, o_glob, corr_mom, ntot, st));
  return 0;
}

int tmq_qkxtm_fixsink_derivative(tmq_ctx *c, const void *d_seq_prop, const void *d_fwd_prop, const void *d_gauge, int prec, int particle, int partflag,
                                 const int *moms, int nmoms, const int src_pos[3], double *corr_noether, double *corr_oneD) {
  TMQ_REQUIRE(c && d_seq_prop && d_fwd_prop && d_gauge && moms && src_pos && corr_noether && corr_oneD, "null argument");
  TMQ_REQUIRE(prec == 8 || prec == 4, "bad precision");
  TMQ_REQUIRE(nmoms > 0, "empty momentum list");
  TMQ_REQUIRE((particle == 0 || particle == 1) && (partflag == 1 || partflag == 2), "bad particle / part");
  TMQ_REQUIRE(c->nranks == 1 && !c->g.part[2] && !c->g.part[3], "the derivative insertions need the neighbours' propagators: not built for a split lattice");
  TMQ_CUDA(cudaSetDevice(c->device));
  static OpTables ops;
  static DerivTables der;
  cudaStream_t st = c->stream;
  TMQ_CUDA(cudaStreamSynchronize(st));                   // a previous call's uploads of the static tables have completed
  operator_tables(&ops, particle, partflag);
  {
    Mat4 g[5];
    ukqcd_gammas(g);
    for (int d = 0; d < 4; d++)
      for (int k = 0; k < 4; k++)
        for (int l = 0; l < 4; l++) {          // operators 16+d / 20+d of gammas_tm_base.h: 1 +- g_d, not rotated
          const cd one = k == l ? 1.0 : 0.0;
          der.pg_re[d][k * 4 + l] = (one + g[d].m[k][l]).real(); der.pg_im[d][k * 4 + l] = (one + g[d].m[k][l]).imag();
          der.mg_re[d][k * 4 + l] = (one - g[d].m[k][l]).real(); der.mg_im[d][k * 4 + l] = (one - g[d].m[k][l]).imag();
        }
  }
  TMQ_CUDA(cudaMemcpyToSymbolAsync(c_ops, &ops, sizeof(ops), 0, cudaMemcpyHostToDevice, st));
  TMQ_CUDA(cudaMemcpyToSymbolAsync(c_deriv, &der, sizeof(der), 0, cudaMemcpyHostToDevice, st));
  const int X = c->g.X[0], Y = c->g.X[1], Z = c->g.X[2], T = c->g.X[3];
  const size_t V = (size_t)2 * c->g.Vh, V3 = (size_t)X * Y * Z;
  TMQ_REQUIRE(X <= 32 * DFT_MAXK && Y <= 32 * DFT_MAXK && Z <= 32 * DFT_MAXK, "spatial extent above %d not supported", 32 * DFT_MAXK);
  const int NCH = 68;                                    // 4 noether + 4 x 16 one-derivative
  int nt = (int)(((size_t)2 << 30) / (V3 * NCH * sizeof(CplxT<double>)));
  if (c->opt_contract_slices > 0 && c->opt_contract_slices < nt) nt = c->opt_contract_slices;
  if (nt < 1) nt = 1;
  if (nt > T) nt = T;
  MomProjector mp;
  mp.build(c, moms, nmoms, src_pos, +1);
  WsPlan plan;
  const size_t o_csite = plan.add((size_t)nt * V3 * NCH * sizeof(CplxT<double>));
  mp.plan(plan, NCH, nt);
  TMQ_TRY(ensure_contract_ws(c, plan.total));
  char *ws = (char *)c->contract_ws;
  CplxT<double> *csite = (CplxT<double> *)(ws + o_csite);
  TMQ_TRY(mp.upload_tables(ws, st));
  std::vector<double> corr((size_t)T * nmoms * NCH * 2, 0.0);
  for (int t0 = 0; t0 < T; t0 += nt) {
    const int n = t0 + nt <= T ? nt : T - t0;
    const size_t nsites = (size_t)n * V3, site0 = (size_t)t0 * V3;
    const dim3 grid((unsigned int)((nsites + CT_BLOCK - 1) / CT_BLOCK), 4);
    if (prec == 8) fixsink_deriv_site_kernel<double><<<grid, CT_BLOCK, 0, st>>>(csite, (const CplxT<double> *)d_fwd_prop, (const CplxT<double> *)d_seq_prop, (const CplxT<double> *)d_gauge, V, site0, nsites, X, Y, Z, T);
    else fixsink_deriv_site_kernel<float><<<grid, CT_BLOCK, 0, st>>>(csite, (const CplxT<float> *)d_fwd_prop, (const CplxT<float> *)d_seq_prop, (const CplxT<float> *)d_gauge, V, site0, nsites, X, Y, Z, T);
    TMQ_CUDA(cudaGetLastError());
    c->launches++;
    TMQ_TRY(mp.project(ws, csite, NCH, t0, n, corr.data(), st));
  }
  for (int t = 0; t < T; t++)
    for (int m = 0; m < nmoms; m++) {
      const double *src = &corr[(((size_t)t * nmoms + m) * NCH) * 2];
      for (int q = 0; q < 8; q++) corr_noether[(((size_t)t * nmoms + m) * 4) * 2 + q] = src[q];
      for (int q = 0; q < 128; q++) corr_oneD[(((size_t)t * nmoms + m) * 64) * 2 + q] = src[8 + q];
    }
  return 0;
}

}  // extern "C"
