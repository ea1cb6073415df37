// tmq_eig.cu -- the other heavy consumer of M^dag M on this path: the Chebyshev-accelerated
// operator of QKXTM_Deflation::polynomialOperator (reference lib/qudaQKXTM_Deflation.cpp:997-1063), the eigensolver
// behind QKXTM_Deflation::eigenSolver (:1069-1475; the reference drives ARPACK's p?naupd by reverse communication
// with host-staged vectors) and the exact-deflation projector of deflateVector (:614-800; host zgemv in the reference).
//
// B200-first design, not a port:
//   * one polynomial degree = the four fused Dslash launches of M^dag M; the three-term recurrence
//     T_{n+1} = d1 (M^dag M T_n) + d2 T_n + d3 T_{n-1} is the epilogue of the fourth launch (EPI_CHEB) and the
//     buffers rotate by pointer, so the reference's ax + cxpaypbz + 2 copies per degree (10 spinor streams) vanish:
//     5376 B/site/degree (fp64, recon-12) instead of 6912;
//   * the Krylov basis stays resident in HBM (180 GB): a thick-restart Lanczos with full re-orthogonalisation
//     (classical Gram-Schmidt applied twice) built from two streaming kernels -- a block of up to 8 inner products per
//     sweep of w, and a block update w -= sum_j c_j v_j with the coefficients read from device memory -- plus an
//     in-place basis rotation V <- V Q staged through shared memory.  The host only sees the 2m coefficients;
//   * M^dag M is hermitian, so the projected matrix is real symmetric (arrowhead + tridiagonal after a restart) and is
//     diagonalised on the host by cyclic Jacobi (m <= a few hundred); ARPACK's non-hermitian Arnoldi is not needed.
// Results are defined by the operator, not by the algorithm: eigenvalues / residuals are recomputed with the true
// M^dag M exactly as the reference does after zneupd (:1426-1439).
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "../../include/tmq.h"
#include "tmq_internal.h"

namespace tmq {

// ---- Chebyshev polynomial operator ------------------------------------------------------------------------------------
// Same recurrence and constants as the reference (delta, theta, sigma1, d1..d3 of Deflation.cpp:1003-1057); T_0 = in,
// T_1 = d2 in + d1 M^dag M in with d1 = sigma1/delta, d2 = 1.
static int cheb_step(tmq_ctx *c, int prec, void *out, const void *y, const void *tm1, double d1, double d2, double d3) {
  // out = d1 M^dag M y + d2 y + d3 tm1 ; K1..K3 as in op_mdagm, K4 with the recurrence fused in
  const int p = c->matpc & 1, q = 1 - p;
  const bool asym = c->matpc >= 2;
  const double k2 = -c->kappa * c->kappa;
  void *t0 = scr(c, prec, 0), *t1 = scr(c, prec, 1);
  HopSpec k1, k2s, k3, k4;
  k1.epi = EPI_TW; k1.out_parity = q; k1.t1 = tw_Ainv(c, 0);
  TMQ_TRY(apply_hop(c, prec, t0, y, k1));
  if (!asym) {
    // w = A^-dag (y - k^2 A^-1 D t0)   (= A^-dag M y)
    k2s.epi = EPI_MDAGM2; k2s.out_parity = p; k2s.t1 = tw_Ainv(c, 0); k2s.k = k2; k2s.x = y; k2s.t3 = tw_Ainv(c, 1);
    k2s.red_slot = SC_T3;
  } else {
    // w = A y - k^2 D t0   (= M_asym y)
    k2s.epi = EPI_TWX_XPAY; k2s.out_parity = p; k2s.tx = tw_A(c, 0); k2s.k = k2; k2s.x = y;
  }
  void *ybuf = nullptr;
  if (!asym && c->clover_on) { TMQ_TRY(ensure_scratch(c, prec, 7)); ybuf = scr(c, prec, 6); k2s.out2 = ybuf; }   // see op_mdagm
  TMQ_TRY(apply_hop(c, prec, t1, t0, k2s));
  k3.epi = EPI_TW; k3.out_parity = q; k3.dagger = 1; k3.t1 = tw_Ainv(c, 1);
  TMQ_TRY(apply_hop(c, prec, t0, t1, k3));
  k4.epi = EPI_CHEB; k4.out_parity = p; k4.dagger = 1; k4.tx = tw_A(c, 1); k4.k = k2; k4.x = t1;
  if (ybuf) { k4.x = ybuf; k4.cl_plain_x = 1; }
  k4.y = y; k4.r = const_cast<void *>(tm1); k4.d1 = d1; k4.d2 = d2; k4.d3 = d3;
  return apply_hop(c, prec, out, t0, k4);
}

// the same step for the FULL (unpreconditioned) operator M_full = A - kappa D on [even | odd] fields:
//   w = M y (two launches), out_p = d1 (A^dag w_p - kappa D^dag w_q) + d2 y_p + d3 tm1_p (two launches, recurrence fused)
static int cheb_step_full(tmq_ctx *c, int prec, void *out, const void *y, const void *tm1, double d1, double d2, double d3, void *w) {
  const size_t pb = parity_bytes(c, prec);
  for (int p = 0; p < 2; p++) {
    HopSpec s; s.epi = EPI_TWX_XPAY; s.out_parity = p; s.dagger = 0; s.tx = tw_A(c, 0); s.k = -c->kappa;
    s.x = (const char *)y + (size_t)p * pb;
    TMQ_TRY(apply_hop(c, prec, (char *)w + (size_t)p * pb, (const char *)y + (size_t)(1 - p) * pb, s));
  }
  for (int p = 0; p < 2; p++) {
    HopSpec s; s.epi = EPI_CHEB; s.out_parity = p; s.dagger = 1; s.tx = tw_A(c, 1); s.k = -c->kappa;
    s.x = (const char *)w + (size_t)p * pb;
    s.y = (const char *)y + (size_t)p * pb;
    s.r = tm1 ? (char *)const_cast<void *>(tm1) + (size_t)p * pb : nullptr;
    s.d1 = d1; s.d2 = d2; s.d3 = d3;
    TMQ_TRY(apply_hop(c, prec, (char *)out + (size_t)p * pb, (const char *)w + (size_t)(1 - p) * pb, s));
  }
  return 0;
}

static int eig_full_scratch(tmq_ctx *c, int prec, void *buf[3]);

// subset = 1: p(M_pc^dag M_pc) on PARITY fields; 2: p(M_full^dag M_full) on FULL fields (the reference's isFullOp).
// deg = 0 copies; deg = -1 applies the bare operator M^dag M (no recurrence)
int op_poly(tmq_ctx *c, int prec, int subset, void *out, const void *in, int deg, double amin, double amax) {
  const size_t pb = parity_bytes(c, prec) * subset;
  if (deg == 0) { TMQ_CUDA(cudaMemcpyAsync(out, in, pb, cudaMemcpyDeviceToDevice, c->stream)); return 0; }
  TMQ_REQUIRE(out != in, "polynomial operator: out must not alias in");
  void *B[2], *w = nullptr;
  if (subset == 1) {
    TMQ_TRY(ensure_scratch(c, prec, 4));
    B[0] = scr(c, prec, 2); B[1] = scr(c, prec, 3);
  } else {
    void *f[3];
    TMQ_TRY(eig_full_scratch(c, prec, f));
    B[0] = f[0]; B[1] = f[1]; w = f[2];
  }
  auto step = [&](void *o, const void *y, const void *tm1, double d1, double d2, double d3) -> int {
    return subset == 1 ? cheb_step(c, prec, o, y, tm1, d1, d2, d3) : cheb_step_full(c, prec, o, y, tm1, d1, d2, d3, w);
  };
  if (deg < 0) return step(out, in, nullptr, 1.0, 0.0, 0.0);            // M^dag M itself
  const double delta = (amax - amin) / 2.0, theta = (amax + amin) / 2.0;
  const double sigma1 = -delta / theta;
  // degree 1
  void *dst = (deg == 1) ? out : B[0];
  TMQ_TRY(step(dst, in, nullptr, sigma1 / delta, 1.0, 0.0));
  if (deg == 1) return 0;
  const void *tm1 = in;     // T_{i-2}
  void *tm2 = B[0];         // T_{i-1}
  double sigma_old = sigma1;
  for (int i = 2; i <= deg; i++) {
    const double sigma = 1.0 / (2.0 / sigma1 - sigma_old);
    const double d1 = 2.0 * sigma / delta, d2 = -d1 * theta, d3 = -sigma * sigma_old;
    // destination: the user's buffer on the last step, otherwise the buffer holding T_{i-2} (in place) -- except at
    // i = 2, where T_0 is the caller's input and must survive
    void *o = (i == deg) ? out : (i == 2 ? B[1] : const_cast<void *>(tm1));
    TMQ_TRY(step(o, tm2, tm1, d1, d2, d3));
    tm1 = tm2; tm2 = o;
    sigma_old = sigma;
  }
  return 0;
}
int op_poly_mdagm(tmq_ctx *c, int prec, void *out, const void *in, int deg, double amin, double amax) {
  return op_poly(c, prec, 1, out, in, deg < 0 ? 0 : deg, amin, amax);
}

// ---- block inner products / block updates -------------------------------------------------------------------------------
constexpr int EIG_BLOCK = 256;
constexpr int EIG_NB = 8;          // vectors per sweep of w

template <typename F, int NB> struct VecPtrs { const VecT<F> *v[NB]; };

// out[2j], out[2j+1] = Re, Im sum_i conj(V_j[i]) w[i]
template <typename F, int NB>
__global__ void __launch_bounds__(EIG_BLOCK) multi_cdot_kernel(VecPtrs<F, NB> V, const VecT<F> *w, size_t n, double *partials,
                                                              unsigned int *ticket, double *out) {
  double red[2 * NB];
#pragma unroll
  for (int j = 0; j < 2 * NB; j++) red[j] = 0.0;
  const size_t stride = (size_t)gridDim.x * EIG_BLOCK;
  for (size_t i = (size_t)blockIdx.x * EIG_BLOCK + threadIdx.x; i < n; i += stride) {
    const VecT<F> wv = w[i];
#pragma unroll
    for (int j = 0; j < NB; j++) {
      const VecT<F> u = V.v[j][i];
      red[2 * j] += (double)u.a * wv.a + (double)u.b * wv.b + (double)u.c * wv.c + (double)u.d * wv.d;
      red[2 * j + 1] += (double)u.a * wv.b - (double)u.b * wv.a + (double)u.c * wv.d - (double)u.d * wv.c;
    }
  }
  block_reduce_finalize<2 * NB>(red, partials, ticket, out, 0);
}

// w += sign * sum_j coef_j V_j  (coef complex, device resident)
template <typename F, int NB>
__global__ void __launch_bounds__(EIG_BLOCK) multi_caxpy_kernel(VecPtrs<F, NB> V, const double *coef, double sign, VecT<F> *w, size_t n) {
  F cr[NB], ci[NB];
#pragma unroll
  for (int j = 0; j < NB; j++) { cr[j] = (F)(sign * coef[2 * j]); ci[j] = (F)(sign * coef[2 * j + 1]); }
  const size_t stride = (size_t)gridDim.x * EIG_BLOCK;
  for (size_t i = (size_t)blockIdx.x * EIG_BLOCK + threadIdx.x; i < n; i += stride) {
    VecT<F> wv = w[i];
#pragma unroll
    for (int j = 0; j < NB; j++) {
      const VecT<F> u = V.v[j][i];
      wv.a += cr[j] * u.a - ci[j] * u.b; wv.b += cr[j] * u.b + ci[j] * u.a;
      wv.c += cr[j] * u.c - ci[j] * u.d; wv.d += cr[j] * u.d + ci[j] * u.c;
    }
    w[i] = wv;
  }
}

template <typename F, int NB>
static cudaError_t launch_cdot(void *const *vp, const void *w, size_t n, double *partials, unsigned int *ticket, double *out,
                               cudaStream_t st) {
  VecPtrs<F, NB> V;
  for (int j = 0; j < NB; j++) V.v[j] = (const VecT<F> *)vp[j];
  int grid = blas_grid();
  const size_t need = (n + EIG_BLOCK - 1) / EIG_BLOCK;
  if (need < (size_t)grid) grid = (int)(need ? need : 1);
  multi_cdot_kernel<F, NB><<<grid, EIG_BLOCK, 0, st>>>(V, (const VecT<F> *)w, n, partials, ticket, out);
  return cudaGetLastError();
}
template <typename F, int NB>
static cudaError_t launch_caxpy(void *const *vp, const double *coef, double sign, void *w, size_t n, cudaStream_t st) {
  VecPtrs<F, NB> V;
  for (int j = 0; j < NB; j++) V.v[j] = (const VecT<F> *)vp[j];
  int grid = blas_grid();
  const size_t need = (n + EIG_BLOCK - 1) / EIG_BLOCK;
  if (need < (size_t)grid) grid = (int)(need ? need : 1);
  multi_caxpy_kernel<F, NB><<<grid, EIG_BLOCK, 0, st>>>(V, coef, sign, (VecT<F> *)w, n);
  return cudaGetLastError();
}
#define EIG_NB_SWITCH(nb, CALL)      \
  switch (nb) {                      \
    case 1: return CALL(1);          \
    case 2: return CALL(2);          \
    case 3: return CALL(3);          \
    case 4: return CALL(4);          \
    case 5: return CALL(5);          \
    case 6: return CALL(6);          \
    case 7: return CALL(7);          \
    case 8: return CALL(8);          \
  }                                  \
  return cudaErrorInvalidValue;

static cudaError_t cdot_block(int prec, int nb, void *const *vp, const void *w, size_t n, double *partials, unsigned int *ticket,
                              double *out, cudaStream_t st) {
  if (prec == 8) {
#define CALL(NB) launch_cdot<double, NB>(vp, w, n, partials, ticket, out, st)
    EIG_NB_SWITCH(nb, CALL)
#undef CALL
  } else {
#define CALL(NB) launch_cdot<float, NB>(vp, w, n, partials, ticket, out, st)
    EIG_NB_SWITCH(nb, CALL)
#undef CALL
  }
}
static cudaError_t caxpy_block(int prec, int nb, void *const *vp, const double *coef, double sign, void *w, size_t n, cudaStream_t st) {
  if (prec == 8) {
#define CALL(NB) launch_caxpy<double, NB>(vp, coef, sign, w, n, st)
    EIG_NB_SWITCH(nb, CALL)
#undef CALL
  } else {
#define CALL(NB) launch_caxpy<float, NB>(vp, coef, sign, w, n, st)
    EIG_NB_SWITCH(nb, CALL)
#undef CALL
  }
}

// ---- in-place basis rotation V[:, 0:k] <- V[:, 0:m] Q (Q real m x k, row-major Q[j*k + jo]) -------------------------------
// A CTA stages `rows` consecutive 2-complex vectors of all m basis fields in shared memory, then every thread forms
// outputs for its (row, jo) pairs and stores them back over fields 0..k-1.  Rows are independent, so in place is safe.
template <typename F>
__global__ void __launch_bounds__(256) rotate_kernel(VecT<F> *const *V, const double *Q, int m, int k, int rows, size_t n) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  VecT<F> *S = (VecT<F> *)smem_raw;                 // S[j * rows + r]
  const size_t i0 = (size_t)blockIdx.x * rows;
  const int nr = (int)((n - i0) < (size_t)rows ? (n - i0) : (size_t)rows);
  for (int e = threadIdx.x; e < m * rows; e += blockDim.x) {
    const int j = e / rows, r = e - j * rows;
    if (r < nr) S[e] = V[j][i0 + r];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < k * rows; e += blockDim.x) {
    const int jo = e / rows, r = e - jo * rows;
    if (r >= nr) continue;
    double a = 0, b = 0, cc = 0, d = 0;
    for (int j = 0; j < m; j++) {
      const double q = __ldg(Q + (size_t)j * k + jo);
      const VecT<F> s = S[j * rows + r];
      a += q * (double)s.a; b += q * (double)s.b; cc += q * (double)s.c; d += q * (double)s.d;
    }
    VecT<F> o; o.a = (F)a; o.b = (F)b; o.c = (F)cc; o.d = (F)d;
    V[jo][i0 + r] = o;
  }
}

// ---- deterministic start vector: uniform(-1,1) keyed by (seed, rank, element) (splitmix64) ---------------------------------
__device__ __forceinline__ double u01(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
template <typename F> __global__ void random_fill_kernel(VecT<F> *x, size_t n, unsigned long long key) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    VecT<F> v;
    v.a = (F)(2.0 * u01(key + 4 * i) - 1.0); v.b = (F)(2.0 * u01(key + 4 * i + 1) - 1.0);
    v.c = (F)(2.0 * u01(key + 4 * i + 2) - 1.0); v.d = (F)(2.0 * u01(key + 4 * i + 3) - 1.0);
    x[i] = v;
  }
}

// ---- host: cyclic Jacobi for a real symmetric matrix (n <= a few hundred) ------------------------------------------------
// A (n x n, row-major, destroyed) -> eigenvalues w[i], eigenvectors in the COLUMNS of Q (row-major Q[r*n + c]).
static void sym_eig_jacobi(int n, std::vector<double> &A, std::vector<double> &Q, std::vector<double> &w) {
  Q.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) Q[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; i++) {
      diag += A[(size_t)i * n + i] * A[(size_t)i * n + i];
      for (int j = i + 1; j < n; j++) off += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    }
    if (off <= 1e-30 * (diag + off) || off == 0.0) break;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        const double apq = A[(size_t)p * n + q];
        if (apq == 0.0) continue;
        const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
        if (fabs(apq) < 1e-300) continue;
        const double tau = (aqq - app) / (2.0 * apq);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        const double cs = 1.0 / sqrt(1.0 + t * t), sn = t * cs;
        for (int r = 0; r < n; r++) {   // columns p, q of A
          const double arp = A[(size_t)r * n + p], arq = A[(size_t)r * n + q];
          A[(size_t)r * n + p] = cs * arp - sn * arq;
          A[(size_t)r * n + q] = sn * arp + cs * arq;
        }
        for (int r = 0; r < n; r++) {   // rows p, q of A
          const double apr = A[(size_t)p * n + r], aqr = A[(size_t)q * n + r];
          A[(size_t)p * n + r] = cs * apr - sn * aqr;
          A[(size_t)q * n + r] = sn * apr + cs * aqr;
        }
        for (int r = 0; r < n; r++) {
          const double qrp = Q[(size_t)r * n + p], qrq = Q[(size_t)r * n + q];
          Q[(size_t)r * n + p] = cs * qrp - sn * qrq;
          Q[(size_t)r * n + q] = sn * qrp + cs * qrq;
        }
      }
  }
  w.resize(n);
  for (int i = 0; i < n; i++) w[i] = A[(size_t)i * n + i];
}

}  // namespace tmq

using namespace tmq;

struct tmq_eigset {
  tmq_ctx *ctx;
  int prec;
  int subset;      // TMQ_SUBSET_PARITY: even-odd operator; TMQ_SUBSET_FULL: the unpreconditioned operator (isFullOp)
  std::vector<tmq_spinor *> vec;
};

namespace tmq {

// per-context scratch of the eigensolver, grown on demand
struct EigWork {
  double *partials = nullptr;     // [blas_grid()][2*EIG_NB]
  double *coef = nullptr;         // device coefficients [2*cap]
  double *h_coef = nullptr;       // pinned mirror
  void **d_ptrs = nullptr;        // device array of basis pointers [cap]
  double *d_Q = nullptr;          // rotation matrix [cap*cap]
  int cap = 0;
  void *full[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};   // FULL-field scratch per precision (full-operator filter)
};
static std::vector<std::pair<tmq_ctx *, EigWork>> g_work;

static EigWork *eig_work(tmq_ctx *c, int cap) {
  EigWork *w = nullptr;
  for (auto &p : g_work) if (p.first == c) w = &p.second;
  if (!w) { g_work.push_back({c, EigWork()}); w = &g_work.back().second; }
  if (w->cap >= cap) return w;
  cudaStreamSynchronize(c->stream);
  if (w->coef) cudaFree(w->coef);
  if (w->h_coef) cudaFreeHost(w->h_coef);
  if (w->d_ptrs) cudaFree(w->d_ptrs);
  if (w->d_Q) cudaFree(w->d_Q);
  bool ok = true;
  if (!w->partials) ok = ok && cudaMalloc(&w->partials, (size_t)blas_grid() * 2 * EIG_NB * sizeof(double)) == cudaSuccess;
  ok = ok && cudaMalloc(&w->coef, (size_t)2 * cap * sizeof(double)) == cudaSuccess;
  ok = ok && cudaMallocHost(&w->h_coef, (size_t)2 * cap * sizeof(double)) == cudaSuccess;
  ok = ok && cudaMalloc(&w->d_ptrs, (size_t)cap * sizeof(void *)) == cudaSuccess;
  ok = ok && cudaMalloc(&w->d_Q, (size_t)cap * cap * sizeof(double)) == cudaSuccess;
  if (!ok) { set_error("eigensolver workspace allocation failed: %s", cudaGetErrorString(cudaGetLastError())); w->cap = 0; return nullptr; }
  w->cap = cap;
  return w;
}
static int eig_full_scratch(tmq_ctx *c, int prec, void *buf[3]) {
  EigWork *w = eig_work(c, 1);
  if (!w) return 1;
  const int pi = prec == 8 ? 0 : 1;
  for (int i = 0; i < 3; i++) {
    if (!w->full[pi][i]) {
      TMQ_CUDA(cudaMalloc(&w->full[pi][i], 2 * parity_bytes(c, prec)));
      TMQ_CUDA(cudaMemsetAsync(w->full[pi][i], 0, 2 * parity_bytes(c, prec), c->stream));
    }
    buf[i] = w->full[pi][i];
  }
  return 0;
}
void eig_release(tmq_ctx *c) {
  for (size_t i = 0; i < g_work.size(); i++)
    if (g_work[i].first == c) {
      EigWork &w = g_work[i].second;
      if (w.partials) cudaFree(w.partials);
      if (w.coef) cudaFree(w.coef);
      if (w.h_coef) cudaFreeHost(w.h_coef);
      if (w.d_ptrs) cudaFree(w.d_ptrs);
      if (w.d_Q) cudaFree(w.d_Q);
      for (int pi = 0; pi < 2; pi++) for (int k = 0; k < 3; k++) if (w.full[pi][k]) cudaFree(w.full[pi][k]);
      g_work.erase(g_work.begin() + i);
      return;
    }
}

// coef[0..2nv) = V^H w for the first nv vectors of `vp` (all-reduced over ranks); result left on the device and
// mirrored into h_coef (host-visible after the returned sync)
static int basis_cdot(tmq_ctx *c, EigWork *W, int prec, size_t n, void *const *vp, int nv, const void *w, bool to_host) {
  for (int j0 = 0; j0 < nv; j0 += EIG_NB) {
    const int nb = std::min(EIG_NB, nv - j0);
    TMQ_CUDA(cdot_block(prec, nb, vp + j0, w, n, W->partials, c->ticket, W->coef + 2 * j0, c->stream));
    c->launches++;
  }
  if (c->multi && c->nranks > 1) TMQ_TRY(comm_allreduce(c, W->coef, 2 * nv, c->stream));
  if (to_host) {
    TMQ_CUDA(cudaMemcpyAsync(W->h_coef, W->coef, (size_t)2 * nv * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    TMQ_CUDA(cudaStreamSynchronize(c->stream));
  }
  return 0;
}
// w += sign * sum_j coef_j v_j
static int basis_caxpy(tmq_ctx *c, EigWork *W, int prec, size_t n, void *const *vp, int nv, double sign, void *w) {
  for (int j0 = 0; j0 < nv; j0 += EIG_NB) {
    const int nb = std::min(EIG_NB, nv - j0);
    TMQ_CUDA(caxpy_block(prec, nb, vp + j0, W->coef + 2 * j0, sign, w, n, c->stream));
    c->launches++;
  }
  return 0;
}
static int basis_rotate(tmq_ctx *c, EigWork *W, int prec, size_t n, void *const *vp, int m, int k, const std::vector<double> &Qmk) {
  // Qmk: m x k row-major
  TMQ_CUDA(cudaMemcpyAsync(W->d_ptrs, vp, (size_t)m * sizeof(void *), cudaMemcpyHostToDevice, c->stream));
  TMQ_CUDA(cudaMemcpyAsync(W->d_Q, Qmk.data(), (size_t)m * k * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  const size_t vb = vec_bytes(prec);
  int rows = 32;
  while (rows > 1 && (size_t)m * rows * vb > (size_t)200 * 1024) rows >>= 1;
  const size_t smem = (size_t)m * rows * vb;
  TMQ_REQUIRE(smem <= (size_t)220 * 1024, "Krylov space too large for the in-place rotation (m = %d)", m);
  const unsigned int grid = (unsigned int)((n + rows - 1) / rows);
  if (prec == 8) {
    TMQ_CUDA(cudaFuncSetAttribute(rotate_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rotate_kernel<double><<<grid, 256, smem, c->stream>>>((VecT<double> *const *)W->d_ptrs, W->d_Q, m, k, rows, n);
  } else {
    TMQ_CUDA(cudaFuncSetAttribute(rotate_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rotate_kernel<float><<<grid, 256, smem, c->stream>>>((VecT<float> *const *)W->d_ptrs, W->d_Q, m, k, rows, n);
  }
  TMQ_CUDA(cudaGetLastError());
  c->launches++;
  // the pageable host buffers above must stay alive until the copies have been consumed
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

static int vec_norm2(tmq_ctx *c, int prec, size_t n, const void *x, double *out) {
  TMQ_CUDA(blas_norm2(prec, x, n, BlasRed{c->partials, c->ticket, c->scal, SC_T0}, c->stream)); c->launches++;
  TMQ_TRY(reduce_finish(c, SC_T0, 1));
  return fetch_scal(c, SC_T0, 1, out);
}

}  // namespace tmq

extern "C" {

int tmq_poly_mdagm(tmq_spinor *out, const tmq_spinor *in, int deg, double amin, double amax) {
  TMQ_REQUIRE(out && in, "null spinor");
  TMQ_REQUIRE(out->subset == in->subset, "out and in must both be PARITY (even-odd operator) or both FULL (full operator)");
  TMQ_REQUIRE(out->ctx == in->ctx && out->prec == in->prec, "fields must share context and precision");
  tmq_ctx *c = out->ctx;
  TMQ_REQUIRE(c->op_set, "operator not set (tmq_op_set)");
  TMQ_REQUIRE(deg >= -1, "polynomial degree must be >= 0 (or -1 for the bare M^dag M)");
  TMQ_REQUIRE(deg <= 0 || (amax > amin && amax + amin != 0.0), "Chebyshev window needs amax > amin");
  TMQ_CUDA(cudaSetDevice(c->device));
  TMQ_TRY(op_poly(c, out->prec, out->subset, out->d, in->d, deg, amin, amax));
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  return check_device_error(c);
}

tmq_eigset *tmq_eigset_alloc(tmq_ctx *c, int nvec, int prec, int subset) {
  if (!c || nvec < 1 || (prec != 8 && prec != 4) || (subset != TMQ_SUBSET_PARITY && subset != TMQ_SUBSET_FULL)) {
    set_error("tmq_eigset_alloc: bad arguments");
    return nullptr;
  }
  tmq_eigset *s = new tmq_eigset();
  s->ctx = c; s->prec = prec; s->subset = subset;
  for (int i = 0; i < nvec; i++) {
    tmq_spinor *v = tmq_spinor_alloc(c, prec, subset);
    if (!v) { for (tmq_spinor *u : s->vec) tmq_spinor_free(u); delete s; return nullptr; }
    s->vec.push_back(v);
  }
  return s;
}
int tmq_eigset_free(tmq_eigset *s) {
  if (!s) return 0;
  for (tmq_spinor *v : s->vec)
    if (s->ctx->spinors.count(v)) tmq_spinor_free(v);
  delete s;
  return 0;
}
int tmq_eigset_size(const tmq_eigset *s) { return s ? (int)s->vec.size() : 0; }
tmq_spinor *tmq_eigset_vector(tmq_eigset *s, int i) {
  if (!s || i < 0 || i >= (int)s->vec.size()) { set_error("tmq_eigset_vector: index out of range"); return nullptr; }
  return s->vec[i];
}

int tmq_eigensolve(tmq_eigset *set, int nev, int nkv, int poly_deg, double amin, double amax, double tol, int max_restarts,
                   int which, unsigned long long seed, double *evals, double *resid, int *nconv_out, int *nrestart_out,
                   int *nmatvec_out) {
  TMQ_REQUIRE(set && evals, "null argument");
  tmq_ctx *c = set->ctx;
  const int prec = set->prec, m = nkv;
  TMQ_REQUIRE(c->op_set, "operator not set (tmq_op_set)");
  TMQ_REQUIRE(nev >= 1 && m >= nev + 2, "need NkV >= NeV + 2 (NeV = %d, NkV = %d)", nev, m);
  TMQ_REQUIRE((int)set->vec.size() >= m + 1, "the eigenvector set must hold NkV + 1 = %d vectors (has %d)", m + 1, (int)set->vec.size());
  TMQ_REQUIRE(which == 0 || which == 1, "which: 0 = smallest (SR), 1 = largest (LR) eigenvalues of M^dag M");
  TMQ_REQUIRE(poly_deg >= 0 && tol > 0 && max_restarts >= 1, "bad polynomial degree / tolerance / restart count");
  TMQ_REQUIRE(poly_deg == 0 || (amax > amin && amax + amin != 0.0), "Chebyshev window needs amax > amin");
  TMQ_CUDA(cudaSetDevice(c->device));
  EigWork *W = eig_work(c, m + 1);
  if (!W) return 1;
  const int subset = set->subset;
  if (subset == TMQ_SUBSET_PARITY) TMQ_TRY(ensure_scratch(c, prec, 5));
  const size_t n = (size_t)6 * c->g.Vh * subset, pb = parity_bytes(c, prec) * subset;
  std::vector<void *> vp(m + 1);
  for (int j = 0; j <= m; j++) vp[j] = set->vec[j]->d;
  // with the Chebyshev filter the wanted end of the spectrum becomes the dominant one of p(M^dag M), as the
  // reference's SR <-> LR swap (Deflation.cpp:1093-1101); without it we iterate M^dag M itself
  const bool acc = poly_deg > 0;
  const bool largest = acc ? (which == 0) : (which == 1);
  auto apply_B = [&](void *out, const void *in) -> int {
    if (acc) return op_poly(c, prec, subset, out, in, poly_deg, amin, amax);
    return subset == TMQ_SUBSET_PARITY ? op_mdagm(c, prec, out, in, SC_T3) : op_poly(c, prec, subset, out, in, -1, 0, 0);
  };
  int nmatvec = 0;

  // start vector
  {
    const unsigned long long key = (seed * 0x9E3779B97F4A7C15ull) ^ ((unsigned long long)(c->rank + 1) << 40);
    if (prec == 8) random_fill_kernel<double><<<blas_grid(), 256, 0, c->stream>>>((VecT<double> *)vp[0], n, key);
    else random_fill_kernel<float><<<blas_grid(), 256, 0, c->stream>>>((VecT<float> *)vp[0], n, key);
    TMQ_CUDA(cudaGetLastError()); c->launches++;
    double nrm;
    TMQ_TRY(vec_norm2(c, prec, n, vp[0], &nrm));
    TMQ_CUDA(blas_ax(prec, 1.0 / sqrt(nrm), vp[0], n, c->stream)); c->launches++;
  }

  std::vector<double> theta(m, 0.0), s(m, 0.0), alpha(m, 0.0), beta(m, 0.0);
  std::vector<double> T, Q, ritz;
  std::vector<int> order(m);
  int k0 = 0, restarts = 0, nconv = 0;
  const double eps23 = pow(2.220446049250313e-16, 2.0 / 3.0);
  for (;;) {
    for (int j = k0; j < m; j++) {
      void *w = vp[j + 1];
      TMQ_TRY(apply_B(w, vp[j]));
      nmatvec++;
      // full re-orthogonalisation against v_0..v_j, classical Gram-Schmidt twice; alpha_j = Re <v_j, w>
      TMQ_TRY(basis_cdot(c, W, prec, n, vp.data(), j + 1, w, true));
      alpha[j] = W->h_coef[2 * j];
      TMQ_TRY(basis_caxpy(c, W, prec, n, vp.data(), j + 1, -1.0, w));
      TMQ_TRY(basis_cdot(c, W, prec, n, vp.data(), j + 1, w, true));
      alpha[j] += W->h_coef[2 * j];
      TMQ_TRY(basis_caxpy(c, W, prec, n, vp.data(), j + 1, -1.0, w));
      double nrm;
      TMQ_TRY(vec_norm2(c, prec, n, w, &nrm));
      beta[j] = sqrt(nrm);
      TMQ_REQUIRE(beta[j] > 0.0 && beta[j] == beta[j], "Lanczos breakdown at step %d (invariant subspace or NaN)", j);
      TMQ_CUDA(blas_ax(prec, 1.0 / beta[j], w, n, c->stream)); c->launches++;
    }
    // projected matrix: diag(theta_0..theta_{k0-1}) with couplings s_i to row k0, tridiagonal (alpha, beta) beyond
    T.assign((size_t)m * m, 0.0);
    for (int i = 0; i < k0; i++) { T[(size_t)i * m + i] = theta[i]; T[(size_t)i * m + k0] = s[i]; T[(size_t)k0 * m + i] = s[i]; }
    for (int j = k0; j < m; j++) {
      T[(size_t)j * m + j] = alpha[j];
      if (j + 1 < m) { T[(size_t)j * m + j + 1] = beta[j]; T[(size_t)(j + 1) * m + j] = beta[j]; }
    }
    sym_eig_jacobi(m, T, Q, ritz);
    for (int i = 0; i < m; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return largest ? ritz[a] > ritz[b] : ritz[a] < ritz[b]; });
    nconv = 0;
    for (int i = 0; i < nev; i++) {
      const int e = order[i];
      const double est = fabs(beta[m - 1] * Q[(size_t)(m - 1) * m + e]);
      if (est <= tol * std::max(eps23, fabs(ritz[e]))) nconv++;
    }
    restarts++;
    const bool done = nconv >= nev || restarts >= max_restarts;
    int keep = done ? nev : nev + std::min(nconv, (m - nev) / 2);
    if (!done && keep < nev + 1) keep = std::min(nev + (m - nev) / 2, m - 1);   // keep a buffer of unwanted-but-close Ritz pairs
    if (keep > m - 1) keep = m - 1;
    std::vector<double> Qmk((size_t)m * keep);
    for (int r = 0; r < m; r++)
      for (int i = 0; i < keep; i++) Qmk[(size_t)r * keep + i] = Q[(size_t)r * m + order[i]];
    TMQ_TRY(basis_rotate(c, W, prec, n, vp.data(), m, keep, Qmk));
    if (done) break;
    for (int i = 0; i < keep; i++) { theta[i] = ritz[order[i]]; s[i] = beta[m - 1] * Q[(size_t)(m - 1) * m + order[i]]; }
    TMQ_CUDA(cudaMemcpyAsync(vp[keep], vp[m], pb, cudaMemcpyDeviceToDevice, c->stream));
    k0 = keep;
  }

  // eigenvalues of the actual operator and their residuals, as the reference does after zneupd (Deflation.cpp:1426-1439)
  std::vector<double> lam(nev), res(nev);
  void *t = vp[m];
  for (int i = 0; i < nev; i++) {
    TMQ_TRY(subset == TMQ_SUBSET_PARITY ? op_mdagm(c, prec, t, vp[i], SC_T3) : op_poly(c, prec, subset, t, vp[i], -1, 0, 0));
    nmatvec++;
    void *one[1] = {vp[i]};
    TMQ_TRY(basis_cdot(c, W, prec, n, one, 1, t, true));
    lam[i] = W->h_coef[0];
    TMQ_CUDA(blas_axpby(prec, -lam[i], vp[i], 1.0, t, n, c->stream)); c->launches++;
    double nrm;
    TMQ_TRY(vec_norm2(c, prec, n, t, &nrm));
    res[i] = sqrt(nrm);
  }
  // ascending eigenvalue order; the set's handles are permuted, no data moves
  std::vector<int> perm(nev);
  for (int i = 0; i < nev; i++) perm[i] = i;
  std::sort(perm.begin(), perm.end(), [&](int a, int b) { return lam[a] < lam[b]; });
  std::vector<tmq_spinor *> sorted(nev);
  for (int i = 0; i < nev; i++) { sorted[i] = set->vec[perm[i]]; evals[i] = lam[perm[i]]; if (resid) resid[i] = res[perm[i]]; }
  for (int i = 0; i < nev; i++) set->vec[i] = sorted[i];
  if (nconv_out) *nconv_out = nconv;
  if (nrestart_out) *nrestart_out = restarts;
  if (nmatvec_out) *nmatvec_out = nmatvec;
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  return check_device_error(c);
}

int tmq_deflate(tmq_spinor *out, const tmq_spinor *in, tmq_eigset *set, const double *evals, int nvec) {
  TMQ_REQUIRE(out && in && set && evals, "null argument");
  tmq_ctx *c = set->ctx;
  TMQ_REQUIRE(out->ctx == c && in->ctx == c, "fields belong to another context");
  TMQ_REQUIRE(out->subset == set->subset && in->subset == set->subset, "fields must have the site subset of the eigenvector set");
  TMQ_REQUIRE(out->prec == set->prec && in->prec == set->prec, "fields must have the precision of the eigenvector set");
  TMQ_REQUIRE(nvec >= 0 && nvec <= (int)set->vec.size(), "bad number of eigenvectors");
  TMQ_REQUIRE(out->d != in->d, "out must not alias in");
  TMQ_CUDA(cudaSetDevice(c->device));
  const size_t pb = parity_bytes(c, set->prec) * set->subset, n = (size_t)6 * c->g.Vh * set->subset;
  TMQ_CUDA(cudaMemsetAsync(out->d, 0, pb, c->stream));
  if (nvec == 0) { TMQ_CUDA(cudaStreamSynchronize(c->stream)); return 0; }
  EigWork *W = eig_work(c, nvec);
  if (!W) return 1;
  std::vector<void *> vp(nvec);
  for (int j = 0; j < nvec; j++) vp[j] = set->vec[j]->d;
  // U^dag in -> Lambda^-1 -> U (.)   (reference: zgemv ConjTrans, MPI_Allreduce, divide, zgemv NoTrans)
  TMQ_TRY(basis_cdot(c, W, set->prec, n, vp.data(), nvec, in->d, true));
  for (int j = 0; j < nvec; j++) {
    TMQ_REQUIRE(evals[j] != 0.0, "zero eigenvalue %d in deflation", j);
    W->h_coef[2 * j] /= evals[j]; W->h_coef[2 * j + 1] /= evals[j];
  }
  TMQ_CUDA(cudaMemcpyAsync(W->coef, W->h_coef, (size_t)2 * nvec * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  TMQ_TRY(basis_caxpy(c, W, set->prec, n, vp.data(), nvec, 1.0, out->d));
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int tmq_project(tmq_spinor *out, const tmq_spinor *in, tmq_eigset *set, int nvec) {
  TMQ_REQUIRE(out && in && set, "null argument");
  tmq_ctx *c = set->ctx;
  TMQ_REQUIRE(out->ctx == c && in->ctx == c, "fields belong to another context");
  TMQ_REQUIRE(out->subset == set->subset && in->subset == set->subset, "fields must have the site subset of the eigenvector set");
  TMQ_REQUIRE(out->prec == set->prec && in->prec == set->prec, "fields must have the precision of the eigenvector set");
  TMQ_REQUIRE(nvec >= 0 && nvec <= (int)set->vec.size(), "bad number of eigenvectors");
  TMQ_CUDA(cudaSetDevice(c->device));
  const size_t pb = parity_bytes(c, set->prec) * set->subset, n = (size_t)6 * c->g.Vh * set->subset;
  if (out->d != in->d) TMQ_CUDA(cudaMemcpyAsync(out->d, in->d, pb, cudaMemcpyDeviceToDevice, c->stream));
  if (nvec > 0) {
    EigWork *W = eig_work(c, nvec);
    if (!W) return 1;
    std::vector<void *> vp(nvec);
    for (int j = 0; j < nvec; j++) vp[j] = set->vec[j]->d;
    // vec_in - U (U^dag vec_in): zgemv ConjTrans, all-reduce, zgemv NoTrans, zaxpy in the reference; the coefficients
    // never leave the device here
    TMQ_TRY(basis_cdot(c, W, set->prec, n, vp.data(), nvec, in->d, false));
    TMQ_TRY(basis_caxpy(c, W, set->prec, n, vp.data(), nvec, -1.0, out->d));
  }
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

}  // extern "C"
