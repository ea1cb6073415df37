// tmq_smear.cu -- Gaussian (Wuppertal) smearing of a QKXTM_Vector on the plug-in's own device layouts
// (SURVEY.md 8f row 2).  Replaces QKXTM_Vector::gaussianSmearing (reference lib/qudaQKXTM_Vector.cpp:386-421) and
// its kernel body lib/code_pieces/Gauss_core.h:
//     out(x) = ( psi(x) + alpha sum_{mu = x,y,z} [ U_mu(x) psi(x+mu) + U_mu(x-mu)^dag psi(x-mu) ] ) / (1 + 6 alpha)
// on   vector d[(s*3+c)*V + x]            (complex, x lexicographic)           lib/qudaQKXTM_Vector.cpp:72-81
//      gauge  d[((dir*3+c1)*3+c2)*V + x]  (the APE-smeared links)              lib/qudaQKXTM_Gauge.cpp:73-89
// The reference runs nsmear (= 50) texture-fetch launches, each followed by a synchronous host-staged ghost exchange
// of the whole vector.  Here:
//   * the time direction does not hop, so time slices are independent: a T-sharded lattice needs NO exchange at all
//     (z-sharding is refused for this entry point), and
//   * optionally (TMQ_OPT_SMEAR_BLOCK_T) the sweep order is (block of time slices) outer, (smearing step) inner, so that
//     the ping-pong vectors and the three spatial link directions of a block (90 MB for one 48^3 slice in fp64) stay
//     resident in the 126 MB L2 across the nsmear steps.  MEASURED on B200 at 48^3x96 (profiles/r03_smear_bench.jsonl):
//     the blocked order is SLOWER (1.95 ms vs 1.67 ms per step in fp64): one slice is only 864 CTAs (1.2 waves) and the
//     4800 short launches expose launch gaps and tails that outweigh the L2 hit rate.  The default is therefore the
//     plain streaming order (0.79 of the measured HBM peak); a persistent per-block kernel with grid syncs is the
//     way to make the L2 residency pay and is left for a later round.
// HBM-bound streaming stencil: algorithmic bytes per site and step = (24 in + 24 out + 3*18 links) * sizeof(real)
// = 816 B in fp64; no tensor cores.
#include <cuda_runtime.h>
#include "../../include/tmq.h"
#include "tmq_internal.h"

namespace tmq {

constexpr int SMEAR_BLOCK = 128;

struct SmearGeom {
  int X[4];
  FastDiv dX, dY, dZ;
  size_t V;       // local 4-volume = component stride
  int t0, nt;     // time slices [t0, t0 + nt) handled by this launch
};

// acc[s][a] += sum_b U[a][b] psi[s][b]   (apply_U_on_S, lib/code_pieces/core_def.h:498-513)
// acc[s][a] += sum_b conj(U[b][a]) psi[s][b]   (apply_U_DAG_on_S, :515-530)
template <typename F, bool DAG>
__device__ __forceinline__ void su3_acc(F (&acc)[12][2], const CplxT<F> *__restrict__ gauge, size_t glink, const CplxT<F> *__restrict__ in,
                                        size_t nsite, size_t V) {
  F u[9][2];
#pragma unroll
  for (int k = 0; k < 9; k++) { const CplxT<F> g = gauge[glink + (size_t)k * V]; u[k][0] = g.re; u[k][1] = g.im; }
#pragma unroll
  for (int s = 0; s < 4; s++) {
    F p[3][2];
#pragma unroll
    for (int b = 0; b < 3; b++) { const CplxT<F> v = in[(size_t)(s * 3 + b) * V + nsite]; p[b][0] = v.re; p[b][1] = v.im; }
#pragma unroll
    for (int a = 0; a < 3; a++) {
      F re = acc[s * 3 + a][0], im = acc[s * 3 + a][1];
#pragma unroll
      for (int b = 0; b < 3; b++) {
        if (!DAG) {
          const F ur = u[a * 3 + b][0], ui = u[a * 3 + b][1];
          re += ur * p[b][0]; re -= ui * p[b][1]; im += ur * p[b][1]; im += ui * p[b][0];
        } else {
          const F ur = u[b * 3 + a][0], ui = u[b * 3 + a][1];
          re += ur * p[b][0]; re += ui * p[b][1]; im += ur * p[b][1]; im -= ui * p[b][0];
        }
      }
      acc[s * 3 + a][0] = re; acc[s * 3 + a][1] = im;
    }
  }
}

template <typename F>
__global__ void __launch_bounds__(SMEAR_BLOCK, (sizeof(F) == 8 ? 4 : 8)) gauss_smear_kernel(CplxT<F> *__restrict__ out, const CplxT<F> *__restrict__ in,
                                                                 const CplxT<F> *__restrict__ gauge, SmearGeom g, F alpha, F normalize) {
  const uint32_t e = blockIdx.x * SMEAR_BLOCK + threadIdx.x;
  const uint32_t nsl = (uint32_t)(g.X[0] * g.X[1] * g.X[2]);
  if (e >= nsl * (uint32_t)g.nt) return;
  // decode (x, y, z, t): x fastest, as the reference's sid (Gauss_core.h:5-13)
  uint32_t r = fd_div(e, g.dX); const int x = (int)(e - r * g.dX.d);
  uint32_t q = fd_div(r, g.dY); const int y = (int)(r - q * g.dY.d); r = q;
  q = fd_div(r, g.dZ); const int z = (int)(r - q * g.dZ.d);
  const int t = g.t0 + (int)q;
  const size_t V = g.V;
  const size_t sx = 1, sy = (size_t)g.X[0], sz = (size_t)g.X[0] * g.X[1];
  const size_t sid = (((size_t)t * g.X[2] + z) * g.X[1] + y) * g.X[0] + x;
  const size_t xp = x == g.X[0] - 1 ? sid - (size_t)(g.X[0] - 1) * sx : sid + sx, xm = x == 0 ? sid + (size_t)(g.X[0] - 1) * sx : sid - sx;
  const size_t yp = y == g.X[1] - 1 ? sid - (size_t)(g.X[1] - 1) * sy : sid + sy, ym = y == 0 ? sid + (size_t)(g.X[1] - 1) * sy : sid - sy;
  const size_t zp = z == g.X[2] - 1 ? sid - (size_t)(g.X[2] - 1) * sz : sid + sz, zm = z == 0 ? sid + (size_t)(g.X[2] - 1) * sz : sid - sz;

  F acc[12][2];
#pragma unroll
  for (int k = 0; k < 12; k++) { acc[k][0] = 0; acc[k][1] = 0; }
  su3_acc<F, false>(acc, gauge, (size_t)0 * 9 * V + sid, in, xp, V);
  su3_acc<F, true>(acc, gauge, (size_t)0 * 9 * V + xm, in, xm, V);
  su3_acc<F, false>(acc, gauge, (size_t)1 * 9 * V + sid, in, yp, V);
  su3_acc<F, true>(acc, gauge, (size_t)1 * 9 * V + ym, in, ym, V);
  su3_acc<F, false>(acc, gauge, (size_t)2 * 9 * V + sid, in, zp, V);
  su3_acc<F, true>(acc, gauge, (size_t)2 * 9 * V + zm, in, zm, V);
#pragma unroll
  for (int k = 0; k < 12; k++) {
    const CplxT<F> s = in[(size_t)k * V + sid];
    CplxT<F> o;
    o.re = normalize * (s.re + alpha * acc[k][0]);      // Gauss_core.h:199-215
    o.im = normalize * (s.im + alpha * acc[k][1]);
    out[(size_t)k * V + sid] = o;
  }
}

template <typename F>
static cudaError_t smear_launch(void *out, const void *in, const void *gauge, const SmearGeom &g, double alpha, cudaStream_t st) {
  const size_t n = (size_t)g.X[0] * g.X[1] * g.X[2] * g.nt;
  const unsigned int grid = (unsigned int)((n + SMEAR_BLOCK - 1) / SMEAR_BLOCK);
  gauss_smear_kernel<F><<<grid, SMEAR_BLOCK, 0, st>>>((CplxT<F> *)out, (const CplxT<F> *)in, (const CplxT<F> *)gauge, g, (F)alpha,
                                                     (F)(1.0 / (1.0 + 6.0 * alpha)));
  return cudaGetLastError();
}

// nsmear steps, ping-ponging between `a` (holds the input) and `b`; returns which buffer holds the result
int smear_run(tmq_ctx *c, void *a, void *b, const void *gauge, int prec, int nsmear, double alpha, int block_t, void **result) {
  SmearGeom g;
  for (int d = 0; d < 4; d++) g.X[d] = c->g.X[d];
  g.dX = make_fastdiv((uint32_t)g.X[0]); g.dY = make_fastdiv((uint32_t)g.X[1]); g.dZ = make_fastdiv((uint32_t)g.X[2]);
  g.V = (size_t)2 * c->g.Vh;
  const int T = g.X[3];
  if (block_t <= 0) block_t = T;        // default: plain streaming order (measured faster, see the header comment)
  if (block_t > T) block_t = T;
  for (int t0 = 0; t0 < T; t0 += block_t) {
    g.t0 = t0; g.nt = (t0 + block_t <= T) ? block_t : T - t0;
    void *src = a, *dst = b;
    for (int i = 0; i < nsmear; i++) {
      TMQ_CUDA(prec == 8 ? smear_launch<double>(dst, src, gauge, g, alpha, c->stream) : smear_launch<float>(dst, src, gauge, g, alpha, c->stream));
      c->launches++;
      void *tmp = src; src = dst; dst = tmp;
    }
  }
  *result = (nsmear & 1) ? b : a;
  return 0;
}

}  // namespace tmq

using namespace tmq;

extern "C" {

int tmq_qkxtm_gauss_smear(tmq_ctx *c, void *d_out, void *d_in, const void *d_gauge, int prec, int nsmear, double alpha) {
  TMQ_REQUIRE(c && d_out && d_in && d_gauge, "null argument");
  TMQ_REQUIRE(prec == 8 || prec == 4, "bad precision");
  TMQ_REQUIRE(nsmear >= 0, "nsmear must be >= 0");
  TMQ_REQUIRE(d_out != d_in, "out must not alias in");
  TMQ_REQUIRE(c->grid[0] == 1 && c->grid[1] == 1 && c->grid[2] == 1,
              "Gaussian smearing is 3-dimensional: it needs no exchange on a T-sharded lattice, a z split is not supported");
  TMQ_CUDA(cudaSetDevice(c->device));
  void *res = nullptr;
  // the reference ping-pongs between `this` (out) and vecIn, first step vecIn -> this (lib/qudaQKXTM_Vector.cpp:403-417)
  TMQ_TRY(smear_run(c, d_in, d_out, d_gauge, prec, nsmear, alpha, c->opt_smear_block_t, &res));
  if (res != d_out) TMQ_CUDA(cudaMemcpyAsync(d_out, d_in, (size_t)2 * c->g.Vh * 24 * prec, cudaMemcpyDeviceToDevice, c->stream));   // :419
  return 0;
}

int tmq_timer_start(tmq_ctx *c) {
  TMQ_REQUIRE(c, "null context");
  TMQ_CUDA(cudaEventRecord(c->ev_a, c->stream));
  return 0;
}
int tmq_timer_stop(tmq_ctx *c, double *ms) {
  TMQ_REQUIRE(c && ms, "null argument");
  TMQ_CUDA(cudaEventRecord(c->ev_b, c->stream));
  TMQ_CUDA(cudaEventSynchronize(c->ev_b));
  float f = 0;
  TMQ_CUDA(cudaEventElapsedTime(&f, c->ev_a, c->ev_b));
  *ms = f;
  return 0;
}

}  // extern "C"
