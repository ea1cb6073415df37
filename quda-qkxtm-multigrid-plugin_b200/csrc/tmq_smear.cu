// tmq_smear.cu -- Gaussian (Wuppertal) smearing of a QKXTM_Vector on the plug-in's own device layouts
// (SURVEY.md 8f row 2).  Replaces QKXTM_Vector::gaussianSmearing (reference lib/qudaQKXTM_Vector.cpp:386-421) and
// its kernel body lib/code_pieces/Gauss_core.h:
//     out(x) = ( psi(x) + alpha sum_{mu = x,y,z} [ U_mu(x) psi(x+mu) + U_mu(x-mu)^dag psi(x-mu) ] ) / (1 + 6 alpha)
// on   vector d[(s*3+c)*V + x]            (complex, x lexicographic)           lib/qudaQKXTM_Vector.cpp:72-81
//      gauge  d[((dir*3+c1)*3+c2)*V + x]  (the APE-smeared links)              lib/qudaQKXTM_Gauge.cpp:73-89
// The reference runs nsmear (= 50) texture-fetch launches, each followed by a synchronous host-staged ghost exchange
// of the whole vector.  Here:
//   * the time direction does not hop, so time slices are independent: a T-sharded lattice needs NO exchange at all;
//     on a z split the two z faces of the vector (24 reals per face site, as the reference's ghost zone,
//     lib/qudaQKXTM_Vector.cpp:334) are packed on the device and exchanged with the z neighbours before every step
//     (NCCL send/recv; a device copy when the dimension wraps onto this rank), and the U_z links of the neighbour's
//     last z slice are fetched once per call, so the arithmetic -- and hence every bit of the result -- is that of
//     the unsharded sweep, and
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
// The link is read at gauge[glink + k*GV], the neighbour spinor at in[nsite + k*SV]: (GV, SV) = (V, V) in the bulk, the face
// size when the operand comes from a ghost buffer.
template <typename F, bool DAG>
__device__ __forceinline__ void su3_acc(F (&acc)[12][2], const CplxT<F> *__restrict__ gauge, size_t glink, size_t GV,
                                        const CplxT<F> *__restrict__ in, size_t nsite, size_t SV) {
  F u[9][2];
#pragma unroll
  for (int k = 0; k < 9; k++) { const CplxT<F> g = gauge[glink + (size_t)k * GV]; u[k][0] = g.re; u[k][1] = g.im; }
#pragma unroll
  for (int s = 0; s < 4; s++) {
    F p[3][2];
#pragma unroll
    for (int b = 0; b < 3; b++) { const CplxT<F> v = in[(size_t)(s * 3 + b) * SV + nsite]; p[b][0] = v.re; p[b][1] = v.im; }
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

// ghost operands of a z-partitioned sweep, all indexed by the face site f = (t*Y + y)*X + x with component stride F = X*Y*T:
//   psi_fwd: the forward neighbour's z = 0 slice, psi_bwd / uz_bwd: the backward neighbour's z = L-1 slice and its U_z links
template <typename F> struct SmearGhost {
  const CplxT<F> *psi_fwd, *psi_bwd, *uz_bwd;
  size_t F_;
};

template <typename F, bool ZGHOST>
__global__ void __launch_bounds__(SMEAR_BLOCK, (sizeof(F) == 8 ? 4 : 8)) gauss_smear_kernel(CplxT<F> *__restrict__ out, const CplxT<F> *__restrict__ in,
                                                                 const CplxT<F> *__restrict__ gauge, SmearGeom g, F alpha, F normalize,
                                                                 SmearGhost<F> gh) {
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
  su3_acc<F, false>(acc, gauge, (size_t)0 * 9 * V + sid, V, in, xp, V);
  su3_acc<F, true>(acc, gauge, (size_t)0 * 9 * V + xm, V, in, xm, V);
  su3_acc<F, false>(acc, gauge, (size_t)1 * 9 * V + sid, V, in, yp, V);
  su3_acc<F, true>(acc, gauge, (size_t)1 * 9 * V + ym, V, in, ym, V);
  if (ZGHOST) {
    // same operands in the same order as the unsharded sweep, only fetched from the ghost buffers on the two z faces
    const size_t f = ((size_t)t * g.X[1] + y) * g.X[0] + x;
    if (z == g.X[2] - 1) su3_acc<F, false>(acc, gauge, (size_t)2 * 9 * V + sid, V, gh.psi_fwd, f, gh.F_);
    else su3_acc<F, false>(acc, gauge, (size_t)2 * 9 * V + sid, V, in, zp, V);
    if (z == 0) su3_acc<F, true>(acc, gh.uz_bwd, f, gh.F_, gh.psi_bwd, f, gh.F_);
    else su3_acc<F, true>(acc, gauge, (size_t)2 * 9 * V + zm, V, in, zm, V);
  } else {
    su3_acc<F, false>(acc, gauge, (size_t)2 * 9 * V + sid, V, in, zp, V);
    su3_acc<F, true>(acc, gauge, (size_t)2 * 9 * V + zm, V, in, zm, V);
  }
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
static cudaError_t smear_launch(void *out, const void *in, const void *gauge, const SmearGeom &g, double alpha, const SmearGhost<F> *gh,
                                cudaStream_t st) {
  const size_t n = (size_t)g.X[0] * g.X[1] * g.X[2] * g.nt;
  const unsigned int grid = (unsigned int)((n + SMEAR_BLOCK - 1) / SMEAR_BLOCK);
  if (gh)
    gauss_smear_kernel<F, true><<<grid, SMEAR_BLOCK, 0, st>>>((CplxT<F> *)out, (const CplxT<F> *)in, (const CplxT<F> *)gauge, g, (F)alpha,
                                                             (F)(1.0 / (1.0 + 6.0 * alpha)), *gh);
  else
    gauss_smear_kernel<F, false><<<grid, SMEAR_BLOCK, 0, st>>>((CplxT<F> *)out, (const CplxT<F> *)in, (const CplxT<F> *)gauge, g, (F)alpha,
                                                              (F)(1.0 / (1.0 + 6.0 * alpha)), SmearGhost<F>{nullptr, nullptr, nullptr, 0});
  return cudaGetLastError();
}

// gather ncomp components of the z slices 0 and L-1 of a QKXTM SoA field (component stride V) into contiguous face
// buffers lo / hi [ncomp][T*Y*X]; either destination may be null.  One thread per (component, face site).
template <typename F>
__global__ void __launch_bounds__(SMEAR_BLOCK) zface_pack_kernel(CplxT<F> *__restrict__ lo, CplxT<F> *__restrict__ hi, const CplxT<F> *__restrict__ src,
                                                                int ncomp, int X, int Y, int Z, int T, size_t V) {
  const size_t nface = (size_t)X * Y * T;
  const size_t e = (size_t)blockIdx.x * SMEAR_BLOCK + threadIdx.x;
  if (e >= nface * ncomp) return;
  const size_t k = e / nface, f = e - k * nface;
  const size_t xy = f % ((size_t)X * Y), t = f / ((size_t)X * Y);
  const size_t s0 = t * Z * X * Y + xy;                            // z = 0
  if (lo) lo[e] = src[k * V + s0];
  if (hi) hi[e] = src[k * V + s0 + (size_t)(Z - 1) * X * Y];       // z = L-1
}
template <typename F>
static cudaError_t zface_pack(void *lo, void *hi, const void *src, int ncomp, const SmearGeom &g, cudaStream_t st) {
  const size_t n = (size_t)g.X[0] * g.X[1] * g.X[3] * ncomp;
  zface_pack_kernel<F><<<(unsigned int)((n + SMEAR_BLOCK - 1) / SMEAR_BLOCK), SMEAR_BLOCK, 0, st>>>((CplxT<F> *)lo, (CplxT<F> *)hi, (const CplxT<F> *)src,
                                                                                                  ncomp, g.X[0], g.X[1], g.X[2], g.X[3], g.V);
  return cudaGetLastError();
}

// nsmear steps, ping-ponging between `a` (holds the input) and `b`; returns which buffer holds the result
int smear_run(tmq_ctx *c, void *a, void *b, const void *gauge, int prec, int nsmear, double alpha, int block_t, void **result) {
  SmearGeom g;
  for (int d = 0; d < 4; d++) g.X[d] = c->g.X[d];
  g.dX = make_fastdiv((uint32_t)g.X[0]); g.dY = make_fastdiv((uint32_t)g.X[1]); g.dZ = make_fastdiv((uint32_t)g.X[2]);
  g.V = (size_t)2 * c->g.Vh;
  const int T = g.X[3];
  const bool zghost = c->g.part[2] != 0;   // z split across ranks, or forced onto the ghost path (tmq_force_partition)
  if (block_t <= 0 || zghost) block_t = T;  // default: plain streaming order (measured faster, see the header comment)
  if (block_t > T) block_t = T;

  // ghost-zone work space of a z-partitioned sweep: [send lo | send hi | recv from fwd | recv from bwd] spinor faces and the
  // backward neighbour's U_z face (sent once: the links do not change between steps)
  char *ws = nullptr;
  const size_t nface = (size_t)g.X[0] * g.X[1] * T, cb = (size_t)2 * prec;
  const size_t sp_face = 12 * nface * cb, u_face = 9 * nface * cb;
  SmearGhost<double> ghd = {nullptr, nullptr, nullptr, nface};
  SmearGhost<float> ghs = {nullptr, nullptr, nullptr, nface};
  if (zghost && nsmear > 0) {
    TMQ_CUDA(cudaMalloc((void **)&ws, 4 * sp_face + 2 * u_face));
    char *u_send = ws + 4 * sp_face, *u_recv = u_send + u_face;
    const char *uz = (const char *)gauge + (size_t)2 * 9 * g.V * cb;
    TMQ_CUDA(prec == 8 ? zface_pack<double>(nullptr, u_send, uz, 9, g, c->stream) : zface_pack<float>(nullptr, u_send, uz, 9, g, c->stream));
    c->launches++;
    // only the forward-going face carries data; the backward-going message re-sends it and lands in the spinor area, unused
    int rc = comm_sendrecv_dim(c, 2, u_send, u_send, ws, u_recv, u_face, c->stream);
    if (rc) { cudaFree(ws); return rc; }
    ghd.psi_fwd = (const CplxT<double> *)(ws + 2 * sp_face); ghd.psi_bwd = (const CplxT<double> *)(ws + 3 * sp_face); ghd.uz_bwd = (const CplxT<double> *)u_recv;
    ghs.psi_fwd = (const CplxT<float> *)(ws + 2 * sp_face); ghs.psi_bwd = (const CplxT<float> *)(ws + 3 * sp_face); ghs.uz_bwd = (const CplxT<float> *)u_recv;
  }
  int rc = 0;
  for (int t0 = 0; t0 < T && !rc; t0 += block_t) {
    g.t0 = t0; g.nt = (t0 + block_t <= T) ? block_t : T - t0;
    void *src = a, *dst = b;
    for (int i = 0; i < nsmear && !rc; i++) {
      cudaError_t e = cudaSuccess;
      if (zghost) {
        e = prec == 8 ? zface_pack<double>(ws, ws + sp_face, src, 12, g, c->stream) : zface_pack<float>(ws, ws + sp_face, src, 12, g, c->stream);
        c->launches++;
        if (e == cudaSuccess) rc = comm_sendrecv_dim(c, 2, ws, ws + sp_face, ws + 2 * sp_face, ws + 3 * sp_face, sp_face, c->stream);
      }
      if (e == cudaSuccess && !rc)
        e = prec == 8 ? smear_launch<double>(dst, src, gauge, g, alpha, zghost ? &ghd : nullptr, c->stream)
                      : smear_launch<float>(dst, src, gauge, g, alpha, zghost ? &ghs : nullptr, c->stream);
      if (e != cudaSuccess) { set_error("%s:%d CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e)); rc = 1; }
      c->launches++;
      void *tmp = src; src = dst; dst = tmp;
    }
  }
  if (ws) { cudaStreamSynchronize(c->stream); cudaFree(ws); }
  if (rc) return rc;
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
  TMQ_REQUIRE(c->grid[0] == 1 && c->grid[1] == 1, "only z and t may be partitioned");
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
