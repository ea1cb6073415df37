// tmq_blas.cu -- fused streaming BLAS for the CG on M^dag M.  Every kernel is a
// grid-stride loop over 32-byte (fp64) / 16-byte (fp32) vectors with a persistent grid sized to the SM
// count; reductions accumulate in double, finish through block_reduce_finalize (deterministic order)
// and leave their result in the device scalar block, so the CG recurrence never waits on the host.
// Replaces upstream blas::{zero,copy,ax,axpy,axpby,xpay,caxpy,cxpaypbz,norm2,reDotProduct,cDotProduct,
// axpyNorm,xmyNorm,axpyZpbx} as used at lib/qudaQKXTM_interface.cpp:135-136 and
// lib/qudaQKXTM_Deflation.cpp:1015-1056,1431-1435.  HBM-bound: bytes/site = 24*s per stream.
#include "tmq_internal.h"

namespace tmq {

static int g_blas_grid = 0;
int blas_grid() {
  if (!g_blas_grid) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    g_blas_grid = sms * 8;   // 8 CTAs of 256 threads per SM: full occupancy, one wave
  }
  return g_blas_grid;
}
constexpr int BLAS_BLOCK = 256;

template <typename F> __device__ __forceinline__ VecT<F> vld(const VecT<F> *p) { return *p; }
template <typename F> __device__ __forceinline__ void vst(VecT<F> *p, const VecT<F> &v) { *p = v; }

template <int NRED, typename Op>
__global__ void __launch_bounds__(BLAS_BLOCK) blas_kernel(Op op, size_t n, BlasRed r) {
  double red[NRED > 0 ? NRED : 1];
#pragma unroll
  for (int j = 0; j < (NRED > 0 ? NRED : 1); j++) red[j] = 0.0;
  const size_t stride = (size_t)gridDim.x * BLAS_BLOCK;
  for (size_t i = (size_t)blockIdx.x * BLAS_BLOCK + threadIdx.x; i < n; i += stride) op(i, red);
  if (NRED > 0) block_reduce_finalize<(NRED > 0 ? NRED : 1)>(red, r.partials, r.ticket, r.scal, r.slot);
}

template <int NRED, typename Op> static cudaError_t run(Op op, size_t n, const BlasRed &r, cudaStream_t st) {
  int grid = blas_grid();
  size_t need = (n + BLAS_BLOCK - 1) / BLAS_BLOCK;
  if (need < (size_t)grid) grid = (int)(need ? need : 1);
  blas_kernel<NRED, Op><<<grid, BLAS_BLOCK, 0, st>>>(op, n, r);
  return cudaGetLastError();
}
static const BlasRed NORED = {nullptr, nullptr, nullptr, 0};

#define V4(op) \
  { o.a = op(a); o.b = op(b); o.c = op(c); o.d = op(d); }

template <typename F> struct OpAxpby {   // y = a x + b y
  F a, b; const VecT<F> *x; VecT<F> *y;
  __device__ void operator()(size_t i, double *) const {
    VecT<F> xv = x[i], yv = y[i];
    yv.a = a * xv.a + b * yv.a; yv.b = a * xv.b + b * yv.b; yv.c = a * xv.c + b * yv.c; yv.d = a * xv.d + b * yv.d;
    y[i] = yv;
  }
};
template <typename F> struct OpAx {
  F a; VecT<F> *x;
  __device__ void operator()(size_t i, double *) const {
    VecT<F> v = x[i]; v.a *= a; v.b *= a; v.c *= a; v.d *= a; x[i] = v;
  }
};
template <typename F> struct OpCaxpy {   // y += (ar + i ai) x
  F ar, ai; const VecT<F> *x; VecT<F> *y;
  __device__ void operator()(size_t i, double *) const {
    VecT<F> xv = x[i], yv = y[i];
    yv.a += ar * xv.a - ai * xv.b; yv.b += ar * xv.b + ai * xv.a;
    yv.c += ar * xv.c - ai * xv.d; yv.d += ar * xv.d + ai * xv.c;
    y[i] = yv;
  }
};
template <typename F> struct OpCxpaypbz {   // z = x + a y + b z
  F ar, ai, br, bi; const VecT<F> *x, *y; VecT<F> *z;
  __device__ void operator()(size_t i, double *) const {
    VecT<F> xv = x[i], yv = y[i], zv = z[i], o;
    o.a = xv.a + ar * yv.a - ai * yv.b + br * zv.a - bi * zv.b;
    o.b = xv.b + ar * yv.b + ai * yv.a + br * zv.b + bi * zv.a;
    o.c = xv.c + ar * yv.c - ai * yv.d + br * zv.c - bi * zv.d;
    o.d = xv.d + ar * yv.d + ai * yv.c + br * zv.d + bi * zv.c;
    z[i] = o;
  }
};
template <typename F> struct OpNorm2 {
  const VecT<F> *x;
  __device__ void operator()(size_t i, double *red) const {
    VecT<F> v = x[i];
    red[0] += (double)v.a * v.a + (double)v.b * v.b + (double)v.c * v.c + (double)v.d * v.d;
  }
};
template <typename F> struct OpRedot {
  const VecT<F> *x, *y;
  __device__ void operator()(size_t i, double *red) const {
    VecT<F> u = x[i], v = y[i];
    red[0] += (double)u.a * v.a + (double)u.b * v.b + (double)u.c * v.c + (double)u.d * v.d;
  }
};
template <typename F> struct OpCdot {   // sum conj(x) y
  const VecT<F> *x, *y;
  __device__ void operator()(size_t i, double *red) const {
    VecT<F> u = x[i], v = y[i];
    red[0] += (double)u.a * v.a + (double)u.b * v.b + (double)u.c * v.c + (double)u.d * v.d;
    red[1] += (double)u.a * v.b - (double)u.b * v.a + (double)u.c * v.d - (double)u.d * v.c;
  }
};
template <typename F> struct OpAxpyNorm {   // y += a x ; |y|^2
  F a; const VecT<F> *x; VecT<F> *y;
  __device__ void operator()(size_t i, double *red) const {
    VecT<F> xv = x[i], yv = y[i];
    yv.a += a * xv.a; yv.b += a * xv.b; yv.c += a * xv.c; yv.d += a * xv.d;
    y[i] = yv;
    red[0] += (double)yv.a * yv.a + (double)yv.b * yv.b + (double)yv.c * yv.c + (double)yv.d * yv.d;
  }
};
template <typename F> struct OpXmyNorm {   // y = x - y ; |y|^2
  const VecT<F> *x; VecT<F> *y;
  __device__ void operator()(size_t i, double *red) const {
    VecT<F> xv = x[i], yv = y[i];
    yv.a = xv.a - yv.a; yv.b = xv.b - yv.b; yv.c = xv.c - yv.c; yv.d = xv.d - yv.d;
    y[i] = yv;
    red[0] += (double)yv.a * yv.a + (double)yv.b * yv.b + (double)yv.c * yv.c + (double)yv.d * yv.d;
  }
};
template <typename F> struct OpAxpyZpbx {   // y += a x ; x = z + b x
  F a, b; VecT<F> *x, *y; const VecT<F> *z;
  __device__ void operator()(size_t i, double *) const {
    VecT<F> xv = x[i], yv = y[i], zv = z[i];
    yv.a += a * xv.a; yv.b += a * xv.b; yv.c += a * xv.c; yv.d += a * xv.d;
    xv.a = zv.a + b * xv.a; xv.b = zv.b + b * xv.b; xv.c = zv.c + b * xv.c; xv.d = zv.d + b * xv.d;
    y[i] = yv; x[i] = xv;
  }
};
template <typename F> struct OpCgUpdate {   // x += alpha p ; p = r + beta p, scalars read from the device block
  VecT<F> *x, *p; const VecT<F> *r; const double *scal; int an, ad, bn, bd, cg_iter;
  __device__ void operator()(size_t i, double *) const {
    if (cg_iteration_is_stale(scal, cg_iter)) return;
    const F alpha = (F)(scal[an] / scal[ad]), beta = (F)(scal[bn] / scal[bd]);
    VecT<F> xv = x[i], pv = p[i], rv = r[i];
    xv.a += alpha * pv.a; xv.b += alpha * pv.b; xv.c += alpha * pv.c; xv.d += alpha * pv.d;
    pv.a = rv.a + beta * pv.a; pv.b = rv.b + beta * pv.b; pv.c = rv.c + beta * pv.c; pv.d = rv.d + beta * pv.d;
    x[i] = xv; p[i] = pv;
  }
};
template <typename FD, typename FS> struct OpCopy {
  VecT<FD> *dst; const VecT<FS> *src;
  __device__ void operator()(size_t i, double *) const {
    VecT<FS> s = src[i]; VecT<FD> d; d.a = (FD)s.a; d.b = (FD)s.b; d.c = (FD)s.c; d.d = (FD)s.d; dst[i] = d;
  }
};
struct OpXpyMixed {   // y(double) += x(float)
  VecT<double> *y; const VecT<float> *x;
  __device__ void operator()(size_t i, double *) const {
    VecT<float> s = x[i]; VecT<double> d = y[i];
    d.a += s.a; d.b += s.b; d.c += s.c; d.d += s.d; y[i] = d;
  }
};
// site-local twist on a parity block vec[6][Vh]: out = c (in + i a g5 in); g5 pairs vector j with j+3
template <typename F> struct OpTwist {
  VecT<F> *out; const VecT<F> *in; F c, a; size_t Vh;
  __device__ void operator()(size_t i, double *) const {   // i < 3*Vh
    const size_t j = i / Vh, s = i - j * Vh;
    VecT<F> u = in[j * Vh + s], l = in[(j + 3) * Vh + s], ou, ol;
    ou.a = c * (u.a - a * l.b); ou.b = c * (u.b + a * l.a); ou.c = c * (u.c - a * l.d); ou.d = c * (u.d + a * l.c);
    ol.a = c * (l.a - a * u.b); ol.b = c * (l.b + a * u.a); ol.c = c * (l.c - a * u.d); ol.d = c * (l.d + a * u.c);
    out[j * Vh + s] = ou; out[(j + 3) * Vh + s] = ol;
  }
};

#define DISPATCH(prec, expr_d, expr_s) ((prec) == 8 ? (expr_d) : (expr_s))
#define VD(p) ((VecT<double> *)(p))
#define VS(p) ((VecT<float> *)(p))
#define CVD(p) ((const VecT<double> *)(p))
#define CVS(p) ((const VecT<float> *)(p))

cudaError_t blas_zero(void *x, size_t bytes, cudaStream_t st) { return cudaMemsetAsync(x, 0, bytes, st); }

cudaError_t blas_copy(void *dst, int dprec, const void *src, int sprec, size_t n, cudaStream_t st) {
  if (dprec == sprec) return cudaMemcpyAsync(dst, src, n * vec_bytes(dprec), cudaMemcpyDeviceToDevice, st);
  if (dprec == 8) return run<0>(OpCopy<double, float>{VD(dst), CVS(src)}, n, NORED, st);
  return run<0>(OpCopy<float, double>{VS(dst), CVD(src)}, n, NORED, st);
}
cudaError_t blas_axpby(int prec, double a, const void *x, double b, void *y, size_t n, cudaStream_t st) {
  return DISPATCH(prec, run<0>(OpAxpby<double>{a, b, CVD(x), VD(y)}, n, NORED, st),
                  run<0>(OpAxpby<float>{(float)a, (float)b, CVS(x), VS(y)}, n, NORED, st));
}
cudaError_t blas_ax(int prec, double a, void *x, size_t n, cudaStream_t st) {
  return DISPATCH(prec, run<0>(OpAx<double>{a, VD(x)}, n, NORED, st), run<0>(OpAx<float>{(float)a, VS(x)}, n, NORED, st));
}
cudaError_t blas_caxpy(int prec, double ar, double ai, const void *x, void *y, size_t n, cudaStream_t st) {
  return DISPATCH(prec, run<0>(OpCaxpy<double>{ar, ai, CVD(x), VD(y)}, n, NORED, st),
                  run<0>(OpCaxpy<float>{(float)ar, (float)ai, CVS(x), VS(y)}, n, NORED, st));
}
cudaError_t blas_cxpaypbz(int prec, const void *x, double ar, double ai, const void *y, double br, double bi, void *z,
                          size_t n, cudaStream_t st) {
  return DISPATCH(prec, run<0>(OpCxpaypbz<double>{ar, ai, br, bi, CVD(x), CVD(y), VD(z)}, n, NORED, st),
                  run<0>(OpCxpaypbz<float>{(float)ar, (float)ai, (float)br, (float)bi, CVS(x), CVS(y), VS(z)}, n, NORED, st));
}
cudaError_t blas_norm2(int prec, const void *x, size_t n, const BlasRed &r, cudaStream_t st) {
  return DISPATCH(prec, run<1>(OpNorm2<double>{CVD(x)}, n, r, st), run<1>(OpNorm2<float>{CVS(x)}, n, r, st));
}
cudaError_t blas_redot(int prec, const void *x, const void *y, size_t n, const BlasRed &r, cudaStream_t st) {
  return DISPATCH(prec, run<1>(OpRedot<double>{CVD(x), CVD(y)}, n, r, st), run<1>(OpRedot<float>{CVS(x), CVS(y)}, n, r, st));
}
cudaError_t blas_cdot(int prec, const void *x, const void *y, size_t n, const BlasRed &r, cudaStream_t st) {
  return DISPATCH(prec, run<2>(OpCdot<double>{CVD(x), CVD(y)}, n, r, st), run<2>(OpCdot<float>{CVS(x), CVS(y)}, n, r, st));
}
cudaError_t blas_axpy_norm(int prec, double a, const void *x, void *y, size_t n, const BlasRed &r, cudaStream_t st) {
  return DISPATCH(prec, run<1>(OpAxpyNorm<double>{a, CVD(x), VD(y)}, n, r, st),
                  run<1>(OpAxpyNorm<float>{(float)a, CVS(x), VS(y)}, n, r, st));
}
cudaError_t blas_xmy_norm(int prec, const void *x, void *y, size_t n, const BlasRed &r, cudaStream_t st) {
  return DISPATCH(prec, run<1>(OpXmyNorm<double>{CVD(x), VD(y)}, n, r, st), run<1>(OpXmyNorm<float>{CVS(x), VS(y)}, n, r, st));
}
cudaError_t blas_axpy_zpbx(int prec, double a, void *x, void *y, const void *z, double b, size_t n, cudaStream_t st) {
  return DISPATCH(prec, run<0>(OpAxpyZpbx<double>{a, b, VD(x), VD(y), CVD(z)}, n, NORED, st),
                  run<0>(OpAxpyZpbx<float>{(float)a, (float)b, VS(x), VS(y), CVS(z)}, n, NORED, st));
}
cudaError_t blas_cg_update(int prec, void *x, void *p, const void *r, size_t n, const double *scal, int an, int ad,
                           int bn, int bd, cudaStream_t st, int cg_iter) {
  return DISPATCH(prec, run<0>(OpCgUpdate<double>{VD(x), VD(p), CVD(r), scal, an, ad, bn, bd, cg_iter}, n, NORED, st),
                  run<0>(OpCgUpdate<float>{VS(x), VS(p), CVS(r), scal, an, ad, bn, bd, cg_iter}, n, NORED, st));
}
cudaError_t blas_xpy_mixed(void *y_d, const void *x_s, size_t n, cudaStream_t st) {
  return run<0>(OpXpyMixed{VD(y_d), CVS(x_s)}, n, NORED, st);
}
cudaError_t blas_twist(int prec, void *out, const void *in, double c, double a, int Vh, cudaStream_t st) {
  const size_t n = (size_t)3 * Vh;
  return DISPATCH(prec, run<0>(OpTwist<double>{VD(out), CVD(in), c, a, (size_t)Vh}, n, NORED, st),
                  run<0>(OpTwist<float>{VS(out), CVS(in), (float)c, (float)a, (size_t)Vh}, n, NORED, st));
}

}  // namespace tmq
