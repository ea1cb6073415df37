// qkxtm_kernels_host.cpp -- TEST INFRASTRUCTURE: compiles the REFERENCE'S OWN kernel bodies for the CPU.
//
// The plug-in's device kernels are textual includes of lib/code_pieces/*_core.h into __global__ wrappers
// (reference lib/qudaQKXTM_kernels.cu:450-472,1024-1124).  The bodies are plain C over a thread index, so this file
// supplies a minimal host stand-in for what the wrappers provide (blockIdx/blockDim/threadIdx, the __constant__
// geometry c_*, double2/float2 with the operators of lib/qudaQKXTM_kernels.cu:333-398, texture fetches as array loads)
// and #includes the bodies FROM WHERE THEY LIE under /root/reference -- nothing is copied into this repository.
// Built by oracle/Makefile into oracle/_ref/libqkxtm_ref.so (git-ignored; travels to the GPU box prebuilt).
// It gives the parity tests the reference's own arithmetic for:
//   Gauss_core.h ................ QKXTM_Vector::gaussianSmearing           
//   uploadToCuda_core.h ......... QKXTM_Vector::uploadToCuda               (8a row a11)
//   downloadFromCuda_core.h ..... QKXTM_Vector::downloadFromCuda           (8a row a12)
//   scaleVector_core.h .......... QKXTM_Vector::scaleVector                (8a row a12)
//   apply_gamma5_vector_core.h .. QKXTM_Vector::apply_gamma5               (8a row a5: gamma5 in UKQCD)
//   castDoubleToFloat_core.h, castFloatToDouble_core.h
//   plaquette_core.h ............ QKXTM_Gauge::calculatePlaq               (8a row a13): its block reduction uses __shared__ and
//                                 __syncthreads, emulated with one OS thread per CUDA thread of a block and a pthread barrier
//   conjugate_vector_core.h, conjugate_propagator_core.h, apply_gamma5_propagator_core.h, rotateToPhysicalBase_core.h
//                                 QKXTM_Vector::conjugate, QKXTM_Propagator::{conjugate, apply_gamma5, rotateToPhysicalBase_device}
//   contractMesons_core.h, contractMesons_core_PosSpace.h
//                                 QKXTM_Contraction::contractMesons (the two-point step after the solves,
//                                 lib/qudaQKXTM_interface.cpp:1217-1223), with the reference's own channel tables
//                                 (lib/qudaQKXTM_kernels.cu:77-78, extracted by the Makefile into _ref/qkxtm_meson_tables.h)
//   seqSourceFixSinkPart1_core.h, seqSourceFixSinkPart2_core.h (+ projectors_tm_base.h), fixSinkContractions_local_core.h
//   (+ gammas_tm_base.h) ........ QKXTM_Contraction::{seqSourceFixSinkPart1, seqSourceFixSinkPart2, contractFixSink (local part)}: the
//                                 fixed-sink three-point function of calcMG_threepTwop_EvenOdd (lib/qudaQKXTM_interface.cpp:838-1170)
//   contractBaryons_core.h ...... QKXTM_Contraction::contractBaryons (lib/qudaQKXTM_interface.cpp:1220), ten channels x 4x4 spin,
//                                 with the reference's tables (lib/qudaQKXTM_kernels.cu:79-88 -> _ref/qkxtm_baryon_tables.h)
#include <cstddef>
#include <cstring>
#include <cmath>
#include <functional>
#include <thread>
#include <pthread.h>
#include <vector>

struct double2 { double x, y; };
struct float2 { float x, y; };
struct dim3_ { int x; };

// operators of lib/qudaQKXTM_kernels.cu:333-398 (templates there; same arithmetic)
template <typename T2> static inline T2 cmul(const T2 a, const T2 b) { T2 r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
static inline double2 operator*(const double2 a, const double2 b) { return cmul(a, b); }
static inline float2 operator*(const float2 a, const float2 b) { return cmul(a, b); }
static inline double2 operator*(const double a, const double2 b) { double2 r; r.x = a * b.x; r.y = a * b.y; return r; }
static inline float2 operator*(const float a, const float2 b) { float2 r; r.x = a * b.x; r.y = a * b.y; return r; }
static inline float2 operator*(const double a, const float2 b) { float2 r; r.x = a * b.x; r.y = a * b.y; return r; }   // double constant x float2, as nvcc promotes
static inline float2 operator*(const int a, const float2 b) { float2 r; r.x = a * b.x; r.y = a * b.y; return r; }       // the int template of :367-373
static inline double2 operator*(const int a, const double2 b) { double2 r; r.x = a * b.x; r.y = a * b.y; return r; }
static inline double2 operator+(const double2 a, const double2 b) { double2 r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
static inline float2 operator+(const float2 a, const float2 b) { float2 r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
static inline double2 operator-(const double2 a, const double2 b) { double2 r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
static inline float2 operator-(const float2 a, const float2 b) { float2 r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
static inline double2 conj(const double2 a) { double2 r; r.x = a.x; r.y = -a.y; return r; }
static inline float2 conj(const float2 a) { float2 r; r.x = a.x; r.y = -a.y; return r; }

// "textures" are the arrays themselves (lib/qudaQKXTM_kernels.cu:314-331: tex1Dfetch of int4 -> double2)
typedef const double2 *texd_t;
typedef const float2 *texf_t;
static inline double2 fetch_double2(texd_t t, int i) { return t[i]; }
static inline float2 fetch_float2(texf_t t, int i) { return t[i]; }

// __constant__ geometry (lib/qudaQKXTM_kernels.cu:24-45, filled by init_qudaQKXTM :118-297); single rank: no dimension broken
static int c_nColor = 3, c_nDim = 4, c_nSpin = 4;
static int c_stride, c_threads;
static int c_localL[4], c_plusGhost[4], c_minusGhost[4], c_surface[4];
static bool c_dimBreak[4] = {false, false, false, false};
static double c_alphaGauss;

#include "qudaQKXTM_utils_lexic.h"     // LEXIC* macros, generated by the Makefile from include/qudaQKXTM_utils.h:25-29
#include <core_def.h>                  // /root/reference/lib/code_pieces/core_def.h

extern "C" void qref_set_geometry(const int L[4], double alphaGauss) {
  for (int i = 0; i < 4; i++) { c_localL[i] = L[i]; c_plusGhost[i] = 0; c_minusGhost[i] = 0; c_surface[i] = 0; c_dimBreak[i] = false; }
  c_threads = L[0] * L[1] * L[2] * L[3];
  c_stride = c_threads;                // GK_localVolume, no ghost in use on a single rank
  c_alphaGauss = alphaGauss;
}

#define THREAD_PREAMBLE dim3_ blockIdx = {sid_}, blockDim = {1}, threadIdx = {0}; (void)blockIdx; (void)blockDim; (void)threadIdx;

static void gauss_double_thread(int sid_, double2 *out, texd_t vecInTex, texd_t gaugeTex) {
  THREAD_PREAMBLE
#define FLOAT2 double2
#define READGAUGE_FLOAT READGAUGE_double
#define READVECTOR_FLOAT READVECTOR_double
#include <Gauss_core.h>
#undef READGAUGE_FLOAT
#undef READVECTOR_FLOAT
#undef FLOAT2
}
static void gauss_float_thread(int sid_, float2 *out, texf_t vecInTex, texf_t gaugeTex) {
  THREAD_PREAMBLE
#define FLOAT2 float2
#define READGAUGE_FLOAT READGAUGE_float
#define READVECTOR_FLOAT READVECTOR_float
#include <Gauss_core.h>
#undef READGAUGE_FLOAT
#undef READVECTOR_FLOAT
#undef FLOAT2
}
// one smearing step out = G(in) (run_GaussianSmearing, lib/qudaQKXTM_kernels.cu:1009-1021)
extern "C" void qref_gauss_step_double(double *out, const double *in, const double *gauge) {
#pragma omp parallel for
  for (int s = 0; s < c_threads; s++) gauss_double_thread(s, (double2 *)out, (texd_t)in, (texd_t)gauge);
}
extern "C" void qref_gauss_step_float(float *out, const float *in, const float *gauge) {
#pragma omp parallel for
  for (int s = 0; s < c_threads; s++) gauss_float_thread(s, (float2 *)out, (texf_t)in, (texf_t)gauge);
}

static void upload_thread(int sid_, const double2 *in, double2 *outEven, double2 *outOdd) {
  THREAD_PREAMBLE
#include <uploadToCuda_core.h>
}
static void download_thread(int sid_, double2 *out, const double2 *inEven, const double2 *inOdd) {
  THREAD_PREAMBLE
#include <downloadFromCuda_core.h>
}
// run_UploadToCuda / run_DownloadFromCuda (lib/qudaQKXTM_kernels.cu:1024-1108); NULL parity pointers as for parity fields
extern "C" void qref_upload(const double *in, double *outEven, double *outOdd) {
  for (int s = 0; s < c_threads / 2; s++) upload_thread(s, (const double2 *)in, (double2 *)outEven, (double2 *)outOdd);
}
extern "C" void qref_download(double *out, const double *inEven, const double *inOdd) {
  for (int s = 0; s < c_threads / 2; s++) download_thread(s, (double2 *)out, (const double2 *)inEven, (const double2 *)inOdd);
}

static void scale_thread(int sid_, double2 *inOut, double a) {
  THREAD_PREAMBLE
#include <scaleVector_core.h>
}
extern "C" void qref_scale_vector(double *inOut, double a) {
  for (int s = 0; s < c_threads; s++) scale_thread(s, (double2 *)inOut, a);
}

template <typename Float2> static void gamma5_thread(int sid_, Float2 *inOut) {
  THREAD_PREAMBLE
#include <apply_gamma5_vector_core.h>
}
extern "C" void qref_apply_gamma5_double(double *inOut) {
  for (int s = 0; s < c_threads; s++) gamma5_thread<double2>(s, (double2 *)inOut);
}

// ---- calculatePlaq: a block of THREADS_PER_BLOCK CUDA threads = that many OS threads meeting at a barrier ---------------------
#define THREADS_PER_BLOCK 64           // lib/qudaQKXTM_kernels.cu:11
static pthread_barrier_t g_block_barrier;
struct PlaqThreadArg { int block, thread; texd_t gauge; double *partial; };
static void plaq_thread_body(dim3_ blockIdx, dim3_ blockDim, dim3_ threadIdx, texd_t gaugeTexPlaq, double *partial_plaq) {
#define FLOAT2 double2
#define FLOAT double
#define READGAUGE_FLOAT READGAUGE_double
#define __shared__ static               /* one block at a time: a function-local static is shared by the block's threads */
#define __syncthreads() pthread_barrier_wait(&g_block_barrier)
#include <plaquette_core.h>
#undef __syncthreads
#undef __shared__
#undef READGAUGE_FLOAT
#undef FLOAT
#undef FLOAT2
}
static void *plaq_thread_entry(void *p) {
  PlaqThreadArg *a = (PlaqThreadArg *)p;
  dim3_ b = {a->block}, d = {THREADS_PER_BLOCK}, t = {a->thread};
  plaq_thread_body(b, d, t, a->gauge, a->partial);
  return nullptr;
}
// calculatePlaq_kernel<double> (lib/qudaQKXTM_kernels.cu:910-957): grid of V/64 blocks, host sum of the block partials,
// normalisation by totalVolume * nColor * 6
extern "C" double qref_calculate_plaq(const double *gauge) {
  const int nblocks = (c_threads + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK;
  std::vector<double> partial(nblocks, 0.0);
  pthread_barrier_init(&g_block_barrier, nullptr, THREADS_PER_BLOCK);
  for (int b = 0; b < nblocks; b++) {
    pthread_t th[THREADS_PER_BLOCK];
    PlaqThreadArg args[THREADS_PER_BLOCK];
    for (int t = 0; t < THREADS_PER_BLOCK; t++) {
      args[t].block = b; args[t].thread = t; args[t].gauge = (texd_t)gauge; args[t].partial = partial.data();
      pthread_create(&th[t], nullptr, plaq_thread_entry, &args[t]);
    }
    for (int t = 0; t < THREADS_PER_BLOCK; t++) pthread_join(th[t], nullptr);
  }
  pthread_barrier_destroy(&g_block_barrier);
  double plaquette = 0.0;
  for (int i = 0; i < nblocks; i++) plaquette += partial[i];
  return plaquette / ((double)c_threads * c_nColor * 6);
}

// ---- site-local propagator / vector kernels (lib/qudaQKXTM_kernels.cu:826-859) ----------------------------------------------
template <typename Float2> static void conj_vector_thread(int sid_, Float2 *inOut) {
  THREAD_PREAMBLE
#include <conjugate_vector_core.h>
}
template <typename Float2> static void conj_prop_thread(int sid_, Float2 *inOut) {
  THREAD_PREAMBLE
#include <conjugate_propagator_core.h>
}
template <typename Float2> static void gamma5_prop_thread(int sid_, Float2 *inOut) {
  THREAD_PREAMBLE
#include <apply_gamma5_propagator_core.h>
}
template <typename Float2> static void rotate_thread(int sid_, Float2 *inOut, int sign) {
  THREAD_PREAMBLE
#include <rotateToPhysicalBase_core.h>
}
extern "C" void qref_conjugate_vector_double(double *v) { for (int s = 0; s < c_threads; s++) conj_vector_thread<double2>(s, (double2 *)v); }
extern "C" void qref_conjugate_propagator_double(double *p) { for (int s = 0; s < c_threads; s++) conj_prop_thread<double2>(s, (double2 *)p); }
extern "C" void qref_gamma5_propagator_double(double *p) { for (int s = 0; s < c_threads; s++) gamma5_prop_thread<double2>(s, (double2 *)p); }
extern "C" void qref_rotate_physical_double(double *p, int sign) { for (int s = 0; s < c_threads; s++) rotate_thread<double2>(s, (double2 *)p, sign); }
extern "C" void qref_rotate_physical_float(float *p, int sign) { for (int s = 0; s < c_threads; s++) rotate_thread<float2>(s, (float2 *)p, sign); }

// ---- meson two-point contraction (lib/qudaQKXTM_kernels.cu:474-511, launcher :1127-1225) --------------------------------------
#define MAX_NMOMENTA 5000              // include/qudaQKXTM_utils.h:19
#define PI 3.141592653589793           // lib/qudaQKXTM_kernels.cu:12
#include "qkxtm_meson_tables.h"        // GK_mesons_indices / GK_mesons_values, the reference's own initialisers
#define c_mesons_indices GK_mesons_indices
#define c_mesons_values GK_mesons_values
static int c_Nmoms;
static short int c_moms[MAX_NMOMENTA][3];
static int c_procPosition[4] = {0, 0, 0, 0};
static int c_totalL[4];

// a grid of blocks of THREADS_PER_BLOCK threads, one block at a time, its threads as OS threads meeting at g_block_barrier
static void run_grid(int nblocks, const std::function<void(dim3_, dim3_, dim3_, dim3_)> &body) {
  pthread_barrier_init(&g_block_barrier, nullptr, THREADS_PER_BLOCK);
  for (int b = 0; b < nblocks; b++) {
    std::vector<std::thread> th;
    for (int t = 0; t < THREADS_PER_BLOCK; t++)
      th.emplace_back([&, b, t]() { dim3_ bi = {b}, bd = {THREADS_PER_BLOCK}, ti = {t}, gd = {nblocks}; body(bi, bd, ti, gd); });
    for (auto &x : th) x.join();
  }
  pthread_barrier_destroy(&g_block_barrier);
}

#define __shared__ static
#define __syncthreads() pthread_barrier_wait(&g_block_barrier)
static void mesons_mom_float_body(dim3_ blockIdx, dim3_ blockDim, dim3_ threadIdx, dim3_ gridDim, float2 *block, texf_t prop1Tex, texf_t prop2Tex,
                                  int it, int x0, int y0, int z0) {
#define FLOAT2 float2
#define FLOAT float
#define FETCH_FLOAT2 fetch_float2
#include <contractMesons_core.h>
#undef PROP
#undef FETCH_FLOAT2
#undef FLOAT2
#undef FLOAT
}
static void mesons_mom_double_body(dim3_ blockIdx, dim3_ blockDim, dim3_ threadIdx, dim3_ gridDim, double2 *block, texd_t prop1Tex, texd_t prop2Tex,
                                   int it, int x0, int y0, int z0) {
#define FLOAT2 double2
#define FLOAT double
#define FETCH_FLOAT2 fetch_double2
#include <contractMesons_core.h>
#undef PROP
#undef FETCH_FLOAT2
#undef FLOAT2
#undef FLOAT
}
static void mesons_pos_float_body(dim3_ blockIdx, dim3_ blockDim, dim3_ threadIdx, dim3_ gridDim, float2 *block, texf_t prop1Tex, texf_t prop2Tex,
                                  int it, int x0, int y0, int z0) {
#define FLOAT2 float2
#define FLOAT float
#define FETCH_FLOAT2 fetch_float2
#include <contractMesons_core_PosSpace.h>
#undef PROP
#undef FETCH_FLOAT2
#undef FLOAT2
#undef FLOAT
}
#undef __syncthreads
#undef __shared__

extern "C" void qref_set_momenta(const int *moms, int nmoms) {
  c_Nmoms = nmoms;
  for (int i = 0; i < nmoms; i++) for (int d = 0; d < 3; d++) c_moms[i][d] = (short int)moms[3 * i + d];
  for (int d = 0; d < 4; d++) { c_totalL[d] = c_localL[d]; c_procPosition[d] = 0; }      // single rank
}

// contractMesons_kernel<Float2, Float> of the launcher, MOMENTUM_SPACE branch (lib/qudaQKXTM_kernels.cu:1167-1199), for all local
// time slices: out[it][imom][iu][ip][re,im] (the reference's corr[it*Nmoms*2 + imom*2 + ri][iu][ip])
template <typename Float, typename Float2, typename Tex, typename Body>
static void mesons_mom(Float *out, const Float *prop1, const Float *prop2, const int src[3], Body body) {
  const int SpVol = c_threads / c_localL[3];
  const int grid = (SpVol + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK;
  std::vector<Float> h_partial_block((size_t)c_Nmoms * 2 * 10 * grid * 2);
  for (int it = 0; it < c_localL[3]; it++) {
    run_grid(grid, [&](dim3_ b, dim3_ d, dim3_ t, dim3_ g) { body(b, d, t, g, (Float2 *)h_partial_block.data(), (Tex)prop1, (Tex)prop2, it, src[0], src[1], src[2]); });
    for (int imom = 0; imom < c_Nmoms; imom++)
      for (int iu = 0; iu < 2; iu++)
        for (int ip = 0; ip < 10; ip++) {
          Float re = 0, im = 0;           // the launcher's Float `reduction` buffer, blocks summed in order
          for (int i = 0; i < grid; i++) {
            re += h_partial_block[(size_t)imom * 2 * 10 * grid * 2 + iu * 10 * grid * 2 + ip * grid * 2 + i * 2 + 0];
            im += h_partial_block[(size_t)imom * 2 * 10 * grid * 2 + iu * 10 * grid * 2 + ip * grid * 2 + i * 2 + 1];
          }
          Float *o = out + ((((size_t)it * c_Nmoms + imom) * 2 + iu) * 10 + ip) * 2;
          o[0] = re; o[1] = im;
        }
  }
}
extern "C" void qref_contract_mesons_mom_float(float *out, const float *prop1, const float *prop2, const int src[3]) {
  mesons_mom<float, float2, texf_t>(out, prop1, prop2, src, mesons_mom_float_body);
}
// the double instantiation of the same body (contractMesons_kernel_double, lib/qudaQKXTM_kernels.cu:500-511; its launcher refuses
// precision 8, :1222, but the kernel is in the tree): a tight pin for the restatement
extern "C" void qref_contract_mesons_mom_double(double *out, const double *prop1, const double *prop2, const int src[3]) {
  mesons_mom<double, double2, texd_t>(out, prop1, prop2, src, mesons_mom_double_body);
}
// POSITION_SPACE branch (lib/qudaQKXTM_kernels.cu:1142-1166): out[it][sv][iu][ip][re,im]
extern "C" void qref_contract_mesons_pos_float(float *out, const float *prop1, const float *prop2) {
  const int SpVol = c_threads / c_localL[3];
  const int grid = (SpVol + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK;
  const size_t alloc = (size_t)THREADS_PER_BLOCK * grid;
  std::vector<float> blk(alloc * 2 * 10 * 2);
  for (int it = 0; it < c_localL[3]; it++) {
    run_grid(grid, [&](dim3_ b, dim3_ d, dim3_ t, dim3_ g) { mesons_pos_float_body(b, d, t, g, (float2 *)blk.data(), (texf_t)prop1, (texf_t)prop2, it, 0, 0, 0); });
    for (int pt = 0; pt < 2; pt++)
      for (int mes = 0; mes < 10; mes++)
        for (int sv = 0; sv < SpVol; sv++)
          for (int ri = 0; ri < 2; ri++)
            out[((((size_t)it * SpVol + sv) * 2 + pt) * 10 + mes) * 2 + ri] = blk[ri + 2 * sv + 2 * alloc * mes + 2 * alloc * 10 * pt];
  }
}

// ---- baryon two-point contraction (lib/qudaQKXTM_kernels.cu:513-524, launcher :1228-1311) -------------------------------------
#include "qkxtm_baryon_tables.h"       // GK_{NTN,NTR,RTN,RTR,Delta}_{indices,values}, the reference's own initialisers
#define c_NTN_indices GK_NTN_indices
#define c_NTN_values GK_NTN_values
#define c_NTR_indices GK_NTR_indices
#define c_NTR_values GK_NTR_values
#define c_RTN_indices GK_RTN_indices
#define c_RTN_values GK_RTN_values
#define c_RTR_indices GK_RTR_indices
#define c_RTR_values GK_RTR_values
#define c_Delta_indices GK_Delta_indices
#define c_Delta_values GK_Delta_values
// the six permutations of (0,1,2) and their signs: the Levi-Civita symbol (lib/qudaQKXTM_kernels.cu:184-197)
static const int c_eps[6][3] = {{0, 1, 2}, {2, 0, 1}, {1, 2, 0}, {2, 1, 0}, {0, 2, 1}, {1, 0, 2}};
static const int c_sgn_eps[6] = {+1, +1, +1, -1, -1, -1};

#define __shared__ static
#define __syncthreads() pthread_barrier_wait(&g_block_barrier)
static void baryons_mom_float_body(dim3_ blockIdx, dim3_ blockDim, dim3_ threadIdx, dim3_ gridDim, float2 *block, texf_t prop1Tex, texf_t prop2Tex,
                                   int it, int x0, int y0, int z0, int ip) {
#define FLOAT2 float2
#define FLOAT float
#define FETCH_FLOAT2 fetch_float2
#include <contractBaryons_core.h>
#undef PROP
#undef FETCH_FLOAT2
#undef FLOAT2
#undef FLOAT
}
// the same body instantiated in double (the reference keeps this instantiation commented out, lib/qudaQKXTM_kernels.cu:540-551):
// a tight pin for the restatement
static void baryons_mom_double_body(dim3_ blockIdx, dim3_ blockDim, dim3_ threadIdx, dim3_ gridDim, double2 *block, texd_t prop1Tex, texd_t prop2Tex,
                                    int it, int x0, int y0, int z0, int ip) {
#define FLOAT2 double2
#define FLOAT double
#define FETCH_FLOAT2 fetch_double2
#include <contractBaryons_core.h>
#undef PROP
#undef FETCH_FLOAT2
#undef FLOAT2
#undef FLOAT
}
#undef __syncthreads
#undef __shared__

// contractBaryons_kernel<Float2, Float>, MOMENTUM_SPACE branch (lib/qudaQKXTM_kernels.cu:1276-1311), all local time slices:
// out[it][imom][iu][ip][gamma][gammap][re,im] (the reference's corr[it*Nmoms*2 + imom*2 + ri][iu][ip][gamma][gammap])
template <typename Float, typename Float2, typename Tex, typename Body>
static void baryons_mom(Float *out, const Float *prop1, const Float *prop2, const int src[3], Body body) {
  const int SpVol = c_threads / c_localL[3];
  const int grid = (SpVol + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK;
  std::vector<Float> h((size_t)c_Nmoms * 2 * 4 * 4 * grid * 2);
  for (int it = 0; it < c_localL[3]; it++)
    for (int ip = 0; ip < 10; ip++) {
      run_grid(grid, [&](dim3_ b, dim3_ d, dim3_ t, dim3_ g) { body(b, d, t, g, (Float2 *)h.data(), (Tex)prop1, (Tex)prop2, it, src[0], src[1], src[2], ip); });
      for (int imom = 0; imom < c_Nmoms; imom++)
        for (int iu = 0; iu < 2; iu++)
          for (int ga = 0; ga < 4; ga++)
            for (int gap = 0; gap < 4; gap++) {
              Float re = 0, im = 0;
              for (int i = 0; i < grid; i++) {
                const size_t k = (size_t)imom * 2 * 4 * 4 * grid * 2 + (size_t)iu * 4 * 4 * grid * 2 + (size_t)ga * 4 * grid * 2 + (size_t)gap * grid * 2 + i * 2;
                re += h[k]; im += h[k + 1];
              }
              Float *o = out + ((((((size_t)it * c_Nmoms + imom) * 2 + iu) * 10 + ip) * 4 + ga) * 4 + gap) * 2;
              o[0] = re; o[1] = im;
            }
    }
}
extern "C" void qref_contract_baryons_mom_float(float *out, const float *prop1, const float *prop2, const int src[3]) {
  baryons_mom<float, float2, texf_t>(out, prop1, prop2, src, baryons_mom_float_body);
}
extern "C" void qref_contract_baryons_mom_double(double *out, const double *prop1, const double *prop2, const int src[3]) {
  baryons_mom<double, double2, texd_t>(out, prop1, prop2, src, baryons_mom_double_body);
}

// ---- fixed-sink three-point function: sequential sources and the ultra-local insertion (lib/qudaQKXTM_kernels.cu:412-424,552-626) -----
enum WHICHPARTICLE { PROTON, NEUTRON };                      // include/qudaQKXTM_utils.h:128
enum WHICHPROJECTOR { G4, G5G123, G5G1, G5G2, G5G3 };        // include/qudaQKXTM_utils.h:129
static inline float norm(const float2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
static inline double norm(const double2 a) { return sqrt(a.x * a.x + a.y * a.y); }
template <typename Float2> static inline void get_Projector(Float2 projector[4][4], WHICHPARTICLE PARTICLE, WHICHPROJECTOR PID) {
#include <projectors_tm_base.h>
}
template <typename Float2> static inline void get_Operator(Float2 gamma[4][4], int flag, WHICHPARTICLE TESTPARTICLE, int partFlag) {
#include <gammas_tm_base.h>
}
static void seq1_double_thread(int sid_, double2 *out, int timeslice, texd_t tex1, texd_t tex2, int c_nu, int c_c2, WHICHPROJECTOR PID, WHICHPARTICLE PARTICLE) {
  THREAD_PREAMBLE
#define FLOAT2 double2
#define FLOAT double
#define FETCH_FLOAT2 fetch_double2
#include <seqSourceFixSinkPart1_core.h>
#undef FETCH_FLOAT2
#undef FLOAT2
#undef FLOAT
}
static void seq2_double_thread(int sid_, double2 *out, int timeslice, texd_t tex, int c_nu, int c_c2, WHICHPROJECTOR PID, WHICHPARTICLE PARTICLE) {
  THREAD_PREAMBLE
#define FLOAT2 double2
#define FLOAT double
#define FETCH_FLOAT2 fetch_double2
#include <seqSourceFixSinkPart2_core.h>
#undef FETCH_FLOAT2
#undef FLOAT2
#undef FLOAT
}
// run_seqSourceFixSinkPart1 / Part2 (lib/qudaQKXTM_kernels.cu:1428-1460): out is a 4-d vector, only the time slice is written;
// tex1 / tex2 / tex are PROPAGATOR3D fields [4][4][3][3][V3]
extern "C" void qref_seq_source_part1_double(double *out, int timeslice, const double *p3d_1, const double *p3d_2, int nu, int c2, int pid, int particle) {
  const int V3 = c_localL[0] * c_localL[1] * c_localL[2];
  for (int s = 0; s < V3; s++) seq1_double_thread(s, (double2 *)out, timeslice, (texd_t)p3d_1, (texd_t)p3d_2, nu, c2, (WHICHPROJECTOR)pid, (WHICHPARTICLE)particle);
}
extern "C" void qref_seq_source_part2_double(double *out, int timeslice, const double *p3d, int nu, int c2, int pid, int particle) {
  const int V3 = c_localL[0] * c_localL[1] * c_localL[2];
  for (int s = 0; s < V3; s++) seq2_double_thread(s, (double2 *)out, timeslice, (texd_t)p3d, nu, c2, (WHICHPROJECTOR)pid, (WHICHPARTICLE)particle);
}

#define __shared__ static
#define __syncthreads() pthread_barrier_wait(&g_block_barrier)
static void fixsink_local_float_body(dim3_ blockIdx, dim3_ blockDim, dim3_ threadIdx, dim3_ gridDim, float2 *block, texf_t fwdTex, texf_t seqTex,
                                     WHICHPARTICLE TESTPARTICLE, int partflag, int it, int x0, int y0, int z0) {
#define FLOAT2 float2
#define FLOAT float
#define FETCH_FLOAT2 fetch_float2
#include <fixSinkContractions_local_core.h>
#undef PROP
#undef FETCH_FLOAT2
#undef FLOAT2
#undef FLOAT
}
// the same body in double (the reference launches the float instantiation only): a tight pin
static void fixsink_local_double_body(dim3_ blockIdx, dim3_ blockDim, dim3_ threadIdx, dim3_ gridDim, double2 *block, texd_t fwdTex, texd_t seqTex,
                                      WHICHPARTICLE TESTPARTICLE, int partflag, int it, int x0, int y0, int z0) {
#define FLOAT2 double2
#define FLOAT double
#define FETCH_FLOAT2 fetch_double2
#include <fixSinkContractions_local_core.h>
#undef PROP
#undef FETCH_FLOAT2
#undef FLOAT2
#undef FLOAT
}
#undef __syncthreads
#undef __shared__
// the local part of run_fixSinkContractions, MOMENTUM_SPACE: out[it][imom][iop][re,im]
template <typename Float, typename Float2, typename Tex, typename Body>
static void fixsink_local(Float *out, const Float *fwd, const Float *seq, int particle, int partflag, const int src[3], Body body) {
  const int SpVol = c_threads / c_localL[3];
  const int grid = (SpVol + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK;
  std::vector<Float> h((size_t)c_Nmoms * 16 * grid * 2);
  for (int it = 0; it < c_localL[3]; it++) {
    run_grid(grid, [&](dim3_ b, dim3_ d, dim3_ t, dim3_ g) { body(b, d, t, g, (Float2 *)h.data(), (Tex)fwd, (Tex)seq, (WHICHPARTICLE)particle, partflag, it, src[0], src[1], src[2]); });
    for (int imom = 0; imom < c_Nmoms; imom++)
      for (int iop = 0; iop < 16; iop++) {
        Float re = 0, im = 0;
        for (int i = 0; i < grid; i++) { re += h[((size_t)imom * 16 * grid + (size_t)iop * grid + i) * 2]; im += h[((size_t)imom * 16 * grid + (size_t)iop * grid + i) * 2 + 1]; }
        Float *o = out + (((size_t)it * c_Nmoms + imom) * 16 + iop) * 2;
        o[0] = re; o[1] = im;
      }
  }
}
extern "C" void qref_fixsink_local_float(float *out, const float *fwd, const float *seq, int particle, int partflag, const int src[3]) {
  fixsink_local<float, float2, texf_t>(out, fwd, seq, particle, partflag, src, fixsink_local_float_body);
}
extern "C" void qref_fixsink_local_double(double *out, const double *fwd, const double *seq, int particle, int partflag, const int src[3]) {
  fixsink_local<double, double2, texd_t>(out, fwd, seq, particle, partflag, src, fixsink_local_double_body);
}
// the tables themselves, so that the restatement can state them as formulas and check the formulas entry by entry
extern "C" void qref_get_projector(double *out32, int pid, int particle) {
  double2 p[4][4];
  get_Projector(p, (WHICHPARTICLE)particle, (WHICHPROJECTOR)pid);
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { out32[(i * 4 + j) * 2] = p[i][j].x; out32[(i * 4 + j) * 2 + 1] = p[i][j].y; }
}
extern "C" void qref_get_operator(double *out32, int flag, int particle, int partflag) {
  double2 g[4][4];
  get_Operator(g, flag, (WHICHPARTICLE)particle, partflag);
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { out32[(i * 4 + j) * 2] = g[i][j].x; out32[(i * 4 + j) * 2 + 1] = g[i][j].y; }
}

// ---- conserved-current (Noether) and one-derivative insertions (lib/qudaQKXTM_kernels.cu:676-760, launcher :1636-1715) -------------
// single rank: c_dimBreak is false in every direction, so the bodies take their periodic "Str" paths (no ghost zones)
#define __shared__ static
#define __syncthreads() pthread_barrier_wait(&g_block_barrier)
static void noether_double_body(dim3_ blockIdx, dim3_ blockDim, dim3_ threadIdx, dim3_ gridDim, double2 *block, texd_t fwdTex, texd_t seqTex, texd_t gaugeTex,
                                WHICHPARTICLE TESTPARTICLE, int partflag, int it, int x0, int y0, int z0) {
#define FLOAT2 double2
#define FLOAT double
#define FETCH_FLOAT2 fetch_double2
#include <fixSinkContractions_noether_core.h>
#undef PROP
#undef GAUGE
#undef PROPplusSur
#undef PROPminusSur
#undef GAUGEminusSur
#undef PROPplusStr
#undef PROPminusStr
#undef GAUGEminusStr
#undef FETCH_FLOAT2
#undef FLOAT2
#undef FLOAT
}
static void oneD_double_body(dim3_ blockIdx, dim3_ blockDim, dim3_ threadIdx, dim3_ gridDim, double2 *block, texd_t fwdTex, texd_t seqTex, texd_t gaugeTex,
                             WHICHPARTICLE TESTPARTICLE, int partflag, int it, int dir, int x0, int y0, int z0) {
#define FLOAT2 double2
#define FLOAT double
#define FETCH_FLOAT2 fetch_double2
#include <fixSinkContractions_oneD_core.h>
#undef PROP
#undef GAUGE
#undef PROPplusSur
#undef PROPminusSur
#undef GAUGEminusSur
#undef PROPplusStr
#undef PROPminusStr
#undef GAUGEminusStr
#undef FETCH_FLOAT2
#undef FLOAT2
#undef FLOAT
}
#undef __syncthreads
#undef __shared__
// noether: out_n[it][imom][dir][re,im]; oneD: out_d[it][imom][dir][iop][re,im] (the reference's corrThp_oneD[it*Nmoms*4*16*2 + imom*4*16*2 +
// dir*16*2 + iop*2 + ri], lib/qudaQKXTM_kernels.cu:1694-1712)
extern "C" void qref_fixsink_derivative_double(double *out_n, double *out_d, const double *fwd, const double *seq, const double *gauge, int particle,
                                               int partflag, const int src[3]) {
  const int SpVol = c_threads / c_localL[3];
  const int grid = (SpVol + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK;
  std::vector<double> h((size_t)c_Nmoms * 16 * grid * 2);
  for (int it = 0; it < c_localL[3]; it++) {
    run_grid(grid, [&](dim3_ b, dim3_ d, dim3_ t, dim3_ g) {
      noether_double_body(b, d, t, g, (double2 *)h.data(), (texd_t)fwd, (texd_t)seq, (texd_t)gauge, (WHICHPARTICLE)particle, partflag, it, src[0], src[1], src[2]);
    });
    for (int imom = 0; imom < c_Nmoms; imom++)
      for (int dir = 0; dir < 4; dir++) {
        double re = 0, im = 0;
        for (int i = 0; i < grid; i++) { re += h[((size_t)imom * 4 * grid + (size_t)dir * grid + i) * 2]; im += h[((size_t)imom * 4 * grid + (size_t)dir * grid + i) * 2 + 1]; }
        double *o = out_n + (((size_t)it * c_Nmoms + imom) * 4 + dir) * 2;
        o[0] = re; o[1] = im;
      }
    for (int dir = 0; dir < 4; dir++) {
      run_grid(grid, [&](dim3_ b, dim3_ d, dim3_ t, dim3_ g) {
        oneD_double_body(b, d, t, g, (double2 *)h.data(), (texd_t)fwd, (texd_t)seq, (texd_t)gauge, (WHICHPARTICLE)particle, partflag, it, dir, src[0], src[1], src[2]);
      });
      for (int imom = 0; imom < c_Nmoms; imom++)
        for (int iop = 0; iop < 16; iop++) {
          double re = 0, im = 0;
          for (int i = 0; i < grid; i++) { re += h[((size_t)imom * 16 * grid + (size_t)iop * grid + i) * 2]; im += h[((size_t)imom * 16 * grid + (size_t)iop * grid + i) * 2 + 1]; }
          double *o = out_d + ((((size_t)it * c_Nmoms + imom) * 4 + dir) * 16 + iop) * 2;
          o[0] = re; o[1] = im;
        }
    }
  }
}
