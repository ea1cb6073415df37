// tmq_internal.h -- library-internal declarations (context, field handles, kernel launchers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <set>
#include <string>
#include <vector>
#include "tmq_types.h"
#include "tmq_reduce.cuh"

struct tmq_ctx;

struct tmq_spinor {
  tmq_ctx *ctx;
  int prec;      // 8 | 4
  int subset;    // 1 parity | 2 full
  void *d;       // device pointer (parity block 0, then block 1 for FULL)
  size_t bytes;
  bool owns;     // false for Even()/Odd() views
  tmq_spinor *view[2];
};

namespace tmq {

struct Comm;   // NCCL state (tmq_comm.cpp)

// Ghost-zone arena of the peer-memory halo path: one cudaMalloc per context (so that ONE IPC handle exposes it to
// the neighbours), identical layout on every rank: recv[buf 0..1][prec 0..1][dim 2..3][dir 0..1] + arrival flags.
struct HaloArena {
  size_t recv[2][2][4][2];
  size_t flag;            // byte offset of unsigned int flags[2 buf][4 dim][2 dir]
  size_t mbox;            // byte offset of double mailbox[2 buf][4][TMQ_MAX_RANKS] (scalar all-reduce)
  size_t mflag;           // byte offset of unsigned int mflags[2 buf][TMQ_MAX_RANKS]
  size_t bytes;
};
struct GaugeStore {
  void *d = nullptr;     // [2][4][3 or 9][Vh] in vec / cplx units
  size_t bytes = 0;
};

// per-precision scratch used by the operator (temporaries of M, M^dag M, CG)
constexpr int NSCRATCH = 7;
struct Scratch {
  void *tmp[NSCRATCH] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // parity fields
};

}  // namespace tmq

struct tmq_ctx {
  int device;
  cudaStream_t stream;       // compute stream
  cudaStream_t comm_stream;  // halo exchange stream
  cudaEvent_t ev_a, ev_b, ev_pack, ev_halo, ev_r2;
  cudaEvent_t ev_ring[8];    // lagged CG: recorded behind the |r|^2 read-back of iteration k (entry k & 7)
  int cg_iter_cur = 0;       // > 0 while tmq_cg_mdagm enqueues iteration cg_iter_cur of a solve whose stopping test lags by one iteration
  int opt_cg_lag = 1;        // TMQ_OPT_CG_LAG: iterations the host loop runs ahead of the residual it reads (0: synchronous loop)
  tmq::Geom g;
  int grid[4], coord[4];
  int nranks, rank;
  long long Vglobal;
  int tile[3];
  bool multi;                // any dimension partitioned
  // gauge
  int recon;                 // 12 | 18
  int t_boundary;
  tmq::GaugeStore gauge_d, gauge_s;
  // twisted-clover: C and (C + i a g5)^-1 in the chiral basis, vec[2 parity][36][Vh], fp64 + fp32 copies (tmq_clover.cu)
  tmq::GaugeStore clov_c_d, clov_inv_d, clov_c_s, clov_inv_s;
  bool clover_on = false, clov_inv_valid = false;
  double clover_coeff = 0, clov_inv_a = 0;
  // operator
  double kappa, mu;
  int matpc;
  bool op_set;
  // scratch parity fields per precision
  tmq::Scratch scr_d, scr_s;
  // reductions
  double *partials;          // device
  size_t partials_len;
  unsigned int *ticket;      // device
  double *scal;              // device scalar block [SC_COUNT]
  double *h_scal;            // pinned, mapped host mirror [SC_COUNT]
  double *h_scal_dev;        // the same memory as the device sees it (written by scal_to_host_kernel)
  // halo buffers per precision: [dim][dir] send / recv, vec[3][face]
  void *halo_send[2][4][2];
  void *halo_send2[2][4][2];   // second set of send buffers (halo mode 4: the faces of application N+1 are packed while the copies of N still read)
  void *halo_recv[2][4][2];
  tmq::Comm *comm;
  long long launches;
  std::vector<double> cg_hist;
  double cg_loop_secs = 0;     // wall time of the iteration loop of the last solve (without set-up and the true-residual computation)
  int cg_reliable_updates = 0;
  // grow-only device staging buffer for host <-> native conversions
  void *stage;
  size_t stage_bytes;
  // host <-> device pipelining of multi-RHS drivers (tmq_host_prefetch / tmq_spinor_from_prefetch / tmq_spinor_to_host_async): two
  // upload and two download staging slots of one FULL host field each, their own copy streams, and the events that order slot reuse
  cudaStream_t up_stream = nullptr, down_stream = nullptr;
  void *io_up[2] = {nullptr, nullptr}, *io_down[2] = {nullptr, nullptr};
  cudaEvent_t ev_up_done[2] = {nullptr, nullptr}, ev_up_free[2] = {nullptr, nullptr};
  cudaEvent_t ev_down_ready[2] = {nullptr, nullptr}, ev_down_done[2] = {nullptr, nullptr};
  std::vector<void *> host_registered;
  // live spinor handles allocated on this context (freed by tmq_destroy; their handles die with the context)
  std::set<tmq_spinor *> spinors;
  // timing-kernel scratch (tmq_time_kernel)
  int sms;
  int opt_prefetch;
  int opt_debug;             // timing experiments only (option 99)
  int opt_pack_async;        // copy-engine halo path: launch the face pack on the exchange stream, beside the Dslash
  int opt_smear_block_t;     // time slices per L2-resident block of the Gaussian smearing (0 = from the L2 size)
  // peer-memory halo path
  int opt_p2p;               // requested: 0 NCCL send/recv, 1 peer stores from the pack kernel, 2 copy-engine peer copies
  int opt_pre_pct;           // % of the interior CTAs scheduled before the boundary CTAs in a fused launch
  int opt_halo_timeout_ms;   // wall-clock limit of a halo / mailbox wait before the device error scalar is raised
  bool p2p;                  // active: every neighbour's arena is mapped
  char *arena;               // own ghost arena (cudaMalloc)
  tmq::HaloArena arena_layout;
  char *peer_arena[4][2];    // [dim][0 = rank-1, 1 = rank+1]: base of that neighbour's arena in this address space
  char *rank_arena[tmq::TMQ_MAX_RANKS];   // every rank's arena (own for rank == self); scalar all-reduce mailboxes
  unsigned int red_seq;
  std::vector<void *> ipc_opened;
  unsigned int halo_seq;
  // fused halo mode: application `prepacked_seq` has had its faces sent by the launch that produced its input field
  unsigned int prepacked_seq = 0;
  const void *prepacked_in = nullptr;
  int prepacked_dagger = 0, prepacked_parity = 0, prepacked_prec = 0;
  unsigned int *ticket2;     // pack-kernel ticket
  unsigned int *seq_table;   // device table seq_table[i] = i: source of the copy-engine flag writes
  // grow-only device work space of the meson contraction (site values + the stages of the separable Fourier sum)
  int opt_contract_slices = 0;   // TMQ_OPT_CONTRACT_SLICES
  void *contract_ws = nullptr;
  size_t contract_ws_bytes = 0;
};

namespace tmq {

void set_error(const char *fmt, ...);
#define TMQ_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      tmq::set_error("%s:%d CUDA error: %s (%s)", __FILE__, __LINE__, cudaGetErrorString(e__), #call); \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

#define TMQ_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) { tmq::set_error(__VA_ARGS__); return 1; } \
  } while (0)
#define TMQ_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__) return rc__;       \
  } while (0)

inline size_t vec_bytes(int prec) { return prec == 8 ? 32 : 16; }
inline size_t parity_bytes(const tmq_ctx *c, int prec) { return (size_t)6 * c->g.Vh * vec_bytes(prec); }

Enum make_enum(const Geom &g, const int lo[3], const int ext[3], const int tile_pref[3], const int step[3] = nullptr);

// ---- one Dslash-class application (tmq_api.cu: possibly split into interior + boundary launches) ----------------
struct Tw { double c, a; int inv = 0, dag = 0; };   // out = c (1 + i a g5) in; inv / dag say which of A, A^-1, A^dag, A^-dag it is (clover path)
struct HopSpec {
  int epi = EPI_PLAIN;
  int out_parity = 0;
  int dagger = 0;
  Tw t1 = {1, 0};      // post-hop twist
  Tw tx = {1, 0};      // twist on the x term
  Tw t3 = {1, 0};      // final twist
  double k = 0;
  const void *x = nullptr;
  void *r = nullptr;
  const void *y = nullptr;          // EPI_CHEB
  void *out2 = nullptr;             // clover EPI_MDAGM2: second output y = M p
  int cl_plain_x = 0;               // clover: use the x term as it is
  double d1 = 0, d2 = 0, d3 = 0;    // EPI_CHEB
  int red_slot = SC_T3;
  int alpha_num = SC_ONE, alpha_den = SC_ONE;
  // fused halo mode: this application's output is the input of the NEXT application (with dagger next_dagger): its boundary CTAs pack
  // and send the faces themselves.  Only set inside chains of applications that this library issues back to back.
  int pack_next = 0, next_dagger = 0;
  // the faces of THIS application may have been sent ahead by the launch issued immediately before it.  Pointer equality of the field
  // is not enough (a work field is rewritten between solves), so only the chains set it, for the links they issue back to back.
  int accept_ahead = 0;
};
inline double tw_a(const tmq_ctx *c) { return 2.0 * c->kappa * c->mu; }
inline Tw tw_A(const tmq_ctx *c, int dag) { return {1.0, dag ? -tw_a(c) : tw_a(c), 0, dag}; }
inline Tw tw_Ainv(const tmq_ctx *c, int dag) {
  const double a = tw_a(c);
  return {1.0 / (1.0 + a * a), dag ? a : -a, 1, dag};
}
// site-local A / A^-1 / A^dag / A^-dag on the parity block `parity`: the constant twist, or the site's clover blocks
int site_op(tmq_ctx *c, int prec, void *out, const void *in, int parity, const Tw &w);
cudaError_t clover_apply(int prec, void *out, const void *in, const void *M, int Vh, int dag, double a, cudaStream_t st);
int clover_update_inverse(tmq_ctx *c);

int apply_hop(tmq_ctx *c, int prec, void *out, const void *in, const HopSpec &s);
int ensure_scratch(tmq_ctx *c, int prec, int n);
inline void *scr(tmq_ctx *c, int prec, int i) { return (prec == 8 ? c->scr_d : c->scr_s).tmp[i]; }
int reduce_finish(tmq_ctx *c, int slot, int n);
int fetch_scal(tmq_ctx *c, int slot, int n, double *out);
int scal_to_host(tmq_ctx *c, int slot, int n);
int check_device_error(tmq_ctx *c);
int op_matpc(tmq_ctx *c, int prec, void *out, const void *in, int dagger);
int op_mdagm(tmq_ctx *c, int prec, void *out, const void *in, int pap_slot);
// out = p(M^dag M) in, the Chebyshev filter of the eigensolver (tmq_eig.cu)
int op_poly_mdagm(tmq_ctx *c, int prec, void *out, const void *in, int deg, double amin, double amax);
void eig_release(tmq_ctx *c);   // frees the eigensolver workspace of a context

// dslash launchers (one TU per precision x recon)
cudaError_t launch_dslash_d8(int epi, bool multi, const DslashArgs<double> &A, cudaStream_t st);
cudaError_t launch_dslash_s8(int epi, bool multi, const DslashArgs<float> &A, cudaStream_t st);
cudaError_t launch_dslash_d12(int epi, bool multi, const DslashArgs<double> &A, cudaStream_t st);
cudaError_t launch_dslash_d18(int epi, bool multi, const DslashArgs<double> &A, cudaStream_t st);
cudaError_t launch_dslash_s12(int epi, bool multi, const DslashArgs<float> &A, cudaStream_t st);
cudaError_t launch_dslash_s18(int epi, bool multi, const DslashArgs<float> &A, cudaStream_t st);

// blas launchers (tmq_blas.cu).  n = number of vectors (6*Vh per parity block).  prec = 8 | 4.
struct BlasRed { double *partials; unsigned int *ticket; double *scal; int slot; };
cudaError_t blas_zero(void *x, size_t bytes, cudaStream_t st);
cudaError_t blas_copy(void *dst, int dprec, const void *src, int sprec, size_t n, cudaStream_t st);
cudaError_t blas_axpby(int prec, double a, const void *x, double b, void *y, size_t n, cudaStream_t st);   // y = a x + b y
cudaError_t blas_ax(int prec, double a, void *x, size_t n, cudaStream_t st);
cudaError_t blas_caxpy(int prec, double ar, double ai, const void *x, void *y, size_t n, cudaStream_t st);
cudaError_t blas_cxpaypbz(int prec, const void *x, double ar, double ai, const void *y, double br, double bi, void *z,
                          size_t n, cudaStream_t st);
cudaError_t blas_norm2(int prec, const void *x, size_t n, const BlasRed &r, cudaStream_t st);
cudaError_t blas_redot(int prec, const void *x, const void *y, size_t n, const BlasRed &r, cudaStream_t st);
cudaError_t blas_cdot(int prec, const void *x, const void *y, size_t n, const BlasRed &r, cudaStream_t st);  // slot, slot+1
cudaError_t blas_axpy_norm(int prec, double a, const void *x, void *y, size_t n, const BlasRed &r, cudaStream_t st);
cudaError_t blas_xmy_norm(int prec, const void *x, void *y, size_t n, const BlasRed &r, cudaStream_t st);
cudaError_t blas_axpy_zpbx(int prec, double a, void *x, void *y, const void *z, double b, size_t n, cudaStream_t st);
// CG update with device-resident scalars: alpha = s[an]/s[ad], beta = s[bn]/s[bd];  x += alpha p ; p = r + beta p
cudaError_t blas_cg_update(int prec, void *x, void *p, const void *r, size_t n, const double *scal, int an, int ad,
                           int bn, int bd, cudaStream_t st, int cg_iter = 0);
// mixed precision accumulate: y(double) += x(float)
cudaError_t blas_xpy_mixed(void *y_d, const void *x_s, size_t n, cudaStream_t st);
// site-local twist / gamma5 on a parity block: out = c (in + i a g5 in)
cudaError_t blas_twist(int prec, void *out, const void *in, double c, double a, int Vh, cudaStream_t st);
int blas_grid();

// field kernels (tmq_fields.cu)
cudaError_t gauge_reorder(int prec, int recon, void *dst, const double *src_mu, int mu, int Vh, cudaStream_t st);
cudaError_t spinor_from_qkxtm(int prec, void *even, void *odd, const void *qk, int qprec, const Geom &g, cudaStream_t st);
cudaError_t spinor_to_qkxtm(void *qk, int qprec, int prec, const void *even, const void *odd, double scale,
                            const Geom &g, cudaStream_t st);
cudaError_t spinor_from_host_eo(int prec, void *dst, const double *d_aos, int Vh, cudaStream_t st);
cudaError_t spinor_to_host_eo(double *d_aos, int prec, const void *src, int Vh, cudaStream_t st);
cudaError_t spinor_from_host_lex(int prec, void *even, void *odd, const double *d_aos, const Geom &g, cudaStream_t st);
cudaError_t spinor_to_host_lex(double *d_aos, int prec, const void *even, const void *odd, double scale, const Geom &g, cudaStream_t st);
cudaError_t plaquette_launch(int recon, const void *gauge_d, const Geom &g, const BlasRed &r, cudaStream_t st);
cudaError_t qkxtm_plaquette(const void *gq, int prec, const Geom &g, const BlasRed &r, cudaStream_t st);
// ghost zones of the QKXTM containers: site offsets (in units of sites: multiply by the number of components) of the plus / minus ghost
// of every partitioned dimension behind the local volume, and the face sizes
struct QkGhost { size_t plus[4], minus[4], surf[4], total_sites; };
QkGhost qk_ghost_layout(const Geom &g);
cudaError_t qkxtm_face_gather(void *lo, void *hi, const void *d, int prec, const Geom &g, int dim, int ncomp, cudaStream_t st);
cudaError_t qkxtm_scale(void *d, int prec, double a, size_t ncplx, cudaStream_t st);
cudaError_t qkxtm_cast(void *dst, int dprec, const void *src, int sprec, size_t ncplx, cudaStream_t st);
cudaError_t qkxtm_gamma5(void *d, int prec, int V, cudaStream_t st);
// halo pack: project (and for the forward-going face multiply by U^dag) the boundary slices of `in`
cudaError_t halo_pack(int prec, int recon, const DslashArgs<double> *Ad, const DslashArgs<float> *As, int dim,
                      void *send_bwd, void *send_fwd, cudaStream_t st);

cudaError_t halo_pack_p2p(int recon, const DslashArgs<double> &A, const PackDst<double> &D, cudaStream_t st);
cudaError_t halo_pack_p2p(int recon, const DslashArgs<float> &A, const PackDst<float> &D, cudaStream_t st);
cudaError_t p2p_allreduce(const P2PRed &R, cudaStream_t st);
cudaError_t cg_update_pack(int recon, void *x, void *p, const void *r, const double *scal, int an, int ad, int bn, int bd, const DslashArgs<double> &A, cudaStream_t st);
cudaError_t cg_update_pack(int recon, void *x, void *p, const void *r, const double *scal, int an, int ad, int bn, int bd, const DslashArgs<float> &A, cudaStream_t st);
// x += alpha p ; p = r + beta p with device-resident scalars; in the fused halo mode the same launch sends the faces of the new p
int cg_update(tmq_ctx *c, int prec, void *x, void *p, const void *r, int an, int ad, int bn, int bd);

HaloArena halo_arena_layout(const Geom &g);
inline size_t arena_flag_off(const HaloArena &L, int buf, int d, int dir) { return L.flag + sizeof(unsigned int) * (size_t)((buf * 4 + d) * 2 + dir); }

// comm layer (tmq_comm.cpp)
int comm_setup_p2p(tmq_ctx *c);
int comm_unique_id(char id128[128]);
int comm_init(tmq_ctx *c, const char id128[128], int nranks, int rank);
void comm_destroy(tmq_ctx *c);
int comm_exchange(tmq_ctx *c, int pi, int prec, cudaStream_t st);
int comm_allreduce(tmq_ctx *c, double *d_ptr, int n, cudaStream_t st);
int comm_barrier(tmq_ctx *c);
int comm_sendrecv_dim(tmq_ctx *c, int dim, const void *send_bwd, const void *send_fwd, void *recv_from_fwd, void *recv_from_bwd,
                      size_t nbytes, cudaStream_t st);

}  // namespace tmq

// ---- guarded device allocations (TMQ_GUARD_BYTES=n, n a multiple of 256) ----------------------------------------------------------------
// compute-sanitizer is not available on the GPU pool, so the library carries its own out-of-bounds net: with the environment variable set
// every device allocation of the library gets a red zone of n bytes on either side, filled with 0xFF -- a NaN pattern in fp32 and fp64, so
// an out-of-bounds READ poisons the result and fails the parity checks, and tmq_guard_check() finds every out-of-bounds WRITE.  Off by
// default (zero overhead: a plain cudaMalloc).  Every translation unit of the library allocates through these two.
namespace tmq {
cudaError_t guard_malloc(void **p, size_t n);
cudaError_t guard_free(void *p);
}
#define cudaMalloc(p, n) tmq::guard_malloc((void **)(p), (n))
#define cudaFree(p) tmq::guard_free((void *)(p))
