// tmq_types.h -- plain-data types shared by the sm_100a kernels, their launchers and the host
// emulation harness (tests only).  No torch, no QUDA types.
//
// Native device layout (B200-first; differs from QUDA's FLOAT2/FLOAT4 orders):
//   * a "vector" is two complex numbers: fp64 -> 32 bytes (one LDG.E.256 on sm_100a),
//     fp32 -> 16 bytes (one LDG.E.128).
//   * parity spinor : vec[6][stride]   (complex k = 3*spin + colour; vector j holds k = 2j, 2j+1)
//   * full spinor   : parity-0 block followed by parity-1 block
//   * gauge, recon-12: vec[2 parity][4 mu][3][stride] holding rows 0,1 of the 3x3 link
//     gauge, recon-18: cplx[2 parity][4 mu][9][stride]
//   * site index inside a parity block: checkerboard-lexicographic, cb = (x + X(y + Y(z + Z t)))/2,
//     the same order the reference's upload kernel writes (lib/code_pieces/uploadToCuda_core.h:7-28)
//   * ghost half-spinor faces: vec[3][face_sites] (6 complex = 2 spin x 3 colour)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TMQ_HD __host__ __device__ __forceinline__
#define TMQ_D __device__ __forceinline__
#else
#define TMQ_HD inline
#define TMQ_D inline
#endif

namespace tmq {

template <typename F> struct VecT;
template <> struct alignas(32) VecT<double> { double a, b, c, d; };  // (re0, im0, re1, im1)
template <> struct alignas(16) VecT<float> { float a, b, c, d; };
template <typename F> struct CplxT;
template <> struct alignas(16) CplxT<double> { double re, im; };
template <> struct alignas(8) CplxT<float> { float re, im; };

// exact unsigned division by a runtime constant (n < 2^31): q = (mulhi(n, m) + n) >> s
struct FastDiv {
  uint32_t d, m, s;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  uint32_t s = 0;
  while ((1ull << s) < d) s++;
  f.s = s;
  f.m = (uint32_t)(((1ull << 32) * ((1ull << s) - d)) / d + 1);
  return f;
}
TMQ_HD uint32_t fd_div(uint32_t n, const FastDiv &f) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)(((uint64_t)__umulhi(n, f.m) + n) >> f.s);
#else
  return (uint32_t)(((((uint64_t)n * f.m) >> 32) + n) >> f.s);
#endif
}

enum { TMQ_MODE_ALL = 0, TMQ_MODE_INTERIOR = 1, TMQ_MODE_BOUNDARY = 2 };

// Local lattice geometry.
struct Geom {
  int X[4];          // local extents
  int Xh;            // X[0]/2
  int Vh;            // local parity volume
  int part[4];       // 1 if the dimension is partitioned across ranks (ghost zone in use)
  int face[4];       // parity sites on a face orthogonal to dim d
  int tb_first;      // 1 if this rank holds global t = 0        (recon-12 boundary sign on backward t links)
  int tb_last;       // 1 if this rank holds global t = T-1       (recon-12 boundary sign on forward t links)
  int tb_sign;       // +1 periodic, -1 anti-periodic folded into U_t(T-1) (qkxtm/QKXTM_util.cpp:698-705)
};

// Site enumeration used to map threads to sites of the sub-box lo <= (y,z,t) < lo+ext (x is always
// the full row).  Order (fastest first): xh, ly<Ty, lz<Tz, lt<Tt, tile_y, tile_z, tile_t.  A CTA takes a
// contiguous chunk of the enumeration, so its sites form a compact 4-d block whose neighbour spinors
// are shared through L1.  Tile extents must divide ext.  The whole lattice is lo=0, ext=X[1..3];
// interior / boundary slabs of a sharded lattice are further sub-boxes.
struct Enum {
  int lo[3];                              // y, z, t origin
  int step[3];                            // coordinate stride per enumerated index (1, or L-1 to visit the two
                                          // boundary slices 0 and L-1 of a partitioned dimension in one launch)
  int nsites;                             // Xh * ext_y * ext_z * ext_t
  FastDiv dXh, dTy, dTz, dTt, dNy, dNz;   // divisors for the decode chain
};

// Epilogue applied to the hop result h (all stages optional, selected by template flags):
//   t = c1 (h + i a1 g5 h)                 post-hop twist (A^-1 or A^-dag)
//   y = cx (x + i ax g5 x) + k t           x term (plain or twisted)
//   z = c3 (y + i a3 g5 y)                 post twist
// reductions: RED=1 -> sum |y|^2 ; RED=2 -> r -= alpha z, sum |r|^2 (z not stored)
// destinations of one fused pack launch: per partitioned-dimension slot, the neighbours' ghost buffers and flags
template <typename F> struct PackDst {
  VecT<F> *dst[2][2];            // [slot][0 = backward-going face -> rank-1, 1 = forward-going face -> rank+1]
  unsigned int *flag[2][2];
  int dim[2];                    // lattice dimension of each slot (2 = z, 3 = t)
  int nslot;
  unsigned int seq;
  unsigned int *ticket;
};

// scalar all-reduce over peer memory (one tiny launch): every rank stores its partial into every peer's mailbox,
// publishes a sequence number, waits for all peers and sums in rank order (bit-identical result on every rank)
enum { TMQ_MAX_RANKS = 16 };
struct P2PRed {
  double *scal;                        // own device scalar block
  int slot, n;                         // scal[slot .. slot+n) is reduced in place (n <= 4)
  int rank, nranks;
  unsigned int seq;
  double *mbox[TMQ_MAX_RANKS];         // mbox[r]: base of rank r's mailbox [2 buf][4][TMQ_MAX_RANKS] (own for r == rank)
  unsigned int *mflag[TMQ_MAX_RANKS];  // mflag[r]: rank r's flags [2 buf][TMQ_MAX_RANKS]
  double *err;
  unsigned long long timeout_ns;       // wall-clock limit of the wait for the other ranks' contributions
  int cg_iter;                         // as DslashArgs::cg_iter
  int cg_stop;                         // 1: slot holds |r|^2 of CG iteration cg_iter: apply the stopping test to the global sum
};

template <typename F> struct Epi {
  F c1, a1, k, cx, ax, c3, a3;
  F d1, d2, d3;   // EPI_CHEB: out = d1 z + d2 y + d3 r  (three-term Chebyshev recurrence)
};

enum {
  EPI_PLAIN = 0,        // out = h
  EPI_TW = 1,           // out = t
  EPI_TW_XPAY = 2,      // out = x + k t
  EPI_XPAY = 3,         // out = x + k h
  EPI_XPAY_TW3 = 4,     // out = c3 (1 + i a3 g5)(x + k h)
  EPI_MDAGM2 = 5,       // out = c3 (1 + i a3 g5)(x + k t), reduce |x + k t|^2
  EPI_TWX_XPAY = 6,     // out = cx (1 + i ax g5) x + k h
  EPI_CG4 = 7,          // z = cx (1 + i ax g5) x + k h ; r -= alpha z ; reduce |r|^2
  EPI_CHEB = 8,         // z = cx (1 + i ax g5) x + k h ; out = d1 z + d2 y + d3 r   (r read-only, out may alias r)
  EPI_COUNT = 9
};

// Arrival flags the boundary CTAs of a fused sharded launch wait on (peer-memory halo path): the neighbour's pack
// kernel stores the face straight into this rank's ghost buffer over NVLink and then publishes `seq` in the flag.
struct HaloWait {
  const unsigned int *flag[4];   // up to 2 partitioned dims x 2 directions
  int n;                         // 0: no waiting (NCCL path / interior-only launch)
  unsigned int seq;
  int exact;                     // 1: wait for flag == seq (copy-engine path, sequence numbers modulo the table size)
  double *err;                   // device scalar set to 1 if a wait times out (never hang the GPU)
  unsigned long long timeout_ns; // wall-clock limit of one wait (%globaltimer), TMQ_OPT_HALO_TIMEOUT_MS
};

template <typename F> struct DslashArgs {
  Geom g;
  Enum en;                 // segment 0: whole lattice, or the interior of a sharded lattice
  Enum en_b[2];            // fused sharded launch: boundary segments ({0,T-1} slices, {0,Z-1} slices)
  int nblk[3];             // CTAs per segment; gridDim.x = nblk[0] + nblk[1] + nblk[2]
  int npre;                // interior CTAs scheduled BEFORE the boundary CTAs (the rest come after them)
  int all_boundary;        // 1: every CTA of this launch works on boundary sites (separate-launch NCCL path)
  HaloWait hw;
  VecT<F> *out;
  const VecT<F> *in;
  const VecT<F> *x;        // x term (may alias nothing else)
  VecT<F> *r;              // EPI_CG4: residual updated in place; EPI_CHEB: T_{n-1} term (read only)
  const VecT<F> *y;        // EPI_CHEB: T_n term
  const void *gauge;       // base of the [2][4][..][stride] array
  int parity;              // parity of the output sites
  F dsign;                 // +1: D, -1: D^dagger (projector signs)
  Epi<F> e;
  // ghost faces (multi-GPU): [dim][0 = from backward neighbour (used by backward hop), 1 = from forward]
  const VecT<F> *ghost[4][2];
  // reductions
  double *partials;        // [gridDim.x]
  unsigned int *ticket;
  double *scal;            // device scalar block (see tmq_blas.cuh)
  int red_slot;            // where the finished sum goes
  int red_accum;           // 1: add to the slot (second and later launches of a split application)
  int alpha_num, alpha_den; // EPI_CG4: alpha = scal[alpha_num] / scal[alpha_den]
  // twisted-clover: site matrices of the OUTPUT parity in the chiral basis chi^(+-)_s = psi_s +- psi_{s+2},
  // two 6x6 complex blocks per site, vec[36][stride] (complex k = block*36 + row*6 + col; vector k/2)
  const VecT<F> *cl_inv;   // (C + i a g5)^-1 ; its conjugate transpose is (C - i a g5)^-1
  const VecT<F> *cl_c;     // C = 1 + i csw kappa sum_{mu<nu} sigma_munu F_munu
  int cl_dag1, cl_dag3;    // apply the conjugate transpose in the post-hop (t1) / final (t3) position
  VecT<F> *out2;           // clover EPI_MDAGM2: also store y = x + k t (= M p) here, so that the M^dag step can take it as a
                           // plain x term instead of re-applying A^dag = C - i a g5 to w (192 B/site instead of 1152)
  int cl_plain_x;          // clover: the x term is used as it is (no site matrix) although the epilogue has TWX
  int prefetch;            // unused (the L2-prefetch experiment was removed: no gain, and it cost the 4th resident CTA)
  // fused compute + halo exchange (TMQ_OPT_HALO_P2P = 3): the boundary CTAs of THIS launch pack the faces of its output for the NEXT
  // application and store them into the neighbours' ghost arenas; the last boundary CTA publishes pk.seq in the neighbours' flags
  int cg_iter;             // > 0: this launch belongs to CG iteration cg_iter (1-based) of a solve whose host loop runs one iteration
                           // ahead of the residual read-back; it exits at once when an earlier iteration has converged (SC_DONE)
  int cg_local_stop;       // 1: the reduction of this launch is the global |r|^2 (single rank): apply the stopping test on the device
  int pk_on;
  F pk_dsign;              // projector sign of the next application (+1: D, -1: D^dagger)
  PackDst<F> pk;
};

}  // namespace tmq
