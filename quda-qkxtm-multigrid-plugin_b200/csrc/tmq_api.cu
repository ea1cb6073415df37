// tmq_api.cu -- the C ABI of libtmq.so (include/tmq.h): context, resident gauge field, spinor handles, the
// even-odd twisted-mass operator composed from the fused Dslash kernels, and the CG on M^dag M.
//
// Operator algebra (kappa normalisation, a = 2 kappa mu, A = 1 + i a g5):
//   M_sym      = 1 - k^2 A^-1 D A^-1 D            M_sym^dag  = 1 - k^2 D^dag A^-dag D^dag A^-dag
//   M_asym     = A - k^2 D A^-1 D                 M_asym^dag = A^dag - k^2 D^dag A^-dag D^dag
// A CG iteration on M_sym^dag M_sym is four Dslash launches and one update launch:
//   K1  t = A^-1 D p                                           (EPI_TW)
//   K2  w = A^-dag (p - k^2 A^-1 D t),  <p,Ap> = |M p|^2        (EPI_MDAGM2: M p never touches HBM)
//   K3  u = A^-dag D^dag w                                      (EPI_TW, dagger)
//   K4  z = A^dag w - k^2 D^dag u ;  r -= alpha z ; |r|^2       (EPI_CG4: A p never touches HBM)
//   U   x += alpha p ; p = r + beta p                           (blas_cg_update, scalars stay on the device)
// i.e. 16 spinor streams + 4 gauge sweeps per iteration instead of the 20 + 4 of an unfused CG.
// There is no CPU path anywhere in this file: without a CUDA device every entry point fails.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <map>
#include "../../include/tmq.h"
#include "tmq_internal.h"

namespace tmq {

static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- guarded allocations (see tmq_internal.h) ----
#undef cudaMalloc
#undef cudaFree
struct GuardRec { char *base; size_t n; };
static std::map<void *, GuardRec> g_guard;
static long long g_guard_bytes = -1;
static size_t guard_size() {
  if (g_guard_bytes < 0) {
    const char *e = getenv("TMQ_GUARD_BYTES");
    long long v = e ? atoll(e) : 0;
    g_guard_bytes = v > 0 ? ((v + 255) / 256) * 256 : 0;
  }
  return (size_t)g_guard_bytes;
}
cudaError_t guard_malloc(void **p, size_t n) {
  const size_t G = guard_size();
  if (G == 0) return cudaMalloc(p, n);
  char *base = nullptr;
  const size_t body = ((n + 255) / 256) * 256;
  cudaError_t e = cudaMalloc((void **)&base, body + 2 * G);
  if (e != cudaSuccess) return e;
  e = cudaMemset(base, 0xFF, body + 2 * G);          // the body too: reading memory the library never wrote shows up as NaN
  if (e != cudaSuccess) return e;
  g_guard[base + G] = GuardRec{base, n};
  *p = base + G;
  return cudaSuccess;
}
cudaError_t guard_free(void *p) {
  if (guard_size() == 0 || p == nullptr) return cudaFree(p);
  auto it = g_guard.find(p);
  if (it == g_guard.end()) return cudaFree(p);
  char *base = it->second.base;
  g_guard.erase(it);
  return cudaFree(base);
}
// number of live allocations whose red zones were written to (0 = clean); the first offender is described in tmq_last_error()
static int guard_check_all() {
  const size_t G = guard_size();
  if (G == 0) return 0;
  cudaDeviceSynchronize();
  std::vector<unsigned char> h(G);
  int bad = 0;
  for (auto &kv : g_guard) {
    const size_t body = ((kv.second.n + 255) / 256) * 256;
    for (int side = 0; side < 2; side++) {
      const char *zone = side == 0 ? kv.second.base : kv.second.base + G + body;
      if (cudaMemcpy(h.data(), zone, G, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
      size_t first = G;
      for (size_t i = 0; i < G; i++) if (h[i] != 0xFF) { first = i; break; }
      // the slack between n and the 256-byte rounded body belongs to the tail zone as well
      if (first == G && side == 1 && body > kv.second.n) {
        std::vector<unsigned char> t(body - kv.second.n);
        if (cudaMemcpy(t.data(), kv.second.base + G + kv.second.n, t.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        for (size_t i = 0; i < t.size(); i++) if (t[i] != 0xFF) { first = 0; break; }
      }
      if (first != G) {
        if (bad == 0) set_error("out-of-bounds write %s an allocation of %zu bytes (offset %zu into the red zone)", side == 0 ? "before" : "behind", kv.second.n, first);
        bad++;
      }
    }
  }
  return bad;
}

#define cudaMalloc(p, n) tmq::guard_malloc((void **)(p), (n))
#define cudaFree(p) tmq::guard_free((void *)(p))

static int largest_divisor_le(int n, int pref) {
  if (pref < 1) pref = 1;
  for (int d = pref < n ? pref : n; d >= 1; d--)
    if (n % d == 0) return d;
  return 1;
}

Enum make_enum(const Geom &g, const int lo[3], const int ext[3], const int tile_pref[3], const int step[3]) {
  Enum en;
  const int Ty = largest_divisor_le(ext[0], tile_pref[0]);
  const int Tz = largest_divisor_le(ext[1], tile_pref[1]);
  const int Tt = largest_divisor_le(ext[2], tile_pref[2]);
  for (int i = 0; i < 3; i++) { en.lo[i] = lo[i]; en.step[i] = step ? step[i] : 1; }
  en.nsites = g.Xh * ext[0] * ext[1] * ext[2];
  en.dXh = make_fastdiv((uint32_t)g.Xh);
  en.dTy = make_fastdiv((uint32_t)Ty);
  en.dTz = make_fastdiv((uint32_t)Tz);
  en.dTt = make_fastdiv((uint32_t)Tt);
  en.dNy = make_fastdiv((uint32_t)(ext[0] / Ty));
  en.dNz = make_fastdiv((uint32_t)(ext[1] / Tz));
  return en;
}

static inline int pidx(int prec) { return prec == 8 ? 0 : 1; }
static const unsigned int SEQ_TABLE = 4096;

HaloArena halo_arena_layout(const Geom &g) {
  HaloArena L;
  size_t off = 0;
  for (int buf = 0; buf < 2; buf++)
    for (int pi = 0; pi < 2; pi++)
      for (int d = 0; d < 4; d++)
        for (int dir = 0; dir < 2; dir++) {
          L.recv[buf][pi][d][dir] = off;
          if (d >= 2) off += ((size_t)3 * g.face[d] * vec_bytes(pi == 0 ? 8 : 4) + 255) & ~(size_t)255;
        }
  L.flag = off;
  off += 2 * 4 * 2 * sizeof(unsigned int);
  off = (off + 255) & ~(size_t)255;
  L.mbox = off;
  off += (size_t)2 * 4 * TMQ_MAX_RANKS * sizeof(double);
  L.mflag = off;
  off += (size_t)2 * TMQ_MAX_RANKS * sizeof(unsigned int);
  L.bytes = (off + 255) & ~(size_t)255;
  return L;
}

int ensure_scratch(tmq_ctx *c, int prec, int n) {
  Scratch &s = prec == 8 ? c->scr_d : c->scr_s;
  for (int i = 0; i < n && i < NSCRATCH; i++)
    if (!s.tmp[i]) {
      TMQ_CUDA(cudaMalloc(&s.tmp[i], parity_bytes(c, prec)));
      TMQ_CUDA(cudaMemsetAsync(s.tmp[i], 0, parity_bytes(c, prec), c->stream));
    }
  return 0;
}

static int ensure_stage(tmq_ctx *c, size_t bytes) {
  if (c->stage_bytes >= bytes) return 0;
  if (c->stage) { TMQ_CUDA(cudaStreamSynchronize(c->stream)); TMQ_CUDA(cudaFree(c->stage)); c->stage = nullptr; c->stage_bytes = 0; }
  TMQ_CUDA(cudaMalloc(&c->stage, bytes));
  c->stage_bytes = bytes;
  return 0;
}

template <typename F> static cudaError_t launch_any(tmq_ctx *c, int epi, bool multi, const DslashArgs<F> &A, cudaStream_t st);
template <> cudaError_t launch_any<double>(tmq_ctx *c, int epi, bool multi, const DslashArgs<double> &A, cudaStream_t st) {
  return c->recon == 8 ? launch_dslash_d8(epi, multi, A, st) : (c->recon == 12 ? launch_dslash_d12(epi, multi, A, st) : launch_dslash_d18(epi, multi, A, st));
}
template <> cudaError_t launch_any<float>(tmq_ctx *c, int epi, bool multi, const DslashArgs<float> &A, cudaStream_t st) {
  return c->recon == 8 ? launch_dslash_s8(epi, multi, A, st) : (c->recon == 12 ? launch_dslash_s12(epi, multi, A, st) : launch_dslash_s18(epi, multi, A, st));
}
template <typename F> static cudaError_t pack_any(tmq_ctx *c, const DslashArgs<F> &A, int dim, void *sb, void *sf, cudaStream_t st);
template <> cudaError_t pack_any<double>(tmq_ctx *c, const DslashArgs<double> &A, int dim, void *sb, void *sf, cudaStream_t st) {
  return halo_pack(8, c->recon, &A, nullptr, dim, sb, sf, st);
}
template <> cudaError_t pack_any<float>(tmq_ctx *c, const DslashArgs<float> &A, int dim, void *sb, void *sf, cudaStream_t st) {
  return halo_pack(4, c->recon, nullptr, &A, dim, sb, sf, st);
}

// destinations of the faces of application `sq` in the neighbours' arenas (peer-store modes)
template <typename F> static void fill_pack_dst(tmq_ctx *c, PackDst<F> &D, unsigned int sq) {
  const int pi = sizeof(F) == 8 ? 0 : 1;
  const HaloArena &L = c->arena_layout;
  memset(&D, 0, sizeof(D));
  const int b2 = (int)(sq & 1u);
  D.seq = sq; D.ticket = c->ticket2;
  for (int d = 2; d < 4; d++) {
    if (!c->g.part[d]) continue;
    const int sl = D.nslot++;
    D.dim[sl] = d;
    // slice 0 is the "from forward neighbour" ghost (dir 1) of rank-1; slice L-1 the dir-0 ghost of rank+1
    D.dst[sl][0] = (VecT<F> *)(c->peer_arena[d][0] + L.recv[b2][pi][d][1]);
    D.flag[sl][0] = (unsigned int *)(c->peer_arena[d][0] + arena_flag_off(L, b2, d, 1));
    D.dst[sl][1] = (VecT<F> *)(c->peer_arena[d][1] + L.recv[b2][pi][d][0]);
    D.flag[sl][1] = (unsigned int *)(c->peer_arena[d][1] + arena_flag_off(L, b2, d, 0));
  }
}

// halo mode 4: the producer packs into THIS rank's send-buffer set (sq & 1); no flags, no ticket -- the copy engines publish
template <typename F> static void fill_pack_local(tmq_ctx *c, PackDst<F> &D, unsigned int sq) {
  const int pi = sizeof(F) == 8 ? 0 : 1;
  memset(&D, 0, sizeof(D));
  D.seq = sq;
  for (int d = 2; d < 4; d++) {
    if (!c->g.part[d]) continue;
    const int sl = D.nslot++;
    D.dim[sl] = d;
    for (int dir = 0; dir < 2; dir++) D.dst[sl][dir] = (VecT<F> *)((sq & 1u) ? c->halo_send2[pi][d][dir] : c->halo_send[pi][d][dir]);
  }
}

template <typename F>
static int apply_hop_t(tmq_ctx *c, void *out, const void *in, const HopSpec &s) {
  const int prec = (int)sizeof(F);
  const int pi = pidx(prec);
  const GaugeStore &gs = prec == 8 ? c->gauge_d : c->gauge_s;
  TMQ_REQUIRE(gs.d != nullptr, "no gauge field loaded (tmq_gauge_load)");
  DslashArgs<F> A;
  memset(&A, 0, sizeof(A));
  A.g = c->g;
  A.out = (VecT<F> *)out;
  A.in = (const VecT<F> *)in;
  A.x = (const VecT<F> *)s.x;
  A.r = (VecT<F> *)s.r;
  A.y = (const VecT<F> *)s.y;
  A.e.d1 = (F)s.d1; A.e.d2 = (F)s.d2; A.e.d3 = (F)s.d3;
  A.gauge = gs.d;
  A.parity = s.out_parity;
  A.dsign = s.dagger ? (F)-1 : (F)1;
  A.e.c1 = (F)s.t1.c; A.e.a1 = (F)s.t1.a; A.e.k = (F)s.k;
  A.e.cx = (F)s.tx.c; A.e.ax = (F)s.tx.a; A.e.c3 = (F)s.t3.c; A.e.a3 = (F)s.t3.a;
  for (int d = 0; d < 4; d++) {
    A.ghost[d][0] = (const VecT<F> *)c->halo_recv[pi][d][0];
    A.ghost[d][1] = (const VecT<F> *)c->halo_recv[pi][d][1];
  }
  A.partials = c->partials; A.ticket = c->ticket; A.scal = c->scal;
  A.red_slot = s.red_slot; A.red_accum = 0;
  A.alpha_num = s.alpha_num; A.alpha_den = s.alpha_den;
  A.prefetch = c->opt_prefetch;
  A.cg_iter = c->cg_iter_cur;
  A.cg_local_stop = (c->cg_iter_cur > 0 && c->nranks == 1) ? 1 : 0;
  if (c->clover_on) {
    // site matrices of the OUTPUT parity (every site operator of the epilogues acts on the output site)
    const size_t off = (size_t)s.out_parity * 36 * c->g.Vh;
    A.cl_inv = (const VecT<F> *)(prec == 8 ? c->clov_inv_d.d : c->clov_inv_s.d) + off;
    A.cl_c = (const VecT<F> *)(prec == 8 ? c->clov_c_d.d : c->clov_c_s.d) + off;
    A.cl_dag1 = s.t1.dag; A.cl_dag3 = s.t3.dag;
    A.out2 = (VecT<F> *)s.out2; A.cl_plain_x = s.cl_plain_x;
  }

  const Geom &g = c->g;
  const bool has_red = (s.epi == EPI_MDAGM2 || s.epi == EPI_CG4);
  auto nblocks = [](const Enum &en) { return (en.nsites + 127) / 128; };
  A.nblk[0] = A.nblk[1] = A.nblk[2] = 0;
  A.npre = 0x7fffffff;
  A.hw.n = 0; A.hw.seq = 0; A.hw.err = c->scal + SC_ERR;
  A.hw.timeout_ns = (unsigned long long)c->opt_halo_timeout_ms * 1000000ull;
  if (!c->multi) {
    const int lo[3] = {0, 0, 0}, ext[3] = {g.X[1], g.X[2], g.X[3]};
    A.en = make_enum(g, lo, ext, c->tile);
    A.nblk[0] = nblocks(A.en);
    TMQ_CUDA(launch_any<F>(c, s.epi, false, A, c->stream));
    c->launches++;
    return 0;
  }
  const int zlo = g.part[2] ? 1 : 0, zhi = g.part[2] ? g.X[2] - 1 : g.X[2];   // interior range [lo, hi)
  const int tlo = g.part[3] ? 1 : 0, thi = g.part[3] ? g.X[3] - 1 : g.X[3];
  // a box of the (y,z,t) index space; nz / nt enumerated z / t values spaced by sz / st
  auto box = [&](int z0, int nz, int sz, int t0, int nt, int st) -> Enum {
    const int lo[3] = {0, z0, t0}, ext[3] = {g.X[1], nz > 0 ? nz : 0, nt > 0 ? nt : 0}, step[3] = {1, sz, st};
    if (nz <= 0 || nt <= 0) { Enum e; memset(&e, 0, sizeof(e)); e.dXh = e.dTy = e.dTz = e.dTt = e.dNy = e.dNz = make_fastdiv(1); return e; }
    return make_enum(g, lo, ext, c->tile, step);
  };
  const Enum en_int = box(zlo, zhi - zlo, 1, tlo, thi - tlo, 1);                      // no ghost needed
  const Enum en_t = g.part[3] ? box(0, g.X[2], 1, 0, 2, g.X[3] - 1) : box(0, 0, 1, 0, 0, 1);      // t slices {0, T-1}
  const Enum en_z = g.part[2] ? box(0, 2, g.X[2] - 1, tlo, thi - tlo, 1) : box(0, 0, 1, 0, 0, 1);  // z slices {0, Z-1}, interior t

  if (c->p2p && c->opt_p2p == 2) {
    // ---- peer-memory path, copy-engine variant: pack into local send buffers (HBM speed), then the DMA engines
    //      push faces + arrival flags into the neighbours' ghost arenas over NVLink while ONE Dslash launch runs
    //      (interior CTAs first; boundary CTAs wait on the flags).  No SM is needed for the transfer, so spinning
    //      boundary CTAs can never starve it.
    const unsigned int seq = ++c->halo_seq;
    const int buf = (int)(seq & 1u);
    const HaloArena &L = c->arena_layout;
    if (c->opt_pack_async) {
      // pack on the (high-priority) exchange stream as well: it only needs the input field, so the Dslash launch below
      // no longer queues behind it -- the interior CTAs start at once and the pack + copies run beside them.  In-order
      // execution on the exchange stream also protects the send buffers from the previous application's copies.
      TMQ_CUDA(cudaEventRecord(c->ev_pack, c->stream));
      TMQ_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_pack, 0));
      for (int d = 2; d < 4; d++)
        if (g.part[d]) {
          TMQ_CUDA(pack_any<F>(c, A, d, c->halo_send[pi][d][0], c->halo_send[pi][d][1], c->comm_stream));
          c->launches++;
        }
    } else {
      TMQ_CUDA(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));     // our previous outgoing copies have left the send buffers
      for (int d = 2; d < 4; d++)
        if (g.part[d]) {
          TMQ_CUDA(pack_any<F>(c, A, d, c->halo_send[pi][d][0], c->halo_send[pi][d][1], c->stream));
          c->launches++;
        }
      TMQ_CUDA(cudaEventRecord(c->ev_pack, c->stream));
      TMQ_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_pack, 0));
    }
    for (int d = 2; d < 4; d++) {
      if (!g.part[d]) continue;
      const size_t nbytes = (size_t)3 * g.face[d] * vec_bytes(prec);
      for (int dir = 0; dir < 2; dir++) {
        // send[dir 0] = slice 0 -> rank-1, where it is the "from forward neighbour" ghost (1); send[1] -> rank+1, ghost 0
        char *peer = c->peer_arena[d][dir];
        TMQ_CUDA(cudaMemcpyAsync(peer + L.recv[buf][pi][d][1 - dir], c->halo_send[pi][d][dir], nbytes, cudaMemcpyDeviceToDevice, c->comm_stream));
        TMQ_CUDA(cudaMemcpyAsync(peer + arena_flag_off(L, buf, d, 1 - dir), c->seq_table + (seq & (SEQ_TABLE - 1)), sizeof(unsigned int),
                                 cudaMemcpyDeviceToDevice, c->comm_stream));
        A.ghost[d][dir] = (const VecT<F> *)(c->arena + L.recv[buf][pi][d][dir]);
        A.hw.flag[A.hw.n++] = (const unsigned int *)(c->arena + arena_flag_off(L, buf, d, dir));
      }
    }
    TMQ_CUDA(cudaEventRecord(c->ev_halo, c->comm_stream));
    A.hw.seq = seq & (SEQ_TABLE - 1); A.hw.exact = 1;
    A.en = en_int; A.en_b[0] = en_t; A.en_b[1] = en_z;
    A.nblk[0] = nblocks(en_int); A.nblk[1] = nblocks(en_t); A.nblk[2] = nblocks(en_z);
    A.npre = (int)((long long)A.nblk[0] * c->opt_pre_pct / 100);
    TMQ_CUDA(launch_any<F>(c, s.epi, true, A, c->stream));
    c->launches++;
    if (has_red) TMQ_TRY(comm_allreduce(c, c->scal + s.red_slot, 1, c->stream));
    return 0;
  }

  if (c->p2p && c->opt_p2p == 4) {
    // ---- fused pack + copy-engine push.  As mode 2, but inside a chain of launches the stand-alone pack launch is gone: the boundary
    //      CTAs of the launch that PRODUCES a field project the faces of their output for the next application into this rank's own
    //      send buffers (plain local stores: no peer store, no system fence, nothing published), and the next application only has the
    //      copy engines push those buffers and the arrival flags.  Nothing leaves this rank before the consuming application is issued,
    //      so faces packed ahead for an application that never comes are simply dropped.
    //      The compute stream never waits for the exchange stream.  Two sets of send buffers, set (N & 1) for application N, make that
    //      safe: set (N+1)&1 is written by the boundary CTAs of kernel N only after they have seen the flags of application N from
    //      every neighbour (or by a pack / update launch queued behind kernel N); a neighbour sends N only after ITS kernel N-1 has
    //      completed, whose boundary CTAs waited for our faces of N-1 -- so our copies of N-1, the last readers of that set, are done.
    //      (Launches that exit early after convergence of a lagged CG do not wait: cg_double drains both streams and all ranks then.)
    const unsigned int seq = ++c->halo_seq;
    const int buf = (int)(seq & 1u);
    const HaloArena &L = c->arena_layout;
    const bool sent_ahead = s.accept_ahead && c->prepacked_seq == seq && c->prepacked_in == in && c->prepacked_dagger == s.dagger &&
                            c->prepacked_parity == s.out_parity && c->prepacked_prec == prec;
    c->prepacked_seq = 0;
    void *(*snd)[4][2] = buf ? c->halo_send2 : c->halo_send;
    if (!sent_ahead) {
      for (int d = 2; d < 4; d++)
        if (g.part[d]) {
          TMQ_CUDA(pack_any<F>(c, A, d, snd[pi][d][0], snd[pi][d][1], c->stream));
          c->launches++;
        }
    }
    const int dbg = c->opt_debug;       // timing experiments only (results are wrong): 1 = boundary CTAs do not wait, 2 = no copies
    if (!(dbg & 2)) {
      TMQ_CUDA(cudaEventRecord(c->ev_pack, c->stream));
      TMQ_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_pack, 0));
    }
    for (int d = 2; d < 4; d++) {
      if (!g.part[d]) continue;
      const size_t nbytes = (size_t)3 * g.face[d] * vec_bytes(prec);
      for (int dir = 0; dir < 2; dir++) {
        char *peer = c->peer_arena[d][dir];
        if (!(dbg & 2)) {
        TMQ_CUDA(cudaMemcpyAsync(peer + L.recv[buf][pi][d][1 - dir], snd[pi][d][dir], nbytes, cudaMemcpyDeviceToDevice, c->comm_stream));
        TMQ_CUDA(cudaMemcpyAsync(peer + arena_flag_off(L, buf, d, 1 - dir), c->seq_table + (seq & (SEQ_TABLE - 1)), sizeof(unsigned int),
                                 cudaMemcpyDeviceToDevice, c->comm_stream));
        }
        A.ghost[d][dir] = (const VecT<F> *)(c->arena + L.recv[buf][pi][d][dir]);
        A.hw.flag[A.hw.n++] = (const unsigned int *)(c->arena + arena_flag_off(L, buf, d, dir));
      }
    }
    A.hw.seq = seq & (SEQ_TABLE - 1); A.hw.exact = 1;
    if (dbg & 1) A.hw.n = 0;
    if (s.pack_next && out != nullptr && s.epi != EPI_CG4) {
      fill_pack_local<F>(c, A.pk, seq + 1);
      A.pk_on = 2;
      A.pk_dsign = s.next_dagger ? (F)-1 : (F)1;
      c->prepacked_seq = seq + 1; c->prepacked_in = out; c->prepacked_dagger = s.next_dagger ? 1 : 0;
      c->prepacked_parity = 1 - s.out_parity; c->prepacked_prec = prec;
    }
    A.en = en_int; A.en_b[0] = en_t; A.en_b[1] = en_z;
    A.nblk[0] = nblocks(en_int); A.nblk[1] = nblocks(en_t); A.nblk[2] = nblocks(en_z);
    A.npre = (int)((long long)A.nblk[0] * c->opt_pre_pct / 100);
    TMQ_CUDA(launch_any<F>(c, s.epi, true, A, c->stream));
    c->launches++;
    if (has_red) TMQ_TRY(comm_allreduce(c, c->scal + s.red_slot, 1, c->stream));
    return 0;
  }

  if (c->p2p && c->opt_p2p == 3) {
    // ---- fused compute + halo exchange: ONE launch per application.  Its boundary CTAs wait for this application's faces
    //      (arrival flags in our own arena), compute, and then pack the faces of their OUTPUT for the NEXT application straight into
    //      the neighbours' arenas over NVLink; the last boundary CTA publishes the next sequence number.  The boundary CTAs sit in
    //      the middle of the grid, so the peer stores overlap the remaining interior CTAs.  Only the first application of a chain
    //      (whose input was not produced by a Dslash launch, e.g. the CG search direction) needs the stand-alone pack launch.
    //      Buffer reuse: faces for application N+1 go into buffer (N+1)&1 = (N-1)&1 of the neighbour, which it read in application
    //      N-1; a boundary CTA only packs after it has seen the neighbour's flag N, which the neighbour publishes after ALL its
    //      boundary CTAs of application N-1 are done.
    unsigned int seq = ++c->halo_seq;
    const HaloArena &L = c->arena_layout;
    bool sent_ahead = false, discard = false;
    if (c->prepacked_seq == seq) {
      sent_ahead = s.accept_ahead && c->prepacked_in == in && c->prepacked_dagger == s.dagger && c->prepacked_parity == s.out_parity && c->prepacked_prec == prec;
      if (!sent_ahead) {
        // faces were sent ahead for an application that is not this one (the update of the LAST CG iteration sends the search direction
        // of an iteration that never runs): that sequence number is skipped on every rank alike.  Its flags were published by the
        // neighbours' launches too, and only after they were done with application seq - 1 -- so the stand-alone pack below waits for
        // them before it overwrites the buffers of application seq - 1.
        discard = true;
        seq = ++c->halo_seq;
      }
    }
    const int buf = (int)(seq & 1u);
    if (!sent_ahead) {
      PackDst<F> D;
      fill_pack_dst<F>(c, D, seq);
      if (discard) {
        const int pbuf = (int)((seq - 1) & 1u);
        for (int d = 2; d < 4; d++) {
          if (!g.part[d]) continue;
          for (int dir = 0; dir < 2; dir++) A.hw.flag[A.hw.n++] = (const unsigned int *)(c->arena + arena_flag_off(L, pbuf, d, dir));
        }
        A.hw.seq = seq - 1;
      }
      TMQ_CUDA(halo_pack_p2p(c->recon, A, D, c->stream));
      c->launches++;
      A.hw.n = 0;
    }
    for (int d = 2; d < 4; d++) {
      if (!g.part[d]) continue;
      for (int dir = 0; dir < 2; dir++) {
        A.ghost[d][dir] = (const VecT<F> *)(c->arena + L.recv[buf][pi][d][dir]);
        A.hw.flag[A.hw.n++] = (const unsigned int *)(c->arena + arena_flag_off(L, buf, d, dir));
      }
    }
    A.hw.seq = seq;
    if (s.pack_next && out != nullptr && s.epi != EPI_CG4) {
      fill_pack_dst<F>(c, A.pk, seq + 1);
      A.pk_on = 1;
      A.pk_dsign = s.next_dagger ? (F)-1 : (F)1;
      c->prepacked_seq = seq + 1; c->prepacked_in = out; c->prepacked_dagger = s.next_dagger ? 1 : 0;
      c->prepacked_parity = 1 - s.out_parity; c->prepacked_prec = prec;
    }
    A.en = en_int; A.en_b[0] = en_t; A.en_b[1] = en_z;
    A.nblk[0] = nblocks(en_int); A.nblk[1] = nblocks(en_t); A.nblk[2] = nblocks(en_z);
    A.npre = (int)((long long)A.nblk[0] * c->opt_pre_pct / 100);
    TMQ_CUDA(launch_any<F>(c, s.epi, true, A, c->stream));
    c->launches++;
    if (has_red) TMQ_TRY(comm_allreduce(c, c->scal + s.red_slot, 1, c->stream));
    return 0;
  }

  if (c->p2p) {
    // ---- peer-memory path: faces are stored straight into the neighbours' ghost arenas by ONE pack launch;
    //      ONE Dslash launch computes interior CTAs first and boundary CTAs last, which wait on arrival flags.
    const unsigned int seq = ++c->halo_seq;
    const int buf = (int)(seq & 1u);
    const HaloArena &L = c->arena_layout;
    PackDst<F> D;
    memset(&D, 0, sizeof(D));
    D.seq = seq; D.ticket = c->ticket2;
    for (int d = 2; d < 4; d++) {
      if (!g.part[d]) continue;
      const int sl = D.nslot++;
      D.dim[sl] = d;
      // my slice 0 is the "from forward neighbour" ghost (dir 1) of rank-1; my slice L-1 the dir-0 ghost of rank+1
      D.dst[sl][0] = (VecT<F> *)(c->peer_arena[d][0] + L.recv[buf][pi][d][1]);
      D.flag[sl][0] = (unsigned int *)(c->peer_arena[d][0] + arena_flag_off(L, buf, d, 1));
      D.dst[sl][1] = (VecT<F> *)(c->peer_arena[d][1] + L.recv[buf][pi][d][0]);
      D.flag[sl][1] = (unsigned int *)(c->peer_arena[d][1] + arena_flag_off(L, buf, d, 0));
      for (int dir = 0; dir < 2; dir++) {
        A.ghost[d][dir] = (const VecT<F> *)(c->arena + L.recv[buf][pi][d][dir]);
        A.hw.flag[A.hw.n++] = (const unsigned int *)(c->arena + arena_flag_off(L, buf, d, dir));
      }
    }
    A.hw.seq = seq;
    {
      const int nwait = A.hw.n;
      A.hw.n = 0;                    // the pack launch itself waits for nothing (its A.hw is only used after discarded faces, fused mode)
      TMQ_CUDA(halo_pack_p2p(c->recon, A, D, c->stream));
      A.hw.n = nwait;
    }
    c->launches++;
    A.en = en_int; A.en_b[0] = en_t; A.en_b[1] = en_z;
    A.nblk[0] = nblocks(en_int); A.nblk[1] = nblocks(en_t); A.nblk[2] = nblocks(en_z);
    A.npre = (int)((long long)A.nblk[0] * c->opt_pre_pct / 100);
    TMQ_CUDA(launch_any<F>(c, s.epi, true, A, c->stream));
    c->launches++;
    if (has_red) TMQ_TRY(comm_allreduce(c, c->scal + s.red_slot, 1, c->stream));
    return 0;
  }

  // ---- NCCL path: pack faces -> send/recv on the comm stream, overlapped with the interior launch
  for (int d = 2; d < 4; d++)
    if (g.part[d]) {
      TMQ_CUDA(pack_any<F>(c, A, d, c->halo_send[pi][d][0], c->halo_send[pi][d][1], c->stream));
      c->launches++;
    }
  TMQ_CUDA(cudaEventRecord(c->ev_pack, c->stream));
  TMQ_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_pack, 0));
  TMQ_TRY(comm_exchange(c, pi, prec, c->comm_stream));
  TMQ_CUDA(cudaEventRecord(c->ev_halo, c->comm_stream));
  bool first = true;
  auto launch_seg = [&](const Enum &en, int is_boundary) -> int {
    if (en.nsites <= 0) return 0;
    A.en = en;
    A.all_boundary = is_boundary;
    A.nblk[0] = nblocks(en);
    A.npre = A.nblk[0];
    A.red_accum = (has_red && !first) ? 1 : 0;
    TMQ_CUDA(launch_any<F>(c, s.epi, true, A, c->stream));
    c->launches++;
    first = false;
    return 0;
  };
  TMQ_TRY(launch_seg(en_int, 0));
  TMQ_CUDA(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
  TMQ_TRY(launch_seg(en_t, 1));
  TMQ_TRY(launch_seg(en_z, 1));
  if (has_red) TMQ_TRY(comm_allreduce(c, c->scal + s.red_slot, 1, c->stream));
  return 0;
}

int apply_hop(tmq_ctx *c, int prec, void *out, const void *in, const HopSpec &s) {
  if (c->clover_on) TMQ_TRY(clover_update_inverse(c));
  return prec == 8 ? apply_hop_t<double>(c, out, in, s) : apply_hop_t<float>(c, out, in, s);
}

int site_op(tmq_ctx *c, int prec, void *out, const void *in, int parity, const Tw &w) {
  if (!c->clover_on) {
    TMQ_CUDA(blas_twist(prec, out, in, w.c, w.a, c->g.Vh, c->stream)); c->launches++;
    return 0;
  }
  TMQ_TRY(clover_update_inverse(c));
  const size_t off = (size_t)parity * 36 * c->g.Vh * vec_bytes(prec);
  const char *M = (const char *)(w.inv ? (prec == 8 ? c->clov_inv_d.d : c->clov_inv_s.d) : (prec == 8 ? c->clov_c_d.d : c->clov_c_s.d)) + off;
  // A = C + i a g5 (w.a carries the dagger sign); A^-1 is stored whole, its conjugate transpose is A^-dag
  TMQ_CUDA(clover_apply(prec, out, in, M, c->g.Vh, w.inv ? w.dag : 0, w.inv ? 0.0 : w.a, c->stream)); c->launches++;
  return 0;
}

static inline BlasRed red_at(tmq_ctx *c, int slot) { return BlasRed{c->partials, c->ticket, c->scal, slot}; }
static inline size_t nvec(const tmq_ctx *c) { return (size_t)6 * c->g.Vh; }

// reduction helpers: run the blas reduction, all-reduce across ranks, leave the result on the device
int reduce_finish(tmq_ctx *c, int slot, int n) {
  if (c->multi && c->nranks > 1) TMQ_TRY(comm_allreduce(c, c->scal + slot, n, c->stream));
  return 0;
}
// Scalars travel to the host through MAPPED pinned memory, written by a one-warp kernel, not through a copy engine: a cudaMemcpy of 8
// bytes would queue behind whatever bulk download is in flight on the same DMA engine (the 2 GB solution of the previous column of a
// propagator: +40 ms on the first read-back of every solve, measured in profiles/r2_e2e_probe.md).
__global__ void scal_to_host_kernel(double *h, const double *d, int n) {
  if ((int)threadIdx.x < n) h[threadIdx.x] = d[threadIdx.x];
  __threadfence_system();
}
int scal_to_host(tmq_ctx *c, int slot, int n) {
  scal_to_host_kernel<<<1, 32, 0, c->stream>>>(c->h_scal_dev + slot, c->scal + slot, n);
  TMQ_CUDA(cudaGetLastError());
  return 0;
}
// The lagged CG's read-back of |r|^2 of one iteration: entry `ring` of the host ring behind the scalar block, and the event the host
// waits on.  (Moving the store to a stream of its own was tried against the slow-down of the iteration next to bulk PCIe copies on
// 8 GPUs and changed nothing, profiles/r2_e2e_n8.md; it stays on the compute stream.)
static int scal_to_host_ring(tmq_ctx *c, int slot, int ring) {
  scal_to_host_kernel<<<1, 32, 0, c->stream>>>(c->h_scal_dev + SC_COUNT + ring, c->scal + slot, 1);
  TMQ_CUDA(cudaGetLastError());
  TMQ_CUDA(cudaEventRecord(c->ev_ring[ring], c->stream));
  return 0;
}
int fetch_scal(tmq_ctx *c, int slot, int n, double *out) {
  TMQ_TRY(scal_to_host(c, slot, n));
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < n; i++) out[i] = c->h_scal[slot + i];
  return 0;
}

// the Dslash kernels raise scal[SC_ERR] when a halo wait timed out (a neighbour never delivered its face)
int check_device_error(tmq_ctx *c) {
  if (!c->multi) return 0;
  double e = 0;
  TMQ_TRY(fetch_scal(c, SC_ERR, 1, &e));
  if (e != 0.0) {
    const double zero = 0.0;
    cudaMemcpyAsync(c->scal + SC_ERR, &zero, sizeof(double), cudaMemcpyHostToDevice, c->stream);
    cudaStreamSynchronize(c->stream);
    set_error("halo exchange timed out: a neighbour rank did not deliver its ghost face");
    return 1;
  }
  return 0;
}

// x += alpha p ; p = r + beta p (alpha = scal[an]/scal[ad], beta = scal[bn]/scal[bd]).  Fused halo mode: the same launch packs the
// new p for the first Dslash launch of the next iteration (K1: D, output parity q) and sends it to the neighbours.
template <typename F> static int cg_update_fused(tmq_ctx *c, void *x, void *p, const void *r, int an, int ad, int bn, int bd) {
  const int prec = (int)sizeof(F);
  const GaugeStore &gs = prec == 8 ? c->gauge_d : c->gauge_s;
  DslashArgs<F> A;
  memset(&A, 0, sizeof(A));
  A.g = c->g;
  A.gauge = gs.d;
  A.parity = c->matpc & 1;                           // p lives on the parity the preconditioned operator acts on
  const unsigned int next = c->halo_seq + 1;
  if (c->opt_p2p == 4) {
    fill_pack_local<F>(c, A.pk, next);
    A.pk_on = 2;
  } else {
    fill_pack_dst<F>(c, A.pk, next);
    A.pk_on = 1;
  }
  A.pk_dsign = (F)1;
  A.cg_iter = c->cg_iter_cur;
  TMQ_CUDA(cg_update_pack(c->recon, x, p, r, c->scal, an, ad, bn, bd, A, c->stream));
  c->launches++;
  c->prepacked_seq = next; c->prepacked_in = p; c->prepacked_dagger = 0; c->prepacked_parity = 1 - (c->matpc & 1); c->prepacked_prec = prec;
  return 0;
}
int cg_update(tmq_ctx *c, int prec, void *x, void *p, const void *r, int an, int ad, int bn, int bd) {
  if (c->multi && c->p2p && c->opt_p2p >= 3 && c->matpc < 2)
    return prec == 8 ? cg_update_fused<double>(c, x, p, r, an, ad, bn, bd) : cg_update_fused<float>(c, x, p, r, an, ad, bn, bd);
  TMQ_CUDA(blas_cg_update(prec, x, p, r, (size_t)6 * c->g.Vh, c->scal, an, ad, bn, bd, c->stream, c->cg_iter_cur));
  c->launches++;
  return 0;
}

// ---- operator compositions on raw parity blocks -----------------------------------------------------------------
// p = parity the preconditioned operator acts on; q = 1 - p
int op_matpc(tmq_ctx *c, int prec, void *out, const void *in, int dagger) {
  TMQ_TRY(ensure_scratch(c, prec, 2));
  const int p = c->matpc & 1, q = 1 - p;
  const bool asym = c->matpc >= 2;
  const double k2 = -c->kappa * c->kappa;
  void *t0 = scr(c, prec, 0), *t1 = scr(c, prec, 1);
  HopSpec a, b;
  if (!asym && !dagger) {
    a.epi = EPI_TW; a.out_parity = q; a.t1 = tw_Ainv(c, 0); a.pack_next = 1; a.next_dagger = 0;
    TMQ_TRY(apply_hop(c, prec, t0, in, a));
    b.epi = EPI_TW_XPAY; b.out_parity = p; b.t1 = tw_Ainv(c, 0); b.k = k2; b.x = in; b.accept_ahead = 1;
    return apply_hop(c, prec, out, t0, b);
  }
  if (!asym && dagger) {
    TMQ_TRY(site_op(c, prec, t1, in, p, tw_Ainv(c, 1)));
    a.epi = EPI_TW; a.out_parity = q; a.dagger = 1; a.t1 = tw_Ainv(c, 1); a.pack_next = 1; a.next_dagger = 1;
    TMQ_TRY(apply_hop(c, prec, t0, t1, a));
    b.epi = EPI_XPAY; b.out_parity = p; b.dagger = 1; b.k = k2; b.x = in; b.accept_ahead = 1;
    return apply_hop(c, prec, out, t0, b);
  }
  // asymmetric, either direction: t = A^-(dag) D(dag) in ; out = A(dag) in - k^2 D(dag) t
  a.epi = EPI_TW; a.out_parity = q; a.dagger = dagger; a.t1 = tw_Ainv(c, dagger); a.pack_next = 1; a.next_dagger = dagger;
  TMQ_TRY(apply_hop(c, prec, t0, in, a));
  b.epi = EPI_TWX_XPAY; b.out_parity = p; b.dagger = dagger; b.tx = tw_A(c, dagger); b.k = k2; b.x = in; b.accept_ahead = 1;
  return apply_hop(c, prec, out, t0, b);
}

// out = M^dag M in.  Symmetric: the four fused launches K1..K4 (without the CG tail); leaves |M in|^2 in
// pap_slot.  Asymmetric: two op_matpc calls through scratch 2.
int op_mdagm(tmq_ctx *c, int prec, void *out, const void *in, int pap_slot) {
  const int p = c->matpc & 1, q = 1 - p;
  const double k2 = -c->kappa * c->kappa;
  if (c->matpc >= 2) {
    TMQ_TRY(ensure_scratch(c, prec, 3));
    TMQ_TRY(op_matpc(c, prec, scr(c, prec, 2), in, 0));
    TMQ_CUDA(blas_norm2(prec, scr(c, prec, 2), nvec(c), red_at(c, pap_slot), c->stream)); c->launches++;
    TMQ_TRY(reduce_finish(c, pap_slot, 1));
    return op_matpc(c, prec, out, scr(c, prec, 2), 1);
  }
  TMQ_TRY(ensure_scratch(c, prec, 2));
  void *t0 = scr(c, prec, 0), *t1 = scr(c, prec, 1);
  HopSpec k1, k2s, k3, k4;
  k1.epi = EPI_TW; k1.out_parity = q; k1.t1 = tw_Ainv(c, 0); k1.pack_next = 1; k1.next_dagger = 0;
  TMQ_TRY(apply_hop(c, prec, t0, in, k1));
  k2s.epi = EPI_MDAGM2; k2s.out_parity = p; k2s.t1 = tw_Ainv(c, 0); k2s.k = k2; k2s.x = in; k2s.t3 = tw_Ainv(c, 1);
  k2s.red_slot = pap_slot; k2s.pack_next = 1; k2s.next_dagger = 1; k2s.accept_ahead = 1;
  // twisted-clover: K2 also stores y = M in, and K4 forms y - k^2 D^dag u from it instead of applying A^dag to w
  void *ybuf = nullptr;
  if (c->clover_on) { TMQ_TRY(ensure_scratch(c, prec, 7)); ybuf = scr(c, prec, 6); k2s.out2 = ybuf; }
  TMQ_TRY(apply_hop(c, prec, t1, t0, k2s));
  k3.epi = EPI_TW; k3.out_parity = q; k3.dagger = 1; k3.t1 = tw_Ainv(c, 1); k3.pack_next = 1; k3.next_dagger = 1; k3.accept_ahead = 1;
  TMQ_TRY(apply_hop(c, prec, t0, t1, k3));
  k4.accept_ahead = 1;
  k4.epi = EPI_TWX_XPAY; k4.out_parity = p; k4.dagger = 1; k4.tx = tw_A(c, 1); k4.k = k2; k4.x = t1;
  if (ybuf) { k4.x = ybuf; k4.cl_plain_x = 1; }
  return apply_hop(c, prec, out, t0, k4);
}

// one fused CG iteration body on M_sym^dag M_sym (no host interaction): K1..K4.
//   reads  r2_old from scal[r2_old], writes <p,Ap> to SC_PAP and the new |r|^2 to scal[r2_new]
static int cg_fused_matvec(tmq_ctx *c, int prec, void *r, const void *p_, int r2_old, int r2_new, bool first) {
  const int p = c->matpc & 1, q = 1 - p;
  const double k2 = -c->kappa * c->kappa;
  void *t0 = scr(c, prec, 0), *t1 = scr(c, prec, 1);
  HopSpec k1, k2s, k3, k4;
  k1.epi = EPI_TW; k1.out_parity = q; k1.t1 = tw_Ainv(c, 0); k1.pack_next = 1; k1.next_dagger = 0;
  k1.accept_ahead = first ? 0 : 1;          // the update of the previous iteration sent the faces of the new search direction
  TMQ_TRY(apply_hop(c, prec, t0, p_, k1));
  k2s.epi = EPI_MDAGM2; k2s.out_parity = p; k2s.t1 = tw_Ainv(c, 0); k2s.k = k2; k2s.x = p_; k2s.t3 = tw_Ainv(c, 1);
  k2s.red_slot = SC_PAP; k2s.pack_next = 1; k2s.next_dagger = 1; k2s.accept_ahead = 1;
  void *ybuf = nullptr;
  if (c->clover_on) { TMQ_TRY(ensure_scratch(c, prec, 7)); ybuf = scr(c, prec, 6); k2s.out2 = ybuf; }
  TMQ_TRY(apply_hop(c, prec, t1, t0, k2s));
  k3.epi = EPI_TW; k3.out_parity = q; k3.dagger = 1; k3.t1 = tw_Ainv(c, 1); k3.pack_next = 1; k3.next_dagger = 1; k3.accept_ahead = 1;
  TMQ_TRY(apply_hop(c, prec, t0, t1, k3));
  k4.accept_ahead = 1;
  k4.epi = EPI_CG4; k4.out_parity = p; k4.dagger = 1; k4.tx = tw_A(c, 1); k4.k = k2; k4.x = t1; k4.r = r;
  if (ybuf) { k4.x = ybuf; k4.cl_plain_x = 1; }
  k4.red_slot = r2_new; k4.alpha_num = r2_old; k4.alpha_den = SC_PAP;
  return apply_hop(c, prec, nullptr, t0, k4);
}

}  // namespace tmq

using namespace tmq;

// =====================================================================================================================
extern "C" { static void io_release(tmq_ctx *c); }

extern "C" {

const char *tmq_last_error(void) { return tmq::g_err; }
int tmq_version(void) { return 100; }
int tmq_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

tmq_ctx *tmq_create(int device, const int localX[4], const int grid[4], const int coord[4]) {
  int ndev = tmq_device_count();
  if (ndev <= 0) { set_error("no usable CUDA device: libtmq has no CPU fallback"); return nullptr; }
  if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return nullptr; }
  for (int d = 0; d < 4; d++) {
    if (localX[d] < 2 || (localX[d] & 1)) { set_error("local extent %d of dimension %d must be even and >= 2", localX[d], d); return nullptr; }
    if (grid[d] < 1 || coord[d] < 0 || coord[d] >= grid[d]) { set_error("bad process grid / coordinate in dimension %d", d); return nullptr; }
  }
  if (grid[0] != 1 || grid[1] != 1) { set_error("only z and t may be partitioned (grid = %d %d %d %d)", grid[0], grid[1], grid[2], grid[3]); return nullptr; }
  const long long V = (long long)localX[0] * localX[1] * localX[2] * localX[3];
  if (V / 2 > 0x7fffffffLL / 32) { set_error("local volume too large for 32-bit site indices"); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(cudaGetLastError())); return nullptr; }

  tmq_ctx *c = new tmq_ctx();
  c->device = device;
  c->stream = nullptr; c->comm_stream = nullptr; memset(c->ev_ring, 0, sizeof(c->ev_ring));
  c->partials = nullptr; c->ticket = nullptr; c->scal = nullptr; c->h_scal = nullptr;
  c->comm = nullptr; c->launches = 0; c->stage = nullptr; c->stage_bytes = 0;
  c->recon = 0; c->t_boundary = 1; c->kappa = 0; c->mu = 0; c->matpc = 0; c->op_set = false;
  memset(c->halo_send, 0, sizeof(c->halo_send));
  memset(c->halo_send2, 0, sizeof(c->halo_send2));
  memset(c->halo_recv, 0, sizeof(c->halo_recv));
  Geom &g = c->g;
  memset(&g, 0, sizeof(g));
  for (int d = 0; d < 4; d++) { g.X[d] = localX[d]; c->grid[d] = grid[d]; c->coord[d] = coord[d]; g.part[d] = grid[d] > 1; }
  g.Xh = localX[0] / 2;
  g.Vh = (int)(V / 2);
  for (int d = 0; d < 4; d++) g.face[d] = g.Vh / g.X[d];
  g.tb_first = coord[3] == 0;
  g.tb_last = coord[3] == grid[3] - 1;
  g.tb_sign = 1;
  c->nranks = grid[0] * grid[1] * grid[2] * grid[3];
  c->rank = ((coord[0] * grid[1] + coord[1]) * grid[2] + coord[2]) * grid[3] + coord[3];
  c->Vglobal = V * c->nranks;
  c->multi = g.part[2] || g.part[3];
  c->tile[0] = 4; c->tile[1] = 4; c->tile[2] = 2;
  c->opt_prefetch = 0;
  c->opt_smear_block_t = 0;
  c->opt_pack_async = 0; c->opt_debug = 0;
  c->opt_halo_timeout_ms = 120000;
  if (const char *e = getenv("TMQ_HALO_TIMEOUT_MS")) { const int v = atoi(e); if (v > 0) c->opt_halo_timeout_ms = v; }
  c->opt_pre_pct = 50; c->red_seq = 0; memset(c->rank_arena, 0, sizeof(c->rank_arena));
  c->opt_p2p = 4; c->p2p = false; c->seq_table = nullptr; c->arena = nullptr; c->halo_seq = 0; c->ticket2 = nullptr;
  if (const char *e = getenv("TMQ_CG_LAG")) { const int v = atoi(e); if (v >= 0 && v <= 6) c->opt_cg_lag = v; }       // as TMQ_OPT_CG_LAG
  if (const char *e = getenv("TMQ_HALO_P2P")) { const int v = atoi(e); if (v >= 0 && v <= 4) c->opt_p2p = v; }   // as TMQ_OPT_HALO_P2P
  memset(c->peer_arena, 0, sizeof(c->peer_arena));

  bool ok = true;
  ok = ok && cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
  {
    // the exchange must not queue behind the interior kernel's CTAs: give its stream the highest priority
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    ok = ok && cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
  }
  ok = ok && cudaEventCreate(&c->ev_a) == cudaSuccess && cudaEventCreate(&c->ev_b) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&c->ev_pack, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&c->ev_r2, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < 8; i++) ok = ok && cudaEventCreateWithFlags(&c->ev_ring[i], cudaEventDisableTiming) == cudaSuccess;
  c->sms = 148;
  cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, device);
  size_t nblk = (size_t)(g.Vh + 127) / 128 + 1;
  if (nblk < (size_t)blas_grid()) nblk = (size_t)blas_grid();
  c->partials_len = nblk * 4;
  ok = ok && cudaMalloc(&c->partials, c->partials_len * sizeof(double)) == cudaSuccess;
  ok = ok && cudaMalloc(&c->ticket, sizeof(unsigned int)) == cudaSuccess;
  ok = ok && cudaMalloc(&c->ticket2, sizeof(unsigned int)) == cudaSuccess;
  ok = ok && cudaMalloc(&c->seq_table, SEQ_TABLE * sizeof(unsigned int)) == cudaSuccess;
  ok = ok && cudaMalloc(&c->scal, SC_COUNT * sizeof(double)) == cudaSuccess;
  ok = ok && cudaHostAlloc((void **)&c->h_scal, (SC_COUNT + 8) * sizeof(double), cudaHostAllocMapped) == cudaSuccess;
  ok = ok && cudaHostGetDevicePointer((void **)&c->h_scal_dev, c->h_scal, 0) == cudaSuccess;
  if (ok) {
    double init[SC_COUNT];
    for (int i = 0; i < SC_COUNT; i++) init[i] = 0.0;
    init[SC_ONE] = 1.0;
    ok = ok && cudaMemset(c->ticket, 0, sizeof(unsigned int)) == cudaSuccess;
    ok = ok && cudaMemset(c->ticket2, 0, sizeof(unsigned int)) == cudaSuccess;
    std::vector<unsigned int> tab(SEQ_TABLE);
    for (unsigned int i = 0; i < SEQ_TABLE; i++) tab[i] = i;
    ok = ok && cudaMemcpy(c->seq_table, tab.data(), SEQ_TABLE * sizeof(unsigned int), cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(c->scal, init, sizeof(init), cudaMemcpyHostToDevice) == cudaSuccess;
  }
  if (!ok) {
    set_error("tmq_create: CUDA resource allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    tmq_destroy(c);
    return nullptr;
  }
  if (c->multi) {
    int part[4] = {0, 0, g.part[2], g.part[3]};
    if (tmq_force_partition(c, part)) { tmq_destroy(c); return nullptr; }
  }
  return c;
}

int tmq_force_partition(tmq_ctx *c, const int part[4]) {
  TMQ_REQUIRE(c, "null context");
  TMQ_REQUIRE(!part[0] && !part[1], "only z and t can carry a ghost zone");
  for (int d = 2; d < 4; d++) {
    if (!part[d] && c->grid[d] > 1) { set_error("dimension %d is split across ranks and cannot be un-partitioned", d); return 1; }
    c->g.part[d] = part[d] ? 1 : 0;
  }
  c->multi = c->g.part[2] || c->g.part[3];
  for (int pi = 0; pi < 2; pi++)
    for (int d = 2; d < 4; d++)
      for (int dir = 0; dir < 2; dir++) {
        if (!c->g.part[d] || c->halo_send[pi][d][dir]) continue;
        const size_t nbytes = (size_t)3 * c->g.face[d] * vec_bytes(pi == 0 ? 8 : 4);
        TMQ_CUDA(cudaMalloc(&c->halo_send[pi][d][dir], nbytes));
        TMQ_CUDA(cudaMalloc(&c->halo_send2[pi][d][dir], nbytes));
        TMQ_CUDA(cudaMalloc(&c->halo_recv[pi][d][dir], nbytes));
      }
  // ghost arena of the peer-memory path; a partitioned dimension on a grid of extent 1 wraps onto this rank, so
  // its "neighbour" arena is our own.  Remote neighbours are mapped by tmq_comm_init (CUDA IPC).
  if (c->multi && !c->arena) {
    c->arena_layout = halo_arena_layout(c->g);
    TMQ_CUDA(cudaMalloc((void **)&c->arena, c->arena_layout.bytes));
    TMQ_CUDA(cudaMemset(c->arena, 0, c->arena_layout.bytes));
  }
  bool all_mapped = c->multi;
  for (int d = 2; d < 4; d++) {
    if (!c->g.part[d]) continue;
    if (c->grid[d] == 1) c->peer_arena[d][0] = c->peer_arena[d][1] = c->arena;
    all_mapped = all_mapped && c->peer_arena[d][0] && c->peer_arena[d][1];
  }
  c->p2p = c->opt_p2p && all_mapped;
  return 0;
}

int tmq_destroy(tmq_ctx *c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
  while (!c->spinors.empty()) tmq_spinor_free(*c->spinors.begin());
  io_release(c);
  comm_destroy(c);
  eig_release(c);
  tmq_clover_free(c);
  tmq_gauge_free(c);
  for (int i = 0; i < NSCRATCH; i++) { if (c->scr_d.tmp[i]) cudaFree(c->scr_d.tmp[i]); if (c->scr_s.tmp[i]) cudaFree(c->scr_s.tmp[i]); }
  for (int pi = 0; pi < 2; pi++)
    for (int d = 0; d < 4; d++)
      for (int dir = 0; dir < 2; dir++) {
        if (c->halo_send[pi][d][dir]) cudaFree(c->halo_send[pi][d][dir]);
        if (c->halo_send2[pi][d][dir]) cudaFree(c->halo_send2[pi][d][dir]);
        if (c->halo_recv[pi][d][dir]) cudaFree(c->halo_recv[pi][d][dir]);
      }
  for (void *p : c->ipc_opened) cudaIpcCloseMemHandle(p);
  if (c->arena) cudaFree(c->arena);
  if (c->ticket2) cudaFree(c->ticket2);
  if (c->seq_table) cudaFree(c->seq_table);
  if (c->stage) cudaFree(c->stage);
  if (c->contract_ws) cudaFree(c->contract_ws);
  if (c->partials) cudaFree(c->partials);
  if (c->ticket) cudaFree(c->ticket);
  if (c->scal) cudaFree(c->scal);
  if (c->h_scal) cudaFreeHost(c->h_scal);
  if (c->ev_a) cudaEventDestroy(c->ev_a);
  if (c->ev_b) cudaEventDestroy(c->ev_b);
  if (c->ev_pack) cudaEventDestroy(c->ev_pack);
  if (c->ev_halo) cudaEventDestroy(c->ev_halo);
  if (c->ev_r2) cudaEventDestroy(c->ev_r2);
  for (int i = 0; i < 8; i++) if (c->ev_ring[i]) cudaEventDestroy(c->ev_ring[i]);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  delete c;
  return 0;
}

int tmq_sync(tmq_ctx *c) {
  TMQ_REQUIRE(c, "null context");
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  TMQ_CUDA(cudaStreamSynchronize(c->comm_stream));
  return check_device_error(c);
}

int tmq_comm_unique_id(char id128[128]) { return comm_unique_id(id128); }
int tmq_comm_init(tmq_ctx *c, const char id128[128], int nranks, int rank) {
  TMQ_REQUIRE(c, "null context");
  TMQ_REQUIRE(nranks == c->nranks && rank == c->rank, "communicator (%d of %d) does not match the process grid (%d of %d)",
              rank, nranks, c->rank, c->nranks);
  TMQ_TRY(comm_init(c, id128, nranks, rank));
  return comm_setup_p2p(c);     // map the neighbours' ghost arenas (CUDA IPC); falls back to NCCL send/recv
}
// in-place sum of a host array over all ranks (staged through the device: NCCL all-reduce over NVLink)
int tmq_allreduce_host(tmq_ctx *c, double *h, size_t n) {
  TMQ_REQUIRE(c && h, "null argument");
  if (c->nranks == 1 || n == 0) return 0;
  TMQ_REQUIRE(n < ((size_t)1 << 31), "array too long");
  TMQ_CUDA(cudaSetDevice(c->device));
  TMQ_TRY(ensure_stage(c, n * sizeof(double)));
  TMQ_CUDA(cudaMemcpyAsync(c->stage, h, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  TMQ_TRY(comm_allreduce(c, (double *)c->stage, (int)n, c->stream));
  TMQ_CUDA(cudaMemcpyAsync(h, c->stage, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
int tmq_barrier(tmq_ctx *c) {
  TMQ_REQUIRE(c, "null context");
  TMQ_CUDA(cudaSetDevice(c->device));
  return comm_barrier(c);
}
int tmq_guard_check(void) { return tmq::guard_check_all(); }
int tmq_halo_mode(tmq_ctx *c) { return c ? (c->multi ? (c->p2p ? 1 + c->opt_p2p : 1) : 0) : -1; }

int tmq_set_tile(tmq_ctx *c, int ty, int tz, int tt) {
  TMQ_REQUIRE(c, "null context");
  if (ty > 0) c->tile[0] = ty;
  if (tz > 0) c->tile[1] = tz;
  if (tt > 0) c->tile[2] = tt;
  return 0;
}

int tmq_set_option(tmq_ctx *c, int option, int value) {
  TMQ_REQUIRE(c, "null context");
  switch (option) {
    case TMQ_OPT_PREFETCH: c->opt_prefetch = value ? 1 : 0; return 0;
    case TMQ_OPT_PACK_ASYNC: c->opt_pack_async = value ? 1 : 0; return 0;
    case 99: c->opt_debug = value; return 0;   // timing experiments (tools/shard_shape_study.py); results are wrong when set
    case TMQ_OPT_CONTRACT_SLICES: c->opt_contract_slices = value < 0 ? 0 : value; return 0;
    case TMQ_OPT_SMEAR_BLOCK_T: c->opt_smear_block_t = value < 0 ? 0 : value; return 0;
    case TMQ_OPT_CG_LAG: c->opt_cg_lag = value < 0 ? 0 : (value > 6 ? 6 : value); return 0;
    case TMQ_OPT_HALO_TIMEOUT_MS: c->opt_halo_timeout_ms = value > 0 ? value : 120000; return 0;
    case TMQ_OPT_BOUNDARY_AT_PCT: c->opt_pre_pct = value < 0 ? 0 : (value > 100 ? 100 : value); return 0;
    case TMQ_OPT_HALO_P2P: {
      c->opt_p2p = value < 0 ? 0 : (value > 4 ? 4 : value);
      c->prepacked_seq = 0;
      bool all_mapped = c->multi;
      for (int d = 2; d < 4; d++)
        if (c->g.part[d]) all_mapped = all_mapped && c->peer_arena[d][0] && c->peer_arena[d][1];
      TMQ_CUDA(cudaStreamSynchronize(c->stream));
      if (c->comm_stream) TMQ_CUDA(cudaStreamSynchronize(c->comm_stream));
      c->p2p = c->opt_p2p && all_mapped;
      return 0;
    }
  }
  set_error("unknown option %d", option);
  return 1;
}

// ---- gauge ----------------------------------------------------------------------------------------------------------
int tmq_gauge_load(tmq_ctx *c, const void *const qdp[4], int t_boundary, int recon) {
  TMQ_REQUIRE(c, "null context");
  TMQ_REQUIRE(recon == 8 || recon == 12 || recon == 18, "reconstruct must be 8, 12 or 18 (got %d)", recon);
  TMQ_REQUIRE(t_boundary == 1 || t_boundary == -1, "t_boundary must be +1 or -1");
  if (recon == 8) {
    // the 8-real format divides by |U01|^2 + |U02|^2: links with a (nearly) vanishing rest of the first row -- a unit field -- cannot be stored
    const size_t nlinks = (size_t)2 * c->g.Vh;
    for (int mu = 0; mu < 4; mu++) {
      TMQ_REQUIRE(qdp[mu], "gauge[%d] is null", mu);
      const double *u = (const double *)qdp[mu];
      for (size_t i = 0; i < nlinks; i++) {
        const double *l = u + i * 18;
        TMQ_REQUIRE(l[2] * l[2] + l[3] * l[3] + l[4] * l[4] + l[5] * l[5] >= 1e-6,
                    "reconstruct 8 cannot store link (mu %d, site %zu): |U01|^2 + |U02|^2 < 1e-6 (unit or near-diagonal field; use 12)", mu, i);
      }
    }
  }
  TMQ_CUDA(cudaSetDevice(c->device));
  tmq_gauge_free(c);
  c->recon = recon;
  c->t_boundary = t_boundary;
  c->g.tb_sign = t_boundary;
  const size_t Vh = c->g.Vh;
  const size_t nreal = (size_t)2 * 4 * recon * Vh;
  c->gauge_d.bytes = nreal * 8;
  c->gauge_s.bytes = nreal * 4;
  TMQ_CUDA(cudaMalloc(&c->gauge_d.d, c->gauge_d.bytes));
  TMQ_CUDA(cudaMalloc(&c->gauge_s.d, c->gauge_s.bytes));
  const size_t mu_bytes = (size_t)2 * Vh * 18 * sizeof(double);
  TMQ_TRY(ensure_stage(c, mu_bytes));
  for (int mu = 0; mu < 4; mu++) {
    TMQ_REQUIRE(qdp[mu], "gauge[%d] is null", mu);
    TMQ_CUDA(cudaMemcpyAsync(c->stage, qdp[mu], mu_bytes, cudaMemcpyHostToDevice, c->stream));
    TMQ_CUDA(gauge_reorder(8, recon, c->gauge_d.d, (const double *)c->stage, mu, (int)Vh, c->stream));
    TMQ_CUDA(gauge_reorder(4, recon, c->gauge_s.d, (const double *)c->stage, mu, (int)Vh, c->stream));
    c->launches += 2;
    TMQ_CUDA(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

int tmq_gauge_free(tmq_ctx *c) {
  if (!c) return 0;
  if (c->gauge_d.d) { cudaFree(c->gauge_d.d); c->gauge_d.d = nullptr; }
  if (c->gauge_s.d) { cudaFree(c->gauge_s.d); c->gauge_s.d = nullptr; }
  return 0;
}

int tmq_plaquette(tmq_ctx *c, double *plaq) {
  TMQ_REQUIRE(c && plaq, "null argument");
  TMQ_REQUIRE(c->gauge_d.d, "no gauge field loaded");
  TMQ_REQUIRE(c->nranks == 1, "tmq_plaquette is a single-rank sanity check");
  TMQ_CUDA(plaquette_launch(c->recon, c->gauge_d.d, c->g, red_at(c, SC_T0), c->stream));
  c->launches++;
  double s;
  TMQ_TRY(fetch_scal(c, SC_T0, 1, &s));
  *plaq = s / ((double)c->Vglobal * 3.0 * 6.0);
  return 0;
}

// ---- spinors --------------------------------------------------------------------------------------------------------
tmq_spinor *tmq_spinor_alloc(tmq_ctx *c, int prec, int subset) {
  if (!c) { set_error("null context"); return nullptr; }
  if ((prec != 8 && prec != 4) || (subset != TMQ_SUBSET_PARITY && subset != TMQ_SUBSET_FULL)) { set_error("bad precision / subset"); return nullptr; }
  tmq_spinor *s = new tmq_spinor();
  s->ctx = c; s->prec = prec; s->subset = subset; s->owns = true; s->view[0] = s->view[1] = nullptr;
  s->bytes = parity_bytes(c, prec) * (size_t)subset;
  cudaSetDevice(c->device);
  if (cudaMalloc(&s->d, s->bytes) != cudaSuccess) {
    set_error("cudaMalloc of %zu bytes failed: %s", s->bytes, cudaGetErrorString(cudaGetLastError()));
    delete s;
    return nullptr;
  }
  cudaMemsetAsync(s->d, 0, s->bytes, c->stream);
  c->spinors.insert(s);
  return s;
}
int tmq_spinor_free(tmq_spinor *s) {
  if (!s) return 0;
  if (!s->owns) { set_error("Even()/Odd() views are freed with their FULL field"); return 1; }
  s->ctx->spinors.erase(s);
  for (int i = 0; i < 2; i++) if (s->view[i]) delete s->view[i];
  if (s->owns && s->d) { cudaStreamSynchronize(s->ctx->stream); cudaFree(s->d); }
  delete s;
  return 0;
}
size_t tmq_spinor_bytes(const tmq_spinor *s) { return s ? s->bytes : 0; }

static tmq_spinor *make_view(tmq_spinor *f, int which) {
  if (!f || f->subset != TMQ_SUBSET_FULL) { set_error("Even()/Odd() need a FULL field"); return nullptr; }
  if (!f->view[which]) {
    tmq_spinor *v = new tmq_spinor();
    v->ctx = f->ctx; v->prec = f->prec; v->subset = TMQ_SUBSET_PARITY; v->owns = false; v->view[0] = v->view[1] = nullptr;
    v->bytes = f->bytes / 2;
    v->d = (char *)f->d + (size_t)which * v->bytes;
    f->view[which] = v;
  }
  return f->view[which];
}
tmq_spinor *tmq_spinor_even(tmq_spinor *f) { return make_view(f, 0); }
tmq_spinor *tmq_spinor_odd(tmq_spinor *f) { return make_view(f, 1); }

int tmq_spinor_from_qkxtm(tmq_spinor *dst, const void *d_qk, int qprec, int parity) {
  TMQ_REQUIRE(dst && d_qk, "null argument");
  TMQ_REQUIRE(qprec == 8 || qprec == 4, "bad QKXTM precision");
  tmq_ctx *c = dst->ctx;
  void *ev = nullptr, *od = nullptr;
  if (dst->subset == TMQ_SUBSET_FULL) { ev = dst->d; od = (char *)dst->d + dst->bytes / 2; }
  else { TMQ_REQUIRE(parity == 0 || parity == 1, "a PARITY field needs parity 0 or 1"); (parity ? od : ev) = dst->d; }
  TMQ_CUDA(spinor_from_qkxtm(dst->prec, ev, od, d_qk, qprec, c->g, c->stream));
  c->launches++;
  return 0;
}
int tmq_spinor_to_qkxtm(void *d_qk, int qprec, const tmq_spinor *src, int parity, double scale) {
  TMQ_REQUIRE(src && d_qk, "null argument");
  TMQ_REQUIRE(qprec == 8 || qprec == 4, "bad QKXTM precision");
  tmq_ctx *c = src->ctx;
  const void *ev = nullptr, *od = nullptr;
  if (src->subset == TMQ_SUBSET_FULL) { ev = src->d; od = (const char *)src->d + src->bytes / 2; }
  else { TMQ_REQUIRE(parity == 0 || parity == 1, "a PARITY field needs parity 0 or 1"); (parity ? od : ev) = src->d; }
  TMQ_CUDA(spinor_to_qkxtm(d_qk, qprec, src->prec, ev, od, scale, c->g, c->stream));
  c->launches++;
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int tmq_spinor_from_host(tmq_spinor *dst, const double *h) {
  TMQ_REQUIRE(dst && h, "null argument");
  tmq_ctx *c = dst->ctx;
  const size_t Vh = c->g.Vh, blk = Vh * 24 * sizeof(double);
  TMQ_TRY(ensure_stage(c, blk * dst->subset));
  TMQ_CUDA(cudaMemcpyAsync(c->stage, h, blk * dst->subset, cudaMemcpyHostToDevice, c->stream));
  for (int p = 0; p < dst->subset; p++) {
    TMQ_CUDA(spinor_from_host_eo(dst->prec, (char *)dst->d + (size_t)p * parity_bytes(c, dst->prec),
                                 (const double *)((char *)c->stage + (size_t)p * blk), (int)Vh, c->stream));
    c->launches++;
  }
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
int tmq_spinor_to_host(double *h, const tmq_spinor *src) {
  TMQ_REQUIRE(src && h, "null argument");
  tmq_ctx *c = src->ctx;
  const size_t Vh = c->g.Vh, blk = Vh * 24 * sizeof(double);
  TMQ_TRY(ensure_stage(c, blk * src->subset));
  for (int p = 0; p < src->subset; p++) {
    TMQ_CUDA(spinor_to_host_eo((double *)((char *)c->stage + (size_t)p * blk), src->prec,
                               (const char *)src->d + (size_t)p * parity_bytes(c, src->prec), (int)Vh, c->stream));
    c->launches++;
  }
  TMQ_CUDA(cudaMemcpyAsync(h, c->stage, blk * src->subset, cudaMemcpyDeviceToHost, c->stream));
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

#define REQ_PARITY(s) TMQ_REQUIRE((s) && (s)->subset == TMQ_SUBSET_PARITY, #s " must be a PARITY field")
#define REQ_FULL(s) TMQ_REQUIRE((s) && (s)->subset == TMQ_SUBSET_FULL, #s " must be a FULL field")
#define REQ_SAME(a, b) TMQ_REQUIRE((a)->ctx == (b)->ctx && (a)->prec == (b)->prec, #a " and " #b " must share context and precision")
#define REQ_OP(c) TMQ_REQUIRE((c)->op_set, "operator parameters not set (tmq_op_set)")

// ---- host <-> device pipelining for multi-RHS drivers -----------------------------------------------------------------------
// Column k of a propagator needs all of its source before the solve can start and has its solution only when the solve ends, so the
// PCIe copies of ONE solve cannot hide behind it -- but the upload of column k+1 and the download of column k-1 can run behind the
// solve of column k.  Uploads and downloads have their own streams (both DMA directions busy at once) and two staging slots each.
static int io_ensure(tmq_ctx *c) {
  if (c->up_stream) return 0;
  TMQ_CUDA(cudaSetDevice(c->device));
  TMQ_CUDA(cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
  TMQ_CUDA(cudaStreamCreateWithFlags(&c->down_stream, cudaStreamNonBlocking));
  const size_t bytes = (size_t)2 * c->g.Vh * 24 * sizeof(double);
  for (int i = 0; i < 2; i++) {
    TMQ_CUDA(cudaMalloc(&c->io_up[i], bytes));
    TMQ_CUDA(cudaMalloc(&c->io_down[i], bytes));
    TMQ_CUDA(cudaEventCreateWithFlags(&c->ev_up_done[i], cudaEventDisableTiming));
    TMQ_CUDA(cudaEventCreateWithFlags(&c->ev_up_free[i], cudaEventDisableTiming));
    TMQ_CUDA(cudaEventCreateWithFlags(&c->ev_down_ready[i], cudaEventDisableTiming));
    TMQ_CUDA(cudaEventCreateWithFlags(&c->ev_down_done[i], cudaEventDisableTiming));
  }
  return 0;
}
static void io_release(tmq_ctx *c) {
  if (!c->up_stream) return;
  cudaStreamSynchronize(c->up_stream);
  cudaStreamSynchronize(c->down_stream);
  for (int i = 0; i < 2; i++) {
    if (c->io_up[i]) cudaFree(c->io_up[i]);
    if (c->io_down[i]) cudaFree(c->io_down[i]);
    if (c->ev_up_done[i]) cudaEventDestroy(c->ev_up_done[i]);
    if (c->ev_up_free[i]) cudaEventDestroy(c->ev_up_free[i]);
    if (c->ev_down_ready[i]) cudaEventDestroy(c->ev_down_ready[i]);
    if (c->ev_down_done[i]) cudaEventDestroy(c->ev_down_done[i]);
    c->io_up[i] = c->io_down[i] = nullptr;
  }
  cudaStreamDestroy(c->up_stream);
  cudaStreamDestroy(c->down_stream);
  c->up_stream = c->down_stream = nullptr;
  for (void *p : c->host_registered) cudaHostUnregister(p);
  c->host_registered.clear();
}

int tmq_host_prefetch(tmq_ctx *c, int slot, const double *h_full) {
  TMQ_REQUIRE(c && h_full, "null argument");
  TMQ_REQUIRE(slot == 0 || slot == 1, "slot must be 0 or 1");
  TMQ_TRY(io_ensure(c));
  // the slot is free once the conversion that read its previous content has run (first use: the event is unrecorded = complete)
  TMQ_CUDA(cudaStreamWaitEvent(c->up_stream, c->ev_up_free[slot], 0));
  TMQ_CUDA(cudaMemcpyAsync(c->io_up[slot], h_full, (size_t)2 * c->g.Vh * 24 * sizeof(double), cudaMemcpyHostToDevice, c->up_stream));
  TMQ_CUDA(cudaEventRecord(c->ev_up_done[slot], c->up_stream));
  return 0;
}
int tmq_spinor_from_prefetch(tmq_spinor *dst, int slot, int host_order) {
  REQ_FULL(dst);
  tmq_ctx *c = dst->ctx;
  TMQ_REQUIRE(slot == 0 || slot == 1, "slot must be 0 or 1");
  TMQ_REQUIRE(c->up_stream, "tmq_host_prefetch has not been called");
  TMQ_REQUIRE(host_order == TMQ_HOST_ORDER_EO || host_order == TMQ_HOST_ORDER_LEX, "bad host order");
  const size_t pb = parity_bytes(c, dst->prec), blk = (size_t)c->g.Vh * 24 * sizeof(double);
  TMQ_CUDA(cudaStreamWaitEvent(c->stream, c->ev_up_done[slot], 0));
  if (host_order == TMQ_HOST_ORDER_EO) {
    for (int p = 0; p < 2; p++) {
      TMQ_CUDA(spinor_from_host_eo(dst->prec, (char *)dst->d + (size_t)p * pb, (const double *)((const char *)c->io_up[slot] + (size_t)p * blk), c->g.Vh, c->stream));
      c->launches++;
    }
  } else {
    TMQ_CUDA(spinor_from_host_lex(dst->prec, dst->d, (char *)dst->d + pb, (const double *)c->io_up[slot], c->g, c->stream));
    c->launches++;
  }
  TMQ_CUDA(cudaEventRecord(c->ev_up_free[slot], c->stream));
  return 0;
}
int tmq_spinor_to_host_async(double *h_full, const tmq_spinor *src, int slot, int host_order, double scale) {
  TMQ_REQUIRE(h_full, "null argument");
  REQ_FULL(src);
  tmq_ctx *c = src->ctx;
  TMQ_REQUIRE(slot == 0 || slot == 1, "slot must be 0 or 1");
  TMQ_REQUIRE(host_order == TMQ_HOST_ORDER_EO || host_order == TMQ_HOST_ORDER_LEX, "bad host order");
  TMQ_TRY(io_ensure(c));
  const size_t pb = parity_bytes(c, src->prec), blk = (size_t)c->g.Vh * 24 * sizeof(double);
  TMQ_CUDA(cudaStreamWaitEvent(c->stream, c->ev_down_done[slot], 0));       // the slot's previous download has left it
  if (host_order == TMQ_HOST_ORDER_EO) {
    if (scale != 1.0) {
      // (the eo converter has no scale: apply it in place on a scratch-free path -- the lexicographic converter multiplies on the fly)
      for (int p = 0; p < 2; p++) { TMQ_CUDA(blas_ax(src->prec, scale, (char *)src->d + (size_t)p * pb, nvec(c), c->stream)); c->launches++; }
    }
    for (int p = 0; p < 2; p++) {
      TMQ_CUDA(spinor_to_host_eo((double *)((char *)c->io_down[slot] + (size_t)p * blk), src->prec, (const char *)src->d + (size_t)p * pb, c->g.Vh, c->stream));
      c->launches++;
    }
  } else {
    TMQ_CUDA(spinor_to_host_lex((double *)c->io_down[slot], src->prec, src->d, (const char *)src->d + pb, scale, c->g, c->stream));
    c->launches++;
  }
  TMQ_CUDA(cudaEventRecord(c->ev_down_ready[slot], c->stream));
  TMQ_CUDA(cudaStreamWaitEvent(c->down_stream, c->ev_down_ready[slot], 0));
  TMQ_CUDA(cudaMemcpyAsync(h_full, c->io_down[slot], (size_t)2 * blk, cudaMemcpyDeviceToHost, c->down_stream));
  TMQ_CUDA(cudaEventRecord(c->ev_down_done[slot], c->down_stream));
  return 0;
}
int tmq_host_wait(tmq_ctx *c) {
  TMQ_REQUIRE(c, "null context");
  if (c->up_stream) { TMQ_CUDA(cudaStreamSynchronize(c->up_stream)); TMQ_CUDA(cudaStreamSynchronize(c->down_stream)); }
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
int tmq_host_link_probe(tmq_ctx *c, const double *h_src, double *h_dst, int reps, double secs[3]) {
  TMQ_REQUIRE(c && h_src && h_dst && secs && reps > 0, "bad argument");
  TMQ_TRY(io_ensure(c));
  TMQ_TRY(tmq_host_wait(c));
  const size_t bytes = (size_t)2 * c->g.Vh * 24 * sizeof(double);
  for (int phase = 0; phase < 3; phase++) {
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; r++) {
      if (phase != 1) TMQ_CUDA(cudaMemcpyAsync(c->io_up[r & 1], h_src, bytes, cudaMemcpyHostToDevice, c->up_stream));
      if (phase != 0) TMQ_CUDA(cudaMemcpyAsync(h_dst, c->io_down[r & 1], bytes, cudaMemcpyDeviceToHost, c->down_stream));
    }
    TMQ_CUDA(cudaStreamSynchronize(c->up_stream));
    TMQ_CUDA(cudaStreamSynchronize(c->down_stream));
    secs[phase] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  return 0;
}
int tmq_host_alloc_pinned(tmq_ctx *c, void **ptr, size_t bytes) {
  TMQ_REQUIRE(c && ptr, "null argument");
  TMQ_CUDA(cudaSetDevice(c->device));
  TMQ_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
  return 0;
}
int tmq_host_free_pinned(tmq_ctx *c, void *ptr) {
  TMQ_REQUIRE(c, "null context");
  TMQ_CUDA(cudaFreeHost(ptr));
  return 0;
}
int tmq_host_register(tmq_ctx *c, void *ptr, size_t bytes) {
  TMQ_REQUIRE(c && ptr, "null argument");
  for (void *p : c->host_registered) if (p == ptr) return 0;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type == cudaMemoryTypeHost) return 0;      // already page-locked
  cudaGetLastError();
  TMQ_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  c->host_registered.push_back(ptr);
  return 0;
}
int tmq_host_unregister(tmq_ctx *c, void *ptr) {
  TMQ_REQUIRE(c && ptr, "null argument");
  for (size_t i = 0; i < c->host_registered.size(); i++)
    if (c->host_registered[i] == ptr) {
      TMQ_CUDA(cudaHostUnregister(ptr));
      c->host_registered.erase(c->host_registered.begin() + i);
      return 0;
    }
  return 0;
}

// ---- operator -------------------------------------------------------------------------------------------------------
int tmq_op_set(tmq_ctx *c, double kappa, double mu, int matpc) {
  TMQ_REQUIRE(c, "null context");
  TMQ_REQUIRE(matpc >= 0 && matpc <= 3, "bad matpc type %d", matpc);
  c->kappa = kappa; c->mu = mu; c->matpc = matpc; c->op_set = true;
  return 0;
}


int tmq_dslash(tmq_spinor *out, const tmq_spinor *in, int out_parity, int dagger) {
  REQ_PARITY(out); REQ_PARITY(in); REQ_SAME(out, in);
  TMQ_REQUIRE(out->d != in->d, "out must not alias in");
  HopSpec s; s.epi = EPI_PLAIN; s.out_parity = out_parity & 1; s.dagger = dagger ? 1 : 0;
  return apply_hop(out->ctx, out->prec, out->d, in->d, s);
}

int tmq_dslash_twist_xpay(tmq_spinor *out, const tmq_spinor *in, int out_parity, int dagger, const tmq_spinor *x, double k) {
  REQ_PARITY(out); REQ_PARITY(in); REQ_SAME(out, in);
  tmq_ctx *c = out->ctx; REQ_OP(c);
  TMQ_REQUIRE(out->d != in->d, "out must not alias in");
  HopSpec s; s.out_parity = out_parity & 1; s.dagger = dagger ? 1 : 0; s.t1 = tw_Ainv(c, s.dagger);
  if (x) { REQ_PARITY(x); REQ_SAME(out, x); s.epi = EPI_TW_XPAY; s.x = x->d; s.k = k; }
  else s.epi = EPI_TW;
  return apply_hop(c, out->prec, out->d, in->d, s);
}

int tmq_matpc(tmq_spinor *out, const tmq_spinor *in, int dagger) {
  REQ_PARITY(out); REQ_PARITY(in); REQ_SAME(out, in);
  REQ_OP(out->ctx);
  TMQ_REQUIRE(out->d != in->d, "out must not alias in");
  return op_matpc(out->ctx, out->prec, out->d, in->d, dagger ? 1 : 0);
}

int tmq_mdagm(tmq_spinor *out, const tmq_spinor *in) {
  REQ_PARITY(out); REQ_PARITY(in); REQ_SAME(out, in);
  REQ_OP(out->ctx);
  TMQ_REQUIRE(out->d != in->d, "out must not alias in");
  return op_mdagm(out->ctx, out->prec, out->d, in->d, SC_T3);
}

int tmq_mat_full(tmq_spinor *out, const tmq_spinor *in, int dagger) {
  REQ_FULL(out); REQ_FULL(in); REQ_SAME(out, in);
  tmq_ctx *c = out->ctx; REQ_OP(c);
  TMQ_REQUIRE(out->d != in->d, "out must not alias in");
  const size_t pb = parity_bytes(c, out->prec);
  for (int p = 0; p < 2; p++) {
    HopSpec s; s.epi = EPI_TWX_XPAY; s.out_parity = p; s.dagger = dagger ? 1 : 0; s.tx = tw_A(c, s.dagger); s.k = -c->kappa;
    s.x = (const char *)in->d + (size_t)p * pb;
    TMQ_TRY(apply_hop(c, out->prec, (char *)out->d + (size_t)p * pb, (const char *)in->d + (size_t)(1 - p) * pb, s));
  }
  return 0;
}

// Dirac::prepare for the MAT solution type (lib/qudaQKXTM_interface.cpp:2020):
//   sym : src_p = A^-1 (b_p + kappa D A^-1 b_q) ; asym: src_p = b_p + kappa D A^-1 b_q
int tmq_prepare(tmq_spinor *src, const tmq_spinor *b) {
  REQ_PARITY(src); REQ_FULL(b); REQ_SAME(src, b);
  tmq_ctx *c = src->ctx; REQ_OP(c);
  const int prec = src->prec, p = c->matpc & 1, q = 1 - p;
  const bool asym = c->matpc >= 2;
  const size_t pb = parity_bytes(c, prec);
  TMQ_TRY(ensure_scratch(c, prec, 1));
  TMQ_TRY(site_op(c, prec, scr(c, prec, 0), (const char *)b->d + (size_t)q * pb, q, tw_Ainv(c, 0)));
  HopSpec s; s.epi = asym ? EPI_XPAY : EPI_XPAY_TW3; s.out_parity = p; s.k = c->kappa; s.x = (const char *)b->d + (size_t)p * pb;
  s.t3 = tw_Ainv(c, 0);
  return apply_hop(c, prec, src->d, scr(c, prec, 0), s);
}

// Dirac::reconstruct (lib/qudaQKXTM_interface.cpp:2040): x_p = x_pc ; x_q = A^-1 (b_q + kappa D x_p)
int tmq_reconstruct(tmq_spinor *x, const tmq_spinor *xpc, const tmq_spinor *b) {
  REQ_FULL(x); REQ_PARITY(xpc); REQ_FULL(b); REQ_SAME(x, b); REQ_SAME(x, xpc);
  tmq_ctx *c = x->ctx; REQ_OP(c);
  const int prec = x->prec, p = c->matpc & 1, q = 1 - p;
  const size_t pb = parity_bytes(c, prec);
  void *xp = (char *)x->d + (size_t)p * pb, *xq = (char *)x->d + (size_t)q * pb;
  if (xp != xpc->d) TMQ_CUDA(cudaMemcpyAsync(xp, xpc->d, pb, cudaMemcpyDeviceToDevice, c->stream));
  HopSpec s; s.epi = EPI_XPAY_TW3; s.out_parity = q; s.k = c->kappa; s.x = (const char *)b->d + (size_t)q * pb; s.t3 = tw_Ainv(c, 0);
  return apply_hop(c, prec, xq, xp, s);
}

// ---- CG on M^dag M ----------------------------------------------------------------------------------------------------
// Flop model per parity site: M^dag M = 5664, CG blas = 240.
static const double FLOPS_MDAGM = 5664.0, FLOPS_CG_BLAS = 240.0;

static int cg_double(tmq_ctx *c, tmq_spinor *x, const tmq_spinor *b, double tol, int maxiter, int *iters, double *true_res) {
  const int prec = 8;
  const size_t n = nvec(c), pb = parity_bytes(c, prec);
  const bool fused = c->matpc < 2;
  TMQ_TRY(ensure_scratch(c, prec, fused ? 4 : 5));
  void *r = scr(c, prec, fused ? 2 : 3), *p = scr(c, prec, fused ? 3 : 4);
  TMQ_CUDA(cudaMemsetAsync(x->d, 0, pb, c->stream));
  TMQ_CUDA(cudaMemcpyAsync(r, b->d, pb, cudaMemcpyDeviceToDevice, c->stream));
  TMQ_CUDA(cudaMemcpyAsync(p, b->d, pb, cudaMemcpyDeviceToDevice, c->stream));
  TMQ_CUDA(blas_norm2(prec, b->d, n, red_at(c, SC_R2_0), c->stream)); c->launches++;
  TMQ_TRY(reduce_finish(c, SC_R2_0, 1));
  double b2;
  TMQ_TRY(fetch_scal(c, SC_R2_0, 1, &b2));
  c->cg_hist.clear();
  c->cg_hist.push_back(b2);
  const double stop = tol * tol * b2;
  double r2 = b2;
  int k = 0;
  if (b2 == 0.0) { *iters = 0; *true_res = 0.0; return 0; }
  const auto tl0 = std::chrono::steady_clock::now();
  c->cg_reliable_updates = 0;
  // The host needs |r|^2 only for the stopping test.  Waiting for it every iteration leaves the GPU idle for a launch latency per
  // iteration -- nothing at 48^3x96 on one GPU, a tenth of the iteration on an 8-way shard.  So the host runs ONE ITERATION AHEAD: it
  // enqueues iteration k+1, then reads |r|^2 of iteration k.  The test itself also runs on the device (the launch that completes the
  // global |r|^2 sets SC_DONE); the launches of an iteration enqueued after convergence exit at once, so x, r, p and the iteration
  // count are exactly those of the synchronous loop.  Not with NCCL all-reduces (their result is not tested on the device).
  // (nor with the split interior / boundary launches of halo mode 0, whose first launch only holds a partial |r|^2)
  const bool lag = fused && c->opt_cg_lag && ((!c->multi && c->nranks == 1) || (c->multi && c->p2p));
  if (lag) {
    c->h_scal[SC_STOP] = stop; c->h_scal[SC_DONE] = 0.0;
    TMQ_CUDA(cudaMemcpyAsync(c->scal + SC_STOP, c->h_scal + SC_STOP, 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    // the host runs L iterations ahead of the |r|^2 it reads (L = TMQ_OPT_CG_LAG): iteration k is enqueued, then the residual of
    // iteration k - L is looked at.  The device takes the same test in the reduction itself, so nothing is computed past convergence.
    const int L = c->opt_cg_lag < 1 ? 1 : (c->opt_cg_lag > 6 ? 6 : c->opt_cg_lag);
    const double *ring = c->h_scal + SC_COUNT;
    int done = -1, seen = 0;              // seen: iterations whose residual the host has read
    auto look = [&](int j) -> int {       // residual after iteration j (0-based); 1 = converged, -1 = broke down
      if (cudaEventSynchronize(c->ev_ring[j & 7]) != cudaSuccess) { set_error("cudaEventSynchronize failed"); return -1; }
      r2 = ring[j & 7];
      c->cg_hist.push_back(r2);
      seen = j + 1;
      if (!(r2 == r2)) { set_error("CG broke down (NaN residual) at iteration %d", j + 1); return -1; }
      return r2 <= stop ? 1 : 0;
    };
    while (k < maxiter) {
      const int so = SC_R2_0 + (k & 1), sn = SC_R2_0 + ((k + 1) & 1);
      c->cg_iter_cur = k + 1;
      int rc = cg_fused_matvec(c, prec, r, p, so, sn, k == 0);
      if (!rc) rc = scal_to_host_ring(c, sn, k & 7);
      if (!rc) rc = cg_update(c, prec, x->d, p, r, so, SC_PAP, sn, so);
      c->cg_iter_cur = 0;
      if (rc) return rc;
      k++;
      if (k > L) {
        const int v = look(k - 1 - L);
        if (v < 0) return 1;
        if (v > 0) { done = seen; break; }      // `seen` iterations count; the device skipped the ones enqueued after them
      }
    }
    while (done < 0 && seen < k) {              // maxiter reached: the residuals of the last L iterations have not been read yet
      const int v = look(seen);
      if (v < 0) return 1;
      if (v > 0) done = seen;
    }
    if (done >= 0) k = done;
    TMQ_CUDA(cudaStreamSynchronize(c->stream));
    if (c->multi) {
      // launches that exited early published no arrival flags: forget the faces "sent ahead" and let every rank drain before the
      // sequence numbers continue (the skipped ones are never waited for)
      c->prepacked_seq = 0;
      if (c->comm_stream) TMQ_CUDA(cudaStreamSynchronize(c->comm_stream));
      if (c->nranks > 1) TMQ_TRY(comm_barrier(c));
    }
  } else
  while (r2 > stop && k < maxiter) {
    const int so = SC_R2_0 + (k & 1), sn = SC_R2_0 + ((k + 1) & 1);
    if (fused) {
      TMQ_TRY(cg_fused_matvec(c, prec, r, p, so, sn, k == 0));
    } else {
      // generic path (asymmetric preconditioning): Ap = M^dag M p ; <p,Ap> ; r -= alpha Ap ; |r|^2
      void *Ap = scr(c, prec, 2);
      TMQ_TRY(op_mdagm(c, prec, Ap, p, SC_PAP));
      double pap;
      TMQ_TRY(fetch_scal(c, SC_PAP, 1, &pap));
      TMQ_CUDA(blas_axpy_norm(prec, -r2 / pap, Ap, r, n, red_at(c, sn), c->stream)); c->launches++;
      TMQ_TRY(reduce_finish(c, sn, 1));
    }
    // |r|^2 travels to the host while the update kernel runs
    TMQ_TRY(scal_to_host(c, sn, 1));
    TMQ_CUDA(cudaEventRecord(c->ev_r2, c->stream));
    TMQ_TRY(cg_update(c, prec, x->d, p, r, so, SC_PAP, sn, so));
    TMQ_CUDA(cudaEventSynchronize(c->ev_r2));
    r2 = c->h_scal[sn];
    k++;
    c->cg_hist.push_back(r2);
    if (!(r2 == r2)) { set_error("CG broke down (NaN residual) at iteration %d", k); return 1; }
  }
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  c->cg_loop_secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - tl0).count();
  // true residual |b - M^dag M x| / |b|
  void *Ax = p;
  TMQ_TRY(op_mdagm(c, prec, Ax, x->d, SC_T3));
  TMQ_CUDA(blas_xmy_norm(prec, b->d, Ax, n, red_at(c, SC_T0), c->stream)); c->launches++;
  TMQ_TRY(reduce_finish(c, SC_T0, 1));
  double t2;
  TMQ_TRY(fetch_scal(c, SC_T0, 1, &t2));
  *iters = k;
  *true_res = sqrt(t2 / b2);
  return 0;
}

// fp32 inner iterations with reliable updates (structure of upstream inv_cg_quda.cpp):
// the residual is recomputed in fp64 whenever the sloppy |r| has dropped by `delta` relative to its
// running maximum, the accumulated fp32 solution is flushed into the fp64 one, and p is restarted
// against the new residual.
static int cg_mixed(tmq_ctx *c, tmq_spinor *x, const tmq_spinor *b, double tol, int maxiter, double delta, int *iters,
                    double *true_res) {
  TMQ_REQUIRE(c->matpc < 2, "mixed-precision CG supports the symmetric preconditioning only");
  const size_t n = nvec(c), pbd = parity_bytes(c, 8), pbs = parity_bytes(c, 4);
  TMQ_TRY(ensure_scratch(c, 8, 4));
  TMQ_TRY(ensure_scratch(c, 4, 5));
  void *rD = scr(c, 8, 2), *AyD = scr(c, 8, 3);
  void *rS = scr(c, 4, 2), *pS = scr(c, 4, 3), *xS = scr(c, 4, 4);
  void *y = x->d;
  TMQ_CUDA(cudaMemsetAsync(y, 0, pbd, c->stream));
  TMQ_CUDA(cudaMemsetAsync(xS, 0, pbs, c->stream));
  TMQ_CUDA(cudaMemcpyAsync(rD, b->d, pbd, cudaMemcpyDeviceToDevice, c->stream));
  TMQ_CUDA(blas_copy(rS, 4, rD, 8, n, c->stream)); c->launches++;
  TMQ_CUDA(cudaMemcpyAsync(pS, rS, pbs, cudaMemcpyDeviceToDevice, c->stream));
  TMQ_CUDA(blas_norm2(8, b->d, n, red_at(c, SC_R2_0), c->stream)); c->launches++;
  TMQ_TRY(reduce_finish(c, SC_R2_0, 1));
  double b2;
  TMQ_TRY(fetch_scal(c, SC_R2_0, 1, &b2));
  c->cg_hist.clear();
  c->cg_hist.push_back(b2);
  if (b2 == 0.0) { *iters = 0; *true_res = 0.0; return 0; }
  const double stop = tol * tol * b2;
  double r2 = b2, rNorm = sqrt(r2), r0Norm = rNorm, maxrx = rNorm, maxrr = rNorm;
  int k = 0, cur = 0;   // scal[SC_R2_0 + cur] holds r2
  bool fresh_p = true;  // the search direction was (re)built by something other than the fused update: its faces have not been sent ahead
  const auto tl0 = std::chrono::steady_clock::now();
  c->cg_reliable_updates = 0;
  while (r2 > stop && k < maxiter) {
    const int so = SC_R2_0 + cur, sn = SC_R2_0 + (1 - cur);
    TMQ_TRY(cg_fused_matvec(c, 4, rS, pS, so, sn, fresh_p));
    fresh_p = false;
    double sc[3];   // SC_R2_0, SC_R2_1, SC_PAP are adjacent
    TMQ_TRY(fetch_scal(c, SC_R2_0, 3, sc));
    const double r2n = sc[sn - SC_R2_0], pap = sc[SC_PAP - SC_R2_0];
    if (!(r2n == r2n)) { set_error("CG broke down (NaN residual) at iteration %d", k + 1); return 1; }
    rNorm = sqrt(r2n);
    if (rNorm > maxrx) maxrx = rNorm;
    if (rNorm > maxrr) maxrr = rNorm;
    const bool updateX = rNorm < delta * r0Norm && r0Norm <= maxrx;
    const bool updateR = (rNorm < delta * maxrr && r0Norm <= maxrr) || updateX;
    if (!updateR && r2n > stop) {
      TMQ_TRY(cg_update(c, 4, xS, pS, rS, so, SC_PAP, sn, so));
      r2 = r2n;
    } else {
      // reliable update (also taken when the sloppy residual claims convergence, so that the loop only
      // ends on a residual computed in fp64): xS += alpha p ; y += xS ; r = b - A y ; p = r + beta p
      TMQ_CUDA(blas_axpby(4, r2 / pap, pS, 1.0, xS, n, c->stream)); c->launches++;
      TMQ_CUDA(blas_xpy_mixed(y, xS, n, c->stream)); c->launches++;
      TMQ_CUDA(cudaMemsetAsync(xS, 0, pbs, c->stream));
      TMQ_TRY(op_mdagm(c, 8, AyD, y, SC_T3));
      TMQ_CUDA(cudaMemcpyAsync(rD, b->d, pbd, cudaMemcpyDeviceToDevice, c->stream));
      TMQ_CUDA(blas_axpy_norm(8, -1.0, AyD, rD, n, red_at(c, sn), c->stream)); c->launches++;
      TMQ_TRY(reduce_finish(c, sn, 1));
      double r2t;
      TMQ_TRY(fetch_scal(c, sn, 1, &r2t));
      TMQ_CUDA(blas_copy(rS, 4, rD, 8, n, c->stream)); c->launches++;
      TMQ_CUDA(blas_axpby(4, 1.0, rS, r2t / r2, pS, n, c->stream)); c->launches++;
      r2 = r2t;
      rNorm = sqrt(r2);
      r0Norm = rNorm; maxrr = rNorm; maxrx = rNorm;
      c->cg_reliable_updates++;
      fresh_p = true;
    }
    cur = 1 - cur;
    k++;
    c->cg_hist.push_back(r2);
  }
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  c->cg_loop_secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - tl0).count();
  // flush what is left in the sloppy accumulator and compute the true residual
  TMQ_CUDA(blas_xpy_mixed(y, xS, n, c->stream)); c->launches++;
  TMQ_TRY(op_mdagm(c, 8, AyD, y, SC_T3));
  TMQ_CUDA(blas_xmy_norm(8, b->d, AyD, n, red_at(c, SC_T0), c->stream)); c->launches++;
  TMQ_TRY(reduce_finish(c, SC_T0, 1));
  double t2;
  TMQ_TRY(fetch_scal(c, SC_T0, 1, &t2));
  *iters = k;
  *true_res = sqrt(t2 / b2);
  return 0;
}

int tmq_cg_mdagm(tmq_spinor *x, const tmq_spinor *b, double tol, int maxiter, double reliable_delta, int sloppy_prec,
                 int *iters, double *true_res, double *secs, double *gflops) {
  REQ_PARITY(x); REQ_PARITY(b); REQ_SAME(x, b);
  tmq_ctx *c = x->ctx; REQ_OP(c);
  TMQ_REQUIRE(x->prec == 8, "the solution field must be fp64 (cuda_prec = double, lib/qudaQKXTM_kernels.cu:1031)");
  TMQ_REQUIRE(sloppy_prec == 8 || sloppy_prec == 4, "sloppy precision must be 8 or 4");
  TMQ_REQUIRE(x->d != b->d, "x must not alias b");
  TMQ_REQUIRE(tol > 0 && maxiter >= 0, "bad tolerance / maxiter");
  TMQ_CUDA(cudaSetDevice(c->device));
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  const auto t0 = std::chrono::steady_clock::now();
  int it = 0;
  double tr = 0;
  if (sloppy_prec == 8) TMQ_TRY(cg_double(c, x, b, tol, maxiter, &it, &tr));
  else TMQ_TRY(cg_mixed(c, x, b, tol, maxiter, reliable_delta > 0 ? reliable_delta : 1e-4, &it, &tr));
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  TMQ_TRY(check_device_error(c));
  const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (iters) *iters = it;
  if (true_res) *true_res = tr;
  if (secs) *secs = s;
  if (gflops) *gflops = s > 0 ? (FLOPS_MDAGM + FLOPS_CG_BLAS) * (double)(c->Vglobal / 2) * it / s * 1e-9 : 0.0;
  return 0;
}

int tmq_cg_stats(tmq_ctx *c, double *loop_secs, int *reliable_updates) {
  TMQ_REQUIRE(c, "null context");
  if (loop_secs) *loop_secs = c->cg_loop_secs;
  if (reliable_updates) *reliable_updates = c->cg_reliable_updates;
  return 0;
}
int tmq_cg_history(tmq_ctx *c, double *r2, int n) {
  TMQ_REQUIRE(c && r2, "null argument");
  const int m = (int)c->cg_hist.size();
  for (int i = 0; i < n && i < m; i++) r2[i] = c->cg_hist[i];
  return 0;
}

// ---- blas -----------------------------------------------------------------------------------------------------------
static inline size_t nvec_of(const tmq_spinor *s) { return (size_t)6 * s->ctx->g.Vh * (size_t)s->subset; }
#define REQ_SHAPE(a, b) TMQ_REQUIRE((a) && (b) && (a)->ctx == (b)->ctx && (a)->subset == (b)->subset && (a)->prec == (b)->prec, #a " and " #b " must have the same shape and precision")

int tmq_zero(tmq_spinor *x) {
  TMQ_REQUIRE(x, "null argument");
  TMQ_CUDA(blas_zero(x->d, x->bytes, x->ctx->stream));
  return 0;
}
int tmq_copy(tmq_spinor *dst, const tmq_spinor *src) {
  TMQ_REQUIRE(dst && src && dst->ctx == src->ctx && dst->subset == src->subset, "copy needs fields of the same shape");
  TMQ_CUDA(blas_copy(dst->d, dst->prec, src->d, src->prec, nvec_of(dst), dst->ctx->stream)); dst->ctx->launches++;
  return 0;
}
int tmq_ax(double a, tmq_spinor *x) {
  TMQ_REQUIRE(x, "null argument");
  TMQ_CUDA(blas_ax(x->prec, a, x->d, nvec_of(x), x->ctx->stream)); x->ctx->launches++;
  return 0;
}
int tmq_axpy(double a, const tmq_spinor *x, tmq_spinor *y) {
  REQ_SHAPE(x, y);
  TMQ_CUDA(blas_axpby(y->prec, a, x->d, 1.0, y->d, nvec_of(y), y->ctx->stream)); y->ctx->launches++;
  return 0;
}
int tmq_axpby(double a, const tmq_spinor *x, double b, tmq_spinor *y) {
  REQ_SHAPE(x, y);
  TMQ_CUDA(blas_axpby(y->prec, a, x->d, b, y->d, nvec_of(y), y->ctx->stream)); y->ctx->launches++;
  return 0;
}
int tmq_xpay(const tmq_spinor *x, double a, tmq_spinor *y) {
  REQ_SHAPE(x, y);
  TMQ_CUDA(blas_axpby(y->prec, 1.0, x->d, a, y->d, nvec_of(y), y->ctx->stream)); y->ctx->launches++;
  return 0;
}
int tmq_caxpy(const double a[2], const tmq_spinor *x, tmq_spinor *y) {
  REQ_SHAPE(x, y);
  TMQ_CUDA(blas_caxpy(y->prec, a[0], a[1], x->d, y->d, nvec_of(y), y->ctx->stream)); y->ctx->launches++;
  return 0;
}
int tmq_cxpaypbz(const tmq_spinor *x, const double a[2], const tmq_spinor *y, const double b[2], tmq_spinor *z) {
  REQ_SHAPE(x, z); REQ_SHAPE(y, z);
  TMQ_CUDA(blas_cxpaypbz(z->prec, x->d, a[0], a[1], y->d, b[0], b[1], z->d, nvec_of(z), z->ctx->stream)); z->ctx->launches++;
  return 0;
}
int tmq_norm2(const tmq_spinor *x, double *out) {
  TMQ_REQUIRE(x && out, "null argument");
  tmq_ctx *c = x->ctx;
  TMQ_CUDA(blas_norm2(x->prec, x->d, nvec_of(x), red_at(c, SC_T0), c->stream)); c->launches++;
  TMQ_TRY(reduce_finish(c, SC_T0, 1));
  return fetch_scal(c, SC_T0, 1, out);
}
int tmq_redot(const tmq_spinor *x, const tmq_spinor *y, double *out) {
  REQ_SHAPE(x, y); TMQ_REQUIRE(out, "null argument");
  tmq_ctx *c = x->ctx;
  TMQ_CUDA(blas_redot(x->prec, x->d, y->d, nvec_of(x), red_at(c, SC_T0), c->stream)); c->launches++;
  TMQ_TRY(reduce_finish(c, SC_T0, 1));
  return fetch_scal(c, SC_T0, 1, out);
}
int tmq_cdot(const tmq_spinor *x, const tmq_spinor *y, double out[2]) {
  REQ_SHAPE(x, y); TMQ_REQUIRE(out, "null argument");
  tmq_ctx *c = x->ctx;
  TMQ_CUDA(blas_cdot(x->prec, x->d, y->d, nvec_of(x), red_at(c, SC_T0), c->stream)); c->launches++;
  TMQ_TRY(reduce_finish(c, SC_T0, 2));
  return fetch_scal(c, SC_T0, 2, out);
}
int tmq_axpy_norm(double a, const tmq_spinor *x, tmq_spinor *y, double *out) {
  REQ_SHAPE(x, y); TMQ_REQUIRE(out, "null argument");
  tmq_ctx *c = x->ctx;
  TMQ_CUDA(blas_axpy_norm(y->prec, a, x->d, y->d, nvec_of(y), red_at(c, SC_T0), c->stream)); c->launches++;
  TMQ_TRY(reduce_finish(c, SC_T0, 1));
  return fetch_scal(c, SC_T0, 1, out);
}
int tmq_xmy_norm(const tmq_spinor *x, tmq_spinor *y, double *out) {
  REQ_SHAPE(x, y); TMQ_REQUIRE(out, "null argument");
  tmq_ctx *c = x->ctx;
  TMQ_CUDA(blas_xmy_norm(y->prec, x->d, y->d, nvec_of(y), red_at(c, SC_T0), c->stream)); c->launches++;
  TMQ_TRY(reduce_finish(c, SC_T0, 1));
  return fetch_scal(c, SC_T0, 1, out);
}
int tmq_axpy_zpbx(double a, tmq_spinor *x, tmq_spinor *y, const tmq_spinor *z, double b) {
  REQ_SHAPE(x, y); REQ_SHAPE(z, y);
  TMQ_CUDA(blas_axpy_zpbx(y->prec, a, x->d, y->d, z->d, b, nvec_of(y), y->ctx->stream)); y->ctx->launches++;
  return 0;
}
int tmq_gamma5(tmq_spinor *x) {
  TMQ_REQUIRE(x, "null argument");
  tmq_ctx *c = x->ctx;
  // UKQCD gamma5 = spin swap 0<->2, 1<->3 (apply_gamma5_vector_core.h:1-16) = swap of the vector halves
  // j <-> j+3 of each parity block
  const size_t pb = parity_bytes(c, x->prec), half = pb / 2;
  TMQ_TRY(ensure_stage(c, half));
  for (int p = 0; p < x->subset; p++) {
    char *blk = (char *)x->d + (size_t)p * pb;
    TMQ_CUDA(cudaMemcpyAsync(c->stage, blk, half, cudaMemcpyDeviceToDevice, c->stream));
    TMQ_CUDA(cudaMemcpyAsync(blk, blk + half, half, cudaMemcpyDeviceToDevice, c->stream));
    TMQ_CUDA(cudaMemcpyAsync(blk + half, c->stage, half, cudaMemcpyDeviceToDevice, c->stream));
  }
  return 0;
}

// ---- QKXTM container kernels ------------------------------------------------------------------------------------------
size_t tmq_qkxtm_ghost_sites(tmq_ctx *c) {
  if (!c) return 0;
  const QkGhost gh = qk_ghost_layout(c->g);
  return gh.total_sites - (size_t)2 * c->g.Vh;
}
// the containers' ghost exchange: replaces ghostToHost -> cpuExchangeGhost -> ghostToDevice (lib/qudaQKXTM_Gauge.cpp:143-373,
// lib/qudaQKXTM_Vector.cpp:172-382, lib/qudaQKXTM_Propagator.cpp) by ONE device-side exchange per partitioned dimension: the two boundary
// slices are gathered on the device and sent to the neighbours (ncclSend / ncclRecv over NVLink), which receive them straight into the
// ghost region behind their local volume.  Nothing is staged through the host.
int tmq_qkxtm_exchange_ghost(tmq_ctx *c, void *d_elem, int prec, int ncomp) {
  TMQ_REQUIRE(c && d_elem, "null argument");
  TMQ_REQUIRE(prec == 8 || prec == 4, "bad precision");
  TMQ_REQUIRE(ncomp > 0, "bad number of components");
  const QkGhost gh = qk_ghost_layout(c->g);
  const size_t cb = (size_t)2 * prec;            // bytes per complex
  for (int d = 2; d < 4; d++) {
    if (!c->g.part[d]) continue;
    const size_t nbytes = gh.surf[d] * ncomp * cb;
    char *lo = nullptr;
    TMQ_CUDA(cudaMalloc((void **)&lo, 2 * nbytes));
    char *hi = lo + nbytes;
    TMQ_CUDA(qkxtm_face_gather(lo, hi, d_elem, prec, c->g, d, ncomp, c->stream)); c->launches++;
    // my slice 0 -> rank-1 (its plus ghost), my slice L-1 -> rank+1 (its minus ghost)
    int rc = comm_sendrecv_dim(c, d, lo, hi, (char *)d_elem + gh.plus[d] * ncomp * cb, (char *)d_elem + gh.minus[d] * ncomp * cb, nbytes, c->stream);
    cudaStreamSynchronize(c->stream);
    cudaFree(lo);
    if (rc) return rc;
  }
  return 0;
}
// sum Re tr P / (V_global * 3 * 6) of a gauge container.  On a partitioned lattice d_gauge must own its ghost region
// ((V + tmq_qkxtm_ghost_sites) * 36 complex): it is exchanged here, as QKXTM_Gauge::calculatePlaq does before its kernel.
int tmq_qkxtm_plaquette(tmq_ctx *c, const void *d_gauge, int prec, double *plaq) {
  TMQ_REQUIRE(c && d_gauge && plaq, "null argument");
  TMQ_REQUIRE(prec == 8 || prec == 4, "bad precision");
  TMQ_REQUIRE((size_t)(c->g.Vh / 64 + 1) * 1 <= c->partials_len, "partials too small");
  if (c->multi) TMQ_TRY(tmq_qkxtm_exchange_ghost(c, (void *)d_gauge, prec, 36));
  TMQ_CUDA(qkxtm_plaquette(d_gauge, prec, c->g, red_at(c, SC_T0), c->stream)); c->launches++;
  TMQ_TRY(reduce_finish(c, SC_T0, 1));
  double s;
  TMQ_TRY(fetch_scal(c, SC_T0, 1, &s));
  *plaq = s / ((double)c->Vglobal * 3.0 * 6.0);
  return 0;
}
int tmq_qkxtm_scale(tmq_ctx *c, void *d, int prec, double a) {
  TMQ_REQUIRE(c && d, "null argument");
  TMQ_CUDA(qkxtm_scale(d, prec, a, (size_t)24 * c->g.Vh, c->stream)); c->launches++;
  return 0;
}
int tmq_qkxtm_cast(tmq_ctx *c, void *dst, int dprec, const void *src, int sprec) {
  TMQ_REQUIRE(c && dst && src, "null argument");
  TMQ_CUDA(qkxtm_cast(dst, dprec, src, sprec, (size_t)24 * c->g.Vh, c->stream)); c->launches++;
  return 0;
}
int tmq_qkxtm_gamma5(tmq_ctx *c, void *d, int prec) {
  TMQ_REQUIRE(c && d, "null argument");
  TMQ_CUDA(qkxtm_gamma5(d, prec, 2 * c->g.Vh, c->stream)); c->launches++;
  return 0;
}
// propagator[(mu*4+nu)*9 + c1*3+c2][x] <- vector[mu*3+c1][x] for the fixed column (nu, c2)
// (lib/qudaQKXTM_Propagator.cpp:90-106: 12 strided device-to-device copies)
int tmq_qkxtm_absorb(tmq_ctx *c, void *d_prop, const void *d_vec, int prec, int nu, int c2) {
  TMQ_REQUIRE(c && d_prop && d_vec, "null argument");
  TMQ_REQUIRE(nu >= 0 && nu < 4 && c2 >= 0 && c2 < 3, "bad column");
  const size_t V = (size_t)2 * c->g.Vh, cb = V * 2 * prec;
  for (int mu = 0; mu < 4; mu++)
    for (int c1 = 0; c1 < 3; c1++) {
      char *dst = (char *)d_prop + ((size_t)(mu * 4 + nu) * 9 + c1 * 3 + c2) * cb;
      const char *src = (const char *)d_vec + (size_t)(mu * 3 + c1) * cb;
      TMQ_CUDA(cudaMemcpyAsync(dst, src, cb, cudaMemcpyDeviceToDevice, c->stream));
    }
  return 0;
}

int tmq_dev_malloc(tmq_ctx *c, void **ptr, size_t bytes) {
  TMQ_REQUIRE(c && ptr, "null argument");
  TMQ_CUDA(cudaSetDevice(c->device));
  TMQ_CUDA(cudaMalloc(ptr, bytes));
  return 0;
}
int tmq_dev_free(tmq_ctx *c, void *ptr) {
  TMQ_REQUIRE(c, "null context");
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  TMQ_CUDA(cudaFree(ptr));
  return 0;
}
int tmq_dev_memset(tmq_ctx *c, void *ptr, int value, size_t bytes) {
  TMQ_REQUIRE(c && ptr, "null argument");
  TMQ_CUDA(cudaMemsetAsync(ptr, value, bytes, c->stream));
  return 0;
}
int tmq_h2d(tmq_ctx *c, void *dst, const void *src, size_t bytes) {
  TMQ_REQUIRE(c && dst && src, "null argument");
  TMQ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
int tmq_d2h(tmq_ctx *c, void *dst, const void *src, size_t bytes) {
  TMQ_REQUIRE(c && dst && src, "null argument");
  TMQ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
int tmq_d2d(tmq_ctx *c, void *dst, const void *src, size_t bytes) {
  TMQ_REQUIRE(c && dst && src, "null argument");
  TMQ_CUDA(cudaSetDevice(c->device));
  TMQ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}

// ---- measurement ------------------------------------------------------------------------------------------------------
int tmq_time_kernel(tmq_ctx *c, int kind, int prec, int reps, const tmq_spinor *in, int flush_l2, double *ms_per_app,
                    long long *launches) {
  TMQ_REQUIRE(c && ms_per_app, "null argument");
  REQ_PARITY(in); REQ_OP(c);
  TMQ_REQUIRE(prec == 8 || prec == 4, "bad precision");
  TMQ_REQUIRE(kind >= 0 && kind <= 5 && reps > 0, "bad kind / reps");
  TMQ_REQUIRE(c->matpc < 2 || kind == 5, "timing kinds 0-4 are defined for the symmetric preconditioning");
  TMQ_TRY(ensure_scratch(c, prec, 6));
  const size_t n = nvec(c);
  void *src = scr(c, prec, 4), *dst = scr(c, prec, 5), *r = scr(c, prec, 2), *p = scr(c, prec, 3);
  TMQ_CUDA(blas_copy(src, prec, in->d, in->prec, n, c->stream));
  TMQ_CUDA(blas_copy(p, prec, in->d, in->prec, n, c->stream));
  TMQ_CUDA(blas_copy(r, prec, in->d, in->prec, n, c->stream));
  TMQ_CUDA(blas_norm2(prec, src, n, red_at(c, SC_R2_0), c->stream));
  TMQ_TRY(reduce_finish(c, SC_R2_0, 1));
  void *flush = nullptr;
  const size_t flush_bytes = (size_t)256 << 20;
  if (flush_l2) TMQ_CUDA(cudaMalloc(&flush, flush_bytes));
  const int pq = c->matpc & 1;
  if (kind == 5) {
    // Chebyshev filter of degree `reps` (4 Dslash launches per degree, recurrence fused): ms per degree
    if (flush) { cudaFree(flush); flush = nullptr; }
    const long long l0 = c->launches;
    TMQ_TRY(op_poly_mdagm(c, prec, dst, src, 2, 0.1, 4.0));
    const long long per = (c->launches - l0) / 2;
    TMQ_CUDA(cudaEventRecord(c->ev_a, c->stream));
    TMQ_TRY(op_poly_mdagm(c, prec, dst, src, reps, 0.1, 4.0));
    TMQ_CUDA(cudaEventRecord(c->ev_b, c->stream));
    TMQ_CUDA(cudaEventSynchronize(c->ev_b));
    float ms = 0;
    TMQ_CUDA(cudaEventElapsedTime(&ms, c->ev_a, c->ev_b));
    TMQ_TRY(check_device_error(c));
    *ms_per_app = ms / reps;
    if (launches) *launches = per;
    return 0;
  }
  auto one = [&](int it) -> int {
    HopSpec s;
    switch (kind) {
      case 0: s.epi = EPI_PLAIN; s.out_parity = 1 - pq; return apply_hop(c, prec, dst, src, s);
      case 1: s.epi = EPI_TW; s.out_parity = 1 - pq; s.t1 = tw_Ainv(c, 0); return apply_hop(c, prec, dst, src, s);
      case 2: s.epi = EPI_TW_XPAY; s.out_parity = pq; s.t1 = tw_Ainv(c, 0); s.k = -c->kappa * c->kappa; s.x = p;
              return apply_hop(c, prec, dst, src, s);
      case 3: return op_mdagm(c, prec, dst, src, SC_T3);
      default: {
        const int so = SC_R2_0 + (it & 1), sn = SC_R2_0 + ((it + 1) & 1);
        TMQ_TRY(cg_fused_matvec(c, prec, r, p, so, sn, it == 0));   // it == 0: p was just (re)written by a copy
        TMQ_TRY(cg_update(c, prec, dst, p, r, so, SC_PAP, sn, so));
        return 0;
      }
    }
  };
  const long long l0 = c->launches;
  TMQ_TRY(one(0));   // warm-up (also loads the module)
  const long long per = c->launches - l0;
  if (kind == 4) {   // restore the CG state the warm-up advanced
    TMQ_CUDA(blas_copy(p, prec, in->d, in->prec, n, c->stream));
    TMQ_CUDA(blas_copy(r, prec, in->d, in->prec, n, c->stream));
    TMQ_CUDA(blas_norm2(prec, src, n, red_at(c, SC_R2_0), c->stream));
    TMQ_TRY(reduce_finish(c, SC_R2_0, 1));
  }
  double total = 0.0;
  if (!flush_l2) {
    TMQ_CUDA(cudaEventRecord(c->ev_a, c->stream));
    for (int i = 0; i < reps; i++) TMQ_TRY(one(i));
    TMQ_CUDA(cudaEventRecord(c->ev_b, c->stream));
    TMQ_CUDA(cudaEventSynchronize(c->ev_b));
    float ms = 0;
    TMQ_CUDA(cudaEventElapsedTime(&ms, c->ev_a, c->ev_b));
    total = ms;
  } else {
    for (int i = 0; i < reps; i++) {
      TMQ_CUDA(cudaMemsetAsync(flush, i & 0xff, flush_bytes, c->stream));
      TMQ_CUDA(cudaEventRecord(c->ev_a, c->stream));
      TMQ_TRY(one(i));
      TMQ_CUDA(cudaEventRecord(c->ev_b, c->stream));
      TMQ_CUDA(cudaEventSynchronize(c->ev_b));
      float ms = 0;
      TMQ_CUDA(cudaEventElapsedTime(&ms, c->ev_a, c->ev_b));
      total += ms;
    }
  }
  if (flush) cudaFree(flush);
  TMQ_TRY(check_device_error(c));
  *ms_per_app = total / reps;
  if (launches) *launches = per;
  return 0;
}

long long tmq_launch_count(tmq_ctx *c) { return c ? c->launches : 0; }

}  // extern "C"
