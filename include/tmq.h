/*
 * tmq.h -- C ABI of libtmq.so: the B200-native even-odd twisted-mass Wilson Dslash + CG on M^dag M.
 *
 * This is the drop-in boundary for ONE hot path of ETMC-QUDA/quda-QKXTM-Multigrid-PlugIn.  The plug-in
 * reaches that path through upstream-QUDA C++ internals (lib/qudaQKXTM_interface.cpp:1 textually
 * includes interface_quda.cpp); each entry point below names the reference call site(s) it replaces.
 * The C++ QKXTM shim in quda-qkxtm-multigrid-plugin_b200/host/ (QKXTM_Vector/Gauge/Propagator,
 * init_qudaQKXTM, loadGaugeQuda, the calc_loops / MG_bench solve skeletons) is the only intended caller.
 *
 * Conventions
 *   - all functions return 0 on success, non-zero on error; tmq_last_error() gives the text.
 *     The shim converts non-zero into the reference's errorQuda() abort behaviour.
 *   - plain pointers and sizes only; handles are opaque; one context per GPU / per process.
 *   - calls on one context are serialised by the caller (the reference is single-threaded per rank).
 *   - device work is enqueued on the context's stream; a call returns after the result it promises is
 *     complete (host-visible) unless documented otherwise.
 *   - there is no CPU fallback: every compute entry point fails if no CUDA device is usable.
 *   - operator: kappa normalisation, M_full = A - kappa D, A = 1 + i (2 kappa mu) gamma5, UKQCD basis
 *     (the only basis the plug-in accepts: lib/qudaQKXTM_interface.cpp:64-67).
 */
#ifndef TMQ_H
#define TMQ_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tmq_ctx tmq_ctx;
typedef struct tmq_spinor tmq_spinor;

enum { TMQ_PREC_SINGLE = 4, TMQ_PREC_DOUBLE = 8 };          /* bytes, as QKXTM_Field::Precision() (include/qudaQKXTM.h:150) */
enum { TMQ_SUBSET_PARITY = 1, TMQ_SUBSET_FULL = 2 };          /* QUDA_PARITY_SITE_SUBSET / QUDA_FULL_SITE_SUBSET            */
enum { TMQ_MATPC_EVEN_EVEN = 0, TMQ_MATPC_ODD_ODD = 1,        /* qkxtm/Calc_Loops.cpp:443-450                               */
       TMQ_MATPC_EVEN_EVEN_ASYM = 2, TMQ_MATPC_ODD_ODD_ASYM = 3 };
enum { TMQ_RECON_8 = 8, TMQ_RECON_12 = 12, TMQ_RECON_18 = 18 }; /* --recon 8 / 12 / 18 (qkxtm/misc.cpp:661-683)               */
enum { TMQ_SOLUTION_MAT = 0, TMQ_SOLUTION_MATPC = 1 };        /* QUDA_MAT_SOLUTION / QUDA_MATPC_SOLUTION                    */

const char *tmq_last_error(void);
int tmq_version(void);
int tmq_device_count(void);

/* ---- context: replaces initQuda(device) + initCommsGridQuda + init_qudaQKXTM geometry
 *      (qkxtm/Calc_Loops.cpp:554,753-755; lib/qudaQKXTM_kernels.cu:118-297) ------------------------------
 * localX: local lattice extents (x,y,z,t), all even.  grid/coord: process grid and this rank's
 * coordinate; only z and t may be partitioned (grid[0] = grid[1] = 1).                                   */
tmq_ctx *tmq_create(int device, const int localX[4], const int grid[4], const int coord[4]);
int tmq_destroy(tmq_ctx *);
int tmq_sync(tmq_ctx *);
/* NCCL bootstrap for the halo exchange and the CG all-reduces: rank 0 fills a 128-byte unique id, the
 * caller broadcasts it (MPI / torch.distributed / file), every rank then joins.  Replaces the MPI/QMP
 * communicator QUDA builds in initCommsGridQuda (qkxtm/QKXTM_util.cpp:48-68).                             */
int tmq_comm_unique_id(char id128[128]);
int tmq_comm_init(tmq_ctx *, const char id128[128], int nranks, int rank);
/* rendezvous of all ranks of the communicator: an NCCL all-reduce (which waits without a time limit) followed by a host wait on
 * the context's streams.  Replaces the MPI_Barrier / MPI_Bcast / MPI_Gather ordering points of the reference's host code
 * (e.g. lib/qudaQKXTM_Vector.cpp:625, lib/qudaQKXTM_Contraction.cpp:1580-1584): call it after any phase in which ranks diverge
 * (file I/O by one rank), so that the bounded halo waits of the next Dslash never see that skew.  No-op on one rank.            */
int tmq_barrier(tmq_ctx *);
/* in-place sum of n host doubles over all ranks: the MPI_Reduce / MPI_Gather of the correlator writers
 * (lib/qudaQKXTM_Contraction.cpp:869-875,1580-1584), staged through the device.  No-op on one rank.                             */
int tmq_allreduce_host(tmq_ctx *, double *h, size_t n);
/* force the ghost-zone (pack -> exchange -> interior/boundary) path in dimension d even when grid[d] = 1,
 * where the exchange wraps onto this rank: the reference's --partition mask (qkxtm/QKXTM_util.cpp:1717-1720).
 * Only z (part[2]) and t (part[3]) may be set.  Must be called before any field is created.  A test / debugging aid: do not keep the
 * host pipeline (tmq_host_prefetch / tmq_spinor_to_host_async) busy next to a self-exchanged solve in the copy-engine halo modes --
 * the same-device face copies can queue behind the bulk copies until the halo wait times out (profiles/r2_e2e_n8.md).       */
int tmq_force_partition(tmq_ctx *, const int part[4]);
/* tuning knobs (tile of the thread->site map); 0 keeps the default                                        */
int tmq_set_tile(tmq_ctx *, int ty, int tz, int tt);
enum { TMQ_OPT_PREFETCH = 1, TMQ_OPT_HALO_P2P = 2, TMQ_OPT_BOUNDARY_AT_PCT = 3, TMQ_OPT_SMEAR_BLOCK_T = 4, TMQ_OPT_PACK_ASYNC = 5,
       TMQ_OPT_CONTRACT_SLICES = 6 /* > 0: time slices per pass of the baryon / derivative contractions (default: what fits 2 GiB) */,
       TMQ_OPT_CG_LAG = 8 /* L = 1 (default) .. 6: the fp64 CG's host loop runs L iterations ahead of the |r|^2 it reads, the stopping test is
                             also taken on the device and launches enqueued past convergence exit at once -- same iterates, same iteration
                             count as the synchronous loop (0).  Environment: TMQ_CG_LAG */,
       TMQ_OPT_HALO_TIMEOUT_MS = 7 /* wall-clock limit (ms, default 120000; env TMQ_HALO_TIMEOUT_MS) of a device-side wait for a neighbour's
                                      ghost face or all-reduce contribution; on expiry nothing is computed from stale ghosts, the device
                                      error scalar is raised and the enclosing call (tmq_sync, tmq_cg_mdagm, ...) fails */ };   /* TMQ_OPT_PREFETCH: accepted and ignored (the L2-prefetch experiment was removed: no gain) */
int tmq_set_option(tmq_ctx *, int option, int value);
/* TMQ_OPT_HALO_P2P = 3: FUSED compute + halo exchange.  Inside the chains of applications the library issues back to back (M_pc,
 * M^dag M, the CG iteration) the boundary CTAs of the launch that PRODUCES a field pack its faces for the next application and store
 * them straight into the neighbours' arenas over NVLink, then publish the arrival flags: one kernel per application, no pack launch,
 * no copy-engine transfer, no NCCL call.  Only the first application of a chain uses the stand-alone pack launch of mode 1.
 * tmq_halo_mode returns 4.
 * TMQ_OPT_HALO_P2P = 4: fused PACK + copy-engine push.  As mode 2, but inside those chains the producing launch's boundary CTAs project
 * the faces of their output into this rank's own send buffers (local stores: no peer store, no system-scope fence, nothing published);
 * the next application only has the copy engines push the buffers and the arrival flags, so the stand-alone pack launch and its place
 * on the critical path are gone while no SM ever waits on NVLink.  tmq_halo_mode returns 5.  This is the DEFAULT (measured fastest on 2
 * and 8 GPUs); the environment variable TMQ_HALO_P2P = 0..4 sets the initial mode of every context.  */
/* TMQ_OPT_HALO_P2P selects the ghost exchange.  0: ncclSend/ncclRecv on a separate stream + interior / boundary
 * launches.  1: the pack kernel stores the faces straight into the neighbours' ghost arenas over NVLink peer mappings
 * (CUDA IPC, set up by tmq_comm_init).  2: faces are packed locally and pushed by the copy engines into the
 * neighbours' arenas, overlapping the Dslash.  In modes 1 to 4 the Dslash is ONE launch whose boundary CTAs wait on
 * arrival flags and the CG scalars are all-reduced through peer-memory mailboxes.  tmq_halo_mode returns 0 (not
 * sharded), 1 (NCCL), 2 (peer stores), 3 (peer copies), 4, 5 (see above); all but 1 need every rank's arena to be mappable.       */
int tmq_halo_mode(tmq_ctx *);

/* ---- gauge: replaces loadGaugeQuda / freeGaugeQuda (qkxtm/Calc_Loops.cpp:759,806) ----------------------
 * qdp_eo_gauge[mu]: host, double, [even Vh | odd Vh] x 3x3 complex row-major (QDP order,
 * qkxtm/QKXTM_util.cpp:840-857), with the T boundary condition ALREADY folded into U_t on the last
 * global time slice when t_boundary = -1 (applyGaugeFieldScaling, qkxtm/QKXTM_util.cpp:698-705).
 * Creates resident fp64 and fp32 copies with the requested reconstruct: 18 (all nine entries), 12 (rows 0, 1; row 2 rebuilt in
 * registers) or 8 (U01, U02, U10 and tan(arg U00 / 4), tan(arg U20 / 4): a trig-free 8-real format, 2 square roots and 3 divisions
 * to unpack; fields with |U01|^2 + |U02|^2 < 1e-6 on some link -- a unit field -- cannot be stored and are refused).               */
int tmq_gauge_load(tmq_ctx *, const void *const qdp_eo_gauge[4], int t_boundary, int recon);
int tmq_gauge_free(tmq_ctx *);
/* sum Re tr P / (V_global * 3 * 6): QKXTM_Gauge::calculatePlaq (lib/qudaQKXTM_Gauge.cpp:376-386,
 * lib/qudaQKXTM_kernels.cu:941-957).  Single rank only.                                                    */
int tmq_plaquette(tmq_ctx *, double *plaq);

/* ---- spinor fields: replace cudaColorSpinorField (lib/qudaQKXTM_interface.cpp:1923-1937) --------------- */
tmq_spinor *tmq_spinor_alloc(tmq_ctx *, int prec, int subset);
int tmq_spinor_free(tmq_spinor *);
size_t tmq_spinor_bytes(const tmq_spinor *);
/* QKXTM device layout <-> native: d_qkxtm[(s*3+c)*V*2 + x_lex*2 + ri] (lib/qudaQKXTM_Vector.cpp:72-81).
 * Replaces run_UploadToCuda / run_DownloadFromCuda / run_ScaleVector
 * (lib/qudaQKXTM_kernels.cu:1024-1124).  For a PARITY field `parity` selects the parity it holds
 * (isEven -> 0); download zero-fills the absent parity (downloadFromCuda_core.h) and multiplies by `scale`
 * (the 2*kappa rescale of lib/qudaQKXTM_interface.cpp:200-203 fused in).                                   */
int tmq_spinor_from_qkxtm(tmq_spinor *dst, const void *d_qkxtm, int qkxtm_prec, int parity);
int tmq_spinor_to_qkxtm(void *d_qkxtm, int qkxtm_prec, const tmq_spinor *src, int parity, double scale);
/* host even-odd order [cb][spin][colour][re,im], double (QUDA_DIRAC_ORDER host spinor; a FULL field is
 * [even Vh | odd Vh], include/QKXTM_mapping_parity.h:67-110).  Used by dslash_test-style drivers.          */
int tmq_spinor_from_host(tmq_spinor *dst, const double *h_eo);
int tmq_spinor_to_host(double *h_eo, const tmq_spinor *src);
/* views of the two parities of a FULL field (qudaVec.Even()/Odd(), lib/qudaQKXTM_kernels.cu:1055)          */
tmq_spinor *tmq_spinor_even(tmq_spinor *full);
tmq_spinor *tmq_spinor_odd(tmq_spinor *full);

/* ---- operator: replaces createDirac + Dirac::{Dslash,M,Mdag,MdagM,prepare,reconstruct}
 *      (lib/qudaQKXTM_interface.cpp:1886-1889,2020-2041; lib/qudaQKXTM_Deflation.cpp:153-155,217) -------- */
int tmq_op_set(tmq_ctx *, double kappa, double mu, int matpc);
/* twisted-clover (dslash_type = twisted-clover, qkxtm/MG_Bench.cpp:243-251,605-608): replaces loadCloverQuda(NULL, NULL,
 * &inv_param), which builds the clover field on the device from the resident gauge field with inv_param.clover_coeff =
 * csw * kappa.  Afterwards every operator entry point uses A = C + i (2 kappa mu) gamma5 with C = 1 + i clover_coeff
 * sum_{mu<nu} sigma_munu F_munu in place of the constant twist; (C + i a gamma5)^-1 is rebuilt whenever tmq_op_set changes
 * kappa or mu.  On a sharded lattice the gauge halo the clover leaves need (incl. the z-t corners) is exchanged once,
 * inside this call.  tmq_clover_free returns to plain twisted mass.                                                  */
int tmq_clover_load(tmq_ctx *, double clover_coeff);
int tmq_clover_free(tmq_ctx *);
/* out(parity) = D in or D^dag in : the bare hop (a4)                                                      */
int tmq_dslash(tmq_spinor *out, const tmq_spinor *in, int out_parity, int dagger);
/* out = A^-1 D in (dagger = 0) / A^-dag D^dag in (dagger = 1); with x != NULL: out = x + k * (that)         */
int tmq_dslash_twist_xpay(tmq_spinor *out, const tmq_spinor *in, int out_parity, int dagger,
                          const tmq_spinor *x, double k);
int tmq_matpc(tmq_spinor *out, const tmq_spinor *in, int dagger);       /* M_pc, sym or asym per tmq_op_set */
int tmq_mdagm(tmq_spinor *out, const tmq_spinor *in);                   /* M_pc^dag M_pc (DiracMdagM)        */
int tmq_mat_full(tmq_spinor *out, const tmq_spinor *in, int dagger);    /* A - kappa D on FULL fields        */
int tmq_prepare(tmq_spinor *src_pc, const tmq_spinor *b_full);          /* Dirac::prepare, MAT solution      */
int tmq_reconstruct(tmq_spinor *x_full, const tmq_spinor *x_pc, const tmq_spinor *b_full); /* Dirac::reconstruct */

/* ---- solver: replaces Solver::create(CG) + (*solve)(out,in) on DiracMdagM
 *      (lib/qudaQKXTM_interface.cpp:2031-2037).  Solves M^dag M x = b for PARITY fields, x0 = 0.
 * sloppy_prec = 8: pure fp64; 4: fp32 inner iterations with reliable updates (delta; <= 0 selects the reference drivers'
 * 1e-4, qkxtm/Calc_Loops.cpp:481), fp64 true residual.
 * Outputs mirror QudaInvertParam::{iter,true_res,secs,gflops} (updateInvertParam).                         */
int tmq_cg_mdagm(tmq_spinor *x, const tmq_spinor *b, double tol, int maxiter, double reliable_delta,
                 int sloppy_prec, int *iters, double *true_res, double *secs, double *gflops);
/* statistics of the last solve: wall time of the iteration loop alone (set-up and the final true-residual computation excluded),
 * number of reliable updates (fp64 residual recomputations) of a mixed-precision solve; either pointer may be NULL             */
int tmq_cg_stats(tmq_ctx *, double *loop_secs, int *reliable_updates);
/* optional residual history of the last solve (|r|^2 per iteration), up to n entries                       */
int tmq_cg_history(tmq_ctx *, double *r2, int n);

/* ---- blas: replaces blas::* used by the plug-in (lib/qudaQKXTM_interface.cpp:135-136,
 *      lib/qudaQKXTM_Deflation.cpp:1015-1056,1431-1435) and by CG ----------------------------------------- */
int tmq_zero(tmq_spinor *x);
int tmq_copy(tmq_spinor *dst, const tmq_spinor *src);                                  /* converts precision */
int tmq_ax(double a, tmq_spinor *x);
int tmq_axpy(double a, const tmq_spinor *x, tmq_spinor *y);                             /* y += a x           */
int tmq_axpby(double a, const tmq_spinor *x, double b, tmq_spinor *y);                  /* y = a x + b y      */
int tmq_xpay(const tmq_spinor *x, double a, tmq_spinor *y);                             /* y = x + a y        */
int tmq_caxpy(const double a[2], const tmq_spinor *x, tmq_spinor *y);                   /* y += a x (complex) */
int tmq_cxpaypbz(const tmq_spinor *x, const double a[2], const tmq_spinor *y, const double b[2], tmq_spinor *z);
                                                                                        /* z = x + a y + b z  */
int tmq_norm2(const tmq_spinor *x, double *out);
int tmq_redot(const tmq_spinor *x, const tmq_spinor *y, double *out);
int tmq_cdot(const tmq_spinor *x, const tmq_spinor *y, double out[2]);                  /* sum conj(x) y      */
int tmq_axpy_norm(double a, const tmq_spinor *x, tmq_spinor *y, double *norm2_y);
int tmq_xmy_norm(const tmq_spinor *x, tmq_spinor *y, double *norm2_y);                  /* y = x - y          */
int tmq_axpy_zpbx(double a, tmq_spinor *x, tmq_spinor *y, const tmq_spinor *z, double b);
                                                                                        /* y += a x; x = z + b x */
int tmq_gamma5(tmq_spinor *x);                                                          /* apply_gamma5_vector */

/* ---- eigensolver on M^dag M: replaces QKXTM_Deflation::{polynomialOperator, eigenSolver, deflateVector}
 *      (lib/qudaQKXTM_Deflation.cpp:997-1063, 1069-1475, 614-800), where the reference drives ARPACK p?naupd by
 *      reverse communication with host-staged vectors and deflates with a host zgemv ------------------------------- */
typedef struct tmq_eigset tmq_eigset;
/* out = p(M^dag M) in: the Chebyshev filter with the reference's recurrence (delta = (amax-amin)/2, theta = (amax+amin)/2,
 * sigma_1 = -delta/theta; degree 0 copies, degree -1 applies the bare M^dag M).  PARITY fields: the even-odd operator,
 * FULL fields: the unpreconditioned one.  One degree = 4 fused Dslash launches, the recurrence is their epilogue.       */
int tmq_poly_mdagm(tmq_spinor *out, const tmq_spinor *in, int deg, double amin, double amax);
/* a set of vectors resident in HBM (Krylov basis / eigenvectors), replaces the host h_elem array.  subset =
 * TMQ_SUBSET_PARITY: the even-odd operator M_pc^dag M_pc; TMQ_SUBSET_FULL: the unpreconditioned M_full^dag M_full on
 * [even | odd] fields (the reference's isFullOp, the branch calc_loops uses: lib/qudaQKXTM_interface.cpp:1725-1736)  */
tmq_eigset *tmq_eigset_alloc(tmq_ctx *, int nvec, int prec, int subset);
int tmq_eigset_free(tmq_eigset *);              /* before tmq_destroy of its context                                  */
int tmq_eigset_size(const tmq_eigset *);
tmq_spinor *tmq_eigset_vector(tmq_eigset *, int i);   /* handle of the i-th vector, owned by the set                  */
/* nev eigenpairs of M^dag M (even-odd or full, per the set's subset) by thick-restart Lanczos in a Krylov space of nkv vectors (set size >= nkv + 1).
 * which = 0: smallest (the reference's SR; with poly_deg > 0 the filter turns them into the dominant ones, as the
 * reference's SR <-> LR swap), 1: largest.  tol as ARPACK's: |beta q_m| <= tol max(eps^(2/3), |theta|).
 * On return vectors 0..nev-1 of the set hold orthonormal eigenvectors sorted by ascending eigenvalue, evals / resid
 * (nev each; resid may be NULL) the Rayleigh quotients <v, M^dag M v> and |M^dag M v - lambda v| computed with the true
 * operator (Deflation.cpp:1426-1439).  Handles obtained from tmq_eigset_vector before the call are stale afterwards.  */
int tmq_eigensolve(tmq_eigset *set, int nev, int nkv, int poly_deg, double amin, double amax, double tol, int max_restarts,
                   int which, unsigned long long seed, double *evals, double *resid, int *nconv, int *nrestarts,
                   int *nmatvec);
/* out = U Lambda^-1 U^dag in over the first nvec vectors of the set (deflateVector)                                  */
int tmq_deflate(tmq_spinor *out, const tmq_spinor *in, tmq_eigset *set, const double *evals, int nvec);
/* out = in - U U^dag in over the first nvec vectors (projectVector, lib/qudaQKXTM_Deflation.cpp:1926-2060); out may
 * alias in                                                                                                         */
int tmq_project(tmq_spinor *out, const tmq_spinor *in, tmq_eigset *set, int nvec);

/* ---- QKXTM container kernels on the QKXTM device layout (lib/qudaQKXTM_kernels.cu:1110-1124,1353-1365,
 *      lib/code_pieces/apply_gamma5_vector_core.h, lib/qudaQKXTM_Propagator.cpp:90-106) ------------------- */
/* The containers' ghost zones (lib/qudaQKXTM_Field.cpp:116-125, lib/qudaQKXTM_kernels.cu:160-170): a container's device array is
 * [ncomp][V] complex followed, for every partitioned dimension in ascending order, by the "plus" ghost (the forward neighbour's slice 0)
 * and the "minus" ghost (the backward neighbour's slice L-1), each [ncomp][surface]; tmq_qkxtm_ghost_sites = the number of ghost sites.
 * tmq_qkxtm_exchange_ghost replaces ghostToHost -> cpuExchangeGhost -> ghostToDevice (lib/qudaQKXTM_Gauge.cpp:143-373,
 * lib/qudaQKXTM_Vector.cpp:172-382) by one device-side exchange per partitioned dimension (nothing is staged through the host);
 * ncomp = 36 (gauge), 12 (vector), 144 (propagator).                                                                              */
size_t tmq_qkxtm_ghost_sites(tmq_ctx *);
int tmq_qkxtm_exchange_ghost(tmq_ctx *, void *d_elem, int prec, int ncomp);
/* plaquette of a gauge field in the QKXTM device layout d[((dir*3+c1)*3+c2)*V + x_lex] (lib/qudaQKXTM_Gauge.cpp:73-89):
 * QKXTM_Gauge::calculatePlaq (lib/qudaQKXTM_Gauge.cpp:376-386).  On a partitioned lattice the array must include its ghost region,
 * which is exchanged here first, and the sum is all-reduced over the ranks.                                      */
int tmq_qkxtm_plaquette(tmq_ctx *, const void *d_gauge_qkxtm, int prec, double *plaq);
int tmq_qkxtm_scale(tmq_ctx *, void *d_qkxtm, int prec, double a);
int tmq_qkxtm_cast(tmq_ctx *, void *d_dst, int dst_prec, const void *d_src, int src_prec);
int tmq_qkxtm_gamma5(tmq_ctx *, void *d_qkxtm, int prec);
int tmq_qkxtm_absorb(tmq_ctx *, void *d_prop, const void *d_vec, int prec, int nu, int c2);
/* QKXTM_Vector::gaussianSmearing (lib/qudaQKXTM_Vector.cpp:386-421, kernel body lib/code_pieces/Gauss_core.h):
 * nsmear steps of out = (psi + alpha sum_{mu=x,y,z} [U_mu(x) psi(x+mu) + U_mu(x-mu)^dag psi(x-mu)]) / (1 + 6 alpha) on the
 * QKXTM device layouts (vector d[(s*3+c)*V + x], gauge d[((dir*3+c1)*3+c2)*V + x]).  Like the reference it ping-pongs
 * between the two vectors: d_in is CLOBBERED, the result ends in d_out (nsmear = 0 copies).  The time direction does
 * not hop, so a T-sharded lattice needs no exchange; on a z split (or with tmq_force_partition in z) the two z faces of
 * the vector are exchanged with the z neighbours before every step and the neighbour's U_z face once per call, and the
 * result is bit-identical to the unsharded sweep.  TMQ_OPT_SMEAR_BLOCK_T > 0 sweeps the time slices in blocks of that
 * many slices (block outer, step inner: L2-resident; ignored on a z split); 0 (default) = streaming order.          */
int tmq_qkxtm_gauss_smear(tmq_ctx *, void *d_out, void *d_in, const void *d_gauge, int prec, int nsmear, double alpha);

/* ---- propagator container kernels and the meson two-point contraction: what calcMG_threepTwop_EvenOdd does with the solved
 *      columns (lib/qudaQKXTM_interface.cpp:1190-1223).  Propagator layout d[((mu*4+nu)*9 + c1*3+c2)*V + x_lex]. ------------ */
/* QKXTM_Vector::conjugate / QKXTM_Propagator::conjugate (lib/code_pieces/conjugate_{vector,propagator}_core.h): ncomp = 12 | 144 */
int tmq_qkxtm_conjugate(tmq_ctx *, void *d, int prec, int ncomp);
/* QKXTM_Propagator::apply_gamma5 (lib/code_pieces/apply_gamma5_propagator_core.h): gamma5 on the sink spin index              */
int tmq_qkxtm_gamma5_prop(tmq_ctx *, void *d_prop, int prec);
/* QKXTM_Propagator::rotateToPhysicalBase_device(sign) (lib/qudaQKXTM_Propagator.cpp:108-112, rotateToPhysicalBase_core.h):
 * P <- 1/2 (1 + i sign gamma5) P (1 + i sign gamma5), sign = +-1                                                            */
int tmq_qkxtm_rotate_physical(tmq_ctx *, void *d_prop, int prec, int sign);
/* one (nu, c2) column between a propagator-like and a vector-like array whose component strides are prop_sites / vec_sites,
 * for nsites sites starting at prop_site0 / vec_site0: covers absorbVectorToDevice and copyPropagator (all V sites) and
 * copyPropagator3D / absorbVectorTimeSlice (one time slice of a 4-d field <-> a 3-d one), lib/qudaQKXTM_Vector.cpp:464-512,
 * lib/qudaQKXTM_Propagator.cpp:90-106,533-550.  to_prop != 0: vector -> propagator, else propagator -> vector.            */
int tmq_qkxtm_column_copy(tmq_ctx *, void *d_prop, long long prop_sites, long long prop_site0, void *d_vec, long long vec_sites,
                          long long vec_site0, long long nsites, int prec, int nu, int c2, int to_prop);
/* QKXTM_Contraction::contractMesons (lib/qudaQKXTM_Contraction.cpp:1606-1648, kernel bodies contractMesons_core{,_PosSpace}.h):
 * the ten meson channels (pseudoscalar, scalar, g5g1..g5g4, g1..g4) C_G(x) = s_G tr[G S G^dag g5 S^dag g5] of prop1 (iu = 0) and
 * of prop2 (iu = 1).  corr_pos (host, may be NULL): [x_lex local][iu][ip][re,im].  corr_mom (host, may be NULL):
 * [t GLOBAL][imom][iu][ip][re,im] = sum_xvec exp(-2 pi i p.(x - src_pos)/L) C(x), moms = nmoms x (px,py,pz), src_pos global
 * (x0,y0,z0); on a sharded lattice every rank gets the complete reduced result.                                              */
int tmq_qkxtm_contract_mesons(tmq_ctx *, const void *d_prop1, const void *d_prop2, int prec, const int *moms, int nmoms,
                              const int src_pos[3], double *corr_mom, double *corr_pos);
/* QKXTM_Contraction::contractBaryons, MOMENTUM_SPACE (lib/qudaQKXTM_Contraction.cpp:906-960, kernel body contractBaryons_core.h):
 * the ten baryon channels (nucl_nucl, nucl_roper, roper_nucl, roper_roper, deltapp_deltamm_11/22/33, deltap_deltaz_11/22/33), each a
 * 4 x 4 matrix in the open spin indices, for the two propagator assignments iu = 0 (prop1 as the doubly occurring flavour, prop2 as
 * the other one) and iu = 1 (swapped).  corr_mom (host): [t GLOBAL][imom][iu][ip][gamma][gamma'][re,im]; every rank gets the
 * complete reduced result.  The reference's launcher accepts float propagators only; double is accepted here.               */
int tmq_qkxtm_contract_baryons(tmq_ctx *, const void *d_prop1, const void *d_prop2, int prec, const int *moms, int nmoms,
                               const int src_pos[3], double *corr_mom);
/* fixed-sink three-point function of the nucleon (lib/qudaQKXTM_interface.cpp:764-1170), ultra-local insertion.
 * tmq_qkxtm_seq_source: QKXTM_Contraction::seqSourceFixSinkPart1 / Part2 (lib/qudaQKXTM_Contraction.cpp:1652-1688, kernel bodies
 * seqSourceFixSinkPart{1,2}_core.h + projectors_tm_base.h): the sequential source for source spin / colour (nu, c2) written into time
 * slice `timeslice` of the 4-d vector d_vec_out (the rest of the vector is not touched: zero it first, as the reference does) from
 * the 3-d propagators of that time slice d[((mu*4+nu)*9 + c1*3+c2)*V3 + x].  part = 1: the quark line that occurs twice in the
 * nucleon (two propagators: tex1, tex2 of the reference), part = 2: the line that occurs once (d_prop3d_2 ignored).
 * pid: 0 G4, 1 G5G123, 2 G5G1, 3 G5G2, 4 G5G3 (WHICHPROJECTOR); particle: 0 proton, 1 neutron (WHICHPARTICLE).                    */
int tmq_qkxtm_seq_source(tmq_ctx *, void *d_vec_out, int timeslice, const void *d_prop3d_1, const void *d_prop3d_2, int prec, int nu, int c2,
                         int pid, int particle, int part);
/* the ultra-local part of QKXTM_Contraction::contractFixSink (lib/qudaQKXTM_Contraction.cpp:3008-3110, kernel body
 * fixSinkContractions_local_core.h + gammas_tm_base.h): 16 insertions (1, g1..g4, g5, g5g1..g5g4, g5 sigma) in the twisted basis,
 * corr_mom (host): [t GLOBAL][imom][iop][re,im] = sum_xvec exp(+2 pi i p.(x - src_pos)/L) sum Gamma_iop[n][r] F[r][m]^{ba} S[n][m]^{ba};
 * partflag 1 | 2 selects the rotation sign together with the particle, as in the reference.                                        */
int tmq_qkxtm_fixsink_local(tmq_ctx *, const void *d_seq_prop, const void *d_fwd_prop, int prec, int particle, int partflag, const int *moms,
                            int nmoms, const int src_pos[3], double *corr_mom);
/* the conserved-current (Noether) and one-derivative parts of contractFixSink (kernel bodies fixSinkContractions_noether_core.h,
 * fixSinkContractions_oneD_core.h): d_gauge is the QKXTM gauge container d[((dir*3+c1)*3+c2)*V + x] of the links the current is built
 * from.  corr_noether (host): [t][imom][dir][re,im]; corr_oneD (host): [t][imom][dir][iop][re,im] (the reference's
 * corrThp_oneD[it*Nmoms*4*16*2 + imom*4*16*2 + dir*16*2 + iop*2 + ri], lib/qudaQKXTM_kernels.cu:1694-1712), both with the reference's
 * 1/4.  Periodic neighbours on this rank only: refused on a split lattice (the reference exchanges propagator ghost zones).      */
int tmq_qkxtm_fixsink_derivative(tmq_ctx *, const void *d_seq_prop, const void *d_fwd_prop, const void *d_gauge, int prec, int particle, int partflag,
                                 const int *moms, int nmoms, const int src_pos[3], double *corr_noether, double *corr_oneD);

/* ---- raw device memory for the containers (QKXTM_Field::create_device, lib/qudaQKXTM_Field.cpp:172) ----- */
int tmq_dev_malloc(tmq_ctx *, void **ptr, size_t bytes);
int tmq_dev_free(tmq_ctx *, void *ptr);
int tmq_dev_memset(tmq_ctx *, void *ptr, int value, size_t bytes);
int tmq_h2d(tmq_ctx *, void *dst, const void *src, size_t bytes);
int tmq_d2h(tmq_ctx *, void *dst, const void *src, size_t bytes);
int tmq_d2d(tmq_ctx *, void *dst, const void *src, size_t bytes);     /* asynchronous on the context's stream */

/* ---- host <-> device pipelining for multi-RHS drivers (MG_bench's 12 columns, invertMultiSrcQuda): the PCIe copies of column k+-1
 *      run behind the solve of column k.  Replaces the blocking cudaMemcpy of QKXTM_Vector::loadVector / unloadVector
 *      (lib/qudaQKXTM_Vector.cpp:120-133) and of invertQuda's host fields around each solve (lib/qudaQKXTM_interface.cpp:118-131,205).
 * Host fields are FULL, V*24 doubles, in one of two orders.  Two staging slots per direction; a slot may be reused as soon as the call
 * that consumes it has been issued (the library orders the reuse with events).  Host buffers should be page-locked
 * (tmq_host_alloc_pinned / tmq_host_register) or the copies degrade to synchronous staged ones (still correct).                    */
enum { TMQ_HOST_ORDER_EO = 0,   /* [even Vh | odd Vh][spin][colour][re,im]: invertQuda's host spinors                                 */
       TMQ_HOST_ORDER_LEX = 1   /* [x_lex][spin][colour][re,im]: the plug-in's host vectors before packVector / after unpackVector    */ };
int tmq_host_prefetch(tmq_ctx *, int slot, const double *h_full);                  /* start the upload of a host field into slot       */
int tmq_spinor_from_prefetch(tmq_spinor *dst_full, int slot, int host_order);      /* compute stream: wait for the upload, convert     */
/* convert (times scale) into download slot `slot` on the compute stream, then copy to h_full on the download stream; returns at once */
int tmq_spinor_to_host_async(double *h_full, const tmq_spinor *src_full, int slot, int host_order, double scale);
int tmq_host_wait(tmq_ctx *);                                                     /* every outstanding upload / download is complete  */
/* measurement aid: `reps` uploads of a full host field alone, `reps` downloads alone, then both at once, on the streams and staging
 * slots of the pipeline above; secs[0..2] = wall time of the three phases on this rank (bench.py reports the host-link bandwidth the
 * end-to-end numbers are bounded by, per GPU and with every rank of a node copying at once)                                          */
int tmq_host_link_probe(tmq_ctx *, const double *h_src_full, double *h_dst_full, int reps, double secs[3]);
int tmq_host_alloc_pinned(tmq_ctx *, void **ptr, size_t bytes);
int tmq_host_free_pinned(tmq_ctx *, void *ptr);
int tmq_host_register(tmq_ctx *, void *ptr, size_t bytes);                        /* page-lock a caller-owned buffer (idempotent)      */
int tmq_host_unregister(tmq_ctx *, void *ptr);

/* ---- measurement helpers (CUDA events on the context's stream) ------------------------------------------ */
/* run `reps` back-to-back applications of one kernel flavour on (a copy of) the PARITY field `in` and return
 * the average device time per application in milliseconds (CUDA events on the launching stream).
 * kind: 0 = K1 hop, 1 = K2 hop+A^-1, 2 = K3 hop+A^-1+xpay, 3 = M^dag M (4 kernels), 4 = one fused CG
 * iteration (4 Dslash kernels + update, no host sync), 5 = Chebyshev filter of degree `reps` (ms per degree).
 * prec selects fp64/fp32 arithmetic.  When
 * flush_l2 != 0 a buffer larger than L2 is rewritten before every application and each application is
 * timed separately.                                                                                          */
int tmq_time_kernel(tmq_ctx *, int kind, int prec, int reps, const tmq_spinor *in, int flush_l2,
                    double *ms_per_app, long long *launches);
/* CUDA-event stopwatch on the context's stream, for the asynchronous container kernels: start records an event, stop
 * records a second one, waits for it and returns the device time between them in milliseconds                */
int tmq_timer_start(tmq_ctx *);
int tmq_timer_stop(tmq_ctx *, double *ms);
/* Out-of-bounds net of the library's own (compute-sanitizer is not available on every GPU pool): with the environment variable
 * TMQ_GUARD_BYTES = n every device allocation of the library carries red zones of n bytes filled with a NaN pattern; an out-of-bounds
 * read then poisons the result, and this call returns the number of allocations whose red zones have been written to (0 = clean,
 * < 0 = CUDA error; details in tmq_last_error()).  Always 0 when the variable is unset.                                           */
int tmq_guard_check(void);
/* number of kernels this library has launched on the context since creation                                 */
long long tmq_launch_count(tmq_ctx *);

#ifdef __cplusplus
}
#endif
#endif
