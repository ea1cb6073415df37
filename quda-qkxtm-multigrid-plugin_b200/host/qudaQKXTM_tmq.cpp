// qudaQKXTM_tmq.cpp -- the QKXTM host layer (containers + solve skeletons) over the libtmq.so C ABI.
// See include/qudaQKXTM_tmq.h for the reference interfaces each piece mirrors.  No arithmetic on lattice
// fields happens here except the reference's own host-side repacking loops (packVector, packGauge, ...),
// which the reference also runs on the CPU; everything else is a call into the CUDA library.
#include "../../include/qudaQKXTM_tmq.h"
#include "../../include/tmq_host.h"
#include "qkxtm_internal.h"
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

using namespace quda;

// ---- library-global state (the reference's GK_* globals + QUDA's gaugePrecise etc.) ----------------------------------
namespace {
struct Globals {
  int device = -1;
  bool quda_initialized = false;
  bool qkxtm_initialized = false;     // GK_init_qudaQKXTM_flag (lib/qudaQKXTM_kernels.cu:120)
  bool gauge_loaded = false;
  bool clover_loaded = false;
  int grid[4] = {1, 1, 1, 1};
  int coord[4] = {0, 0, 0, 0};
  int rank = 0, nranks = 1;           // comm_rank() / comm_size(): from the launcher's environment (initCommsGridQuda)
  int localL[4] = {0, 0, 0, 0};
  long long localVolume = 0;
  tmq_ctx *ctx = nullptr;
  QudaVerbosity verbosity = QUDA_SUMMARIZE;
  qkxtm_error_handler on_error = nullptr;
  double op_kappa = 0, op_mu = 0;
  int nsmearGauss = 0;                // GK_nsmearGauss, GK_alphaGauss (lib/qudaQKXTM_kernels.cu:60-64)
  double alphaGauss = 0;
  std::vector<int> moms;              // GK_moms [GK_Nmoms][3] (lib/qudaQKXTM_kernels.cu:75-76, createMomenta :98-116)
  std::vector<int> sourcePosition;    // GK_sourcePosition [Nsources][4]
  int op_matpc = -1;
  bool moms_overflow = false;         // init_qudaQKXTM saw more than MAX_NMOMENTA momenta
  quda::ColorSpinorField *work[3] = {nullptr, nullptr, nullptr};   // parity work fields of solve_device
  quda::ColorSpinorField *io_b = nullptr, *io_x = nullptr;         // FULL fields of invertQuda / invertMultiSrcQuda
} G;

void default_error(const char *msg) {
  fprintf(stderr, "%s\n", msg);
  fflush(stderr);
  std::abort();      // errorQuda aborts the job (SURVEY.md 8b "Errors")
}
}  // namespace

#define errorQuda(...)                                                                          \
  do {                                                                                          \
    char b__[1024];                                                                             \
    int n__ = snprintf(b__, sizeof(b__), "ERROR: ");                                            \
    n__ += snprintf(b__ + n__, sizeof(b__) - n__, __VA_ARGS__);                                 \
    snprintf(b__ + n__, sizeof(b__) - n__, " (%s:%d in %s())", __FILE__, __LINE__, __func__);   \
    (G.on_error ? G.on_error : default_error)(b__);                                             \
  } while (0)
#define printfQuda(...)                                              \
  do {                                                               \
    if (G.verbosity > QUDA_SILENT && G.rank == 0) { printf(__VA_ARGS__); fflush(stdout); } \
  } while (0)
#define TMQ_OK(call)                                                        \
  do {                                                                      \
    if ((call) != 0) errorQuda("libtmq: %s", tmq_last_error());             \
  } while (0)

// ---- process grid ----------------------------------------------------------------------------------------------------------------
// The reference builds its communicator with MPI (qkxtm/QKXTM_util.cpp:48-68).  Here every rank is one process started by any
// launcher that exports a rank and a world size (torchrun --no-python, mpirun, srun); the 128-byte NCCL id travels from rank 0 to
// the others through a file that all ranks of ONE launch agree on (same parent process id, or TMQ_COMM_ID_FILE).
static int env_int(const char *const *names, int dflt) {
  for (int i = 0; names[i]; i++) {
    const char *v = getenv(names[i]);
    if (v && *v) return atoi(v);
  }
  return dflt;
}
// One launch = one nonce.  Every rank derives it independently from what the launcher gives all of them: the run / job id
// (TORCHELASTIC_RUN_ID, SLURM_JOB_ID, TMQ_COMM_NONCE), MASTER_PORT, the world size and -- when the file is node-local -- the parent
// process and ITS start time (a later launch from a recycled pid differs).  A rank only accepts a file that carries its own nonce, so an
// id file left behind by a run that died can never be taken for the current one.
static unsigned long long parent_start_time() {
  char path[64], buf[1024];
  snprintf(path, sizeof(path), "/proc/%ld/stat", (long)getppid());
  FILE *fp = fopen(path, "r");
  if (!fp) return 0;
  const size_t n = fread(buf, 1, sizeof(buf) - 1, fp);
  fclose(fp);
  buf[n] = 0;
  const char *p = strrchr(buf, ')');            // the command name may contain spaces: fields are counted after it
  if (!p) return 0;
  unsigned long long v = 0;
  int field = 2;
  for (p++; *p; ) {
    while (*p == ' ') p++;
    field++;
    if (field == 22) { v = strtoull(p, NULL, 10); break; }
    while (*p && *p != ' ') p++;
  }
  return v;
}
static bool ranks_span_nodes() {
  static const char *local_size[] = {"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "SLURM_NTASKS_PER_NODE", NULL};
  static const char *nnodes[] = {"GROUP_WORLD_SIZE", "SLURM_NNODES", NULL};
  const int ls = env_int(local_size, -1), nn = env_int(nnodes, -1);
  return (ls > 0 && ls < G.nranks) || nn > 1;
}
static void comm_bootstrap() {
  char path[512], nonce[256];
  const char *f = getenv("TMQ_COMM_ID_FILE");
  const char *port = getenv("MASTER_PORT") ? getenv("MASTER_PORT") : "0";
  const char *run = getenv("TMQ_COMM_NONCE") ? getenv("TMQ_COMM_NONCE")
                    : (getenv("TORCHELASTIC_RUN_ID") ? getenv("TORCHELASTIC_RUN_ID") : (getenv("SLURM_JOB_ID") ? getenv("SLURM_JOB_ID") : "-"));
  const bool shared = f && *f;
  memset(nonce, 0, sizeof(nonce));
  if (shared) {
    snprintf(path, sizeof(path), "%s", f);
    snprintf(nonce, sizeof(nonce), "tmq1|%s|%s|%d", run, port, G.nranks);
  } else {
    // /tmp is node-local and the parent differs per node: a launch that spans nodes must name a file on a shared filesystem
    if (ranks_span_nodes())
      errorQuda("the ranks span several nodes: set TMQ_COMM_ID_FILE to a path on a filesystem all of them share (the default id file is node-local)");
    // a directory only this user can write to; the name inside is predictable, the directory makes that harmless
    char dir[256];
    snprintf(dir, sizeof(dir), "/tmp/tmq-%ld", (long)getuid());
    if (mkdir(dir, 0700) != 0 && errno != EEXIST) errorQuda("cannot create %s", dir);
    struct stat st;
    if (lstat(dir, &st) != 0 || !S_ISDIR(st.st_mode) || st.st_uid != getuid() || (st.st_mode & 077) != 0)
      errorQuda("%s must be a directory owned by this user with mode 0700", dir);
    snprintf(path, sizeof(path), "%s/nccl_id_%ld_%s", dir, (long)getppid(), port);
    snprintf(nonce, sizeof(nonce), "tmq1|%s|%s|%d|%ld|%llu", run, port, G.nranks, (long)getppid(), parent_start_time());
  }
  char id[128], rec[sizeof(nonce) + 128];
  if (G.rank == 0) {
    TMQ_OK(tmq_comm_unique_id(id));
    char tmp[600];
    snprintf(tmp, sizeof(tmp), "%s.tmp", path);
    unlink(path);                                  // whatever an earlier run left behind
    unlink(tmp);
    const int fd = open(tmp, O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW, 0600);
    if (fd < 0) errorQuda("cannot create the communicator id file %s", tmp);
    memcpy(rec, nonce, sizeof(nonce));
    memcpy(rec + sizeof(nonce), id, 128);
    const bool ok = write(fd, rec, sizeof(rec)) == (ssize_t)sizeof(rec) && fsync(fd) == 0;
    close(fd);
    if (!ok) errorQuda("cannot write the communicator id file %s", tmp);
    if (rename(tmp, path) != 0) errorQuda("cannot publish the communicator id file %s", path);
  } else {
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
      const int fd = open(path, O_RDONLY | O_NOFOLLOW);
      if (fd >= 0) {
        struct stat st;
        const bool mine = fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && (shared || st.st_uid == getuid());
        const ssize_t n = mine ? read(fd, rec, sizeof(rec)) : -1;
        close(fd);
        if (n == (ssize_t)sizeof(rec) && memcmp(rec, nonce, sizeof(nonce)) == 0) { memcpy(id, rec + sizeof(nonce), 128); break; }
      }
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 300.0)
        errorQuda("rank %d: no communicator id for this launch from rank 0 after 300 s (%s)", G.rank, path);
      std::this_thread::sleep_for(std::chrono::milliseconds(20));
    }
  }
  TMQ_OK(tmq_comm_init(G.ctx, id, G.nranks, G.rank));     // collective: returns once every rank has joined
  if (G.rank == 0) unlink(path);
}

static void ensure_context(const int X[4]) {
  if (!G.quda_initialized) errorQuda("initQuda must be called first");
  if (G.ctx) {
    for (int d = 0; d < 4; d++)
      if (G.localL[d] != X[d]) errorQuda("local lattice %d %d %d %d does not match the initialised one", X[0], X[1], X[2], X[3]);
    return;
  }
  for (int d = 0; d < 4; d++) G.localL[d] = X[d];
  G.localVolume = (long long)X[0] * X[1] * X[2] * X[3];
  G.ctx = tmq_create(G.device, X, G.grid, G.coord);
  if (!G.ctx) errorQuda("libtmq: %s", tmq_last_error());
  if (G.nranks > 1) comm_bootstrap();
}

// ---- QUDA C API slice ---------------------------------------------------------------------------------------------------
QudaGaugeParam newQudaGaugeParam(void) {
  QudaGaugeParam p;
  memset(&p, 0, sizeof(p));
  p.anisotropy = 1.0;
  p.t_boundary = QUDA_ANTI_PERIODIC_T;
  p.cpu_prec = p.cuda_prec = QUDA_DOUBLE_PRECISION;
  p.cuda_prec_sloppy = p.cuda_prec_precondition = QUDA_DOUBLE_PRECISION;
  p.reconstruct = p.reconstruct_sloppy = p.reconstruct_precondition = QUDA_RECONSTRUCT_NO;
  return p;
}
QudaInvertParam newQudaInvertParam(void) {
  QudaInvertParam p;
  memset(&p, 0, sizeof(p));
  p.kappa = -1.0;
  p.dslash_type = QUDA_TWISTED_MASS_DSLASH;
  p.twist_flavor = QUDA_TWIST_SINGLET;
  p.inv_type = QUDA_INVALID_INVERTER;
  p.inv_type_precondition = QUDA_INVALID_INVERTER;
  p.solve_type = QUDA_NORMOP_PC_SOLVE;
  p.solution_type = QUDA_MAT_SOLUTION;
  p.gamma_basis = QUDA_UKQCD_GAMMA_BASIS;
  p.cpu_prec = p.cuda_prec = p.cuda_prec_sloppy = p.cuda_prec_precondition = QUDA_DOUBLE_PRECISION;
  p.input_location = p.output_location = QUDA_CPU_FIELD_LOCATION;
  p.Ls = 1;
  p.tol = 1e-7;
  p.maxiter = 100;
  p.reliable_delta = 1e-4;
  p.residual_type = QUDA_L2_RELATIVE_RESIDUAL;
  p.verbosity = QUDA_SUMMARIZE;
  return p;
}

void setVerbosityQuda(QudaVerbosity v, const char *, FILE *) { G.verbosity = v; }

QudaMultigridParam newQudaMultigridParam(void) {
  QudaMultigridParam p;
  memset(&p, 0, sizeof(p));
  return p;
}
// The multigrid solver is out of scope.  qkxtm/MG_Bench.cpp:614 and CalcMG_2pt3pt_EvenOdd.cpp call newMultigridQuda
// unconditionally and only pass the handle on in inv_param.preconditioner, so it must not abort: it warns and returns an inert handle;
// the solves then run CG on the normal operator (--inv-type cg --solve-type normop-pc), anything else is refused by the entry points.
static int g_mg_sentinel;
void *newMultigridQuda(QudaMultigridParam *) {
  if (G.verbosity > QUDA_SILENT && G.rank == 0)
    fprintf(stderr, "WARNING: newMultigridQuda: the multigrid preconditioner is not provided by this library; the handle is inert and the solves use "
                    "CG on M^dag M (run the drivers with --inv-type cg --solve-type normop-pc)\n");
  return &g_mg_sentinel;
}
void destroyMultigridQuda(void *mg_instance) {
  if (mg_instance && mg_instance != &g_mg_sentinel) errorQuda("destroyMultigridQuda: not a handle of this library");
}

void initCommsGridQuda(int nDim, const int *dims, QudaCommsMap func, void *) {
  if (func) errorQuda("initCommsGridQuda: a custom rank map is not supported (pass NULL: lexicographic, t fastest)");
  if (nDim != 4) errorQuda("Number of communication grid dimensions must be 4");
  if (G.ctx) errorQuda("initCommsGridQuda must come before the first field is created");
  if (dims[0] != 1 || dims[1] != 1) errorQuda("only z and t may be partitioned (gridsize %d %d %d %d)", dims[0], dims[1], dims[2], dims[3]);
  static const char *rank_names[] = {"RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID", NULL};
  static const char *size_names[] = {"WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS", NULL};
  G.rank = env_int(rank_names, 0);
  G.nranks = env_int(size_names, 1);
  long long prod = 1;
  for (int d = 0; d < 4; d++) { if (dims[d] < 1) errorQuda("bad gridsize"); G.grid[d] = dims[d]; prod *= dims[d]; }
  if (prod != G.nranks) errorQuda("gridsize %d x %d x %d x %d needs %lld ranks, the launcher started %d", dims[0], dims[1], dims[2], dims[3], prod, G.nranks);
  if (G.rank < 0 || G.rank >= G.nranks) errorQuda("bad rank %d of %d", G.rank, G.nranks);
  // rank <-> coordinate: t fastest, rank = ((cx gy + cy) gz + cz) gt + ct (include/tmq.h; QUDA's default lexicographic map)
  int r = G.rank;
  for (int d = 3; d >= 0; d--) { G.coord[d] = r % G.grid[d]; r /= G.grid[d]; }
}
// the slice of upstream's comm_quda.h / util_quda.h that the reference's driver-side sources use (qkxtm/QKXTM_util.cpp:55-68,
// include/QKXTM_read_conf.h:98): declared in include/compat/{comm_quda.h,util_quda.h}
struct Topology { int dims[4]; int coords[4]; };
static Topology g_topo = {{1, 1, 1, 1}, {0, 0, 0, 0}};
Topology *default_topo = &g_topo;
extern "C" const int *comm_coords(const Topology *) { return G.coord; }
extern "C" const int *comm_dims(const Topology *) { return G.grid; }
extern "C" void comm_dim_partitioned_set(int) {}      // the --partition mask of the drivers: tmq_force_partition is the test-only equivalent
extern "C" QudaVerbosity getVerbosity(void) { return G.verbosity; }
extern "C" void qkxtm_error_at(const char *file, int line, const char *func, const char *fmt, ...) {
  char b[1024];
  int n = snprintf(b, sizeof(b), "ERROR: ");
  va_list ap;
  va_start(ap, fmt);
  n += vsnprintf(b + n, sizeof(b) - n, fmt, ap);
  va_end(ap);
  if (n < (int)sizeof(b)) snprintf(b + n, sizeof(b) - n, " (%s:%d in %s())", file, line, func);
  (G.on_error ? G.on_error : default_error)(b);
}
// upstream's tests/misc.cpp has it, the reference's copy (qkxtm/misc.cpp) does not although qkxtm/QKXTM_util.cpp calls it
QudaSchwarzType get_schwarz_type(char *s) {
  if (strcmp(s, "additive") == 0) return QUDA_ADDITIVE_SCHWARZ;
  if (strcmp(s, "multiplicative") == 0) return QUDA_MULTIPLICATIVE_SCHWARZ;
  fprintf(stderr, "Error: invalid Schwarz type %s\n", s);
  exit(1);
}
int comm_rank(void) { return G.rank; }
int comm_size(void) { return G.nranks; }
int comm_coord(int dim) { return (dim >= 0 && dim < 4) ? G.coord[dim] : 0; }
int comm_dim(int dim) { return (dim >= 0 && dim < 4) ? G.grid[dim] : 1; }
int comm_dim_partitioned(int dim) { return (dim >= 0 && dim < 4) ? (G.grid[dim] > 1) : 0; }
// a real rendezvous of all ranks (the reference gets its ordering from MPI collectives, e.g. lib/qudaQKXTM_Vector.cpp:625): phases in
// which ranks diverge -- rank 0 writing a file, reading a configuration -- end with one, so that no rank runs ahead into a halo exchange
void comm_barrier(void) {
  if (G.nranks > 1 && G.ctx) TMQ_OK(tmq_barrier(G.ctx));
}

void initQuda(int device) {
  if (G.quda_initialized) return;
  if (tmq_device_count() <= 0) errorQuda("no CUDA device (there is no CPU fallback)");
  static const char *local_names[] = {"LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID", NULL};
  G.device = device < 0 ? env_int(local_names, 0) % tmq_device_count() : device;     // initQuda(-1): one GPU per local rank
  G.quda_initialized = true;
}

static void release_work_fields() {
  for (int i = 0; i < 3; i++) { delete G.work[i]; G.work[i] = nullptr; }
  delete G.io_b; delete G.io_x;
  G.io_b = G.io_x = nullptr;
}
void endQuda(void) {
  release_work_fields();
  if (G.ctx) { tmq_destroy(G.ctx); G.ctx = nullptr; }
  G.quda_initialized = G.qkxtm_initialized = G.gauge_loaded = false;
  G.op_matpc = -1;
}

void loadGaugeQuda(void *h_gauge, QudaGaugeParam *param) {
  if (!param || !h_gauge) errorQuda("null argument");
  if (param->gauge_order != QUDA_QDP_GAUGE_ORDER) errorQuda("only QUDA_QDP_GAUGE_ORDER host links are accepted (qkxtm/Calc_Loops.cpp:197)");
  if (param->cpu_prec != QUDA_DOUBLE_PRECISION) errorQuda("host links must be double precision");
  if (param->anisotropy != 1.0) errorQuda("anisotropy != 1 is not supported");
  if (param->type == QUDA_SMEARED_LINKS) {
    // the 2pt/3pt driver loads smeared links first (qkxtm/CalcMG_2pt3pt_EvenOdd.cpp:675-676); they feed the
    // smearing containers only, which are not on this path
    printfQuda("loadGaugeQuda: smeared links are not used by the solver path; ignored\n");
    return;
  }
  ensure_context(param->X);
  const int recon = (int)param->reconstruct;
  if (recon != 8 && recon != 12 && recon != 18) errorQuda("reconstruct must be 8, 12 or 18");
  TMQ_OK(tmq_gauge_load(G.ctx, (const void *const *)h_gauge, (int)param->t_boundary, recon));
  G.gauge_loaded = true;
}
void loadCloverQuda(void *h_clover, void *h_clovinv, QudaInvertParam *inv_param) {
  if (!G.gauge_loaded) errorQuda("loadCloverQuda needs a resident gauge field (loadGaugeQuda)");
  if (h_clover || h_clovinv) errorQuda("host clover fields are not supported: pass NULL to have the field built on the device");
  if (!inv_param) errorQuda("null argument");
  TMQ_OK(tmq_clover_load(G.ctx, inv_param->clover_coeff));
  G.clover_loaded = true;
  G.op_matpc = -1;              // force createDirac to push kappa / mu again (the inverse depends on them)
}
void freeCloverQuda(void) {
  if (G.ctx) tmq_clover_free(G.ctx);
  G.clover_loaded = false;
}
void freeGaugeQuda(void) {
  release_work_fields();
  if (G.ctx) tmq_gauge_free(G.ctx);
  G.gauge_loaded = false;
}

// checkInvertParam + the drivers' own guards (lib/qudaQKXTM_interface.cpp:64-67, qkxtm/Calc_Loops.cpp:685-689,774-776)
static int matpc_of(const QudaInvertParam *p) { return (int)p->matpc_type; }
static void check_param(const QudaInvertParam *p) {
  if (!G.gauge_loaded) errorQuda("no gauge field resident (loadGaugeQuda)");
  if (p->dslash_type != QUDA_TWISTED_MASS_DSLASH && p->dslash_type != QUDA_TWISTED_CLOVER_DSLASH)
    errorQuda("This routine is for twisted mass or twisted clover operators only");        // qkxtm/Calc_Loops.cpp:685-689
  if (p->dslash_type == QUDA_TWISTED_CLOVER_DSLASH && !G.clover_loaded) errorQuda("twisted-clover needs loadCloverQuda first");
  if (p->dslash_type == QUDA_TWISTED_MASS_DSLASH && G.clover_loaded) errorQuda("a clover field is resident: call freeCloverQuda for plain twisted mass");
  if (p->gamma_basis != QUDA_UKQCD_GAMMA_BASIS) errorQuda("This function works only with ukqcd gamma basis");
  if (p->dirac_order != QUDA_DIRAC_ORDER) errorQuda("This function works only with colors inside the spins");
  if (p->cuda_prec != QUDA_DOUBLE_PRECISION) errorQuda("cuda_prec must be double (the QKXTM upload kernel writes double2, lib/qudaQKXTM_kernels.cu:1031)");
  if (p->sp_pad != 0) errorQuda("sp_pad must be 0 (lib/code_pieces/uploadToCuda_core.h:5)");
  if (p->kappa <= 0) errorQuda("kappa must be set");
}
static void check_solver(const QudaInvertParam *p) {
  check_param(p);
  if (p->inv_type != QUDA_CG_INVERTER) errorQuda("This path provides the CG inverter only (inv_type)");
  if (p->solve_type != QUDA_NORMOP_PC_SOLVE) errorQuda("CG requires a normal-operator pc solve (qkxtm/Calc_Loops.cpp:774-776)");
  if (p->solution_type != QUDA_MAT_SOLUTION) errorQuda("solution_type must be QUDA_MAT_SOLUTION (qkxtm/Calc_Loops.cpp:439)");
  if (p->tol <= 0 || p->maxiter <= 0) errorQuda("tol and maxiter must be positive");
}
// The reference's propagator drivers hard-wire GCR preconditioned by multigrid (qkxtm/MG_Bench.cpp:361,443; lib/qudaQKXTM_interface.cpp:32,
// 267 refuse anything else).  That solver is out of scope; the SYSTEM it solves is not: the entry points substitute CG on the even-odd
// normal operator for it -- the same solution to the same tolerance -- say so once, and restore the caller's parameters on return.
struct SolverSubstitution {
  QudaInvertParam *p;
  QudaInverterType inv;
  QudaSolveType solve;
  SolverSubstitution(QudaInvertParam *param, const char *who) : p(param), inv(param->inv_type), solve(param->solve_type) {
    if (p->inv_type != QUDA_GCR_INVERTER) return;
    static bool told = false;
    if (!told && G.rank == 0 && G.verbosity > QUDA_SILENT)
      fprintf(stderr, "WARNING: %s: GCR (+ multigrid) was requested; this library solves the same system with CG on M^dag M (even-odd, fp64 or fp32/fp64 mixed)\n", who);
    told = true;
    p->inv_type = QUDA_CG_INVERTER;
    p->solve_type = QUDA_NORMOP_PC_SOLVE;
  }
  ~SolverSubstitution() { p->inv_type = inv; p->solve_type = solve; }
};
// createDirac: the operator parameters live in the context
static void create_dirac(const QudaInvertParam *p) {
  const int m = matpc_of(p);
  if (G.op_kappa != p->kappa || G.op_mu != p->mu || G.op_matpc != m) {
    TMQ_OK(tmq_op_set(G.ctx, p->kappa, p->mu, m));
    G.op_kappa = p->kappa; G.op_mu = p->mu; G.op_matpc = m;
  }
}

// the solve of lib/qudaQKXTM_interface.cpp:2020-2041 on device fields: x = M_full^-1 b
static void solve_device(ColorSpinorField &x, ColorSpinorField &b, QudaInvertParam *param) {
  create_dirac(param);
  param->secs = 0; param->gflops = 0; param->iter = 0; param->true_res = 0;       // interface.cpp:95-97
  param->spinorGiB = (double)G.localVolume * 24 * 8 * 5 / (1024.0 * 1024.0 * 1024.0);
  // the three parity work fields live as long as the gauge field: the reference allocates them per call (interface.cpp:1923-1937), but a
  // cudaFree per solve is a device-wide synchronisation that would also stall the copies running behind it on the other streams
  for (int i = 0; i < 3; i++)
    if (!G.work[i]) G.work[i] = new ColorSpinorField(QUDA_PARITY_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
  ColorSpinorField &in = *G.work[0], &out = *G.work[1], &tmp = *G.work[2];
  const auto t0 = std::chrono::steady_clock::now();
  TMQ_OK(tmq_prepare(tmp.handle(), b.handle()));                                   // dirac.prepare            (:2020)
  TMQ_OK(tmq_matpc(in.handle(), tmp.handle(), 1));                                 // in <- M^dag in           (:2034)
  int iters = 0;
  double true_res = 0, secs = 0, gflops = 0;
  TMQ_OK(tmq_cg_mdagm(out.handle(), in.handle(), param->tol, param->maxiter, param->reliable_delta,
                      (int)param->cuda_prec_sloppy, &iters, &true_res, &secs, &gflops));   // (*solve)(*out,*in) (:2036)
  TMQ_OK(tmq_reconstruct(x.handle(), out.handle(), b.handle()));                   // dirac.reconstruct        (:2040)
  TMQ_OK(tmq_sync(G.ctx));
  param->iter = iters; param->true_res = true_res; param->gflops = gflops;         // updateInvertParam
  param->secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  double loop_secs = 0;
  int nrel = 0;
  tmq_cg_stats(G.ctx, &loop_secs, &nrel);
  if (G.verbosity >= QUDA_SUMMARIZE)
    printfQuda("CG: Convergence at %d iterations, L2 relative residual: true = %e (tol %e); %.3f secs (solver %.3f, its iteration loop %.3f, %d reliable updates), %.1f Gflops\n",
               iters, true_res, param->tol, param->secs, secs, loop_secs, nrel, gflops);
  if (iters >= param->maxiter && G.verbosity > QUDA_SILENT)
    fprintf(stderr, "WARNING: Exceeded maximum iterations %d\n", param->maxiter);   // warningQuda continues
}

void invertQuda(void *h_x, void *h_b, QudaInvertParam *param) {
  if (!h_x || !h_b || !param) errorQuda("null argument");
  check_solver(param);
  if (!G.io_b) G.io_b = new ColorSpinorField(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
  if (!G.io_x) G.io_x = new ColorSpinorField(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
  ColorSpinorField &b = *G.io_b, &x = *G.io_x;
  TMQ_OK(tmq_spinor_from_host(b.handle(), (const double *)h_b));
  solve_device(x, b, param);
  // x *= 2 kappa for the mass normalisations (lib/qudaQKXTM_interface.cpp:200-203), on the device
  if (param->mass_normalization == QUDA_MASS_NORMALIZATION || param->mass_normalization == QUDA_ASYMMETRIC_MASS_NORMALIZATION)
    TMQ_OK(tmq_ax(2.0 * param->kappa, x.handle()));
  TMQ_OK(tmq_spinor_to_host((double *)h_x, x.handle()));
}

// invertMultiSrcQuda (upstream quda.h): param->num_src right-hand sides, hp_b[k] -> hp_x[k], same host order as invertQuda.  The
// solves run one after the other on the compute stream; the upload of source k+1 and the download of solution k-1 run behind the
// solve of column k on their own streams (tmq_host_prefetch / tmq_spinor_to_host_async), so that in the steady state a column costs
// its solve and nothing else.  param->iter / secs / gflops are summed over the columns, true_res is the largest one.
void invertMultiSrcQuda(void **hp_x, void **hp_b, QudaInvertParam *param) {
  if (!hp_x || !hp_b || !param) errorQuda("null argument");
  const int n = param->num_src;
  if (n < 1) errorQuda("invertMultiSrcQuda: num_src = %d", n);
  for (int k = 0; k < n; k++) if (!hp_x[k] || !hp_b[k]) errorQuda("invertMultiSrcQuda: null host field %d", k);
  check_solver(param);
  const bool massnorm = param->mass_normalization == QUDA_MASS_NORMALIZATION || param->mass_normalization == QUDA_ASYMMETRIC_MASS_NORMALIZATION;
  const double scale = massnorm ? 2.0 * param->kappa : 1.0;
  if (getenv("TMQ_HOST_REGISTER") && atoi(getenv("TMQ_HOST_REGISTER")) > 0)      // page-lock the caller's buffers (kept until endQuda)
    for (int k = 0; k < n; k++) {
      TMQ_OK(tmq_host_register(G.ctx, hp_b[k], (size_t)G.localVolume * 24 * sizeof(double)));
      TMQ_OK(tmq_host_register(G.ctx, hp_x[k], (size_t)G.localVolume * 24 * sizeof(double)));
    }
  if (!G.io_b) G.io_b = new ColorSpinorField(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
  if (!G.io_x) G.io_x = new ColorSpinorField(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
  ColorSpinorField &b = *G.io_b, &x = *G.io_x;
  int iter = 0;
  double secs = 0, flops = 0, worst = 0;
  TMQ_OK(tmq_host_prefetch(G.ctx, 0, (const double *)hp_b[0]));
  for (int k = 0; k < n; k++) {
    if (k + 1 < n) TMQ_OK(tmq_host_prefetch(G.ctx, (k + 1) & 1, (const double *)hp_b[k + 1]));
    TMQ_OK(tmq_spinor_from_prefetch(b.handle(), k & 1, TMQ_HOST_ORDER_EO));
    solve_device(x, b, param);
    iter += param->iter; secs += param->secs; flops += param->gflops * param->secs;
    worst = param->true_res > worst ? param->true_res : worst;
    TMQ_OK(tmq_spinor_to_host_async((double *)hp_x[k], x.handle(), k & 1, TMQ_HOST_ORDER_EO, scale));
  }
  TMQ_OK(tmq_host_wait(G.ctx));
  param->iter = iter; param->secs = secs; param->gflops = secs > 0 ? flops / secs : 0; param->true_res = worst;
}

void MatQuda(void *h_out, void *h_in, QudaInvertParam *param) {
  if (!h_out || !h_in || !param) errorQuda("null argument");
  check_param(param);
  create_dirac(param);
  ColorSpinorField in(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION), out(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
  TMQ_OK(tmq_spinor_from_host(in.handle(), (const double *)h_in));
  TMQ_OK(tmq_mat_full(out.handle(), in.handle(), param->dagger == QUDA_DAG_YES ? 1 : 0));
  TMQ_OK(tmq_spinor_to_host((double *)h_out, out.handle()));
}

// ---- quda:: ---------------------------------------------------------------------------------------------------------------
namespace quda {

tmq_ctx *qkxtm_context() { return G.ctx; }
const int *qkxtm_local_extent() { return G.localL; }
long long qkxtm_local_volume() { return G.localVolume; }
void qkxtm_raise(const char *msg) { errorQuda("%s", msg); }
int qkxtm_Nmoms() { return (int)G.moms.size() / 3; }
const int *qkxtm_moms() { return G.moms.data(); }
void qkxtm_set_error_handler(qkxtm_error_handler h) { G.on_error = h; }

void init_qudaQKXTM(qudaQKXTMinfo *info) {
  if (G.qkxtm_initialized) return;                       // one-shot (lib/qudaQKXTM_kernels.cu:120,288)
  if (!info) errorQuda("null info");
  ensure_context(info->lL);
  G.nsmearGauss = info->nsmearGauss; G.alphaGauss = info->alphaGauss;      // GK_nsmearGauss, GK_alphaGauss (:125-127)
  // createMomenta(info->Q_sq) (lib/qudaQKXTM_kernels.cu:98-116,128): all integer momenta with p^2 <= Q_sq, shell by shell.  More than
  // MAX_NMOMENTA is an error in the reference; qkxtm/MG_Bench.cpp:534-544 never sets Q_sq (stack garbage), so the error is deferred to the
  // first call that needs momenta: here the list is left empty with a warning
  G.moms.clear();
  G.moms_overflow = false;
  for (int iQ = 0; iQ <= info->Q_sq && !G.moms_overflow; iQ++) {
    const int r = (int)std::floor(std::sqrt((double)iQ));
    for (int nx = r; nx >= -r && !G.moms_overflow; nx--)
      for (int ny = r; ny >= -r && !G.moms_overflow; ny--)
        for (int nz = r; nz >= -r; nz--)
          if (nx * nx + ny * ny + nz * nz == iQ) {
            if ((int)G.moms.size() / 3 >= MAX_NMOMENTA) { G.moms_overflow = true; break; }
            G.moms.push_back(nx); G.moms.push_back(ny); G.moms.push_back(nz);
          }
  }
  if (G.moms_overflow) {
    G.moms.clear();
    if (G.rank == 0 && G.verbosity > QUDA_SILENT)
      fprintf(stderr, "WARNING: init_qudaQKXTM: Q_sq = %d gives more than %d momenta (uninitialised info.Q_sq?): no momenta are kept, contractions will refuse\n", info->Q_sq, MAX_NMOMENTA);
  }
  // qkxtm/MG_Bench.cpp:534-544 and CalcLowModeProjection.cpp leave Nsources (and more) of their stack-allocated info unset; the reference
  // copies that many positions without a check (lib/qudaQKXTM_kernels.cu:129-132).  Out-of-range values are taken as "no sources".
  int nsrc = info->Nsources;
  if (nsrc < 0 || nsrc > MAX_NSOURCES) {
    if (G.rank == 0 && G.verbosity > QUDA_SILENT) fprintf(stderr, "WARNING: init_qudaQKXTM: info.Nsources = %d is out of range (uninitialised?): no source positions are kept\n", nsrc);
    nsrc = 0;
  }
  G.sourcePosition.assign(&info->sourcePosition[0][0], &info->sourcePosition[0][0] + (size_t)nsrc * 4);   // :129-132
  G.qkxtm_initialized = true;
  printfQuda("qudaQKXTM has been initialized\n");
}

void printf_qudaQKXTM() {
  printfQuda("Number of colors is %d\nNumber of spins is %d\nNumber of dimensions is %d\n", 3, 4, 4);
  printfQuda("Number of process in each direction is (x,y,z,t) %d x %d x %d x %d\n", G.grid[0], G.grid[1], G.grid[2], G.grid[3]);
  printfQuda("Local lattice is (x,y,z,t) %d x %d x %d x %d\n", G.localL[0], G.localL[1], G.localL[2], G.localL[3]);
  printfQuda("Local volume is %lld\n", G.localVolume);
}

// lib/qudaQKXTM_utils.cpp:87-109: the plaquette of the same links through two container objects
void testPlaquette(void **gauge) {
  for (int rep = 0; rep < 2; rep++) {
    QKXTM_Gauge<double> *gauge_object = new QKXTM_Gauge<double>(BOTH, GAUGE);
    gauge_object->printInfo();
    gauge_object->packGauge(gauge);
    gauge_object->loadGauge();
    gauge_object->calculatePlaq();
    delete gauge_object;
  }
}
// lib/qudaQKXTM_utils.cpp:116-141: Gaussian smearing of a point source at the origin; the 12 components at the origin are
// printed like the reference does
void testGaussSmearing(void **gauge) {
  QKXTM_Gauge<double> *gauge_object = new QKXTM_Gauge<double>(BOTH, GAUGE);
  gauge_object->printInfo();
  gauge_object->packGauge(gauge);
  gauge_object->loadGauge();
  gauge_object->calculatePlaq();
  QKXTM_Vector<double> *vecIn = new QKXTM_Vector<double>(BOTH, VECTOR);
  QKXTM_Vector<double> *vecOut = new QKXTM_Vector<double>(BOTH, VECTOR);
  double *input_vector = (double *)calloc((size_t)G.localVolume * 24, sizeof(double));   // (the reference leaves the rest uninitialised)
  if (!input_vector) errorQuda("Error allocating memory for the host source");
  input_vector[0] = 1.;
  vecIn->packVector(input_vector);
  vecIn->loadVector();
  vecOut->gaussianSmearing(*vecIn, *gauge_object);
  vecOut->download();
  for (int mu = 0; mu < 4; mu++)
    for (int c1 = 0; c1 < 3; c1++) printf("%+e %+e\n", vecOut->H_elem()[mu * 3 * 2 + c1 * 2 + 0], vecOut->H_elem()[mu * 3 * 2 + c1 * 2 + 1]);
  free(input_vector);
  delete vecOut;
  delete vecIn;
  delete gauge_object;
}

ColorSpinorField::ColorSpinorField(QudaSiteSubset subset, QudaPrecision prec) : h_(nullptr), subset_(subset) {
  if (!G.ctx) errorQuda("no context: call loadGaugeQuda / init_qudaQKXTM first");
  h_ = tmq_spinor_alloc(G.ctx, (int)prec, subset == QUDA_FULL_SITE_SUBSET ? TMQ_SUBSET_FULL : TMQ_SUBSET_PARITY);
  if (!h_) errorQuda("libtmq: %s", tmq_last_error());
}
ColorSpinorField::~ColorSpinorField() { if (h_ && G.ctx) tmq_spinor_free(h_); }
tmq_spinor *ColorSpinorField::Even() const { return tmq_spinor_even(h_); }
tmq_spinor *ColorSpinorField::Odd() const { return tmq_spinor_odd(h_); }

// ---- QKXTM_Field ------------------------------------------------------------------------------------------------------------
template <typename Float>
QKXTM_Field<Float>::QKXTM_Field(ALLOCATION_FLAG alloc_flag, CLASS_ENUM classT)
    : h_elem(NULL), h_elem_backup(NULL), d_elem(NULL), isAllocHost(false), isAllocDevice(false), isAllocHostBackup(false) {
  if (!G.qkxtm_initialized) errorQuda("You must initialize init_qudaQKXTM first");
  switch (classT) {
    case FIELD: field_length = 1; total_length = G.localVolume; break;
    case GAUGE: field_length = 4 * 3 * 3; total_length = G.localVolume; break;
    case VECTOR: field_length = 4 * 3; total_length = G.localVolume; break;
    case PROPAGATOR: field_length = 4 * 3 * 4 * 3; total_length = G.localVolume; break;
    case PROPAGATOR3D: field_length = 4 * 3 * 4 * 3; total_length = G.localVolume / G.localL[3]; break;
    case VECTOR3D: field_length = 4 * 3; total_length = G.localVolume / G.localL[3]; break;
  }
  bytes_total_length = (size_t)total_length * field_length * 2 * sizeof(Float);
  // ghost zones behind the local volume, one pair per partitioned dimension (lib/qudaQKXTM_Field.cpp:116-125); 4-d containers only
  ghost_length = (classT == PROPAGATOR3D || classT == VECTOR3D) ? 0 : (long long)tmq_qkxtm_ghost_sites(G.ctx);
  bytes_ghost_length = (size_t)ghost_length * field_length * 2 * sizeof(Float);
  bytes_total_plus_ghost_length = bytes_total_length + bytes_ghost_length;
  if (alloc_flag == BOTH) { create_host(); create_device(); }
  else if (alloc_flag == HOST) create_host();
  else if (alloc_flag == DEVICE) create_device();
  else if (alloc_flag == BOTH_EXTRA) { create_host(); create_host_backup(); create_device(); }
}
template <typename Float> QKXTM_Field<Float>::~QKXTM_Field() {
  if (h_elem != NULL) destroy_host();
  if (h_elem_backup != NULL) destroy_host_backup();
  if (d_elem != NULL) destroy_device();
}
template <typename Float> void QKXTM_Field<Float>::create_host() {
  h_elem = (Float *)malloc(bytes_total_length);
  if (h_elem == NULL) errorQuda("Error with allocation host memory");
  isAllocHost = true;
  zero_host();
}
template <typename Float> void QKXTM_Field<Float>::create_host_backup() {
  h_elem_backup = (Float *)malloc(bytes_total_length);
  if (h_elem_backup == NULL) errorQuda("Error with allocation host memory");
  isAllocHostBackup = true;
  zero_host_backup();
}
template <typename Float> void QKXTM_Field<Float>::create_device() {
  void *p = nullptr;
  TMQ_OK(tmq_dev_malloc(G.ctx, &p, bytes_total_plus_ghost_length));
  d_elem = (Float *)p;
  isAllocDevice = true;
  zero_device();
}
template <typename Float> void QKXTM_Field<Float>::destroy_host() { free(h_elem); h_elem = NULL; }
template <typename Float> void QKXTM_Field<Float>::destroy_host_backup() { free(h_elem_backup); h_elem_backup = NULL; }   // (the reference nulls h_elem here: App. C quirk, not replicated)
template <typename Float> void QKXTM_Field<Float>::destroy_device() {
  if (G.ctx) tmq_dev_free(G.ctx, d_elem);
  d_elem = NULL;
}
template <typename Float> void QKXTM_Field<Float>::zero_host() { memset(h_elem, 0, bytes_total_length); }
template <typename Float> void QKXTM_Field<Float>::zero_host_backup() { memset(h_elem_backup, 0, bytes_total_length); }
template <typename Float> void QKXTM_Field<Float>::zero_device() { TMQ_OK(tmq_dev_memset(G.ctx, d_elem, 0, bytes_total_plus_ghost_length)); }
// The ghost trio (include/qudaQKXTM.h:177-179,199-201,244-246; lib/qudaQKXTM_Gauge.cpp:143-373, lib/qudaQKXTM_Vector.cpp:172-382).  The reference
// gathers the boundary slices with cudaMemcpy2D into the host array, exchanges them over MPI and copies the received faces back; the
// contract is "after the three calls the ghost region behind the local volume holds the neighbours' slices".  Here the middle call does
// all of it on the device (tmq_qkxtm_exchange_ghost: gather kernel + ncclSend/Recv straight into the ghost region); the outer two have
// nothing left to do.
template <typename Float> void QKXTM_Field<Float>::exchange_ghost_device() {
  if (!isAllocDevice) errorQuda("the ghost exchange needs the device copy of the container");
  if (ghost_length == 0) return;
  TMQ_OK(tmq_qkxtm_exchange_ghost(G.ctx, d_elem, (int)sizeof(Float), field_length));
}
template <typename Float> void QKXTM_Field<Float>::printInfo() {
  printfQuda("GPU memory needed is %f MB \n", bytes_total_length / (1024.0 * 1024.0));
}

// ---- QKXTM_Gauge --------------------------------------------------------------------------------------------------------------
template <typename Float> QKXTM_Gauge<Float>::QKXTM_Gauge(ALLOCATION_FLAG a, CLASS_ENUM c) : QKXTM_Field<Float>(a, c) {}

template <typename Float> static void pack_gauge_into(Float *dst, void **gauge, long long V) {
  double **pg = (double **)gauge;
  for (int dir = 0; dir < 4; dir++)
#pragma omp parallel for
    for (long long iv = 0; iv < V; iv++)
      for (int c1 = 0; c1 < 3; c1++)
        for (int c2 = 0; c2 < 3; c2++)
          for (int part = 0; part < 2; part++)
            dst[(((size_t)dir * 3 + c1) * 3 + c2) * V * 2 + iv * 2 + part] = (Float)pg[dir][iv * 18 + c1 * 6 + c2 * 2 + part];
}
template <typename Float> void QKXTM_Gauge<Float>::packGauge(void **gauge) { pack_gauge_into(this->h_elem, gauge, this->total_length); }
template <typename Float> void QKXTM_Gauge<Float>::packGaugeToBackup(void **gauge) {
  if (this->h_elem_backup == NULL) errorQuda("Error you can call this method only if you allocate memory for h_elem_backup");
  pack_gauge_into(this->h_elem_backup, gauge, this->total_length);
}
template <typename Float> void QKXTM_Gauge<Float>::ghostToHost() {}
template <typename Float> void QKXTM_Gauge<Float>::cpuExchangeGhost() { this->exchange_ghost_device(); }
template <typename Float> void QKXTM_Gauge<Float>::ghostToDevice() {}
template <typename Float> void QKXTM_Gauge<Float>::loadGauge() { TMQ_OK(tmq_h2d(G.ctx, this->d_elem, this->h_elem, this->bytes_total_length)); }
template <typename Float> void QKXTM_Gauge<Float>::loadGaugeFromBackup() {
  if (this->h_elem_backup == NULL) errorQuda("Error you can call this method only if you allocate memory for h_elem_backup");
  TMQ_OK(tmq_h2d(G.ctx, this->d_elem, this->h_elem_backup, this->bytes_total_length));
}
template <typename Float> void QKXTM_Gauge<Float>::justDownloadGauge() { TMQ_OK(tmq_d2h(G.ctx, this->h_elem, this->d_elem, this->bytes_total_length)); }
template <typename Float> double QKXTM_Gauge<Float>::calculatePlaq() {
  double plaq = 0;
  ghostToHost();                      // lib/qudaQKXTM_Gauge.cpp:379-381
  cpuExchangeGhost();
  ghostToDevice();
  TMQ_OK(tmq_qkxtm_plaquette(G.ctx, this->d_elem, (int)sizeof(Float), &plaq));     // all-reduced over the ranks
  if (sizeof(Float) == 4) printfQuda("Calculated plaquette in single precision is %f\n", plaq);
  else printfQuda("Calculated plaquette in double precision is %lf\n", plaq);
  return plaq;
}

// ---- QKXTM_Vector --------------------------------------------------------------------------------------------------------------
template <typename Float> QKXTM_Vector<Float>::QKXTM_Vector(ALLOCATION_FLAG a, CLASS_ENUM c) : QKXTM_Field<Float>(a, c) {}

template <typename Float> void QKXTM_Vector<Float>::packVector(Float *vector) {
  const long long V = this->total_length;
  Float *h = this->h_elem;
#pragma omp parallel for
  for (long long iv = 0; iv < V; iv++)
    for (int sc = 0; sc < 12; sc++)       // always colours inside spins
      for (int part = 0; part < 2; part++) h[(size_t)sc * V * 2 + iv * 2 + part] = vector[iv * 24 + sc * 2 + part];
}
template <typename Float> void QKXTM_Vector<Float>::unpackVector(Float *vector) {
  const long long V = this->total_length;
  Float *h = this->h_elem;
#pragma omp parallel for
  for (long long iv = 0; iv < V; iv++)
    for (int sc = 0; sc < 12; sc++)
      for (int part = 0; part < 2; part++) h[iv * 24 + sc * 2 + part] = vector[(size_t)sc * V * 2 + iv * 2 + part];
}
template <typename Float> void QKXTM_Vector<Float>::unpackVector() {
  Float *tmp = (Float *)malloc(this->bytes_total_length);
  if (tmp == NULL) errorQuda("Error in allocate memory of tmp vector in unpackVector");
  memcpy(tmp, this->h_elem, this->bytes_total_length);
  unpackVector(tmp);
  free(tmp);
}
template <typename Float> void QKXTM_Vector<Float>::ghostToHost() {}
template <typename Float> void QKXTM_Vector<Float>::cpuExchangeGhost() { this->exchange_ghost_device(); }
template <typename Float> void QKXTM_Vector<Float>::ghostToDevice() {}
template <typename Float> void QKXTM_Vector<Float>::loadVector() { TMQ_OK(tmq_h2d(G.ctx, this->d_elem, this->h_elem, this->bytes_total_length)); }
template <typename Float> void QKXTM_Vector<Float>::unloadVector() { TMQ_OK(tmq_d2h(G.ctx, this->h_elem, this->d_elem, this->bytes_total_length)); }
template <typename Float> void QKXTM_Vector<Float>::download() { unloadVector(); unpackVector(); }

template <typename Float> void QKXTM_Vector<Float>::uploadToCuda(ColorSpinorField *cv, bool isEv) {
  if (!cv) errorQuda("null field");
  const int parity = cv->SiteSubset() == QUDA_FULL_SITE_SUBSET ? -1 : (isEv ? 0 : 1);
  TMQ_OK(tmq_spinor_from_qkxtm(cv->handle(), this->d_elem, (int)sizeof(Float), parity));
}
template <typename Float> void QKXTM_Vector<Float>::downloadFromCuda(ColorSpinorField *cv, bool isEv) {
  if (!cv) errorQuda("null field");
  const int parity = cv->SiteSubset() == QUDA_FULL_SITE_SUBSET ? -1 : (isEv ? 0 : 1);
  TMQ_OK(tmq_spinor_to_qkxtm(this->d_elem, (int)sizeof(Float), cv->handle(), parity, 1.0));
}
template <typename Float> void QKXTM_Vector<Float>::gaussianSmearing(QKXTM_Vector<Float> &vecIn, QKXTM_Gauge<Float> &gaugeAPE) {
  // GK_nsmearGauss / GK_alphaGauss come from init_qudaQKXTM (lib/qudaQKXTM_kernels.cu:125-127); no ghost exchange is
  // needed: the hop is 3-dimensional and the lattice is sharded along T
  TMQ_OK(tmq_qkxtm_gauss_smear(G.ctx, this->d_elem, vecIn.D_elem(), gaugeAPE.D_elem(), (int)sizeof(Float), G.nsmearGauss, G.alphaGauss));
}
template <typename Float> void QKXTM_Vector<Float>::write(char *filename) {
  // h_elem holds the host AoS vector (after download()), as in the reference (lib/qudaQKXTM_Vector.cpp:676-690)
  // rank 0 creates the file and writes the headers, then every rank writes its sub-block at its own offset.  The reference orders the
  // two steps with the MPI_Bcast of the payload offset (lib/qudaQKXTM_Vector.cpp:625); here: create -> barrier -> blocks -> barrier
  if (G.nranks > 1) {
    if (G.rank == 0 && tmq_lime_write_vector_header(filename, (int)sizeof(Float), G.localL, G.grid)) errorQuda("%s", tmq_lime_last_error());
    comm_barrier();
    if (tmq_lime_write_vector_block(filename, this->h_elem, (int)sizeof(Float), G.localL, G.grid, G.coord)) errorQuda("%s", tmq_lime_last_error());
    comm_barrier();
    return;
  }
  if (tmq_lime_write_vector(filename, this->h_elem, (int)sizeof(Float), G.localL, G.grid, G.coord)) errorQuda("%s", tmq_lime_last_error());
}
template <typename Float> void QKXTM_Vector<Float>::scaleVector(double a) { TMQ_OK(tmq_qkxtm_scale(G.ctx, this->d_elem, (int)sizeof(Float), a)); }
template <typename Float> void QKXTM_Vector<Float>::castDoubleToFloat(QKXTM_Vector<double> &in) {
  if (sizeof(Float) != 4) errorQuda("castDoubleToFloat needs a float vector");
  TMQ_OK(tmq_qkxtm_cast(G.ctx, this->d_elem, 4, in.D_elem(), 8));
}
template <typename Float> void QKXTM_Vector<Float>::castFloatToDouble(QKXTM_Vector<float> &in) {
  if (sizeof(Float) != 8) errorQuda("castFloatToDouble needs a double vector");
  TMQ_OK(tmq_qkxtm_cast(G.ctx, this->d_elem, 8, in.D_elem(), 4));
}
template <typename Float> double QKXTM_Vector<Float>::norm2Host() {
  double res = 0.0;
  const long long n = this->total_length * 24;
  for (long long i = 0; i < n; i++) res += (double)this->h_elem[i] * (double)this->h_elem[i];
  printfQuda("Vector norm2 is %e\n", res);
  return res;
}
template <typename Float> void QKXTM_Vector<Float>::apply_gamma5() { TMQ_OK(tmq_qkxtm_gamma5(G.ctx, this->d_elem, (int)sizeof(Float))); }
template <typename Float> void QKXTM_Vector<Float>::conjugate() { TMQ_OK(tmq_qkxtm_conjugate(G.ctx, this->d_elem, (int)sizeof(Float), 12)); }
template <typename Float> void QKXTM_Vector<Float>::copyPropagator(QKXTM_Propagator<Float> &prop, int nu, int c2) {
  const long long V = this->total_length;
  TMQ_OK(tmq_qkxtm_column_copy(G.ctx, prop.D_elem(), V, 0, this->d_elem, V, 0, V, (int)sizeof(Float), nu, c2, 0));
}
template <typename Float> void QKXTM_Vector<Float>::copyPropagator3D(QKXTM_Propagator3D<Float> &prop, int timeslice, int nu, int c2) {
  const long long V = this->total_length, V3 = V / G.localL[3];
  if (timeslice < 0 || timeslice >= G.localL[3]) errorQuda("time slice %d outside the local lattice", timeslice);
  TMQ_OK(tmq_qkxtm_column_copy(G.ctx, prop.D_elem(), V3, 0, this->d_elem, V, (long long)timeslice * V3, V3, (int)sizeof(Float), nu, c2, 0));
}

// ---- QKXTM_Propagator ------------------------------------------------------------------------------------------------------------
template <typename Float> QKXTM_Propagator<Float>::QKXTM_Propagator(ALLOCATION_FLAG a, CLASS_ENUM c) : QKXTM_Field<Float>(a, c) {}
template <typename Float> void QKXTM_Propagator<Float>::ghostToHost() {}
template <typename Float> void QKXTM_Propagator<Float>::cpuExchangeGhost() { this->exchange_ghost_device(); }
template <typename Float> void QKXTM_Propagator<Float>::ghostToDevice() {}
template <typename Float> void QKXTM_Propagator<Float>::absorbVectorToDevice(QKXTM_Vector<Float> &vec, int nu, int c2) {
  TMQ_OK(tmq_qkxtm_absorb(G.ctx, this->d_elem, vec.D_elem(), (int)sizeof(Float), nu, c2));
}
template <typename Float> void QKXTM_Propagator<Float>::absorbVectorToHost(QKXTM_Vector<Float> &vec, int nu, int c2) {
  const size_t V = (size_t)this->total_length;
  for (int mu = 0; mu < 4; mu++)
    for (int c1 = 0; c1 < 3; c1++)
      TMQ_OK(tmq_d2h(G.ctx, this->h_elem + (((size_t)mu * 4 + nu) * 9 + c1 * 3 + c2) * V * 2, vec.D_elem() + ((size_t)mu * 3 + c1) * V * 2,
                     V * 2 * sizeof(Float)));
}

template <typename Float> void QKXTM_Propagator<Float>::rotateToPhysicalBase_device(int sign) {
  if ((sign != +1) && (sign != -1)) errorQuda("The sign can be only +-1");
  TMQ_OK(tmq_qkxtm_rotate_physical(G.ctx, this->d_elem, (int)sizeof(Float), sign));
}
template <typename Float> void QKXTM_Propagator<Float>::rotateToPhysicalBase_host(int sign) {
  if ((sign != +1) && (sign != -1)) errorQuda("The sign can be only +-1");
  // P <- 1/2 (1 + i s g5) P (1 + i s g5) with g5 the UKQCD spin swap, per site and colour pair
  const size_t V = (size_t)this->total_length;
  const Float s = (Float)sign;
  for (size_t x = 0; x < V; x++)
    for (int cc = 0; cc < 9; cc++) {
      Float P[16][2], R[16][2];
      for (int k = 0; k < 16; k++) { const Float *p = this->h_elem + (((size_t)k * 9 + cc) * V + x) * 2; P[k][0] = p[0]; P[k][1] = p[1]; }
      for (int a = 0; a < 4; a++)
        for (int g = 0; g < 4; g++) {
          const int k = a * 4 + g, ka = (a ^ 2) * 4 + g, kg = a * 4 + (g ^ 2), kag = (a ^ 2) * 4 + (g ^ 2);
          R[k][0] = (Float)0.5 * (P[k][0] - s * P[ka][1] - s * P[kg][1] - P[kag][0]);
          R[k][1] = (Float)0.5 * (P[k][1] + s * P[ka][0] + s * P[kg][0] - P[kag][1]);
        }
      for (int k = 0; k < 16; k++) { Float *p = this->h_elem + (((size_t)k * 9 + cc) * V + x) * 2; p[0] = R[k][0]; p[1] = R[k][1]; }
    }
}
template <typename Float> void QKXTM_Propagator<Float>::conjugate() { TMQ_OK(tmq_qkxtm_conjugate(G.ctx, this->d_elem, (int)sizeof(Float), 144)); }
template <typename Float> void QKXTM_Propagator<Float>::apply_gamma5() { TMQ_OK(tmq_qkxtm_gamma5_prop(G.ctx, this->d_elem, (int)sizeof(Float))); }

// ---- QKXTM_Propagator3D -------------------------------------------------------------------------------------------------------
template <typename Float> QKXTM_Propagator3D<Float>::QKXTM_Propagator3D(ALLOCATION_FLAG a, CLASS_ENUM c) : QKXTM_Field<Float>(a, c) {
  if (c != PROPAGATOR3D) errorQuda("QKXTM_Propagator3D needs the class PROPAGATOR3D");
}
template <typename Float> void QKXTM_Propagator3D<Float>::absorbTimeSlice(QKXTM_Propagator<Float> &prop, int timeslice) {
  const long long V3 = this->total_length, V = V3 * G.localL[3];
  if (timeslice < 0 || timeslice >= G.localL[3]) errorQuda("time slice %d outside the local lattice", timeslice);
  // the 144 components are 12 columns of 12: a propagator is a vector-like array of columns for this purpose
  for (int nu = 0; nu < 4; nu++)
    for (int c2 = 0; c2 < 3; c2++)
      for (int mu = 0; mu < 4; mu++)
        for (int c1 = 0; c1 < 3; c1++) {
          const size_t k = ((size_t)(mu * 4 + nu) * 9 + c1 * 3 + c2);
          TMQ_OK(tmq_d2d(G.ctx, this->d_elem + k * V3 * 2, prop.D_elem() + (k * V + (size_t)timeslice * V3) * 2, (size_t)V3 * 2 * sizeof(Float)));
        }
}
template <typename Float> void QKXTM_Propagator3D<Float>::absorbVectorTimeSlice(QKXTM_Vector<Float> &vec, int timeslice, int nu, int c2) {
  const long long V3 = this->total_length, V = V3 * G.localL[3];
  if (timeslice < 0 || timeslice >= G.localL[3]) errorQuda("time slice %d outside the local lattice", timeslice);
  TMQ_OK(tmq_qkxtm_column_copy(G.ctx, this->d_elem, V3, 0, vec.D_elem(), V, (long long)timeslice * V3, V3, (int)sizeof(Float), nu, c2, 1));
}

// The writers' gather over the time communicator (the reference: MPI_Gather of every rank's buffer, lib/qudaQKXTM_Contraction.cpp:1580-1584):
// the CALLER's buffer [Lt][per_t] -- already summed over the ranks that share its time slices -- is placed at this rank's time offset of a
// zero-padded global-T array and the arrays are summed over all ranks; only the ranks at z coordinate 0 contribute.
template <typename Float> static std::vector<Float> gather_over_t(const Float *local, size_t per_t) {
  const int Lt = G.localL[3], T = Lt * G.grid[3];
  std::vector<double> g((size_t)T * per_t, 0.0);
  if (G.coord[0] == 0 && G.coord[1] == 0 && G.coord[2] == 0)
    for (size_t i = 0; i < (size_t)Lt * per_t; i++) g[(size_t)G.coord[3] * Lt * per_t + i] = (double)local[i];
  TMQ_OK(tmq_allreduce_host(G.ctx, g.data(), g.size()));
  return std::vector<Float>(g.begin(), g.end());
}

// ---- QKXTM_Contraction (mesons) -----------------------------------------------------------------------------------------------
template <typename Float>
void QKXTM_Contraction<Float>::contractMesons(QKXTM_Propagator<Float> &prop1, QKXTM_Propagator<Float> &prop2, void *corrMesons, int isource,
                                              CORR_SPACE CorrSpace) {
  if (!corrMesons) errorQuda("null correlator buffer");
  if (isource < 0 || (size_t)isource * 4 >= G.sourcePosition.size()) errorQuda("source %d was not given to init_qudaQKXTM", isource);
  printfQuda("contractMesons: Will perform in %s precision\n", sizeof(Float) == 4 ? "single" : "double");
  Float *out = (Float *)corrMesons;
  const long long V = G.localVolume;
  const int Lt = G.localL[3], nm = qkxtm_Nmoms();
  if (CorrSpace == POSITION_SPACE) {
    std::vector<double> pos((size_t)V * 40);
    TMQ_OK(tmq_qkxtm_contract_mesons(G.ctx, prop1.D_elem(), prop2.D_elem(), (int)sizeof(Float), NULL, 0, NULL, NULL, pos.data()));
    for (long long x = 0; x < V; x++)
      for (int ch = 0; ch < 20; ch++)
        for (int ri = 0; ri < 2; ri++) out[((size_t)2 * x + ri) * 20 + ch] = (Float)pos[((size_t)x * 20 + ch) * 2 + ri];
  } else if (CorrSpace == MOMENTUM_SPACE) {
    if (G.moms_overflow) errorQuda("Error exceeded max number of momenta");     // lib/qudaQKXTM_kernels.cu:113
    if (nm <= 0) errorQuda("no momenta: init_qudaQKXTM was given Q_sq < 0");
    const int gT = Lt * G.grid[3];
    std::vector<double> mom((size_t)gT * nm * 40);
    TMQ_OK(tmq_qkxtm_contract_mesons(G.ctx, prop1.D_elem(), prop2.D_elem(), (int)sizeof(Float), G.moms.data(), nm, &G.sourcePosition[(size_t)isource * 4],
                                     mom.data(), NULL));
    for (int it = 0; it < Lt; it++)
      for (int im = 0; im < nm; im++)
        for (int ch = 0; ch < 20; ch++)
          for (int ri = 0; ri < 2; ri++)
            out[(((size_t)it * nm + im) * 2 + ri) * 20 + ch] = (Float)mom[((((size_t)(it + G.coord[3] * Lt)) * nm + im) * 20 + ch) * 2 + ri];
  } else errorQuda("contractMesons: Supports only POSITION_SPACE and MOMENTUM_SPACE!");
}

template <typename Float>
void QKXTM_Contraction<Float>::writeTwopMesons_ASCII(void *corrMesons, char *filename_out, int isource, CORR_SPACE CorrSpace) {
  if (CorrSpace != MOMENTUM_SPACE) errorQuda("writeTwopMesons_ASCII: Supports writing only in momentum-space!");
  printfQuda("writeTwopMesons_ASCII: Will write in %s precision\n", sizeof(Float) == 4 ? "single" : "double");
  const int nm = qkxtm_Nmoms(), T = G.localL[3] * G.grid[3];
  // on a t split the caller's buffers are gathered over the time ranks (every rank calls this)
  std::vector<Float> gathered;
  const Float *c = (const Float *)corrMesons;
  if (G.grid[3] != 1) { gathered = gather_over_t(c, (size_t)nm * 40); c = gathered.data(); }
  const int *mv = qkxtm_moms();
  bool root = true;
  for (int d = 0; d < 4; d++) root = root && G.coord[d] == 0;
  if (!root) { comm_barrier(); return; }       // the writing rank joins when its file is complete (the reference has an MPI_Gather here)
  FILE *ptr_out = fopen(filename_out, "w");
  if (ptr_out == NULL) errorQuda("Error opening file for writing");
  for (int ip = 0; ip < 10; ip++)
    for (int it = 0; it < T; it++)
      for (int imom = 0; imom < nm; imom++) {
        const int it_shift = (it + G.sourcePosition[(size_t)isource * 4 + 3]) % T;
        const size_t b = ((size_t)it_shift * nm + imom) * 2;
        fprintf(ptr_out, "%d \t %d \t %+d %+d %+d \t %+e %+e \t %+e %+e\n", ip, it, mv[3 * imom], mv[3 * imom + 1], mv[3 * imom + 2],
                (double)c[(b + 0) * 20 + ip], (double)c[(b + 1) * 20 + ip], (double)c[(b + 0) * 20 + 10 + ip], (double)c[(b + 1) * 20 + 10 + ip]);
      }
  fclose(ptr_out);
  comm_barrier();
}

template <typename Float>
void QKXTM_Contraction<Float>::contractBaryons(QKXTM_Propagator<Float> &prop1, QKXTM_Propagator<Float> &prop2, void *corrBaryons, int isource,
                                               CORR_SPACE CorrSpace) {
  if (!corrBaryons) errorQuda("null correlator buffer");
  if (isource < 0 || (size_t)isource * 4 >= G.sourcePosition.size()) errorQuda("source %d was not given to init_qudaQKXTM", isource);
  if (CorrSpace != MOMENTUM_SPACE) errorQuda("contractBaryons: only MOMENTUM_SPACE is built");
  printfQuda("contractBaryons: Will perform in %s precision\n", sizeof(Float) == 4 ? "single" : "double");
  Float *out = (Float *)corrBaryons;
  const int Lt = G.localL[3], nm = qkxtm_Nmoms(), gT = Lt * G.grid[3];
  if (nm <= 0) errorQuda("no momenta: init_qudaQKXTM was given Q_sq < 0");
  std::vector<double> mom((size_t)gT * nm * 320 * 2);
  TMQ_OK(tmq_qkxtm_contract_baryons(G.ctx, prop1.D_elem(), prop2.D_elem(), (int)sizeof(Float), G.moms.data(), nm, &G.sourcePosition[(size_t)isource * 4],
                                    mom.data()));
  for (int it = 0; it < Lt; it++)
    for (int im = 0; im < nm; im++)
      for (int ch = 0; ch < 320; ch++)
        for (int ri = 0; ri < 2; ri++)
          out[(((size_t)it * nm + im) * 2 + ri) * 320 + ch] = (Float)mom[((((size_t)(it + G.coord[3] * Lt)) * nm + im) * 320 + ch) * 2 + ri];
}

template <typename Float>
void QKXTM_Contraction<Float>::writeTwopBaryons_ASCII(void *corrBaryons, char *filename_out, int isource, CORR_SPACE CorrSpace) {
  if (CorrSpace != MOMENTUM_SPACE) errorQuda("writeTwopBaryons_ASCII: Supports writing only in momentum-space!");
  printfQuda("writeTwopBaryons_ASCII: Will write in %s precision\n", sizeof(Float) == 4 ? "single" : "double");
  const int nm = qkxtm_Nmoms(), T = G.localL[3] * G.grid[3];
  std::vector<Float> gathered;
  const Float *c = (const Float *)corrBaryons;
  if (G.grid[3] != 1) { gathered = gather_over_t(c, (size_t)nm * 640); c = gathered.data(); }
  const int *mv = qkxtm_moms();
  bool root = true;
  for (int d = 0; d < 4; d++) root = root && G.coord[d] == 0;
  if (!root) { comm_barrier(); return; }       // the writing rank joins when its file is complete (the reference has an MPI_Gather here)
  FILE *ptr_out = fopen(filename_out, "w");
  if (ptr_out == NULL) errorQuda("Error opening file for writing");
  const int tsrc = G.sourcePosition[(size_t)isource * 4 + 3];
  for (int ip = 0; ip < 10; ip++)
    for (int it = 0; it < T; it++)
      for (int imom = 0; imom < nm; imom++)
        for (int gamma = 0; gamma < 4; gamma++)
          for (int gammap = 0; gammap < 4; gammap++) {
            const int it_shift = (it + tsrc) % T;
            const int sign = (it + tsrc) >= T ? -1 : +1;       // the baryon picks up the anti-periodic boundary sign when the time wraps
            const size_t b = ((size_t)it_shift * nm + imom) * 2, k = (size_t)ip * 16 + gamma * 4 + gammap;
            fprintf(ptr_out, "%d \t %d \t %+d %+d %+d \t %d %d \t %+e %+e \t %+e %+e\n", ip, it, mv[3 * imom], mv[3 * imom + 1], mv[3 * imom + 2],
                    gamma, gammap, sign * (double)c[(b + 0) * 320 + k], sign * (double)c[(b + 1) * 320 + k], sign * (double)c[(b + 0) * 320 + 160 + k],
                    sign * (double)c[(b + 1) * 320 + 160 + k]);
          }
  fclose(ptr_out);
  comm_barrier();
}

template <typename Float>
void QKXTM_Contraction<Float>::seqSourceFixSinkPart1(QKXTM_Vector<Float> &vec, QKXTM_Propagator3D<Float> &prop1, QKXTM_Propagator3D<Float> &prop2,
                                                     int timeslice, int nu, int c2, WHICHPROJECTOR PID, WHICHPARTICLE testParticle) {
  TMQ_OK(tmq_qkxtm_seq_source(G.ctx, vec.D_elem(), timeslice, prop1.D_elem(), prop2.D_elem(), (int)sizeof(Float), nu, c2, (int)PID, (int)testParticle, 1));
}
template <typename Float>
void QKXTM_Contraction<Float>::seqSourceFixSinkPart2(QKXTM_Vector<Float> &vec, QKXTM_Propagator3D<Float> &prop, int timeslice, int nu, int c2,
                                                     WHICHPROJECTOR PID, WHICHPARTICLE testParticle) {
  TMQ_OK(tmq_qkxtm_seq_source(G.ctx, vec.D_elem(), timeslice, prop.D_elem(), NULL, (int)sizeof(Float), nu, c2, (int)PID, (int)testParticle, 2));
}
template <typename Float>
void QKXTM_Contraction<Float>::contractFixSink(QKXTM_Propagator<Float> &seqProp, QKXTM_Propagator<Float> &prop, QKXTM_Gauge<Float> &gauge,
                                               void *corrThp_local, void *corrThp_noether, void *corrThp_oneD, WHICHPROJECTOR typeProj,
                                               WHICHPARTICLE testParticle, int partflag, int isource, CORR_SPACE CorrSpace) {
  (void)typeProj;
  if (!corrThp_local) errorQuda("null correlator buffer");
  if ((corrThp_noether == NULL) != (corrThp_oneD == NULL)) errorQuda("contractFixSink: give both the Noether and the one-derivative buffer, or neither");
  if (CorrSpace != MOMENTUM_SPACE) errorQuda("contractFixSink: only MOMENTUM_SPACE is built");
  if (isource < 0 || (size_t)isource * 4 >= G.sourcePosition.size()) errorQuda("source %d was not given to init_qudaQKXTM", isource);
  printfQuda("contractFixSink: Will perform in %s precision\n", sizeof(Float) == 4 ? "single" : "double");
  const int Lt = G.localL[3], nm = qkxtm_Nmoms(), gT = Lt * G.grid[3];
  if (nm <= 0) errorQuda("no momenta: init_qudaQKXTM was given Q_sq < 0");
  std::vector<double> mom((size_t)gT * nm * 16 * 2);
  TMQ_OK(tmq_qkxtm_fixsink_local(G.ctx, seqProp.D_elem(), prop.D_elem(), (int)sizeof(Float), (int)testParticle, partflag, G.moms.data(), nm,
                                 &G.sourcePosition[(size_t)isource * 4], mom.data()));
  Float *out = (Float *)corrThp_local;
  for (size_t i = 0; i < (size_t)Lt * nm * 32; i++) out[i] = (Float)mom[(size_t)G.coord[3] * Lt * nm * 32 + i];
  if (corrThp_noether) {
    if (!gauge.D_elem()) errorQuda("contractFixSink: the Noether and one-derivative insertions need the gauge links on the device");
    std::vector<double> cn((size_t)Lt * nm * 4 * 2), co((size_t)Lt * nm * 64 * 2);
    TMQ_OK(tmq_qkxtm_fixsink_derivative(G.ctx, seqProp.D_elem(), prop.D_elem(), gauge.D_elem(), (int)sizeof(Float), (int)testParticle, partflag,
                                        G.moms.data(), nm, &G.sourcePosition[(size_t)isource * 4], cn.data(), co.data()));
    for (size_t i = 0; i < cn.size(); i++) ((Float *)corrThp_noether)[i] = (Float)cn[i];
    for (size_t i = 0; i < co.size(); i++) ((Float *)corrThp_oneD)[i] = (Float)co[i];
  }
}
template <typename Float>
void QKXTM_Contraction<Float>::writeThrp_ASCII(void *corrThp_local, void *corrThp_noether, void *corrThp_oneD, WHICHPARTICLE testParticle, int partflag,
                                               char *filename_out, int isource, int tsinkMtsource, CORR_SPACE CorrSpace) {
  if (CorrSpace != MOMENTUM_SPACE) errorQuda("writeThrp_ASCII: Supports writing only in momentum-space!");
  if ((corrThp_noether == NULL) != (corrThp_oneD == NULL)) errorQuda("writeThrp_ASCII: give both the Noether and the one-derivative buffer, or neither");
  if (partflag != 1 && partflag != 2) errorQuda("writeThrp_ASCII: Got the wrong part! Should be either 1 or 2.");
  printfQuda("writeThrp_ASCII: Will write in %s precision\n", sizeof(Float) == 4 ? "single" : "double");
  const char *particle = testParticle == PROTON ? "proton" : "neutron";
  const char *flavour = (testParticle == PROTON) == (partflag == 1) ? "up" : "down";          // Contraction.cpp:2903-2914
  const int *sp = &G.sourcePosition[(size_t)isource * 4];
  char fname_local[1024], fname_noether[1024], fname_oneD[1024];
  snprintf(fname_local, sizeof(fname_local), "%s.%s.%s.%s.SS.%02d.%02d.%02d.%02d.dat", filename_out, particle, flavour, "ultra_local", sp[0], sp[1], sp[2], sp[3]);
  snprintf(fname_noether, sizeof(fname_noether), "%s.%s.%s.%s.SS.%02d.%02d.%02d.%02d.dat", filename_out, particle, flavour, "noether", sp[0], sp[1], sp[2], sp[3]);
  snprintf(fname_oneD, sizeof(fname_oneD), "%s.%s.%s.%s.SS.%02d.%02d.%02d.%02d.dat", filename_out, particle, flavour, "oneD", sp[0], sp[1], sp[2], sp[3]);
  const int nm = qkxtm_Nmoms(), T = G.localL[3] * G.grid[3];
  const int *mv = qkxtm_moms();
  std::vector<Float> gathered;
  const Float *c = (const Float *)corrThp_local;
  if (G.grid[3] != 1) { gathered = gather_over_t(c, (size_t)nm * 32); c = gathered.data(); }      // collective: before the root test
  bool root = true;
  for (int d = 0; d < 4; d++) root = root && G.coord[d] == 0;
  if (!root) { comm_barrier(); return; }       // the writing rank joins when its file is complete (the reference has an MPI_Gather here)
  FILE *ptr_local = fopen(fname_local, "w");
  if (ptr_local == NULL) errorQuda("Error opening file for writing");
  for (int iop = 0; iop < 16; iop++)
    for (int it = 0; it < T; it++)
      for (int imom = 0; imom < nm; imom++) {
        const int it_shift = (it + sp[3]) % T;
        const int sign = (tsinkMtsource + sp[3]) >= T ? -1 : +1;            // the sink lies beyond the anti-periodic boundary
        const size_t k = ((size_t)it_shift * nm + imom) * 32 + iop * 2;
        fprintf(ptr_local, "%d \t %d \t %+d %+d %+d \t %+e %+e\n", iop, it, mv[3 * imom], mv[3 * imom + 1], mv[3 * imom + 2], sign * (double)c[k],
                sign * (double)c[k + 1]);
      }
  fclose(ptr_local);
  if (!corrThp_noether) { comm_barrier(); return; }
  const Float *cn = (const Float *)corrThp_noether, *co = (const Float *)corrThp_oneD;
  FILE *ptr_noether = fopen(fname_noether, "w"), *ptr_oneD = fopen(fname_oneD, "w");
  if (ptr_noether == NULL || ptr_oneD == NULL) errorQuda("Error opening file for writing");
  const int sign = (tsinkMtsource + sp[3]) >= T ? -1 : +1;
  for (int iop = 0; iop < 4; iop++)                                        // :2960-2975 (iop = direction)
    for (int it = 0; it < T; it++)
      for (int imom = 0; imom < nm; imom++) {
        const size_t k = ((size_t)((it + sp[3]) % T) * nm + imom) * 8 + iop * 2;
        fprintf(ptr_noether, "%d \t %d \t %+d %+d %+d \t %+e %+e\n", iop, it, mv[3 * imom], mv[3 * imom + 1], mv[3 * imom + 2], sign * (double)cn[k],
                sign * (double)cn[k + 1]);
      }
  for (int iop = 0; iop < 16; iop++)                                       // :2976-2995
    for (int dir = 0; dir < 4; dir++)
      for (int it = 0; it < T; it++)
        for (int imom = 0; imom < nm; imom++) {
          const size_t k = ((size_t)((it + sp[3]) % T) * nm + imom) * 128 + dir * 32 + iop * 2;
          fprintf(ptr_oneD, "%d \t %d \t %d \t %+d %+d %+d \t %+e %+e\n", iop, dir, it, mv[3 * imom], mv[3 * imom + 1], mv[3 * imom + 2],
                  sign * (double)co[k], sign * (double)co[k + 1]);
        }
  fclose(ptr_noether);
  fclose(ptr_oneD);
  comm_barrier();
}

// ---- QKXTM_Deflation ----------------------------------------------------------------------------------------------------
template <typename Float>
QKXTM_Deflation<Float>::QKXTM_Deflation(QudaInvertParam *param, qudaQKXTM_arpackInfo ai)
    : eigenValues(NULL), residuals(NULL), set(NULL), nconv(0), nrestarts(0), nmatvec(0) {
  if (!G.qkxtm_initialized) errorQuda("You must initialize QKXTM library first");
  if (typeid(Float) != typeid(double)) errorQuda("Single precision is not implemented in the eigensolver (Deflation.cpp:1001)");
  PolyDeg = ai.PolyDeg; NeV = ai.nEv; NkV = ai.nKv; spectrumPart = ai.spectrumPart; isACC = ai.isACC;
  tolArpack = ai.tolArpack; maxIterArpack = ai.maxIterArpack; amin = ai.amin; amax = ai.amax;
  isEv = ai.isEven; isFullOp = ai.isFullOp; flavor_sign = param->mu; invert_param = param;
  total_length_per_NeV = 0; bytes_total_length_per_NeV = 0;
  if (NeV == 0) { printfQuda("######### Got NeV = 0 #########\n"); return; }

  if (spectrumPart != SR && spectrumPart != LR && spectrumPart != SM && spectrumPart != LM)
    errorQuda("eigenSolver: Option for spectrumPart is suspicious");
  check_param(param);
  if (!isFullOp && ((int)param->matpc_type & 1) != (isEv ? 0 : 1)) errorQuda("matpc_type does not match arpackInfo.isEven");
  invert_param->solve_type = isFullOp ? QUDA_NORMOP_SOLVE : QUDA_NORMOP_PC_SOLVE;     // Deflation.cpp:132-133
  total_length_per_NeV = (G.localVolume / (isFullOp ? 1 : 2)) * 4 * 3 * 2;
  bytes_total_length_per_NeV = (size_t)total_length_per_NeV * sizeof(Float);
  eigenValues = (Float *)calloc((size_t)2 * NkV, sizeof(Float));
  residuals = (double *)calloc((size_t)NkV, sizeof(double));
  if (!eigenValues || !residuals) errorQuda("Error: Out of memory of eigenValues.");
  set = tmq_eigset_alloc(G.ctx, NkV + 1, (int)sizeof(Float), isFullOp ? TMQ_SUBSET_FULL : TMQ_SUBSET_PARITY);
  if (!set) errorQuda("libtmq: %s", tmq_last_error());
}
template <typename Float> QKXTM_Deflation<Float>::~QKXTM_Deflation() {
  if (set) tmq_eigset_free(set);
  free(eigenValues);
  free(residuals);
}
template <typename Float> void QKXTM_Deflation<Float>::printInfo() {
  printfQuda("\n======= DEFLATION INFO =======\n");
  if (isFullOp) printfQuda(" The EigenVectors are for the Full %smu operator\n", (flavor_sign > 0) ? "+" : "-");
  else printfQuda(" Will calculate EigenVectors for the %s %smu operator\n", isEv ? "even-even" : "odd-odd", (flavor_sign > 0) ? "+" : "-");
  printfQuda(" Number of requested EigenVectors is %d in precision %d\n", NeV, (int)sizeof(Float));
  printfQuda(" The Size of Krylov space is %d\n", NkV);
  printfQuda(" Device GB for the Krylov space: %lf\n", (NkV + 1) * ((double)bytes_total_length_per_NeV / (1024. * 1024. * 1024.)));
  printfQuda("==============================\n");
}
template <typename Float> void QKXTM_Deflation<Float>::eigenSolver() {
  if (NeV == 0) { printfQuda("eigenSolver: Got NeV=%d. Returning...\n", NeV); return; }
  create_dirac(invert_param);
  printfQuda("\neigenSolver: Input to the Lanczos solver\n========================================\n");
  printfQuda(" Number of Ritz eigenvalues requested: %d\n Size of Krylov space is: %d\n", NeV, NkV);
  printfQuda(" Polynomial acceleration: %s\n", isACC ? "yes" : "no");
  if (isACC) printfQuda(" Chebyshev polynomial paramaters: Degree = %d, amin = %+e, amax = %+e\n", PolyDeg, amin, amax);
  printfQuda(" The convergence criterion is %+e\n Maximum number of restarts is %d\n========================================\n\n",
             tolArpack, maxIterArpack);
  const int which = (spectrumPart == SR || spectrumPart == SM) ? 0 : 1;
  std::vector<double> ev(NeV), rs(NeV);
  const auto t0 = std::chrono::steady_clock::now();
  TMQ_OK(tmq_eigensolve(set, NeV, NkV, isACC ? PolyDeg : 0, amin, amax, tolArpack, maxIterArpack, which, 1234ull, ev.data(),
                        rs.data(), &nconv, &nrestarts, &nmatvec));
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printfQuda("eigenSolver: Number of converged eigenvalues: %d (restarts %d, operator applications %d)\n", nconv, nrestarts, nmatvec);
  printfQuda("eigenSolver: TIME_REPORT - Eigenvalue calculation: %f sec\n", secs);
  printfQuda("Eigenvalues of the %s Dirac operator:\n===========\n", isFullOp ? "Full" : "Even-Odd");
  for (int i = 0; i < NeV; i++) {
    eigenValues[2 * i] = (Float)ev[i]; eigenValues[2 * i + 1] = 0;
    residuals[i] = rs[i];
    printfQuda("Eval[%04d] = %+e  %+e    Residual: %+e\n", i, ev[i], 0.0, rs[i]);
  }
}
template <typename Float> void QKXTM_Deflation<Float>::polynomialOperator(ColorSpinorField &out, const ColorSpinorField &in) {
  create_dirac(invert_param);
  TMQ_OK(tmq_poly_mdagm(out.handle(), in.handle(), PolyDeg, amin, amax));
}
template <typename Float> void QKXTM_Deflation<Float>::deflateVector(QKXTM_Vector<Float> &vec_defl, QKXTM_Vector<Float> &vec_in) {
  if (NeV == 0) { vec_defl.zero_device(); return; }
  // vec_in.H_elem() holds the host AoS vector (Deflation.cpp:640-668); the parity block isEv is deflated, the other
  // parity of the result is zero
  QKXTM_Vector<Float> stage(BOTH, VECTOR);
  stage.packVector(vec_in.H_elem());
  stage.loadVector();
  const QudaSiteSubset sub = isFullOp ? QUDA_FULL_SITE_SUBSET : QUDA_PARITY_SITE_SUBSET;
  ColorSpinorField in(sub, QUDA_DOUBLE_PRECISION), out(sub, QUDA_DOUBLE_PRECISION);
  stage.uploadToCuda(&in, isEv);
  std::vector<double> ev(NeV);
  for (int i = 0; i < NeV; i++) ev[i] = (double)eigenValues[2 * i];
  TMQ_OK(tmq_deflate(out.handle(), in.handle(), set, ev.data(), NeV));
  vec_defl.downloadFromCuda(&out, isEv);
  vec_defl.unloadVector();        // keep h_elem (SoA) in step with the device copy, as packVector + loadVector leave it
}
template <typename Float> void QKXTM_Deflation<Float>::ApplyMdagM(Float *vec_out, Float *vec_in, QudaInvertParam *param) {
  if (!isFullOp) { ::ApplyMdagM((double *)vec_out, (double *)vec_in, param, isEv); return; }
  // full operator branch (Deflation.cpp:194-228)
  create_dirac(param);
  QKXTM_Vector<double> Kvec(BOTH, VECTOR);
  ColorSpinorField in(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION), out(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
  Kvec.packVector((double *)vec_in);
  Kvec.loadVector();
  Kvec.uploadToCuda(&in, false);
  TMQ_OK(tmq_poly_mdagm(out.handle(), in.handle(), -1, 0.0, 0.0));      // diracOp->MdagM on the full field (:213)
  Kvec.downloadFromCuda(&out, false);
  Kvec.unloadVector();
  Kvec.unpackVector();
  memcpy(vec_out, Kvec.H_elem(), (size_t)G.localVolume * 24 * sizeof(double));
}
// vec_defl = vec_in - U U^dag vec_in over the first NeV_defl eigenvectors (Deflation.cpp:1926-2060; full operator only)
template <typename Float> void QKXTM_Deflation<Float>::projectVector(QKXTM_Vector<Float> &vec_defl, QKXTM_Vector<Float> &vec_in, int is, int NeV_defl) {
  (void)is;
  if (!isFullOp) errorQuda("projectVector: This function only works with the Full Operator");
  if (NeV_defl == 0 || NeV == 0) {                                          // Deflation.cpp:2071-2076
    printfQuda("projectVector: Got NeV = %d. Will not project vector!\n", NeV_defl);
    vec_defl.packVector((Float *)vec_in.H_elem());
    vec_defl.loadVector();
    return;
  }
  QKXTM_Vector<Float> stage(BOTH, VECTOR);
  stage.packVector(vec_in.H_elem());
  stage.loadVector();
  ColorSpinorField in(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
  stage.uploadToCuda(&in, false);
  TMQ_OK(tmq_project(in.handle(), in.handle(), set, NeV_defl < NeV ? NeV_defl : NeV));
  vec_defl.downloadFromCuda(&in, false);
  vec_defl.unloadVector();
}
template <typename Float> void QKXTM_Deflation<Float>::projectVector(QKXTM_Vector<Float> &vec_defl, QKXTM_Vector<Float> &vec_in, int is) {
  projectVector(vec_defl, vec_in, is, NeV);                                 // Deflation.cpp:1931-2059: all NeV vectors
}
// Deflation.cpp:285-385: the reference reorders its HOST copy of the eigenvectors from [even | odd] to lexicographic sites, the order its
// host zgemv in projectVector needs.  Here the basis stays on the device in the native [even | odd] layout and projectVector converts the
// vector it is given instead, so there is nothing to reorder; the calls are kept so that calc_loops reads like the reference.
template <typename Float> void QKXTM_Deflation<Float>::MapEvenOddToFull() {
  if (!isFullOp) { printfQuda("WARNING: MapEvenOddToFull: This function only works with the Full Operator\n"); return; }
  if (NeV == 0) return;
  printfQuda("MapEvenOddToFull: Completed successfully\n");
}
template <typename Float> void QKXTM_Deflation<Float>::MapEvenOddToFull(int i) {
  if (!isFullOp) errorQuda("MapEvenOddToFull: This function only works with the Full Operator");
  if (NeV == 0) return;
  printfQuda("MapEvenOddToFull: Vector %d completed successfully\n", i);
}
template <typename Float> void QKXTM_Deflation<Float>::copyEigenVectorToQKXTM_Vector(int id, Float *vec) {
  if (NeV == 0) return;
  if (id < 0 || id >= NeV) errorQuda("eigenvector index %d out of range", id);
  QKXTM_Vector<Float> stage(BOTH, VECTOR);
  ColorSpinorField v(isFullOp ? QUDA_FULL_SITE_SUBSET : QUDA_PARITY_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
  TMQ_OK(tmq_copy(v.handle(), tmq_eigset_vector(set, id)));
  stage.downloadFromCuda(&v, isEv);
  stage.download();
  memcpy(vec, stage.H_elem(), (size_t)G.localVolume * 24 * sizeof(Float));
}
template class QKXTM_Deflation<double>;

template class QKXTM_Field<double>;
template class QKXTM_Field<float>;
template class QKXTM_Gauge<double>;
template class QKXTM_Gauge<float>;
template class QKXTM_Vector<double>;
template class QKXTM_Vector<float>;
template class QKXTM_Propagator<double>;
template class QKXTM_Propagator<float>;
template class QKXTM_Propagator3D<double>;
template class QKXTM_Propagator3D<float>;
template class QKXTM_Contraction<double>;
template class QKXTM_Contraction<float>;

}  // namespace quda

// ---- solve entry points ---------------------------------------------------------------------------------------------------------
void MG_bench(void **gaugeSmeared, void **gauge, QudaGaugeParam *gauge_param, QudaInvertParam *param, qudaQKXTMinfo info) {
  MG_bench(gaugeSmeared, gauge, gauge_param, param, info, (double *)NULL);
}
void MG_bench(void **gaugeSmeared, void **gauge, QudaGaugeParam *gauge_param, QudaInvertParam *param, qudaQKXTMinfo info,
              double *prop_out) {
  (void)gauge;
  if (!param || !gauge_param) errorQuda("null argument");
  SolverSubstitution subst(param, "MG_bench");
  check_solver(param);
  if (!G.qkxtm_initialized) errorQuda("You must initialize init_qudaQKXTM first");
  if (G.nsmearGauss != 0 && !gaugeSmeared) errorQuda("Gaussian smearing of the source needs the smeared links (gaugeSmeared)");
  const bool flag_eo = info.isEven;     // (the reference leaves this unset; b and x are full fields, so it is moot)
  const auto T0 = std::chrono::steady_clock::now();
  const long long V = G.localVolume;
  QKXTM_Vector<double> *K_vector = new QKXTM_Vector<double>(BOTH, VECTOR);
  QKXTM_Vector<double> *K_guess = new QKXTM_Vector<double>(BOTH, VECTOR);
  QKXTM_Gauge<double> *K_gaugeSmeared = NULL;
  if (gaugeSmeared) {
    K_gaugeSmeared = new QKXTM_Gauge<double>(BOTH, GAUGE);              // interface.cpp:72-77
    K_gaugeSmeared->packGauge(gaugeSmeared);
    K_gaugeSmeared->loadGauge();
    K_gaugeSmeared->calculatePlaq();
  }
  printfQuda("Memory allocation was successfull\n");
  ColorSpinorField *b = new ColorSpinorField(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
  ColorSpinorField *x = new ColorSpinorField(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
  printfQuda("\n ### Calculations for source-position %d - %02d.%02d.%02d.%02d begin now ###\n\nForward Inversions:\n", 0, 0, 0, 0, 0);
  const bool massnorm = param->mass_normalization == QUDA_MASS_NORMALIZATION || param->mass_normalization == QUDA_ASYMMETRIC_MASS_NORMALIZATION;
  for (int isc = 0; isc < 12; isc++) {
    const auto t4 = std::chrono::steady_clock::now();
    if (param->mu < 0) param->mu *= -1.0;                 // "Ensure mu is positive" (interface.cpp:162-163)
    // point source at the origin, spin-colour isc (on the rank that holds it).  The reference zeroes a host vector, sets one entry,
    // packVector()s it and copies V x 24 reals over PCIe (interface.cpp:165-176); the same device vector is made here by a device
    // memset and ONE 16-byte copy into component isc of site 0 of the QKXTM layout d[(s*3+c)*V + x]
    K_vector->zero_device();
    if (G.coord[0] == 0 && G.coord[1] == 0 && G.coord[2] == 0 && G.coord[3] == 0) {
      const double one[2] = {1.0, 0.0};
      TMQ_OK(tmq_h2d(G.ctx, K_vector->D_elem() + (size_t)isc * V * 2, one, sizeof(one)));
    }
    if (K_gaugeSmeared) {
      K_guess->gaussianSmearing(*K_vector, *K_gaugeSmeared);            // interface.cpp:184 (nsmearGauss = 0 copies)
      K_guess->uploadToCuda(b, flag_eo);                                // :185
    } else {
      K_vector->uploadToCuda(b, flag_eo);
    }
    printfQuda(" up - %02d: \n", isc);
    solve_device(*x, *b, param);
    K_vector->downloadFromCuda(x, flag_eo);
    if (massnorm) K_vector->scaleVector(2 * param->kappa);
    // column isc to the caller: converted to the host order on the device and copied on the download stream, behind the next solve
    if (prop_out)
      TMQ_OK(tmq_spinor_to_host_async(prop_out + (size_t)isc * V * 24, x->handle(), isc & 1, TMQ_HOST_ORDER_LEX, massnorm ? 2 * param->kappa : 1.0));
    printfQuda("Inversion up = %d, for source = %d finished in time %f sec\n", isc, 0,
               std::chrono::duration<double>(std::chrono::steady_clock::now() - t4).count());
  }
  if (prop_out) TMQ_OK(tmq_host_wait(G.ctx));
  delete K_vector;
  delete K_guess;
  if (K_gaugeSmeared) delete K_gaugeSmeared;
  delete x;
  delete b;
  printfQuda("...Done (%f sec)\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - T0).count());
}

void calcMG_threepTwop_EvenOdd(void **gaugeSmeared, void **gauge, QudaGaugeParam *gauge_param, QudaInvertParam *param, qudaQKXTMinfo info,
                               char *filename_twop, char *filename_threep, WHICHPARTICLE NUCLEON) {
  if (!param || !gauge_param || !filename_twop) errorQuda("null argument");
  SolverSubstitution subst(param, "calcMG_threepTwop_EvenOdd");
  check_solver(param);
  if (!G.qkxtm_initialized) errorQuda("You must initialize init_qudaQKXTM first");
  if (param->gamma_basis != QUDA_UKQCD_GAMMA_BASIS) errorQuda("This function works only with ukqcd gamma basis");       // interface.cpp:270-273
  if (info.CorrFileFormat != ASCII_FORM) errorQuda("only the ASCII two-point format is built (no HDF5 here)");
  if (info.CorrSpace != MOMENTUM_SPACE) errorQuda("the ASCII two-point writer supports only momentum space");           // Contraction.cpp:1565
  bool any3pt = false;
  for (int i = 0; i < info.Nsources; i++) any3pt = any3pt || info.run3pt_src[i];
  if (any3pt && !filename_threep) errorQuda("null three-point file name");
  if (any3pt && (info.Ntsink < 0 || info.Ntsink > MAX_TSINK)) errorQuda("bad number of sink-source separations %d", info.Ntsink);
  if (G.nsmearGauss != 0 && !gaugeSmeared) errorQuda("Gaussian smearing needs the smeared links (gaugeSmeared)");
  const bool flag_eo = info.isEven;
  const long long V = G.localVolume;
  const int nm = qkxtm_Nmoms();
  const auto T0 = std::chrono::steady_clock::now();
  double *input_vector = (double *)malloc((size_t)V * 24 * sizeof(double));
  if (!input_vector) errorQuda("Error allocating memory for the host source");
  QKXTM_Gauge<double> *K_gaugeSmeared = NULL;
  if (gaugeSmeared) {
    K_gaugeSmeared = new QKXTM_Gauge<double>(BOTH, GAUGE);                       // interface.cpp:344-348
    K_gaugeSmeared->packGauge(gaugeSmeared);
    K_gaugeSmeared->loadGauge();
    K_gaugeSmeared->calculatePlaq();
  }
  QKXTM_Vector<double> *K_vector = new QKXTM_Vector<double>(BOTH, VECTOR);
  QKXTM_Vector<double> *K_guess = new QKXTM_Vector<double>(BOTH, VECTOR);
  QKXTM_Vector<float> *K_temp = new QKXTM_Vector<float>(BOTH, VECTOR);
  QKXTM_Propagator<float> *K_prop_up = new QKXTM_Propagator<float>(BOTH, PROPAGATOR);
  QKXTM_Propagator<float> *K_prop_down = new QKXTM_Propagator<float>(BOTH, PROPAGATOR);
  QKXTM_Contraction<float> *K_contract = new QKXTM_Contraction<float>();
  QKXTM_Propagator<float> *K_seqProp = NULL;
  QKXTM_Propagator3D<float> *K_prop3D_up = NULL, *K_prop3D_down = NULL;
  float *corrThp_local = NULL;
  if (any3pt) {
    K_seqProp = new QKXTM_Propagator<float>(BOTH, PROPAGATOR);                     // interface.cpp:372-376
    K_prop3D_up = new QKXTM_Propagator3D<float>(BOTH, PROPAGATOR3D);
    K_prop3D_down = new QKXTM_Propagator3D<float>(BOTH, PROPAGATOR3D);
    corrThp_local = (float *)calloc((size_t)G.localL[3] * nm * 16 * 2, sizeof(float));
    if (!corrThp_local) errorQuda("Cannot allocate memory for the three-point function");
  }
  static const char *proj_names[5] = {"G4", "G5G123", "G5G1", "G5G2", "G5G3"};  // info.thrp_proj_type (:284-288)
  // the links of the conserved-current / one-derivative insertions (interface.cpp:351-357: K_gaugeContractions->packGauge(gauge))
  QKXTM_Gauge<float> *K_gaugeContractions = NULL;
  float *corrThp_noether = NULL, *corrThp_oneD = NULL;
  // the conserved-current and one-derivative insertions read the neighbours' propagators and links; their halo exchange is not
  // built, so on a split lattice the request is REFUSED (the reference computes them there: a silent skip would lose two of the
  // three output files).  Passing gauge = NULL asks for the ultra-local insertion only, on any process grid.
  if (any3pt && gauge && G.nranks > 1)
    errorQuda("calcMG_threepTwop_EvenOdd: the conserved-current / one-derivative insertions are not available on a split lattice "
              "(%d ranks); pass gauge = NULL for the ultra-local three-point function only", G.nranks);
  if (any3pt && gauge && G.nranks == 1) {
    K_gaugeContractions = new QKXTM_Gauge<float>(BOTH, GAUGE);
    K_gaugeContractions->packGauge(gauge);
    K_gaugeContractions->loadGauge();
    corrThp_noether = (float *)calloc((size_t)G.localL[3] * nm * 4 * 2, sizeof(float));
    corrThp_oneD = (float *)calloc((size_t)G.localL[3] * nm * 4 * 16 * 2, sizeof(float));
    if (!corrThp_noether || !corrThp_oneD) errorQuda("Cannot allocate memory for the three-point function");
  } else if (any3pt) {
    K_gaugeContractions = new QKXTM_Gauge<float>(NONE, GAUGE);
  }
  float *corrMesons = (float *)calloc((size_t)G.localL[3] * nm * 2 * 10 * 2, sizeof(float));
  float *corrBaryons = (float *)calloc((size_t)G.localL[3] * nm * 2 * 10 * 4 * 4 * 2, sizeof(float));
  if (!corrMesons || !corrBaryons) errorQuda("Cannot allocate memory for the two-point functions");
  printfQuda("Memory allocation was successfull\n");
  ColorSpinorField *b = new ColorSpinorField(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
  ColorSpinorField *x = new ColorSpinorField(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
  const double mu_abs = param->mu < 0 ? -param->mu : param->mu;

  for (int isource = 0; isource < info.Nsources; isource++) {
    const int *sp = info.sourcePosition[isource];
    printfQuda("\n ### Calculations for source-position %d - %02d.%02d.%02d.%02d begin now ###\n\n", isource, sp[0], sp[1], sp[2], sp[3]);
    char filename_mesons[1024], filename_baryons[1024];
    snprintf(filename_mesons, sizeof(filename_mesons), "%s.mesons.SS.%02d.%02d.%02d.%02d.dat", filename_twop, sp[0], sp[1], sp[2], sp[3]);   // :602-613
    snprintf(filename_baryons, sizeof(filename_baryons), "%s.baryons.SS.%02d.%02d.%02d.%02d.dat", filename_twop, sp[0], sp[1], sp[2], sp[3]);
    if (info.check_files) {
      FILE *f = fopen(filename_mesons, "r"), *fb = fopen(filename_baryons, "r");
      const bool both = f && fb;
      if (f) fclose(f);
      if (fb) fclose(fb);
      if (both) continue;                                                   // :633-638
    }
    printfQuda("Forward Inversions:\n");
    for (int isc = 0; isc < 12; isc++) {
      // point source at the source position if this rank holds it (interface.cpp:648-665), Gaussian-smeared
      memset(input_vector, 0, (size_t)V * 24 * sizeof(double));
      int my_src[4];
      bool mine = true;
      for (int i = 0; i < 4; i++) { my_src[i] = sp[i] - G.coord[i] * G.localL[i]; mine = mine && my_src[i] >= 0 && my_src[i] < G.localL[i]; }
      if (mine)
        input_vector[((((size_t)my_src[3] * G.localL[2] + my_src[2]) * G.localL[1] + my_src[1]) * G.localL[0] + my_src[0]) * 24 + isc * 2] = 1.0;
      K_vector->packVector(input_vector);
      K_vector->loadVector();
      if (K_gaugeSmeared) K_guess->gaussianSmearing(*K_vector, *K_gaugeSmeared);
      else { K_guess->packVector(input_vector); K_guess->loadVector(); }
      for (int flavour = 0; flavour < 2; flavour++) {
        param->mu = flavour == 0 ? mu_abs : -mu_abs;                      // "Ensure mu is +ve" / "-ve" (:667-668, :733-734)
        K_guess->uploadToCuda(b, flag_eo);
        printfQuda(" %s - %02d: \n", flavour == 0 ? "up" : "dn", isc);
        solve_device(*x, *b, param);
        K_vector->downloadFromCuda(x, flag_eo);
        if (param->mass_normalization == QUDA_MASS_NORMALIZATION || param->mass_normalization == QUDA_ASYMMETRIC_MASS_NORMALIZATION)
          K_vector->scaleVector(2 * param->kappa);
        K_temp->castDoubleToFloat(*K_vector);
        (flavour == 0 ? K_prop_up : K_prop_down)->absorbVectorToDevice(*K_temp, isc / 3, isc % 3);
      }
    }
    // ---- fixed-sink three-point function, ultra-local insertion (interface.cpp:764-1170); uses the forward propagators BEFORE
    //      their sink smearing and rotation, as the reference does -------------------------------------------------------------
    if (info.run3pt_src[isource]) {
      const int T = G.localL[3] * G.grid[3];
      for (int its = 0; its < info.Ntsink; its++) {
        const int my_fixSinkTime = (info.tsinkSource[its] + sp[3]) % T - G.coord[3] * G.localL[3];
        const bool mine_t = my_fixSinkTime >= 0 && my_fixSinkTime < G.localL[3];
        if (mine_t) { K_prop3D_up->absorbTimeSlice(*K_prop_up, my_fixSinkTime); K_prop3D_down->absorbTimeSlice(*K_prop_down, my_fixSinkTime); }
        // sink smearing of the 3-d propagators, column by column (:790-826)
        for (int nu = 0; nu < 4; nu++)
          for (int c2 = 0; c2 < 3; c2++)
            for (int flavour = 0; flavour < 2; flavour++) {
              QKXTM_Propagator3D<float> *P3 = flavour == 0 ? K_prop3D_up : K_prop3D_down;
              K_temp->zero_device();
              if (mine_t) K_temp->copyPropagator3D(*P3, my_fixSinkTime, nu, c2);
              K_vector->castFloatToDouble(*K_temp);
              if (K_gaugeSmeared) K_guess->gaussianSmearing(*K_vector, *K_gaugeSmeared);
              K_temp->castDoubleToFloat(K_gaugeSmeared ? *K_guess : *K_vector);
              if (mine_t) P3->absorbVectorTimeSlice(*K_temp, my_fixSinkTime, nu, c2);
            }
        for (int proj = 0; proj < info.Nproj[its]; proj++) {
          const WHICHPROJECTOR PID = (WHICHPROJECTOR)info.proj_list[its][proj];
          if ((int)PID < 0 || (int)PID > 4) errorQuda("bad projector %d", (int)PID);
          printfQuda("\n# Three-point function calculation for source-position = %d, sink-source = %d, projector %s begins now\n", isource,
                     info.tsinkSource[its], proj_names[(int)PID]);
          char filename_threep_base[1024];
          snprintf(filename_threep_base, sizeof(filename_threep_base), "%s_tsink%d_proj%s", filename_threep, info.tsinkSource[its], proj_names[(int)PID]);   // :838-840
          for (int part = 1; part <= 2; part++) {
            // part 1: the flavour that occurs twice (up for the proton) with the OTHER flavour's operator; part 2 the reverse (:853-866, 1007-1022)
            const bool up_line = (NUCLEON == PROTON) == (part == 1);
            param->mu = up_line ? -mu_abs : mu_abs;
            printfQuda("Sequential Inversions, flavor %s:\n", up_line ? "up" : "dn");
            for (int nu = 0; nu < 4; nu++)
              for (int c2 = 0; c2 < 3; c2++) {
                K_temp->zero_device();
                if (mine_t) {
                  if (part == 1) {
                    if (NUCLEON == PROTON) K_contract->seqSourceFixSinkPart1(*K_temp, *K_prop3D_up, *K_prop3D_down, my_fixSinkTime, nu, c2, PID, NUCLEON);
                    else K_contract->seqSourceFixSinkPart1(*K_temp, *K_prop3D_down, *K_prop3D_up, my_fixSinkTime, nu, c2, PID, NUCLEON);
                  } else {
                    K_contract->seqSourceFixSinkPart2(*K_temp, NUCLEON == PROTON ? *K_prop3D_up : *K_prop3D_down, my_fixSinkTime, nu, c2, PID, NUCLEON);
                  }
                }
                K_temp->conjugate();
                K_temp->apply_gamma5();
                K_vector->castFloatToDouble(*K_temp);
                K_vector->scaleVector(1e+10);                              // "Scale up vector to avoid MP errors" (:880)
                if (K_gaugeSmeared) K_guess->gaussianSmearing(*K_vector, *K_gaugeSmeared);
                (K_gaugeSmeared ? K_guess : K_vector)->uploadToCuda(b, flag_eo);
                printfQuda("%02d - \n", nu * 3 + c2);
                solve_device(*x, *b, param);
                K_vector->downloadFromCuda(x, flag_eo);
                if (param->mass_normalization == QUDA_MASS_NORMALIZATION || param->mass_normalization == QUDA_ASYMMETRIC_MASS_NORMALIZATION)
                  K_vector->scaleVector(2 * param->kappa);
                K_vector->scaleVector(1e-10);                              // "Rescale to normal"
                K_temp->castDoubleToFloat(*K_vector);
                K_seqProp->absorbVectorToDevice(*K_temp, nu, c2);
              }
            // part 1 contracts with the forward propagator of the doubly occurring flavour, part 2 with the other (:937-951, 1095-1109)
            QKXTM_Propagator<float> *fwd = up_line ? K_prop_up : K_prop_down;
            K_contract->contractFixSink(*K_seqProp, *fwd, *K_gaugeContractions, corrThp_local, corrThp_noether, corrThp_oneD, PID, NUCLEON, part, isource,
                                        info.CorrSpace);
            K_contract->writeThrp_ASCII(corrThp_local, corrThp_noether, corrThp_oneD, NUCLEON, part, filename_threep_base, isource, info.tsinkSource[its],
                                        info.CorrSpace);
          }
        }
      }
    }
    // smear the forward propagators at the sink (interface.cpp:1190-1215; the reference stages this through the host and
    // QUDA's performWuppertalnStep, here the same device kernel as at the source)
    if (K_gaugeSmeared)
      for (int nu = 0; nu < 4; nu++)
        for (int c2 = 0; c2 < 3; c2++)
          for (int flavour = 0; flavour < 2; flavour++) {
            QKXTM_Propagator<float> *P = flavour == 0 ? K_prop_up : K_prop_down;
            K_temp->copyPropagator(*P, nu, c2);
            K_vector->castFloatToDouble(*K_temp);
            K_guess->gaussianSmearing(*K_vector, *K_gaugeSmeared);
            K_temp->castDoubleToFloat(*K_guess);
            P->absorbVectorToDevice(*K_temp, nu, c2);
          }
    K_prop_up->rotateToPhysicalBase_device(+1);                            // :1217-1218
    K_prop_down->rotateToPhysicalBase_device(-1);
    const auto t1 = std::chrono::steady_clock::now();
    K_contract->contractBaryons(*K_prop_up, *K_prop_down, corrBaryons, isource, info.CorrSpace);  // :1220
    K_contract->contractMesons(*K_prop_up, *K_prop_down, corrMesons, isource, info.CorrSpace);    // :1222
    printfQuda("TIME_REPORT - Two-point Contractions: %f sec\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());
    printfQuda("The mesons two-point function ASCII filename is: %s\n", filename_mesons);
    printfQuda("The baryons two-point function ASCII filename is: %s\n", filename_baryons);
    K_contract->writeTwopBaryons_ASCII(corrBaryons, filename_baryons, isource, info.CorrSpace);   // :1241-1248
    K_contract->writeTwopMesons_ASCII(corrMesons, filename_mesons, isource, info.CorrSpace);
  }
  param->mu = mu_abs;
  free(corrMesons);
  free(corrBaryons);
  if (corrThp_local) free(corrThp_local);
  if (corrThp_noether) free(corrThp_noether);
  if (corrThp_oneD) free(corrThp_oneD);
  if (K_gaugeContractions) delete K_gaugeContractions;
  if (K_seqProp) delete K_seqProp;
  if (K_prop3D_up) delete K_prop3D_up;
  if (K_prop3D_down) delete K_prop3D_down;
  free(input_vector);
  delete K_contract; delete K_prop_down; delete K_prop_up; delete K_temp; delete K_guess; delete K_vector;
  if (K_gaugeSmeared) delete K_gaugeSmeared;
  delete x; delete b;
  printfQuda("...Done (%f sec)\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - T0).count());
}

void calcLowModeProjection(QudaInvertParam *evInvParam, qudaQKXTM_arpackInfo arpackInfo) {
  calcLowModeProjection(evInvParam, arpackInfo, (int *)NULL, (double *)NULL);
}
void calcLowModeProjection(QudaInvertParam *evInvParam, qudaQKXTM_arpackInfo arpackInfo, int *nconv, double *evals) {
  const char *fname = "calcLowModeProjection";
  if (!evInvParam) errorQuda("null argument");
  if (!G.quda_initialized) errorQuda("%s: QUDA not initialized", fname);
  // checks for the exact deflation part (interface.cpp:1360-1368)
  if ((evInvParam->matpc_type != QUDA_MATPC_EVEN_EVEN_ASYMMETRIC) && (evInvParam->matpc_type != QUDA_MATPC_ODD_ODD_ASYMMETRIC))
    errorQuda("Only asymmetric operators are supported in deflation");
  if (arpackInfo.isEven && (evInvParam->matpc_type != QUDA_MATPC_EVEN_EVEN_ASYMMETRIC)) errorQuda("%s: Inconsistency between operator types!", fname);
  if ((!arpackInfo.isEven) && (evInvParam->matpc_type != QUDA_MATPC_ODD_ODD_ASYMMETRIC)) errorQuda("%s: Inconsistency between operator types!", fname);
  QKXTM_Deflation<double> *deflation = new QKXTM_Deflation<double>(evInvParam, arpackInfo);
  deflation->printInfo();
  const auto t1 = std::chrono::steady_clock::now();
  deflation->eigenSolver();
  printfQuda("%s TIME REPORT:Full Operator EigenVector Calculation: %f sec\n", fname,
             std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());
  if (nconv) *nconv = deflation->Converged();
  if (evals) for (int i = 0; i < deflation->NeVs(); i++) evals[i] = deflation->EigenValues()[2 * i];
  printfQuda("\nCleaning up...\n");
  delete deflation;
  printfQuda("...Done\n");
}

void readLimeGauge(void **gauge, char *fname, QudaGaugeParam *param, QudaInvertParam *inv_param, int gridSize[4]) {
  if (!gauge || !fname || !param || !gridSize) errorQuda("null argument");
  if (param->cpu_prec != QUDA_DOUBLE_PRECISION) errorQuda("Dont support reading confs lime single precision");
  int GX[4], prec = 0;
  double kap = 0, mu = 0;
  if (tmq_lime_gauge_info(fname, GX, &prec, &kap, &mu)) errorQuda("%s", tmq_lime_last_error());
  if (inv_param) {
    printfQuda("Kappa given is : %.8f \t Kappa conf is : %.8f \t check that they agree\n", inv_param->kappa, kap);
    printfQuda("Mu given is : %f \t Mu conf is : %f \t may disagree for heavy quark\n", inv_param->mu, mu);
  }
  printfQuda("Precision:\t%i bit\n", prec);
  for (int d = 0; d < 4; d++) {
    if (gridSize[d] < 1 || GX[d] % gridSize[d]) errorQuda("lattice extent %d is not divisible by the process grid in dimension %d", GX[d], d);
    param->X[d] = GX[d] / gridSize[d];                                                    // QKXTM_read_conf.h:190-207
  }
  printfQuda("Volume:   \t%ix%ix%ix%i\nSubvolume:\t%ix%ix%ix%i\n", GX[0], GX[1], GX[2], GX[3], param->X[0], param->X[1], param->X[2], param->X[3]);
  if (tmq_lime_read_gauge(fname, (double *const *)gauge, param->X, gridSize, G.coord)) errorQuda("%s", tmq_lime_last_error());
  comm_barrier();      // ranks read their blocks at different speeds (the reference's MPI_File_read_all is collective)
}
void readLimeGaugeSmeared(void **gauge, char *fname, QudaGaugeParam *param, QudaInvertParam *inv_param, int gridSize[4]) {
  readLimeGauge(gauge, fname, param, inv_param, gridSize);      // same record layout (QKXTM_read_conf.h:401-675)
}
void applyBoundaryCondition(void **gauge, int Vh, QudaGaugeParam *gauge_param) {
  if (gauge_param->cpu_prec != QUDA_DOUBLE_PRECISION) errorQuda("boundary condition application implement only for double precision");
  if ((long long)Vh * 2 != (long long)gauge_param->X[0] * gauge_param->X[1] * gauge_param->X[2] * gauge_param->X[3]) errorQuda("Vh does not match the lattice");
  tmq_apply_t_boundary((double *const *)gauge, gauge_param->X, G.grid, G.coord, (int)gauge_param->t_boundary);
}

void calc_loops_solve(double *h_solution, double *h_source, QudaInvertParam *param, qudaQKXTMinfo info) {
  if (!h_solution || !h_source || !param) errorQuda("null argument");
  check_solver(param);
  if (!G.qkxtm_initialized) errorQuda("You must initialize init_qudaQKXTM first");
  const bool flag_eo = info.isEven;
  QKXTM_Vector<double> *K_vector = new QKXTM_Vector<double>(BOTH, VECTOR);
  ColorSpinorField b(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION), x(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
  K_vector->packVector(h_source);                                      // interface.cpp:2008
  K_vector->loadVector();                                              // :2009
  K_vector->uploadToCuda(&b, flag_eo);                                 // :2010
  solve_device(x, b, param);                                           // :2020-2041
  K_vector->downloadFromCuda(&x, flag_eo);                             // :2062
  K_vector->download();                                                // :2063
  memcpy(h_solution, K_vector->H_elem(), (size_t)G.localVolume * 24 * sizeof(double));
  delete K_vector;
}

void ApplyMdagM(double *h_out, double *h_in, QudaInvertParam *param, bool isEven) {
  if (!h_out || !h_in || !param) errorQuda("null argument");
  check_param(param);
  if (!G.qkxtm_initialized) errorQuda("You must initialize init_qudaQKXTM first");
  const int want = isEven ? 0 : 1;
  if ((matpc_of(param) & 1) != want) errorQuda("matpc_type does not match the requested parity");
  create_dirac(param);
  QKXTM_Vector<double> *Kvec = new QKXTM_Vector<double>(BOTH, VECTOR);
  ColorSpinorField in(QUDA_PARITY_SITE_SUBSET, QUDA_DOUBLE_PRECISION), out(QUDA_PARITY_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
  Kvec->packVector(h_in);
  Kvec->loadVector();
  Kvec->uploadToCuda(&in, isEven);
  TMQ_OK(tmq_mdagm(out.handle(), in.handle()));                        // diracOp->MdagM (Deflation.cpp:265)
  Kvec->downloadFromCuda(&out, isEven);
  Kvec->unloadVector();
  Kvec->unpackVector();
  memcpy(h_out, Kvec->H_elem(), (size_t)G.localVolume * 24 * sizeof(double));
  delete Kvec;
}

// ---- calc_loops (include/qudaQKXTM.h:501-507, lib/qudaQKXTM_interface.cpp:1409-2233) -----------------------------------------------
namespace {
quda::qkxtm_loop_hook g_loop_hook = nullptr;
void *g_loop_hook_user = nullptr;
}
namespace quda {
void qkxtm_set_loop_hook(qkxtm_loop_hook hook, void *user) { g_loop_hook = hook; g_loop_hook_user = user; }
}

// QKXTM_LOOP_DUMP=<file>: a built-in hook for drivers that install none (the reference's own Calc_Loops.cpp): appends, for every event,
// 8 doubles (kind, is, ih, sc, dstep, NeV_defl, eigenvalue | iterations, true_res), then for kind 1 the host source, then the vector
// handed to the hook (downloaded, plug-in AoS order) -- what a contraction would consume, in a form a test can check
static void dump_hook(const qkxtm_loop_event *ev, void *user) {
  FILE *f = (FILE *)user;
  QKXTM_Vector<double> K(BOTH, VECTOR);
  K.downloadFromCuda(ev->x, false);
  K.download();
  const double head[8] = {(double)ev->kind, (double)ev->is, (double)ev->ih, (double)ev->sc, (double)ev->dstep, (double)ev->NeV_defl,
                          ev->kind == 0 ? ev->eigenvalue : (double)ev->iter, ev->true_res};
  const size_t n = (size_t)G.localVolume * 24;
  bool ok = fwrite(head, sizeof(double), 8, f) == 8;
  if (ev->kind == 1) ok = ok && fwrite(ev->h_source, sizeof(double), n, f) == n;
  ok = ok && fwrite(K.H_elem(), sizeof(double), n, f) == n;
  if (!ok) errorQuda("QKXTM_LOOP_DUMP: short write");
}

void calc_loops(void **gaugeToPlaquette, QudaInvertParam *EvInvParam, QudaInvertParam *param, QudaGaugeParam *gauge_param,
                qudaQKXTM_arpackInfo arpackInfo, qudaQKXTM_loopInfo loopInfo, qudaQKXTMinfo info) {
  const char *fname = "calc_loops";
  if (!EvInvParam || !param || !gauge_param) errorQuda("%s: null argument", fname);
  FILE *dump_file = NULL;
  const qkxtm_loop_hook hook_saved = g_loop_hook;
  void *const hook_user_saved = g_loop_hook_user;
  if (!g_loop_hook && getenv("QKXTM_LOOP_DUMP") && *getenv("QKXTM_LOOP_DUMP")) {
    char path[1024];
    if (G.nranks > 1) snprintf(path, sizeof(path), "%s.rank%d", getenv("QKXTM_LOOP_DUMP"), G.rank);
    else snprintf(path, sizeof(path), "%s", getenv("QKXTM_LOOP_DUMP"));
    dump_file = fopen(path, "wb");
    if (!dump_file) errorQuda("QKXTM_LOOP_DUMP: cannot open %s", path);
    g_loop_hook = dump_hook; g_loop_hook_user = dump_file;
  }
  // ---- parameter checks (:1427-1494) ----
  unsigned short int *Vc = NULL;
  const int k_probing = loopInfo.k_probing;
  const bool spinColorDil = loopInfo.spinColorDil;
  bool isProbing = false, isProbingMstep = false;
  int Nc = 1, Nc_low = 0, Nc_high = 1;
  if (!G.quda_initialized) errorQuda("%s: QUDA not initialized", fname);
  if (!G.qkxtm_initialized) errorQuda("You must initialize init_qudaQKXTM first");
  if (k_probing > 0) {
    Nc = 2 * (int)std::lround(std::pow(2.0, 4 * (k_probing - 1)));
    Vc = hch_coloring(k_probing, 4);                       // 4D hierarchical coloring
    isProbing = true;
    if (loopInfo.hadamLow < 0 || loopInfo.hadamHigh < 0) errorQuda("Error: You cannot give negative values for hadamLow or hadamHigh");
    if (loopInfo.hadamLow > loopInfo.hadamHigh) errorQuda("Error: hadamLow cannot be greater than hadamHigh");
    Nc_low = loopInfo.hadamLow;
    Nc_high = loopInfo.hadamHigh == 0 ? Nc : loopInfo.hadamHigh;
    if (Nc_high > Nc) errorQuda("Error: You cannot choose hadamHigh to be greater than Nc");
    if (Nc_low > 0 || Nc_high < Nc) isProbingMstep = true;
  }
  const int Nsc = spinColorDil ? 12 : 1;
  const QudaVerbosity verbosity_saved = G.verbosity;
  G.verbosity = param->verbosity;                          // pushVerbosity(param->verbosity)
  printfQuda("\n### %s: Loop calculation begins now\n\n", fname);
  if ((EvInvParam->matpc_type != QUDA_MATPC_EVEN_EVEN_ASYMMETRIC) && (EvInvParam->matpc_type != QUDA_MATPC_ODD_ODD_ASYMMETRIC))
    errorQuda("Only asymmetric operators are supported in deflation");
  if (arpackInfo.isEven && (EvInvParam->matpc_type != QUDA_MATPC_EVEN_EVEN_ASYMMETRIC)) errorQuda("%s: Inconsistency between operator types!", fname);
  if ((!arpackInfo.isEven) && (EvInvParam->matpc_type != QUDA_MATPC_ODD_ODD_ASYMMETRIC)) errorQuda("%s: Inconsistency between operator types!", fname);
  if ((param->inv_type != QUDA_GCR_INVERTER) && (param->inv_type != QUDA_CG_INVERTER)) errorQuda("%s: This function works only with GCR/CG solver", fname);
  SolverSubstitution subst(param, fname);          // the GCR branch (:2025-2030) solves the same system
  if (param->gamma_basis != QUDA_UKQCD_GAMMA_BASIS) errorQuda("%s: This function works only with ukqcd gamma basis", fname);
  if (param->dirac_order != QUDA_DIRAC_ORDER) errorQuda("%s: This function works only with color-inside-spin", fname);
  if (loopInfo.FileFormat == HDF5_FORM && g_loop_hook == nullptr) printfQuda("%s: no contraction hook installed: no loop files (HDF5 or ASCII) are written\n", fname);
  check_solver(param);

  const int Nstoch = loopInfo.Nstoch;
  const unsigned long int seed = loopInfo.seed;
  const int Ndump = loopInfo.Ndump;
  loopInfo.Nmoms = qkxtm_Nmoms();
  const int deflSteps = loopInfo.nSteps_defl;
  if (deflSteps < 0 || deflSteps > MAX_DEFLSTEPS) errorQuda("%s: bad number of deflation steps %d", fname, deflSteps);
  if (Nstoch < 0 || Ndump <= 0) errorQuda("%s: bad Nstoch / Ndump", fname);
  // names the reference fills into its by-value copy (:1517-1534); they feed its writers only
  static char lt0[] = "Scalar", lt1[] = "dOp", lt2[] = "Loops", lt3[] = "LoopsCv", lt4[] = "LpsDw", lt5[] = "LpsDwCv";
  char *lts[6] = {lt0, lt1, lt2, lt3, lt4, lt5};
  for (int i = 0; i < 6; i++) { loopInfo.loop_type[i] = lts[i]; loopInfo.loop_oneD[i] = i >= 2; }

  printfQuda("\nLoop Calculation Info\n=====================\n");
  printfQuda(" The seed is: %ld\n", seed);
  printfQuda(" The conf trajectory is: %04d\n", loopInfo.traj);
  printfQuda(" Will produce the loop for %d Momentum Combinations\n", loopInfo.Nmoms);
  printfQuda(" The loop base name is %s\n", loopInfo.loop_fname);
  if (isProbingMstep)
    printfQuda(" %d Stoch vectors, %d Hadamard vectors (using Mstep), %d spin-colour diluted : %04d inversions\n", Nstoch, Nc_high - Nc_low, Nsc,
               Nstoch * (Nc_high - Nc_low) * Nsc);
  else
    printfQuda(" %d Stoch vectors, %d Hadamard vectors, %d spin-colour diluted : %04d inversions\n", Nstoch, Nc, Nsc, Nstoch * Nc * Nsc);
  printfQuda(" Will project\n");
  for (int a = 0; a < deflSteps; a++) printfQuda(" Ndefl %d: %d\n", a, loopInfo.deflStep[a]);
  printfQuda(" exact eigenmodes fom the solutions\n");
  if (info.source_type == RANDOM) printfQuda(" Will use RANDOM stochastic sources\n");
  else if (info.source_type == UNITY) printfQuda(" Will use UNITY stochastic sources\n");
  printfQuda("=====================\n\n");

  // ---- exact part (:1713-1838): the low modes of the operator of EvInvParam; the loop contraction of each eigenpair is the hook ----
  printfQuda("\n ### Exact part calculation ###\n");
  const int NeV_Full = arpackInfo.nEv;
  QKXTM_Deflation<double> *deflation = new QKXTM_Deflation<double>(EvInvParam, arpackInfo);
  deflation->printInfo();
  auto t1 = std::chrono::steady_clock::now();
  deflation->eigenSolver();
  printfQuda("%s TIME REPORT:Full Operator EigenVector Calculation: %f sec\n", fname,
             std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());
  deflation->MapEvenOddToFull();
  if (g_loop_hook && NeV_Full > 0 && arpackInfo.isFullOp) {
    ColorSpinorField ev(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
    for (int n = 0; n < NeV_Full; n++) {
      TMQ_OK(tmq_copy(ev.handle(), tmq_eigset_vector(deflation->EigenSet(), n)));
      qkxtm_loop_event e;
      memset(&e, 0, sizeof(e));
      e.kind = 0; e.is = n; e.x = &ev; e.eigenvalue = (double)deflation->EigenValues()[2 * n];
      g_loop_hook(&e, g_loop_hook_user);                   // Loop_w_One_Der_FullOp_Exact(n, ...) (:1775)
    }
  }
  printfQuda("\n ### Exact part calculation Done ###\n");

  // ---- stochastic part (:1840-2134) ----
  printfQuda("\n ### Stochastic part calculation ###\n\n");
  if (gaugeToPlaquette) {
    QKXTM_Gauge<double> *K_gauge = new QKXTM_Gauge<double>(BOTH, GAUGE);     // :1846-1850
    K_gauge->packGauge(gaugeToPlaquette);
    K_gauge->loadGauge();
    K_gauge->calculatePlaq();
    delete K_gauge;
  }
  bool flag_eo = false;
  if (info.isEven) { printfQuda("%s: Solving for the Even-Even operator\n", fname); flag_eo = true; }
  else printfQuda("%s: Solving for the Odd-Odd operator\n", fname);

  const long long V = G.localVolume;
  double *input_vector = (double *)calloc((size_t)V * 24, sizeof(double));
  double *temp_input_vector = (isProbing || spinColorDil) ? (double *)calloc((size_t)V * 24, sizeof(double)) : NULL;
  if (!input_vector || ((isProbing || spinColorDil) && !temp_input_vector)) errorQuda("%s: Error allocating memory for the host sources", fname);
  ColorSpinorField *b = new ColorSpinorField(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
  ColorSpinorField *x = new ColorSpinorField(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
  ColorSpinorField *sol = new ColorSpinorField(QUDA_FULL_SITE_SUBSET, QUDA_DOUBLE_PRECISION);
  QKXTM_Vector<double> *K_vector = new QKXTM_Vector<double>(BOTH, VECTOR);

  void *rNum = qkxtm_rng_alloc(seed + (unsigned long int)comm_rank() * seed);   // gsl_rng_set(rNum, seed + comm_rank()*seed) (:1951)
  if (!rNum) errorQuda("%s: cannot allocate the random number generator", fname);
  const char *msg_str = "LOOPS";
  int iPrint = -1;
  for (int is = 0; is < Nstoch; is++) {
    t1 = std::chrono::steady_clock::now();
    getStochasticRandomSource<double>(input_vector, rNum, info.source_type);   // :1981-1982
    printfQuda("TIME_REPORT: %s %04d - Source creation: %f sec\n", msg_str, is + 1,
               std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());
    for (int ih = Nc_low; ih < Nc_high; ih++) {
      for (int sc = 0; sc < Nsc; sc++) {
        const auto t3 = std::chrono::steady_clock::now();
        double *src = input_vector;
        if (spinColorDil) {                                                  // :1994-2005
          if (isProbing) get_probing4D_spinColor_dilution<double>(temp_input_vector, input_vector, Vc, ih, sc);
          else get_spinColor_dilution<double>(temp_input_vector, input_vector, sc);
          src = temp_input_vector;
        } else if (isProbing) {
          get_probing4D_dilution<double>(temp_input_vector, input_vector, Vc, ih);
          src = temp_input_vector;
        }
        K_vector->packVector(src);                                           // :2008
        K_vector->loadVector();                                              // :2009
        K_vector->uploadToCuda(b, flag_eo);                                  // :2010
        const double orig_tol = param->tol;
        const int orig_maxiter = param->maxiter;
        solve_device(*sol, *b, param);                                       // :2020-2041 (prepare, M^dag, CG, reconstruct)
        printfQuda("TIME_REPORT: %s Stoch = %02d, HadVec = %02d, Spin-colour = %02d - Full Inversion Time: %f sec\n", msg_str, is, ih, sc, param->secs);
        param->tol = orig_tol;
        param->maxiter = orig_maxiter;
        for (int dstep = 0; dstep < deflSteps; dstep++) {                    // :2056-2112
          const int NeV_defl = loopInfo.deflStep[dstep];
          t1 = std::chrono::steady_clock::now();
          // x <- (1 - U U^dag) sol.  The reference stages this through the host (downloadFromCuda -> download -> projectVector ->
          // uploadToCuda, :2059-2066); the basis lives on the device here, so the projection never leaves it.
          if (NeV_defl > 0 && NeV_Full > 0) {
            if (!arpackInfo.isFullOp) errorQuda("projectVector: This function only works with the Full Operator");
            TMQ_OK(tmq_project(x->handle(), sol->handle(), deflation->EigenSet(), NeV_defl < NeV_Full ? NeV_defl : NeV_Full));
          } else {
            printfQuda("projectVector: Got NeV = %d. Will not project vector!\n", NeV_defl);
            TMQ_OK(tmq_copy(x->handle(), sol->handle()));
          }
          TMQ_OK(tmq_sync(G.ctx));
          printfQuda("TIME_REPORT: %s Stoch = %02d, HadVec = %02d, Spin-colour = %02d, NeV = %04d, Solution projection: %f sec\n", msg_str, is, ih, sc,
                     NeV_defl, std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());
          if (g_loop_hook) {                                                 // oneEndTrick_w_One_Der (:2074-2078): not built, hook only
            qkxtm_loop_event e;
            memset(&e, 0, sizeof(e));
            e.kind = 1; e.is = is; e.ih = ih; e.sc = sc; e.dstep = dstep; e.NeV_defl = NeV_defl; e.x = x; e.h_source = src;
            e.iter = param->iter; e.true_res = param->true_res;
            g_loop_hook(&e, g_loop_hook_user);
          }
          if (((is + 1) % Ndump == 0) && (ih * Nsc + sc == Nc_high * Nsc - 1) && dstep == 0) iPrint++;   // :2087-2089 (the FT + copy follow there)
        }
        printfQuda("TIME_REPORT: %s Stoch = %02d, HadVec = %02d, Spin-colour = %02d - Total Processing Time %f sec\n", msg_str, is, ih, sc,
                   std::chrono::duration<double>(std::chrono::steady_clock::now() - t3).count());
      }
    }
  }
  (void)iPrint;
  qkxtm_rng_free(rNum);
  printfQuda("\n ### Stochastic part calculation Done ###\n");
  printfQuda("\nCleaning up...\n");
  free(input_vector);
  if (temp_input_vector) free(temp_input_vector);
  if (Vc) free(Vc);
  delete deflation;
  delete K_vector;
  delete sol;
  delete x;
  delete b;
  if (dump_file) { fclose(dump_file); g_loop_hook = hook_saved; g_loop_hook_user = hook_user_saved; }
  printfQuda("...Done\n");
  G.verbosity = verbosity_saved;                           // popVerbosity()
}
