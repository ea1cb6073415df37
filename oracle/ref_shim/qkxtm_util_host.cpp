// qkxtm_util_host.cpp -- TEST INFRASTRUCTURE: compiles the REFERENCE'S OWN host utility file qkxtm/QKXTM_util.cpp (geometry,
// boundary conditions, recon-12, random SU(3), and -- through include/QKXTM_read_conf.h, which it includes -- the ILDG / LIME
// configuration reader) FROM WHERE IT LIES under /root/reference, and exposes the helpers of SURVEY.md 8a row a17 / 8f row 4
// to the parity tests.  The upstream-QUDA headers that file includes are replaced by the declaration-only stand-ins in
// ref_shim/stubs/ (written for this repository, not reference code); the externals it references are defined below:
// c-lime's reader on plain stdio (the reference calls limeCreateReader ... limeReaderReadData, include/QKXTM_read_conf.h:
// 148-214), a single-rank comm layer, and aborting stand-ins for everything the exposed helpers never reach.
// Built by oracle/Makefile into oracle/_ref/libqkxtm_util_ref.so.
#include <mpi.h>
#include <QKXTM_util.cpp>      // /root/reference/qkxtm/QKXTM_util.cpp

// ---- single-rank comm layer / never-reached externals -----------------------------------------------------------------
Topology *default_topo = nullptr;
const int *comm_coords(const Topology *) { static const int z[4] = {0, 0, 0, 0}; return z; }
#define NEVER(name) do { fprintf(stderr, "oracle/_ref: %s is a stand-in and must not be reached\n", name); abort(); } while (0)
void initCommsGridQuda(int, const int *, QudaCommsMap, void *) {}
int MPI_Init(int *, char ***) { return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Comm_rank(MPI_Comm, int *rank) { *rank = 0; return 0; }
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }      // one rank: nothing to broadcast
int MPI_Type_create_subarray(int, const int *, const int *, const int *, int, MPI_Datatype, MPI_Datatype *) { NEVER("MPI_Type_create_subarray"); }
int MPI_Type_commit(MPI_Datatype *) { NEVER("MPI_Type_commit"); }
int MPI_File_open(MPI_Comm, const char *, int, MPI_Info, MPI_File *) { NEVER("MPI_File_open"); }
int MPI_File_set_view(MPI_File, MPI_Offset, MPI_Datatype, MPI_Datatype, const char *, MPI_Info) { NEVER("MPI_File_set_view"); }
int MPI_File_read_all(MPI_File, void *, int, MPI_Datatype, MPI_Status *) { NEVER("MPI_File_read_all"); }
int MPI_File_close(MPI_File *) { NEVER("MPI_File_close"); }
QudaPrecision get_prec(char *) { NEVER("get_prec"); }
QudaReconstructType get_recon(char *) { NEVER("get_recon"); }
QudaInverterType get_solver_type(char *) { NEVER("get_solver_type"); }
QudaDslashType get_dslash_type(char *) { NEVER("get_dslash_type"); }
QudaMassNormalization get_mass_normalization_type(char *) { NEVER("get_mass_normalization_type"); }
QudaMatPCType get_matpc_type(char *) { NEVER("get_matpc_type"); }
QudaSolveType get_solve_type(char *) { NEVER("get_solve_type"); }
QudaTwistFlavorType get_flavor_type(char *) { NEVER("get_flavor_type"); }
QudaVerbosity get_verbosity_type(char *) { NEVER("get_verbosity_type"); }
QudaSchwarzType get_schwarz_type(char *) { NEVER("get_schwarz_type"); }
const char *get_quda_ver_str() { return "stand-in"; }

// ---- c-lime reader on stdio: 144-byte big-endian record headers, payload padded to 8 bytes --------------------------------
struct LimeReader_s { FILE *fp; long next; long data; unsigned long long bytes; unsigned long long read; char type[129]; };
LimeReader *limeCreateReader(FILE *fp) { LimeReader *r = (LimeReader *)calloc(1, sizeof(LimeReader)); r->fp = fp; r->next = 0; return r; }
void limeDestroyReader(LimeReader *r) { free(r); }
int limeReaderNextRecord(LimeReader *r) {
  unsigned char h[144];
  if (fseek(r->fp, r->next, SEEK_SET) != 0 || fread(h, 1, 144, r->fp) != 144) return LIME_EOF;
  if (!(h[0] == 0x45 && h[1] == 0x67 && h[2] == 0x89 && h[3] == 0xab)) return LIME_EOF;
  unsigned long long n = 0;
  for (int i = 0; i < 8; i++) n = (n << 8) | h[8 + i];
  memcpy(r->type, h + 16, 128); r->type[128] = 0;
  r->bytes = n; r->read = 0; r->data = r->next + 144; r->next = r->data + (long)((n + 7) & ~7ULL);
  return 0;
}
char *limeReaderType(LimeReader *r) { return r->type; }
n_uint64_t limeReaderBytes(LimeReader *r) { return r->bytes; }
int limeReaderReadData(void *dest, n_uint64_t *nbytes, LimeReader *r) {
  if (fseek(r->fp, r->data + (long)r->read, SEEK_SET) != 0) return -1;
  const size_t got = fread(dest, 1, (size_t)*nbytes, r->fp);
  r->read += got; *nbytes = got;
  return 0;
}

// ---- what the tests call ----------------------------------------------------------------------------------------------
static QudaGaugeParam make_param(const int X[4], int t_boundary) {
  QudaGaugeParam p;
  memset(&p, 0, sizeof(p));
  for (int d = 0; d < 4; d++) p.X[d] = X[d];
  p.anisotropy = 1.0; p.type = QUDA_WILSON_LINKS; p.gauge_order = QUDA_QDP_GAUGE_ORDER;
  p.t_boundary = t_boundary == -1 ? QUDA_ANTI_PERIODIC_T : QUDA_PERIODIC_T;
  p.cpu_prec = p.cuda_prec = QUDA_DOUBLE_PRECISION; p.reconstruct = QUDA_RECONSTRUCT_12; p.gauge_fix = QUDA_GAUGE_FIXED_NO;
  return p;
}
extern "C" {
void qutil_set_dims(const int X[4]) { int x[4] = {X[0], X[1], X[2], X[3]}; setDims(x); }                    // :94-128
int qutil_full_lattice_index(int i, int oddBit) { return fullLatticeIndex(i, oddBit); }                      // :418-442
int qutil_neighbor_index(int i, int oddBit, int dx4, int dx3, int dx2, int dx1) { return neighborIndex(i, oddBit, dx4, dx3, dx2, dx1); }   // :455-470
int qutil_get_odd_bit(int Y) { return getOddBit(Y); }                                                         // :191-197
void qutil_su3_reconstruct12(double *mat18, int dir, int ga_idx, int t_boundary) {                            // :281-295
  int X[4] = {Z[0], Z[1], Z[2], Z[3]};
  QudaGaugeParam p = make_param(X, t_boundary);
  su3Reconstruct12<double>(mat18, dir, ga_idx, &p);
}
void qutil_apply_gauge_field_scaling(double **gauge, int t_boundary) {                                       // :682-725
  int X[4] = {Z[0], Z[1], Z[2], Z[3]};
  QudaGaugeParam p = make_param(X, t_boundary);
  applyGaugeFieldScaling<double>(gauge, Vh, &p);
}
void qutil_construct_gauge_field(double **gauge, int type, unsigned int seed, int t_boundary) {              // :879-955, :1017-1030
  int X[4] = {Z[0], Z[1], Z[2], Z[3]};
  QudaGaugeParam p = make_param(X, t_boundary);
  srand(seed);
  construct_gauge_field((void **)gauge, type, QUDA_DOUBLE_PRECISION, &p);
}
// readLimeGauge (include/QKXTM_read_conf.h:816-822 -> read_custom_binary_gauge_field :107-400), single rank
void qutil_read_lime_gauge(double **gauge, const char *fname, int X_out[4], double kappa, double mu) {
  QudaGaugeParam p; memset(&p, 0, sizeof(p)); p.cpu_prec = QUDA_DOUBLE_PRECISION;
  for (int d = 0; d < 4; d++) p.X[d] = X_out[d];
  QudaInvertParam ip; memset(&ip, 0, sizeof(ip)); ip.kappa = kappa; ip.mu = mu;
  int grid[4] = {1, 1, 1, 1};
  readLimeGauge((void **)gauge, (char *)fname, &p, &ip, grid);
  for (int d = 0; d < 4; d++) X_out[d] = p.X[d];
}
}
