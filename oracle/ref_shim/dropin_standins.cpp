// dropin_standins.cpp -- TEST INFRASTRUCTURE for the drop-in link test (oracle/Makefile, target _ref/dropin/*): the reference's OWN
// drivers (qkxtm/MG_Bench.cpp, Calc_Loops.cpp, CalcMG_2pt3pt_EvenOdd.cpp, CalcLowModeProjection.cpp) and their driver-side sources
// (qkxtm/QKXTM_util.cpp, qkxtm/misc.cpp) are compiled UNMODIFIED from /root/reference against include/compat/ and linked against
// libqkxtm_tmq.so + libtmq.so.  The two third-party libraries those sources need besides QUDA -- MPI and c-lime -- are absent in this
// image; this file supplies one-process stand-ins for the handful of calls they make (a production build links the real libraries).
#include <mpi.h>      // oracle/ref_shim/stubs/mpi.h
extern "C" {
#include <lime.h>     // oracle/ref_shim/stubs/lime.h (the reference includes it inside extern "C", include/QKXTM_read_conf.h)
}
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

// ---- MPI, one process ----------------------------------------------------------------------------------------------------------
int MPI_Init(int *, char ***) { return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Comm_rank(MPI_Comm, int *rank) { *rank = 0; return 0; }
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
// MPI-IO of include/QKXTM_read_conf.h:684-760 (the reader of the binary payload): one process reads the whole block
namespace {
struct File { FILE *fp; long long off; };
}
int MPI_Type_create_subarray(int, const int *, const int *, const int *, int, MPI_Datatype, MPI_Datatype *newt) { *newt = MPI_DOUBLE; return 0; }
int MPI_Type_commit(MPI_Datatype *) { return 0; }
int MPI_File_open(MPI_Comm, const char *fname, int, MPI_Info, MPI_File *f) {
  File *h = new File{fopen(fname, "rb"), 0};
  *f = h;
  return h->fp ? 0 : 1;
}
int MPI_File_set_view(MPI_File f, MPI_Offset off, MPI_Datatype, MPI_Datatype, const char *, MPI_Info) { ((File *)f)->off = off; return 0; }
int MPI_File_read_all(MPI_File f, void *buf, int count, MPI_Datatype, MPI_Status *) {
  File *h = (File *)f;
  if (fseek(h->fp, (long)h->off, SEEK_SET) != 0) return 1;
  return fread(buf, sizeof(double), (size_t)count, h->fp) == (size_t)count ? 0 : 1;
}
int MPI_File_close(MPI_File *f) { File *h = (File *)*f; if (h->fp) fclose(h->fp); delete h; *f = nullptr; return 0; }

// ---- c-lime reader on stdio: 144-byte big-endian record headers, payload padded to 8 bytes -----------------------------------------
extern "C" {
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
}  // extern "C"
