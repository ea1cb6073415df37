// tmq_lime.cpp -- on-disk formats either side of the hot path, host code only:
//   * the ILDG / LIME gauge configuration reader of include/QKXTM_read_conf.h:107-400 (read_custom_binary_gauge_field):
//     records "xlf-info", "ildg-format" (<precision>, <lx>..<lt>) and "ildg-binary-data" = big-endian doubles ordered
//     [t][z][y][x][mu = x,y,z,t][3][3][re,im], scattered into the QDP even-odd host order loadGaugeQuda takes
//     (gauge[mu] = [even Vh | odd Vh] x 18, cb index = local lexicographic index / 2).  No boundary condition or scaling is
//     applied by the reader (:395-397); the drivers call applyGaugeFieldScaling afterwards;
//   * the "DiracFermion_Sink" propagator writer of QKXTM_Vector::write (lib/qudaQKXTM_Vector.cpp:510-702): records
//     "propagator-type", "quda-propagator-format" (etmcFormat XML) and "scidac-binary-data" = big-endian reals ordered
//     [t][z][y][x][spin][colour][re,im].
// The reference goes through c-lime and MPI-IO sub-array views; neither is available here, and neither is needed: a
// LIME record is a 144-byte big-endian header (magic 0x456789ab, version 1, MB/ME flags, 64-bit payload length, 128-byte
// type string) followed by the payload padded to 8 bytes, and every rank reads / writes its own x-rows of the global
// array with plain pread / pwrite at computed offsets.
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <fcntl.h>
#include <unistd.h>
#include "../../include/tmq_host.h"

namespace {

const uint32_t LIME_MAGIC = 0x456789abu;
const size_t LIME_HDR = 144;

thread_local char g_lime_err[512] = "";
int fail(const char *fmt, const char *a = "", long long b = 0) {
  snprintf(g_lime_err, sizeof(g_lime_err), fmt, a, b);
  return 1;
}

inline uint64_t bswap64(uint64_t v) { return __builtin_bswap64(v); }
inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }
bool host_is_big_endian() { const uint16_t one = 1; return *(const uint8_t *)&one == 0; }      // qcd_isBigEndian

struct Record {
  std::string type;
  uint64_t bytes = 0;
  off_t data_off = 0;      // file offset of the payload
  bool mb = false, me = false;
};

// walk the records of a LIME file
int scan(int fd, std::vector<Record> &recs) {
  off_t pos = 0;
  for (;;) {
    unsigned char h[LIME_HDR];
    const ssize_t n = pread(fd, h, LIME_HDR, pos);
    if (n == 0) break;
    if (n != (ssize_t)LIME_HDR) return fail("truncated LIME header%s at offset %lld", "", (long long)pos);
    uint32_t magic; memcpy(&magic, h, 4);
    if (!host_is_big_endian()) magic = bswap32(magic);
    if (magic != LIME_MAGIC) return fail("not a LIME record%s at offset %lld (bad magic)", "", (long long)pos);
    uint64_t len; memcpy(&len, h + 8, 8);
    if (!host_is_big_endian()) len = bswap64(len);
    Record r;
    r.mb = (h[6] & 0x80) != 0; r.me = (h[6] & 0x40) != 0;
    char t[129]; memcpy(t, h + 16, 128); t[128] = 0;
    r.type = t; r.bytes = len; r.data_off = pos + (off_t)LIME_HDR;
    recs.push_back(r);
    pos = r.data_off + (off_t)((len + 7) & ~(uint64_t)7);
  }
  return 0;
}

int write_header(FILE *f, const char *type, uint64_t bytes, bool mb, bool me) {
  unsigned char h[LIME_HDR];
  memset(h, 0, sizeof(h));
  uint32_t magic = LIME_MAGIC; uint16_t ver = 1; uint64_t len = bytes;
  if (!host_is_big_endian()) { magic = bswap32(magic); ver = (uint16_t)((ver >> 8) | (ver << 8)); len = bswap64(len); }
  memcpy(h, &magic, 4); memcpy(h + 4, &ver, 2);
  h[6] = (unsigned char)((mb ? 0x80 : 0) | (me ? 0x40 : 0));
  memcpy(h + 8, &len, 8);
  strncpy((char *)h + 16, type, 127);
  return fwrite(h, 1, LIME_HDR, f) == LIME_HDR ? 0 : 1;
}
int write_record(FILE *f, const char *type, const void *data, uint64_t bytes, bool mb, bool me) {
  if (write_header(f, type, bytes, mb, me)) return 1;
  if (bytes && fwrite(data, 1, bytes, f) != bytes) return 1;
  static const char pad[8] = {0};
  const size_t p = (size_t)((8 - bytes % 8) % 8);
  if (p && fwrite(pad, 1, p, f) != p) return 1;
  return 0;
}

// "<tag>value" lookup of qcd_getParam / "key = value" lookup of qcd_getParamComma (include/QKXTM_read_conf.h:30-98)
bool xml_int(const std::string &s, const char *tag, int *out) {
  const size_t p = s.find(tag);
  if (p == std::string::npos) return false;
  *out = atoi(s.c_str() + p + strlen(tag));
  return true;
}
bool key_double(const std::string &s, const char *key, double *out) {
  const size_t p = s.find(key);
  if (p == std::string::npos) return false;
  *out = atof(s.c_str() + p + strlen(key));
  return true;
}

std::string read_text(int fd, const Record &r) {
  std::string s((size_t)r.bytes, '\0');
  if (r.bytes && pread(fd, &s[0], (size_t)r.bytes, r.data_off) != (ssize_t)r.bytes) s.clear();
  return s;
}

// local x-rows of a global [T][Z][Y][X][site_doubles] array: (t,z,y) outer, contiguous X_loc*site_bytes inner
template <typename Fn>
int for_each_row(const int G[4], const int X[4], const int coord[4], size_t site_bytes, Fn fn) {
  for (int t = 0; t < X[3]; t++)
    for (int z = 0; z < X[2]; z++)
      for (int y = 0; y < X[1]; y++) {
        const long long gt = t + (long long)coord[3] * X[3], gz = z + (long long)coord[2] * X[2], gy = y + (long long)coord[1] * X[1];
        const long long gx0 = (long long)coord[0] * X[0];
        const off_t off = (off_t)((((gt * G[2] + gz) * G[1] + gy) * G[0] + gx0) * (long long)site_bytes);
        const long long lrow = ((long long)t * X[2] + z) * X[1] + y;      // local row index, x = 0
        if (fn(off, lrow * X[0])) return 1;
      }
  return 0;
}

}  // namespace

extern "C" {

const char *tmq_lime_last_error(void) { return g_lime_err; }

/* lattice extents / precision / kappa / mu advertised by a configuration file (ildg-format, xlf-info records) */
int tmq_lime_gauge_info(const char *fname, int globalX[4], int *precision_bits, double *kappa, double *mu) {
  const int fd = open(fname, O_RDONLY);
  if (fd < 0) return fail("Error reading configuration! Could not open %s for reading", fname);
  std::vector<Record> recs;
  if (scan(fd, recs)) { close(fd); return 1; }
  bool have_fmt = false;
  if (kappa) *kappa = 0;
  if (mu) *mu = 0;
  for (const Record &r : recs) {
    if (r.type == "ildg-format") {
      const std::string s = read_text(fd, r);
      int p = 64;
      have_fmt = xml_int(s, "<lx>", &globalX[0]) && xml_int(s, "<ly>", &globalX[1]) && xml_int(s, "<lz>", &globalX[2]) &&
                 xml_int(s, "<lt>", &globalX[3]);
      xml_int(s, "<precision>", &p);
      if (precision_bits) *precision_bits = p;
    } else if (r.type == "xlf-info") {
      const std::string s = read_text(fd, r);
      if (kappa) key_double(s, "kappa =", kappa);
      if (mu) key_double(s, "mu =", mu);
    }
  }
  close(fd);
  if (!have_fmt) return fail("no ildg-format record in %s", fname);
  return 0;
}

/* read_custom_binary_gauge_field: this rank's sub-block of the ildg-binary-data record -> QDP even-odd host order.
 * gauge[mu]: 2*Vh*18 doubles.  localX * grid must equal the extents in the file.                                   */
int tmq_lime_read_gauge(const char *fname, double *const gauge[4], const int localX[4], const int grid[4], const int coord[4]) {
  int G[4], prec = 64;
  if (tmq_lime_gauge_info(fname, G, &prec, nullptr, nullptr)) return 1;
  if (prec == 32) return fail("Unsupported precision 32 bits%s", "");                     // QKXTM_read_conf.h:228-231
  for (int d = 0; d < 4; d++)
    if (localX[d] * grid[d] != G[d]) return fail("lattice in %s does not match local extents x grid (dimension %lld)", fname, d);
  const long long lvol = (long long)G[0] * G[1] * G[2] * G[3];
  if (lvol == 0) return fail("Zero volume%s", "");
  const int fd = open(fname, O_RDONLY);
  if (fd < 0) return fail("Could not open %s", fname);
  std::vector<Record> recs;
  if (scan(fd, recs)) { close(fd); return 1; }
  const Record *bin = nullptr;
  for (const Record &r : recs) if (r.type == "ildg-binary-data") { bin = &r; break; }
  if (!bin) { close(fd); return fail("no ildg-binary-data record in %s", fname); }
  if (bin->bytes != (uint64_t)lvol * 72 * sizeof(double)) { close(fd); return fail("Error, could not read proper amount of data%s (record holds %lld bytes)", "", (long long)bin->bytes); }
  const long long Vh = (long long)localX[0] * localX[1] * localX[2] * localX[3] / 2;
  std::vector<uint64_t> row((size_t)localX[0] * 72);
  const bool swap = !host_is_big_endian();
  const int rc = for_each_row(G, localX, coord, 72 * sizeof(double), [&](off_t off, long long lsite0) -> int {
    const size_t nb = row.size() * 8;
    if (pread(fd, row.data(), nb, bin->data_off + off) != (ssize_t)nb) return fail("short read in %s", fname);
    if (swap) for (uint64_t &v : row) v = bswap64(v);                                       // qcd_swap_8
    // local coordinates of the row start: parity of (x+y+z+t) with GLOBAL coordinates (x1+x2+x3+x4, :354)
    const long long l = lsite0;
    const int y = (int)((l / localX[0]) % localX[1]), z = (int)((l / ((long long)localX[0] * localX[1])) % localX[2]);
    const int t = (int)(l / ((long long)localX[0] * localX[1] * localX[2]));
    const long long gsum0 = (long long)coord[0] * localX[0] + (long long)coord[1] * localX[1] + y + (long long)coord[2] * localX[2] + z +
                            (long long)coord[3] * localX[3] + t;
    for (int x = 0; x < localX[0]; x++) {
      const int odd = (int)((gsum0 + x) & 1);
      const long long cb = (l + x) / 2;
      for (int mu = 0; mu < 4; mu++)
        memcpy(gauge[mu] + ((long long)odd * Vh + cb) * 18, &row[((size_t)x * 4 + mu) * 18], 18 * sizeof(double));
    }
    return 0;
  });
  close(fd);
  return rc;
}

/* the inverse, for tests and for exporting synthetic configurations: rank 0 (coord all zero) creates the file with the
 * xlf-info / ildg-format / ildg-binary-data records, every rank then fills in its sub-block (call rank 0 first).   */
int tmq_lime_write_gauge(const char *fname, const double *const gauge[4], const int localX[4], const int grid[4], const int coord[4],
                         double kappa, double mu) {
  int G[4];
  for (int d = 0; d < 4; d++) G[d] = localX[d] * grid[d];
  const long long lvol = (long long)G[0] * G[1] * G[2] * G[3];
  const bool first = coord[0] == 0 && coord[1] == 0 && coord[2] == 0 && coord[3] == 0;
  if (first) {
    FILE *f = fopen(fname, "wb");
    if (!f) return fail("could not create %s", fname);
    char xlf[512], fmt[512];
    snprintf(xlf, sizeof(xlf), "plaquette = 0.0\n trajectory nr = 0\n beta = 0.0, kappa = %.12f, mu = %.12f, c2_rec = 0.0\n", kappa, mu);
    snprintf(fmt, sizeof(fmt), "<?xml version=\"1.0\" encoding=\"UTF-8\"?>\n<ildgFormat>\n<version>1.0</version>\n<field>su3gauge</field>\n"
             "<precision>64</precision>\n<lx>%d</lx>\n<ly>%d</ly>\n<lz>%d</lz>\n<lt>%d</lt>\n</ildgFormat>", G[0], G[1], G[2], G[3]);
    int rc = write_record(f, "xlf-info", xlf, strlen(xlf), true, true);
    rc |= write_record(f, "ildg-format", fmt, strlen(fmt), true, false);
    rc |= write_header(f, "ildg-binary-data", (uint64_t)lvol * 72 * 8, false, true);
    if (!rc) rc = (fflush(f) != 0) || (ftruncate(fileno(f), ftell(f) + (off_t)lvol * 72 * 8) != 0);
    fclose(f);
    if (rc) return fail("error writing the LIME headers of %s", fname);
  }
  const int fd = open(fname, O_RDWR);
  if (fd < 0) return fail("Could not open %s", fname);
  std::vector<Record> recs;
  if (scan(fd, recs)) { close(fd); return 1; }
  const Record *bin = nullptr;
  for (const Record &r : recs) if (r.type == "ildg-binary-data") { bin = &r; break; }
  if (!bin) { close(fd); return fail("no ildg-binary-data record in %s", fname); }
  const long long Vh = (long long)localX[0] * localX[1] * localX[2] * localX[3] / 2;
  std::vector<uint64_t> row((size_t)localX[0] * 72);
  const bool swap = !host_is_big_endian();
  const int rc = for_each_row(G, localX, coord, 72 * sizeof(double), [&](off_t off, long long l) -> int {
    const int y = (int)((l / localX[0]) % localX[1]), z = (int)((l / ((long long)localX[0] * localX[1])) % localX[2]);
    const int t = (int)(l / ((long long)localX[0] * localX[1] * localX[2]));
    const long long gsum0 = (long long)coord[0] * localX[0] + (long long)coord[1] * localX[1] + y + (long long)coord[2] * localX[2] + z +
                            (long long)coord[3] * localX[3] + t;
    for (int x = 0; x < localX[0]; x++) {
      const int odd = (int)((gsum0 + x) & 1);
      const long long cb = (l + x) / 2;
      for (int mu_ = 0; mu_ < 4; mu_++)
        memcpy(&row[((size_t)x * 4 + mu_) * 18], gauge[mu_] + ((long long)odd * Vh + cb) * 18, 18 * sizeof(double));
    }
    if (swap) for (uint64_t &v : row) v = bswap64(v);
    const size_t nb = row.size() * 8;
    return pwrite(fd, row.data(), nb, bin->data_off + off) == (ssize_t)nb ? 0 : fail("short write in %s", fname);
  });
  close(fd);
  return rc;
}

/* QKXTM_Vector::write: host vector in the plug-in's AoS order [x_lex][spin][colour][re,im] (local) -> "DiracFermion_Sink"
 * file of the GLOBAL lattice; prec = 8 | 4 selects the precision of both the source buffer and the file.           */
/* step 1 (ONE rank): create the file -- an existing one of that name is replaced atomically, never written into -- with the three
 * record headers and room for the payload.  On a process grid every rank must wait (a barrier) before step 2.                 */
int tmq_lime_write_vector_header(const char *fname, int prec, const int localX[4], const int grid[4]) {
  if (prec != 8 && prec != 4) return fail("bad precision%s", "");
  int G[4];
  for (int d = 0; d < 4; d++) G[d] = localX[d] * grid[d];
  const long long lvol = (long long)G[0] * G[1] * G[2] * G[3];
  {
    const std::string tmpname = std::string(fname) + ".tmp";
    FILE *f = fopen(tmpname.c_str(), "wb");
    if (!f) return fail("Error open file to write propagator %s", fname);
    const char *ptype = "DiracFermion_Sink";
    char fmt[1024];
    snprintf(fmt, sizeof(fmt), "<?xml version=\"1.0\" encoding=\"UTF-8\"?>\n<etmcFormat>\n\t<field>diracFermion</field>\n\t<precision>%d</precision>\n"
             "\t<flavours>1</flavours>\n\t<lx>%d</lx>\n\t<ly>%d</ly>\n\t<lz>%d</lz>\n\t<lt>%d</lt>\n\t<spin>4</spin>\n\t<colour>3</colour>\n</etmcFormat>",
             prec * 8, G[0], G[1], G[2], G[3]);
    int rc = write_record(f, "propagator-type", ptype, strlen(ptype), true, true);
    rc |= write_record(f, "quda-propagator-format", fmt, strlen(fmt), true, true);
    rc |= write_header(f, "scidac-binary-data", (uint64_t)lvol * 24 * prec, true, true);
    if (!rc) rc = (fflush(f) != 0) || (ftruncate(fileno(f), ftell(f) + (off_t)lvol * 24 * prec) != 0);
    fclose(f);
    if (!rc) rc = rename(tmpname.c_str(), fname) != 0;
    if (rc) return fail("LIME write header error in %s", fname);
  }
  return 0;
}
/* step 2 (EVERY rank, after step 1 is complete): write this rank's sub-block at its own offsets                                 */
int tmq_lime_write_vector_block(const char *fname, const void *h_aos, int prec, const int localX[4], const int grid[4], const int coord[4]) {
  if (prec != 8 && prec != 4) return fail("bad precision%s", "");
  int G[4];
  for (int d = 0; d < 4; d++) G[d] = localX[d] * grid[d];
  const int fd = open(fname, O_RDWR);
  if (fd < 0) return fail("Could not open %s", fname);
  std::vector<Record> recs;
  if (scan(fd, recs)) { close(fd); return 1; }
  const Record *bin = nullptr;
  for (const Record &r : recs) if (r.type == "scidac-binary-data") { bin = &r; break; }
  if (!bin) { close(fd); return fail("no scidac-binary-data record in %s", fname); }
  if (bin->bytes != (uint64_t)G[0] * G[1] * G[2] * G[3] * 24 * prec) { close(fd); return fail("the payload record of %s does not match the lattice (stale file?)", fname); }
  const size_t site = (size_t)24 * prec;
  std::vector<unsigned char> row((size_t)localX[0] * site);
  const bool swap = !host_is_big_endian();
  const int rc = for_each_row(G, localX, coord, site, [&](off_t off, long long l) -> int {
    memcpy(row.data(), (const unsigned char *)h_aos + (size_t)l * site, row.size());
    if (swap) {
      if (prec == 8) { uint64_t *p = (uint64_t *)row.data(); for (size_t i = 0; i < row.size() / 8; i++) p[i] = bswap64(p[i]); }
      else { uint32_t *p = (uint32_t *)row.data(); for (size_t i = 0; i < row.size() / 4; i++) p[i] = bswap32(p[i]); }
    }
    return pwrite(fd, row.data(), row.size(), bin->data_off + off) == (ssize_t)row.size() ? 0 : fail("short write in %s", fname);
  });
  close(fd);
  return rc;
}
/* both steps in one call, for a single process (or for callers that serialise the ranks themselves, rank (0,0,0,0) first)         */
int tmq_lime_write_vector(const char *fname, const void *h_aos, int prec, const int localX[4], const int grid[4], const int coord[4]) {
  const bool first = coord[0] == 0 && coord[1] == 0 && coord[2] == 0 && coord[3] == 0;
  if (first && tmq_lime_write_vector_header(fname, prec, localX, grid)) return 1;
  return tmq_lime_write_vector_block(fname, h_aos, prec, localX, grid, coord);
}

/* reads a DiracFermion_Sink file back into the local AoS order (the reference has the matching reader for sources in
 * its 2pt/3pt drivers); also reports the extents and precision stored in the file                                  */
int tmq_lime_read_vector(const char *fname, void *h_aos, int prec, const int localX[4], const int grid[4], const int coord[4]) {
  if (prec != 8 && prec != 4) return fail("bad precision%s", "");
  const int fd = open(fname, O_RDONLY);
  if (fd < 0) return fail("Could not open %s", fname);
  std::vector<Record> recs;
  if (scan(fd, recs)) { close(fd); return 1; }
  int G[4] = {0, 0, 0, 0}, fprec = 0;
  const Record *bin = nullptr;
  std::string ptype;
  for (const Record &r : recs) {
    if (r.type == "quda-propagator-format" || r.type == "etmc-propagator-format") {
      const std::string s = read_text(fd, r);
      xml_int(s, "<lx>", &G[0]); xml_int(s, "<ly>", &G[1]); xml_int(s, "<lz>", &G[2]); xml_int(s, "<lt>", &G[3]); xml_int(s, "<precision>", &fprec);
    } else if (r.type == "propagator-type") ptype = read_text(fd, r);
    else if (r.type == "scidac-binary-data" && !bin) bin = &r;
  }
  if (!bin) { close(fd); return fail("no scidac-binary-data record in %s", fname); }
  if (fprec != prec * 8) { close(fd); return fail("precision in %s does not match the request (%lld bits)", fname, fprec); }
  for (int d = 0; d < 4; d++)
    if (localX[d] * grid[d] != G[d]) { close(fd); return fail("lattice in %s does not match local extents x grid (dimension %lld)", fname, d); }
  const size_t site = (size_t)24 * prec;
  const bool swap = !host_is_big_endian();
  const int rc = for_each_row(G, localX, coord, site, [&](off_t off, long long l) -> int {
    unsigned char *dst = (unsigned char *)h_aos + (size_t)l * site;
    const size_t nb = (size_t)localX[0] * site;
    if (pread(fd, dst, nb, bin->data_off + off) != (ssize_t)nb) return fail("short read in %s", fname);
    if (swap) {
      if (prec == 8) { uint64_t *p = (uint64_t *)dst; for (size_t i = 0; i < nb / 8; i++) p[i] = bswap64(p[i]); }
      else { uint32_t *p = (uint32_t *)dst; for (size_t i = 0; i < nb / 4; i++) p[i] = bswap32(p[i]); }
    }
    return 0;
  });
  close(fd);
  return rc;
}

}  // extern "C"
