// qkxtm_loops_host.cpp -- TEST INFRASTRUCTURE: compiles the REFERENCE'S OWN noise / dilution / hierarchical-probing routines of
// lib/qudaQKXTM_utils.cpp (getStochasticRandomSource :148-180, fcb / get_ind2Vec / get_vec2Idx / create_hch_coloring :499-577,
// hch_coloring / HadamardElements / get_*_dilution :666-752).  oracle/Makefile cuts exactly those line ranges out of the reference
// file where it lies into oracle/_ref/qkxtm_utils_noise.inc (git-ignored, nothing is copied into the repository) and this file
// supplies what they reference: the GK_* geometry globals, errorQuda / printfQuda, and a gsl_rng stand-in that replays a stream
// of integers handed over by the test (GSL itself is absent; the RANLUX generator is pinned separately to GSL's known-answer test).
// Built into oracle/_ref/libqkxtm_loops_ref.so; used by tests/test_calc_loops_noise.py to pin host/qkxtm_noise.cpp.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <typeinfo>

static int GK_localL[4], GK_totalL[4];
static long int GK_localVolume;
#define errorQuda(...) do { fprintf(stderr, __VA_ARGS__); abort(); } while (0)
#define printfQuda(...) do { } while (0)
enum SOURCE_T { UNITY, RANDOM };
struct gsl_rng { const int *stream; long n, pos; };
static unsigned long gsl_rng_uniform_int(gsl_rng *r, unsigned long) {
  if (r->pos >= r->n) { fprintf(stderr, "oracle/_ref: random stream exhausted\n"); abort(); }
  return (unsigned long)r->stream[r->pos++];
}

#include "qkxtm_utils_noise.inc"      // cut from /root/reference/lib/qudaQKXTM_utils.cpp by oracle/Makefile

extern "C" {
void qloops_set_lattice(const int L[4]) {
  GK_localVolume = 1;
  for (int d = 0; d < 4; d++) { GK_localL[d] = GK_totalL[d] = L[d]; GK_localVolume *= L[d]; }
}
// stream: GK_localVolume * 12 integers in [0, 4), the values gsl_rng_uniform_int(rNum, 4) would return
void qloops_stochastic_source(double *out, const int *stream, long n, int source_type) {
  gsl_rng r = {stream, n, 0};
  getStochasticRandomSource<double>(out, &r, source_type ? RANDOM : UNITY);
}
void qloops_hch_coloring(unsigned short *out, int k, int d) {
  unsigned short int *Vc = hch_coloring(k, d);
  long len = (long)GK_localL[0] * GK_localL[1] * GK_localL[2] * (d == 4 ? GK_localL[3] : 1);
  memcpy(out, Vc, sizeof(unsigned short) * len);
  free(Vc);
}
int qloops_hadamard(int i, int j) { return HadamardElements(i, j); }
void qloops_probing4D_spinColor_dilution(double *out, double *in, unsigned short *Vc, int ih, int sc) { get_probing4D_spinColor_dilution<double>(out, in, Vc, ih, sc); }
void qloops_spinColor_dilution(double *out, double *in, int sc) { get_spinColor_dilution<double>(out, in, sc); }
void qloops_probing4D_dilution(double *out, double *in, unsigned short *Vc, int ih) { get_probing4D_dilution<double>(out, in, Vc, ih); }
}
