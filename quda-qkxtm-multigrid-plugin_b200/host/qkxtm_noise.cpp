// qkxtm_noise.cpp -- the stochastic sources of calc_loops (lib/qudaQKXTM_interface.cpp:1951,1982-2005): Z4 / unity noise vectors,
// spin-colour dilution and hierarchical probing, on host vectors in the plug-in's AoS order [x_lex][spin][colour][re,im].
// The reference takes its random numbers from GSL (gsl_rng_ranlux, luxury level 223, seeded seed + rank * seed).  GSL is a
// third-party dependency that is absent here, so the generator is restated from the published algorithm (M. Luescher, Comput.
// Phys. Commun. 79 (1994) 100; F. James, ibid. 79 (1994) 111: 24-bit subtract-with-borrow lags (24, 10), 24 numbers delivered
// out of every 223) and pinned to GSL's own known-answer test: seed 314159265, the 10000th number is 12077992
// (tests/test_calc_loops_noise.py).  With it a noise vector is bit-identical to the reference's for the same seed and rank.
#include "../../include/qudaQKXTM_tmq.h"
#include "../../include/tmq_host.h"
#include "qkxtm_internal.h"
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace quda {

namespace {
struct Ranlux {
  unsigned int i, j, n, skip, carry;
  unsigned long int u[24];
};
const unsigned long int kMaskLo = 0x00ffffffUL, kTwo24 = 16777216UL;

inline unsigned long int ranlux_step(Ranlux *s) {
  long int delta = (long int)s->u[s->j] - (long int)s->u[s->i] - (long int)s->carry;
  if (delta < 0) { s->carry = 1; delta += (long int)kTwo24; }
  else s->carry = 0;
  s->u[s->i] = (unsigned long int)delta;
  s->i = s->i == 0 ? 23 : s->i - 1;
  s->j = s->j == 0 ? 23 : s->j - 1;
  return (unsigned long int)delta;
}
}  // namespace

}  // namespace quda

extern "C" {
void *tmq_ranlux_alloc(unsigned long seed_in) {
  using namespace quda;
  Ranlux *s = (Ranlux *)calloc(1, sizeof(Ranlux));
  if (!s) return NULL;
  long int seed = (long int)(seed_in == 0 ? 314159265UL : seed_in);
  for (int k = 0; k < 24; k++) {     // the seeding recurrence of F. James' RANLUX (a 31-bit linear congruential generator)
    const long int q = seed / 53668;
    seed = 40014 * (seed - q * 53668) - q * 12211;
    if (seed < 0) seed += 2147483563;
    s->u[k] = (unsigned long int)seed % kTwo24;
  }
  s->i = 23; s->j = 9; s->n = 0; s->skip = 223 - 24;
  s->carry = (s->u[23] & ~kMaskLo) ? 1 : 0;
  return s;
}
void tmq_ranlux_free(void *rng) { free(rng); }
unsigned long tmq_ranlux_get(void *rng) {
  using namespace quda;
  Ranlux *s = (Ranlux *)rng;
  const unsigned long int r = ranlux_step(s);
  if (++s->n == 24) {
    s->n = 0;
    for (unsigned int k = 0; k < s->skip; k++) ranlux_step(s);
  }
  return r;
}
// gsl_rng_uniform_int: the range [0, 2^24 - 1] is cut into n equal bins, numbers beyond the last bin are rejected
unsigned long tmq_ranlux_uniform_int(void *rng, unsigned long n) {
  const unsigned long int range = quda::kMaskLo;
  if (n == 0 || n > range) return 0;
  const unsigned long int scale = range / n;
  unsigned long int k;
  do { k = tmq_ranlux_get(rng) / scale; } while (k >= n);
  return k;
}
// lib/qudaQKXTM_utils.cpp:148-180 on ncomplex complex numbers: one draw per component in BOTH modes (UNITY discards it)
void tmq_noise_z4(double *out, long long ncomplex, void *rng, int unity) {
  memset(out, 0, (size_t)ncomplex * 2 * sizeof(double));
  for (long long i = 0; i < ncomplex; i++) {
    const unsigned long r = tmq_ranlux_uniform_int(rng, 4);
    if (unity) out[2 * i] = 1.0;
    else if (r == 0) out[2 * i] = 1.0;
    else if (r == 1) out[2 * i] = -1.0;
    else if (r == 2) out[2 * i + 1] = 1.0;
    else out[2 * i + 1] = -1.0;
  }
}
// the colours of the hierarchical probing for a local lattice L[0..d-1] (x fastest); returns 0 on success
int tmq_hch_coloring(unsigned short *Vc, const int *L, int k, int d) {
  if ((d != 3 && d != 4) || k < 1) return 1;
  const int Lu = 1 << (k - 1);
  for (int i = 0; i < d; i++)
    if (L[i] % (2 * Lu) != 0) return 2;
  if (2.0 * std::pow(2.0, d * (k - 1)) > 65536.0) return 3;
  long long len = 1;
  for (int i = 0; i < d; i++) len *= L[i];
  for (long long idx = 0; idx < len; idx++) {
    long long rest = idx;
    int inblock = 0, stride = 1, blocksum = 0;
    for (int i = 0; i < d; i++) {           // x fastest, both in the lattice and inside the block (get_ind2Vec / get_vec2Idx)
      const int xi = (int)(rest % L[i]);
      rest /= L[i];
      blocksum += xi / Lu;
      inblock += (xi % Lu) * stride;
      stride *= Lu;
    }
    Vc[idx] = (unsigned short)(2 * inblock + (blocksum & 1));
  }
  return 0;
}
int tmq_hadamard_element(int i, int j) { return (__builtin_popcount((unsigned)i & (unsigned)j) & 1) ? -1 : 1; }
}  // extern "C"

namespace quda {
void *qkxtm_rng_alloc(unsigned long int seed) { return tmq_ranlux_alloc(seed); }
void qkxtm_rng_free(void *rng) { tmq_ranlux_free(rng); }
unsigned long int qkxtm_rng_get(void *rng) { return tmq_ranlux_get(rng); }
unsigned long int qkxtm_rng_uniform_int(void *rng, unsigned long int n) { return tmq_ranlux_uniform_int(rng, n); }

// lib/qudaQKXTM_utils.cpp:148-180: one draw per complex component in BOTH modes (UNITY discards it, keeping the streams aligned)
template <typename Float> void getStochasticRandomSource(void *spinorIn, void *rng, SOURCE_T source_type) {
  const long long n = qkxtm_local_volume() * 12;
  Float *v = (Float *)spinorIn;
  memset(v, 0, (size_t)n * 2 * sizeof(Float));
  for (long long i = 0; i < n; i++) {
    const unsigned long int r = qkxtm_rng_uniform_int(rng, 4);
    if (source_type == UNITY) v[2 * i] = (Float)1;
    else if (source_type == RANDOM) {
      if (r == 0) v[2 * i] = (Float)1;
      else if (r == 1) v[2 * i] = (Float)-1;
      else if (r == 2) v[2 * i + 1] = (Float)1;
      else v[2 * i + 1] = (Float)-1;
    } else qkxtm_raise("Source type not set correctly!! Aborting.");
  }
}
template void getStochasticRandomSource<double>(void *, void *, SOURCE_T);
template void getStochasticRandomSource<float>(void *, void *, SOURCE_T);

// ---- hierarchical probing (lib/qudaQKXTM_utils.cpp:476-717): 2 * 2^{d(k-1)} colours; the lattice is tiled with blocks of extent
//      Lu = 2^{k-1}, a site's colour is 2 * (its lexicographic position inside the block) + (parity of the block) ---------------
unsigned short int *hch_coloring(int k, int d) {
  if (d != 3 && d != 4) qkxtm_raise("Only 3 and 4 dimensions of coloring are allowed");
  if (k < 1) qkxtm_raise("k must be greater than 1");
  const int *L = qkxtm_local_extent();
  long long len = 1;
  for (int i = 0; i < d; i++) len *= L[i];
  unsigned short int *Vc = (unsigned short int *)malloc(sizeof(unsigned short int) * (size_t)len);
  if (!Vc) qkxtm_raise("hch_coloring: out of memory");
  const int rc = tmq_hch_coloring(Vc, L, k, d);
  if (rc == 2) qkxtm_raise("2*Lu cannot fit in the local lattice extent");
  if (rc == 3) qkxtm_raise("Exceeded maximum number of colors");
  return Vc;
}

int HadamardElements(int i, int j) { return tmq_hadamard_element(i, j); }   // (-1)^{popcount(i & j)}: Sylvester-Hadamard

template <typename Float> void get_probing4D_spinColor_dilution(void *temp_input_vector, void *input_vector, unsigned short int *Vc, int ih, int sc) {
  const long long V = qkxtm_local_volume();
  Float *out = (Float *)temp_input_vector;
  const Float *in = (const Float *)input_vector;
  memset(out, 0, (size_t)V * 24 * sizeof(Float));
  for (long long i = 0; i < V; i++) {
    const Float sg = (Float)HadamardElements(Vc[i], ih);
    for (int ri = 0; ri < 2; ri++) out[i * 24 + sc * 2 + ri] = sg * in[i * 24 + sc * 2 + ri];
  }
}
template <typename Float> void get_spinColor_dilution(void *temp_input_vector, void *input_vector, int sc) {
  const long long V = qkxtm_local_volume();
  Float *out = (Float *)temp_input_vector;
  const Float *in = (const Float *)input_vector;
  memset(out, 0, (size_t)V * 24 * sizeof(Float));
  for (long long i = 0; i < V; i++)
    for (int ri = 0; ri < 2; ri++) out[i * 24 + sc * 2 + ri] = in[i * 24 + sc * 2 + ri];
}
template <typename Float> void get_probing4D_dilution(void *temp_input_vector, void *input_vector, unsigned short int *Vc, int ih) {
  const long long V = qkxtm_local_volume();
  Float *out = (Float *)temp_input_vector;
  const Float *in = (const Float *)input_vector;
  for (long long i = 0; i < V; i++) {
    const Float sg = (Float)HadamardElements(Vc[i], ih);
    for (int k = 0; k < 24; k++) out[i * 24 + k] = sg * in[i * 24 + k];
  }
}
template void get_probing4D_spinColor_dilution<double>(void *, void *, unsigned short int *, int, int);
template void get_spinColor_dilution<double>(void *, void *, int);
template void get_probing4D_dilution<double>(void *, void *, unsigned short int *, int);
template void get_probing4D_spinColor_dilution<float>(void *, void *, unsigned short int *, int, int);
template void get_spinColor_dilution<float>(void *, void *, int);
template void get_probing4D_dilution<float>(void *, void *, unsigned short int *, int);

}  // namespace quda
