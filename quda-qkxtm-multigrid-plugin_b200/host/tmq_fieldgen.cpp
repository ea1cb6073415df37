// tmq_fieldgen.cpp -- synthetic lattice inputs for drivers, tests and bench.py (host code, OpenMP).
//
// Restates the semantics of the reference's test helpers with a portable counter-based RNG:
//   random SU(3): rows 1,2 uniform in [0,1) -> normalise, Gram-Schmidt, normalise; row 0 = conj cross
//   product (qkxtm/QKXTM_util.cpp:879-955); QDP even-odd order [even Vh | odd Vh] x 3x3 row-major
//   (qkxtm/QKXTM_util.cpp:840-857); anti-periodic T folded into U_t on the last global time slice
//   (qkxtm/QKXTM_util.cpp:698-705); Z4 noise source (lib/qudaQKXTM_utils.cpp:148-180).
// The RNG is keyed by (seed, stream, GLOBAL lexicographic site, component), so every sharding of a global
// lattice sees the same field; the reference's libc rand() order dependence is deliberately not reproduced.
// tests/lattice_util.py holds the same generator in numpy; tests/test_host.py checks that they agree.
#include "../../include/tmq_host.h"
#include <cmath>
#include <complex>
#include <cstdint>

namespace {

inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
struct Key {
  uint64_t k;   // mix(mix(seed) ^ stream)
  Key(uint64_t seed, uint64_t stream) { k = mix64(mix64(seed) ^ stream); }
  double uniform(uint64_t site, uint64_t comp) const {
    const uint64_t h = mix64(mix64(k ^ site) ^ comp);
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
  }
};

struct Lat {
  int X[4], G[4], off[4];
  long long V;
  Lat(const int localX[4], const int grid[4], const int coord[4]) {
    V = 1;
    for (int d = 0; d < 4; d++) { X[d] = localX[d]; G[d] = localX[d] * grid[d]; off[d] = coord[d] * localX[d]; V *= localX[d]; }
  }
  // local lexicographic index -> (global lexicographic index, even-odd index, local t)
  inline void map(long long i, uint64_t &gl, long long &eo, int &t) const {
    const int x = (int)(i % X[0]), y = (int)((i / X[0]) % X[1]), z = (int)((i / ((long long)X[0] * X[1])) % X[2]);
    t = (int)(i / ((long long)X[0] * X[1] * X[2]));
    gl = (uint64_t)(x + off[0]) + (uint64_t)G[0] * ((uint64_t)(y + off[1]) + (uint64_t)G[1] * ((uint64_t)(z + off[2]) + (uint64_t)G[2] * (uint64_t)(t + off[3])));
    const int par = (x + y + z + t) & 1;
    eo = (long long)par * (V / 2) + i / 2;
  }
};

typedef std::complex<double> cd;

}  // namespace

extern "C" {

void tmq_fieldgen_gauge_qdp(double *const gauge[4], const int localX[4], const int grid[4], const int coord[4],
                            unsigned long long seed, int t_boundary) {
  const Lat L(localX, grid, coord);
  const bool last_t = coord[3] == grid[3] - 1;
  for (int mu = 0; mu < 4; mu++) {
    const Key key(seed, (uint64_t)mu);
    double *out = gauge[mu];
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < L.V; i++) {
      uint64_t gl; long long eo; int t;
      L.map(i, gl, eo, t);
      cd u[3], v[3], w[3];
      for (int n = 0; n < 3; n++) {
        u[n] = cd(key.uniform(gl, (uint64_t)(n * 2)), key.uniform(gl, (uint64_t)(n * 2 + 1)));
        v[n] = cd(key.uniform(gl, (uint64_t)((3 + n) * 2)), key.uniform(gl, (uint64_t)((3 + n) * 2 + 1)));
      }
      double nu = 0; for (int n = 0; n < 3; n++) nu += std::norm(u[n]);
      nu = std::sqrt(nu); for (int n = 0; n < 3; n++) u[n] /= nu;
      cd dot = 0; for (int n = 0; n < 3; n++) dot += std::conj(u[n]) * v[n];
      for (int n = 0; n < 3; n++) v[n] -= dot * u[n];
      double nv = 0; for (int n = 0; n < 3; n++) nv += std::norm(v[n]);
      nv = std::sqrt(nv); for (int n = 0; n < 3; n++) v[n] /= nv;
      w[0] = std::conj(u[1] * v[2] - u[2] * v[1]);
      w[1] = std::conj(u[2] * v[0] - u[0] * v[2]);
      w[2] = std::conj(u[0] * v[1] - u[1] * v[0]);
      const double sgn = (mu == 3 && t_boundary == -1 && last_t && t == L.X[3] - 1) ? -1.0 : 1.0;
      double *o = out + eo * 18;
      for (int n = 0; n < 3; n++) {
        o[n * 2] = sgn * w[n].real(); o[n * 2 + 1] = sgn * w[n].imag();
        o[6 + n * 2] = sgn * u[n].real(); o[6 + n * 2 + 1] = sgn * u[n].imag();
        o[12 + n * 2] = sgn * v[n].real(); o[12 + n * 2 + 1] = sgn * v[n].imag();
      }
    }
  }
}

void tmq_fieldgen_unit_gauge_qdp(double *const gauge[4], const int localX[4], const int grid[4], const int coord[4],
                                 int t_boundary) {
  const Lat L(localX, grid, coord);
  const bool last_t = coord[3] == grid[3] - 1;
  for (int mu = 0; mu < 4; mu++) {
    double *out = gauge[mu];
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < L.V; i++) {
      uint64_t gl; long long eo; int t;
      L.map(i, gl, eo, t);
      const double sgn = (mu == 3 && t_boundary == -1 && last_t && t == L.X[3] - 1) ? -1.0 : 1.0;
      double *o = out + eo * 18;
      for (int k = 0; k < 18; k++) o[k] = 0.0;
      o[0] = o[8] = o[16] = sgn;
    }
  }
}

// dense Gaussian spinor, written in even-odd order [even Vh | odd Vh][4][3][2] (eo = 1) or in the plug-in's
// host order [x_lex][4][3][2] (eo = 0; lib/qudaQKXTM_Vector.cpp:72-81)
void tmq_fieldgen_spinor_gaussian(double *out, const int localX[4], const int grid[4], const int coord[4],
                                  unsigned long long seed, int eo_order) {
  const Lat L(localX, grid, coord);
  const Key key(seed, 16);
  const double two_pi = 2.0 * 3.14159265358979323846;
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < L.V; i++) {
    uint64_t gl; long long eo; int t;
    L.map(i, gl, eo, t);
    double *o = out + (eo_order ? eo : i) * 24;
    for (int k = 0; k < 24; k++) {
      const double u1 = key.uniform(gl, (uint64_t)(2 * k)), u2 = key.uniform(gl, (uint64_t)(2 * k + 1));
      o[k] = std::sqrt(-2.0 * std::log(1.0 - u1)) * std::cos(two_pi * u2);
    }
  }
}

// Z4 noise: 0 -> +1, 1 -> -1, 2 -> +i, 3 -> -i per spin-colour (lib/qudaQKXTM_utils.cpp:153-174)
void tmq_fieldgen_spinor_z4(double *out, const int localX[4], const int grid[4], const int coord[4],
                            unsigned long long seed, int eo_order) {
  const Lat L(localX, grid, coord);
  const Key key(seed, 17);
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < L.V; i++) {
    uint64_t gl; long long eo; int t;
    L.map(i, gl, eo, t);
    double *o = out + (eo_order ? eo : i) * 24;
    for (int k = 0; k < 12; k++) {
      const int r = (int)std::floor(key.uniform(gl, (uint64_t)k) * 4.0);
      o[2 * k] = r == 0 ? 1.0 : (r == 1 ? -1.0 : 0.0);
      o[2 * k + 1] = r == 2 ? 1.0 : (r == 3 ? -1.0 : 0.0);
    }
  }
}

// applyGaugeFieldScaling with anisotropy 1 (qkxtm/QKXTM_util.cpp:682-725): only the anti-periodic sign on U_t(T-1) remains
void tmq_apply_t_boundary(double *const gauge[4], const int localX[4], const int grid[4], const int coord[4], int t_boundary) {
  if (t_boundary != -1 || coord[3] != grid[3] - 1) return;
  const long long Vh = (long long)localX[0] * localX[1] * localX[2] * localX[3] / 2;
  const long long slice_h = (long long)localX[0] * localX[1] * localX[2] / 2;
  // the last local time slice occupies the last slice_h checkerboard sites of each parity block
  for (int par = 0; par < 2; par++)
    for (long long i = Vh - slice_h; i < Vh; i++) {
      double *u = gauge[3] + (par * Vh + i) * 18;
      for (int k = 0; k < 18; k++) u[k] = -u[k];
    }
}

}  // extern "C"
