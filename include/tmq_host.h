/*
 * tmq_host.h -- C ABI of libqkxtm_tmq.so, the host-side (CPU, no CUDA) companion of libtmq.so:
 * synthetic-field generators with the semantics of the reference's test helpers.  The C++ QKXTM shim
 * (qudaQKXTM_tmq.h: QKXTM_Field / QKXTM_Gauge / QKXTM_Vector / QKXTM_Propagator, init_qudaQKXTM, loadGaugeQuda,
 * the MG_bench / calc_loops solve skeletons) lives in the same library and is C++-only.
 */
#ifndef TMQ_HOST_H
#define TMQ_HOST_H
#ifdef __cplusplus
extern "C" {
#endif

/* random SU(3) field in QDP even-odd order (constructGaugeField semantics, qkxtm/QKXTM_util.cpp:879-955,
 * 840-857) for the rank at `coord` of `grid`, with the T boundary condition folded into U_t on the last
 * GLOBAL time slice when t_boundary = -1 (applyGaugeFieldScaling, qkxtm/QKXTM_util.cpp:698-705).
 * gauge[mu]: 2*Vh*18 doubles each. */
void tmq_fieldgen_gauge_qdp(double *const gauge[4], const int localX[4], const int grid[4], const int coord[4],
                            unsigned long long seed, int t_boundary);
void tmq_fieldgen_unit_gauge_qdp(double *const gauge[4], const int localX[4], const int grid[4], const int coord[4],
                                 int t_boundary);
/* spinors, V*24 doubles; eo_order = 1: [even Vh | odd Vh][4][3][2] (include/QKXTM_mapping_parity.h:67-110),
 * 0: [x_lex][4][3][2] (lib/qudaQKXTM_Vector.cpp:72-81).  Z4 noise: lib/qudaQKXTM_utils.cpp:148-180. */
void tmq_fieldgen_spinor_gaussian(double *out, const int localX[4], const int grid[4], const int coord[4],
                                  unsigned long long seed, int eo_order);
void tmq_fieldgen_spinor_z4(double *out, const int localX[4], const int grid[4], const int coord[4],
                            unsigned long long seed, int eo_order);


/* ---- on-disk formats (host/tmq_lime.cpp): ILDG / LIME gauge configurations (include/QKXTM_read_conf.h:107-400) and the
 * "DiracFermion_Sink" propagator files of QKXTM_Vector::write (lib/qudaQKXTM_Vector.cpp:510-702).  Every function returns
 * 0 on success; tmq_lime_last_error() gives the text.  Each rank reads / writes its own sub-block with pread / pwrite
 * (no c-lime, no MPI-IO); when writing, the rank at coordinate (0,0,0,0) creates the file and must be called first.   */
const char *tmq_lime_last_error(void);
int tmq_lime_gauge_info(const char *fname, int globalX[4], int *precision_bits, double *kappa, double *mu);
/* gauge[mu]: QDP even-odd host order, 2*Vh*18 doubles; no boundary condition is applied (QKXTM_read_conf.h:395-397)      */
int tmq_lime_read_gauge(const char *fname, double *const gauge[4], const int localX[4], const int grid[4], const int coord[4]);
int tmq_lime_write_gauge(const char *fname, const double *const gauge[4], const int localX[4], const int grid[4],
                         const int coord[4], double kappa, double mu);
/* h_aos: the plug-in's host vector order [x_lex][spin][colour][re,im] (local sub-lattice), prec = 8 | 4                  */
int tmq_lime_write_vector(const char *fname, const void *h_aos, int prec, const int localX[4], const int grid[4], const int coord[4]);
int tmq_lime_read_vector(const char *fname, void *h_aos, int prec, const int localX[4], const int grid[4], const int coord[4]);
/* the two steps of tmq_lime_write_vector for a process grid: ONE rank creates the file (an existing file of that name is replaced
 * atomically), all ranks synchronise (comm_barrier), then EVERY rank writes its sub-block.  The reference orders the same two steps
 * with the MPI_Bcast of the payload offset (lib/qudaQKXTM_Vector.cpp:625).                                                     */
int tmq_lime_write_vector_header(const char *fname, int prec, const int localX[4], const int grid[4]);
int tmq_lime_write_vector_block(const char *fname, const void *h_aos, int prec, const int localX[4], const int grid[4], const int coord[4]);
/* applyGaugeFieldScaling restricted to what the path uses (qkxtm/QKXTM_util.cpp:682-725 with anisotropy 1): multiplies
 * U_t on the last GLOBAL time slice by t_boundary (-1: anti-periodic).  gauge[mu]: QDP even-odd host order.               */
void tmq_apply_t_boundary(double *const gauge[4], const int localX[4], const int grid[4], const int coord[4], int t_boundary);

/* ---- the noise vectors of calc_loops (host/qkxtm_noise.cpp; lib/qudaQKXTM_interface.cpp:1951,1982; lib/qudaQKXTM_utils.cpp:148-180,
 * 476-717).  The reference draws them from GSL's gsl_rng_ranlux (absent here): the generator is restated from the published RANLUX
 * algorithm (luxury 223) and reproduces GSL's known-answer test (seed 314159265 -> 10000th number 12077992), so that a noise vector
 * is bit-identical to the reference's for the same seed.                                                                         */
void *tmq_ranlux_alloc(unsigned long seed);                 /* gsl_rng_alloc(gsl_rng_ranlux) + gsl_rng_set                         */
void tmq_ranlux_free(void *rng);
unsigned long tmq_ranlux_get(void *rng);                    /* gsl_rng_get: uniform in [0, 2^24 - 1]                               */
unsigned long tmq_ranlux_uniform_int(void *rng, unsigned long n);   /* gsl_rng_uniform_int                                          */
void tmq_noise_z4(double *out, long long ncomplex, void *rng, int unity);   /* getStochasticRandomSource: +-1, +-i per component     */
/* hierarchical probing: colour of every site of the local lattice L[0..d-1] (x fastest), 2 * 2^{d(k-1)} colours; 0 = ok               */
int tmq_hch_coloring(unsigned short *Vc, const int *L, int k, int d);
int tmq_hadamard_element(int i, int j);                     /* HadamardElements: (-1)^{popcount(i & j)}                            */

#ifdef __cplusplus
}
#endif
#endif
