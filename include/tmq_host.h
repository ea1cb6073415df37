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

#ifdef __cplusplus
}
#endif
#endif
