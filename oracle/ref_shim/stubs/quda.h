// TEST INFRASTRUCTURE: a stand-in for the handful of upstream-QUDA declarations that the reference's own host utility file
// qkxtm/QKXTM_util.cpp (a derivative of upstream tests/test_util.cpp) needs in order to COMPILE here.  Nothing in this
// directory is reference code: these are declarations written for this repository so that oracle/ref_shim/qkxtm_util_host.cpp
// can #include the reference file from where it lies and call its geometry / gauge helpers.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <sys/time.h>
typedef enum { QUDA_HALF_PRECISION=2, QUDA_SINGLE_PRECISION = 4, QUDA_DOUBLE_PRECISION = 8, QUDA_INVALID_PRECISION=-1 } QudaPrecision;
typedef enum { QUDA_RECONSTRUCT_NO = 18, QUDA_RECONSTRUCT_12 = 12, QUDA_RECONSTRUCT_8 = 8, QUDA_RECONSTRUCT_9 = 9, QUDA_RECONSTRUCT_13 = 13, QUDA_RECONSTRUCT_INVALID=-1 } QudaReconstructType;
typedef enum { QUDA_ANTI_PERIODIC_T = -1, QUDA_PERIODIC_T = 1 } QudaTboundary;
typedef enum { QUDA_QDP_GAUGE_ORDER = 0, QUDA_MILC_GAUGE_ORDER, QUDA_CPS_WILSON_GAUGE_ORDER, QUDA_BQCD_GAUGE_ORDER, QUDA_TIFR_GAUGE_ORDER, QUDA_TIFR_PADDED_GAUGE_ORDER } QudaGaugeFieldOrder;
typedef enum { QUDA_WILSON_LINKS = 0, QUDA_SMEARED_LINKS, QUDA_ASQTAD_FAT_LINKS, QUDA_ASQTAD_LONG_LINKS, QUDA_ASQTAD_MOM_LINKS, QUDA_ASQTAD_GENERAL_LINKS, QUDA_SU3_LINKS, QUDA_GENERAL_LINKS, QUDA_THREE_LINKS, QUDA_MOMENTUM, QUDA_COARSE_LINKS } QudaLinkType;
typedef enum { QUDA_GAUGE_FIXED_NO = 0, QUDA_GAUGE_FIXED_YES } QudaGaugeFixed;
typedef enum { QUDA_WILSON_DSLASH, QUDA_CLOVER_WILSON_DSLASH, QUDA_DOMAIN_WALL_DSLASH, QUDA_DOMAIN_WALL_4D_DSLASH, QUDA_MOBIUS_DWF_DSLASH, QUDA_STAGGERED_DSLASH, QUDA_ASQTAD_DSLASH, QUDA_TWISTED_MASS_DSLASH, QUDA_TWISTED_CLOVER_DSLASH, QUDA_LAPLACE_DSLASH, QUDA_COVDEV_DSLASH, QUDA_INVALID_DSLASH=-1 } QudaDslashType;
typedef enum { QUDA_DIRAC_ORDER = 0, QUDA_QDP_DIRAC_ORDER, QUDA_QDPJIT_DIRAC_ORDER, QUDA_CPS_WILSON_DIRAC_ORDER, QUDA_LEX_DIRAC_ORDER, QUDA_TIFR_PADDED_DIRAC_ORDER } QudaDiracFieldOrder;
#define QUDA_MAX_DIM 6
#define QUDA_MAX_MG_LEVEL 4
typedef struct QudaGaugeParam_s { int X[4]; double anisotropy; double tadpole_coeff; double scale; QudaLinkType type; QudaGaugeFieldOrder gauge_order; QudaTboundary t_boundary; QudaPrecision cpu_prec, cuda_prec, cuda_prec_sloppy, cuda_prec_precondition; QudaReconstructType reconstruct, reconstruct_sloppy, reconstruct_precondition; QudaGaugeFixed gauge_fix; int ga_pad; } QudaGaugeParam;
typedef enum { QUDA_CG_INVERTER, QUDA_BICGSTAB_INVERTER, QUDA_GCR_INVERTER, QUDA_MR_INVERTER, QUDA_MPBICGSTAB_INVERTER, QUDA_SD_INVERTER, QUDA_XSD_INVERTER, QUDA_PCG_INVERTER, QUDA_MPCG_INVERTER, QUDA_EIGCG_INVERTER, QUDA_INC_EIGCG_INVERTER, QUDA_GMRESDR_INVERTER, QUDA_GMRESDR_PROJ_INVERTER, QUDA_GMRESDR_SH_INVERTER, QUDA_FGMRESDR_INVERTER, QUDA_MG_INVERTER, QUDA_BICGSTABL_INVERTER, QUDA_INVALID_INVERTER = -1 } QudaInverterType;
typedef enum { QUDA_DIRECT_SOLVE, QUDA_NORMOP_SOLVE, QUDA_DIRECT_PC_SOLVE, QUDA_NORMOP_PC_SOLVE, QUDA_NORMERR_SOLVE, QUDA_NORMERR_PC_SOLVE, QUDA_INVALID_SOLVE = -1 } QudaSolveType;
typedef enum { QUDA_MATPC_EVEN_EVEN, QUDA_MATPC_ODD_ODD, QUDA_MATPC_EVEN_EVEN_ASYMMETRIC, QUDA_MATPC_ODD_ODD_ASYMMETRIC, QUDA_MATPC_INVALID = -1 } QudaMatPCType;
typedef enum { QUDA_KAPPA_NORMALIZATION, QUDA_MASS_NORMALIZATION, QUDA_ASYMMETRIC_MASS_NORMALIZATION, QUDA_INVALID_NORMALIZATION = -1 } QudaMassNormalization;
typedef enum { QUDA_TWIST_SINGLET = 1, QUDA_TWIST_NONDEG_DOUBLET = +2, QUDA_TWIST_DEG_DOUBLET = -2, QUDA_TWIST_NO = 0, QUDA_TWIST_MINUS = -1, QUDA_TWIST_PLUS = +1, QUDA_TWIST_INVALID = -100 } QudaTwistFlavorType;
typedef enum { QUDA_SILENT, QUDA_SUMMARIZE, QUDA_VERBOSE, QUDA_DEBUG_VERBOSE, QUDA_INVALID_VERBOSITY = -1 } QudaVerbosity;
typedef enum { QUDA_ADDITIVE_SCHWARZ, QUDA_MULTIPLICATIVE_SCHWARZ, QUDA_INVALID_SCHWARZ = -1 } QudaSchwarzType;
typedef enum { QUDA_NULL_VECTOR_SETUP, QUDA_TEST_VECTOR_SETUP, QUDA_INVALID_SETUP_TYPE = -1 } QudaSetupType;
typedef enum { QUDA_DAG_NO, QUDA_DAG_YES, QUDA_DAG_INVALID = -1 } QudaDagType;
typedef struct QudaInvertParam_s { double kappa, mu; QudaDiracFieldOrder dirac_order; QudaPrecision cpu_prec; int Ls; } QudaInvertParam;
typedef int (*QudaCommsMap)(const int *coords, void *fdata);
void initCommsGridQuda(int nDim, const int *dims, QudaCommsMap func, void *fdata);
struct Topology;
extern Topology *default_topo;
const int *comm_coords(const Topology *topo);
int comm_dim_partitioned(int dim);
void comm_dim_partitioned_set(int dim);
int comm_rank(void);
int comm_size(void);
int comm_dim(int dim);
int comm_coord(int dim);
#define errorQuda(...) do { fprintf(stderr, __VA_ARGS__); abort(); } while (0)
#define printfQuda(...) printf(__VA_ARGS__)
#define warningQuda(...) printf(__VA_ARGS__)
