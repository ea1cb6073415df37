/*
 * quda_tmq.h -- the slice of upstream QUDA's C API (quda.h / enum_quda.h / quda_constants.h) that the QKXTM plug-in and its
 * drivers touch on the hot path, re-declared for the libtmq drop-in.
 *
 * Upstream QUDA is NOT vendored under the reference tree (CMakeLists.txt:18), so nothing here is a copy: the type, field and
 * enumerator NAMES are the ones the reference's own sources spell out -- qkxtm/Calc_Loops.cpp:187-225 (setGaugeParam),
 * :227-378 (setMultigridParam), :380-497 (setInvertParam), qkxtm/MG_Bench.cpp:190-445, lib/qudaQKXTM_interface.cpp:19-233,
 * 1409-2233 -- so that those parameter-setup blocks compile UNMODIFIED against this header (tests/test_dropin_compile.py
 * compiles them from where they lie).  Enumerator values follow upstream's numbering as far as it is documented by use
 * (precisions are byte sizes, reconstruct types are real counts, t_boundary is the sign, QUDA_INVALID_ENUM is INT_MIN); a
 * driver is recompiled against this header, it is not binary-compatible with a prebuilt libquda.
 *
 * Fields that the built path does not read (multigrid, Schwarz preconditioner, multi-shift, clover precision, ...) are
 * present so that drivers compile, and are ignored; see DESIGN.md "boundary".
 */
#ifndef QUDA_TMQ_H
#define QUDA_TMQ_H

#include <limits.h>
#include <stdio.h>
#include <stddef.h>

#define QUDA_VERSION_MAJOR 0
#define QUDA_VERSION_MINOR 9
#define QUDA_VERSION_SUBMINOR 0
#define QUDA_MAX_DIM 6              /* quda_constants.h */
#define QUDA_MAX_MULTI_SHIFT 32
#define QUDA_MAX_MG_LEVEL 4
#define QUDA_INVALID_ENUM INT_MIN

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enum_quda.h ------------------------------------------------------------------------------------------------------ */
typedef enum QudaPrecision_s { QUDA_HALF_PRECISION = 2, QUDA_SINGLE_PRECISION = 4, QUDA_DOUBLE_PRECISION = 8,
                               QUDA_INVALID_PRECISION = QUDA_INVALID_ENUM } QudaPrecision;
typedef enum QudaReconstructType_s { QUDA_RECONSTRUCT_NO = 18, QUDA_RECONSTRUCT_12 = 12, QUDA_RECONSTRUCT_8 = 8, QUDA_RECONSTRUCT_9 = 9,
                                     QUDA_RECONSTRUCT_13 = 13, QUDA_RECONSTRUCT_10 = 10, QUDA_RECONSTRUCT_INVALID = QUDA_INVALID_ENUM } QudaReconstructType;
typedef enum QudaTboundary_s { QUDA_ANTI_PERIODIC_T = -1, QUDA_PERIODIC_T = 1, QUDA_INVALID_T_BOUNDARY = QUDA_INVALID_ENUM } QudaTboundary;
typedef enum QudaGaugeFieldOrder_s { QUDA_FLOAT_GAUGE_ORDER = 1, QUDA_FLOAT2_GAUGE_ORDER = 2, QUDA_FLOAT4_GAUGE_ORDER = 4, QUDA_QDP_GAUGE_ORDER,
                                     QUDA_QDPJIT_GAUGE_ORDER, QUDA_CPS_WILSON_GAUGE_ORDER, QUDA_MILC_GAUGE_ORDER, QUDA_BQCD_GAUGE_ORDER,
                                     QUDA_TIFR_GAUGE_ORDER, QUDA_TIFR_PADDED_GAUGE_ORDER, QUDA_INVALID_GAUGE_ORDER = QUDA_INVALID_ENUM } QudaGaugeFieldOrder;
typedef enum QudaLinkType_s { QUDA_SU3_LINKS, QUDA_GENERAL_LINKS, QUDA_THREE_LINKS, QUDA_MOMENTUM_LINKS, QUDA_COARSE_LINKS, QUDA_SMEARED_LINKS,
                              QUDA_WILSON_LINKS = QUDA_SU3_LINKS, QUDA_ASQTAD_FAT_LINKS = QUDA_GENERAL_LINKS, QUDA_ASQTAD_LONG_LINKS = QUDA_THREE_LINKS,
                              QUDA_ASQTAD_MOM_LINKS = QUDA_MOMENTUM_LINKS, QUDA_ASQTAD_GENERAL_LINKS = QUDA_GENERAL_LINKS,
                              QUDA_INVALID_LINKS = QUDA_INVALID_ENUM } QudaLinkType;
typedef enum QudaGaugeFixed_s { QUDA_GAUGE_FIXED_NO, QUDA_GAUGE_FIXED_YES, QUDA_GAUGE_FIXED_INVALID = QUDA_INVALID_ENUM } QudaGaugeFixed;
typedef enum QudaDslashType_s { QUDA_WILSON_DSLASH, QUDA_CLOVER_WILSON_DSLASH, QUDA_DOMAIN_WALL_DSLASH, QUDA_DOMAIN_WALL_4D_DSLASH,
                                QUDA_MOBIUS_DWF_DSLASH, QUDA_STAGGERED_DSLASH, QUDA_ASQTAD_DSLASH, QUDA_TWISTED_MASS_DSLASH,
                                QUDA_TWISTED_CLOVER_DSLASH, QUDA_LAPLACE_DSLASH, QUDA_COVDEV_DSLASH, QUDA_INVALID_DSLASH = QUDA_INVALID_ENUM } QudaDslashType;
typedef enum QudaInverterType_s { QUDA_CG_INVERTER, QUDA_BICGSTAB_INVERTER, QUDA_GCR_INVERTER, QUDA_MR_INVERTER, QUDA_MPBICGSTAB_INVERTER,
                                  QUDA_SD_INVERTER, QUDA_XSD_INVERTER, QUDA_PCG_INVERTER, QUDA_MPCG_INVERTER, QUDA_EIGCG_INVERTER,
                                  QUDA_INC_EIGCG_INVERTER, QUDA_GMRESDR_INVERTER, QUDA_GMRESDR_PROJ_INVERTER, QUDA_GMRESDR_SH_INVERTER,
                                  QUDA_FGMRESDR_INVERTER, QUDA_MG_INVERTER, QUDA_BICGSTABL_INVERTER, QUDA_CGNE_INVERTER, QUDA_CGNR_INVERTER,
                                  QUDA_INVALID_INVERTER = QUDA_INVALID_ENUM } QudaInverterType;
typedef enum QudaSolutionType_s { QUDA_MAT_SOLUTION, QUDA_MATDAG_MAT_SOLUTION, QUDA_MATPC_SOLUTION, QUDA_MATPC_DAG_SOLUTION,
                                  QUDA_MATPCDAG_MATPC_SOLUTION, QUDA_MATPCDAG_MATPC_SHIFT_SOLUTION, QUDA_INVALID_SOLUTION = QUDA_INVALID_ENUM } QudaSolutionType;
typedef enum QudaSolveType_s { QUDA_DIRECT_SOLVE, QUDA_NORMOP_SOLVE, QUDA_DIRECT_PC_SOLVE, QUDA_NORMOP_PC_SOLVE, QUDA_NORMERR_SOLVE,
                               QUDA_NORMERR_PC_SOLVE, QUDA_NORMEQ_SOLVE = QUDA_NORMOP_SOLVE, QUDA_NORMEQ_PC_SOLVE = QUDA_NORMOP_PC_SOLVE,
                               QUDA_INVALID_SOLVE = QUDA_INVALID_ENUM } QudaSolveType;
typedef enum QudaMultigridCycleType_s { QUDA_MG_CYCLE_VCYCLE, QUDA_MG_CYCLE_FCYCLE, QUDA_MG_CYCLE_WCYCLE, QUDA_MG_CYCLE_RECURSIVE,
                                        QUDA_MG_CYCLE_INVALID = QUDA_INVALID_ENUM } QudaMultigridCycleType;
typedef enum QudaSchwarzType_s { QUDA_ADDITIVE_SCHWARZ, QUDA_MULTIPLICATIVE_SCHWARZ, QUDA_INVALID_SCHWARZ = QUDA_INVALID_ENUM } QudaSchwarzType;
typedef enum QudaResidualType_s { QUDA_L2_RELATIVE_RESIDUAL = 1, QUDA_L2_ABSOLUTE_RESIDUAL = 2, QUDA_HEAVY_QUARK_RESIDUAL = 4,
                                  QUDA_INVALID_RESIDUAL = QUDA_INVALID_ENUM } QudaResidualType;
/* values 0..3 are what include/tmq.h's TMQ_MATPC_* take */
typedef enum QudaMatPCType_s { QUDA_MATPC_EVEN_EVEN, QUDA_MATPC_ODD_ODD, QUDA_MATPC_EVEN_EVEN_ASYMMETRIC, QUDA_MATPC_ODD_ODD_ASYMMETRIC,
                               QUDA_MATPC_INVALID = QUDA_INVALID_ENUM } QudaMatPCType;
typedef enum QudaDagType_s { QUDA_DAG_NO, QUDA_DAG_YES, QUDA_DAG_INVALID = QUDA_INVALID_ENUM } QudaDagType;
typedef enum QudaMassNormalization_s { QUDA_KAPPA_NORMALIZATION, QUDA_MASS_NORMALIZATION, QUDA_ASYMMETRIC_MASS_NORMALIZATION,
                                       QUDA_INVALID_NORMALIZATION = QUDA_INVALID_ENUM } QudaMassNormalization;
typedef enum QudaSolverNormalization_s { QUDA_DEFAULT_NORMALIZATION, QUDA_SOURCE_NORMALIZATION } QudaSolverNormalization;
typedef enum QudaPreserveSource_s { QUDA_PRESERVE_SOURCE_NO, QUDA_PRESERVE_SOURCE_YES, QUDA_PRESERVE_SOURCE_INVALID = QUDA_INVALID_ENUM } QudaPreserveSource;
typedef enum QudaDiracFieldOrder_s { QUDA_INTERNAL_DIRAC_ORDER, QUDA_DIRAC_ORDER, QUDA_QDP_DIRAC_ORDER, QUDA_QDPJIT_DIRAC_ORDER,
                                     QUDA_CPS_WILSON_DIRAC_ORDER, QUDA_LEX_DIRAC_ORDER, QUDA_TIFR_PADDED_DIRAC_ORDER,
                                     QUDA_INVALID_DIRAC_ORDER = QUDA_INVALID_ENUM } QudaDiracFieldOrder;
typedef enum QudaCloverFieldOrder_s { QUDA_FLOAT_CLOVER_ORDER = 1, QUDA_FLOAT2_CLOVER_ORDER = 2, QUDA_FLOAT4_CLOVER_ORDER = 4, QUDA_PACKED_CLOVER_ORDER,
                                      QUDA_QDPJIT_CLOVER_ORDER, QUDA_BQCD_CLOVER_ORDER, QUDA_INVALID_CLOVER_ORDER = QUDA_INVALID_ENUM } QudaCloverFieldOrder;
typedef enum QudaVerbosity_s { QUDA_SILENT, QUDA_SUMMARIZE, QUDA_VERBOSE, QUDA_DEBUG_VERBOSE, QUDA_INVALID_VERBOSITY = QUDA_INVALID_ENUM } QudaVerbosity;
typedef enum QudaTune_s { QUDA_TUNE_NO, QUDA_TUNE_YES, QUDA_TUNE_INVALID = QUDA_INVALID_ENUM } QudaTune;
typedef enum QudaFieldLocation_s { QUDA_CPU_FIELD_LOCATION = 1, QUDA_CUDA_FIELD_LOCATION = 2, QUDA_INVALID_FIELD_LOCATION = QUDA_INVALID_ENUM } QudaFieldLocation;
typedef enum QudaSiteSubset_s { QUDA_PARITY_SITE_SUBSET = 1, QUDA_FULL_SITE_SUBSET = 2, QUDA_INVALID_SITE_SUBSET = QUDA_INVALID_ENUM } QudaSiteSubset;
typedef enum QudaGammaBasis_s { QUDA_DEGRAND_ROSSI_GAMMA_BASIS, QUDA_UKQCD_GAMMA_BASIS, QUDA_CHIRAL_GAMMA_BASIS,
                                QUDA_INVALID_GAMMA_BASIS = QUDA_INVALID_ENUM } QudaGammaBasis;
typedef enum QudaTwistFlavorType_s { QUDA_TWIST_SINGLET = 1, QUDA_TWIST_NONDEG_DOUBLET = +2, QUDA_TWIST_DEG_DOUBLET = -2, QUDA_TWIST_NO = 0,
                                     QUDA_TWIST_MINUS = -1, QUDA_TWIST_PLUS = +1, QUDA_TWIST_INVALID = QUDA_INVALID_ENUM } QudaTwistFlavorType;
typedef enum QudaUseInitGuess_s { QUDA_USE_INIT_GUESS_NO, QUDA_USE_INIT_GUESS_YES, QUDA_USE_INIT_GUESS_INVALID = QUDA_INVALID_ENUM } QudaUseInitGuess;
typedef enum QudaComputeNullVector_s { QUDA_COMPUTE_NULL_VECTOR_NO, QUDA_COMPUTE_NULL_VECTOR_YES,
                                       QUDA_COMPUTE_NULL_VECTOR_INVALID = QUDA_INVALID_ENUM } QudaComputeNullVector;
typedef enum QudaSetupType_s { QUDA_NULL_VECTOR_SETUP, QUDA_TEST_VECTOR_SETUP, QUDA_INVALID_SETUP_TYPE = QUDA_INVALID_ENUM } QudaSetupType;
typedef enum QudaBoolean_s { QUDA_BOOLEAN_NO = 0, QUDA_BOOLEAN_YES = 1, QUDA_BOOLEAN_INVALID = QUDA_INVALID_ENUM } QudaBoolean;

/* ---- quda.h: QudaGaugeParam (fields set at qkxtm/Calc_Loops.cpp:187-225) ---------------------------------------------------- */
typedef struct QudaGaugeParam_s {
  QudaFieldLocation location;
  int X[4];                     /* LOCAL lattice extents */
  double anisotropy;            /* must be 1 on this path */
  double tadpole_coeff;
  double scale;
  QudaLinkType type;            /* QUDA_WILSON_LINKS feeds the solver; QUDA_SMEARED_LINKS is accepted and ignored */
  QudaGaugeFieldOrder gauge_order;   /* QUDA_QDP_GAUGE_ORDER only */
  QudaTboundary t_boundary;
  QudaPrecision cpu_prec;       /* double only */
  QudaPrecision cuda_prec;
  QudaReconstructType reconstruct;   /* 18 / 12 / 8: storage of the resident links (both precisions) */
  QudaPrecision cuda_prec_sloppy;
  QudaReconstructType reconstruct_sloppy;
  QudaPrecision cuda_prec_precondition;
  QudaReconstructType reconstruct_precondition;
  QudaGaugeFixed gauge_fix;
  int ga_pad;                   /* ignored: faces are exchanged spin-projected, there is no padded gauge ghost */
  int site_ga_pad, staple_pad, llfat_ga_pad, mom_ga_pad;
  double gaugeGiB;
  int overlap, overwrite_mom, use_resident_gauge, use_resident_mom, make_resident_gauge, make_resident_mom, return_result_gauge, return_result_mom;
} QudaGaugeParam;

/* ---- quda.h: QudaInvertParam (fields set at qkxtm/Calc_Loops.cpp:380-497; read at lib/qudaQKXTM_interface.cpp:19-233,1409-2233) -- */
typedef struct QudaInvertParam_s {
  QudaFieldLocation input_location, output_location;
  QudaDslashType dslash_type;   /* twisted-mass or twisted-clover */
  QudaInverterType inv_type;    /* QUDA_CG_INVERTER */
  double mass, kappa;
  double m5;
  int Ls;
  double mu;                    /* sign carries the flavour (lib/qudaQKXTM_interface.cpp:504,515) */
  double epsilon;
  QudaTwistFlavorType twist_flavor;
  double tol;                   /* |r| / |b| */
  double tol_restart;
  double tol_hq;
  int compute_true_res;
  double true_res;              /* OUT */
  double true_res_hq;           /* OUT (not computed: 0) */
  int maxiter;
  double reliable_delta;        /* mixed precision: residual drop that triggers an fp64 update (drivers: 1e-4) */
  int use_sloppy_partial_accumulator, max_res_increase, max_res_increase_total, heavy_quark_check;
  int pipeline;
  int num_offset;
  int num_src;
  int overlap;
  double offset[QUDA_MAX_MULTI_SHIFT];
  double tol_offset[QUDA_MAX_MULTI_SHIFT];
  double tol_hq_offset[QUDA_MAX_MULTI_SHIFT];
  double true_res_offset[QUDA_MAX_MULTI_SHIFT];
  double iter_res_offset[QUDA_MAX_MULTI_SHIFT];
  double true_res_hq_offset[QUDA_MAX_MULTI_SHIFT];
  double residue[QUDA_MAX_MULTI_SHIFT];
  int compute_action;
  double action[2];
  QudaSolutionType solution_type;    /* QUDA_MAT_SOLUTION */
  QudaSolveType solve_type;          /* QUDA_NORMOP_PC_SOLVE */
  QudaMatPCType matpc_type;
  QudaDagType dagger;
  QudaMassNormalization mass_normalization;
  QudaSolverNormalization solver_normalization;
  QudaPreserveSource preserve_source;
  QudaPrecision cpu_prec;            /* double */
  QudaPrecision cuda_prec;           /* double (the QKXTM upload kernel writes double2, lib/qudaQKXTM_kernels.cu:1031) */
  QudaPrecision cuda_prec_sloppy;    /* double: pure fp64 CG; single: fp32 inner iterations + reliable updates */
  QudaPrecision cuda_prec_precondition;
  QudaDiracFieldOrder dirac_order;   /* QUDA_DIRAC_ORDER (colour inside spin) */
  QudaGammaBasis gamma_basis;        /* QUDA_UKQCD_GAMMA_BASIS */
  QudaFieldLocation clover_location;
  QudaPrecision clover_cpu_prec, clover_cuda_prec, clover_cuda_prec_sloppy, clover_cuda_prec_precondition;
  QudaCloverFieldOrder clover_order;
  QudaUseInitGuess use_init_guess;
  double clover_coeff;               /* csw * kappa (qkxtm/MG_Bench.cpp:249), used by loadCloverQuda */
  int compute_clover_trlog;
  double trlogA[2];
  int compute_clover, compute_clover_inverse, return_clover, return_clover_inverse;
  QudaVerbosity verbosity;
  int sp_pad, cl_pad;                /* must be 0 (lib/code_pieces/uploadToCuda_core.h:5) */
  int iter;                          /* OUT */
  double gflops;                     /* OUT */
  double secs;                       /* OUT */
  QudaTune tune;
  int Nsteps;
  int gcrNkrylov;
  QudaInverterType inv_type_precondition;
  void *preconditioner;              /* multigrid handle: accepted, unused (CG on the normal operator is the built solver) */
  void *preconditionerUP, *preconditionerDN;   /* the plug-in's own patch to quda.h (README:84-108) */
  void *deflation_op;
  QudaDslashType dslash_type_precondition;
  QudaVerbosity verbosity_precondition;
  double tol_precondition;
  int maxiter_precondition;
  double omega;
  QudaSchwarzType schwarz_type;
  int precondition_cycle;
  QudaResidualType residual_type;
  double spinorGiB;                  /* OUT (lib/qudaQKXTM_interface.cpp:84-94) */
  double cloverGiB;
  double gaugeGiB;
} QudaInvertParam;

/* ---- quda.h: QudaMultigridParam (set at qkxtm/Calc_Loops.cpp:227-378).  The multigrid solver is NOT built (out of
 *      scope); the struct exists so that the drivers' setMultigridParam compiles, and newMultigridQuda refuses at run time. -------- */
typedef struct QudaMultigridParam_s {
  QudaInvertParam *invert_param;
  int n_level;
  int geo_block_size[QUDA_MAX_MG_LEVEL][QUDA_MAX_DIM];
  int spin_block_size[QUDA_MAX_MG_LEVEL];
  int n_vec[QUDA_MAX_MG_LEVEL];
  QudaPrecision precision_null[QUDA_MAX_MG_LEVEL];
  QudaVerbosity verbosity[QUDA_MAX_MG_LEVEL];
  QudaInverterType setup_inv_type[QUDA_MAX_MG_LEVEL];
  int num_setup_iter[QUDA_MAX_MG_LEVEL];
  double setup_tol[QUDA_MAX_MG_LEVEL];
  QudaSetupType setup_type;
  QudaBoolean pre_orthonormalize, post_orthonormalize;
  QudaInverterType coarse_solver[QUDA_MAX_MG_LEVEL];
  double coarse_solver_tol[QUDA_MAX_MG_LEVEL];
  int coarse_solver_maxiter[QUDA_MAX_MG_LEVEL];
  QudaInverterType smoother[QUDA_MAX_MG_LEVEL];
  double smoother_tol[QUDA_MAX_MG_LEVEL];
  int nu_pre[QUDA_MAX_MG_LEVEL], nu_post[QUDA_MAX_MG_LEVEL];
  double omega[QUDA_MAX_MG_LEVEL];
  QudaSchwarzType smoother_schwarz_type[QUDA_MAX_MG_LEVEL];
  int smoother_schwarz_cycle[QUDA_MAX_MG_LEVEL];
  QudaSolutionType coarse_grid_solution_type[QUDA_MAX_MG_LEVEL];
  QudaSolveType smoother_solve_type[QUDA_MAX_MG_LEVEL];
  QudaMultigridCycleType cycle_type[QUDA_MAX_MG_LEVEL];
  QudaBoolean global_reduction[QUDA_MAX_MG_LEVEL];
  QudaFieldLocation location[QUDA_MAX_MG_LEVEL];
  QudaComputeNullVector compute_null_vector;
  QudaBoolean generate_all_levels;
  QudaBoolean run_verify;
  double mu_factor[QUDA_MAX_MG_LEVEL];
  double gflops, secs;
  char vec_infile[256], vec_outfile[256];
} QudaMultigridParam;

typedef int (*QudaCommsMap)(const int *coords, void *fdata);

/* ---- the C API entry points the drivers call (qkxtm/Calc_Loops.cpp:692-708,753-759,797-806; qkxtm/QKXTM_util.cpp:48-68) ---- */
QudaGaugeParam newQudaGaugeParam(void);
QudaInvertParam newQudaInvertParam(void);
QudaMultigridParam newQudaMultigridParam(void);
/* dims = the process grid (x, y, z, t); only z and t may exceed 1.  One process per rank; rank and world size come from the
 * launcher's environment (RANK / WORLD_SIZE of torchrun --no-python, OMPI_COMM_WORLD_*, PMI_*, SLURM_*), the rank <-> coordinate map
 * has t fastest (func / fdata: a custom map is not supported and must be NULL).  The NCCL communicator is created when the first
 * field is (loadGaugeQuda / init_qudaQKXTM): rank 0 passes the id to the others through a file, see TMQ_COMM_ID_FILE in INTEGRATION.md. */
void initCommsGridQuda(int nDim, const int *dims, QudaCommsMap func, void *fdata);
int comm_rank(void);
int comm_size(void);
int comm_coord(int dim);
int comm_dim(int dim);
int comm_dim_partitioned(int dim);
void comm_barrier(void);                                                       /* a real rendezvous of all ranks (NCCL all-reduce + host wait) */
void initQuda(int device);                                                     /* device < 0: the launcher's local rank */
void loadGaugeQuda(void *h_gauge, QudaGaugeParam *param);                      /* void *gauge[4], QDP even-odd order, double */
void freeGaugeQuda(void);
/* loadCloverQuda(NULL, NULL, &inv_param) (qkxtm/MG_Bench.cpp:605-608): the clover field is BUILT on the device from the resident
 * gauge field with inv_param->clover_coeff; host clover fields (h_clover / h_clovinv != NULL) are not supported */
void loadCloverQuda(void *h_clover, void *h_clovinv, QudaInvertParam *inv_param);
void freeCloverQuda(void);
void endQuda(void);
/* host spinors: full lattice, even-odd site order [even Vh | odd Vh][spin][colour][re,im], double */
void invertQuda(void *h_x, void *h_b, QudaInvertParam *param);
/* param->num_src right-hand sides; uploads / downloads of neighbouring columns run behind each solve */
void invertMultiSrcQuda(void **hp_x, void **hp_b, QudaInvertParam *param);
void MatQuda(void *h_out, void *h_in, QudaInvertParam *param);                 /* full operator */
void setVerbosityQuda(QudaVerbosity verbosity, const char prefix[], FILE *outfile);
/* multigrid is out of scope: newMultigridQuda aborts through errorQuda, destroyMultigridQuda(NULL) is a no-op */
void *newMultigridQuda(QudaMultigridParam *param);
void destroyMultigridQuda(void *mg_instance);

#ifdef __cplusplus
}
/* one-argument form used by this repository's own drivers and tests */
inline void setVerbosityQuda(QudaVerbosity verbosity) { setVerbosityQuda(verbosity, "", (FILE *)0); }
#endif

#endif
