// qudaQKXTM_tmq.h -- the C++ host side of the drop-in: the QKXTM containers, parameter structs and solve entry
// points of ETMC-QUDA/quda-QKXTM-Multigrid-PlugIn for ONE hot path (even-odd twisted-mass Dslash + CG on
// M^dag M), re-implemented over the libtmq.so C ABI (include/tmq.h).  Same names, argument meaning and
// error behaviour (errorQuda aborts) as the reference, so that a driver written against
//   include/qudaQKXTM.h:104-277 (containers), :484-513 (entry points),
//   include/qudaQKXTM_utils.h:45-75,126-139 (qudaQKXTMinfo, enums, init_qudaQKXTM)
// and the handful of QUDA C-API calls the drivers make (qkxtm/Calc_Loops.cpp:692-708,753-759,797-806)
// compiles against this header for that path.  Everything the path does not touch (contractions,
// loops, file I/O, ghost exchange of the containers) is deliberately absent -- see DESIGN.md.
//
// Threading / state: like the reference, one host thread per rank and library-global state (one context,
// one resident gauge field, one-shot init_qudaQKXTM); not re-entrant.  Multi-GPU: one process per rank, see initCommsGridQuda.
#ifndef QUDAQKXTM_TMQ_H
#define QUDAQKXTM_TMQ_H

#include <cstddef>
#include <typeinfo>
#include "tmq.h"

#define QUDAQKXTM_DIM 4

// ---- the slice of quda.h / enum_quda.h the path reads ----------------------------------------------
typedef enum { QUDA_SINGLE_PRECISION = 4, QUDA_DOUBLE_PRECISION = 8 } QudaPrecision;
typedef enum { QUDA_RECONSTRUCT_NO = 18, QUDA_RECONSTRUCT_12 = 12, QUDA_RECONSTRUCT_8 = 8 } QudaReconstructType;
typedef enum { QUDA_ANTI_PERIODIC_T = -1, QUDA_PERIODIC_T = 1 } QudaTboundary;
typedef enum { QUDA_QDP_GAUGE_ORDER = 0 } QudaGaugeFieldOrder;
typedef enum { QUDA_WILSON_LINKS = 0, QUDA_SMEARED_LINKS = 1 } QudaLinkType;
typedef enum { QUDA_GAUGE_FIXED_NO = 0 } QudaGaugeFixed;
typedef enum { QUDA_TWISTED_MASS_DSLASH = 0, QUDA_WILSON_DSLASH = 1, QUDA_TWISTED_CLOVER_DSLASH = 2 } QudaDslashType;
typedef enum { QUDA_CG_INVERTER = 0, QUDA_GCR_INVERTER = 1, QUDA_BICGSTAB_INVERTER = 2, QUDA_INVALID_INVERTER = -1 } QudaInverterType;
typedef enum { QUDA_MAT_SOLUTION = 0, QUDA_MATPC_SOLUTION = 1, QUDA_MATPCDAG_MATPC_SOLUTION = 2 } QudaSolutionType;
typedef enum { QUDA_DIRECT_SOLVE = 0, QUDA_NORMOP_SOLVE = 1, QUDA_DIRECT_PC_SOLVE = 2, QUDA_NORMOP_PC_SOLVE = 3 } QudaSolveType;
typedef enum { QUDA_MATPC_EVEN_EVEN = 0, QUDA_MATPC_ODD_ODD = 1, QUDA_MATPC_EVEN_EVEN_ASYMMETRIC = 2,
               QUDA_MATPC_ODD_ODD_ASYMMETRIC = 3 } QudaMatPCType;
typedef enum { QUDA_KAPPA_NORMALIZATION = 0, QUDA_MASS_NORMALIZATION = 1, QUDA_ASYMMETRIC_MASS_NORMALIZATION = 2 } QudaMassNormalization;
typedef enum { QUDA_DEGRAND_ROSSI_GAMMA_BASIS = 0, QUDA_UKQCD_GAMMA_BASIS = 1 } QudaGammaBasis;
typedef enum { QUDA_DIRAC_ORDER = 0 } QudaDiracFieldOrder;
typedef enum { QUDA_TWIST_SINGLET = 1, QUDA_TWIST_NO = 0 } QudaTwistFlavorType;
typedef enum { QUDA_DAG_NO = 0, QUDA_DAG_YES = 1 } QudaDagType;
typedef enum { QUDA_PRESERVE_SOURCE_NO = 0, QUDA_PRESERVE_SOURCE_YES = 1 } QudaPreserveSource;
typedef enum { QUDA_CPU_FIELD_LOCATION = 1, QUDA_CUDA_FIELD_LOCATION = 2 } QudaFieldLocation;
typedef enum { QUDA_L2_RELATIVE_RESIDUAL = 1 } QudaResidualType;
typedef enum { QUDA_SILENT = 0, QUDA_SUMMARIZE = 1, QUDA_VERBOSE = 2 } QudaVerbosity;
typedef enum { QUDA_PARITY_SITE_SUBSET = 1, QUDA_FULL_SITE_SUBSET = 2 } QudaSiteSubset;

typedef struct QudaGaugeParam_s {       // fields set at qkxtm/Calc_Loops.cpp:189-225
  int X[4];
  double anisotropy;
  QudaLinkType type;
  QudaGaugeFieldOrder gauge_order;
  QudaTboundary t_boundary;
  QudaPrecision cpu_prec, cuda_prec, cuda_prec_sloppy, cuda_prec_precondition;
  QudaReconstructType reconstruct, reconstruct_sloppy, reconstruct_precondition;
  QudaGaugeFixed gauge_fix;
  int ga_pad;
} QudaGaugeParam;

typedef struct QudaInvertParam_s {      // fields read / written on the path
  double kappa, mu, mass;
  double clover_coeff;                 // csw * kappa (qkxtm/MG_Bench.cpp:249), used by loadCloverQuda
  QudaDslashType dslash_type;
  QudaTwistFlavorType twist_flavor;
  QudaMatPCType matpc_type;
  QudaSolveType solve_type;
  QudaSolutionType solution_type;
  QudaInverterType inv_type, inv_type_precondition;
  QudaMassNormalization mass_normalization;
  QudaGammaBasis gamma_basis;
  QudaDiracFieldOrder dirac_order;
  QudaPrecision cpu_prec, cuda_prec, cuda_prec_sloppy, cuda_prec_precondition;
  QudaPreserveSource preserve_source;
  QudaFieldLocation input_location, output_location;
  int sp_pad, cl_pad, Ls;
  QudaDagType dagger;
  double tol, tol_hq;
  QudaResidualType residual_type;
  int maxiter;
  double reliable_delta;
  int pipeline, gcrNkrylov;
  QudaVerbosity verbosity, verbosity_precondition;
  void *preconditioner;
  // outputs (zeroed before each solve, lib/qudaQKXTM_interface.cpp:95-97; filled like updateInvertParam)
  double spinorGiB, secs, gflops, true_res;
  int iter;
} QudaInvertParam;

QudaGaugeParam newQudaGaugeParam(void);
QudaInvertParam newQudaInvertParam(void);
// qkxtm/QKXTM_util.cpp:48-68.  dims = the process grid (x, y, z, t), only z and t may exceed 1.  One process per rank; rank and world
// size come from the launcher's environment (RANK / WORLD_SIZE of torchrun --no-python, OMPI_COMM_WORLD_*, PMI_*, SLURM_*), the rank <->
// coordinate map has t fastest.  The NCCL communicator is created when the first field is (loadGaugeQuda / init_qudaQKXTM): rank 0
// passes the id to the others through a file (TMQ_COMM_ID_FILE, default /tmp/tmq_nccl_id_<parent pid>_<MASTER_PORT>).
void initCommsGridQuda(int nDim, const int *dims, void *func, void *fdata);
int comm_rank(void);
int comm_size(void);
int comm_coord(int dim);
void initQuda(int device);                                                     // qkxtm/Calc_Loops.cpp:753 (device < 0: the launcher's local rank)
void loadGaugeQuda(void *h_gauge, QudaGaugeParam *param);                      // qkxtm/Calc_Loops.cpp:759 (void *gauge[4], QDP order)
void freeGaugeQuda(void);
// loadCloverQuda(NULL, NULL, &inv_param) (qkxtm/MG_Bench.cpp:605-608): the clover field is BUILT on the device from the resident
// gauge field with inv_param->clover_coeff; host clover fields (h_clover / h_clovinv != NULL) are not supported
void loadCloverQuda(void *h_clover, void *h_clovinv, QudaInvertParam *inv_param);
void freeCloverQuda(void);
void endQuda(void);
// host spinors: full lattice, even-odd site order [even Vh | odd Vh][spin][colour][re,im], double
void invertQuda(void *h_x, void *h_b, QudaInvertParam *param);
void MatQuda(void *h_out, void *h_in, QudaInvertParam *param);                 // full operator (dslash_test-style check)
void setVerbosityQuda(QudaVerbosity v);

namespace quda {

enum ALLOCATION_FLAG { NONE, HOST, DEVICE, BOTH, BOTH_EXTRA };                 // include/qudaQKXTM_utils.h:126
enum CLASS_ENUM { FIELD, GAUGE, VECTOR, PROPAGATOR, PROPAGATOR3D, VECTOR3D };  // include/qudaQKXTM_utils.h:127

#define MAX_NSOURCES 1000                                                     // include/qudaQKXTM_utils.h:16
#define MAX_NMOMENTA 5000                                                     // include/qudaQKXTM_utils.h:19
enum CORR_SPACE { POSITION_SPACE, MOMENTUM_SPACE };                            // include/qudaQKXTM_utils.h:41
enum FILE_WRITE_FORMAT { ASCII_FORM, HDF5_FORM };                              // include/qudaQKXTM_utils.h:42
enum WHICHPARTICLE { PROTON, NEUTRON };                                        // include/qudaQKXTM_utils.h:128
enum WHICHPROJECTOR { G4, G5G123, G5G1, G5G2, G5G3 };                          // include/qudaQKXTM_utils.h:129
#define MAX_TSINK 10                                                          // include/qudaQKXTM_utils.h:20
#define MAX_PROJS 5                                                           // include/qudaQKXTM_utils.h:23

typedef struct {                       // the slice of qudaQKXTMinfo the built paths read (include/qudaQKXTM_utils.h:45-75)
  int nsmearAPE, nsmearGauss;
  double alphaAPE, alphaGauss;
  int lL[QUDAQKXTM_DIM];
  int Nsources;
  int sourcePosition[MAX_NSOURCES][QUDAQKXTM_DIM];    // global (x, y, z, t)
  QudaPrecision Precision;
  int Q_sq;                            // momenta with p^2 <= Q_sq (createMomenta, lib/qudaQKXTM_kernels.cu:98-116)
  int traj;
  bool check_files;
  int Ntsink;                          // sink-source separations of the three-point function
  int Nproj[MAX_TSINK];
  int tsinkSource[MAX_TSINK];
  int proj_list[MAX_TSINK][MAX_PROJS]; // WHICHPROJECTOR values
  int run3pt_src[MAX_NSOURCES];        // != 0: also the fixed-sink three-point function (ultra-local insertion) for this source
  FILE_WRITE_FORMAT CorrFileFormat;    // ASCII_FORM only (no HDF5 here)
  CORR_SPACE CorrSpace;
  bool isEven;
  double kappa, mu, csw, inv_tol;
} qudaQKXTMinfo;

void init_qudaQKXTM(qudaQKXTMinfo *info);      // lib/qudaQKXTM_kernels.cu:118-297 (one-shot; containers need it)
void printf_qudaQKXTM();
// include/qudaQKXTM_utils.h:36-37, lib/qudaQKXTM_utils.cpp:87-141 (gauge: 4 lexicographic link arrays, as packGauge takes)
void testPlaquette(void **gauge);
void testGaussSmearing(void **gauge);

// stand-in for cudaColorSpinorField on this path: a device field in libtmq's native layout
class ColorSpinorField {
  tmq_spinor *h_;
  QudaSiteSubset subset_;
public:
  ColorSpinorField(QudaSiteSubset subset, QudaPrecision prec);   // zero-initialised (QUDA_ZERO_FIELD_CREATE)
  ~ColorSpinorField();
  tmq_spinor *handle() const { return h_; }
  QudaSiteSubset SiteSubset() const { return subset_; }
  tmq_spinor *Even() const;
  tmq_spinor *Odd() const;
};

template <typename Float> class QKXTM_Vector;
template <typename Float> class QKXTM_Propagator;
template <typename Float> class QKXTM_Propagator3D;

template <typename Float> class QKXTM_Field {     // include/qudaQKXTM.h:104-160, lib/qudaQKXTM_Field.cpp:81-258
protected:
  int field_length;
  long long total_length;                          // 64-bit: the reference's int overflows at 64^3x128
  size_t bytes_total_length;
  Float *h_elem, *h_elem_backup, *d_elem;
  bool isAllocHost, isAllocDevice, isAllocHostBackup;
  void create_host();
  void create_host_backup();
  void destroy_host();
  void destroy_host_backup();
  void create_device();
  void destroy_device();
public:
  QKXTM_Field(ALLOCATION_FLAG alloc_flag, CLASS_ENUM classT);
  virtual ~QKXTM_Field();
  void zero_host();
  void zero_host_backup();
  void zero_device();
  Float *H_elem() const { return h_elem; }
  Float *D_elem() const { return d_elem; }
  size_t Bytes_total() const { return bytes_total_length; }
  size_t Bytes_ghost() const { return 0; }         // the containers' own ghost exchange is not on this path
  size_t Bytes_total_plus_ghost() const { return bytes_total_length; }
  int Precision() const { return (int)sizeof(Float); }
  void printInfo();
};

template <typename Float> class QKXTM_Gauge : public QKXTM_Field<Float> {       // include/qudaQKXTM.h:166-183
public:
  QKXTM_Gauge(ALLOCATION_FLAG alloc_flag, CLASS_ENUM classT);
  void packGauge(void **gauge);                    // lexicographic host links -> SoA (lib/qudaQKXTM_Gauge.cpp:73-89)
  void packGaugeToBackup(void **gauge);
  void loadGaugeFromBackup();
  void justDownloadGauge();
  void loadGauge();
  double calculatePlaq();                          // prints like the reference and also returns the value (skipped, 0, on a split lattice)
};

template <typename Float> class QKXTM_Vector : public QKXTM_Field<Float> {      // include/qudaQKXTM.h:189-225
public:
  QKXTM_Vector(ALLOCATION_FLAG alloc_flag, CLASS_ENUM classT);
  void packVector(Float *vector);                  // AoS [x][s][c][ri] -> SoA (lib/qudaQKXTM_Vector.cpp:72-81)
  void unpackVector();
  void unpackVector(Float *vector);
  void loadVector();
  void unloadVector();
  void download();                                 // D2H + SoA -> AoS in h_elem (lib/qudaQKXTM_Vector.cpp:135-156)
  void uploadToCuda(ColorSpinorField *cudaVector, bool isEv = false);      // lib/qudaQKXTM_Vector.cpp:424-427
  void downloadFromCuda(ColorSpinorField *cudaVector, bool isEv = false);  // lib/qudaQKXTM_Vector.cpp:430-432
  void gaussianSmearing(QKXTM_Vector<Float> &vecIn, QKXTM_Gauge<Float> &gaugeAPE);   // lib/qudaQKXTM_Vector.cpp:386-421 (vecIn is clobbered, as in the reference)
  void scaleVector(double a);
  void conjugate();                                // lib/code_pieces/conjugate_vector_core.h
  void copyPropagator(QKXTM_Propagator<Float> &prop, int nu, int c2);                         // lib/qudaQKXTM_Vector.cpp:490-512
  void copyPropagator3D(QKXTM_Propagator3D<Float> &prop, int timeslice, int nu, int c2);      // lib/qudaQKXTM_Vector.cpp:463-488
  void write(char *filename);                      // "DiracFermion_Sink" LIME file from h_elem (AoS), lib/qudaQKXTM_Vector.cpp:510-702
  void castDoubleToFloat(QKXTM_Vector<double> &vecIn);
  void castFloatToDouble(QKXTM_Vector<float> &vecIn);
  double norm2Host();
  void apply_gamma5();
};

template <typename Float> class QKXTM_Propagator : public QKXTM_Field<Float> {  // include/qudaQKXTM.h:244-265
public:
  QKXTM_Propagator(ALLOCATION_FLAG alloc_flag, CLASS_ENUM classT);
  void absorbVectorToHost(QKXTM_Vector<Float> &vec, int nu, int c2);
  void absorbVectorToDevice(QKXTM_Vector<Float> &vec, int nu, int c2);
  void rotateToPhysicalBase_host(int sign);        // lib/qudaQKXTM_Propagator.cpp:117-180 (on h_elem)
  void rotateToPhysicalBase_device(int sign);      // lib/qudaQKXTM_Propagator.cpp:108-112: sign = +1 (up) / -1 (down)
  void conjugate();                                // lib/code_pieces/conjugate_propagator_core.h
  void apply_gamma5();                             // lib/code_pieces/apply_gamma5_propagator_core.h
};

template <typename Float> class QKXTM_Propagator3D : public QKXTM_Field<Float> {   // include/qudaQKXTM.h:267-277
public:
  QKXTM_Propagator3D(ALLOCATION_FLAG alloc_flag, CLASS_ENUM classT);
  void absorbTimeSlice(QKXTM_Propagator<Float> &prop, int timeslice);                 // lib/qudaQKXTM_Propagator.cpp:509-531
  void absorbVectorTimeSlice(QKXTM_Vector<Float> &vec, int timeslice, int nu, int c2); // lib/qudaQKXTM_Propagator.cpp:533-550
};

// the two-point part of QKXTM_Contraction (include/qudaQKXTM.h:283-389; lib/qudaQKXTM_Contraction.cpp:1563-1648)
template <typename Float> class QKXTM_Contraction {
public:
  QKXTM_Contraction() {}
  // corrMesons, MOMENTUM_SPACE: Float[T_local * Nmoms * 2][2][10], entry [it*Nmoms*2 + imom*2 + ri][iu][ip], already summed over
  // the ranks that share this rank's time slices (the reference leaves that sum on the space communicator's root);
  // POSITION_SPACE: Float[2 * V_local][2][10], entry [2*x_lex + ri][iu][ip].  Float = float or double (the reference's launcher
  // refuses double, lib/qudaQKXTM_kernels.cu:1222).
  void contractMesons(QKXTM_Propagator<Float> &prop1, QKXTM_Propagator<Float> &prop2, void *corrMesons, int isource, CORR_SPACE CorrSpace);
  // "ip it px py pz  re(up) im(up)  re(down) im(down)", time relative to the source, lib/qudaQKXTM_Contraction.cpp:1586-1599;
  // every rank calls it, rank 0 writes; on a t split the complete correlator is taken from the preceding contractMesons call (libtmq
  // returns the reduced result for all time slices to every rank: no MPI_Gather over the time communicator)
  void writeTwopMesons_ASCII(void *corrMesons, char *filename_out, int isource, CORR_SPACE CorrSpace);
  // corrBaryons, MOMENTUM_SPACE only: Float[T_local * Nmoms * 2][2][10][4][4], entry [it*Nmoms*2 + imom*2 + ri][iu][ip][gamma][gammap]
  // (lib/qudaQKXTM_Contraction.cpp:906-960), summed over the ranks that share this rank's time slices
  void contractBaryons(QKXTM_Propagator<Float> &prop1, QKXTM_Propagator<Float> &prop2, void *corrBaryons, int isource, CORR_SPACE CorrSpace);
  // "ip it px py pz gamma gammap  re(iu=0) im(iu=0)  re(iu=1) im(iu=1)", time relative to the source, sign flip where the time wraps
  // (anti-periodic boundary), lib/qudaQKXTM_Contraction.cpp:877-901
  void writeTwopBaryons_ASCII(void *corrBaryons, char *filename_out, int isource, CORR_SPACE CorrSpace);
  // fixed-sink sequential sources (lib/qudaQKXTM_Contraction.cpp:1652-1688): written into time slice `timeslice` of vec
  void seqSourceFixSinkPart1(QKXTM_Vector<Float> &vec, QKXTM_Propagator3D<Float> &prop1, QKXTM_Propagator3D<Float> &prop2, int timeslice, int nu,
                             int c2, WHICHPROJECTOR PID, WHICHPARTICLE testParticle);
  void seqSourceFixSinkPart2(QKXTM_Vector<Float> &vec, QKXTM_Propagator3D<Float> &prop, int timeslice, int nu, int c2, WHICHPROJECTOR PID,
                             WHICHPARTICLE testParticle);
  // lib/qudaQKXTM_Contraction.cpp:3008-3110: corrThp_local Float[T_local * Nmoms * 16 * 2], entry [it*Nmoms*16*2 + imom*16*2 + iop*2 + ri];
  // corrThp_noether Float[T_local * Nmoms * 4 * 2] (entry [.. + dir*2 + ri]) and corrThp_oneD Float[T_local * Nmoms * 4 * 16 * 2] (entry
  // [.. + dir*16*2 + iop*2 + ri]) may both be NULL (ultra-local insertion only); when given, gauge must hold the links on the device
  // and the lattice must not be split (the conserved-current and one-derivative insertions read the neighbours' propagators)
  void contractFixSink(QKXTM_Propagator<Float> &seqProp, QKXTM_Propagator<Float> &prop, QKXTM_Gauge<Float> &gauge, void *corrThp_local,
                       void *corrThp_noether, void *corrThp_oneD, WHICHPROJECTOR typeProj, WHICHPARTICLE testParticle, int partflag, int isource,
                       CORR_SPACE CorrSpace);
  // "<filename_out>.<proton|neutron>.<up|down>.{ultra_local,noether,oneD}.SS.xx.yy.zz.tt.dat": "iop it px py pz re im" (ultra_local and
  // noether, iop = direction there) and "iop dir it px py pz re im" (oneD), lib/qudaQKXTM_Contraction.cpp:2842-3000; the noether and oneD
  // files are written when their buffers are given
  void writeThrp_ASCII(void *corrThp_local, void *corrThp_noether, void *corrThp_oneD, WHICHPARTICLE testParticle, int partflag, char *filename_out,
                       int isource, int tsinkMtsource, CORR_SPACE CorrSpace);
};
int qkxtm_Nmoms();                                  // GK_Nmoms / GK_moms after init_qudaQKXTM
const int *qkxtm_moms();                            // [Nmoms][3]

// ---- exact deflation (include/qudaQKXTM.h:391-475, include/qudaQKXTM_utils.h:76-94) ------------------------------------
enum WHICHSPECTRUM { SR, LR, SM, LM, SI, LI };
typedef struct {
  int PolyDeg;                 // degree of the Chebyshev polynomial
  int nEv;                     // number of eigenvectors wanted
  int nKv;                     // size of the Krylov space
  WHICHSPECTRUM spectrumPart;  // SR or LR (eigenvalues of M^dag M are real and positive: SM = SR, LM = LR)
  bool isACC;
  double tolArpack;
  int maxIterArpack;
  char arpack_logfile[512];    // unused (no ARPACK)
  double amin, amax;
  bool isEven;
  bool isFullOp;               // false: even-odd M_pc^dag M_pc; true: the unpreconditioned M^dag M (what calc_loops deflates with)
  int modeArpack;              // unused
} qudaQKXTM_arpackInfo;

// QKXTM_Deflation for the even-odd or the full M^dag M: the Krylov basis and the eigenvectors stay resident in HBM (the reference
// keeps NkV host vectors and stages every ARPACK reverse-communication step through PCIe).  The eigensolver is a
// thick-restart Lanczos inside libtmq (tmq_eigensolve); eigenvalues/residuals are recomputed with the true operator as
// the reference does after zneupd.
template <typename Float> class QKXTM_Deflation {
  int PolyDeg, NeV, NkV;
  WHICHSPECTRUM spectrumPart;
  bool isACC, isEv, isFullOp;
  double tolArpack, amin, amax, flavor_sign;
  int maxIterArpack;
  long long total_length_per_NeV;
  size_t bytes_total_length_per_NeV;
  Float *eigenValues;          // [2 * NkV] (re, im) like the reference
  double *residuals;
  tmq_eigset *set;
  QudaInvertParam *invert_param;
  int nconv, nrestarts, nmatvec;
public:
  QKXTM_Deflation(QudaInvertParam *param, qudaQKXTM_arpackInfo arpackInfo);
  ~QKXTM_Deflation();
  Float *EigenValues() const { return eigenValues; }
  const double *Residuals() const { return residuals; }
  size_t Bytes_Per_NeV() const { return bytes_total_length_per_NeV; }
  long long Length_Per_NeV() const { return total_length_per_NeV; }
  int NeVs() const { return NeV; }
  int Converged() const { return nconv; }
  int MatVecs() const { return nmatvec; }
  void printInfo();
  void eigenSolver();                                                                  // Deflation.cpp:1069-1475
  void polynomialOperator(ColorSpinorField &out, const ColorSpinorField &in);         // :997-1063
  void deflateVector(QKXTM_Vector<Float> &vec_defl, QKXTM_Vector<Float> &vec_in);     // :614-800 (vec_in: host AoS)
  void projectVector(QKXTM_Vector<Float> &vec_defl, QKXTM_Vector<Float> &vec_in, int is, int NeV_defl);   // :1926-2060 (isFullOp)
  void ApplyMdagM(Float *vec_out, Float *vec_in, QudaInvertParam *param);             // :189-281
  void copyEigenVectorToQKXTM_Vector(int eigenVector_id, Float *vec);                 // :449-536 (full volume, AoS)
  tmq_eigset *EigenSet() const { return set; }
};

// accessors for drivers / tests
tmq_ctx *qkxtm_context();
typedef void (*qkxtm_error_handler)(const char *msg);
void qkxtm_set_error_handler(qkxtm_error_handler h);   // default: print and abort, like errorQuda

}  // namespace quda

// ---- solve entry points (include/qudaQKXTM.h:484-513) --------------------------------------------------------------
// MG_bench: the 12-column point-source propagator skeleton of lib/qudaQKXTM_interface.cpp:19-233 with the
// solver swapped for CG on M^dag M, as the calc_loops CG branch does (lib/qudaQKXTM_interface.cpp:2031-2038):
//   point source -> packVector -> loadVector -> uploadToCuda -> prepare -> in <- M^dag in -> CG -> reconstruct
//   -> downloadFromCuda -> scaleVector(2 kappa) if mass-normalised.
// gaugeSmeared: lexicographic links for the plaquette print (may be NULL); gauge: unused, as in the reference
// (the solver uses the field resident since loadGaugeQuda).  prop_out (optional, not in the reference, which
// discards the columns): 12 x V x 24 doubles, column-major host AoS.
void MG_bench(void **gaugeSmeared, void **gauge, QudaGaugeParam *gauge_param, QudaInvertParam *param,
              quda::qudaQKXTMinfo info, double *prop_out = nullptr);
// the per-RHS solve of calc_loops' CG branch (lib/qudaQKXTM_interface.cpp:2008-2041,2062) for one host source in
// the plug-in's AoS order [x_lex][s][c][ri]; the solution comes back in the same order.
void calc_loops_solve(double *h_solution, double *h_source, QudaInvertParam *param, quda::qudaQKXTMinfo info);
// QKXTM_Deflation::ApplyMdagM (lib/qudaQKXTM_Deflation.cpp:189-281, preconditioned branch): full-volume host
// vector in the plug-in's AoS order [x_lex][s][c][ri] -> packVector/loadVector/uploadToCuda(parity isEven) ->
// M_pc^dag M_pc -> downloadFromCuda/unloadVector/unpackVector; the other parity comes back zero-filled.
void ApplyMdagM(double *h_out, double *h_in, QudaInvertParam *param, bool isEven);

// calcMG_threepTwop_EvenOdd (include/qudaQKXTM.h:494-499, lib/qudaQKXTM_interface.cpp:236-1290), the TWO-POINT part:
// for every source position, 12 + 12 point-source solves (up: +mu, down: -mu; Gaussian-smeared source), columns cast to float
// and absorbed into K_prop_up / K_prop_down, sink smearing, rotateToPhysicalBase_device(+-1), contractBaryons, contractMesons, and
// the ASCII files "<filename_twop>.baryons.SS.xx.yy.zz.tt.dat" / "<filename_twop>.mesons.SS.xx.yy.zz.tt.dat".  CG on M^dag M replaces the reference's GCR + multigrid solver (the
// preconditionerUP/DN patch to quda.h, README:84-108).  Three-point functions (info.run3pt_src != 0) and HDF5 output are not built and
// are refused.
void calcMG_threepTwop_EvenOdd(void **gaugeSmeared, void **gauge, QudaGaugeParam *gauge_param, QudaInvertParam *param,
                               quda::qudaQKXTMinfo info, char *filename_twop, char *filename_threep, quda::WHICHPARTICLE NUCLEON);

// calcLowModeProjection (include/qudaQKXTM.h:508-510, lib/qudaQKXTM_interface.cpp:1342-1401): the low modes of M^dag M through
// QKXTM_Deflation::eigenSolver, with the reference's consistency checks (asymmetric operators only, parity of the operator and of
// arpackInfo must agree).  Returns nothing, like the reference; nconv / evals (optional, not in the reference) report the result.
void calcLowModeProjection(QudaInvertParam *evInvParam, quda::qudaQKXTM_arpackInfo arpackInfo, int *nconv = nullptr, double *evals = nullptr);

// ---- configuration I/O (include/QKXTM_read_conf.h:816-848) -----------------------------------------------------------
// reads this rank's sub-block of an ILDG / LIME configuration into the QDP even-odd host order of loadGaugeQuda and sets
// param->X to the local extents; no boundary condition is applied (call applyBoundaryCondition, as the drivers do)
void readLimeGauge(void **gauge, char *fname, QudaGaugeParam *param, QudaInvertParam *inv_param, int gridSize[4]);
void readLimeGaugeSmeared(void **gauge, char *fname, QudaGaugeParam *param, QudaInvertParam *inv_param, int gridSize[4]);
void applyBoundaryCondition(void **gauge, int Vh, QudaGaugeParam *gauge_param);

#endif
