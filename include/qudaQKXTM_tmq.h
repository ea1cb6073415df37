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

// ---- the slice of quda.h / enum_quda.h the path reads: enums, QudaGaugeParam, QudaInvertParam, QudaMultigridParam and the
//      C API entry points (initQuda, loadGaugeQuda, invertQuda, ...) -----------------------------------------------------------------
#include "quda_tmq.h"

namespace quda {

enum ALLOCATION_FLAG { NONE, HOST, DEVICE, BOTH, BOTH_EXTRA };                 // include/qudaQKXTM_utils.h:126
enum CLASS_ENUM { FIELD, GAUGE, VECTOR, PROPAGATOR, PROPAGATOR3D, VECTOR3D };  // include/qudaQKXTM_utils.h:127

// include/qudaQKXTM_utils.h:16-23
#define MAX_NSOURCES 1000
#define MAX_NMOMENTA 5000
#define MAX_TSINK 10
#define MAX_DEFLSTEPS 10
#define MAX_LP_CRIT 10
#define MAX_PROJS 5
// include/qudaQKXTM_utils.h:25-29
#define LEXIC(it,iz,iy,ix,L) ( (it)*L[0]*L[1]*L[2] + (iz)*L[0]*L[1] + (iy)*L[0] + (ix) )
#define LEXIC_TZY(it,iz,iy,L) ( (it)*L[1]*L[2] + (iz)*L[1] + (iy) )
#define LEXIC_TZX(it,iz,ix,L) ( (it)*L[0]*L[2] + (iz)*L[0] + (ix) )
#define LEXIC_TYX(it,iy,ix,L) ( (it)*L[0]*L[1] + (iy)*L[0] + (ix) )
#define LEXIC_ZYX(iz,iy,ix,L) ( (iz)*L[0]*L[1] + (iy)*L[0] + (ix) )

enum SOURCE_T { UNITY, RANDOM };                                               // include/qudaQKXTM_utils.h:41-43
enum CORR_SPACE { POSITION_SPACE, MOMENTUM_SPACE };
enum FILE_WRITE_FORMAT { ASCII_FORM, HDF5_FORM };
enum WHICHPARTICLE { PROTON, NEUTRON };                                        // include/qudaQKXTM_utils.h:128-132
enum WHICHPROJECTOR { G4, G5G123, G5G1, G5G2, G5G3 };
enum THRP_TYPE { THRP_LOCAL, THRP_NOETHER, THRP_ONED };
enum APEDIM { D3, D4 };

// qudaQKXTMinfo, FIELD FOR FIELD as include/qudaQKXTM_utils.h:45-75: every entry point takes it BY VALUE, so the layout is part of the
// ABI (tests/test_dropin_compile.py static_asserts sizeof / offsetof of every member against the reference header).  Members the built
// paths do not read (nsmearAPE / alphaAPE, the name tables, HighMomForm, csw, ...) are present and ignored.
typedef struct {
  int nsmearAPE;
  int nsmearGauss;
  double alphaAPE;
  double alphaGauss;
  int lL[QUDAQKXTM_DIM];
  int Nsources;
  int sourcePosition[MAX_NSOURCES][QUDAQKXTM_DIM];    // global (x, y, z, t)
  QudaPrecision Precision;
  int Q_sq;                            // momenta with p^2 <= Q_sq (createMomenta, lib/qudaQKXTM_kernels.cu:98-116)
  int Ntsink;                          // sink-source separations of the three-point function
  int Nproj[MAX_TSINK];
  int traj;
  bool check_files;
  char *thrp_type[3];                  // filled by the entry points themselves (lib/qudaQKXTM_interface.cpp:276-288), never read from the caller
  char *thrp_proj_type[5];
  char *baryon_type[10];
  char *meson_type[10];
  int tsinkSource[MAX_TSINK];
  int proj_list[MAX_TSINK][MAX_PROJS]; // WHICHPROJECTOR values
  int run3pt_src[MAX_NSOURCES];        // != 0: also the fixed-sink three-point function for this source
  FILE_WRITE_FORMAT CorrFileFormat;    // ASCII_FORM only (no HDF5 here)
  SOURCE_T source_type;                // stochastic sources of calc_loops: RANDOM (Z4 noise) or UNITY
  CORR_SPACE CorrSpace;
  bool HighMomForm;
  bool isEven;
  double kappa;
  double mu;
  double csw;
  double inv_tol;
} qudaQKXTMinfo;

// include/qudaQKXTM_utils.h:77-94 (the reference guards it with HAVE_ARPACK; the eigensolver here is libtmq's own, so it is always there)
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
  double amin;
  double amax;
  bool isEven;
  bool isFullOp;               // false: even-odd M_pc^dag M_pc; true: the unpreconditioned M^dag M (what calc_loops deflates with)
  int modeArpack;              // unused
} qudaQKXTM_arpackInfo;

// include/qudaQKXTM_utils.h:96-124 (with HAVE_ARPACK: nSteps_defl / deflStep are members)
typedef struct {
  int Nstoch;                  // number of stochastic sources
  unsigned long int seed;      // gsl_rng_ranlux seed; rank r uses seed + r * seed (lib/qudaQKXTM_interface.cpp:1951)
  int Ndump;                   // dump every Ndump sources
  char loop_fname[512];
  int nSteps_defl;
  int deflStep[MAX_DEFLSTEPS];
  int traj;
  int Nprint;
  int Nmoms;
  int Qsq = 0;
  FILE_WRITE_FORMAT FileFormat;
  char *loop_type[6];          // filled by calc_loops itself (lib/qudaQKXTM_interface.cpp:1517-1534)
  bool loop_oneD[6];
  int k_probing;               // <= 0: hierarchical probing off
  int hadamLow;
  int hadamHigh;
  bool spinColorDil;           // spin-colour dilution: 12 solves per noise vector
  bool HighMomForm;
  double kappa;
  double mu;
  double csw;
  double inv_tol;
} qudaQKXTM_loopInfo;

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
  long long ghost_length;                          // ghost sites behind the local volume (0 on an unpartitioned lattice)
  size_t bytes_ghost_length, bytes_total_plus_ghost_length;
  Float *h_elem, *h_elem_backup, *d_elem;
  bool isAllocHost, isAllocDevice, isAllocHostBackup;
  void create_host();
  void create_host_backup();
  void destroy_host();
  void destroy_host_backup();
  void create_device();
  void destroy_device();
  void exchange_ghost_device();                    // the device-side body of the ghost trio of the derived containers
public:
  QKXTM_Field(ALLOCATION_FLAG alloc_flag, CLASS_ENUM classT);
  virtual ~QKXTM_Field();
  void zero_host();
  void zero_host_backup();
  void zero_device();
  Float *H_elem() const { return h_elem; }
  Float *D_elem() const { return d_elem; }
  size_t Bytes_total() const { return bytes_total_length; }
  size_t Bytes_ghost() const { return bytes_ghost_length; }
  size_t Bytes_total_plus_ghost() const { return bytes_total_plus_ghost_length; }
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
  // include/qudaQKXTM.h:177-179: after the three calls the ghost region of the DEVICE array ([ncomp][V] followed by the plus / minus ghost of
  // every partitioned dimension) holds the neighbours' boundary slices.  The exchange runs on the device inside cpuExchangeGhost.
  void ghostToHost();
  void cpuExchangeGhost();
  void ghostToDevice();
  void loadGauge();
  double calculatePlaq();                          // prints like the reference and also returns the value (all-reduced over the ranks)
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
  void ghostToHost();                              // include/qudaQKXTM.h:199-201 (see QKXTM_Gauge)
  void cpuExchangeGhost();
  void ghostToDevice();
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
  void ghostToHost();                              // include/qudaQKXTM.h:244-246 (see QKXTM_Gauge)
  void cpuExchangeGhost();
  void ghostToDevice();
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
  void projectVector(QKXTM_Vector<Float> &vec_defl, QKXTM_Vector<Float> &vec_in, int is);                 // :1931-2059 (isFullOp; all NeV vectors)
  void projectVector(QKXTM_Vector<Float> &vec_defl, QKXTM_Vector<Float> &vec_in, int is, int NeV_defl);   // :2063-2190 (isFullOp)
  void MapEvenOddToFull();                                                                                // :285-335 (nothing to do here: see the .cpp)
  void MapEvenOddToFull(int i);
  void ApplyMdagM(Float *vec_out, Float *vec_in, QudaInvertParam *param);             // :189-281
  void copyEigenVectorToQKXTM_Vector(int eigenVector_id, Float *vec);                 // :449-536 (full volume, AoS)
  tmq_eigset *EigenSet() const { return set; }
};

// accessors for drivers / tests
tmq_ctx *qkxtm_context();
typedef void (*qkxtm_error_handler)(const char *msg);
void qkxtm_set_error_handler(qkxtm_error_handler h);   // default: print and abort, like errorQuda

}  // namespace quda

// ---- solve entry points: EXACTLY the reference's signatures (include/qudaQKXTM.h:484-513; structs by value), plus overloads that
//      hand results back to the caller.  A driver written against the reference header links against these symbols unchanged. -------
// MG_bench: the 12-column point-source propagator skeleton of lib/qudaQKXTM_interface.cpp:19-233 with the
// solver swapped for CG on M^dag M, as the calc_loops CG branch does (lib/qudaQKXTM_interface.cpp:2031-2038):
//   point source -> packVector -> loadVector -> uploadToCuda -> prepare -> in <- M^dag in -> CG -> reconstruct
//   -> downloadFromCuda -> scaleVector(2 kappa) if mass-normalised.
// gaugeSmeared: lexicographic links for the plaquette print (may be NULL); gauge: unused, as in the reference
// (the solver uses the field resident since loadGaugeQuda).
void MG_bench(void **gaugeSmeared, void **gauge, QudaGaugeParam *gauge_param, QudaInvertParam *param, quda::qudaQKXTMinfo info);
// overload (not in the reference, which discards the columns): prop_out = 12 x V x 24 doubles, column-major host AoS; the D2H copy of
// column k runs behind the solve of column k+1
void MG_bench(void **gaugeSmeared, void **gauge, QudaGaugeParam *gauge_param, QudaInvertParam *param, quda::qudaQKXTMinfo info, double *prop_out);

// calc_loops (include/qudaQKXTM.h:501-507, lib/qudaQKXTM_interface.cpp:1409-2233), everything except the loop CONTRACTIONS (SURVEY.md 2:
// out of scope; upstream's contract / covariant-derivative conventions are not in the tree):
//   parameter checks (:1473-1494) -> QKXTM_Deflation::eigenSolver on the operator of EVparam (:1725-1736) -> [exact part: hook only]
//   -> plaquette of gaugeToPlaquette (:1843-1848) -> for every stochastic source is < Nstoch: Z4 / unity noise from gsl_rng_ranlux
//   seeded seed + rank * seed (:1951,1982; the generator is restated and pinned to GSL's own known-answer test), for every Hadamard
//   vector of the hierarchical probing (k_probing > 0, :1438-1458, lib/qudaQKXTM_utils.cpp:476-717) and every spin-colour component
//   (spinColorDil, :1994-2005): packVector -> loadVector -> uploadToCuda -> prepare -> M^dag -> CG -> reconstruct (:2008-2041), then for
//   every deflation step downloadFromCuda -> projectVector(NeV_defl) -> uploadToCuda (:2056-2067) -> [oneEndTrick_w_One_Der: hook only].
// Where the reference contracts, the installed hook (if any) receives the projected solution; without a hook the step is a no-op and
// no loop files are written.  HDF5 output and the GCR + multigrid solver are refused.
void calc_loops(void **gaugeToPlaquette, QudaInvertParam *EVparam, QudaInvertParam *param, QudaGaugeParam *gauge_param,
                quda::qudaQKXTM_arpackInfo arpackInfo, quda::qudaQKXTM_loopInfo loopInfo, quda::qudaQKXTMinfo info);
namespace quda {
// what calc_loops hands to the hook in place of the reference's contraction calls
struct qkxtm_loop_event {
  int kind;                    // 0: eigenpair n of the exact part (Loop_w_One_Der_FullOp_Exact, :1775); 1: projected solution (oneEndTrick_w_One_Der, :2074)
  int is, ih, sc;              // stochastic source, Hadamard vector, spin-colour component (kind 1); n in `is` for kind 0
  int dstep, NeV_defl;         // deflation step and its number of projected-out eigenvectors (kind 1)
  ColorSpinorField *x;         // kind 1: the projected full-volume solution on the device; kind 0: eigenvector n (FULL field)
  double eigenvalue;           // kind 0
  const double *h_source;      // kind 1: the (diluted) host source of this solve, plug-in AoS order [x_lex][s][c][ri]
  int iter;                    // kind 1: CG iterations of this solve
  double true_res;
};
typedef void (*qkxtm_loop_hook)(const qkxtm_loop_event *ev, void *user);
void qkxtm_set_loop_hook(qkxtm_loop_hook hook, void *user);
// the noise / dilution helpers of lib/qudaQKXTM_utils.cpp on host vectors in the plug-in's AoS order (exposed for tests and drivers)
void *qkxtm_rng_alloc(unsigned long int seed);                              // gsl_rng_alloc(gsl_rng_ranlux) + gsl_rng_set
void qkxtm_rng_free(void *rng);
unsigned long int qkxtm_rng_get(void *rng);                                 // gsl_rng_get
unsigned long int qkxtm_rng_uniform_int(void *rng, unsigned long int n);    // gsl_rng_uniform_int
template <typename Float> void getStochasticRandomSource(void *spinorIn, void *rng, SOURCE_T source_type);                  // :148-180
unsigned short int *hch_coloring(int k, int d);                                                                              // :666-717 (malloc'ed)
int HadamardElements(int i, int j);                                                                                          // :696-708
template <typename Float> void get_probing4D_spinColor_dilution(void *temp_input_vector, void *input_vector, unsigned short int *Vc, int ih, int sc);
template <typename Float> void get_spinColor_dilution(void *temp_input_vector, void *input_vector, int sc);
template <typename Float> void get_probing4D_dilution(void *temp_input_vector, void *input_vector, unsigned short int *Vc, int ih);
}  // namespace quda

// the per-RHS solve of calc_loops' CG branch (lib/qudaQKXTM_interface.cpp:2008-2041,2062) for one host source in
// the plug-in's AoS order [x_lex][s][c][ri]; the solution comes back in the same order.
void calc_loops_solve(double *h_solution, double *h_source, QudaInvertParam *param, quda::qudaQKXTMinfo info);
// QKXTM_Deflation::ApplyMdagM (lib/qudaQKXTM_Deflation.cpp:189-281, preconditioned branch): full-volume host
// vector in the plug-in's AoS order [x_lex][s][c][ri] -> packVector/loadVector/uploadToCuda(parity isEven) ->
// M_pc^dag M_pc -> downloadFromCuda/unloadVector/unpackVector; the other parity comes back zero-filled.
void ApplyMdagM(double *h_out, double *h_in, QudaInvertParam *param, bool isEven);

// calcMG_threepTwop_EvenOdd (include/qudaQKXTM.h:494-499, lib/qudaQKXTM_interface.cpp:236-1290): for every source position, 12 + 12
// point-source solves (up: +mu, down: -mu; Gaussian-smeared source), columns cast to float and absorbed into K_prop_up / K_prop_down,
// the fixed-sink three-point functions where info.run3pt_src != 0, sink smearing, rotateToPhysicalBase_device(+-1), contractBaryons,
// contractMesons, and the ASCII files "<filename_twop>.baryons.SS.xx.yy.zz.tt.dat" / "<filename_twop>.mesons.SS.xx.yy.zz.tt.dat".  CG on
// M^dag M replaces the reference's GCR + multigrid solver (the preconditionerUP/DN patch to quda.h, README:84-108).  HDF5 output is refused.
void calcMG_threepTwop_EvenOdd(void **gaugeSmeared, void **gauge, QudaGaugeParam *gauge_param, QudaInvertParam *param,
                               quda::qudaQKXTMinfo info, char *filename_twop, char *filename_threep, quda::WHICHPARTICLE NUCLEON);

// calcLowModeProjection (include/qudaQKXTM.h:508-510, lib/qudaQKXTM_interface.cpp:1342-1401): the low modes of M^dag M through
// QKXTM_Deflation::eigenSolver, with the reference's consistency checks (asymmetric operators only, parity of the operator and of
// arpackInfo must agree).  Returns nothing, like the reference.
void calcLowModeProjection(QudaInvertParam *evInvParam, quda::qudaQKXTM_arpackInfo arpackInfo);
// overload: nconv / evals report the result
void calcLowModeProjection(QudaInvertParam *evInvParam, quda::qudaQKXTM_arpackInfo arpackInfo, int *nconv, double *evals);

// ---- configuration I/O (include/QKXTM_read_conf.h:816-848) -----------------------------------------------------------
// reads this rank's sub-block of an ILDG / LIME configuration into the QDP even-odd host order of loadGaugeQuda and sets
// param->X to the local extents; no boundary condition is applied (call applyBoundaryCondition, as the drivers do)
void readLimeGauge(void **gauge, char *fname, QudaGaugeParam *param, QudaInvertParam *inv_param, int gridSize[4]);
void readLimeGaugeSmeared(void **gauge, char *fname, QudaGaugeParam *param, QudaInvertParam *inv_param, int gridSize[4]);
void applyBoundaryCondition(void **gauge, int Vh, QudaGaugeParam *gauge_param);

#endif
