/* misc.h -- drop-in stand-in for upstream QUDA's tests/misc.h, which the reference drivers include (qkxtm/Calc_Loops.cpp:13) but the
 * reference tree does not ship: declarations of the string <-> enum helpers whose DEFINITIONS are the reference's own qkxtm/misc.cpp
 * (:613-1217; get_schwarz_type is referenced by qkxtm/QKXTM_util.cpp but defined nowhere in the tree: libqkxtm_tmq.so provides it). */
#pragma once
#include "quda.h"
/* direction indices used by the link sanity checks of qkxtm/misc.cpp:233-249 */
#define XUP 0
#define YUP 1
#define ZUP 2
#define TUP 3
void display_spinor(void *spinor, int len, int precision);
void display_link(void *link, int len, int precision);
int link_sanity_check(void *link, int len, int precision, int dir, QudaGaugeParam *gaugeParam);
int site_link_sanity_check(void *link, int len, int precision, QudaGaugeParam *gaugeParam);
QudaVerbosity get_verbosity_type(char *s);
const char *get_verbosity_str(QudaVerbosity type);
QudaReconstructType get_recon(char *s);
QudaPrecision get_prec(char *s);
const char *get_prec_str(QudaPrecision prec);
const char *get_unitarization_str(bool svd_only);
const char *get_gauge_order_str(QudaGaugeFieldOrder order);
const char *get_recon_str(QudaReconstructType recon);
const char *get_test_type(int t);
QudaDslashType get_dslash_type(char *s);
const char *get_dslash_str(QudaDslashType type);
QudaMassNormalization get_mass_normalization_type(char *s);
const char *get_mass_normalization_str(QudaMassNormalization type);
QudaMatPCType get_matpc_type(char *s);
const char *get_matpc_str(QudaMatPCType type);
QudaSolveType get_solve_type(char *s);
const char *get_solve_str(QudaSolveType type);
QudaTwistFlavorType get_flavor_type(char *s);
const char *get_flavor_str(QudaTwistFlavorType type);
QudaInverterType get_solver_type(char *s);
const char *get_solver_str(QudaInverterType type);
QudaSchwarzType get_schwarz_type(char *s);
const char *get_quda_ver_str();
