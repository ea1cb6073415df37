// TEST INFRASTRUCTURE stand-in (see quda.h in this directory): declarations of upstream tests/misc.h used by qkxtm/QKXTM_util.cpp
#pragma once
#include <quda.h>
QudaPrecision get_prec(char *s);
QudaReconstructType get_recon(char *s);
QudaInverterType get_solver_type(char *s);
QudaDslashType get_dslash_type(char *s);
QudaMassNormalization get_mass_normalization_type(char *s);
QudaMatPCType get_matpc_type(char *s);
QudaSolveType get_solve_type(char *s);
QudaTwistFlavorType get_flavor_type(char *s);
QudaVerbosity get_verbosity_type(char *s);
QudaSchwarzType get_schwarz_type(char *s);
const char *get_quda_ver_str();
const char *get_prec_str(QudaPrecision prec);
const char *get_recon_str(QudaReconstructType recon);
