/* blas_reference.h -- drop-in stand-in: the reference drivers include this upstream-QUDA header (qkxtm/Calc_Loops.cpp:7-15) without using anything
 * from it on the path this library provides.  See include/compat/README.md. */
#pragma once
#include "quda.h"
