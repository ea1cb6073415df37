/* quda.h -- drop-in stand-in for upstream QUDA's public header: the slice of the C API that the QKXTM drivers use, served by
 * libqkxtm_tmq.so (include/quda_tmq.h).  A driver written against the reference (qkxtm/MG_Bench.cpp, qkxtm/Calc_Loops.cpp, ...) compiles
 * with -I include/compat -I include and links against -lqkxtm_tmq -ltmq.  See include/compat/README.md. */
#pragma once
#include "../quda_tmq.h"
