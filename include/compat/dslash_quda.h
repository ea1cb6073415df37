/* dslash_quda.h -- drop-in stand-in: upstream QUDA header included by qkxtm/QKXTM_util.cpp (:16-17) for its logging macros and the comm layer */
#pragma once
#include <sys/time.h>
#include "util_quda.h"
#include "comm_quda.h"
