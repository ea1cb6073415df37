/* invert_quda.h -- drop-in stand-in: included by qkxtm/misc.cpp:5, nothing of it is used */
#pragma once
#include "quda.h"
