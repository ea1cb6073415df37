/* contractQuda.h -- drop-in stand-in: included by qkxtm/CalcLowModeProjection.cpp:29, nothing of it is used on the built path */
#pragma once
#include "quda.h"
