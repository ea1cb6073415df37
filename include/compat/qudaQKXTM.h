/* qudaQKXTM.h -- drop-in stand-in for the plug-in's header (reference include/qudaQKXTM.h): containers, parameter structs and the
 * entry points MG_bench / calcMG_threepTwop_EvenOdd / calc_loops / calcLowModeProjection, served by libqkxtm_tmq.so. */
#pragma once
#include "../qudaQKXTM_tmq.h"
