/* qudaQKXTM_utils.h -- drop-in stand-in (reference include/qudaQKXTM_utils.h): see qudaQKXTM.h in this directory */
#pragma once
#include "../qudaQKXTM_tmq.h"
