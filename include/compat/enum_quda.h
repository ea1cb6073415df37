/* enum_quda.h -- drop-in stand-in (see quda.h in this directory) */
#pragma once
#include "../quda_tmq.h"
