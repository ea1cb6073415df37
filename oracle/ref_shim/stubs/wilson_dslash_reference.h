// TEST INFRASTRUCTURE stand-in (see quda.h in this directory): nothing of this upstream header is needed
#pragma once
