/* defs.h -- the name the reference's headers and its users include (utils/defs.h); everything lives in
 * corrfunc_b200_defs.h, whose struct layouts are binary-identical to the reference's (tests/test_cpu_abi.py). */
#ifndef CORRFUNC_B200_DEFS_COMPAT_H
#define CORRFUNC_B200_DEFS_COMPAT_H
#include "corrfunc_b200_defs.h"
#endif
