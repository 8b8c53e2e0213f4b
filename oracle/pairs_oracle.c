/* pairs_oracle.c -- TEST INFRASTRUCTURE ONLY.
 * CPU restatement (oracle) of the reference's gridded pair-counting hot path; see oracle_impl.h.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load the library built
 * from this file.  The product (corrfunc_b200/) never links, imports or calls it.
 * Build: make -C oracle   ->  oracle/libpairs_oracle.so
 */
#define _GNU_SOURCE
#include <float.h>
#include <inttypes.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_DD 0
#define ORC_XI 1
#define ORC_RPPI 2
#define ORC_WP 3
#define ORC_SMU 4
#define ORC_THETA 5
#define ORC_RPPI_MOCKS 6 /* X,Y,Z arguments carry RA, DEC (deg) and the comoving distance */
#define ORC_SMU_MOCKS 7

/* utils/function_precision.h:20-21 */
#define ORC_PI_OVER_180 0.017453292519943295769236907684886127134428718885417254560971
#define ORC_INV_PI_OVER_180 57.29577951308232087679815481410517033240547246656432154916024

/* 1: wp / DDrppi follow the AVX-512 kernels' z-sorted, chunked control flow (see oracle_impl.h) */
static int orc_literal_kernels = 0;
void oracle_set_literal_kernels(int on) { orc_literal_kernels = on; }

#define REAL float
#define SUFFIX float
#define REAL_IS_DOUBLE 0
#include "oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef REAL_IS_DOUBLE

#define REAL double
#define SUFFIX double
#define REAL_IS_DOUBLE 1
#include "oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef REAL_IS_DOUBLE

void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
