/* countpairs_xi.h -- drop-in C ABI for 3-D correlation function xi(r) in a periodic box.
 * Replaces the reference interface theory/xi/countpairs_xi.h:20-36 (Corrfunc v2.5.3): same symbol names, argument
 * order/meaning, result layout and error behaviour (EXIT_SUCCESS / EXIT_FAILURE + stderr message).
 * Inputs are HOST pointers of element size options->float_type (4 or 8); result arrays are
 * malloc'ed by the callee and released with the matching free_results* call.
 * The pair counting itself runs on the GPU (sm_100a); there is no CPU fallback.
 */
#ifndef CORRFUNC_B200_COUNTPAIRS_XI_H
#define CORRFUNC_B200_COUNTPAIRS_XI_H
#include <stdint.h>
#include "corrfunc_b200_defs.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint64_t *npairs;
    double *xi;
    double *rupp;
    double *ravg;
    double *weightavg;
    int nbin;
} results_countpairs_xi;

extern int countpairs_xi(const int64_t ND1, void *X1, void *Y1, void *Z1, const double boxsize, const int numthreads,
                         const char *binfile, results_countpairs_xi *results, struct config_options *options,
                         struct extra_options *extra);
extern void free_results_xi(results_countpairs_xi *results);
/* theory/xi/countpairs_xi_impl.h.src:35-41 */
extern int countpairs_xi_float(const int64_t ND1, float *X1, float *Y1, float *Z1, const double boxsize,
                               const int numthreads, const char *binfile, results_countpairs_xi *results,
                               struct config_options *options, struct extra_options *extra);
extern int countpairs_xi_double(const int64_t ND1, double *X1, double *Y1, double *Z1, const double boxsize,
                                const int numthreads, const char *binfile, results_countpairs_xi *results,
                                struct config_options *options, struct extra_options *extra);


#ifdef __cplusplus
}
#endif
#endif
