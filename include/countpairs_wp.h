/* countpairs_wp.h -- drop-in C ABI for projected correlation function wp(rp) in a periodic box.
 * Replaces the reference interface theory/wp/countpairs_wp.h:20-39 (Corrfunc v2.5.3): same symbol names, argument
 * order/meaning, result layout and error behaviour (EXIT_SUCCESS / EXIT_FAILURE + stderr message).
 * Inputs are HOST pointers of element size options->float_type (4 or 8); result arrays are
 * malloc'ed by the callee and released with the matching free_results* call.
 * The pair counting itself runs on the GPU (sm_100a); there is no CPU fallback.
 */
#ifndef CORRFUNC_B200_COUNTPAIRS_WP_H
#define CORRFUNC_B200_COUNTPAIRS_WP_H
#include <stdint.h>
#include "corrfunc_b200_defs.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint64_t *npairs;
    double *wp;
    double *rupp;
    double *rpavg;
    double *weightavg;
    double pimax;
    int nbin;
} results_countpairs_wp;

extern int countpairs_wp(const int64_t ND1, void *X1, void *Y1, void *Z1, const double boxsize, const int numthreads,
                         const char *binfile, const double pimax, results_countpairs_wp *result,
                         struct config_options *options, struct extra_options *extra)
    __attribute__((warn_unused_result));
extern void free_results_wp(results_countpairs_wp *results);
/* theory/wp/countpairs_wp_impl.h.src:37-44 */
extern int countpairs_wp_float(const int64_t ND1, float *X1, float *Y1, float *Z1, const double boxsize,
                               const int numthreads, const char *binfile, const double pimax,
                               results_countpairs_wp *result, struct config_options *options,
                               struct extra_options *extra);
extern int countpairs_wp_double(const int64_t ND1, double *X1, double *Y1, double *Z1, const double boxsize,
                                const int numthreads, const char *binfile, const double pimax,
                                results_countpairs_wp *result, struct config_options *options,
                                struct extra_options *extra);


#ifdef __cplusplus
}
#endif
#endif
