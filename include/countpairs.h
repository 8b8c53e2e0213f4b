/* countpairs.h -- drop-in C ABI for 3-D pair counts DD(r).
 * Replaces the reference interface theory/DD/countpairs.h:20-37 (Corrfunc v2.5.3): same symbol names, argument
 * order/meaning, result layout and error behaviour (EXIT_SUCCESS / EXIT_FAILURE + stderr message).
 * Inputs are HOST pointers of element size options->float_type (4 or 8); result arrays are
 * malloc'ed by the callee and released with the matching free_results* call.
 * The pair counting itself runs on the GPU (sm_100a); there is no CPU fallback.
 */
#ifndef CORRFUNC_B200_COUNTPAIRS_H
#define CORRFUNC_B200_COUNTPAIRS_H
#include <stdint.h>
#include "corrfunc_b200_defs.h"
#ifdef __cplusplus
extern "C" {
#endif

/* npairs/rupp/rpavg/weightavg have nbin entries; entry 0 is junk, entries 1..nbin-1 are the bins
 * [rupp[i-1], rupp[i]) (theory/DD/countpairs_impl.c.src:667-690). */
typedef struct {
    uint64_t *npairs;
    double *rupp;
    double *rpavg;
    double *weightavg;
    int nbin;
} results_countpairs;

extern int countpairs(const int64_t ND1, void *X1, void *Y1, void *Z1, const int64_t ND2, void *X2, void *Y2,
                      void *Z2, const int numthreads, const int autocorr, const char *binfile,
                      results_countpairs *results, struct config_options *options,
                      struct extra_options *extra) __attribute__((warn_unused_result));
extern void free_results(results_countpairs *results);
/* Precision-suffixed entry points of the reference's static libraries (theory/DD/countpairs_impl.h.src:37-44):
 * typed pointers, options->float_type is ignored (set to the suffix's width for the duration of the call). */
extern int countpairs_float(const int64_t ND1, float *X1, float *Y1, float *Z1, const int64_t ND2, float *X2, float *Y2,
                            float *Z2, const int numthreads, const int autocorr, const char *binfile,
                            results_countpairs *results, struct config_options *options, struct extra_options *extra);
extern int countpairs_double(const int64_t ND1, double *X1, double *Y1, double *Z1, const int64_t ND2, double *X2,
                             double *Y2, double *Z2, const int numthreads, const int autocorr, const char *binfile,
                             results_countpairs *results, struct config_options *options, struct extra_options *extra);


#ifdef __cplusplus
}
#endif
#endif
