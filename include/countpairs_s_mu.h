/* countpairs_s_mu.h -- drop-in C ABI for pair counts DD(s, mu).
 * Replaces the reference interface theory/DDsmu/countpairs_s_mu.h:19-41 (Corrfunc v2.5.3): same symbol names, argument
 * order/meaning, result layout and error behaviour (EXIT_SUCCESS / EXIT_FAILURE + stderr message).
 * Inputs are HOST pointers of element size options->float_type (4 or 8); result arrays are
 * malloc'ed by the callee and released with the matching free_results* call.
 * The pair counting itself runs on the GPU (sm_100a); there is no CPU fallback.
 */
#ifndef CORRFUNC_B200_COUNTPAIRS_S_MU_H
#define CORRFUNC_B200_COUNTPAIRS_S_MU_H
#include <stdint.h>
#include "corrfunc_b200_defs.h"
#ifdef __cplusplus
extern "C" {
#endif

/* 2-D arrays have (nsbin+1)*(nmu_bins+1) entries, index i*(nmu_bins+1)+j, only j<nmu_bins is
 * meaningful (theory/DDsmu/countpairs_s_mu_impl.c.src:668-700). */
typedef struct {
    uint64_t *npairs;
    double *supp;
    double *savg;
    double mu_max;
    double mu_min; /* always 0 */
    double *weightavg;
    int nsbin;
    int nmu_bins;
} results_countpairs_s_mu;

extern int countpairs_s_mu(const int64_t ND1, void *X1, void *Y1, void *Z1, const int64_t ND2, void *X2, void *Y2,
                           void *Z2, const int numthreads, const int autocorr, const char *sbinfile,
                           const double mu_max, const int nmu_bins, results_countpairs_s_mu *results,
                           struct config_options *options, struct extra_options *extra);
extern void free_results_s_mu(results_countpairs_s_mu *results);
/* theory/DDsmu/countpairs_s_mu_impl.h.src:39-48 */
extern int countpairs_s_mu_float(const int64_t ND1, float *X1, float *Y1, float *Z1, const int64_t ND2, float *X2,
                                 float *Y2, float *Z2, const int numthreads, const int autocorr, const char *sbinfile,
                                 const double mu_max, const int nmu_bins, results_countpairs_s_mu *results,
                                 struct config_options *options, struct extra_options *extra);
extern int countpairs_s_mu_double(const int64_t ND1, double *X1, double *Y1, double *Z1, const int64_t ND2, double *X2,
                                  double *Y2, double *Z2, const int numthreads, const int autocorr,
                                  const char *sbinfile, const double mu_max, const int nmu_bins,
                                  results_countpairs_s_mu *results, struct config_options *options,
                                  struct extra_options *extra);


#ifdef __cplusplus
}
#endif
#endif
