/* countpairs_theta_mocks.h -- drop-in C ABI for angular pair counts DD(theta) from RA/DEC (degrees).
 * Replaces the reference interface mocks/DDtheta_mocks/countpairs_theta_mocks.h:19-35 (Corrfunc v2.5.3): same symbol names, argument
 * order/meaning, result layout and error behaviour (EXIT_SUCCESS / EXIT_FAILURE + stderr message).
 * Inputs are HOST pointers of element size options->float_type (4 or 8); result arrays are
 * malloc'ed by the callee and released with the matching free_results* call.
 * The pair counting itself runs on the GPU (sm_100a); there is no CPU fallback.
 */
#ifndef CORRFUNC_B200_COUNTPAIRS_THETA_MOCKS_H
#define CORRFUNC_B200_COUNTPAIRS_THETA_MOCKS_H
#include <stdint.h>
#include "corrfunc_b200_defs.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint64_t *npairs;
    double *theta_upp;
    double *theta_avg;
    double *weightavg;
    int nbin;
} results_countpairs_theta;

/* phi = RA, theta = DEC, both in degrees.  As in the reference, RA in [-180,180] / DEC in [0,180]
 * inputs are shifted IN PLACE to [0,360] / [-90,90] (countpairs_theta_mocks_impl.c.src:42-84). */
extern int countpairs_theta_mocks(const int64_t ND1, void *phi1, void *theta1, const int64_t ND2, void *phi2,
                                  void *theta2, const int numthreads, const int autocorr, const char *binfile,
                                  results_countpairs_theta *results, struct config_options *options,
                                  struct extra_options *extra);
extern void free_results_countpairs_theta(results_countpairs_theta *results);
/* mocks/DDtheta_mocks/countpairs_theta_mocks_impl.h.src:39-45 */
extern int countpairs_theta_mocks_float(const int64_t ND1, float *phi1, float *theta1, const int64_t ND2, float *phi2,
                                        float *theta2, const int numthreads, const int autocorr, const char *binfile,
                                        results_countpairs_theta *results, struct config_options *options,
                                        struct extra_options *extra);
extern int countpairs_theta_mocks_double(const int64_t ND1, double *phi1, double *theta1, const int64_t ND2,
                                         double *phi2, double *theta2, const int numthreads, const int autocorr,
                                         const char *binfile, results_countpairs_theta *results,
                                         struct config_options *options, struct extra_options *extra);


#ifdef __cplusplus
}
#endif
#endif
