/* countpairs_rp_pi.h -- drop-in C ABI for pair counts DD(rp, pi).
 * Replaces the reference interface theory/DDrppi/countpairs_rp_pi.h:19-39 (Corrfunc v2.5.3): same symbol names, argument
 * order/meaning, result layout and error behaviour (EXIT_SUCCESS / EXIT_FAILURE + stderr message).
 * Inputs are HOST pointers of element size options->float_type (4 or 8); result arrays are
 * malloc'ed by the callee and released with the matching free_results* call.
 * The pair counting itself runs on the GPU (sm_100a); there is no CPU fallback.
 */
#ifndef CORRFUNC_B200_COUNTPAIRS_RP_PI_H
#define CORRFUNC_B200_COUNTPAIRS_RP_PI_H
#include <stdint.h>
#include "corrfunc_b200_defs.h"
#ifdef __cplusplus
extern "C" {
#endif

/* 2-D arrays have (nbin+1)*(npibin+1) entries, index i*(npibin+1)+j, only j<npibin is meaningful
 * (theory/DDrppi/countpairs_rp_pi_impl.c.src:667-697). */
typedef struct {
    uint64_t *npairs;
    double *rupp;
    double *rpavg;
    double *weightavg;
    double pimax;
    int nbin;
    int npibin;
} results_countpairs_rp_pi;

extern int countpairs_rp_pi(const int64_t ND1, void *X1, void *Y1, void *Z1, const int64_t ND2, void *X2, void *Y2,
                            void *Z2, const int numthreads, const int autocorr, const char *binfile,
                            const double pimax, results_countpairs_rp_pi *results, struct config_options *options,
                            struct extra_options *extra);
extern void free_results_rp_pi(results_countpairs_rp_pi *results);
/* theory/DDrppi/countpairs_rp_pi_impl.h.src:37-45 */
extern int countpairs_rp_pi_float(const int64_t ND1, float *X1, float *Y1, float *Z1, const int64_t ND2, float *X2,
                                  float *Y2, float *Z2, const int numthreads, const int autocorr, const char *binfile,
                                  const double pimax, results_countpairs_rp_pi *results,
                                  struct config_options *options, struct extra_options *extra);
extern int countpairs_rp_pi_double(const int64_t ND1, double *X1, double *Y1, double *Z1, const int64_t ND2, double *X2,
                                   double *Y2, double *Z2, const int numthreads, const int autocorr,
                                   const char *binfile, const double pimax, results_countpairs_rp_pi *results,
                                   struct config_options *options, struct extra_options *extra);


#ifdef __cplusplus
}
#endif
#endif
