/* countpairs_rp_pi_mocks.h -- drop-in C ABI for survey-geometry pair counts DD(rp, pi) from (RA, DEC, distance).
 * Replaces the reference interface mocks/DDrppi_mocks/countpairs_rp_pi_mocks.h:18-39 (Corrfunc v2.5.3): same symbol
 * names, argument order/meaning, result layout and error behaviour (EXIT_SUCCESS / EXIT_FAILURE + stderr message).
 * Inputs are HOST pointers of element size options->float_type (4 or 8): RA and DEC in degrees and the third array
 * as cz in km/s (options->is_comoving_dist = 0) or as a comoving distance (= 1).  cz is converted like the reference
 * does: its own redshift -> distance table (utils/set_cosmo_dist.c:27-75, plain Simpson's rule, reproduced bit for
 * bit) and GSL's linear interpolation, restated from GSL's source because GSL itself is absent (see
 * corrfunc_b200_cz_to_comoving in corrfunc_b200.h); a redshift outside the table returns EXIT_FAILURE where GSL's
 * error handler would abort.  RA / DEC out of range are shifted, and redshifts passed as cz are scaled, in place
 * (countpairs_rp_pi_mocks_impl.c.src:43-110).  Line of sight = pair midpoint.  The pair counting runs on the GPU
 * (sm_100a); there is no CPU fallback.
 */
#ifndef CORRFUNC_B200_COUNTPAIRS_RP_PI_MOCKS_H
#define CORRFUNC_B200_COUNTPAIRS_RP_PI_MOCKS_H
#include <stdint.h>
#include "corrfunc_b200_defs.h"
#ifdef __cplusplus
extern "C" {
#endif

/* 2-D arrays have (nbin+1)*(npibin+1) entries, index i*(npibin+1)+j, only j<npibin is meaningful. */
typedef struct {
    uint64_t *npairs;
    double *rupp;
    double *rpavg;
    double *weightavg;
    double pimax;
    int nbin;
    int npibin;
} results_countpairs_mocks;

extern int countpairs_mocks(const int64_t ND1, void *phi1 /* RA */, void *theta1 /* DEC */, void *czD1,
                            const int64_t ND2, void *phi2, void *theta2, void *czD2, const int numthreads,
                            const int autocorr, const char *binfile, const double pimax, const int cosmology,
                            results_countpairs_mocks *results, struct config_options *options,
                            struct extra_options *extra);
extern void free_results_mocks(results_countpairs_mocks *results);
/* mocks/DDrppi_mocks/countpairs_rp_pi_mocks_impl.h.src */
extern int countpairs_mocks_float(const int64_t ND1, float *phi1, float *theta1, float *czD1, const int64_t ND2,
                                  float *phi2, float *theta2, float *czD2, const int numthreads, const int autocorr,
                                  const char *binfile, const float pimax, const int cosmology,
                                  results_countpairs_mocks *results, struct config_options *options,
                                  struct extra_options *extra);
extern int countpairs_mocks_double(const int64_t ND1, double *phi1, double *theta1, double *czD1, const int64_t ND2,
                                   double *phi2, double *theta2, double *czD2, const int numthreads,
                                   const int autocorr, const char *binfile, const double pimax, const int cosmology,
                                   results_countpairs_mocks *results, struct config_options *options,
                                   struct extra_options *extra);

#ifdef __cplusplus
}
#endif
#endif
