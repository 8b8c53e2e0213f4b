/* countpairs_s_mu_mocks.h -- drop-in C ABI for survey-geometry pair counts DD(s, mu) from (RA, DEC, distance).
 * Replaces the reference interface mocks/DDsmu_mocks/countpairs_s_mu_mocks.h:19-43 (Corrfunc v2.5.3): same symbol
 * names, argument order/meaning, result layout and error behaviour.  See countpairs_rp_pi_mocks.h for the input
 * contract (host pointers, degrees, cz in km/s or comoving distance per options->is_comoving_dist).
 */
#ifndef CORRFUNC_B200_COUNTPAIRS_S_MU_MOCKS_H
#define CORRFUNC_B200_COUNTPAIRS_S_MU_MOCKS_H
#include <stdint.h>
#include "corrfunc_b200_defs.h"
#ifdef __cplusplus
extern "C" {
#endif

/* 2-D arrays have (nsbin+1)*(nmu_bins+1) entries, index i*(nmu_bins+1)+j, only j<nmu_bins is meaningful. */
typedef struct {
    uint64_t *npairs;
    double *supp;
    double *savg;
    double mu_max;
    double mu_min; /* not used -> 0.0 */
    double *weightavg;
    int nsbin;
    int nmu_bins;
} results_countpairs_mocks_s_mu;

extern int countpairs_mocks_s_mu(const int64_t ND1, void *phi1 /* RA */, void *theta1 /* DEC */, void *czD1,
                                 const int64_t ND2, void *phi2, void *theta2, void *czD2, const int numthreads,
                                 const int autocorr, const char *sbinfile, const double mu_max, const int nmu_bins,
                                 const int cosmology, results_countpairs_mocks_s_mu *results,
                                 struct config_options *options, struct extra_options *extra);
extern void free_results_mocks_s_mu(results_countpairs_mocks_s_mu *results);
/* mocks/DDsmu_mocks/countpairs_s_mu_mocks_impl.h.src */
extern int countpairs_mocks_s_mu_float(const int64_t ND1, float *phi1, float *theta1, float *czD1, const int64_t ND2,
                                       float *phi2, float *theta2, float *czD2, const int numthreads,
                                       const int autocorr, const char *sbinfile, const double mu_max,
                                       const int nmu_bins, const int cosmology,
                                       results_countpairs_mocks_s_mu *results, struct config_options *options,
                                       struct extra_options *extra);
extern int countpairs_mocks_s_mu_double(const int64_t ND1, double *phi1, double *theta1, double *czD1,
                                        const int64_t ND2, double *phi2, double *theta2, double *czD2,
                                        const int numthreads, const int autocorr, const char *sbinfile,
                                        const double mu_max, const int nmu_bins, const int cosmology,
                                        results_countpairs_mocks_s_mu *results, struct config_options *options,
                                        struct extra_options *extra);

#ifdef __cplusplus
}
#endif
#endif
