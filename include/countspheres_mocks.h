/* countspheres_mocks.h -- drop-in C ABI for counts-in-spheres / the void probability function on survey catalogues.
 * Replaces the reference interface mocks/vpf_mocks/countspheres_mocks.h:19-41 (Corrfunc v2.5.3): same symbol names,
 * argument order/meaning, result layout and error behaviour (EXIT_SUCCESS / EXIT_FAILURE + stderr message).
 * Inputs are HOST pointers of element size options->float_type (4 or 8): RA, DEC in degrees and cz in km/s (or the
 * comoving distance with options->is_comoving_dist = 1), for the galaxies and for the randoms.  Sphere centres come
 * from `centers_file` when it holds at least `nc` centres of radius >= rmax ("x y z r" per line, in the shifted frame
 * the reference writes); otherwise they are the first `nc` randoms with more than `threshold_neighbors` randoms within
 * rmax, and the file is rewritten with them -- exactly the reference's behaviour.  pN[ibin][i] = fraction of spheres of
 * radius (ibin+1)*rmax/nbin holding exactly i galaxies.  The counting runs on the GPU (sm_100a); no CPU fallback.
 */
#ifndef CORRFUNC_B200_COUNTSPHERES_MOCKS_H
#define CORRFUNC_B200_COUNTSPHERES_MOCKS_H
#include <stdint.h>
#include "corrfunc_b200_defs.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    double **pN; /* [nbin][num_pN]: one malloc'ed row per radius, released by free_results_countspheres_mocks */
    double rmax;
    int nbin;
    int nc;
    int num_pN;
} results_countspheres_mocks;

extern int countspheres_mocks(const int64_t Ngal, void *xgal /* RA */, void *ygal /* DEC */, void *zgal /* cz */,
                              const int64_t Nran, void *xran, void *yran, void *zran, const int threshold_neighbors,
                              const double rmax, const int nbin, const int nc, const int num_pN, const char *centers_file,
                              const int cosmology, results_countspheres_mocks *results, struct config_options *options,
                              struct extra_options *extra);
extern void free_results_countspheres_mocks(results_countspheres_mocks *results);

#ifdef __cplusplus
}
#endif
#endif
