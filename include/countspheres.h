/* countspheres.h -- drop-in C ABI for counts-in-spheres / the void probability function in a simulation box.
 * Replaces the reference interface theory/vpf/countspheres.h:19-41 (Corrfunc v2.5.3): same symbol names, argument
 * order/meaning, result layout and error behaviour.  X, Y, Z are HOST pointers of element size options->float_type.
 * `nc` sphere centres are drawn uniformly over the box (options->periodic with boxsize) or over the particle extent,
 * rejecting spheres that reach past it, from MT19937 seeded with `seed` -- the stream GSL's gsl_rng_mt19937 produces,
 * restated from the published algorithm because GSL is absent here (checked against numpy's independent MT19937).
 * pN[ibin][i] = fraction of spheres of radius (ibin+1)*rmax/nbin holding exactly i particles.  GPU only (sm_100a).
 */
#ifndef CORRFUNC_B200_COUNTSPHERES_H
#define CORRFUNC_B200_COUNTSPHERES_H
#include <stdint.h>
#include "corrfunc_b200_defs.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    double **pN; /* [nbin][num_pN]: one malloc'ed row per radius, released by free_results_countspheres */
    double rmax;
    int nbin;
    int nc;
    int num_pN;
} results_countspheres;

extern int countspheres(const int64_t np, void *X, void *Y, void *Z, const double rmax, const int nbin, const int nc,
                        const int num_pN, unsigned long seed, results_countspheres *results,
                        struct config_options *options, struct extra_options *extra);
extern void free_results_countspheres(results_countspheres *results);

#ifdef __cplusplus
}
#endif
#endif
