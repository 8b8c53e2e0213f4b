/* TEST INFRASTRUCTURE ONLY -- replaces utils/set_cosmo_dist.c + utils/set_cosmology.c (which integrate with
 * GSL, absent here) when oracle/build_ref.sh links the unmodified mocks sources.  Reached only with
 * options->is_comoving_dist == 0, which the parity tests never use. */
#include <stdio.h>
int set_cosmo_dist(const double zmax, const int max_size, double *zc, double *dc, const int lasdamas_cosmology)
{
    (void)zmax, (void)max_size, (void)zc, (void)dc, (void)lasdamas_cosmology;
    fprintf(stderr, "oracle/_ref: cz -> comoving distance needs GSL, which this image lacks; pass comoving distances "
                    "with is_comoving_dist = 1\n");
    return -1;
}
