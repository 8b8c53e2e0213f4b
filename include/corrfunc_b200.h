/* corrfunc_b200.h -- extensions of libcorrfunc_b200.so beyond the reference's C API.
 * None of these exist in the reference; they expose what a GPU build needs in addition:
 * multi-GPU work sharding (one process per GPU), per-call measurements, and tuning knobs.
 */
#ifndef CORRFUNC_B200_H
#define CORRFUNC_B200_H
#include <stdint.h>
#include "corrfunc_b200_device.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Multi-GPU: each process (rank) counts a disjoint share of the primary cells against a full replica
 * of the particles; the small per-bin histograms are then summed across ranks by `fn` (e.g. an NCCL
 * all-reduce issued through torch.distributed) BEFORE the host epilogue (x2, self pairs, averages,
 * xi/wp estimators) runs, so every rank returns the complete result.  Integer sums are exact, hence
 * npairs does not depend on the number of GPUs. fn returns 0 on success. */
typedef int (*corrfunc_b200_reduce_fn)(uint64_t *npairs, double *sum_sep, double *sum_w, int64_t nslots, void *user);
void corrfunc_b200_set_shard(int rank, int nranks);
void corrfunc_b200_set_reduce_hook(corrfunc_b200_reduce_fn fn, void *user);

/* Measurements of the most recent countpairs* call of this process. */
typedef struct {
    cfb_stats dev;        /* device-side timings and counters */
    double ms_host_total; /* wall time of the whole C call, host clock */
    double ms_upload;     /* wall time spent in cfb_upload (H2D) */
    int nmesh[3];         /* reference lattice */
    int refine[3];        /* bin refine factors after the reference's heuristics */
} corrfunc_b200_stats;
const corrfunc_b200_stats *corrfunc_b200_last_stats(void);

const char *corrfunc_b200_version(void);

/* DD / DR / RR in one context: between corrfunc_b200_catalog_cache(1) and (0), catalogues passed again under the same
 * pointers are uploaded once and sorted once per lattice (see cfb_set_catalog_cache).  The caller must not modify the
 * arrays in between. */
void corrfunc_b200_catalog_cache(int on);

/* cz (km/s) -> comoving distance (Mpc/h) exactly as countpairs_mocks / countpairs_mocks_s_mu do it for
 * is_comoving_dist == 0 (mocks/DDrppi_mocks/countpairs_rp_pi_mocks_impl.c.src:326-362): the redshift -> distance table
 * of utils/set_cosmo_dist.c:27-75 (Simpson's rule, 10000 points per unit redshift; plain C despite its GSL include),
 * then GSL's linear interpolation (interpolation/linear.c of GSL 2.x: y_lo + (x - x_lo) / dx * (y_hi - y_lo) on the
 * bracket found by bisection).  prec = 4 | 8 selects float / double arrays; cosmology = 1 (LasDamas) | 2 (Planck).
 * Returns 0, or 1 with a message on stderr (unknown cosmology, redshift outside the table's [1e-4, zmax] domain --
 * where GSL's error handler would abort the reference).  Pure host code: usable without a GPU. */
int corrfunc_b200_cz_to_comoving(int prec, int64_t n, const void *cz, int cosmology, void *dist);
/* The table itself (for tests): fills zc[], dc[] (max_size entries each) and returns the number of entries, -1 on error. */
int corrfunc_b200_cosmo_dist_table(double zmax, int max_size, double *zc, double *dc, int cosmology);
/* The first n uniform deviates of the MT19937 stream countspheres draws its sphere centres from (what GSL's
 * gsl_rng_mt19937 + gsl_rng_uniform give for this seed); for tests. */
void corrfunc_b200_mt19937_uniform(unsigned long seed, int64_t n, double *out);

#ifdef __cplusplus
}
#endif
#endif
