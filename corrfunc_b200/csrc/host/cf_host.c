/* cf_host.c -- C host layer of libcorrfunc_b200: the reference's public C API, with everything the
 * reference decides on the host (bins, extents, wrap, refine heuristics, lattice sizes, epilogues)
 * kept on the host and everything it does per particle / per pair handed to the CUDA layer through
 * include/corrfunc_b200_device.h.  Precision-specific code lives in cf_host_impl.h, included twice.
 */
#define _GNU_SOURCE
#include <float.h>
#include <inttypes.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <ctype.h>

#include "corrfunc_b200.h"
#include "corrfunc_b200_defs.h"
#include "corrfunc_b200_device.h"
#include "countpairs.h"
#include "countpairs_rp_pi.h"
#include "countpairs_rp_pi_mocks.h"
#include "countspheres_mocks.h"
#include "countspheres.h"
#include "countpairs_s_mu.h"
#include "countpairs_s_mu_mocks.h"
#include "countpairs_theta_mocks.h"
#include "countpairs_wp.h"
#include "countpairs_xi.h"

#define CF_PI_OVER_180 0.017453292519943295769236907684886127134428718885417254560971
#define CF_INV_PI_OVER_180 57.29577951308232087679815481410517033240547246656432154916024

static corrfunc_b200_stats g_stats;
static corrfunc_b200_reduce_fn g_reduce = NULL;
static void *g_reduce_user = NULL;

const corrfunc_b200_stats *corrfunc_b200_last_stats(void) { return &g_stats; }
const char *corrfunc_b200_version(void) { return "corrfunc_b200 0.1.0 (API " CORRFUNC_API_VERSION ")"; }
void corrfunc_b200_set_shard(int rank, int nranks) { cfb_set_shard(rank, nranks); }
void corrfunc_b200_catalog_cache(int on) { cfb_set_catalog_cache(on); }
/* utils/cpu_features.c:19-120: nothing to detect here, the kernels run on the GPU whatever the host CPU is */
int runtime_instrset_detect(void) { return AVX512F; }
int get_max_usable_isa(void) { return AVX512F; }
void corrfunc_b200_set_reduce_hook(corrfunc_b200_reduce_fn fn, void *user)
{
    g_reduce = fn;
    g_reduce_user = user;
}

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* The reference stringifies a quoted macro, so its own version string is `"2.5.3"` including the
 * quote characters (utils/defs.h:27 + common.mk:175); accept both spellings. */
static int version_ok(const struct config_options *o)
{
    const char *v = o->version;
    if (v[0] == '"') v++;
    return strncmp(v, CORRFUNC_API_VERSION, strlen(CORRFUNC_API_VERSION)) == 0;
}

/* Bin file parser: "lo hi" per line, '#' comments; nbin = lines+1, rupp[0] = first low edge
 * (utils/utils.c:62-106 and :715-745). Returns malloc'ed rupp[nbin+1]. */
static int64_t count_data_lines(const char *fname)
{
    FILE *fp = fopen(fname, "rt");
    if (!fp) {
        fprintf(stderr, "Error: could not open bin file `%s'\n", fname);
        return -1;
    }
    char line[10000];
    int64_t n = 0;
    while (fgets(line, sizeof(line), fp)) {
        const char *c = line;
        while (*c != '\0' && isspace((unsigned char)*c)) c++;
        if (*c != '\0' && *c != '#') n++;
    }
    fclose(fp);
    return n;
}

static int cf_setup_bins(const char *fname, double *rmin, double *rmax, int *nbin, double **rupp, const int as_float)
{
    const int64_t nl = count_data_lines(fname);
    if (nl < 0) return EXIT_FAILURE;
    *nbin = (int)nl + 1;
    *rupp = calloc((size_t)*nbin + 1, sizeof(double));
    if (!*rupp) return EXIT_FAILURE;
    FILE *fp = fopen(fname, "r");
    if (!fp) {
        free(*rupp);
        *rupp = NULL;
        return EXIT_FAILURE;
    }
    char buf[1000];
    int index = 1;
    *rmin = 0.0;
    while (fgets(buf, sizeof(buf), fp)) {
        double lo, hi;
        int nread;
        if (as_float) { /* setup_bins_float parses with %f (utils/utils.c:154-196) */
            float flo, fhi;
            nread = sscanf(buf, "%f %f", &flo, &fhi);
            lo = flo;
            hi = fhi;
        } else {
            nread = sscanf(buf, "%lf %lf", &lo, &hi);
        }
        if (nread == 2 && index <= *nbin) {
            if (index == 1) {
                *rmin = lo;
                (*rupp)[0] = lo;
            }
            (*rupp)[index++] = hi;
        }
    }
    fclose(fp);
    *rmax = (*rupp)[index - 1];
    (*rupp)[*nbin] = *rmax;
    (*rupp)[*nbin - 1] = *rmax;
    return EXIT_SUCCESS;
}

/* ---- interrupts (utils/macros.h:145-167; theory/DD/countpairs_impl.c.src:31-37, 193, 475-477, 554-569, 695) ----
 * The reference installs handlers for SIGTERM, SIGINT and SIGHUP for the duration of a call ("mostly useful during the
 * python execution"), its loop over cell pairs polls the flag they set, and an interrupted call returns EXIT_FAILURE.
 * Here the flag lives in mapped pinned memory (cfb_abort_flag): the persistent pair kernels read it at every tile fetch. */
#include <signal.h>
typedef void (*cf_sig_t)(int);
static const int cf_signals[3] = {SIGTERM, SIGINT, SIGHUP};
static volatile sig_atomic_t cf_interrupt_status = 0;
static volatile int *cf_abort = NULL;
static void cf_interrupt_handler(int signo)
{
    fprintf(stderr, "Received signal = `%s' (signo = %d). Aborting \n", strsignal(signo), signo);
    cf_interrupt_status = 1;
    if (cf_abort) *cf_abort = 1;
}
static void cf_signals_install(cf_sig_t prev[3])
{
    cf_abort = cfb_abort_flag();
    *cf_abort = 0;
    cf_interrupt_status = 0;
    for (int i = 0; i < 3; i++) {
        prev[i] = signal(cf_signals[i], cf_interrupt_handler);
        if (prev[i] == SIG_ERR) fprintf(stderr, "Can not handle signal = %d\n", cf_signals[i]);
    }
}
/* returns 1 when the call was interrupted */
static int cf_signals_restore(const cf_sig_t prev[3])
{
    for (int i = 0; i < 3; i++) {
        if (prev[i] == SIG_IGN || prev[i] == SIG_ERR) continue;
        if (signal(cf_signals[i], prev[i]) == SIG_ERR)
            fprintf(stderr, "Could not reset signal handler to default for signal = %d\n", cf_signals[i]);
    }
    const int hit = cf_interrupt_status != 0;
    if (cf_abort) *cf_abort = 0;
    cf_interrupt_status = 0;
    return hit;
}

/* raw device histograms -> optional cross-rank sum */
/* local_status: what this rank's device count returned.  Every rank enters the collective whatever it returned -- a rank
 * that failed locally (out of memory, a particle outside the box of its replica) must not leave the others waiting in
 * the all-reduce for ever -- and announces the failure in slot 0 of the counts, which no statistic ever writes (bins start
 * at 1); afterwards all ranks fail together. */
static int reduce_across_ranks(int local_status, uint64_t *np, double *ss, double *sw, int64_t nslots)
{
    int rank = 0, nranks = 1;
    cfb_get_shard(&rank, &nranks);
    if (nranks <= 1) return local_status;
    if (!g_reduce) {
        fprintf(stderr, "corrfunc_b200> work is sharded over %d ranks but no reduce hook is set; "
                        "results would be partial\n", nranks);
        return 1;
    }
    const uint64_t failed_mark = (uint64_t)1 << 62;
    if (nslots > 0) np[0] = local_status ? failed_mark : 0;
    const int rc = g_reduce(np, ss, sw, nslots, g_reduce_user);
    const int any_failed = nslots > 0 && np[0] >= failed_mark;
    if (nslots > 0) np[0] = 0;
    if (any_failed && !local_status) fprintf(stderr, "corrfunc_b200> another rank failed; this rank's result is discarded too\n");
    return (rc || any_failed || local_status) ? 1 : 0;
}

/* what every statistic hands back to its public wrapper (arrays are malloc'ed, caller owns them) */
/* ---- redshift -> comoving distance (mocks statistics with is_comoving_dist == 0) ---------------------------- */
#define CF_SPEED_OF_LIGHT 299800.0 /* utils/set_cosmo_dist.h:19 */

/* utils/set_cosmo_dist.c:27-75 with the parameters of utils/cosmology_params.c:21-54; same statements in the same
 * order (compiled with -ffp-contract=off like the reference's -std=c99), so the table is bit-identical */
int corrfunc_b200_cosmo_dist_table(double zmax, int max_size, double *zc, double *dc, int cosmology)
{
    double OMEGA_M;
    switch (cosmology) {
    case 1: OMEGA_M = 0.25; break;
    case 2: OMEGA_M = 0.302; break;
    default: fprintf(stderr, "ERROR: In %s> Cosmology=%d not implemented\n", "init_cosmology", cosmology); return -1;
    }
    const double OMEGA_L = 1.0 - OMEGA_M;
    int i = 0;
    const double smallh = 1.0;
    const double Omegak = 1.0 - OMEGA_M - OMEGA_L;
    const double Dh = CF_SPEED_OF_LIGHT * 0.01 / smallh;
    const double Deltaz = 1.0 / max_size;
    const double dz = 1e-2 * Deltaz;
    const double epsilon = 1e-10;
    double Eint = 0.0, E2 = 1.0, z2 = Deltaz;
#define CF_CUBE(x) ((x) * (x) * (x))
    for (double z = 2.0 * dz; z < zmax; z += 2.0 * dz) {
        const double E0 = E2;
        const double E1 = 1.0 / sqrt(OMEGA_M * CF_CUBE(1 + z - dz) + Omegak * (1 + z - dz) + OMEGA_L);
        E2 = 1.0 / sqrt(OMEGA_M * CF_CUBE(1 + z) + Omegak * (1 + z) + OMEGA_L);
        Eint += dz * (E0 + 4. * E1 + E2) / 3.;
        if (z > (z2 - epsilon) && z < (z2 + epsilon)) {
            if (i >= max_size) break;
            zc[i] = z;
            dc[i] = Eint * Dh;
            z2 += Deltaz;
            i++;
        }
    }
#undef CF_CUBE
    return i;
}

/* GSL 2.x interpolation/linear.c:linear_eval + interpolation/bsearch.c:gsl_interp_bsearch (GSL is a dependency of the
 * reference that is absent here; restated from its published source).  Returns 1 when x is outside the table. */
static int cf_interp_linear(const double *xa, const double *ya, const int n, const double x, double *y)
{
    if (n < 2 || x < xa[0] || x > xa[n - 1]) return 1; /* gsl_interp_eval: GSL_EDOM -> the default handler aborts */
    size_t ilo = 0, ihi = (size_t)n - 1;
    while (ihi > ilo + 1) {
        const size_t i = (ihi + ilo) / 2;
        if (xa[i] > x) ihi = i; else ilo = i;
    }
    const double x_lo = xa[ilo], x_hi = xa[ilo + 1], y_lo = ya[ilo], y_hi = ya[ilo + 1];
    const double dx = x_hi - x_lo;
    if (!(dx > 0.0)) return 1;
    *y = y_lo + (x - x_lo) / dx * (y_hi - y_lo);
    return 0;
}

typedef struct {
    int nbin;         /* number of edges = reference's nbin */
    int n2;           /* npibin / nmu_bins for the 2-D statistics */
    uint64_t *npairs; /* [nslots] */
    double *rupp;     /* [nbin] */
    double *avg;      /* [nslots] */
    double *wavg;     /* [nslots] */
    double *cf;       /* [nbin] xi or wp, NULL otherwise */
} cf_box_out;

/* ---- MT19937 as GSL's gsl_rng_mt19937 runs it (theory/vpf draws its sphere centres from it, countspheres_impl.c.src:
 * 190-192, 303-305).  GSL is absent here; this is the published generator of Matsumoto & Nishimura with the 2002
 * initialisation that GSL 2.x's rng/mt.c implements (seed 0 -> 4357), gsl_rng_uniform = next word / 2^32. */
typedef struct {
    uint32_t mt[624];
    int mti;
} cf_mt19937;

static void cf_mt_set(cf_mt19937 *r, unsigned long seed)
{
    uint32_t s = (uint32_t)(seed & 0xffffffffUL);
    if (seed == 0) s = 4357u;
    r->mt[0] = s;
    for (int i = 1; i < 624; i++) r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
    r->mti = 624;
}

static double cf_mt_uniform(cf_mt19937 *r)
{
    uint32_t *const mt = r->mt;
    if (r->mti >= 624) {
        for (int kk = 0; kk < 624; kk++) {
            const uint32_t y = (mt[kk] & 0x80000000u) | (mt[(kk + 1) % 624] & 0x7fffffffu);
            mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        r->mti = 0;
    }
    uint32_t k = mt[r->mti++];
    k ^= (k >> 11);
    k ^= (k << 7) & 0x9d2c5680u;
    k ^= (k << 15) & 0xefc60000u;
    k ^= (k >> 18);
    return k / 4294967296.0;
}

/* the first n uniforms of that stream (for tests) */
void corrfunc_b200_mt19937_uniform(unsigned long seed, int64_t n, double *out)
{
    cf_mt19937 r;
    cf_mt_set(&r, seed);
    for (int64_t i = 0; i < n; i++) out[i] = cf_mt_uniform(&r);
}

typedef struct { /* counts-in-spheres result (cf_vpf_mocks) */
    int nbin, num_pN;
    double **pN;
} cf_vpf_out;

#define REAL float
#define SFX f32
#define REAL_IS_DOUBLE 0
#include "cf_host_impl.h"
#undef REAL
#undef SFX
#undef REAL_IS_DOUBLE

#define REAL double
#define SFX f64
#define REAL_IS_DOUBLE 1
#include "cf_host_impl.h"
#undef REAL
#undef SFX
#undef REAL_IS_DOUBLE

/* ------------------------------------------------------------------------------------------ */
/* public entry points: float/double dispatch like theory/DD/countpairs.c:34-77                 */

static int check_common(const struct config_options *options, const char *fn)
{
    if (options == NULL) {
        fprintf(stderr, "Error: In %s> options can not be NULL\n", fn);
        return EXIT_FAILURE;
    }
    if (!(options->float_type == sizeof(float) || options->float_type == sizeof(double))) {
        fprintf(stderr, "ERROR: In %s> Can only handle doubles or floats. Got an array of size = %zu\n", fn,
                options->float_type);
        return EXIT_FAILURE;
    }
    if (!version_ok(options)) {
        fprintf(stderr, "Error: Do not know this API version = `%s'. Expected version = `%s'\n", options->version,
                CORRFUNC_API_VERSION);
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}

void free_results(results_countpairs *r)
{
    if (!r) return;
    free(r->npairs); free(r->rupp); free(r->rpavg); free(r->weightavg);
    r->npairs = NULL; r->rupp = NULL; r->rpavg = NULL; r->weightavg = NULL;
}
void free_results_rp_pi(results_countpairs_rp_pi *r)
{
    if (!r) return;
    free(r->npairs); free(r->rupp); free(r->rpavg); free(r->weightavg);
    r->npairs = NULL; r->rupp = NULL; r->rpavg = NULL; r->weightavg = NULL;
}
void free_results_s_mu(results_countpairs_s_mu *r)
{
    if (!r) return;
    free(r->npairs); free(r->supp); free(r->savg); free(r->weightavg);
    r->npairs = NULL; r->supp = NULL; r->savg = NULL; r->weightavg = NULL;
}
void free_results_wp(results_countpairs_wp *r)
{
    if (!r) return;
    free(r->npairs); free(r->rupp); free(r->rpavg); free(r->weightavg); free(r->wp);
    r->npairs = NULL; r->rupp = NULL; r->rpavg = NULL; r->weightavg = NULL; r->wp = NULL;
}
void free_results_xi(results_countpairs_xi *r)
{
    if (!r) return;
    free(r->npairs); free(r->rupp); free(r->ravg); free(r->weightavg); free(r->xi);
    r->npairs = NULL; r->rupp = NULL; r->ravg = NULL; r->weightavg = NULL; r->xi = NULL;
}
void free_results_countpairs_theta(results_countpairs_theta *r)
{
    if (!r) return;
    free(r->npairs); free(r->theta_upp); free(r->theta_avg); free(r->weightavg);
    r->npairs = NULL; r->theta_upp = NULL; r->theta_avg = NULL; r->weightavg = NULL;
}

int countpairs(const int64_t ND1, void *X1, void *Y1, void *Z1, const int64_t ND2, void *X2, void *Y2, void *Z2,
               const int numthreads, const int autocorr, const char *binfile, results_countpairs *results,
               struct config_options *options, struct extra_options *extra)
{
    if (check_common(options, __func__)) return EXIT_FAILURE;
    cf_box_out o;
    memset(&o, 0, sizeof(o));
    const int st = options->float_type == sizeof(float)
                       ? cf_box_f32(CFB_DD, ND1, X1, Y1, Z1, ND2, X2, Y2, Z2, numthreads, autocorr, binfile, 0.0, 0.0,
                                    0, 0.0, options, extra, &o)
                       : cf_box_f64(CFB_DD, ND1, X1, Y1, Z1, ND2, X2, Y2, Z2, numthreads, autocorr, binfile, 0.0, 0.0,
                                    0, 0.0, options, extra, &o);
    if (st != EXIT_SUCCESS) return st;
    results->nbin = o.nbin;
    results->npairs = o.npairs;
    results->rupp = o.rupp;
    results->rpavg = o.avg;
    results->weightavg = o.wavg;
    free(o.cf);
    return EXIT_SUCCESS;
}

int countpairs_rp_pi(const int64_t ND1, void *X1, void *Y1, void *Z1, const int64_t ND2, void *X2, void *Y2, void *Z2,
                     const int numthreads, const int autocorr, const char *binfile, const double pimax,
                     results_countpairs_rp_pi *results, struct config_options *options, struct extra_options *extra)
{
    if (check_common(options, __func__)) return EXIT_FAILURE;
    cf_box_out o;
    memset(&o, 0, sizeof(o));
    const int st = options->float_type == sizeof(float)
                       ? cf_box_f32(CFB_RPPI, ND1, X1, Y1, Z1, ND2, X2, Y2, Z2, numthreads, autocorr, binfile, pimax,
                                    0.0, 0, 0.0, options, extra, &o)
                       : cf_box_f64(CFB_RPPI, ND1, X1, Y1, Z1, ND2, X2, Y2, Z2, numthreads, autocorr, binfile, pimax,
                                    0.0, 0, 0.0, options, extra, &o);
    if (st != EXIT_SUCCESS) return st;
    results->nbin = o.nbin;
    results->npibin = o.n2;
    results->pimax = pimax;
    results->npairs = o.npairs;
    results->rupp = o.rupp;
    results->rpavg = o.avg;
    results->weightavg = o.wavg;
    free(o.cf);
    return EXIT_SUCCESS;
}

int countpairs_s_mu(const int64_t ND1, void *X1, void *Y1, void *Z1, const int64_t ND2, void *X2, void *Y2, void *Z2,
                    const int numthreads, const int autocorr, const char *sbinfile, const double mu_max,
                    const int nmu_bins, results_countpairs_s_mu *results, struct config_options *options,
                    struct extra_options *extra)
{
    if (check_common(options, __func__)) return EXIT_FAILURE;
    cf_box_out o;
    memset(&o, 0, sizeof(o));
    const int st = options->float_type == sizeof(float)
                       ? cf_box_f32(CFB_SMU, ND1, X1, Y1, Z1, ND2, X2, Y2, Z2, numthreads, autocorr, sbinfile, 0.0,
                                    mu_max, nmu_bins, 0.0, options, extra, &o)
                       : cf_box_f64(CFB_SMU, ND1, X1, Y1, Z1, ND2, X2, Y2, Z2, numthreads, autocorr, sbinfile, 0.0,
                                    mu_max, nmu_bins, 0.0, options, extra, &o);
    if (st != EXIT_SUCCESS) return st;
    results->nsbin = o.nbin;
    results->nmu_bins = nmu_bins;
    results->mu_max = mu_max; /* the double the caller passed (countpairs_s_mu_impl.c.src:670) */
    results->mu_min = 0.0;
    results->npairs = o.npairs;
    results->supp = o.rupp;
    results->savg = o.avg;
    results->weightavg = o.wavg;
    free(o.cf);
    return EXIT_SUCCESS;
}

int countpairs_wp(const int64_t ND, void *X, void *Y, void *Z, const double boxsize, const int numthreads,
                  const char *binfile, const double pimax, results_countpairs_wp *results,
                  struct config_options *options, struct extra_options *extra)
{
    if (check_common(options, __func__)) return EXIT_FAILURE;
    cf_box_out o;
    memset(&o, 0, sizeof(o));
    const int st = options->float_type == sizeof(float)
                       ? cf_box_f32(CFB_WP, ND, X, Y, Z, 0, NULL, NULL, NULL, numthreads, 1, binfile, pimax, 0.0, 0,
                                    boxsize, options, extra, &o)
                       : cf_box_f64(CFB_WP, ND, X, Y, Z, 0, NULL, NULL, NULL, numthreads, 1, binfile, pimax, 0.0, 0,
                                    boxsize, options, extra, &o);
    if (st != EXIT_SUCCESS) return st;
    results->nbin = o.nbin;
    results->pimax = pimax;
    results->npairs = o.npairs;
    results->rupp = o.rupp;
    results->rpavg = o.avg;
    results->weightavg = o.wavg;
    results->wp = o.cf;
    return EXIT_SUCCESS;
}

int countpairs_xi(const int64_t ND, void *X, void *Y, void *Z, const double boxsize, const int numthreads,
                  const char *binfile, results_countpairs_xi *results, struct config_options *options,
                  struct extra_options *extra)
{
    if (check_common(options, __func__)) return EXIT_FAILURE;
    cf_box_out o;
    memset(&o, 0, sizeof(o));
    const int st = options->float_type == sizeof(float)
                       ? cf_box_f32(CFB_XI, ND, X, Y, Z, 0, NULL, NULL, NULL, numthreads, 1, binfile, 0.0, 0.0, 0,
                                    boxsize, options, extra, &o)
                       : cf_box_f64(CFB_XI, ND, X, Y, Z, 0, NULL, NULL, NULL, numthreads, 1, binfile, 0.0, 0.0, 0,
                                    boxsize, options, extra, &o);
    if (st != EXIT_SUCCESS) return st;
    results->nbin = o.nbin;
    results->npairs = o.npairs;
    results->rupp = o.rupp;
    results->ravg = o.avg;
    results->weightavg = o.wavg;
    results->xi = o.cf;
    return EXIT_SUCCESS;
}

int countpairs_theta_mocks(const int64_t ND1, void *phi1, void *theta1, const int64_t ND2, void *phi2, void *theta2,
                           const int numthreads, const int autocorr, const char *binfile,
                           results_countpairs_theta *results, struct config_options *options,
                           struct extra_options *extra)
{
    if (ND1 == 0 || ND2 == 0) { /* countpairs_theta_mocks_impl.c.src:463-476 */
        fprintf(stderr, "Warning: Received 0 particles in at least one of the arrays. len(array1) = %" PRId64
                        " len(array2) = %" PRId64 "\n", ND1, ND2);
        if (results != NULL) {
            results->npairs = NULL;
            results->theta_avg = NULL;
            results->weightavg = NULL;
            results->theta_upp = NULL;
        }
        return EXIT_SUCCESS;
    }
    if (check_common(options, __func__)) return EXIT_FAILURE;
    cf_box_out o;
    memset(&o, 0, sizeof(o));
    const int st = options->float_type == sizeof(float)
                       ? cf_theta_f32(ND1, phi1, theta1, ND2, phi2, theta2, numthreads, autocorr, binfile, options,
                                      extra, &o)
                       : cf_theta_f64(ND1, phi1, theta1, ND2, phi2, theta2, numthreads, autocorr, binfile, options,
                                      extra, &o);
    if (st != EXIT_SUCCESS) return st;
    results->nbin = o.nbin;
    results->npairs = o.npairs;
    results->theta_upp = o.rupp;
    results->theta_avg = o.avg;
    results->weightavg = o.wavg;
    free(o.cf);
    return EXIT_SUCCESS;
}

/* ------------------------------------------------------------------------------------------------
 * Precision-suffixed entry points (the reference's *_impl.h.src prototypes): typed pointers; the
 * options' float_type is overridden for the call and restored afterwards. */
int corrfunc_b200_cz_to_comoving(int prec, int64_t n, const void *cz, int cosmology, void *dist)
{
    if (prec == 4) {
        float czmax = 0.0f;
        for (int64_t i = 0; i < n; i++)
            if (((const float *)cz)[i] > czmax) czmax = ((const float *)cz)[i];
        return cf_cz_to_dist_f32(n, (const float *)cz, czmax, cosmology, (float *)dist);
    }
    if (prec == 8) {
        double czmax = 0.0;
        for (int64_t i = 0; i < n; i++)
            if (((const double *)cz)[i] > czmax) czmax = ((const double *)cz)[i];
        return cf_cz_to_dist_f64(n, (const double *)cz, czmax, cosmology, (double *)dist);
    }
    fprintf(stderr, "Error: In %s> element size must be 4 or 8 (got %d)\n", __func__, prec);
    return EXIT_FAILURE;
}

/* ---- survey geometry: mocks/DDrppi_mocks/countpairs_rp_pi_mocks.c:33-77, DDsmu_mocks ---- */
void free_results_mocks(results_countpairs_mocks *r)
{
    if (r == NULL) return;
    free(r->npairs); free(r->rupp); free(r->rpavg); free(r->weightavg);
    r->npairs = NULL; r->rupp = NULL; r->rpavg = NULL; r->weightavg = NULL;
}

void free_results_mocks_s_mu(results_countpairs_mocks_s_mu *r)
{
    if (r == NULL) return;
    free(r->npairs); free(r->supp); free(r->savg); free(r->weightavg);
    r->npairs = NULL; r->supp = NULL; r->savg = NULL; r->weightavg = NULL;
}

int countpairs_mocks(const int64_t ND1, void *phi1, void *theta1, void *czD1, const int64_t ND2, void *phi2, void *theta2,
                     void *czD2, const int numthreads, const int autocorr, const char *binfile, const double pimax,
                     const int cosmology, results_countpairs_mocks *results, struct config_options *options,
                     struct extra_options *extra)
{
    if (check_common(options, __func__)) return EXIT_FAILURE;
    cf_box_out o;
    memset(&o, 0, sizeof(o));
    const int st = options->float_type == sizeof(float)
                       ? cf_mocks_f32(CFB_RPPI_MOCKS, ND1, phi1, theta1, czD1, ND2, phi2, theta2, czD2, numthreads, autocorr,
                                      binfile, (double)(float)pimax, 0.0, 0, cosmology, options, extra, &o)
                       : cf_mocks_f64(CFB_RPPI_MOCKS, ND1, phi1, theta1, czD1, ND2, phi2, theta2, czD2, numthreads, autocorr,
                                      binfile, pimax, 0.0, 0, cosmology, options, extra, &o);
    if (st != EXIT_SUCCESS || o.npairs == NULL) return st; /* empty input: success, results untouched */
    results->nbin = o.nbin;
    results->npibin = o.n2;
    results->pimax = options->float_type == sizeof(float) ? (double)(float)pimax : pimax;
    results->npairs = o.npairs;
    results->rupp = o.rupp;
    results->rpavg = o.avg;
    results->weightavg = o.wavg;
    free(o.cf);
    return EXIT_SUCCESS;
}

int countpairs_mocks_s_mu(const int64_t ND1, void *phi1, void *theta1, void *czD1, const int64_t ND2, void *phi2,
                          void *theta2, void *czD2, const int numthreads, const int autocorr, const char *sbinfile,
                          const double mu_max, const int nmu_bins, const int cosmology,
                          results_countpairs_mocks_s_mu *results, struct config_options *options,
                          struct extra_options *extra)
{
    if (check_common(options, __func__)) return EXIT_FAILURE;
    cf_box_out o;
    memset(&o, 0, sizeof(o));
    const int st = options->float_type == sizeof(float)
                       ? cf_mocks_f32(CFB_SMU_MOCKS, ND1, phi1, theta1, czD1, ND2, phi2, theta2, czD2, numthreads, autocorr,
                                      sbinfile, 0.0, mu_max, nmu_bins, cosmology, options, extra, &o)
                       : cf_mocks_f64(CFB_SMU_MOCKS, ND1, phi1, theta1, czD1, ND2, phi2, theta2, czD2, numthreads, autocorr,
                                      sbinfile, 0.0, mu_max, nmu_bins, cosmology, options, extra, &o);
    if (st != EXIT_SUCCESS || o.npairs == NULL) return st;
    results->nsbin = o.nbin;
    results->nmu_bins = nmu_bins;
    results->mu_max = mu_max;
    results->mu_min = 0.0;
    results->npairs = o.npairs;
    results->supp = o.rupp;
    results->savg = o.avg;
    results->weightavg = o.wavg;
    free(o.cf);
    return EXIT_SUCCESS;
}

/* ---- counts-in-spheres in a simulation box: theory/vpf/countspheres.c ---- */
void free_results_countspheres(results_countspheres *r)
{
    if (r == NULL || r->pN == NULL) return;
    for (int i = 0; i < r->nbin; i++) free(r->pN[i]);
    free(r->pN);
    r->pN = NULL;
}

int countspheres(const int64_t np, void *X, void *Y, void *Z, const double rmax, const int nbin, const int nc, const int num_pN,
                 unsigned long seed, results_countspheres *results, struct config_options *options, struct extra_options *extra)
{
    (void)extra;
    if (check_common(options, __func__)) return EXIT_FAILURE;
    cf_vpf_out o;
    memset(&o, 0, sizeof(o));
    const int st = options->float_type == sizeof(float)
                       ? cf_vpf_theory_f32(np, X, Y, Z, rmax, nbin, nc, num_pN, seed, options, &o)
                       : cf_vpf_theory_f64(np, X, Y, Z, rmax, nbin, nc, num_pN, seed, options, &o);
    if (st != EXIT_SUCCESS) {
        if (o.pN) {
            for (int i = 0; i < nbin; i++) free(o.pN[i]);
            free(o.pN);
        }
        return st;
    }
    results->rmax = rmax;
    results->nbin = nbin;
    results->nc = nc;
    results->num_pN = num_pN;
    results->pN = o.pN;
    return EXIT_SUCCESS;
}

/* ---- counts-in-spheres: mocks/vpf_mocks/countspheres_mocks.c ---- */
void free_results_countspheres_mocks(results_countspheres_mocks *r)
{
    if (r == NULL || r->pN == NULL) return;
    for (int i = 0; i < r->nbin; i++) free(r->pN[i]);
    free(r->pN);
    r->pN = NULL;
}

int countspheres_mocks(const int64_t Ngal, void *xgal, void *ygal, void *zgal, const int64_t Nran, void *xran, void *yran,
                       void *zran, const int threshold_neighbors, const double rmax, const int nbin, const int nc,
                       const int num_pN, const char *centers_file, const int cosmology, results_countspheres_mocks *results,
                       struct config_options *options, struct extra_options *extra)
{
    (void)extra;
    if (check_common(options, __func__)) return EXIT_FAILURE;
    cf_vpf_out o;
    memset(&o, 0, sizeof(o));
    const int st = options->float_type == sizeof(float)
                       ? cf_vpf_mocks_f32(Ngal, xgal, ygal, zgal, Nran, xran, yran, zran, threshold_neighbors, (float)rmax, nbin, nc,
                                          num_pN, centers_file, cosmology, options, &o)
                       : cf_vpf_mocks_f64(Ngal, xgal, ygal, zgal, Nran, xran, yran, zran, threshold_neighbors, rmax, nbin, nc,
                                          num_pN, centers_file, cosmology, options, &o);
    if (st != EXIT_SUCCESS) {
        if (o.pN) {
            for (int i = 0; i < nbin; i++) free(o.pN[i]);
            free(o.pN);
        }
        return st;
    }
    results->rmax = rmax;
    results->nbin = nbin;
    results->nc = nc;
    results->num_pN = num_pN;
    results->pN = o.pN;
    return EXIT_SUCCESS;
}

#define CF_TYPED(call)                                                      \
    do {                                                                    \
        if (!options) {                                                     \
            fprintf(stderr, "Error: In %s> options can not be NULL\n", __func__); \
            return EXIT_FAILURE;                                            \
        }                                                                   \
        const size_t saved = options->float_type;                           \
        options->float_type = sizeof(*X1);                                  \
        const int st = (call);                                              \
        options->float_type = saved;                                        \
        return st;                                                          \
    } while (0)

#define CF_TYPED_FUNCS(SUF, REAL)                                                                                        \
    int countpairs_##SUF(const int64_t ND1, REAL *X1, REAL *Y1, REAL *Z1, const int64_t ND2, REAL *X2, REAL *Y2,         \
                         REAL *Z2, const int numthreads, const int autocorr, const char *binfile,                       \
                         results_countpairs *results, struct config_options *options, struct extra_options *extra)      \
    {                                                                                                                    \
        CF_TYPED(countpairs(ND1, X1, Y1, Z1, ND2, X2, Y2, Z2, numthreads, autocorr, binfile, results, options, extra));  \
    }                                                                                                                    \
    int countpairs_rp_pi_##SUF(const int64_t ND1, REAL *X1, REAL *Y1, REAL *Z1, const int64_t ND2, REAL *X2, REAL *Y2,   \
                               REAL *Z2, const int numthreads, const int autocorr, const char *binfile,                 \
                               const double pimax, results_countpairs_rp_pi *results, struct config_options *options,   \
                               struct extra_options *extra)                                                              \
    {                                                                                                                    \
        CF_TYPED(countpairs_rp_pi(ND1, X1, Y1, Z1, ND2, X2, Y2, Z2, numthreads, autocorr, binfile, pimax, results,       \
                                  options, extra));                                                                      \
    }                                                                                                                    \
    int countpairs_s_mu_##SUF(const int64_t ND1, REAL *X1, REAL *Y1, REAL *Z1, const int64_t ND2, REAL *X2, REAL *Y2,    \
                              REAL *Z2, const int numthreads, const int autocorr, const char *sbinfile,                 \
                              const double mu_max, const int nmu_bins, results_countpairs_s_mu *results,                \
                              struct config_options *options, struct extra_options *extra)                              \
    {                                                                                                                    \
        CF_TYPED(countpairs_s_mu(ND1, X1, Y1, Z1, ND2, X2, Y2, Z2, numthreads, autocorr, sbinfile, mu_max, nmu_bins,     \
                                 results, options, extra));                                                              \
    }                                                                                                                    \
    int countpairs_wp_##SUF(const int64_t ND1, REAL *X1, REAL *Y1, REAL *Z1, const double boxsize, const int numthreads, \
                            const char *binfile, const double pimax, results_countpairs_wp *result,                     \
                            struct config_options *options, struct extra_options *extra)                                \
    {                                                                                                                    \
        CF_TYPED(countpairs_wp(ND1, X1, Y1, Z1, boxsize, numthreads, binfile, pimax, result, options, extra));           \
    }                                                                                                                    \
    int countpairs_xi_##SUF(const int64_t ND1, REAL *X1, REAL *Y1, REAL *Z1, const double boxsize, const int numthreads, \
                            const char *binfile, results_countpairs_xi *results, struct config_options *options,       \
                            struct extra_options *extra)                                                                 \
    {                                                                                                                    \
        CF_TYPED(countpairs_xi(ND1, X1, Y1, Z1, boxsize, numthreads, binfile, results, options, extra));                 \
    }                                                                                                                    \
    int countpairs_theta_mocks_##SUF(const int64_t ND1, REAL *X1, REAL *theta1, const int64_t ND2, REAL *phi2,           \
                                     REAL *theta2, const int numthreads, const int autocorr, const char *binfile,       \
                                     results_countpairs_theta *results, struct config_options *options,                 \
                                     struct extra_options *extra)                                                        \
    {                                                                                                                    \
        CF_TYPED(countpairs_theta_mocks(ND1, X1, theta1, ND2, phi2, theta2, numthreads, autocorr, binfile, results,      \
                                        options, extra));                                                                \
    }

CF_TYPED_FUNCS(float, float)
CF_TYPED_FUNCS(double, double)

#define CF_TYPED_MOCKS(SUF, REAL)                                                                                        \
    int countpairs_mocks_##SUF(const int64_t ND1, REAL *X1, REAL *theta1, REAL *czD1, const int64_t ND2, REAL *phi2,     \
                               REAL *theta2, REAL *czD2, const int numthreads, const int autocorr, const char *binfile,  \
                               const REAL pimax, const int cosmology, results_countpairs_mocks *results,                \
                               struct config_options *options, struct extra_options *extra)                             \
    {                                                                                                                    \
        CF_TYPED(countpairs_mocks(ND1, X1, theta1, czD1, ND2, phi2, theta2, czD2, numthreads, autocorr, binfile,         \
                                  (double)pimax, cosmology, results, options, extra));                                  \
    }                                                                                                                    \
    int countpairs_mocks_s_mu_##SUF(const int64_t ND1, REAL *X1, REAL *theta1, REAL *czD1, const int64_t ND2,            \
                                    REAL *phi2, REAL *theta2, REAL *czD2, const int numthreads, const int autocorr,      \
                                    const char *sbinfile, const double mu_max, const int nmu_bins, const int cosmology,  \
                                    results_countpairs_mocks_s_mu *results, struct config_options *options,             \
                                    struct extra_options *extra)                                                         \
    {                                                                                                                    \
        CF_TYPED(countpairs_mocks_s_mu(ND1, X1, theta1, czD1, ND2, phi2, theta2, czD2, numthreads, autocorr, sbinfile,   \
                                       mu_max, nmu_bins, cosmology, results, options, extra));                          \
    }

CF_TYPED_MOCKS(float, float)
CF_TYPED_MOCKS(double, double)
