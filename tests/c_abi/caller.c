/* caller.c -- a C program written against the REFERENCE's public headers (utils/defs.h, theory/x/countpairs_x.h,
 * mocks/DDtheta_mocks/countpairs_theta_mocks.h, included from /root/reference at build time: tests/c_abi/build.sh) and
 * linked against libcorrfunc_b200.so.  It is what a user of the reference's static-library interface
 * (docs/source/staticlibrary-interface.rst:33-117) writes: get_config_options(), get_extra_options(), one call, read
 * the results struct, free_results*().  Test infrastructure: tests/test_gpu_dropin.py feeds it a particle file and
 * compares what it prints with the CPU oracle.
 *
 *   caller <stat> <prec 4|8> <particle file> <bin file> <boxsize> [pimax] [mu_max nmu_bins]
 * particle file: int64 N1, int64 N2, then x|y|z|w of set 1 (N1 each) and of set 2 (N2 each; N2 = 0: autocorrelation),
 * all of element size prec; for DDtheta the "x" and "y" arrays are RA and DEC in degrees.
 * output: one line per bin / slot "npairs ravg weightavg [xi|wp]". */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "defs.h"
#include "countpairs.h"
#include "countpairs_rp_pi.h"
#include "countpairs_s_mu.h"
#include "countpairs_wp.h"
#include "countpairs_xi.h"
#include "countpairs_theta_mocks.h"

static void *rd(FILE *f, size_t n, size_t sz)
{
    void *p = malloc(n * sz + 1);
    if (n && fread(p, sz, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
    return p;
}

int main(int argc, char **argv)
{
    if (argc < 6) { fprintf(stderr, "usage: caller stat prec particles bins boxsize [pimax] [mu_max nmu]\n"); return 2; }
    const char *stat = argv[1];
    const int prec = atoi(argv[2]);
    FILE *f = fopen(argv[3], "rb");
    if (!f) { perror(argv[3]); return 2; }
    int64_t n[2];
    if (fread(n, 8, 2, f) != 2) return 2;
    void *a[2][4];
    for (int s = 0; s < 2; s++)
        for (int k = 0; k < 4; k++) a[s][k] = rd(f, (size_t)n[s], (size_t)prec);
    fclose(f);
    const char *binfile = argv[4];
    const double boxsize = atof(argv[5]);
    const int autocorr = n[1] == 0;
    if (autocorr) {  /* like the reference's own callers: the first set again (DDtheta returns early on ND2 == 0) */
        n[1] = n[0];
        for (int k = 0; k < 4; k++) a[1][k] = a[0][k];
    }

    struct config_options options = get_config_options();
    options.float_type = (uint8_t)prec;
    options.verbose = 0;
    options.periodic = boxsize > 0;
    options.need_avg_sep = 1;
    options.boxsize_x = options.boxsize_y = options.boxsize_z = boxsize > 0 ? boxsize : BOXSIZE_NOTGIVEN;
    options.c_api_timer = 1;
    struct extra_options extra = get_extra_options(PAIR_PRODUCT);
    extra.weights0.weights[0] = a[0][3];
    extra.weights1.weights[0] = autocorr ? NULL : a[1][3];
    const int nthreads = 2;
    int status = EXIT_FAILURE;

    if (strcmp(stat, "DD") == 0) {
        results_countpairs r;
        status = countpairs(n[0], a[0][0], a[0][1], a[0][2], n[1], a[1][0], a[1][1], a[1][2], nthreads, autocorr, binfile, &r,
                            &options, &extra);
        if (status == EXIT_SUCCESS) {
            for (int i = 1; i < r.nbin; i++) printf("%" PRIu64 " %.17g %.17g\n", r.npairs[i], r.rpavg[i], r.weightavg[i]);
            free_results(&r);
        }
    } else if (strcmp(stat, "DDrppi") == 0) {
        const double pimax = atof(argv[6]);
        results_countpairs_rp_pi r;
        status = countpairs_rp_pi(n[0], a[0][0], a[0][1], a[0][2], n[1], a[1][0], a[1][1], a[1][2], nthreads, autocorr,
                                  binfile, pimax, &r, &options, &extra);
        if (status == EXIT_SUCCESS) {
            for (int i = 1; i < r.nbin; i++)
                for (int j = 0; j < r.npibin; j++) {
                    const int k = i * (r.npibin + 1) + j;
                    printf("%" PRIu64 " %.17g %.17g\n", r.npairs[k], r.rpavg[k], r.weightavg[k]);
                }
            free_results_rp_pi(&r);
        }
    } else if (strcmp(stat, "DDsmu") == 0) {
        const double mu_max = atof(argv[6]);
        const int nmu = atoi(argv[7]);
        results_countpairs_s_mu r;
        status = countpairs_s_mu(n[0], a[0][0], a[0][1], a[0][2], n[1], a[1][0], a[1][1], a[1][2], nthreads, autocorr,
                                 binfile, mu_max, nmu, &r, &options, &extra);
        if (status == EXIT_SUCCESS) {
            for (int i = 1; i < r.nsbin; i++)
                for (int j = 0; j < r.nmu_bins; j++) {
                    const int k = i * (r.nmu_bins + 1) + j;
                    printf("%" PRIu64 " %.17g %.17g\n", r.npairs[k], r.savg[k], r.weightavg[k]);
                }
            free_results_s_mu(&r);
        }
    } else if (strcmp(stat, "wp") == 0) {
        const double pimax = atof(argv[6]);
        results_countpairs_wp r;
        status = countpairs_wp(n[0], a[0][0], a[0][1], a[0][2], boxsize, nthreads, binfile, pimax, &r, &options, &extra);
        if (status == EXIT_SUCCESS) {
            for (int i = 1; i < r.nbin; i++)
                printf("%" PRIu64 " %.17g %.17g %.17g\n", r.npairs[i], r.rpavg[i], r.weightavg[i], r.wp[i]);
            free_results_wp(&r);
        }
    } else if (strcmp(stat, "xi") == 0) {
        results_countpairs_xi r;
        status = countpairs_xi(n[0], a[0][0], a[0][1], a[0][2], boxsize, nthreads, binfile, &r, &options, &extra);
        if (status == EXIT_SUCCESS) {
            for (int i = 1; i < r.nbin; i++)
                printf("%" PRIu64 " %.17g %.17g %.17g\n", r.npairs[i], r.ravg[i], r.weightavg[i], r.xi[i]);
            free_results_xi(&r);
        }
    } else if (strcmp(stat, "DDtheta") == 0) {
        results_countpairs_theta r;
        options.link_in_dec = 1;
        options.link_in_ra = 1;
        status = countpairs_theta_mocks(n[0], a[0][0], a[0][1], n[1], a[1][0], a[1][1], nthreads, autocorr, binfile, &r,
                                        &options, &extra);
        if (status == EXIT_SUCCESS) {
            for (int i = 1; i < r.nbin; i++) printf("%" PRIu64 " %.17g %.17g\n", r.npairs[i], r.theta_avg[i], r.weightavg[i]);
            free_results_countpairs_theta(&r);
        }
    } else {
        fprintf(stderr, "unknown statistic %s\n", stat);
        return 2;
    }
    if (status != EXIT_SUCCESS) { fprintf(stderr, "%s failed\n", stat); return 1; }
    fprintf(stderr, "c_api_time %g s\n", options.c_api_time);
    return 0;
}
