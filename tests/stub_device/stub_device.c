/* TEST INFRASTRUCTURE ONLY -- a CPU stand-in for the CUDA layer's C ABI (include/corrfunc_b200_device.h), linked with
 * the product's HOST layer (corrfunc_b200/csrc/host/cf_host.c) into a throw-away library by tests/test_cpu_host_vpf.py,
 * so that the host-side driver of countspheres_mocks (centres file, cz -> distance, shift, centre selection, pN) can be
 * exercised without a GPU.  Never shipped, never loaded by the product; only the particle upload, the pinned scratch
 * and the counts-in-spheres entry points do anything (brute force), the pair-counting entry points fail. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "corrfunc_b200_device.h"

static struct {
    int prec;
    int64_t n;
    void *x, *y, *z;
} g_set[2];
static void *g_scratch[6];

const char *cfb_last_error(void) { return "stub device layer"; }
int cfb_init(void) { return 0; }
void cfb_set_shard(int rank, int nranks) { (void)rank, (void)nranks; }
void cfb_get_shard(int *rank, int *nranks) { *rank = 0, *nranks = 1; }
void *cfb_host_scratch(int which, size_t bytes)
{
    if (which < 0 || which >= 6) return NULL;
    free(g_scratch[which]);
    g_scratch[which] = malloc(bytes ? bytes : 1);
    return g_scratch[which];
}
int cfb_upload(int slot, int prec, int64_t n, const void *x, const void *y, const void *z, const void *w, const void *ra,
               const void *dec)
{
    (void)w, (void)ra, (void)dec;
    const void *src[3] = {x, y, z};
    void **dst[3] = {&g_set[slot].x, &g_set[slot].y, &g_set[slot].z};
    for (int a = 0; a < 3; a++) {
        free(*dst[a]);
        *dst[a] = malloc((size_t)(n > 0 ? n : 1) * prec);
        memcpy(*dst[a], src[a], (size_t)n * prec);
    }
    g_set[slot].prec = prec;
    g_set[slot].n = n;
    return 0;
}
int cfb_extent(int slot, int which, double lohi[6]) { (void)slot, (void)which, (void)lohi; return 1; }
int cfb_count_box(const cfb_binning *b, const cfb_box_lattice *l, cfb_hist *o, cfb_stats *s) { (void)b, (void)l, (void)o, (void)s; return 1; }
int cfb_theta_subdivision(int64_t nmax, int64_t ncells) { (void)nmax, (void)ncells; return 1; }
int cfb_theta_gridlink(int slot, int prec, const cfb_theta_lattice *lat, int64_t ncells, int64_t *counts, double *ra_bounds,
                       double *xyz_bounds)
{
    (void)slot, (void)prec, (void)lat, (void)ncells, (void)counts, (void)ra_bounds, (void)xyz_bounds;
    return 1;
}
int cfb_count_theta(const cfb_binning *bin, int64_t ncells, const int64_t *off, const int32_t *cells, cfb_hist *out, cfb_stats *st)
{
    (void)bin, (void)ncells, (void)off, (void)cells, (void)out, (void)st;
    return 1;
}

#define SPHERES(T, FMA)                                                                                              \
    for (int64_t c = 0; c < ncen; c++) {                                                                             \
        for (int k = 0; k < nbin; k++) counts[c * nbin + k] = 0;                                                     \
        const T X = ((const T *)xc)[c], Y = ((const T *)yc)[c], Z = ((const T *)zc)[c];                              \
        for (int64_t j = 0; j < g_set[slot].n; j++) {                                                                \
            const T dx = X - ((const T *)g_set[slot].x)[j], dy = Y - ((const T *)g_set[slot].y)[j],                  \
                    dz = Z - ((const T *)g_set[slot].z)[j];                                                          \
            if (shells) {                                                                                            \
                const T r2 = FMA(dz, dz, FMA(dy, dy, dx * dx));                                                      \
                if (!(r2 < (T)rmax_sqr)) continue;                                                                   \
                int left = 1;                                                                                        \
                for (int k = nbin - 1; k >= 1; k--)                                                                  \
                    if (r2 < (T)edges[k] && r2 >= (T)edges[k - 1]) {                                                 \
                        counts[c * nbin + k]++;                                                                      \
                        left = 0;                                                                                    \
                        break;                                                                                       \
                    }                                                                                                \
                if (left && nbin >= 2) counts[c * nbin]++;                                                           \
            } else {                                                                                                 \
                const T r2 = dx * dx + dy * dy + dz * dz;                                                            \
                if (r2 < (T)rmax_sqr) counts[c * nbin]++;                                                            \
            }                                                                                                        \
        }                                                                                                            \
    }

int cfb_count_spheres(int slot, int prec, double extent, int regrid, int64_t ncen, const void *xc, const void *yc,
                      const void *zc, double rmax, double rmax_sqr, int nbin, const double *edges, int shells,
                      uint32_t *counts)
{
    (void)extent, (void)regrid, (void)rmax;
    if (g_set[slot].prec != prec) return 1;
    if (prec == 4) {
        SPHERES(float, fmaf)
    } else {
        SPHERES(double, fma)
    }
    return 0;
}
