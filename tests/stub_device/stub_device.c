/* TEST INFRASTRUCTURE ONLY -- a CPU stand-in for the CUDA layer's C ABI (include/corrfunc_b200_device.h), linked with
 * the product's HOST layer (corrfunc_b200/csrc/host/cf_host.c) into a throw-away library by tests/test_cpu_host_layer.py,
 * so that the host-side driver of countspheres_mocks (centres file, cz -> distance, shift, centre selection, pN) can be
 * and of countpairs_mocks / countpairs_mocks_s_mu (angles and cz -> Cartesian, extents, bins, epilogue) can be exercised
 * without a GPU.  Never shipped, never loaded by the product; the particle upload, the pinned scratch, the extent, the
 * counts-in-spheres and the two survey-geometry pair modes are brute force, every other entry point fails. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <signal.h>
#include <string.h>

#include "corrfunc_b200_device.h"

static struct {
    int prec;
    int64_t n;
    void *x, *y, *z, *w;
} g_set[2];
static void *g_scratch[6];

const char *cfb_last_error(void) { return "stub device layer"; }
int cfb_init(void) { return 0; }
int cfb_is_device_ptr(const void *p) { (void)p; return 0; }
int cfb_copy_to_host(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); return 0; }
int cfb_last_device_count(void) { return 1; }
static volatile int stub_abort;
volatile int *cfb_abort_flag(void) { return &stub_abort; }
void cfb_set_catalog_cache(int on) { (void)on; }
long long cfb_catalog_cache_hits(void) { return 0; }
void cfb_shutdown(void) {}
void cfb_set_target_occupancy(int n) { (void)n; }
void cfb_force_kernel(int k) { (void)k; }
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
    (void)ra, (void)dec;
    const void *src[4] = {x, y, z, w};
    void **dst[4] = {&g_set[slot].x, &g_set[slot].y, &g_set[slot].z, &g_set[slot].w};
    for (int a = 0; a < 4; a++) {
        free(*dst[a]);
        *dst[a] = NULL;
        if (!src[a]) continue;
        *dst[a] = malloc((size_t)(n > 0 ? n : 1) * prec);
        memcpy(*dst[a], src[a], (size_t)n * prec);
    }
    g_set[slot].prec = prec;
    g_set[slot].n = n;
    return 0;
}
int cfb_extent(int slot, int which, double lohi[6])
{
    if (which != 0) return 1;
    const void *a[3] = {g_set[slot].x, g_set[slot].y, g_set[slot].z};
    for (int k = 0; k < 3; k++)
        for (int64_t i = 0; i < g_set[slot].n; i++) {
            const double v = g_set[slot].prec == 4 ? (double)((const float *)a[k])[i] : ((const double *)a[k])[i];
            if (v < lohi[k]) lohi[k] = v;
            if (v > lohi[3 + k]) lohi[3 + k] = v;
        }
    return 0;
}

/* the two survey-geometry modes only, every pair tried (the lattice only prunes): the per-pair arithmetic of
 * countpairs_rp_pi_mocks_kernels.c.src:200-300 / countpairs_s_mu_mocks_kernels.c.src:196-290 */
#define MOCKS_PAIRS(T, FMA, SQRT)                                                                                     \
    {                                                                                                                 \
        const T *x0 = g_set[0].x, *y0 = g_set[0].y, *z0 = g_set[0].z, *w0 = g_set[0].w;                               \
        const T *x1 = g_set[s1].x, *y1 = g_set[s1].y, *z1 = g_set[s1].z, *w1 = g_set[s1].w;                           \
        const int ne = b->nedges;                                                                                     \
        T E[4096];                                                                                                    \
        for (int k = 0; k < ne; k++) E[k] = (T)b->edges[k];                                                           \
        const T pimax = (T)b->pimax, sqr_pimax = pimax * pimax, sqr_max_sep = E[ne - 1] + sqr_pimax;                  \
        const T inv_dpi = (T)b->inv_dpi, inv_dmu = (T)b->inv_dmu, sqr_mumax = (T)b->sqr_mumax;                        \
        _Pragma("omp parallel for schedule(dynamic, 64)")                                                             \
        for (int64_t i = 0; i < g_set[0].n; i++)                                                                      \
            for (int64_t j = b->autocorr ? i + 1 : 0; j < g_set[s1].n; j++) {                                         \
                const T dx = x1[j] - x0[i], dy = y1[j] - y0[i], dz = z1[j] - z0[i];                                   \
                const T px = x1[j] + x0[i], py = y1[j] + y0[i], pz = z1[j] + z0[i];                                   \
                const T t1 = px * dx, t2 = py * dy;                                                                   \
                const T sl = FMA(pz, dz, t1 + t2), sl2 = sl * sl;                                                     \
                const T s2 = FMA(dx, dx, FMA(dy, dy, dz * dz));                                                       \
                T key, second, sep2;                                                                                  \
                int n2;                                                                                               \
                if (b->mode == CFB_RPPI_MOCKS) {                                                                      \
                    if (!(s2 < sqr_max_sep)) continue;                                                                \
                    const T l2 = FMA(px, px, FMA(py, py, pz * pz));                                                   \
                    if (!(sl2 < sqr_pimax * l2)) continue;                                                            \
                    const T dpar2 = sl2 / l2, dperp2 = s2 - dpar2;                                                    \
                    if (!(dpar2 < sqr_pimax && dperp2 < E[ne - 1] && dperp2 >= E[0])) continue;                       \
                    key = dperp2, second = SQRT(dpar2) * inv_dpi, n2 = b->npibin, sep2 = dperp2;                      \
                } else {                                                                                              \
                    if (!(s2 < E[ne - 1] && s2 >= E[0])) continue;                                                    \
                    const T l2 = FMA(px, px, FMA(py, py, pz * pz));                                                   \
                    const T mu2 = sl2 / (l2 * s2);                                                                    \
                    if (!(mu2 < sqr_mumax)) continue;                                                                 \
                    key = s2, second = SQRT(mu2) * inv_dmu, n2 = b->nmu_bins, sep2 = s2;                              \
                }                                                                                                     \
                int kb;                                                                                               \
                for (kb = ne - 1; kb >= 1; kb--)                                                                      \
                    if (key >= E[kb - 1]) break;                                                                      \
                const T fin = (T)kb * (T)(n2 + 1) + second;                                                           \
                const int64_t slot = (int64_t)(int)fin;                                                               \
                _Pragma("omp critical")                                                                               \
                {                                                                                                     \
                    o->npairs[slot]++;                                                                                \
                    if (b->need_avg) o->sum_sep[slot] += (double)SQRT(sep2);                                          \
                    if (b->need_weights) o->sum_w[slot] += (double)(T)(w0[i] * w1[j]);                                \
                }                                                                                                     \
            }                                                                                                         \
    }

int cfb_count_box(const cfb_binning *b, const cfb_box_lattice *l, cfb_hist *o, cfb_stats *s)
{
    (void)l;
    if (s) memset(s, 0, sizeof(*s));
    /* test hook: a signal arrives while the device counts (tests/test_cpu_host_layer.py::test_host_layer_interrupt) */
    if (getenv("CFB_STUB_RAISE")) raise(atoi(getenv("CFB_STUB_RAISE")));
    if (!(b->mode == CFB_RPPI_MOCKS || b->mode == CFB_SMU_MOCKS) || b->nedges > 4096) return 1;
    const int s1 = b->autocorr ? 0 : 1;
    if (b->prec == 4) MOCKS_PAIRS(float, fmaf, sqrtf)
    else MOCKS_PAIRS(double, fma, sqrt)
    return 0;
}
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
        const T Cc[3] = {((const T *)xc)[c], ((const T *)yc)[c], ((const T *)zc)[c]};                                \
        for (int64_t j = 0; j < g_set[slot].n; j++) {                                                                \
            const T Pp[3] = {((const T *)g_set[slot].x)[j], ((const T *)g_set[slot].y)[j], ((const T *)g_set[slot].z)[j]}; \
            T dd[3];                                                                                                 \
            for (int a = 0; a < 3; a++) { /* nearest image on a periodic axis, as the device kernel does */          \
                T cen = Cc[a];                                                                                       \
                if (periodic[a]) {                                                                                   \
                    const T raw = Pp[a] - Cc[a], half = (T)0.5 * (T)wrap[a];                                         \
                    if (raw > half) cen = Cc[a] + (T)wrap[a];                                                        \
                    else if (raw < -half) cen = Cc[a] - (T)wrap[a];                                                  \
                }                                                                                                    \
                dd[a] = cen - Pp[a];                                                                                 \
            }                                                                                                        \
            const T dx = dd[0], dy = dd[1], dz = dd[2];                                                              \
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

int cfb_count_spheres(int slot, int prec, const double lo[3], const double ext[3], const int periodic[3],
                      const double wrap[3], int regrid, int64_t ncen, const void *xc, const void *yc, const void *zc,
                      double rmax, double rmax_sqr, int nbin, const double *edges, int shells, uint32_t *counts)
{
    (void)lo, (void)ext, (void)regrid, (void)rmax;
    if (g_set[slot].prec != prec) return 1;
    if (prec == 4) {
        SPHERES(float, fmaf)
    } else {
        SPHERES(double, fma)
    }
    return 0;
}
