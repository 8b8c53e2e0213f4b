/* TEST INFRASTRUCTURE ONLY -- stand-in for <gsl/gsl_interp.h> (GSL is not installed in this image).
 * The reference's mocks/DDrppi_mocks and mocks/DDsmu_mocks include this header for the cz -> comoving
 * distance table lookup, a branch that runs only when options->is_comoving_dist == 0.  oracle/build_ref.sh
 * puts this directory on the include path so the UNMODIFIED reference sources compile.  Restated from GSL 2.x's
 * published source -- interpolation/linear.c:linear_eval (y_lo + (x - x_lo) / dx * (y_hi - y_lo)) and
 * interpolation/bsearch.c:gsl_interp_bsearch (bisection to x_array[i] <= x < x_array[i+1]); the accelerator only
 * caches the bracket, so it is ignored.  Out-of-range x: GSL raises GSL_EDOM (its default handler aborts); here NaN.
 * Parity for cz input is therefore pinned on the reference's own table code (utils/set_cosmo_dist.c, compiled
 * unmodified) plus this restatement of a 10-line GSL routine -- stated as such in DESIGN.md section 4c. */
#include <math.h>
#pragma once
#include <stdlib.h>

typedef struct { int kind; } gsl_interp_type;
static const gsl_interp_type gsl_shim_linear_type = {0};
#define gsl_interp_linear (&gsl_shim_linear_type)
typedef struct { size_t size; } gsl_interp;
typedef struct { size_t cache; } gsl_interp_accel;

static inline gsl_interp_accel *gsl_interp_accel_alloc(void) { return (gsl_interp_accel *)calloc(1, sizeof(gsl_interp_accel)); }
static inline void gsl_interp_accel_free(gsl_interp_accel *a) { free(a); }
static inline gsl_interp *gsl_interp_alloc(const gsl_interp_type *t, size_t n)
{
    (void)t;
    gsl_interp *p = (gsl_interp *)calloc(1, sizeof(gsl_interp));
    if (p) p->size = n;
    return p;
}
static inline int gsl_interp_init(gsl_interp *p, const double *x, const double *y, size_t n)
{
    (void)x, (void)y;
    p->size = n;
    return 0;
}
static inline void gsl_interp_free(gsl_interp *p) { free(p); }
static inline double gsl_interp_eval(const gsl_interp *p, const double *x, const double *y, double xv, gsl_interp_accel *a)
{
    size_t lo = 0, hi = p->size - 1;
    (void)a;
    if (xv < x[0] || xv > x[p->size - 1]) return NAN;
    while (hi - lo > 1) {
        const size_t mid = (lo + hi) / 2;
        if (x[mid] > xv) hi = mid; else lo = mid;
    }
    const double dx = x[lo + 1] - x[lo];
    return dx > 0.0 ? y[lo] + (xv - x[lo]) / dx * (y[lo + 1] - y[lo]) : y[lo];
}
