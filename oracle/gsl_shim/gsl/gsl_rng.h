/* TEST INFRASTRUCTURE ONLY -- stand-in for <gsl/gsl_rng.h> (GSL is not installed in this image), so that
 * oracle/build_ref.sh can compile the reference's theory/vpf UNMODIFIED.  theory/vpf draws its sphere centres from
 * gsl_rng_mt19937 (countspheres_impl.c.src:190-192, 303-305).  Restated from the published algorithm that GSL 2.x's
 * rng/mt.c implements: MT19937 of Matsumoto & Nishimura with the 2002 initialisation
 * (mt[i] = 1812433253 * (mt[i-1] ^ (mt[i-1] >> 30)) + i, seed 0 -> 4357), gsl_rng_uniform = next 32-bit word / 2^32.
 * The same stream comes out of numpy's independent MT19937 with legacy seeding (tests/test_cpu_host_layer.py). */
#pragma once
#include <stdlib.h>

typedef struct { int dummy; } gsl_rng_type;
static const gsl_rng_type gsl_shim_mt19937_type = {0};
#define gsl_rng_mt19937 (&gsl_shim_mt19937_type)
typedef struct {
    unsigned long mt[624];
    int mti;
} gsl_rng;

static inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *t)
{
    (void)t;
    gsl_rng *r = (gsl_rng *)calloc(1, sizeof(gsl_rng));
    if (r) r->mti = 625;
    return r;
}
static inline void gsl_rng_free(gsl_rng *r) { free(r); }
static inline void gsl_rng_set(gsl_rng *r, unsigned long s)
{
    if (s == 0) s = 4357;
    r->mt[0] = s & 0xffffffffUL;
    for (int i = 1; i < 624; i++) r->mt[i] = (1812433253UL * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (unsigned long)i) & 0xffffffffUL;
    r->mti = 624;
}
static inline unsigned long gsl_shim_mt_get(gsl_rng *r)
{
    unsigned long *const mt = r->mt;
    if (r->mti >= 624) {
        int kk;
        for (kk = 0; kk < 624 - 397; kk++) {
            const unsigned long y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
        }
        for (; kk < 623; kk++) {
            const unsigned long y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL);
            mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
        }
        {
            const unsigned long y = (mt[623] & 0x80000000UL) | (mt[0] & 0x7fffffffUL);
            mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
        }
        r->mti = 0;
    }
    unsigned long k = mt[r->mti++];
    k ^= (k >> 11);
    k ^= (k << 7) & 0x9d2c5680UL;
    k ^= (k << 15) & 0xefc60000UL;
    k ^= (k >> 18);
    return k & 0xffffffffUL;
}
static inline double gsl_rng_uniform(gsl_rng *r) { return gsl_shim_mt_get(r) / 4294967296.0; }
