/* corrfunc_b200_defs.h -- option / weight structures of the drop-in C ABI.
 *
 * Binary layout is that of the reference's `struct config_options` and `struct extra_options`
 * (reference: utils/defs.h:53-156 and utils/defs.h:353-402, Corrfunc v2.5.3): both are frozen at
 * 1024 bytes, so a caller compiled against the reference's headers can pass its structs to this
 * library unchanged.  Field order and widths therefore have to be what they are; everything else
 * in this file (helpers, comments, checks) is written for this project.
 *
 * Fields that only steer the reference's CPU kernels are accepted and ignored by the B200 path
 * (results do not depend on them, see DESIGN.md "accepted-and-ignored options"):
 *   instruction_set, copy_particles, use_heap_sort, sort_on_z, enable_min_sep_opt,
 *   c_cell_timer, fast_divide_and_NR_steps (always a true IEEE divide on the GPU).
 */
#ifndef CORRFUNC_B200_DEFS_H
#define CORRFUNC_B200_DEFS_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CORRFUNC_API_VERSION "2.5.3" /* options->version must start with this (theory/DD/countpairs.c:52-55) */
#ifndef API_VERSION
#define API_VERSION CORRFUNC_API_VERSION
#endif

/* low nibble of binning_flags: are the bin-refine factors user-pinned? (utils/defs.h:31-35) */
#define BINNING_REF_MASK 0x0000000F
#define BINNING_ORD_MASK 0x000000F0
#define BINNING_DFL 0x0
#define BINNING_CUST 0x1

#define OPTIONS_HEADER_SIZE 1024
#define EXTRA_OPTIONS_HEADER_SIZE 1024
#define MAX_FAST_DIVIDE_NR_STEPS 3
#define BOXSIZE_NOTGIVEN (-2.)
#define MAX_NUM_WEIGHTS 10

/* lattice heuristics constants (utils/macros.h:3-7) */
#ifndef NLATMAX
#define NLATMAX 100
#endif
#define BOOST_CELL_THRESH 10
#define BOOST_NUMPART_THRESH 250
#define BOOST_BIN_REF 1

/* ISA selector values of the reference (utils/cpu_features.h:20-35); parsed, never used. */
typedef enum {
    DEFAULT = -42,
    FALLBACK = 0,
    SSE = 1,
    SSE2 = 2,
    SSE3 = 3,
    SSSE3 = 4,
    SSE4 = 5,
    SSE42 = 6,
    AVX = 7,
    AVX2 = 8,
    AVX512F = 9,
    ARM64 = 10,
    NUM_ISA
} isa;

/* utils/cpu_features.h:36-37 -- the reference's extension modules ask which kernels the CPU can run and pass the answer
 * back as options->instruction_set.  This library has one code path (the GPU's); every value is accepted and ignored, and
 * the "maximum" reported is the reference's own maximum so that isa="fastest" resolves to a value the reference knows. */
extern int runtime_instrset_detect(void);
extern int get_max_usable_isa(void);

struct api_cell_timings {
    int64_t N1;
    int64_t N2;
    int64_t time_in_ns;
    int first_cellindex;
    int second_cellindex;
    int tid;
};

struct config_options {
    union {
        double boxsize;
        double boxsize_x;
    };
    double boxsize_y;
    double boxsize_z;

    struct { /* cosmology block: only used by mocks routines that are out of scope here */
        double OMEGA_M;
        double OMEGA_B;
        double OMEGA_L;
        double HUBBLE;
        double LITTLE_H;
        double SIGMA_8;
        double NS;
    };

    double c_api_time; /* wall time of the call when c_api_timer is set */

    struct api_cell_timings *cell_timings; /* never filled by the GPU path */
    int64_t totncells_timings;

    size_t float_type;       /* 4 or 8: element size of every input array of the call */
    int32_t instruction_set; /* ignored */

    char version[32];
    uint8_t verbose;
    uint8_t c_api_timer;
    uint8_t c_cell_timer;

    uint8_t need_avg_sep;
    uint8_t autocorr;

    uint8_t periodic;
    uint8_t sort_on_z;

    uint8_t is_comoving_dist;

    uint8_t link_in_dec;
    uint8_t link_in_ra;

    uint8_t fast_divide_and_NR_steps;
    uint8_t fast_acos;
    uint8_t enable_min_sep_opt;

    int8_t bin_refine_factors[3];
    uint16_t max_cells_per_dim;

    uint8_t copy_particles;
    uint8_t use_heap_sort;
    union {
        uint32_t binning_flags;
        uint8_t bin_masks[4];
    };

    uint8_t reserved[OPTIONS_HEADER_SIZE - 33 * sizeof(char) - sizeof(size_t) - 11 * sizeof(double) -
                     3 * sizeof(int) - sizeof(uint16_t) - 16 * sizeof(uint8_t) -
                     sizeof(struct api_cell_timings *) - sizeof(int64_t)];
};

typedef struct {
    void *weights[MAX_NUM_WEIGHTS]; /* weights[w] -> N reals of the positions' dtype */
    int64_t num_weights;
} weight_struct;

typedef enum { NONE = -42, PAIR_PRODUCT = 0, NUM_WEIGHT_TYPE } weight_method_t;

struct extra_options {
    weight_struct weights0;
    weight_struct weights1;
    weight_method_t weight_method;
    uint8_t reserved[EXTRA_OPTIONS_HEADER_SIZE - 2 * sizeof(weight_struct) - sizeof(weight_method_t)];
};

#if defined(__cplusplus)
static_assert(sizeof(struct config_options) == OPTIONS_HEADER_SIZE, "config_options must stay 1024 bytes");
static_assert(sizeof(struct extra_options) == EXTRA_OPTIONS_HEADER_SIZE, "extra_options must stay 1024 bytes");
#else
_Static_assert(sizeof(struct config_options) == OPTIONS_HEADER_SIZE, "config_options must stay 1024 bytes");
_Static_assert(sizeof(struct extra_options) == EXTRA_OPTIONS_HEADER_SIZE, "extra_options must stay 1024 bytes");
#endif

/* ---- helpers mirroring the reference's inline API (utils/defs.h:158-346, 365-436) ---- */

static inline void set_bin_refine_scheme(struct config_options *o, const int8_t flag)
{
    o->binning_flags = (o->binning_flags & ~BINNING_REF_MASK) | ((uint32_t)flag & BINNING_REF_MASK);
}
static inline void reset_bin_refine_scheme(struct config_options *o) { set_bin_refine_scheme(o, BINNING_DFL); }
static inline int8_t get_bin_refine_scheme(const struct config_options *o)
{
    return (int8_t)(o->binning_flags & BINNING_REF_MASK);
}
static inline void reset_bin_refine_factors(struct config_options *o)
{
    o->bin_refine_factors[0] = 2;
    o->bin_refine_factors[1] = 2;
    o->bin_refine_factors[2] = 1;
    reset_bin_refine_scheme(o);
}
static inline void set_bin_refine_factors(struct config_options *o, const int f[3])
{
    for (int i = 0; i < 3; i++) {
        int v = f[i];
        if (v > INT8_MAX) {
            fprintf(stderr, "Warning: bin refine factor[%d] can be at most %d. Found %d instead\n", i, INT8_MAX, v);
            v = 1;
        }
        o->bin_refine_factors[i] = (int8_t)v;
    }
    reset_bin_refine_scheme(o);
}
static inline void set_custom_bin_refine_factors(struct config_options *o, const int f[3])
{
    set_bin_refine_factors(o, f);
    set_bin_refine_scheme(o, BINNING_CUST);
}
static inline void set_max_cells(struct config_options *o, const int max)
{
    if (max <= 0) {
        fprintf(stderr, "Warning: max. cells per dimension must be positive (got %d); unchanged\n", max);
        return;
    }
    if (max > INT16_MAX)
        fprintf(stderr, "Warning: max. cells per dimension = %d does not fit the 2-byte field (max %d)\n", max, INT16_MAX);
    o->max_cells_per_dim = (uint16_t)max;
}
static inline void reset_max_cells(struct config_options *o) { o->max_cells_per_dim = NLATMAX; }

/* Defaults follow the reference's shipped theory.options / mocks.options:
 * periodic, need_avg_sep (OUTPUT_RPAVG), copy_particles, enable_min_sep_opt, link_in_dec+ra, double. */
static inline struct config_options get_config_options(void)
{
    struct config_options o;
    memset(&o, 0, sizeof(o));
    snprintf(o.version, sizeof(o.version) - 1, "%s", CORRFUNC_API_VERSION);
    o.boxsize_x = BOXSIZE_NOTGIVEN;
    o.boxsize_y = BOXSIZE_NOTGIVEN;
    o.boxsize_z = BOXSIZE_NOTGIVEN;
    o.float_type = sizeof(double);
    o.verbose = 0;
    o.need_avg_sep = 1;
    o.periodic = 1;
    o.instruction_set = -1; /* 'fastest'; ignored by the GPU path */
    o.link_in_dec = 1;
    o.link_in_ra = 1;
    o.enable_min_sep_opt = 1;
    o.copy_particles = 1;
    o.totncells_timings = 0;
    o.cell_timings = NULL;
    reset_max_cells(&o);
    reset_bin_refine_factors(&o);
    return o;
}

static inline int get_num_weights_by_method(const weight_method_t method)
{
    return method == PAIR_PRODUCT ? 1 : 0;
}
static inline int get_weight_method_by_name(const char *name, weight_method_t *method)
{
    if (name == NULL || name[0] == '\0') {
        *method = NONE;
        return EXIT_SUCCESS;
    }
    if (strcmp(name, "pair_product") == 0 || strcmp(name, "p") == 0) {
        *method = PAIR_PRODUCT;
        return EXIT_SUCCESS;
    }
    return EXIT_FAILURE;
}
static inline struct extra_options get_extra_options(const weight_method_t weight_method)
{
    struct extra_options e;
    memset(&e, 0, sizeof(e));
    e.weight_method = weight_method;
    e.weights0.num_weights = get_num_weights_by_method(weight_method);
    e.weights1.num_weights = e.weights0.num_weights;
    return e;
}
static inline void free_cell_timings(struct config_options *o)
{
    if (o->totncells_timings > 0 && o->cell_timings != NULL) free(o->cell_timings);
    o->cell_timings = NULL;
    o->totncells_timings = 0;
}

#ifdef __cplusplus
}
#endif
#endif /* CORRFUNC_B200_DEFS_H */
