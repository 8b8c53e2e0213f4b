/* corrfunc_b200_device.h -- the thin C ABI between the C host layer (csrc/host/ *.c) and the CUDA
 * layer (csrc/cuda/ *.cu).  Plain pointers and sizes only.  The host layer owns everything the
 * reference decides on the host (bins, extents, wrap, refine heuristics, nmesh, epilogues); the CUDA
 * layer owns what the reference's gridlink_* and *_kernels.c.src do:
 *
 *   cfb_upload        <- the copy_particles copy in gridlink   (utils/gridlink_impl.c.src:130-141)
 *   cfb_extent        <- get_max_min_DOUBLE                    (utils/gridlink_utils.c.src:52-70)
 *   cfb_count_box     <- gridlink_DOUBLE + generate_cell_pairs_DOUBLE + the per-cell-pair kernels
 *                        (utils/gridlink_impl.c.src:65-625, theory/x/x_kernels.c.src)
 *   cfb_count_theta   <- gridlink_mocks_theta_ra_dec_DOUBLE + generate_cell_pairs_mocks_theta_ra_dec
 *                        + countpairs_theta_mocks kernels (utils/gridlink_mocks_impl.c.src:1006-1650)
 *
 * All real-valued scalars cross this boundary as doubles that hold values already rounded to the
 * run precision `prec` (4 = float, 8 = double), so the device reproduces the reference's arithmetic.
 */
#ifndef CORRFUNC_B200_DEVICE_H
#define CORRFUNC_B200_DEVICE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum cfb_mode {
    CFB_DD = 0, CFB_XI = 1, CFB_RPPI = 2, CFB_WP = 3, CFB_SMU = 4, CFB_THETA = 5,
    /* survey geometry, line of sight = pair midpoint (mocks/DDrppi_mocks, mocks/DDsmu_mocks): the box lattice is the
     * non-periodic one over the Cartesian positions, max_sep[0] = the 3-D pruning radius */
    CFB_RPPI_MOCKS = 6, CFB_SMU_MOCKS = 7
};

#define CFB_MAX_EDGES 4096

/* What to count and how to bin it (one per call). */
typedef struct {
    int mode;      /* enum cfb_mode */
    int prec;      /* 4 | 8 */
    int autocorr;  /* 1: secondaries are set 0 again, count each unordered pair once */
    int nedges;    /* reference's `nbin` = number of bin edges */
    const double *edges; /* rupp_sqr[] (box modes) or costheta_upp[] (theta), REAL-valued */
    double pimax;        /* wp, rppi, rppi mocks (REAL-valued) */
    int npibin;          /* rppi */
    double inv_dpi;      /* rppi */
    double sqr_mumax;    /* smu */
    int nmu_bins;        /* smu */
    double inv_dmu;      /* smu */
    int need_avg;        /* accumulate sum of separations */
    int need_weights;    /* accumulate sum of w0*w1 (PAIR_PRODUCT) */
    int fast_acos;       /* theta: polynomial acos for the average */
    int64_t nslots;      /* histogram slots: nedges (1-D) or (nedges+1)*(n2+1) (2-D) */
} cfb_binning;

/* Reference box lattice (theory statistics). */
typedef struct {
    int nmesh[3];    /* reference lattice, decides first/second roles and wrap assignment */
    int refine[3];   /* neighbour reach in reference cells (bin_refine_factors after heuristics) */
    int periodic[3];
    double lo[3];    /* xmin, ymin, zmin */
    double inv[3];   /* 1/binsize (0 if flat) */
    double wrap[3];  /* periodic wrap per axis, 0 if that axis is not periodic */
    double max_sep[3]; /* pruning radii: [0]=3-D (rmax), [1]=2-D (rpmax), [2]=1-D in z (pimax); <=0: unused */
} cfb_box_lattice;

/* Reference RA/DEC lattice (DDtheta). */
typedef struct {
    int ngrid_dec;
    const int *ngrid_ra; /* [ngrid_dec] RA cells per DEC band (1 everywhere when !link_in_ra) */
    double dec_min, inv_dec_diff;
    double ra_min, ra_max, inv_ra_diff;
    int ra_refine, dec_refine;
    double sqr_max_chord; /* 2(1-cos(thetamax)) */
    int enable_min_sep;
    int sub;             /* device-only refinement: every reference cell is split sub x sub ways (in DEC and RA);
                          * from cfb_theta_subdivision, the same value for both particle sets */
} cfb_theta_lattice;

typedef struct {
    uint64_t *npairs; /* [nslots] raw counts (no x2, no self pairs) */
    double *sum_sep;  /* [nslots] or NULL */
    double *sum_w;    /* [nslots] or NULL */
} cfb_hist;

/* Per-call measurements (filled by every cfb_count_*). Times in milliseconds (CUDA events). */
typedef struct {
    double ms_h2d, ms_gridlink, ms_pairs, ms_total_device;
    uint64_t n_eval;      /* pair separations computed and range-tested by the pair kernel (padding excluded) */
    uint64_t n_tilepairs; /* (primary tile, secondary cell) pairs that survived pruning */
    int64_t n_cells, n_tiles;
    int fine[3];          /* fine lattice used on the device */
    int kernel_launches;  /* CUDA kernels launched by this call */
    int kernel_kind;      /* 0 = legacy generic, 1 = fast 1-D, 2 = per-pair-sum */
    uint64_t n_analytic;  /* pairs binned from cell bounding boxes alone, without evaluating a separation */
    uint64_t n_levelpairs; /* sum over evaluated pairs of the number of edge compares (levels) each one took */
} cfb_stats;

/* Lifetime: a lazily created per-process context on the current CUDA device (or CORRFUNC_B200_DEVICE). */
int cfb_init(void);
void cfb_shutdown(void);
const char *cfb_last_error(void);

/* Particle sets: slot 0 = first/primary set, slot 1 = second set (cross-correlations).
 * Pointers may be host or device memory (detected with cudaPointerGetAttributes). w/ra/dec may be NULL. */
int cfb_upload(int slot, int prec, int64_t n, const void *x, const void *y, const void *z, const void *w,
               const void *ra, const void *dec);
/* 1 when p is device or managed memory (cudaPointerGetAttributes), else 0; blocking device-to-host copy (0 = ok).  The host
 * layer's epilogue reads weight arrays on the host and uses these when the caller passed device pointers. */
int cfb_is_device_ptr(const void *p);
int cfb_copy_to_host(void *dst, const void *src, size_t bytes);
/* Catalogue cache: while on, a particle set passed again under the same pointers / length / element size is taken to be
 * unchanged -- it is not copied again (whichever slot it was in) and keeps its sorted form while the lattice stays the
 * same.  For workflows that count DD, DR and RR over two catalogues (Corrfunc/utils.py:27-165). */
/* Interrupt flag (utils/macros.h:145-167, theory/DD/countpairs_impl.c.src:31-37,475-477: the reference installs SIGINT /
 * SIGTERM / SIGHUP handlers for the duration of a call and its loop over cell pairs polls the flag they set).  Returns the
 * HOST address of one int in mapped pinned memory (a plain static int without a CUDA device): the host layer's signal
 * handler stores 1 there; the persistent pair kernels read it at every 64th tile fetch (a read over PCIe) and, when it is
 * set, push the tile counter past the end, which stops every warp at its next fetch. */
volatile int *cfb_abort_flag(void);
void cfb_set_catalog_cache(int on);
long long cfb_catalog_cache_hits(void);
/* Folds the device-side min/max of slot's x,y,z (or ra,dec when which==1) into lohi[6]={min3,max3}. */
int cfb_extent(int slot, int which, double lohi[6]);

/* Work sharding across ranks (one process per GPU): this process handles the primary cells c with
 * (c / 8) % nranks == rank -- by cell, not by tile, because the particle order inside a cell differs between
 * the ranks' replicas; every rank holds all particles and the histograms are summed by the reduce hook. */
void cfb_set_shard(int rank, int nranks);
void cfb_get_shard(int *rank, int *nranks);

int cfb_count_box(const cfb_binning *bin, const cfb_box_lattice *lat, cfb_hist *out, cfb_stats *stats);
/* Devices the last cfb_count_box ran on.  CORRFUNC_B200_NGPUS = n shards one call over n devices of this process (default:
 * all visible ones once a particle set holds 4 M points; 1 under cfb_set_shard with nranks > 1 or CORRFUNC_B200_DEVICE). */
int cfb_last_device_count(void);

/* Theta: two-phase because the reference's neighbour search needs per-cell RA bounds.
 * cfb_theta_gridlink sorts slot(s) into the lattice and returns per-cell counts and bounds (host
 * arrays, caller-allocated, ncells entries; bounds are {lo,hi} pairs).  cfb_count_theta then takes
 * the CSR neighbour list built by the host. */
/* Persistent pinned host buffer number `which` (0..5) of at least `bytes` bytes, owned by the context (NULL on failure). */
void *cfb_host_scratch(int which, size_t bytes);
/* sub x sub fine cells per reference cell such that a fine cell holds about the target occupancy. */
int cfb_theta_subdivision(int64_t nmax, int64_t ncells);
int cfb_theta_gridlink(int slot, int prec, const cfb_theta_lattice *lat, int64_t ncells, int64_t *counts,
                       double *ra_bounds, double *xyz_bounds /* [ncells][6] */);
int cfb_count_theta(const cfb_binning *bin, int64_t ncells, const int64_t *ngb_offsets /* [ncells+1] */,
                    const int32_t *ngb_cells, cfb_hist *out, cfb_stats *stats);

/* Counts-in-spheres (theory/vpf, mocks/vpf_mocks): for each of ncen centres (host arrays of element size prec), the
 * number of particles of `slot` per radial shell -> counts[ncen][nbin] (host).  The particles must lie in
 * [lo, lo + ext] per axis; they are gridded on an internal lattice with cells a little larger than rmax (regrid = 0
 * reuses the lattice of the previous call on this slot); the counts do not depend on it.  On a periodic axis the centre
 * is shifted by -+wrap[axis] towards a particle more than half a wrap away (the reference does the same per neighbour
 * cell, theory/vpf/countspheres_impl.c.src:331-380).  shells = 1: edges[nbin] are the squared upper shell edges
 * (REAL-valued), r2 = fma(dz,dz, fma(dy,dy, dx*dx)), shell assignment as in vpf_mocks_kernels.c.src:63-92.
 * shells = 0: nbin must be 1, r2 = dx*dx + dy*dy + dz*dz, counts = #{r2 < rmax_sqr} (count_neighbors). */
int cfb_count_spheres(int slot, int prec, const double lo[3], const double ext[3], const int periodic[3],
                      const double wrap[3], int regrid, int64_t ncen, const void *xc, const void *yc, const void *zc,
                      double rmax, double rmax_sqr, int nbin, const double *edges, int shells, uint32_t *counts);

/* Tunables (mostly for tests / benchmarks). */
void cfb_set_target_occupancy(int particles_per_fine_cell); /* 0 = default */
void cfb_force_kernel(int kind);                            /* -1 auto, 0 no fast kernel (per-pair-sum, else generic), 1 fast, 3 legacy generic only */

#ifdef __cplusplus
}
#endif
#endif
