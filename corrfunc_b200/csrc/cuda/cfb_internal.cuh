// cfb_internal.cuh -- shared declarations of the CUDA layer (context, particle sets, launch params).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "corrfunc_b200_device.h"

#define CFB_SHARD_GROUP 8    // fine cells per shard group: rank r owns the cells c with (c / 8) % nranks == r
#define CFB_PAD 4            // every fine cell's run in the sorted SoA starts at a multiple of this (16 B for float)
#define CFB_TILE 128         // primaries per tile in the generic kernel (one per thread)
#define CFB_FAST_MAX_EDGES 64 // the fast kernel keeps edges and the block histogram in static shared memory

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

// One particle set, raw (as uploaded) and cell-sorted.
struct ParticleSet {
    int prec = 0;
    int64_t n = 0;
    DevBuf rawbuf[6];            // owned copies of x,y,z,w,ra,dec when the caller passed host memory
    const void *raw[6] = {0};    // pointers actually used (may alias caller's device memory)
    DevBuf sorted[4];            // x,y,z,w cell-sorted, cells padded to CFB_PAD with NaN
    DevBuf cidx, rank;           // int32 per particle: fine cell index, arrival rank within the cell
    DevBuf count, start, tstart; // int32 per fine cell: particles, padded offset, first tile id
    DevBuf bounds;               // per fine cell: 6 reals {xlo,xhi,ylo,yhi,zlo,zhi} (+ {ralo,rahi} for theta)
    DevBuf tile_cell, tile_off;  // int32 per tile
    int64_t ncells = 0, npad = 0, ntiles = 0;
    bool gridded = false;
    // catalogue cache (cfb_set_catalog_cache): the caller's pointers this set was uploaded from, and the lattice it was
    // last sorted into -- an identical request is served without a copy / without a sort
    const void *src[6] = {0};
    bool src_valid = false;
    unsigned char grid_sig[160];
    bool grid_sig_valid = false;
};

struct Ctx {
    bool ready = false;
    int dev = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8];
    ParticleSet set[2];
    DevBuf scratch;  // reductions, scalars
    DevBuf sort_rec, sort_cid, sort_cur;  // two-pass scatter of large sets (gridlink.cu): records, their cells, cursors
    DevBuf hist;     // npairs | sum_sep | sum_w | n_eval, n_tilepairs
    DevBuf edges;
    DevBuf list_off, list_cells, ngrid_ra, ra_off;
    void *pinned = nullptr;  // small pinned host staging area
    size_t pinned_cap = 0;
    int shard_rank = 0, shard_n = 1;
    int target_occ = 0;
    int theta_sub = 1;  // sub x sub fine cells per reference RA/DEC cell of the last theta gridlink
    int force_kernel = -1;
    bool prefer_legacy = false;  // this count is better served by the legacy generic kernel (set per call)
    int last_kind = 0;  // kernel the last count ran: 0 legacy generic, 1 fast, 2 per-pair-sum
    int launches = 0;
    char err[512];
};

Ctx &cfb_ctx();
int cfb_fail(const char *fmt, ...);
int cfb_ensure(DevBuf &b, size_t bytes);

#define CK(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return cfb_fail("CUDA error %s at %s:%d (%s)", cudaGetErrorName(e__), __FILE__, __LINE__, \
                            cudaGetErrorString(e__));                                               \
    } while (0)

// Device view of a gridded particle set.
template <typename T>
struct SetView {
    const T *x, *y, *z, *w;
    const int *count, *start;
    const T *bounds;  // stride CFB_NB
};
#define CFB_NB 8  // bounds stride (reals per cell)

// Fine lattice = reference lattice with every cell split sub[] ways.
struct FineGeom {
    int n[3];    // reference nmesh
    int s[3];    // subdivision
    int ng[3];   // n*s
    int refine[3]; // neighbour reach in reference cells
    int reach[3];  // neighbour reach in fine cells: (refine+1)*s-1, whole reference cells on both sides
    int periodic[3];
};

struct PairParams {
    // exact-division magics for the fast kernel's candidate decode: ceil(2^32/d), 0 for d == 1
    unsigned m_wz, m_s[3];
    // binning
    int mode, nedges, npibin, nmu_bins, autocorr, cross;
    int64_t nslots;
    double pimax, inv_dpi, sqr_mumax, inv_dmu;
    int fast_acos;
    const void *edges;  // device, T[nedges]
    // lattice
    FineGeom g;
    double wrap[3];
    double max_sep[3];
    // neighbour list (theta): CSR over REFERENCE cells; a fine cell is reference cell * list_sub2 + sub-cell
    const int64_t *list_off;
    const int32_t *list_cells;
    int list_sub2;
    // tiles of the primary set
    const int *tile_cell, *tile_off;
    int64_t ntiles;
    int64_t tail_first;  // fast kernel: tiles from this one on are handed out in tail_parts pieces (groups of neighbour rows)
    int tail_parts;
    int shard_rank, shard_n;
    // outputs
    unsigned long long *npairs;
    double *sum_sep, *sum_w;
    const double *wmax;  // device: max |weight| of the first and of the second set (generic kernel, weights on)
    unsigned long long *counters;  // [0]=n_eval [1]=n_tilepairs [2]=pairs binned without evaluation [3]=sum of evaluated pairs x levels [4]=next tile (persistent warps)
    const volatile int *abort;  // mapped host flag set by the host layer's signal handler (cfb_abort_flag); polled at every 64th tile fetch
    int hist_in_smem;
    int sum_copies_shift;  // per-pair-sum kernel: log2 of the number of shared-memory histogram copies
    int sum_norm_drains;   // per-pair-sum kernel: drains of a warp between two normalisations of the block's histogram
    int sum_cmask;           // copies - 1
    unsigned sum_hstride_b, sum_hist_off;  // bytes from one histogram slot to the next; offset of the histogram in dynamic shared memory
    float sum_inv_dmu_f;
    int sum_kmin, sum_nkeys; // per-pair-sum kernel: separation-bin table (first float key, number of keys; 0 = no table)
    double coord_max;      // largest |coordinate| a particle (with its periodic shift) can have: float-filter margins of the per-pair-sum kernel
};

// Multi-rank sharding is by primary CELL, never by tile: every rank sorts its own replica and the order of the
// particles inside a cell (atomic arrival ranks) differs from rank to rank, so the tiles of one cell only
// partition its particles consistently when one rank handles all of them.  Groups of 8 consecutive cells are dealt to the
// ranks in turn: every rank sees the same mix of cells whatever the catalogue and whatever the reference's role filter
// does to the cells near the ends of the index range (contiguous cost-balanced cell ranges were tried in round 2: the
// filter gives the low-index cells a fraction of the pairs of the high-index ones, a neighbour-sum cost model still left
// rank 0 of 8 with 0.119 and rank 7 with 0.131 of config 5's evaluations, and the interleaving itself costs nothing --
// tools/exp_shard.py: what a rank loses against an n-th of the full kernel is the tail of its last tiles).
__host__ __device__ __forceinline__ bool cfb_owns_cell(const int cell, const int rank, const int nranks)
{
    return nranks <= 1 || (cell / CFB_SHARD_GROUP) % nranks == rank;
}

// gridlink entry points (gridlink.cu)
// scale: power of two applied to the sorted copy of the positions (1 unless the fast float kernel runs)
int cfb_gridlink_box_set(ParticleSet &S, const cfb_box_lattice *lat, const int sub[3], double scale);
int cfb_gridlink_theta_set(ParticleSet &S, const cfb_theta_lattice *lat, int64_t ncells /* reference cells */);
// counts-in-spheres buffers (spheres.cu)
void cfb_spheres_release();
// pair kernels (pairs_generic.cu)
int cfb_launch_pairs_generic(const cfb_binning *bin, const PairParams &P, int prec, bool list_mode);
// per-pair-sum kernel (pairs_sum.cu): 0 launched, -1 not applicable (the caller falls back to the generic kernel), 1 error
int cfb_launch_pairs_sum(const cfb_binning *bin, const PairParams &P, int prec, bool list_mode);
// fast 1-D kernel (pairs_fast.cu); P.edges / P.wrap / P.pimax are in the kernel's scaled units
int cfb_launch_pairs_fast(const cfb_binning *bin, const PairParams &P, int prec, bool list_mode);
