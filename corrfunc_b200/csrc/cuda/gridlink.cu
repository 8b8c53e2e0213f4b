// gridlink.cu -- GPU counting sort of particles into lattice cells.
//
// Replaces gridlink_DOUBLE (utils/gridlink_impl.c.src:65-436) and the particle-assignment part of
// gridlink_mocks_theta_ra_dec_DOUBLE (utils/gridlink_mocks_impl.c.src:1246-1350):
//   k_cellindex_*  cell index per particle (the reference's truncating formula, bit-for-bit)
//                  + per-cell histogram; for small sets the atomic's return value is the particle's rank in its cell
//   k_scan_sums / k_scan_cells   exclusive scan of the (padded) cell counts -> cell start offsets, and of the
//                  per-cell tile counts -> tile ids (<= 64 blocks, one segment of cells each)
//   k_partition / k_place   sets of 4 M points and more: bucket partition of {x, y, z, cell} records staged in shared
//                  memory, then in-order placement at start[cell] + arrival rank (see "two-pass scatter" below)
//   k_scatter      smaller sets: SoA scatter x|y|z|w into cell order, one pass
//   k_bounds_pad   per-cell min/max bounds (the reference's xbounds/ybounds/zbounds/ra_bounds)
//                  and NaN fill of the padding slots
//   k_fill_tiles   (cell, offset) table of primary tiles
// All of it is HBM-bound byte shuffling; DESIGN.md section 3a has the per-kernel times and DRAM traffic at 100 M points.
#include <stdlib.h>

#include "cfb_internal.cuh"

template <typename T>
struct BoxGeomT {
    T lo[3], inv[3];
    int n[3], s[3], ng[3];
};

// Reference cell index: ix=(int)((X-xmin)*xinv); if(ix>nmesh-1) ix--  (gridlink_impl.c.src:165-181).
// The fine index subdivides the reference cell by the fractional position; it only steers pruning
// (cell bounds are measured from the particles), never the results.
template <typename T>
__device__ __forceinline__ bool fine_coord(const T v, const T lo, const T inv, const int n, const int s, int &g)
{
    const T u = (v - lo) * inv;  // compiled with -fmad=false: one rounded subtract, one rounded multiply
    if (!(u == u)) return false;
    int i = (int)u;  // truncation toward zero, like the C cast
    if (i > n - 1) i--;
    if (i < 0 || i >= n) return false;
    int sub = 0;
    if (s > 1) {
        sub = (int)((u - (T)i) * (T)s);
        sub = sub < 0 ? 0 : (sub > s - 1 ? s - 1 : sub);
    }
    g = i * s + sub;
    return true;
}

template <typename T>
__global__ void k_cellindex_box(const int64_t n, const T *__restrict__ x, const T *__restrict__ y,
                                const T *__restrict__ z, const BoxGeomT<T> G, int *__restrict__ cidx,
                                int *__restrict__ rank, int *__restrict__ count, unsigned long long *oob)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int gx, gy, gz;
    const bool ok = fine_coord(x[i], G.lo[0], G.inv[0], G.n[0], G.s[0], gx) &
                    fine_coord(y[i], G.lo[1], G.inv[1], G.n[1], G.s[1], gy) &
                    fine_coord(z[i], G.lo[2], G.inv[2], G.n[2], G.s[2], gz);
    if (!ok) {
        atomicAdd(oob, 1ULL);
        cidx[i] = -1;
        return;
    }
    const int c = (gx * G.ng[1] + gy) * G.ng[2] + gz;
    cidx[i] = c;
    if (rank) rank[i] = atomicAdd(&count[c], 1);
    else atomicAdd(&count[c], 1);  // the two-pass scatter ranks by arrival itself: no return value, no second array
}

// DDtheta lattice: idec=(int)(ngrid_dec*(DEC-dec_min)*inv_dec_diff); if(idec>=ngrid_dec) idec--;
// ira=(int)(ngrid_ra[idec]*(RA-ra_min)*inv_ra_diff); if(ira>=ngrid_ra[idec]) ira--
// (gridlink_mocks_impl.c.src:1249-1263)
template <typename T>
__global__ void k_cellindex_theta(const int64_t n, const T *__restrict__ ra, const T *__restrict__ dec,
                                  const int ngrid_dec, const int *__restrict__ ngrid_ra,
                                  const int *__restrict__ ra_off, const T dec_min, const T inv_dec_diff,
                                  const T ra_min, const T inv_ra_diff, const int sub, int *__restrict__ cidx,
                                  int *__restrict__ rank, int *__restrict__ count, unsigned long long *oob)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    // reference cell: gridlink_mocks_impl.c.src:1249-1263
    const T ud = (T)ngrid_dec * (dec[i] - dec_min) * inv_dec_diff;
    int idec = (int)ud;
    if (idec >= ngrid_dec) idec--;
    bool ok = (ud == ud) && idec >= 0 && idec < ngrid_dec;
    int ira = 0, nra = 1;
    T ur = 0;
    if (ok) {
        nra = ngrid_ra[idec];
        ur = (T)nra * (ra[i] - ra_min) * inv_ra_diff;
        ira = (int)ur;
        if (ira >= nra) ira--;
        ok = (ur == ur) && ira >= 0 && ira < nra;
    }
    if (!ok) {
        atomicAdd(oob, 1ULL);
        cidx[i] = -1;
        return;
    }
    // device-only refinement: position inside the reference cell -> one of sub x sub fine cells.  It only groups
    // particles (tighter bounding boxes, smaller work units); which reference cells are paired is untouched.
    int sd = (int)((ud - (T)idec) * (T)sub), sr = (int)((ur - (T)ira) * (T)sub);
    sd = sd < 0 ? 0 : (sd >= sub ? sub - 1 : sd);
    sr = sr < 0 ? 0 : (sr >= sub ? sub - 1 : sr);
    const int c = ((ra_off[idec] + ira) * sub + sd) * sub + sr;
    cidx[i] = c;
    rank[i] = atomicAdd(&count[c], 1);
}

// Exclusive scan of padded counts (-> start) and tile counts (-> tstart) in two launches: every block owns `seg`
// consecutive cells; k_scan_sums leaves the segment totals, k_scan_cells adds up the totals of the segments before its
// own and scans its segment (1024 threads x 8 cells at a time, warp shuffles).  totals[0] = padded particle total,
// totals[1] = tile total.  (One block over all cells took 1.15 ms of config 5's gridlink for 3.5 MB of counts.)
#define CFB_SCAN_ITEMS 8
__global__ void __launch_bounds__(1024)
k_scan_sums(const int64_t ncells, const int64_t seg, const int *__restrict__ count, long long *__restrict__ bsum)
{
    __shared__ long long s_a[32], s_b[32];
    const int64_t lo = blockIdx.x * seg, hi = min(ncells, lo + seg);
    long long sa = 0, sb = 0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += 1024) {
        const int cnt = count[i];
        sa += (cnt + CFB_PAD - 1) / CFB_PAD * CFB_PAD;
        sb += (cnt + CFB_TILE - 1) / CFB_TILE;
    }
    for (int off = 16; off > 0; off >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, off);
        sb += __shfl_xor_sync(0xffffffffu, sb, off);
    }
    if ((threadIdx.x & 31) == 0) s_a[threadIdx.x >> 5] = sa, s_b[threadIdx.x >> 5] = sb;
    __syncthreads();
    if (threadIdx.x < 32) {
        sa = s_a[threadIdx.x], sb = s_b[threadIdx.x];
        for (int off = 16; off > 0; off >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, off);
            sb += __shfl_xor_sync(0xffffffffu, sb, off);
        }
        if (threadIdx.x == 0) bsum[2 * blockIdx.x] = sa, bsum[2 * blockIdx.x + 1] = sb;
    }
}

__global__ void __launch_bounds__(1024)
k_scan_cells(const int64_t ncells, const int64_t seg, const int *__restrict__ count, int *__restrict__ start,
             int *__restrict__ tstart, const long long *__restrict__ bsum, long long *totals)
{
    __shared__ long long s_a[1024], s_b[1024];
    __shared__ long long carry_a, carry_b;
    {
        // totals of the segments before this one
        long long pa = 0, pb = 0;
        for (int j = threadIdx.x; j < (int)blockIdx.x; j += 1024) pa += bsum[2 * j], pb += bsum[2 * j + 1];
        for (int off = 16; off > 0; off >>= 1) {
            pa += __shfl_xor_sync(0xffffffffu, pa, off);
            pb += __shfl_xor_sync(0xffffffffu, pb, off);
        }
        if ((threadIdx.x & 31) == 0) s_a[threadIdx.x >> 5] = pa, s_b[threadIdx.x >> 5] = pb;
        __syncthreads();
        if (threadIdx.x == 0) {
            pa = 0, pb = 0;
            for (int w = 0; w < 32; w++) pa += s_a[w], pb += s_b[w];
            carry_a = pa;
            carry_b = pb;
        }
        __syncthreads();
    }
    const int ITEMS = CFB_SCAN_ITEMS;
    const int64_t lo = blockIdx.x * seg, hi = min(ncells, lo + seg);
    for (int64_t base = lo; base < hi; base += 1024 * ITEMS) {
        long long va[ITEMS], vb[ITEMS], sa = 0, sb = 0;
        const int64_t i0 = base + (int64_t)threadIdx.x * ITEMS;
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            const int64_t i = i0 + k;
            const int cnt = i < hi ? count[i] : 0;
            va[k] = (cnt + CFB_PAD - 1) / CFB_PAD * CFB_PAD;
            vb[k] = (cnt + CFB_TILE - 1) / CFB_TILE;
            sa += va[k];
            sb += vb[k];
        }
        // inclusive scan of the 1024 thread sums: shuffles within the warps, then over the 32 warp totals
        {
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
            long long ia = sa, ib = sb;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const long long ta = __shfl_up_sync(0xffffffffu, ia, off), tb = __shfl_up_sync(0xffffffffu, ib, off);
                if (lane >= off) ia += ta, ib += tb;
            }
            if (lane == 31) s_a[wid] = ia, s_b[wid] = ib;  // warp totals
            __syncthreads();
            if (wid == 0) {
                long long wa = s_a[lane], wb = s_b[lane];
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const long long ta = __shfl_up_sync(0xffffffffu, wa, off), tb = __shfl_up_sync(0xffffffffu, wb, off);
                    if (lane >= off) wa += ta, wb += tb;
                }
                s_a[32 + lane] = wa, s_b[32 + lane] = wb;  // inclusive scan of the warp totals
            }
            __syncthreads();
            const long long pa = wid ? s_a[32 + wid - 1] : 0, pb = wid ? s_b[32 + wid - 1] : 0;
            __syncthreads();
            s_a[threadIdx.x] = ia + pa;
            s_b[threadIdx.x] = ib + pb;
            __syncthreads();
        }
        long long ea = carry_a + s_a[threadIdx.x] - sa, eb = carry_b + s_b[threadIdx.x] - sb;
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            const int64_t i = i0 + k;
            if (i < hi) {
                start[i] = (int)ea;
                tstart[i] = (int)eb;
            }
            ea += va[k];
            eb += vb[k];
        }
        __syncthreads();
        if (threadIdx.x == 1023) {
            carry_a += s_a[1023];
            carry_b += s_b[1023];
        }
        __syncthreads();
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        totals[0] = carry_a;
        totals[1] = carry_b;
    }
}

template <typename T>
__global__ void k_scatter(const int64_t n, const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z,
                          const T *__restrict__ w, const int *__restrict__ cidx, const int *__restrict__ rank,
                          const int *__restrict__ start, T *__restrict__ xs, T *__restrict__ ys, T *__restrict__ zs,
                          T *__restrict__ ws, const T scale)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cidx[i];
    if (c < 0) return;
    const int p = start[c] + rank[i];
    // scale is a power of two (exact): the fast float kernel works on pre-scaled positions
    xs[p] = x[i] * scale;
    ys[p] = y[i] * scale;
    zs[p] = z[i] * scale;
    if (w) ws[p] = w[i];
}

// ---- two-pass scatter for large sets -------------------------------------------------------------------------------
// k_scatter writes every particle's three 4-byte coordinates to random places of a 1.2 GB target (config 5): the L2
// cannot hold the partially written 32-byte sectors, and ncu counts 8.9 GB read + 8.4 GB written for 1.2 GB of payload
// (10.6 of the 13.3 ms of the whole gridlink).  Large sets therefore go through a coarse partition first:
//   pass 1 (k_partition): buckets of 2^shift consecutive cells (<= 512 buckets).  A block takes a chunk of particles
//          (64 KB of records), ranks them within their buckets with shared-memory atomics, orders the records by bucket
//          in shared memory (one block scan of the bucket counts), reserves one run per bucket with a single global
//          atomic and writes the staged records out in order -- consecutive threads, consecutive 16-byte records of one
//          run -- into the bucket's region of a temporary array that has the layout of the final one.  A record is
//          {x, y, z, cell} (without weights: one 16-byte store per particle, nothing else) or {x, y, z, w} + the cell in a
//          second array;
//   pass 2 (k_place): walks the temporary array in order -- at any time the blocks in flight work on a few neighbouring
//          buckets, whose final positions (a few MB) stay in L2 until their sectors are complete -- and places every
//          record at start[cell] + arrival rank.  Which slots of the temporary array hold a record follows from the
//          position alone (a bucket's records are contiguous from its base; the padding is at its end): nothing is cleared.
// The order of the particles inside a cell is, as before, the order of arrival.
#define CFB_PART_MAXB 512
template <typename T>
struct alignas(4 * sizeof(T)) Rec4 {
    T x, y, z, w;
};
template <typename T>
__host__ __device__ constexpr int part_chunk() { return 65536 / (int)sizeof(Rec4<T>); }  // records per block and round: 4096 / 2048
__device__ __forceinline__ float cell_as_real(const int c, float) { return __int_as_float(c); }
__device__ __forceinline__ double cell_as_real(const int c, double) { return __longlong_as_double((long long)c); }
__device__ __forceinline__ int real_as_cell(const float v) { return __float_as_int(v); }
__device__ __forceinline__ int real_as_cell(const double v) { return (int)__double_as_longlong(v); }
template <typename T, bool WGT>
__host__ __device__ constexpr size_t part_smem() { return (size_t)part_chunk<T>() * (sizeof(Rec4<T>) + 2 + (WGT ? 4 : 0)); }

template <typename T, bool WGT>
__global__ void __launch_bounds__(256)
k_partition(const int64_t n, const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z,
            const T *__restrict__ w, const int *__restrict__ cidx, const int shift, const int nbuckets,
            const int *__restrict__ start, int *__restrict__ bcursor, Rec4<T> *__restrict__ rec, int *__restrict__ rcid,
            const T scale)
{
    constexpr int CH = part_chunk<T>(), PER = CH / 256;
    extern __shared__ __align__(16) unsigned char part_sm[];
    Rec4<T> *s_rec = (Rec4<T> *)part_sm;
    unsigned short *s_bk = (unsigned short *)(part_sm + (size_t)CH * sizeof(Rec4<T>));
    int *s_cid = (int *)(part_sm + (size_t)CH * (sizeof(Rec4<T>) + 2));
    __shared__ int s_cnt[CFB_PART_MAXB], s_off[CFB_PART_MAXB], s_base[CFB_PART_MAXB], s_wsum[8], s_total;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int64_t chunk = blockIdx.x; chunk * CH < n; chunk += gridDim.x) {
        for (int b = tid; b < CFB_PART_MAXB; b += 256) s_cnt[b] = 0;
        __syncthreads();
        int c[PER], lr[PER];
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int64_t i = chunk * CH + k * 256 + tid;
            c[k] = i < n ? __ldcs(cidx + i) : -1;
            lr[k] = c[k] >= 0 ? atomicAdd(&s_cnt[c[k] >> shift], 1) : 0;
        }
        __syncthreads();
        {
            // exclusive scan of the 512 bucket counts (two per thread) -> the buckets' offsets in the staging area;
            // one global atomic per non-empty bucket reserves its run
            const int c0 = s_cnt[2 * tid], c1 = s_cnt[2 * tid + 1];
            int incl = c0 + c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_wsum[wid] = incl;
            __syncthreads();
            int pre = 0;
#pragma unroll
            for (int v = 0; v < 8; v++) pre += v < wid ? s_wsum[v] : 0;
            const int excl = pre + incl - (c0 + c1);
            s_off[2 * tid] = excl;
            s_off[2 * tid + 1] = excl + c0;
            if (c0) s_base[2 * tid] = start[(int64_t)(2 * tid) << shift] + atomicAdd(&bcursor[2 * tid], c0);
            if (c1) s_base[2 * tid + 1] = start[(int64_t)(2 * tid + 1) << shift] + atomicAdd(&bcursor[2 * tid + 1], c1);
            if (tid == 255) s_total = pre + incl;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PER; k++) {
            if (c[k] < 0) continue;
            const int64_t i = chunk * CH + k * 256 + tid;
            const int bk = c[k] >> shift;
            const int slot = s_off[bk] + lr[k];
            Rec4<T> r;
            // scale is a power of two (exact): the fast float kernel works on pre-scaled positions
            r.x = __ldcs(x + i) * scale, r.y = __ldcs(y + i) * scale, r.z = __ldcs(z + i) * scale;
            r.w = WGT ? __ldcs(w + i) : cell_as_real(c[k], (T)0);
            s_rec[slot] = r;
            s_bk[slot] = (unsigned short)bk;
            if (WGT) s_cid[slot] = c[k];
        }
        __syncthreads();
        const int total = s_total;
        for (int idx = tid; idx < total; idx += 256) {
            const int bk = s_bk[idx];
            const int p = s_base[bk] + (idx - s_off[bk]);
            rec[p] = s_rec[idx];
            if (WGT) rcid[p] = s_cid[idx];
        }
        __syncthreads();
    }
}

// cur[c] starts as start[c] (a device copy): the atomic's return value is the record's final position.  (1024 threads per
// block over the same window of records measured 0.7 ms slower than 256 on config 5: the pass is bound by the L2's
// handling of the scattered 4-byte stores, not by latency.)
#ifndef CFB_PLACE_THREADS
#define CFB_PLACE_THREADS 256
#endif
// streaming (evict-first) load of a record: the 1.6 GB of records pass through the L2 once and must not push out the
// partially written target sectors
__device__ __forceinline__ Rec4<float> load_rec_cs(const Rec4<float> *p)
{
    const float4 v = __ldcs(reinterpret_cast<const float4 *>(p));
    Rec4<float> r;
    r.x = v.x, r.y = v.y, r.z = v.z, r.w = v.w;
    return r;
}
__device__ __forceinline__ Rec4<double> load_rec_cs(const Rec4<double> *p)
{
    const double2 a = __ldcs(reinterpret_cast<const double2 *>(p)), b = __ldcs(reinterpret_cast<const double2 *>(p) + 1);
    Rec4<double> r;
    r.x = a.x, r.y = a.y, r.z = b.x, r.w = b.y;
    return r;
}
template <typename T, bool WGT>
__global__ void __launch_bounds__(CFB_PLACE_THREADS)
k_place(const int64_t npad, const Rec4<T> *__restrict__ rec, const int *__restrict__ rcid, const int *__restrict__ start,
        const int shift, const int nbuckets, const int *__restrict__ bcount, int *__restrict__ cur, T *__restrict__ xs,
        T *__restrict__ ys, T *__restrict__ zs, T *__restrict__ ws)
{
    constexpr int CH = 4096, PER = CH / CFB_PLACE_THREADS;
    __shared__ int s_base[CFB_PART_MAXB], s_cnt[CFB_PART_MAXB];
    for (int b = threadIdx.x; b < nbuckets; b += CFB_PLACE_THREADS) {
        s_base[b] = start[(int64_t)b << shift];
        s_cnt[b] = bcount[b];
    }
    __syncthreads();
    for (int64_t chunk = blockIdx.x; chunk * CH < npad; chunk += gridDim.x) {
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int64_t p = chunk * CH + k * CFB_PLACE_THREADS + threadIdx.x;
            if (p >= npad) continue;
            int lo = 0, hi = nbuckets;  // last bucket whose base is <= p (of equal bases the last: the others are empty)
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if ((int64_t)s_base[mid] <= p) lo = mid; else hi = mid;
            }
            if (p - s_base[lo] >= s_cnt[lo]) continue;  // padding at the end of the bucket's region
            const Rec4<T> r = load_rec_cs(rec + p);
            const int c = WGT ? __ldcs(rcid + p) : real_as_cell(r.w);
            const int q = atomicAdd(&cur[c], 1);
            xs[q] = r.x;
            ys[q] = r.y;
            zs[q] = r.z;
            if (WGT) ws[q] = r.w;
        }
    }
}

// One warp per cell: min/max bounds of x,y,z (and of a fourth, unsorted-by-value array `ra` given
// through the particle permutation) + NaN padding.
template <typename T>
__global__ void k_bounds_pad(const int64_t ncells, const int *__restrict__ count, const int *__restrict__ start,
                             T *__restrict__ xs, T *__restrict__ ys, T *__restrict__ zs, T *__restrict__ ws,
                             T *__restrict__ bounds)
{
    const int64_t cell = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (cell >= ncells) return;
    const int n = count[cell], s = start[cell];
    const T big = sizeof(T) == 4 ? (T)3.402823466e+38F : (T)1.7976931348623157e+308;
    T lo[3] = {big, big, big}, hi[3] = {-big, -big, -big};
    for (int k = lane; k < n; k += 32) {
        const T v[3] = {xs[s + k], ys[s + k], zs[s + k]};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            lo[a] = v[a] < lo[a] ? v[a] : lo[a];
            hi[a] = v[a] > hi[a] ? v[a] : hi[a];
        }
    }
#pragma unroll
    for (int a = 0; a < 3; a++)
        for (int off = 16; off > 0; off >>= 1) {
            const T l2 = __shfl_xor_sync(0xffffffffu, lo[a], off), h2 = __shfl_xor_sync(0xffffffffu, hi[a], off);
            lo[a] = l2 < lo[a] ? l2 : lo[a];
            hi[a] = h2 > hi[a] ? h2 : hi[a];
        }
    if (lane == 0) {
        T *b = bounds + cell * CFB_NB;
        for (int a = 0; a < 3; a++) {
            b[2 * a] = lo[a];
            b[2 * a + 1] = hi[a];
        }
    }
    const int npad = (n + CFB_PAD - 1) / CFB_PAD * CFB_PAD;
    const T nanv = sizeof(T) == 4 ? (T)__int_as_float(0x7fc00000) : (T)__longlong_as_double(0x7ff8000000000000LL);
    if (lane < npad - n) {
        xs[s + n + lane] = nanv;
        ys[s + n + lane] = nanv;
        zs[s + n + lane] = nanv;
        if (ws) ws[s + n + lane] = (T)0;
    }
}

// RA bounds per theta cell (the reference keeps ra_bounds of each cell's particles,
// gridlink_mocks_impl.c.src:1336-1350); RA is not carried into the sorted SoA, so use atomics on an
// order-preserving integer image of the value.
template <typename T>
__global__ void k_ra_bounds_init(const int64_t ncells, T *bounds)
{
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const T big = sizeof(T) == 4 ? (T)3.402823466e+38F : (T)1.7976931348623157e+308;
    bounds[c * CFB_NB + 6] = big;
    bounds[c * CFB_NB + 7] = -big;
}
__device__ __forceinline__ void atomic_min_real(float *a, float v)
{  // valid for any sign: compare as signed ints when >=0, as unsigned reversed when negative
    if (v >= 0) atomicMin((int *)a, __float_as_int(v));
    else atomicMax((unsigned int *)a, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_real(float *a, float v)
{
    if (v >= 0) atomicMax((int *)a, __float_as_int(v));
    else atomicMin((unsigned int *)a, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_real(double *a, double v)
{
    if (v >= 0) atomicMin((long long *)a, __double_as_longlong(v));
    else atomicMax((unsigned long long *)a, (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ void atomic_max_real(double *a, double v)
{
    if (v >= 0) atomicMax((long long *)a, __double_as_longlong(v));
    else atomicMin((unsigned long long *)a, (unsigned long long)__double_as_longlong(v));
}
template <typename T>
__global__ void k_ra_bounds(const int64_t n, const T *__restrict__ ra, const int *__restrict__ cidx, T *bounds)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cidx[i];
    if (c < 0) return;
    atomic_min_real(&bounds[(int64_t)c * CFB_NB + 6], ra[i]);
    atomic_max_real(&bounds[(int64_t)c * CFB_NB + 7], ra[i]);
}

__global__ void k_fill_tiles(const int64_t ncells, const int *__restrict__ count, const int *__restrict__ tstart,
                             int *__restrict__ tile_cell, int *__restrict__ tile_off)
{
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const int n = count[c];
    const int nt = (n + CFB_TILE - 1) / CFB_TILE;
    const int t0 = tstart[c];
    for (int t = 0; t < nt; t++) {
        tile_cell[t0 + t] = (int)c;
        tile_off[t0 + t] = t * CFB_TILE;
    }
}

static inline unsigned int nblocks(int64_t n, int bs) { return (unsigned int)((n + bs - 1) / bs); }

static int64_t sort2_min()
{
    static int64_t v = -1;  // sets at least this large take the two-pass scatter
    if (v < 0) {
        const char *e = getenv("CORRFUNC_B200_SORT2_MIN");
        v = (e && *e) ? atoll(e) : 4000000;
    }
    return v;
}

template <typename T, bool WGT>
static int scatter_two_pass(Ctx &c, ParticleSet &S, int64_t ncells, double scale)
{
    int shift = 0;
    while (((ncells + ((int64_t)1 << shift) - 1) >> shift) > CFB_PART_MAXB) shift++;
    const int nbuckets = (int)((ncells + ((int64_t)1 << shift) - 1) >> shift);
    if (cfb_ensure(c.sort_rec, (size_t)S.npad * sizeof(Rec4<T>))) return 1;
    if (WGT && cfb_ensure(c.sort_cid, (size_t)S.npad * 4)) return 1;
    if (cfb_ensure(c.sort_cur, (size_t)(ncells + CFB_PART_MAXB) * 4)) return 1;
    int *bcursor = (int *)c.sort_cur.p, *ccursor = (int *)c.sort_cur.p + CFB_PART_MAXB;
    CK(cudaMemsetAsync(bcursor, 0, (size_t)CFB_PART_MAXB * 4, c.stream));
    CK(cudaMemcpyAsync(ccursor, S.start.p, (size_t)ncells * 4, cudaMemcpyDeviceToDevice, c.stream));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto kp = k_partition<T, WGT>;
    constexpr size_t sm = part_smem<T, WGT>();
    CK(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const unsigned g1 = (unsigned)min((int64_t)nblocks(S.n, part_chunk<T>()), (int64_t)sms * 8);
    kp<<<g1, 256, sm, c.stream>>>(S.n, (const T *)S.raw[0], (const T *)S.raw[1], (const T *)S.raw[2], (const T *)S.raw[3],
                                  (const int *)S.cidx.p, shift, nbuckets, (const int *)S.start.p, bcursor,
                                  (Rec4<T> *)c.sort_rec.p, (int *)c.sort_cid.p, (T)scale);
    // few blocks in flight: their targets (a few neighbouring buckets) must stay L2-resident until complete
    static int place_blocks = -1;
    if (place_blocks < 0) {
        const char *e = getenv("CORRFUNC_B200_PLACE_BLOCKS");
        place_blocks = (e && atoi(e) > 0) ? atoi(e) : 2;
    }
    const unsigned g2 = (unsigned)min((int64_t)nblocks(S.npad, 4096), (int64_t)sms * place_blocks);
    k_place<T, WGT><<<g2, CFB_PLACE_THREADS, 0, c.stream>>>(S.npad, (const Rec4<T> *)c.sort_rec.p, (const int *)c.sort_cid.p,
                                               (const int *)S.start.p, shift, nbuckets, bcursor, ccursor,
                                               (T *)S.sorted[0].p, (T *)S.sorted[1].p, (T *)S.sorted[2].p,
                                               WGT ? (T *)S.sorted[3].p : nullptr);
    c.launches += 2;
    CK(cudaGetLastError());
    return 0;
}

// have_rank: the cell-index kernel left every particle's arrival rank in S.rank (needed by the one-pass scatter)
template <typename T>
static int finish_sort(Ctx &c, ParticleSet &S, int64_t ncells, double scale, bool have_rank)
{
    // scan -> starts / tile ids
    if (cfb_ensure(S.start, (size_t)ncells * 4)) return 1;
    if (cfb_ensure(S.tstart, (size_t)ncells * 4)) return 1;
    // scratch: [0] out-of-bounds counter (written by the cell-index kernel), +64: {padded total, tile total},
    // +1024: segment totals of the scan (2 x 64)
    long long *totals = (long long *)((char *)c.scratch.p + 64), *bsum = (long long *)((char *)c.scratch.p + 1024);
    {
        const int64_t per = 1024 * CFB_SCAN_ITEMS;
        int64_t seg = (ncells + 63) / 64;
        seg = (seg + per - 1) / per * per;
        const unsigned g = (unsigned)((ncells + seg - 1) / seg);  // <= 64
        k_scan_sums<<<g, 1024, 0, c.stream>>>(ncells, seg, (const int *)S.count.p, bsum);
        k_scan_cells<<<g, 1024, 0, c.stream>>>(ncells, seg, (const int *)S.count.p, (int *)S.start.p, (int *)S.tstart.p,
                                               bsum, totals);
        c.launches += 2;
    }
    CK(cudaGetLastError());
    long long *h = (long long *)c.pinned;
    CK(cudaMemcpyAsync(h, c.scratch.p, 64 + 16, cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    const unsigned long long n_oob = *(unsigned long long *)h;
    if (n_oob != 0)
        return cfb_fail("%llu particles are out of bounds. Check periodic wrapping?", n_oob);  // gridlink_impl.c.src:183
    S.npad = h[8];
    S.ntiles = h[9];
    if (S.npad >= 2147483647LL) return cfb_fail("padded particle count %lld exceeds 2^31", (long long)S.npad);
    const size_t sb = (size_t)(S.npad > 0 ? S.npad : 1) * sizeof(T);
    const bool hasw = S.raw[3] != nullptr;
    for (int a = 0; a < (hasw ? 4 : 3); a++)
        if (cfb_ensure(S.sorted[a], sb)) return 1;
    if (S.n > 0 && (S.n >= sort2_min() || !have_rank)) {
        if (hasw ? scatter_two_pass<T, true>(c, S, ncells, scale) : scatter_two_pass<T, false>(c, S, ncells, scale)) return 1;
    } else if (S.n > 0) {
        k_scatter<T><<<nblocks(S.n, 256), 256, 0, c.stream>>>(
            S.n, (const T *)S.raw[0], (const T *)S.raw[1], (const T *)S.raw[2], (const T *)S.raw[3],
            (const int *)S.cidx.p, (const int *)S.rank.p, (const int *)S.start.p, (T *)S.sorted[0].p,
            (T *)S.sorted[1].p, (T *)S.sorted[2].p, hasw ? (T *)S.sorted[3].p : nullptr, (T)scale);
        c.launches++;
        CK(cudaGetLastError());
    }
    k_bounds_pad<T><<<nblocks(ncells * 32, 256), 256, 0, c.stream>>>(
        ncells, (const int *)S.count.p, (const int *)S.start.p, (T *)S.sorted[0].p, (T *)S.sorted[1].p,
        (T *)S.sorted[2].p, hasw ? (T *)S.sorted[3].p : nullptr, (T *)S.bounds.p);
    c.launches++;
    CK(cudaGetLastError());
    if (cfb_ensure(S.tile_cell, (size_t)(S.ntiles > 0 ? S.ntiles : 1) * 4)) return 1;
    if (cfb_ensure(S.tile_off, (size_t)(S.ntiles > 0 ? S.ntiles : 1) * 4)) return 1;
    k_fill_tiles<<<nblocks(ncells, 256), 256, 0, c.stream>>>(ncells, (const int *)S.count.p, (const int *)S.tstart.p,
                                                            (int *)S.tile_cell.p, (int *)S.tile_off.p);
    c.launches++;
    CK(cudaGetLastError());
    S.ncells = ncells;
    S.gridded = true;
    return 0;
}

template <typename T>
static int gridlink_box_T(Ctx &c, ParticleSet &S, const cfb_box_lattice *lat, const int sub[3], double scale)
{
    BoxGeomT<T> G;
    int64_t ncells = 1;
    for (int k = 0; k < 3; k++) {
        G.lo[k] = (T)lat->lo[k];
        G.inv[k] = (T)lat->inv[k];
        G.n[k] = lat->nmesh[k];
        G.s[k] = sub[k];
        G.ng[k] = lat->nmesh[k] * sub[k];
        ncells *= G.ng[k];
    }
    if (ncells >= 2147483647LL) return cfb_fail("fine lattice too large (%lld cells)", (long long)ncells);
    if (cfb_ensure(c.scratch, 4096)) return 1;
    if (cfb_ensure(S.count, (size_t)ncells * 4)) return 1;
    if (cfb_ensure(S.bounds, (size_t)ncells * CFB_NB * sizeof(T))) return 1;
    if (cfb_ensure(S.cidx, (size_t)(S.n > 0 ? S.n : 1) * 4)) return 1;
    // large sets take the two-pass scatter, which ranks by arrival itself: no rank array (400 MB at 100 M points)
    const bool want_rank = S.n < sort2_min();
    if (want_rank && cfb_ensure(S.rank, (size_t)(S.n > 0 ? S.n : 1) * 4)) return 1;
    CK(cudaMemsetAsync(S.count.p, 0, (size_t)ncells * 4, c.stream));
    CK(cudaMemsetAsync(c.scratch.p, 0, 256, c.stream));
    if (S.n > 0) {
        k_cellindex_box<T><<<nblocks(S.n, 256), 256, 0, c.stream>>>(
            S.n, (const T *)S.raw[0], (const T *)S.raw[1], (const T *)S.raw[2], G, (int *)S.cidx.p,
            want_rank ? (int *)S.rank.p : nullptr, (int *)S.count.p, (unsigned long long *)c.scratch.p);
        c.launches++;
        CK(cudaGetLastError());
    }
    return finish_sort<T>(c, S, ncells, scale, want_rank);
}

int cfb_gridlink_box_set(ParticleSet &S, const cfb_box_lattice *lat, const int sub[3], double scale)
{
    Ctx &c = cfb_ctx();
    // the sorted form of an unchanged catalogue (catalogue cache) in an unchanged lattice is still good
    unsigned char sig[sizeof(S.grid_sig)];
    static_assert(sizeof(cfb_box_lattice) + 3 * sizeof(int) + sizeof(double) <= sizeof(S.grid_sig), "signature buffer");
    memset(sig, 0, sizeof(sig));
    memcpy(sig, lat, sizeof(*lat));
    memcpy(sig + sizeof(*lat), sub, 3 * sizeof(int));
    memcpy(sig + sizeof(*lat) + 3 * sizeof(int), &scale, sizeof(double));
    if (S.gridded && S.grid_sig_valid && memcmp(sig, S.grid_sig, sizeof(sig)) == 0) return 0;
    S.grid_sig_valid = false;
    const int rc = S.prec == 4 ? gridlink_box_T<float>(c, S, lat, sub, scale) : gridlink_box_T<double>(c, S, lat, sub, scale);
    if (rc == 0) {
        memcpy(S.grid_sig, sig, sizeof(sig));
        S.grid_sig_valid = true;
    }
    return rc;
}

template <typename T>
static int gridlink_theta_T(Ctx &c, ParticleSet &S, const cfb_theta_lattice *lat, int64_t ncells)
{
    if (!S.raw[4] || !S.raw[5]) return cfb_fail("theta gridlink needs RA and DEC on the device");
    if (cfb_ensure(c.scratch, 4096)) return 1;
    if (cfb_ensure(c.ngrid_ra, (size_t)lat->ngrid_dec * 4)) return 1;
    if (cfb_ensure(c.ra_off, (size_t)lat->ngrid_dec * 4)) return 1;
    int *h = (int *)c.pinned;
    int off = 0;
    for (int i = 0; i < lat->ngrid_dec; i++) {
        h[i] = lat->ngrid_ra[i];
        h[lat->ngrid_dec + i] = off;
        off += lat->ngrid_ra[i];
    }
    if (off != ncells) return cfb_fail("theta lattice: cell count mismatch (%d vs %lld)", off, (long long)ncells);
    const int sub = lat->sub > 0 ? lat->sub : 1;
    const int64_t nfine = ncells * sub * sub;
    if (nfine >= 2147483647LL / 2) return cfb_fail("theta lattice: too many fine cells (%lld)", (long long)nfine);
    CK(cudaMemcpyAsync(c.ngrid_ra.p, h, (size_t)lat->ngrid_dec * 4, cudaMemcpyHostToDevice, c.stream));
    CK(cudaMemcpyAsync(c.ra_off.p, h + lat->ngrid_dec, (size_t)lat->ngrid_dec * 4, cudaMemcpyHostToDevice, c.stream));
    if (cfb_ensure(S.count, (size_t)nfine * 4)) return 1;
    if (cfb_ensure(S.bounds, (size_t)nfine * CFB_NB * sizeof(T))) return 1;
    if (cfb_ensure(S.cidx, (size_t)(S.n > 0 ? S.n : 1) * 4)) return 1;
    if (cfb_ensure(S.rank, (size_t)(S.n > 0 ? S.n : 1) * 4)) return 1;
    CK(cudaMemsetAsync(S.count.p, 0, (size_t)nfine * 4, c.stream));
    CK(cudaMemsetAsync(c.scratch.p, 0, 256, c.stream));
    k_cellindex_theta<T><<<nblocks(S.n, 256), 256, 0, c.stream>>>(
        S.n, (const T *)S.raw[4], (const T *)S.raw[5], lat->ngrid_dec, (const int *)c.ngrid_ra.p,
        (const int *)c.ra_off.p, (T)lat->dec_min, (T)lat->inv_dec_diff, (T)lat->ra_min, (T)lat->inv_ra_diff, sub,
        (int *)S.cidx.p, (int *)S.rank.p, (int *)S.count.p, (unsigned long long *)c.scratch.p);
    c.launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c.stream));  // pinned staging reused by finish_sort
    if (finish_sort<T>(c, S, nfine, 1.0, true)) return 1;
    k_ra_bounds_init<T><<<nblocks(nfine, 256), 256, 0, c.stream>>>(nfine, (T *)S.bounds.p);
    k_ra_bounds<T><<<nblocks(S.n, 256), 256, 0, c.stream>>>(S.n, (const T *)S.raw[4], (const int *)S.cidx.p,
                                                           (T *)S.bounds.p);
    c.launches += 2;
    CK(cudaGetLastError());
    return 0;
}

int cfb_gridlink_theta_set(ParticleSet &S, const cfb_theta_lattice *lat, int64_t ncells)
{
    Ctx &c = cfb_ctx();
    S.grid_sig_valid = false;
    return S.prec == 4 ? gridlink_theta_T<float>(c, S, lat, ncells) : gridlink_theta_T<double>(c, S, lat, ncells);
}
