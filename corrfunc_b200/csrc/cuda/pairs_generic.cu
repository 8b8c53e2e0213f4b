// pairs_generic.cu -- the general pair-counting kernel: every statistic, weights and averages.
//
// Replaces the per-cell-pair CPU kernels (theory/DD/countpairs_kernels.c.src:25-279 and siblings)
// together with the cell-pair enumeration of generate_cell_pairs_DOUBLE
// (utils/gridlink_impl.c.src:439-625), which is done on the fly from lattice coordinates:
//
//   block  = one primary tile: up to 128 particles of one fine cell, one per thread, in registers
//   phase 1: the 128 threads test 128 candidate neighbour cells in parallel (periodic wrap, the
//            reference's first/second role filter, bounds-based pruning) and queue survivors
//   phase 2: per queued neighbour, its particles are staged through shared memory (coalesced SoA
//            loads) and every thread runs its primary against them with the reference's arithmetic:
//            same subtraction order (second - (first + wrap)), same FMA association per statistic,
//            comparisons against the squared bin edges, 2-D bin index evaluated in floating point.
//   histograms are privatised per block in shared memory and merged with 64-bit global atomics.
//
// Compiled with -fmad=false: an FMA appears exactly where the reference's AVX-512 kernels have one.
#include <math_constants.h>
#include <stdlib.h>

#include "cfb_internal.cuh"

#define GEN_CH 64   // secondaries staged per chunk, per warp (shared memory per block decides the occupancy)
#define GEN_WARPS (CFB_TILE / 32)
#define GEN_DIRZ 64  // queue-entry flag next to the 6-bit wrap code: primary and secondary lie in different reference cells

template <typename T>
__device__ __forceinline__ T fma_t(T a, T b, T c);
template <>
__device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template <>
__device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return __fma_rn(a, b, c); }
template <typename T>
__device__ __forceinline__ T sqrt_t(T a);
template <>
__device__ __forceinline__ float sqrt_t<float>(float a) { return __fsqrt_rn(a); }
template <>
__device__ __forceinline__ double sqrt_t<double>(double a) { return __dsqrt_rn(a); }
template <typename T>
__device__ __forceinline__ T divi_t(T a, T b);
template <>
__device__ __forceinline__ float divi_t<float>(float a, float b) { return __fdiv_rn(a, b); }
template <>
__device__ __forceinline__ double divi_t<double>(double a, double b) { return __ddiv_rn(a, b); }

// utils/fast_acos.h:57-101 (degree-8 estimate, |err| < 3.7e-9)
template <typename T>
__device__ __forceinline__ T fast_acos_t(const T x)
{
    const T xa = x < 0 ? -x : x;
    T poly = (T) + 7.1796493341480527e-04;
    poly = (T)-4.1160981058965262e-03 + poly * xa;
    poly = (T) + 1.1272900916992512e-02 + poly * xa;
    poly = (T)-2.0949278766238422e-02 + poly * xa;
    poly = (T) + 3.2683762943179318e-02 + poly * xa;
    poly = (T)-5.0625279962389413e-02 + poly * xa;
    poly = (T) + 8.9034700107934128e-02 + poly * xa;
    poly = (T)-2.1460143648688035e-01 + poly * xa;
    poly = (T) + 1.5707963267948966 + poly * xa;
    poly = poly * sqrt_t<T>((T)1.0 - xa);
    return (x < 0) ? (T)(3.14159265358979323846 - (double)poly) : poly;
}

// Shared-memory accumulation without compare-and-swap loops.  On sm_100a only the 32-bit integer add is a
// native shared-memory atomic (64-bit, float and double adds compile to ATOMS.CAST.SPIN loops), so a
// block keeps its counts as two and its sums as three 32-bit words per slot: sums are 96-bit fixed point,
// value * 2^k rounded to an integer below 2^60 in magnitude (k from the bin's upper edge, or from the largest
// weight product), added word by word with the carries of the returned old values.  Integer sums are
// exact and order independent, so the averages are reproducible from run to run.
__device__ __forceinline__ void add_count(unsigned *w0, unsigned *w1)
{
    if (atomicAdd(w0, 1u) == 0xffffffffu) atomicAdd(w1, 1u);
}
__device__ __forceinline__ void add_fixed96(unsigned *w0, unsigned *w1, unsigned *w2, const long long q)
{
    const unsigned lo = (unsigned)q, mid = (unsigned)((unsigned long long)q >> 32), hi = q < 0 ? 0xffffffffu : 0u;
    const unsigned old0 = atomicAdd(w0, lo);
    const unsigned long long t = (unsigned long long)mid + ((unsigned)(old0 + lo) < lo ? 1u : 0u);
    const unsigned m = (unsigned)t;
    unsigned h = hi + (unsigned)(t >> 32);
    if (m) {
        const unsigned old1 = atomicAdd(w1, m);
        h += (unsigned)(old1 + m) < m ? 1u : 0u;
    }
    if (h) atomicAdd(w2, h);
}
// value of a signed 96-bit word triple
__device__ __forceinline__ double fixed96_to_double(const unsigned w0, const unsigned w1, const unsigned w2)
{
    return (double)(int)w2 * 18446744073709551616.0 + (double)w1 * 4294967296.0 + (double)w0;
}
// 2^k such that |v| * 2^k < 2^59 for every |v| <= vmax
__device__ __forceinline__ double fixed_scale(const double vmax)
{
    int e = 0;
    if (vmax > 0.0 && vmax < 1.0e300) frexp(vmax, &e);  // vmax < 2^e
    return ldexp(1.0, 59 - e);
}

template <typename T>
struct GenShared {
    // staging buffers are per WARP: the warps of a tile walk the queue independently (block-wide barriers around a
    // shared buffer cost 13 % of all warp samples, the warps' accepted-pair work differs too much)
    T sx[GEN_WARPS][GEN_CH], sy[GEN_WARPS][GEN_CH], sz[GEN_WARPS][GEN_CH], sw[GEN_WARPS][GEN_CH];
    int q_cell[CFB_TILE];
    int q_code[CFB_TILE];
    int q_n;
};

// (explicit minimum-blocks launch bounds of 8, 6 or 5 were measured 15-35 % slower than leaving the choice to ptxas)
template <typename T, int MODE, bool AVG, bool WGT, bool LIST>
__global__ void __launch_bounds__(CFB_TILE)
k_pairs_generic(const PairParams P, const SetView<T> A, const SetView<T> B)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    {
        // interrupt (see cfb_abort_flag): blocks that start after the signal do nothing
        // every 64th block reads the host's flag (over PCIe) and latches it in device memory, all blocks read the latch
        __shared__ int s_abort;
        if (threadIdx.x == 0) {
            if ((blockIdx.x & 63) == 0 && P.abort && *P.abort) atomicExch(&P.counters[5], 1ULL);
            s_abort = *(volatile unsigned long long *)&P.counters[5] != 0ULL;
        }
        __syncthreads();
        if (s_abort) return;
    }
    // dynamic smem layout: edges | sep scale per edge (double) | npairs (2 words) | sum_sep (3 words) | sum_w (3 words)
    T *s_edges = (T *)smem_raw;
    const int nedges = P.nedges;
    size_t off = ((size_t)nedges * sizeof(T) + 15) & ~(size_t)15;
    double *s_scale = (double *)(smem_raw + off);
    off += (size_t)nedges * 8;
    unsigned *s_np = (unsigned *)(smem_raw + off);
    unsigned *s_sep = nullptr, *s_w = nullptr;
    const int64_t ns = P.nslots;
    if (P.hist_in_smem) {
        off += (size_t)ns * 8;
        if (AVG) {
            s_sep = (unsigned *)(smem_raw + off);
            off += (size_t)ns * 12;
        }
        if (WGT) s_w = (unsigned *)(smem_raw + off);
    }
    __shared__ GenShared<T> S;

    // one block per primary tile; across ranks the work is sharded by primary cell (cfb_owns_cell)
    const int64_t tile = blockIdx.x;
    if (tile >= P.ntiles) return;
    if (!cfb_owns_cell(P.tile_cell[tile], P.shard_rank, P.shard_n)) return;
    const int tid = threadIdx.x;

    for (int i = tid; i < nedges; i += CFB_TILE) {
        const T e = ((const T *)P.edges)[i];
        s_edges[i] = e;
        // largest separation a pair of the bin below edge i can have: sqrt(edge) (edges are squared), or the
        // angle in degrees whose cosine the edge is
        double vmax;
        if (MODE == CFB_THETA) vmax = acos(fmin(1.0, fmax(-1.0, (double)e))) * 57.29577951308232087679815481410517 + 1e-9;
        else vmax = sqrt(fmax(0.0, (double)e));
        s_scale[i] = fixed_scale(vmax);
    }
    double w_scale = 1.0;
    if (WGT) w_scale = fixed_scale(P.wmax[0] * P.wmax[1]);
    if (P.hist_in_smem)
        for (int64_t i = tid; i < ns; i += CFB_TILE) {
            s_np[i] = 0u;
            s_np[ns + i] = 0u;
            if (AVG) s_sep[i] = s_sep[ns + i] = s_sep[2 * ns + i] = 0u;
            if (WGT) s_w[i] = s_w[ns + i] = s_w[2 * ns + i] = 0u;
        }

    const int cellP = P.tile_cell[tile];
    const int toff = P.tile_off[tile];
    const int nP = A.count[cellP];
    const int startP = A.start[cellP];
    const int iloc = toff + tid;
    const bool valid = iloc < nP;
    const T nanv = sizeof(T) == 4 ? (T)CUDART_NAN_F : (T)CUDART_NAN;
    T xp = nanv, yp = nanv, zp = nanv, wp = 0;
    if (valid) {
        xp = A.x[startP + iloc];
        yp = A.y[startP + iloc];
        zp = A.z[startP + iloc];
        if (WGT) wp = A.w[startP + iloc];
    }
    // primary cell bounds (whole fine cell: conservative for the tile)
    double pb[6];
#pragma unroll
    for (int k = 0; k < 6; k++) pb[k] = (double)A.bounds[(int64_t)cellP * CFB_NB + k];

    // lattice coordinates of the primary fine cell and its reference cell
    int gx = 0, gy = 0, gz = 0;
    long long refA = 0;
    int ncand;
    int64_t list0 = 0;
    const int sub2 = LIST ? P.list_sub2 : 1;  // fine cells per reference cell of the RA/DEC lattice
    const int refP = LIST ? cellP / sub2 : 0;
    if (LIST) {
        list0 = P.list_off[refP];
        ncand = (int)(P.list_off[refP + 1] - list0) * sub2;
    } else {
        gz = cellP % P.g.ng[2];
        gy = (cellP / P.g.ng[2]) % P.g.ng[1];
        gx = cellP / (P.g.ng[2] * P.g.ng[1]);
        refA = ((long long)(gx / P.g.s[0]) * P.g.n[1] + gy / P.g.s[1]) * P.g.n[2] + gz / P.g.s[2];
        ncand = (2 * P.g.reach[0] + 1) * (2 * P.g.reach[1] + 1) * (2 * P.g.reach[2] + 1);
    }
    const int wy = 2 * P.g.reach[1] + 1, wz = 2 * P.g.reach[2] + 1;

    const T pimax = (T)P.pimax;
    const T inv_dpi = (T)P.inv_dpi, inv_dmu = (T)P.inv_dmu, sqr_mumax = (T)P.sqr_mumax;
    const T npi_p1 = (T)(P.npibin + 1), nmu_p1 = (T)(P.nmu_bins + 1);
    const T sqr_pimax = pimax * pimax;  // rppi mocks (countpairs_rp_pi_mocks_kernels.c.src:56-57)
    unsigned long long my_eval = 0, my_tp = 0;
    __syncthreads();
    const T e_lo = s_edges[0], e_hi = s_edges[nedges - 1];
    const T sqr_max_sep = e_hi + sqr_pimax;
    const int lane = tid & 31, wid = tid >> 5;
    T *sx = S.sx[wid], *sy = S.sy[wid], *sz = S.sz[wid], *sw = S.sw[wid];
    const bool warp_has_work = __any_sync(0xffffffffu, valid);

    for (int base = 0; base < ncand; base += CFB_TILE) {
        if (tid == 0) S.q_n = 0;
        __syncthreads();
        // ---------------- phase 1: candidate test, one candidate per thread ----------------
        const int cand = base + tid;
        if (cand < ncand) {
            int cellQ = -1, code = 0;
            bool keep = true;
            double offd[3] = {0.0, 0.0, 0.0};
            if (LIST) {
                const int li = cand / sub2;
                const int refQ = P.list_cells[list0 + li];
                cellQ = refQ * sub2 + (cand - li * sub2);
                if (P.autocorr && refQ == refP && cellQ > cellP) keep = false;  // fine pairs of one reference cell: once
                if (!(P.autocorr && refQ == refP)) code |= GEN_DIRZ;
            } else {
                const int dz = cand % wz - P.g.reach[2];
                const int dy = (cand / wz) % wy - P.g.reach[1];
                const int dx = cand / (wz * wy) - P.g.reach[0];
                const int t[3] = {gx + dx, gy + dy, gz + dz};
                int q[3];
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const int ng = P.g.ng[a];
                    {
                        // the neighbour must sit in a reference cell within +-refine of the primary's
                        // (gridlink_impl.c.src:499-518); the fine reach is padded to whole reference cells
                        const int sa = P.g.s[a];
                        const int rt = t[a] >= 0 ? t[a] / sa : -((-t[a] + sa - 1) / sa);
                        const int dref = rt - (a == 0 ? gx : (a == 1 ? gy : gz)) / sa;
                        if (dref > P.g.refine[a] || dref < -P.g.refine[a]) keep = false;
                    }
                    if (P.g.periodic[a]) {
                        // one image per side at most; further images are exact duplicates the
                        // reference suppresses (gridlink_utils.h.src:46-72), or fall outside its
                        // single-wrap index map (gridlink_impl.c.src:500-504)
                        if (t[a] < -ng || t[a] >= 2 * ng) keep = false;
                        if (t[a] < 0) {
                            q[a] = t[a] + ng;
                            code |= 1 << (2 * a);  // +wrap on the first particle
                            offd[a] = P.wrap[a];
                        } else if (t[a] >= ng) {
                            q[a] = t[a] - ng;
                            code |= 2 << (2 * a);  // -wrap
                            offd[a] = -P.wrap[a];
                        } else
                            q[a] = t[a];
                    } else {
                        if (t[a] < 0 || t[a] >= ng) keep = false;
                        q[a] = t[a];
                    }
                }
                if (keep) {
                    cellQ = (q[0] * P.g.ng[1] + q[1]) * P.g.ng[2] + q[2];
                    if (P.autocorr) {
                        // reference: skip icell2 > icell (gridlink_impl.c.src:525); inside one reference
                        // cell the pair (i<j) has i first, so only secondaries at or after the primary
                        const long long refB =
                            ((long long)(q[0] / P.g.s[0]) * P.g.n[1] + q[1] / P.g.s[1]) * P.g.n[2] + q[2] / P.g.s[2];
                        if (refB > refA || (refB == refA && cellQ < cellP)) keep = false;
                        if (refB != refA) code |= GEN_DIRZ;
                    } else
                        code |= GEN_DIRZ;
                }
            }
            if (keep && B.count[cellQ] == 0) keep = false;
            if (keep) {
                // bounds-based pruning (pure optimisation; margins keep it conservative); theta: chord vs max chord
                double qb[6];
#pragma unroll
                for (int k = 0; k < 6; k++) qb[k] = (double)B.bounds[(int64_t)cellQ * CFB_NB + k];
                double md[3];
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const double plo = pb[2 * a] + offd[a], phi = pb[2 * a + 1] + offd[a];
                    double d = 0.0;
                    if (qb[2 * a] > phi) d = qb[2 * a] - phi;
                    else if (plo > qb[2 * a + 1]) d = plo - qb[2 * a + 1];
                    md[a] = d;
                }
                const double slack = sizeof(T) == 4 ? 1.0e-5 : 1.0e-12;
                if (P.max_sep[0] > 0) {
                    const double s = md[0] * md[0] + md[1] * md[1] + md[2] * md[2];
                    if (s > P.max_sep[0] * P.max_sep[0] * (1.0 + slack)) keep = false;
                }
                if (P.max_sep[1] > 0) {
                    const double s = md[0] * md[0] + md[1] * md[1];
                    if (s > P.max_sep[1] * P.max_sep[1] * (1.0 + slack)) keep = false;
                }
                if (P.max_sep[2] > 0) {
                    if (md[2] > P.max_sep[2] * (1.0 + slack)) keep = false;
                }
            }
            if (keep) {
                const int slot = atomicAdd(&S.q_n, 1);
                S.q_cell[slot] = cellQ;
                S.q_code[slot] = code;
            }
        }
        __syncthreads();
        const int nq_cells = S.q_n;
        if (tid == 0) my_tp += nq_cells;
        // ---------------- phase 2: run the tile against each queued neighbour ----------------
        for (int e = 0; e < nq_cells; e++) {
            const int cellQ = S.q_cell[e];
            const int code = S.q_code[e];
            T xpos = xp, ypos = yp, zpos = zp;
            if (!LIST) {
                // first particle gets the wrap: xpos = x0 + off_xwrap (countpairs_kernels.c.src:79-81)
                const int cx = code & 3, cy = (code >> 2) & 3, cz = (code >> 4) & 3;
                if (cx) xpos = xp + (cx == 1 ? (T)P.wrap[0] : -(T)P.wrap[0]);
                if (cy) ypos = yp + (cy == 1 ? (T)P.wrap[1] : -(T)P.wrap[1]);
                if (cz) zpos = zp + (cz == 1 ? (T)P.wrap[2] : -(T)P.wrap[2]);
            }
            const bool tri = P.autocorr && (cellQ == cellP);  // same cell: only j > i
            const bool dirz = (code & GEN_DIRZ) != 0;
            const T tz = zpos - pimax;
            const int nQ = B.count[cellQ];
            const int startQ = B.start[cellQ];
            if (valid) my_eval += tri ? (unsigned long long)(nQ - 1 - iloc > 0 ? nQ - 1 - iloc : 0) : (unsigned long long)nQ;
            if (!warp_has_work) continue;  // warp-uniform: no primary in this warp
            for (int c0 = 0; c0 < nQ; c0 += GEN_CH) {
                const int m = min(GEN_CH, nQ - c0);
                __syncwarp();  // everyone is done with the previous chunk
                for (int k = lane; k < m; k += 32) {
                    sx[k] = B.x[startQ + c0 + k];
                    sy[k] = B.y[startQ + c0 + k];
                    sz[k] = B.z[startQ + c0 + k];
                    if (WGT) sw[k] = B.w[startQ + c0 + k];
                }
                __syncwarp();
                if (!valid) continue;
                int k0 = 0;
                if (tri) {
                    k0 = iloc + 1 - c0;
                    if (k0 < 0) k0 = 0;
                }
                // (Deferring the accepted pairs -- a lane parks the indices of up to 2 or 4 accepted secondaries and
                // the warp retires one parked pair per lane when some lane is full, so that the bin search / sqrt /
                // histogram code runs with most lanes active -- was measured 15-50 % SLOWER on configs 2 and 3.)
                for (int k = k0; k < m; k++) {
                    const T dx = sx[k] - xpos, dy = sy[k] - ypos, dz = sz[k] - zpos;
                    int64_t slot;
                    int kbin = 0;  // the separation bin (upper-edge index) of the pair
                    T sep = 0;
                    if (MODE == CFB_DD || MODE == CFB_XI) {
                        const T r2 = fma_t<T>(dz, dz, fma_t<T>(dy, dy, dx * dx));
                        if (!(r2 < e_hi && r2 >= e_lo)) continue;
                        int kb;
                        for (kb = nedges - 1; kb >= 1; kb--)
                            if (r2 >= s_edges[kb - 1]) break;
                        slot = kb;
                        kbin = kb;
                        if (AVG) sep = sqrt_t<T>(r2);
                    } else if (MODE == CFB_WP) {
                        const T r2 = fma_t<T>(dy, dy, dx * dx);
                        // two reference cells: the reference fast-forwards over z1 <= zpos - pimax and then masks with
                        // the SIGNED dz < pimax (wp_kernels.c.src:139-142, 207-221), so a survivor whose dz rounds to
                        // exactly -pimax counts; inside one reference cell j follows i in z order (dz >= 0)
                        if (dirz ? !(sz[k] > tz && dz < pimax) : !(dz > -pimax && dz < pimax)) continue;
                        if (!(r2 < e_hi && r2 >= e_lo)) continue;
                        int kb;
                        for (kb = nedges - 1; kb >= 1; kb--)
                            if (r2 >= s_edges[kb - 1]) break;
                        slot = kb;
                        kbin = kb;
                        if (AVG) sep = sqrt_t<T>(r2);
                    } else if (MODE == CFB_RPPI) {
                        const T r2 = fma_t<T>(dy, dy, dx * dx);
                        // two reference cells: survivors of the fast-forward over z1 <= zpos - pimax
                        // (countpairs_rp_pi_kernels.c.src:139-142), then |dz| < pimax (:196-207)
                        if (dirz ? !(sz[k] > tz) : !(dz > -pimax)) continue;
                        const T adz = dz < 0 ? -dz : dz;
                        if (!(adz < pimax)) continue;
                        if (!(r2 < e_hi && r2 >= e_lo)) continue;
                        int kb;
                        for (kb = nedges - 1; kb >= 1; kb--)
                            if (r2 >= s_edges[kb - 1]) break;
                        // rpbin*(npibin+1) + |dz|*inv_dpi evaluated in T, then truncated
                        // (countpairs_rp_pi_kernels.c.src:249-256)
                        const T fin = (T)kb * npi_p1 + adz * inv_dpi;
                        slot = (int64_t)(int)fin;
                        kbin = kb;
                        if (AVG) sep = sqrt_t<T>(r2);
                    } else if (MODE == CFB_SMU) {
                        const T sqr_dz = dz * dz;
                        const T s2 = fma_t<T>(dx, dx, fma_t<T>(dy, dy, sqr_dz));
                        if (!(sqr_dz < s2 * sqr_mumax)) continue;
                        if (!(s2 < e_hi && s2 >= e_lo)) continue;
                        const T mu = sqrt_t<T>(divi_t<T>(sqr_dz, s2));
                        int kb;
                        for (kb = nedges - 1; kb >= 1; kb--)
                            if (s2 >= s_edges[kb - 1]) break;
                        const T fin = (T)kb * nmu_p1 + mu * inv_dmu;
                        slot = (int64_t)(int)fin;
                        kbin = kb;
                        if (AVG) sep = sqrt_t<T>(s2);
                    } else if (MODE == CFB_RPPI_MOCKS) {
                        // line of sight = pair midpoint: pi^2 = (s.l)^2 / l^2, rp^2 = s^2 - pi^2
                        // (countpairs_rp_pi_mocks_kernels.c.src:200-262)
                        const T parx = sx[k] + xpos, pary = sy[k] + ypos, parz = sz[k] + zpos;
                        const T term1 = parx * dx, term2 = pary * dy;
                        const T s_dot_l = fma_t<T>(parz, dz, term1 + term2);
                        const T sqr_s_dot_l = s_dot_l * s_dot_l;
                        const T sqr_sep = fma_t<T>(dx, dx, fma_t<T>(dy, dy, dz * dz));
                        if (!(sqr_sep < sqr_max_sep)) continue;
                        const T sqr_norm_l = fma_t<T>(parx, parx, fma_t<T>(pary, pary, parz * parz));
                        if (!(sqr_s_dot_l < sqr_pimax * sqr_norm_l)) continue;
                        const T sqr_Dpar = divi_t<T>(sqr_s_dot_l, sqr_norm_l);  // fast_divide_and_NR_steps == 0
                        const T sqr_Dperp = sqr_sep - sqr_Dpar;
                        if (!(sqr_Dpar < sqr_pimax && sqr_Dperp < e_hi && sqr_Dperp >= e_lo)) continue;
                        const T Dpar = sqrt_t<T>(sqr_Dpar);
                        int kb;
                        for (kb = nedges - 1; kb >= 1; kb--)
                            if (sqr_Dperp >= s_edges[kb - 1]) break;
                        const T fin = (T)kb * npi_p1 + Dpar * inv_dpi;  // :283-290, evaluated in T, then truncated
                        slot = (int64_t)(int)fin;
                        kbin = kb;
                        if (AVG) sep = sqrt_t<T>(sqr_Dperp);
                    } else if (MODE == CFB_SMU_MOCKS) {
                        // mu^2 = (s.l)^2 / (l^2 s^2) (countpairs_s_mu_mocks_kernels.c.src:196-290)
                        const T parx = sx[k] + xpos, pary = sy[k] + ypos, parz = sz[k] + zpos;
                        const T term1 = parx * dx, term2 = pary * dy;
                        const T s_dot_l = fma_t<T>(parz, dz, term1 + term2);
                        const T sqr_s_dot_l = s_dot_l * s_dot_l;
                        const T s2 = fma_t<T>(dx, dx, fma_t<T>(dy, dy, dz * dz));
                        if (!(s2 < e_hi && s2 >= e_lo)) continue;
                        const T sqr_norm_l = fma_t<T>(parx, parx, fma_t<T>(pary, pary, parz * parz));
                        const T sqr_mu = divi_t<T>(sqr_s_dot_l, sqr_norm_l * s2);
                        if (!(sqr_mu < sqr_mumax)) continue;
                        const T mu = sqrt_t<T>(sqr_mu);
                        int kb;
                        for (kb = nedges - 1; kb >= 1; kb--)
                            if (s2 >= s_edges[kb - 1]) break;
                        const T fin = (T)kb * nmu_p1 + mu * inv_dmu;
                        slot = (int64_t)(int)fin;
                        kbin = kb;
                        if (AVG) sep = sqrt_t<T>(s2);
                    } else {  // CFB_THETA: cos(theta) = 1 - chord^2/2, edges are cos(theta_upp) (decreasing)
                        const T chord2 = fma_t<T>(dz, dz, fma_t<T>(dy, dy, dx * dx));
                        const T ct = (T)1.0 - (T)0.5 * chord2;
                        if (!(ct > e_hi && ct <= e_lo)) continue;
                        int kb;
                        for (kb = nedges - 1; kb >= 1; kb--)
                            if (ct <= s_edges[kb - 1]) break;
                        slot = kb;
                        kbin = kb;
                        if (AVG) {
                            const T cc = ct >= (T)1.0 ? (T)1.0 : ct;
                            const T th = P.fast_acos ? fast_acos_t<T>(cc) : (T)acos(cc);
                            sep = (T)(th * (T)57.29577951308232087679815481410517);
                        }
                    }
                    if (P.hist_in_smem) {
                        add_count(&s_np[slot], &s_np[ns + slot]);
                        if (AVG) add_fixed96(&s_sep[slot], &s_sep[ns + slot], &s_sep[2 * ns + slot], __double2ll_rn((double)sep * s_scale[kbin]));
                        if (WGT)
                            add_fixed96(&s_w[slot], &s_w[ns + slot], &s_w[2 * ns + slot],
                                        __double2ll_rn((double)(T)(wp * sw[k]) * w_scale));
                    } else {
                        atomicAdd(&P.npairs[slot], 1ULL);
                        if (AVG) atomicAdd(&P.sum_sep[slot], (double)sep);
                        if (WGT) atomicAdd(&P.sum_w[slot], (double)(T)(wp * sw[k]));
                    }
                }
            }
        }
        __syncthreads();
    }
    // ---------------- merge ----------------
    __syncthreads();
    if (P.hist_in_smem) {
        for (int64_t i = tid; i < ns; i += CFB_TILE) {
            const unsigned long long v = (unsigned long long)s_np[i] | ((unsigned long long)s_np[ns + i] << 32);
            if (v) {
                atomicAdd(&P.npairs[i], v);
                if (AVG) {
                    const int kb = (MODE == CFB_RPPI || MODE == CFB_RPPI_MOCKS)
                                       ? (int)(i / (P.npibin + 1))
                                       : ((MODE == CFB_SMU || MODE == CFB_SMU_MOCKS) ? (int)(i / (P.nmu_bins + 1)) : (int)i);
                    atomicAdd(&P.sum_sep[i], fixed96_to_double(s_sep[i], s_sep[ns + i], s_sep[2 * ns + i]) / s_scale[kb < nedges ? kb : nedges - 1]);
                }
                if (WGT) atomicAdd(&P.sum_w[i], fixed96_to_double(s_w[i], s_w[ns + i], s_w[2 * ns + i]) / w_scale);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) my_eval += __shfl_xor_sync(0xffffffffu, my_eval, o);
    if ((tid & 31) == 0 && my_eval) atomicAdd(&P.counters[0], my_eval);
    if (tid == 0 && my_tp) atomicAdd(&P.counters[1], my_tp);
}

// max |w| over the cell-sorted weights of one set (padding holds zeros), as the bit pattern of a non-negative
// double (ordered like the value) so that one 64-bit atomicMax merges the blocks
template <typename T>
__global__ void k_absmax(const int64_t n, const T *__restrict__ w, unsigned long long *out)
{
    double m = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double a = fabs((double)w[i]);
        if (a > m && a < 1.0e300) m = a;
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

template <typename T>
static SetView<T> view_of(const ParticleSet &S)
{
    SetView<T> v;
    v.x = (const T *)S.sorted[0].p;
    v.y = (const T *)S.sorted[1].p;
    v.z = (const T *)S.sorted[2].p;
    v.w = (const T *)S.sorted[3].p;
    v.count = (const int *)S.count.p;
    v.start = (const int *)S.start.p;
    v.bounds = (const T *)S.bounds.p;
    return v;
}

template <typename T, int MODE, bool AVG, bool WGT, bool LIST>
static int launch_inst(PairParams P, const ParticleSet &SA, const ParticleSet &SB, cudaStream_t st)
{
    auto kern = k_pairs_generic<T, MODE, AVG, WGT, LIST>;
    size_t sm = (((size_t)P.nedges * sizeof(T) + 15) & ~(size_t)15) + (size_t)P.nedges * 8;
    const size_t hist = (size_t)P.nslots * (8 + (AVG ? 12 : 0) + (WGT ? 12 : 0));
    const size_t budget = 160 * 1024;
    P.hist_in_smem = (sm + hist <= budget) ? 1 : 0;
    sm += P.hist_in_smem ? hist : 8;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget + 16384));
    const int64_t nblk = P.ntiles;
    if (nblk <= 0) return 0;
    if (nblk >= 2147483647LL) return cfb_fail("too many tiles (%lld)", (long long)nblk);
    kern<<<(unsigned int)nblk, CFB_TILE, sm, st>>>(P, view_of<T>(SA), view_of<T>(SB));
    cfb_ctx().launches++;
    CK(cudaGetLastError());
    return 0;
}

template <typename T, int MODE, bool LIST>
static int launch_mode(const cfb_binning *bin, const PairParams &P, const ParticleSet &SA, const ParticleSet &SB,
                       cudaStream_t st)
{
    const bool a = bin->need_avg != 0, w = bin->need_weights != 0;
    if (a && w) return launch_inst<T, MODE, true, true, LIST>(P, SA, SB, st);
    if (a) return launch_inst<T, MODE, true, false, LIST>(P, SA, SB, st);
    if (w) return launch_inst<T, MODE, false, true, LIST>(P, SA, SB, st);
    return launch_inst<T, MODE, false, false, LIST>(P, SA, SB, st);
}

template <typename T>
static int launch_T(const cfb_binning *bin, const PairParams &P, bool list_mode)
{
    Ctx &c = cfb_ctx();
    const ParticleSet &SA = c.set[0];
    const ParticleSet &SB = bin->autocorr ? c.set[0] : c.set[1];
    if (list_mode) {
        if (bin->mode != CFB_THETA) return cfb_fail("neighbour-list mode is only used by DDtheta");
        return launch_mode<T, CFB_THETA, true>(bin, P, SA, SB, c.stream);
    }
    switch (bin->mode) {
    case CFB_DD:
    case CFB_XI: return launch_mode<T, CFB_DD, false>(bin, P, SA, SB, c.stream);
    case CFB_WP: return launch_mode<T, CFB_WP, false>(bin, P, SA, SB, c.stream);
    case CFB_RPPI: return launch_mode<T, CFB_RPPI, false>(bin, P, SA, SB, c.stream);
    case CFB_SMU: return launch_mode<T, CFB_SMU, false>(bin, P, SA, SB, c.stream);
    case CFB_RPPI_MOCKS: return launch_mode<T, CFB_RPPI_MOCKS, false>(bin, P, SA, SB, c.stream);
    case CFB_SMU_MOCKS: return launch_mode<T, CFB_SMU_MOCKS, false>(bin, P, SA, SB, c.stream);
    default: return cfb_fail("unknown mode %d", bin->mode);
    }
}

template <typename T>
static int weight_maxima(Ctx &c, const cfb_binning *bin, PairParams &P)
{
    // two doubles at scratch + 2048 (the first 256 bytes belong to gridlink)
    if (cfb_ensure(c.scratch, 4096)) return 1;
    unsigned long long *out = (unsigned long long *)((char *)c.scratch.p + 2048);
    CK(cudaMemsetAsync(out, 0, 16, c.stream));
    const ParticleSet &SA = c.set[0];
    const ParticleSet &SB = bin->autocorr ? c.set[0] : c.set[1];
    const ParticleSet *sets[2] = {&SA, &SB};
    for (int k = 0; k < 2; k++) {
        const ParticleSet &S = *sets[k];
        if (S.npad <= 0 || !S.sorted[3].p) continue;
        int nb = (int)((S.npad + 1023) / 1024);
        if (nb > 1184) nb = 1184;
        k_absmax<T><<<nb, 256, 0, c.stream>>>(S.npad, (const T *)S.sorted[3].p, out + k);
        c.launches++;
    }
    CK(cudaGetLastError());
    P.wmax = (const double *)out;
    return 0;
}

int cfb_launch_pairs_generic(const cfb_binning *bin, const PairParams &P0, int prec, bool list_mode)
{
    PairParams P = P0;
    if (bin->need_weights) {
        if (prec == 4 ? weight_maxima<float>(cfb_ctx(), bin, P) : weight_maxima<double>(cfb_ctx(), bin, P)) return 1;
    }
    // the per-pair-sum kernel (pairs_sum.cu) serves everything it can hold in shared memory; this kernel remains for
    // histograms too large for that, more than 256 edges, and as a cross-check (force_kernel 3 / CORRFUNC_B200_LEGACY_GENERIC)
    static int legacy = -1;
    if (legacy < 0) {
        const char *e = getenv("CORRFUNC_B200_LEGACY_GENERIC");
        legacy = (e && atoi(e) > 0) ? 1 : 0;
    }
    if (!legacy && cfb_ctx().force_kernel != 3 && !(cfb_ctx().prefer_legacy && cfb_ctx().force_kernel != 0)) {
        const int rc = cfb_launch_pairs_sum(bin, P, prec, list_mode);
        if (rc >= 0) {
            cfb_ctx().last_kind = 2;
            return rc;
        }
    }
    cfb_ctx().last_kind = 0;
    return prec == 4 ? launch_T<float>(bin, P, list_mode) : launch_T<double>(bin, P, list_mode);
}
