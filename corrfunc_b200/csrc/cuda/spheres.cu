// spheres.cu -- counts-in-spheres for the void probability function on survey catalogues
// (mocks/vpf_mocks/countspheres_mocks_impl.c.src:206-638, vpf_mocks_kernels.c.src:20-100).
//
// One warp per sphere centre: the particles of the 27 lattice cells around the centre are tried against it, lanes
// striding over a cell's contiguous run in the cell-sorted SoA (coalesced).  The per-centre shell counts live in the
// warp's slice of shared memory (native 32-bit ATOMS.ADD) and are written out once.  HBM/L2 bound in principle
// (each centre re-reads ~27 cells that its neighbours also read, so they come from L2); the work is tiny next to the
// pair counts: 1e5 centres x a few hundred candidates each.
#include <math_constants.h>

#include "cfb_internal.cuh"

static DevBuf g_cen[3], g_out, g_edges;

void cfb_spheres_release()  // called by cfb_shutdown
{
    DevBuf *all[5] = {&g_cen[0], &g_cen[1], &g_cen[2], &g_out, &g_edges};
    for (DevBuf *b : all) {
        if (b->p) cudaFree(b->p);
        b->p = nullptr;
        b->cap = 0;
    }
}

#include "spheres_kernel.cuh"

template <typename T>
static int count_spheres_T(Ctx &c, ParticleSet &S, const double lo[3], const double ext[3], const int periodic[3],
                           const double wrap[3], int regrid, int64_t ncen, const void *xc, const void *yc, const void *zc,
                           double rmax, double rmax_sqr, int nbin, const double *edges, int shells, uint32_t *counts)
{
    // internal lattice over the particles' extent: cells a little larger than rmax, so that the 27 cells around a centre
    // hold every particle within rmax whatever way the two cell indices round; the counts do not depend on this choice
    SphGeom G;
    for (int k = 0; k < 3; k++) {
        int nm = ext[k] > 0 ? (int)floor(ext[k] / (rmax * 1.001)) : 1;
        nm = nm < 1 ? 1 : (nm > 128 ? 128 : nm);
        G.n[k] = nm;
        G.periodic[k] = periodic[k];
        G.lo[k] = (double)(T)lo[k];
        G.inv[k] = ext[k] > 0 ? (double)((T)nm / (T)ext[k]) : 0.0;
        G.wrap[k] = (double)(T)wrap[k];
    }
    if (regrid || !S.gridded) {
        cfb_box_lattice lat;
        memset(&lat, 0, sizeof(lat));
        for (int k = 0; k < 3; k++) {
            lat.nmesh[k] = G.n[k];
            lat.refine[k] = 1;
            lat.lo[k] = G.lo[k];
            lat.inv[k] = G.inv[k];
            lat.max_sep[k] = -1.0;
        }
        const int sub[3] = {1, 1, 1};
        if (cfb_gridlink_box_set(S, &lat, sub, 1.0)) return 1;
    }
    const size_t cb = (size_t)(ncen > 0 ? ncen : 1) * sizeof(T);
    const void *src[3] = {xc, yc, zc};
    for (int a = 0; a < 3; a++) {
        if (cfb_ensure(g_cen[a], cb)) return 1;
        CK(cudaMemcpyAsync(g_cen[a].p, src[a], (size_t)ncen * sizeof(T), cudaMemcpyHostToDevice, c.stream));
    }
    if (cfb_ensure(g_out, (size_t)(ncen > 0 ? ncen : 1) * nbin * 4)) return 1;
    if (cfb_ensure(g_edges, (size_t)nbin * sizeof(T))) return 1;
    if (shells) {
        T *h = (T *)malloc((size_t)nbin * sizeof(T));
        if (!h) return cfb_fail("out of host memory");
        for (int k = 0; k < nbin; k++) h[k] = (T)edges[k];
        cudaError_t e = cudaMemcpyAsync(g_edges.p, h, (size_t)nbin * sizeof(T), cudaMemcpyHostToDevice, c.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
        free(h);
        if (e != cudaSuccess) return cfb_fail("CUDA error %s uploading the shell edges", cudaGetErrorName(e));
    }
    if (ncen > 0) {
        SetView<T> V;
        V.x = (const T *)S.sorted[0].p;
        V.y = (const T *)S.sorted[1].p;
        V.z = (const T *)S.sorted[2].p;
        V.w = nullptr;
        V.count = (const int *)S.count.p;
        V.start = (const int *)S.start.p;
        V.bounds = (const T *)S.bounds.p;
        const int warps = 8;
        const size_t smem = (((size_t)nbin * sizeof(T) + 15) & ~(size_t)15) + (size_t)warps * nbin * 4;
        if (smem > 96 * 1024) return cfb_fail("too many radial bins (%d) for the counts-in-spheres kernel", nbin);
        const int64_t nblk = (ncen + warps - 1) / warps;
        if (nblk >= 2147483647LL) return cfb_fail("too many sphere centres (%lld)", (long long)ncen);
        if (shells) {
            CK(cudaFuncSetAttribute(k_spheres<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            k_spheres<T, true><<<(unsigned)nblk, warps * 32, smem, c.stream>>>(
                ncen, (const T *)g_cen[0].p, (const T *)g_cen[1].p, (const T *)g_cen[2].p, V, G, (T)rmax_sqr, nbin,
                (const T *)g_edges.p, (unsigned *)g_out.p);
        } else {
            CK(cudaFuncSetAttribute(k_spheres<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            k_spheres<T, false><<<(unsigned)nblk, warps * 32, smem, c.stream>>>(
                ncen, (const T *)g_cen[0].p, (const T *)g_cen[1].p, (const T *)g_cen[2].p, V, G, (T)rmax_sqr, nbin,
                (const T *)g_edges.p, (unsigned *)g_out.p);
        }
        c.launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(counts, g_out.p, (size_t)ncen * nbin * 4, cudaMemcpyDeviceToHost, c.stream));
    }
    CK(cudaStreamSynchronize(c.stream));
    return 0;
}

extern "C" int cfb_count_spheres(int slot, int prec, const double lo[3], const double ext[3], const int periodic[3],
                                 const double wrap[3], int regrid, int64_t ncen, const void *xc, const void *yc,
                                 const void *zc, double rmax, double rmax_sqr, int nbin, const double *edges, int shells,
                                 uint32_t *counts)
{
    Ctx &c = cfb_ctx();
    if (!c.ready) return cfb_fail("cfb_count_spheres before cfb_upload");
    if (slot < 0 || slot > 1) return cfb_fail("bad particle slot %d", slot);
    CK(cudaSetDevice(c.dev));
    ParticleSet &S = c.set[slot];
    if (S.prec != prec) return cfb_fail("particle set %d precision mismatch", slot);
    if (!(rmax > 0.0) || nbin < 1) return cfb_fail("bad counts-in-spheres parameters");
    return prec == 4 ? count_spheres_T<float>(c, S, lo, ext, periodic, wrap, regrid, ncen, xc, yc, zc, rmax, rmax_sqr, nbin,
                                              edges, shells, counts)
                     : count_spheres_T<double>(c, S, lo, ext, periodic, wrap, regrid, ncen, xc, yc, zc, rmax, rmax_sqr, nbin,
                                               edges, shells, counts);
}
