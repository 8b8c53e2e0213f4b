// TEST INFRASTRUCTURE ONLY -- runs the text of corrfunc_b200/csrc/cuda/spheres_kernel.cuh on the CPU: one std::thread per
// CUDA thread, blocks one after the other, __syncthreads / __syncwarp as real barriers, shared memory as one array.
// It checks the kernel's indexing (neighbour-cell runs, periodic images, shared-memory layout, output layout) and its
// arithmetic against a brute force over all particles; it cannot check anything the CUDA runtime does (launches, copies).
//   g++ -std=c++20 -O1 -pthread -ffp-contract=off emul_spheres.cpp -o emul_spheres && ./emul_spheres
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <random>
#include <thread>
#include <vector>

// ---- the CUDA constructs the kernel uses ----
#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __shared__
struct Dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local Dim3 threadIdx, blockIdx;
static Dim3 blockDim;
unsigned char sph_smem[96 * 1024];
static std::barrier<> *g_block_barrier = nullptr;
static std::vector<std::unique_ptr<std::barrier<>>> g_warp_barrier;
static inline void __syncthreads() { g_block_barrier->arrive_and_wait(); }
static inline void __syncwarp() { g_warp_barrier[threadIdx.x >> 5]->arrive_and_wait(); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
using std::max;
using std::min;
template <typename T>
struct SetView {  // cfb_internal.cuh
    const T *x, *y, *z, *w;
    const int *count, *start;
    const T *bounds;
};

#include "spheres_kernel.cuh"

template <typename T, bool SHELLS>
static void launch(int nblk, int nthreads, int64_t ncen, const T *xc, const T *yc, const T *zc, SetView<T> B, SphGeom G,
                   T rmax_sqr, int nbin, const T *edges, unsigned *out)
{
    blockDim.x = (unsigned)nthreads;
    for (int b = 0; b < nblk; b++) {
        std::barrier<> bar(nthreads);
        g_block_barrier = &bar;
        g_warp_barrier.clear();
        for (int w = 0; w < nthreads / 32; w++) g_warp_barrier.push_back(std::make_unique<std::barrier<>>(32));
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++)
            th.emplace_back([=]() {
                threadIdx.x = (unsigned)t;
                blockIdx.x = (unsigned)b;
                k_spheres<T, SHELLS>(ncen, xc, yc, zc, B, G, rmax_sqr, nbin, edges, out);
            });
        for (auto &t : th) t.join();
    }
}

template <typename T>
static int run_case(bool periodic, bool shells, int nm_override, unsigned seed)
{
    std::mt19937 rng(seed);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    const int N = 6000, ncen = 37, nbin = shells ? 5 : 1;
    const double lo[3] = {3.0, -7.0, 100.0}, ext[3] = {90.0, 60.0, 75.0};
    const T rmax = (T)11.0, rmax_sqr = rmax * rmax;
    std::vector<T> P[3], C[3];
    for (int a = 0; a < 3; a++) {
        for (int i = 0; i < N; i++) P[a].push_back((T)(lo[a] + ext[a] * U(rng)));
        P[a][0] = (T)lo[a];                       // extent corners are part of the data, as for the host layer
        P[a][1] = (T)(lo[a] + ext[a]);
        for (int i = 0; i < ncen; i++) C[a].push_back((T)(lo[a] + ext[a] * U(rng) * (periodic ? 1.02 : 1.0)));  // slightly past the extent
    }
    // lattice + cell-sorted SoA the way gridlink leaves it: runs padded to 4, start = padded offset
    SphGeom G;
    for (int a = 0; a < 3; a++) {
        int nm = nm_override > 0 ? nm_override : (int)std::floor(ext[a] / ((double)rmax * 1.001));
        G.n[a] = std::max(1, nm);
        G.periodic[a] = periodic;
        G.lo[a] = (double)(T)lo[a];
        G.inv[a] = (double)((T)G.n[a] / (T)ext[a]);
        G.wrap[a] = (double)(T)(ext[a] * (periodic ? 1.02 : 1.0));
    }
    const int ncell = G.n[0] * G.n[1] * G.n[2];
    std::vector<int> cell(N), count(ncell, 0), start(ncell, 0);
    for (int i = 0; i < N; i++) {
        int g[3];
        for (int a = 0; a < 3; a++) {
            int v = (int)((P[a][i] - (T)G.lo[a]) * (T)G.inv[a]);
            if (v > G.n[a] - 1) v--;
            g[a] = std::min(std::max(v, 0), G.n[a] - 1);
        }
        cell[i] = (g[0] * G.n[1] + g[1]) * G.n[2] + g[2];
        count[cell[i]]++;
    }
    int off = 0;
    for (int c = 0; c < ncell; c++) {
        start[c] = off;
        off += (count[c] + 3) & ~3;
    }
    std::vector<T> S[3];
    for (int a = 0; a < 3; a++) S[a].assign(off + 4, std::nanf(""));
    std::vector<int> fill(ncell, 0);
    for (int i = 0; i < N; i++) {
        const int p = start[cell[i]] + fill[cell[i]]++;
        for (int a = 0; a < 3; a++) S[a][p] = P[a][i];
    }
    SetView<T> B{S[0].data(), S[1].data(), S[2].data(), nullptr, count.data(), start.data(), nullptr};
    std::vector<T> E(nbin);
    const T rstep = rmax / (T)nbin;
    for (int k = 0; k < nbin; k++) E[k] = (k + 1) * rstep * rstep * (k + 1);
    std::vector<unsigned> out((size_t)ncen * nbin, 12345u);
    const int warps = 8;
    if (shells) launch<T, true>((ncen + warps - 1) / warps, warps * 32, ncen, C[0].data(), C[1].data(), C[2].data(), B, G, rmax_sqr, nbin, E.data(), out.data());
    else launch<T, false>((ncen + warps - 1) / warps, warps * 32, ncen, C[0].data(), C[1].data(), C[2].data(), B, G, rmax_sqr, nbin, E.data(), out.data());
    // brute force over all particles (what tests/stub_device does)
    int bad = 0;
    for (int c = 0; c < ncen; c++) {
        std::vector<unsigned> want(nbin, 0u);
        for (int j = 0; j < N; j++) {
            T d[3];
            for (int a = 0; a < 3; a++) {
                T cen = C[a][c];
                if (periodic) {
                    const T raw = P[a][j] - C[a][c], half = (T)0.5 * (T)G.wrap[a];
                    if (raw > half) cen = C[a][c] + (T)G.wrap[a];
                    else if (raw < -half) cen = C[a][c] - (T)G.wrap[a];
                }
                d[a] = cen - P[a][j];
            }
            if (shells) {
                const T r2 = std::fma(d[2], d[2], std::fma(d[1], d[1], d[0] * d[0]));
                if (!(r2 < rmax_sqr)) continue;
                bool left = true;
                for (int k = nbin - 1; k >= 1; k--)
                    if (r2 < E[k] && r2 >= E[k - 1]) { want[k]++; left = false; break; }
                if (left && nbin >= 2) want[0]++;
            } else {
                const T r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                if (r2 < rmax_sqr) want[0]++;
            }
        }
        for (int k = 0; k < nbin; k++)
            if (out[(size_t)c * nbin + k] != want[k]) bad++;
    }
    std::printf("%s periodic=%d shells=%d lattice=%dx%dx%d : %s\n", sizeof(T) == 4 ? "float " : "double", (int)periodic, (int)shells,
                G.n[0], G.n[1], G.n[2], bad ? "MISMATCH" : "ok");
    return bad;
}

int main()
{
    int bad = 0;
    for (int per = 0; per < 2; per++)
        for (int sh = 0; sh < 2; sh++) {
            bad += run_case<float>(per, sh, 0, 1u + per + 2 * sh);
            bad += run_case<double>(per, sh, 0, 11u + per + 2 * sh);
        }
    bad += run_case<double>(true, true, 2, 21u);   // two cells per axis: every cell once, images per particle
    bad += run_case<float>(true, true, 1, 22u);    // one cell
    bad += run_case<double>(false, true, 3, 23u);
    std::printf(bad ? "FAILED\n" : "all cases agree\n");
    return bad ? 1 : 0;
}
