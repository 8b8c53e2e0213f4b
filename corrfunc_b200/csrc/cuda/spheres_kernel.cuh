// spheres_kernel.cuh -- the device code of spheres.cu (k_spheres and what it needs), kept in a header of its own so that
// tests/cuda_emul/ can compile the very same text for the CPU (one std::thread per CUDA thread) and check its indexing
// and arithmetic against the brute-force stand-in without a GPU.  Needs SetView<T> (cfb_internal.cuh) declared first.
#pragma once

template <typename T>
__device__ __forceinline__ T sph_fma(T a, T b, T c);
template <>
__device__ __forceinline__ float sph_fma<float>(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template <>
__device__ __forceinline__ double sph_fma<double>(double a, double b, double c) { return __fma_rn(a, b, c); }

struct SphGeom {
    int n[3];         // lattice cells per axis
    int periodic[3];  // axis wraps
    double lo[3], inv[3], wrap[3];  // REAL-valued
};

// SHELLS = true : r2 = fma(dz,dz, fma(dy,dy, dx*dx)) and the AVX-512 kernel's shell assignment (vpf_mocks_kernels:63-92,
//                 theory/vpf/vpf_kernels.c.src alike)
// SHELLS = false: r2 = dx*dx + dy*dy + dz*dz and a plain count of r2 < rmax_sqr (count_neighbors, impl:140-204)
// Periodic axes (theory vpf): the reference shifts the centre by -+wrap for the neighbour cells across the box edge
// (countspheres_impl.c.src:331-380); a particle that can count is always met with its nearest image, so the image is
// chosen per particle here (d > wrap/2 -> centre + wrap, d < -wrap/2 -> centre - wrap), whatever the lattice.
template <typename T, bool SHELLS>
__global__ void k_spheres(const int64_t ncen, const T *__restrict__ xc, const T *__restrict__ yc, const T *__restrict__ zc,
                          const SetView<T> B, const SphGeom G, const T rmax_sqr, const int nbin,
                          const T *__restrict__ edges, unsigned *__restrict__ out)
{
    extern __shared__ unsigned char sph_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    T *s_E = (T *)sph_smem;
    unsigned *s_cnt = (unsigned *)(sph_smem + (((size_t)nbin * sizeof(T) + 15) & ~(size_t)15)) + (size_t)wid * nbin;
    for (int k = threadIdx.x; k < nbin; k += blockDim.x) s_E[k] = SHELLS ? edges[k] : (T)0;
    for (int k = lane; k < nbin; k += 32) s_cnt[k] = 0u;
    __syncthreads();
    const int64_t c = (int64_t)blockIdx.x * nw + wid;
    if (c >= ncen) return;
    const T C[3] = {xc[c], yc[c], zc[c]};
    int first[3], cnt[3];  // per axis: the run of neighbour cells (taken modulo n on a periodic axis)
    T wrap[3], half[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        int i = (int)((C[a] - (T)G.lo[a]) * (T)G.inv[a]);
        i = min(max(i, 0), G.n[a] - 1);
        wrap[a] = (T)G.wrap[a];
        half[a] = (T)0.5 * wrap[a];
        if (G.periodic[a]) {
            if (G.n[a] >= 3) {
                first[a] = i - 1 + G.n[a];
                cnt[a] = 3;
            } else {
                first[a] = 0;
                cnt[a] = G.n[a];
            }
        } else {
            first[a] = max(i - 1, 0);
            cnt[a] = min(i + 1, G.n[a] - 1) - first[a] + 1;
        }
    }
    for (int tx = 0; tx < cnt[0]; tx++)
        for (int ty = 0; ty < cnt[1]; ty++)
            for (int tz = 0; tz < cnt[2]; tz++) {
                const int cx = (first[0] + tx) % G.n[0], cy = (first[1] + ty) % G.n[1], cz = (first[2] + tz) % G.n[2];
                const int cell = (cx * G.n[1] + cy) * G.n[2] + cz;
                const int n = B.count[cell], s0 = B.start[cell];
                for (int j = lane; j < n; j += 32) {
                    const T P[3] = {B.x[s0 + j], B.y[s0 + j], B.z[s0 + j]};
                    T d[3];
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        T cen = C[a];
                        if (G.periodic[a]) {
                            const T raw = P[a] - C[a];
                            if (raw > half[a]) cen = C[a] + wrap[a];
                            else if (raw < -half[a]) cen = C[a] - wrap[a];
                        }
                        d[a] = cen - P[a];
                    }
                    if (SHELLS) {
                        const T r2 = sph_fma<T>(d[2], d[2], sph_fma<T>(d[1], d[1], d[0] * d[0]));
                        if (!(r2 < rmax_sqr)) continue;
                        // lane by lane what the masked loop over k = nbin-1 .. 1 does: the shell with E[k-1] <= r2 < E[k];
                        // whoever is left after k == 1 goes to shell 0; with one shell the loop never runs
                        bool left = true;
                        for (int k = nbin - 1; k >= 1; k--)
                            if (r2 < s_E[k] && r2 >= s_E[k - 1]) {
                                atomicAdd(&s_cnt[k], 1u);
                                left = false;
                                break;
                            }
                        if (left && nbin >= 2) atomicAdd(&s_cnt[0], 1u);
                    } else {
                        const T r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];  // -fmad=false: no contraction
                        if (r2 < rmax_sqr) atomicAdd(&s_cnt[0], 1u);
                    }
                }
            }
    __syncwarp();
    for (int k = lane; k < nbin; k += 32) out[c * nbin + k] = s_cnt[k];
}

