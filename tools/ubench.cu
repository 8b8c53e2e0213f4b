// ubench.cu -- B200 instruction-throughput probes that decide the pair-kernel design.
// Each probe runs the same loop body in every thread of a full-chip grid (148*k CTAs x 256 threads)
// and reports warp-instructions per clock per SM (from cudaEvent time and the SM clock).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define ITERS 4096
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// 0: 8 independent scalar FFMA per iteration
__global__ void k_ffma(float *out, float a, float b) {
    float acc[8]; for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = __fmaf_rn(acc[i], a, b);
    }
    float s = 0; for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 1: 8 independent packed FFMA2 per iteration (16 lane-FMAs)
__global__ void k_ffma2(float *out, float a, float b) {
    unsigned long long acc[8], pa = pack2(a, a), pb = pack2(b, b);
    for (int i = 0; i < 8; i++) acc[i] = pack2(threadIdx.x + i, i);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = fma2(acc[i], pa, pb);
    }
    unsigned long long s = 0; for (int i = 0; i < 8; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(s ^ (s >> 32)));
}
// 2: pair-distance body, scalar: 4 primaries x 1 secondary: 3 sub + mul + 2 fma each (24 FP32 instr)
__global__ void k_dist_scalar(float *out, const float *sec, float e) {
    float xp[4], yp[4], zp[4]; int cnt[4] = {0,0,0,0};
    for (int r = 0; r < 4; r++) { xp[r] = threadIdx.x * 0.01f + r; yp[r] = r * 0.5f; zp[r] = blockIdx.x * 0.001f; }
    for (int it = 0; it < ITERS; it++) {
        const float xs = sec[(it * 3) & 1023], ys = sec[(it * 3 + 1) & 1023], zs = sec[(it * 3 + 2) & 1023];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const float dx = xs - xp[r], dy = ys - yp[r], dz = zs - zp[r];
            const float r2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
            cnt[r] += (r2 >= e);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = cnt[0] + cnt[1] + cnt[2] + cnt[3];
}
// 3: same with NLEV compare+count levels per pair
template <int NLEV>
__global__ void k_dist_levels(float *out, const float *sec, float e0, float de) {
    float xp[4], yp[4], zp[4]; int cnt[NLEV];
    for (int l = 0; l < NLEV; l++) cnt[l] = 0;
    for (int r = 0; r < 4; r++) { xp[r] = threadIdx.x * 0.01f + r; yp[r] = r * 0.5f; zp[r] = blockIdx.x * 0.001f; }
    for (int it = 0; it < ITERS; it++) {
        const float xs = sec[(it * 3) & 1023], ys = sec[(it * 3 + 1) & 1023], zs = sec[(it * 3 + 2) & 1023];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const float dx = xs - xp[r], dy = ys - yp[r], dz = zs - zp[r];
            const float r2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
#pragma unroll
            for (int l = 0; l < NLEV; l++) cnt[l] += (r2 >= e0 + l * de);
        }
    }
    int s = 0; for (int l = 0; l < NLEV; l++) s += cnt[l] * (l + 1);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 4: packed distance: 4 primaries x 2 secondaries per iteration via f32x2 (12 packed instr = 24 lane ops x2)
template <int NLEV>
__global__ void k_dist_packed(float *out, const float *sec, float e0, float de) {
    unsigned long long xp[4], yp[4], zp[4]; int cnt[NLEV];
    for (int l = 0; l < NLEV; l++) cnt[l] = 0;
    for (int r = 0; r < 4; r++) { float a = threadIdx.x * 0.01f + r, b = r * 0.5f, c = blockIdx.x * 0.001f;
        xp[r] = pack2(a, a); yp[r] = pack2(b, b); zp[r] = pack2(c, c); }
    const unsigned long long *s2 = (const unsigned long long *)sec;
    for (int it = 0; it < ITERS; it++) {
        const unsigned long long xs = s2[(it * 3) & 511], ys = s2[(it * 3 + 1) & 511], zs = s2[(it * 3 + 2) & 511];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const unsigned long long dx = sub2(xs, xp[r]), dy = sub2(ys, yp[r]), dz = sub2(zs, zp[r]);
            const unsigned long long r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
            const float r2a = __uint_as_float((unsigned)r2), r2b = __uint_as_float((unsigned)(r2 >> 32));
#pragma unroll
            for (int l = 0; l < NLEV; l++) { cnt[l] += (r2a >= e0 + l * de); cnt[l] += (r2b >= e0 + l * de); }
        }
    }
    int s = 0; for (int l = 0; l < NLEV; l++) s += cnt[l] * (l + 1);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 5: shared-memory private histogram RMW (conflict free: hist[bin][tid]) vs atomicAdd
template <bool ATOMIC>
__global__ void k_smem_hist(float *out, int nb) {
    extern __shared__ unsigned int h[];
    for (int i = threadIdx.x; i < nb * 256; i += 256) h[i] = 0;
    __syncthreads();
    unsigned int x = threadIdx.x * 2654435761u + blockIdx.x;
    for (int it = 0; it < ITERS; it++) {
        x = x * 1664525u + 1013904223u;
        const int b = (x >> 20) % nb;
        if (ATOMIC) atomicAdd(&h[b * 256 + threadIdx.x], 1u); else h[b * 256 + threadIdx.x] += 1u;
    }
    __syncthreads();
    unsigned int s = 0; for (int b = 0; b < nb; b++) s += h[b * 256 + threadIdx.x];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 6: DFMA chain (8 independent) and double distance body with NLEV levels
__global__ void k_dfma(double *out, double a, double b) {
    double acc[8]; for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = __fma_rn(acc[i], a, b);
    }
    double s = 0; for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NLEV>
__global__ void k_ddist_levels(double *out, const double *sec, double e0, double de) {
    double xp[4], yp[4], zp[4]; int cnt[NLEV];
    for (int l = 0; l < NLEV; l++) cnt[l] = 0;
    for (int r = 0; r < 4; r++) { xp[r] = threadIdx.x * 0.01 + r; yp[r] = r * 0.5; zp[r] = blockIdx.x * 0.001; }
    for (int it = 0; it < ITERS; it++) {
        const double xs = sec[(it * 3) & 511], ys = sec[(it * 3 + 1) & 511], zs = sec[(it * 3 + 2) & 511];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const double dx = xs - xp[r], dy = ys - yp[r], dz = zs - zp[r];
            const double r2 = __fma_rn(dz, dz, __fma_rn(dy, dy, dx * dx));
#pragma unroll
            for (int l = 0; l < NLEV; l++) cnt[l] += (r2 >= e0 + l * de);
        }
    }
    int s = 0; for (int l = 0; l < NLEV; l++) s += cnt[l] * (l + 1);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double timeit(F launch) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch(); launch(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a)); for (int i = 0; i < 5; i++) launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms / 5.0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("device %s SMs=%d maxclock=%d kHz\n", p.name, p.multiProcessorCount, clk_khz);
    const int nsm = p.multiProcessorCount, grid = nsm * 8, bs = 256;
    float *out; CK(cudaMalloc(&out, sizeof(double) * grid * bs));
    float *sec; CK(cudaMalloc(&sec, 8192)); CK(cudaMemset(sec, 0, 8192));
    const double warps = (double)grid * bs / 32.0;
    auto report = [&](const char *name, double ms, double pair_evals_per_thread_iter, double instr_per_iter) {
        const double evals = (double)grid * bs * ITERS * pair_evals_per_thread_iter;
        const double winstr = warps * ITERS * instr_per_iter;
        printf("%-34s %8.3f ms  %8.2f Gpair-evals/s  (%.3e warp-instr/s nominal -> %.2f per clk per SM @max clock)\n", name, ms,
               evals / ms / 1e6, winstr / (ms * 1e-3), winstr / (ms * 1e-3) / (clk_khz * 1e3) / nsm);
    };
    double ms;
    ms = timeit([&] { k_ffma<<<grid, bs>>>(out, 1.0001f, 0.5f); }); report("ffma x8 (scalar)", ms, 0, 8);
    ms = timeit([&] { k_ffma2<<<grid, bs>>>(out, 1.0001f, 0.5f); }); report("ffma2 x8 (packed, 16 lane-fma)", ms, 0, 8);
    ms = timeit([&] { k_dist_scalar<<<grid, bs>>>(out, sec, 1e30f); }); report("dist scalar 4x1 + 1 level", ms, 4, 24 + 8 + 3);
    ms = timeit([&] { k_dist_levels<1><<<grid, bs>>>(out, sec, 1e30f, 1.f); }); report("dist scalar 4x1, 1 level", ms, 4, 24 + 8 + 3);
    ms = timeit([&] { k_dist_levels<2><<<grid, bs>>>(out, sec, 1e30f, 1.f); }); report("dist scalar 4x1, 2 levels", ms, 4, 24 + 16 + 3);
    ms = timeit([&] { k_dist_levels<3><<<grid, bs>>>(out, sec, 1e30f, 1.f); }); report("dist scalar 4x1, 3 levels", ms, 4, 24 + 24 + 3);
    ms = timeit([&] { k_dist_levels<4><<<grid, bs>>>(out, sec, 1e30f, 1.f); }); report("dist scalar 4x1, 4 levels", ms, 4, 24 + 32 + 3);
    ms = timeit([&] { k_dist_levels<6><<<grid, bs>>>(out, sec, 1e30f, 1.f); }); report("dist scalar 4x1, 6 levels", ms, 4, 24 + 48 + 3);
    ms = timeit([&] { k_dist_packed<1><<<grid, bs>>>(out, sec, 1e30f, 1.f); }); report("dist packed 4x2, 1 level", ms, 8, 24 + 16 + 3);
    ms = timeit([&] { k_dist_packed<2><<<grid, bs>>>(out, sec, 1e30f, 1.f); }); report("dist packed 4x2, 2 levels", ms, 8, 24 + 32 + 3);
    ms = timeit([&] { k_dist_packed<3><<<grid, bs>>>(out, sec, 1e30f, 1.f); }); report("dist packed 4x2, 3 levels", ms, 8, 24 + 48 + 3);
    ms = timeit([&] { k_dist_packed<4><<<grid, bs>>>(out, sec, 1e30f, 1.f); }); report("dist packed 4x2, 4 levels", ms, 8, 24 + 64 + 3);
    CK(cudaFuncSetAttribute(k_smem_hist<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 256 * 4));
    CK(cudaFuncSetAttribute(k_smem_hist<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 256 * 4));
    ms = timeit([&] { k_smem_hist<false><<<grid, bs, 32 * 256 * 4>>>(out, 32); }); report("smem private hist RMW (ld+add+st)", ms, 1, 1);
    ms = timeit([&] { k_smem_hist<true><<<grid, bs, 32 * 256 * 4>>>(out, 32); }); report("smem private hist atomicAdd", ms, 1, 1);
    ms = timeit([&] { k_dfma<<<grid, bs>>>((double *)out, 1.0001, 0.5); }); report("dfma x8", ms, 0, 8);
    ms = timeit([&] { k_ddist_levels<1><<<grid, bs>>>((double *)out, (double *)sec, 1e300, 1.); }); report("ddist 4x1, 1 level", ms, 4, 24 + 8 + 3);
    ms = timeit([&] { k_ddist_levels<2><<<grid, bs>>>((double *)out, (double *)sec, 1e300, 1.); }); report("ddist 4x1, 2 levels", ms, 4, 24 + 16 + 3);
    ms = timeit([&] { k_ddist_levels<4><<<grid, bs>>>((double *)out, (double *)sec, 1e300, 1.); }); report("ddist 4x1, 4 levels", ms, 4, 24 + 32 + 3);
    return 0;
}
