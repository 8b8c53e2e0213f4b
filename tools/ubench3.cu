// ubench3.cu -- FMA-pipe probes for packed f32x2 instruction mixes on B200 (see profiles/r1_ubench2_b200.txt)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
typedef unsigned long long u64;
#define IT 2048
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fsub(float a, float b) { float d; asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float fmul(float a, float b) { float d; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

// MODE 0: alternating independent FADD2 / FFMA2 chains (8)      -> 8 packed / iter
// MODE 1: 6 groups of the distance sequence (3 sub, mul, 2 fma), loop carried -> 36 packed / iter
// MODE 2: 4 FFMA2 chains + 4 scalar FFMA chains                -> 4 packed + 4 scalar
// MODE 3: 4 FFMA2 chains + 8 scalar FFMA chains                -> 4 packed + 8 scalar
// MODE 4: distance sequence, scalar, 12 groups                  -> 72 scalar / iter
// MODE 5: 6 groups packed distance + 6 groups scalar distance  -> 36 packed + 36 scalar
// MODE 6: 6 groups packed distance + 3 groups scalar distance  -> 36 packed + 18 scalar
// MODE 7: 8 FADD2 chains where the subtrahend is a broadcast scalar register
template <int MODE>
__global__ void __launch_bounds__(128) k(float *out, float a)
{
    const float t = threadIdx.x * 0.001f;
    u64 P[8]; float S[12];
    for (int i = 0; i < 8; i++) P[i] = pk(t + i, a * i);
    for (int i = 0; i < 12; i++) S[i] = t * i + a;
    const u64 pa = pk(a, a + t), pb = pk(t, a);
    const float sa = a + t, sb = a * t;
    u64 xp[6], yp[6], zp[6]; float xs[12], ys[12], zs[12];
    for (int i = 0; i < 6; i++) { xp[i] = pk(i + t, i - t); yp[i] = pk(a * i, t * i); zp[i] = pk(a + i, a - i); }
    for (int i = 0; i < 12; i++) { xs[i] = i + t; ys[i] = a * i + t; zs[i] = a - i * t; }
    for (int it = 0; it < IT; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) P[i] = (i & 1) ? fma2(P[i], pa, pb) : sub2(P[i], pa);
        } else if (MODE == 7) {
#pragma unroll
            for (int i = 0; i < 8; i++) P[i] = sub2(P[i], pk(sa, sa));
        } else if (MODE == 2 || MODE == 3) {
#pragma unroll
            for (int i = 0; i < 4; i++) P[i] = fma2(P[i], pa, pb);
#pragma unroll
            for (int i = 0; i < (MODE == 2 ? 4 : 8); i++) S[i] = ffma(S[i], sa, sb);
        }
        if (MODE == 1 || MODE == 5 || MODE == 6) {
#pragma unroll
            for (int g = 0; g < 6; g++) {
                const u64 dx = sub2(P[g], xp[g]), dy = sub2(pa, yp[g]), dz = sub2(pb, zp[g]);
                P[g] = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
            }
        }
        if (MODE == 4 || MODE == 5 || MODE == 6) {
#pragma unroll
            for (int g = 0; g < (MODE == 6 ? 3 : (MODE == 5 ? 6 : 12)); g++) {
                const float dx = fsub(S[g], xs[g]), dy = fsub(sa, ys[g]), dz = fsub(sb, zs[g]);
                S[g] = ffma(dz, dz, ffma(dy, dy, fmul(dx, dx)));
            }
        }
    }
    u64 s = 0; float f = 0;
    for (int i = 0; i < 8; i++) s ^= P[i];
    for (int i = 0; i < 12; i++) f += S[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = f + __uint_as_float((unsigned)(s ^ (s >> 32)));
}
static int nsm, clk_khz; static float *out;
template <int MODE> static void run(const char *name, int npacked, int nscalar) {
    const int grid = nsm * 8;
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    k<MODE><<<grid, 128>>>(out, 1.0001f); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a)); for (int i = 0; i < 3; i++) k<MODE><<<grid, 128>>>(out, 1.0001f); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); ms /= 3;
    const double warps = grid * 4.0, cyc = ms * 1e-3 * clk_khz * 1e3;  // SM cycles
    const double per_smsp_cycles_per_iter = cyc / (warps / nsm / 4.0) / IT;
    printf("%-52s %7.3f ms  %6.1f SMSP-cycles/iter/warp  for %2d packed + %2d scalar  => lane-ops/clk/SM = %.1f\n", name, ms, per_smsp_cycles_per_iter,
           npacked, nscalar, (npacked * 64.0 + nscalar * 32.0) / per_smsp_cycles_per_iter * 4);
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0)); nsm = p.multiProcessorCount;
    CK(cudaMalloc(&out, 4 * nsm * 8 * 128));
    run<0>("alternating FADD2/FFMA2 chains", 8, 0);
    run<7>("FADD2 with broadcast scalar operand", 8, 0);
    run<1>("packed distance sequence x6", 36, 0);
    run<4>("scalar distance sequence x12", 0, 72);
    run<2>("4 FFMA2 + 4 FFMA", 4, 4);
    run<3>("4 FFMA2 + 8 FFMA", 4, 8);
    run<5>("packed distance x6 + scalar distance x6", 36, 36);
    run<6>("packed distance x6 + scalar distance x3", 36, 18);
    return 0;
}
