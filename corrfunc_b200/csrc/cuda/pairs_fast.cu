// pairs_fast.cu -- the hot kernel: pair counts for the 1-D statistics (DD, xi, wp, DDtheta) when no
// per-pair averages or weights are requested (the BASELINE configs c1, c2-wp, c4, c5).
//
// Replaces the per-cell-pair CPU kernels (theory/DD/countpairs_kernels.c.src:25-279,
// theory/xi/xi_kernels.c.src:23-270, theory/wp/wp_kernels.c.src:23-305,
// mocks/DDtheta_mocks/countpairs_theta_mocks_kernels.c.src:872-1141) and the cell-pair enumeration
// of generate_cell_pairs_DOUBLE (utils/gridlink_impl.c.src:439-625).
//
// Design:
//   * one WARP owns one primary tile: up to 128 particles of one fine cell, 4 per lane in registers; warps are
//     persistent and pull their next tile from a global counter;
//   * phase 1: the 32 lanes test 32 candidate neighbour cells at a time.  From the two cells' particle
//     bounding boxes a lane derives a conservative interval [vlo, vhi] of every separation the pair
//     of cells can produce and, from it, the few bin edges that can actually split those pairs
//     ("levels").  No level -> the whole N1 x N2 block goes to one bin without touching a particle;
//   * phase 2: the neighbour's particles are staged in shared memory per warp, double buffered, with 16-byte
//     cp.async (default) or 1-D TMA bulk copies + mbarrier (CORRFUNC_B200_STAGE=tma), and every lane runs its
//     primaries against them with the reference's arithmetic: same subtraction order (second - (first + wrap)),
//     same FMA association.  Instead of searching a bin per pair, the warp keeps one cumulative
//     counter per level, #{v < edge}, in registers; bin counts are differences of those counters.
//     float : two pairs per instruction (sub/mul/fma.f32x2 with the primary broadcast by the instruction); the
//             indicator [v < edge] is the sign bit of one more packed subtract, added to the counter by LEA.HI;
//     double: plain compare-and-count.
//   * per-warp histogram of signed 32-bit deltas in shared memory (no atomics on the per-job path), flushed into
//     the block's 64-bit histogram before it can overflow; blocks merge with global atomics at the end.
//   The kernel is bound by warp-instruction issue (a packed f32x2 instruction holds the issue port two cycles)
//   and sensitive to the instruction-cache footprint of its inner-loop bodies: see DESIGN.md section 4.
//
// Compiled with -fmad=false: an FMA appears exactly where the reference's AVX-512 kernels have one.
#include <math_constants.h>
#include <stdlib.h>

#include "cfb_internal.cuh"

#define FAST_CH 128     // secondaries per staged chunk
#define FAST_WARPS 4    // warps (= primary tiles) per block
#define FAST_QCAP 128   // job queue entries per warp
#ifndef FAST_LMAX
#define FAST_LMAX 8     // levels per pass over a chunk (more levels -> more passes)
#endif
#ifndef FAST_UNROLL
#define FAST_UNROLL 1   // unroll factor of the hottest inner-loop variants (2 measured 15 % slower: instruction cache)
#endif
#ifndef FAST_SPI2_PA
#define FAST_SPI2_PA 4  // inner-loop bodies with >= this many primaries per lane and >= FAST_SPI2_NL levels
#endif                  // take 2 instead of 4 secondaries per iteration (instruction-cache footprint)
#ifndef FAST_SPI2_NL
#define FAST_SPI2_NL 2
#endif
#ifndef FAST_F64_SPI1
#define FAST_F64_SPI1 999 // double kernel: bodies with primaries x levels >= this take 1 secondary per iteration (off: 4, 8 and 12 measured 1-5 % slower)
#endif
#ifndef FAST_MINB_F32
#define FAST_MINB_F32 4 // resident blocks per SM the float kernel is compiled for (register budget); 5 and 6 measured no faster
#endif
#ifndef FAST_MINB_F64
#define FAST_MINB_F64 3
#endif
#define FAST_PRIM 4     // primaries per lane  (32 * 4 = CFB_TILE)
#ifndef FAST_SPLIT
#define FAST_SPLIT 1    // half-warp split of a last primary row that holds <= 16 particles (float DD / xi, <= 3 levels)
#endif

typedef unsigned long long u64;

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D TMA bulk copy, packed f32x2 arithmetic
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, u64 *b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *b, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ u64 pk(float a, float b)
{
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk(u64 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b)
{
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// a - {b, b}: the scalar operand is broadcast by the instruction (SASS operand form R.F32), no register pair needed
__device__ __forceinline__ u64 sub2s(u64 a, float b)
{
    u64 d;
    asm("{\n.reg .b64 t;\nmov.b64 t, {%2,%2};\nsub.rn.f32x2 %0, %1, t;\n}" : "=l"(d) : "l"(a), "f"(b));
    return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b)
{
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b)
{
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// sat(a - b): 1.0f when a - b >= 1, 0.0f when a <= b (also for b = NaN or a - b = NaN)
__device__ __forceinline__ float subsat(float a, float b)
{
    float d;
    asm("sub.rn.sat.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}

__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// exact n / d for n * d < 2^32, magic = ceil(2^32 / d) (0 stands for d == 1)
__device__ __forceinline__ unsigned fdiv(unsigned n, unsigned magic) { return magic ? __umulhi(n, magic) : n; }

// ------------------------------------------------------------------------------------------------
struct FastJob {
    int start;  // first secondary (index into the sorted arrays, multiple of CFB_PAD)
    int n;      // secondaries
    int meta;   // bits 0-5 wrap code | 6-12 kbase | 13-19 L | 20 tri | 21 zcut | 22 top level measured
};
#define JOB_TRI (1 << 20)
#define JOB_ZCUT (1 << 21)
#define JOB_TOPM (1 << 22)
#define JOB_DIRZ (1 << 23)  // wp, two different reference cells: the reference's one-sided pi cut (see chunk_f32)

template <typename T>
struct FastWarp {
    alignas(16) T buf[2][3][FAST_CH];
    u64 mbar[2];
    FastJob q[FAST_QCAP];
    int wh[CFB_FAST_MAX_EDGES + 2];  // the warp's private histogram of signed 32-bit deltas (flushed before it can overflow)
};

template <typename T>
struct FastShared {
    FastWarp<T> w[FAST_WARPS];
    u64 hist[CFB_FAST_MAX_EDGES + 1];
    double edges_d[CFB_FAST_MAX_EDGES];
    T edges[CFB_FAST_MAX_EDGES + FAST_LMAX];  // padded with +inf
};

// ------------------------------------------------------------------------------------------------
// One chunk of secondaries against the lane's 4 primaries; cnt[l] += #{pairs with v < E[l]}.
// SPLIT: the last of the PA primary rows holds at most 16 particles (a cell of 97-112 particles: 40 % of the tiles of
// config 5), so its upper half-warp would evaluate nothing but padding.  The two half-warps then share that row's
// primaries (lane l and lane l + 16 hold the same one, fetched by a shuffle) and split the SECONDARIES of every
// iteration between them: the row costs half its instructions.  Every (primary, secondary) pair is still evaluated
// exactly once, with the same operations.
template <int MODE, int NL, int PA, int ZCUT, bool SPLIT>
__device__ __forceinline__ void chunk_f32(const float *sx, const float *sy, const float *sz, const int m4,
                                          const float (&xq)[FAST_PRIM], const float (&yq)[FAST_PRIM],
                                          const float (&zq)[FAST_PRIM], const float *Es, const float pimax,
                                          const bool dirz, int (&cnt)[FAST_LMAX])
{
    float E[NL];
    unsigned c[NL];
#pragma unroll
    for (int l = 0; l < NL; l++) {
        E[l] = Es[l];
        c[l] = 0u;
    }
    const u64 th_a = pk(8388608.0f, 8388608.0f), th_b = pk(-16777216.0f, -16777216.0f);
    float tz[PA];  // target_z = zpos - pimax of the reference's fast-forward (ZCUT == 2 only)
#pragma unroll
    for (int p = 0; p < PA; p++) tz[p] = zq[p] - pimax;
    // The kernel holds one loop per (levels, primaries) variant and the warps of an SM run different ones:
    // unrolling them all 4x overflowed the instruction cache (measured: 26 "no instruction" stall cycles per
    // issued instruction, 5x slower).  Only the variants that carry ~95 % of the iterations (3 primaries per
    // lane, 1-3 levels) are unrolled, by 2: the sub-partition is issue bound and the 7 loop-control
    // instructions are 5-7 % of such an iteration.
    constexpr int UNR = (PA == 3 && NL <= 3) ? FAST_UNROLL : 1;
    // Secondaries per iteration: 4 (three LDS.128), or 2 for the long bodies (many primaries x levels) -- a body
    // must stay small enough that the loops of the four warps of a sub-partition fit its instruction cache
    // together (measured with ncu: 24 % of all warp samples were instruction-fetch stalls at every 128-byte
    // line of the 100+ instruction bodies)
    constexpr int SPI = (PA >= FAST_SPI2_PA && NL >= FAST_SPI2_NL) ? 2 : 4;
    constexpr int H = SPI / 2;
    constexpr int PFULL = SPLIT ? PA - 1 : PA;  // rows evaluated by every lane against every secondary
    float xl = 0, yl = 0, zl = 0;
    int hoff = 0;
    if (SPLIT) {
        const int lane = threadIdx.x & 31;
        xl = __shfl_sync(0xffffffffu, xq[PA - 1], lane & 15);
        yl = __shfl_sync(0xffffffffu, yq[PA - 1], lane & 15);
        zl = __shfl_sync(0xffffffffu, zq[PA - 1], lane & 15);
        hoff = lane >> 4;
    }
#pragma unroll UNR
    for (int j = 0; j < m4; j += SPI) {
        u64 xs[H], ys[H], zs[H];
        if constexpr (SPI == 4) {
            const float4 X = *reinterpret_cast<const float4 *>(sx + j);
            const float4 Y = *reinterpret_cast<const float4 *>(sy + j);
            const float4 Z = *reinterpret_cast<const float4 *>(sz + j);
            xs[0] = pk(X.x, X.y), xs[1] = pk(X.z, X.w);
            ys[0] = pk(Y.x, Y.y), ys[1] = pk(Y.z, Y.w);
            zs[0] = pk(Z.x, Z.y), zs[1] = pk(Z.z, Z.w);
        } else {
            const float2 X = *reinterpret_cast<const float2 *>(sx + j);
            const float2 Y = *reinterpret_cast<const float2 *>(sy + j);
            const float2 Z = *reinterpret_cast<const float2 *>(sz + j);
            xs[0] = pk(X.x, X.y);
            ys[0] = pk(Y.x, Y.y);
            zs[0] = pk(Z.x, Z.y);
        }
        if (SPLIT) {
            // the shared last row: this half-warp's half of the iteration's secondaries
            if constexpr (SPI == 4) {
                const float2 X = *reinterpret_cast<const float2 *>(sx + j + 2 * hoff);
                const float2 Y = *reinterpret_cast<const float2 *>(sy + j + 2 * hoff);
                const float2 Z = *reinterpret_cast<const float2 *>(sz + j + 2 * hoff);
                const u64 dx = sub2s(pk(X.x, X.y), xl), dy = sub2s(pk(Y.x, Y.y), yl), dz = sub2s(pk(Z.x, Z.y), zl);
                const u64 v2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
#pragma unroll
                for (int l = 0; l < NL; l++) {
                    const u64 d = sub2s(v2, E[l]);
                    c[l] += (unsigned)d >> 31;
                    c[l] += (unsigned)(d >> 63);
                }
            } else {
                const float dx = sx[j + hoff] - xl, dy = sy[j + hoff] - yl, dz = sz[j + hoff] - zl;
                const float v = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
#pragma unroll
                for (int l = 0; l < NL; l++) c[l] += __float_as_uint(v - E[l]) >> 31;
            }
        }
#pragma unroll
        for (int p = 0; p < PFULL; p++) {
#pragma unroll
            for (int h = 0; h < H; h++) {
                const u64 dx = sub2s(xs[h], xq[p]), dy = sub2s(ys[h], yq[p]), dz = sub2s(zs[h], zq[p]);
                u64 v2;
                if (MODE == CFB_WP) {
                    v2 = fma2(dy, dy, mul2(dx, dx));  // wp_kernels.c.src:197-198
                } else {
                    v2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));  // countpairs_kernels.c.src:188-190
                    // -cos(theta) * 2^24 = chord^2 * 2^23 - 2^24 (countpairs_theta_mocks_kernels.c.src:1062-1066)
                    if (MODE == CFB_THETA) v2 = fma2(v2, th_a, th_b);
                }
                if (ZCUT == 1) {  // same reference cell (j after i in z order, dz >= 0): dz < pimax, here as |dz| < pimax
                    float v0, v1, z0, z1;
                    upk(v2, v0, v1);
                    upk(dz, z0, z1);
                    v0 = fabsf(z0) < pimax ? v0 : CUDART_INF_F;
                    v1 = fabsf(z1) < pimax ? v1 : CUDART_INF_F;
                    v2 = pk(v0, v1);
                }
                if (ZCUT == 2) {
                    // two reference cells: the reference fast-forwards over the secondaries with z1 <= zpos - pimax
                    // (wp_kernels.c.src:139-142) and then masks with the SIGNED dz < pimax (:207-221).  A survivor
                    // whose dz rounds to exactly -pimax is therefore counted; |dz| < pimax would drop it.
                    float v0, v1, z0, z1, s0, s1;
                    upk(v2, v0, v1);
                    upk(dz, z0, z1);
                    upk(zs[h], s0, s1);
                    const bool k0 = (s0 > tz[p]) & (z0 < pimax), k1 = (s1 > tz[p]) & (z1 < pimax);
                    v0 = k0 ? v0 : CUDART_INF_F;
                    v1 = k1 ? v1 : CUDART_INF_F;
                    v2 = pk(v0, v1);
                }
                // [v < E] is the sign bit of the rounded difference v - E (x - x = +0; NaN and +inf give
                // a non-negative result): one packed subtract on the FMA pipe for two pairs, then the
                // sign bits are added to the level's counter on the integer pipe (LEA.HI)
#pragma unroll
                for (int l = 0; l < NL; l++) {
                    const u64 d = sub2s(v2, E[l]);
                    c[l] += (unsigned)d >> 31;
                    c[l] += (unsigned)(d >> 63);
                }
            }
        }
    }
#pragma unroll
    for (int l = 0; l < NL; l++) cnt[l] += (int)c[l];
}

template <int MODE, int NL, int PA, int ZCUT>
__device__ __forceinline__ void chunk_f64(const double *sx, const double *sy, const double *sz, const int m4,
                                          const double (&xq)[FAST_PRIM], const double (&yq)[FAST_PRIM],
                                          const double (&zq)[FAST_PRIM], const double *Es,
                                          const double pimax, const bool dirz, int (&cnt)[FAST_LMAX])
{
    double E[NL];
#pragma unroll
    for (int l = 0; l < NL; l++) E[l] = Es[l];
    double tz[PA];  // see chunk_f32 (ZCUT == 2 only)
#pragma unroll
    for (int p = 0; p < PA; p++) tz[p] = zq[p] - pimax;
    // one secondary per iteration for the long bodies (many primaries x levels): same instruction-cache
    // consideration as in chunk_f32
    constexpr int SPI = (PA * NL >= FAST_F64_SPI1) ? 1 : 2;
#pragma unroll 1
    for (int j = 0; j < m4; j += SPI) {
        double xs[SPI], ys[SPI], zs[SPI];
        if constexpr (SPI == 2) {
            const double2 X = *reinterpret_cast<const double2 *>(sx + j);
            const double2 Y = *reinterpret_cast<const double2 *>(sy + j);
            const double2 Z = *reinterpret_cast<const double2 *>(sz + j);
            xs[0] = X.x, xs[1] = X.y, ys[0] = Y.x, ys[1] = Y.y, zs[0] = Z.x, zs[1] = Z.y;
        } else {
            xs[0] = sx[j], ys[0] = sy[j], zs[0] = sz[j];
        }
#pragma unroll
        for (int h = 0; h < SPI; h++) {
#pragma unroll
            for (int p = 0; p < PA; p++) {
                const double dx = xs[h] - xq[p], dy = ys[h] - yq[p], dz = zs[h] - zq[p];
                double v;
                if (MODE == CFB_WP) {
                    v = __fma_rn(dy, dy, dx * dx);
                } else {
                    v = __fma_rn(dz, dz, __fma_rn(dy, dy, dx * dx));
                    if (MODE == CFB_THETA) v = __fma_rn(v, 0.5, -1.0);  // -(1 - chord^2/2), exactly
                }
                // masked pairs get the high word of +inf (with the old low word that is +inf or a NaN: never below an edge):
                // one select instead of two
                if (ZCUT == 1) v = __hiloint2double(fabs(dz) < pimax ? __double2hiint(v) : 0x7ff00000, __double2loint(v));
                if (ZCUT == 2) {  // see chunk_f32
                    const bool kz = (zs[h] > tz[p]) & (dz < pimax);
                    v = __hiloint2double(kz ? __double2hiint(v) : 0x7ff00000, __double2loint(v));
                }
#pragma unroll
                for (int l = 0; l < NL; l++) cnt[l] += (v < E[l]) ? 1 : 0;
            }
        }
    }
}

template <typename T, int MODE, int NL, int PA, int ZCUT, bool SPLIT = false>
__device__ __forceinline__ void chunk_T(const T *sx, const T *sy, const T *sz, const int m4, const T (&xq)[FAST_PRIM],
                                        const T (&yq)[FAST_PRIM], const T (&zq)[FAST_PRIM], const T *E,
                                        const T pimax, const bool dirz, int (&cnt)[FAST_LMAX])
{
    if constexpr (sizeof(T) == 4)
        chunk_f32<MODE, NL, PA, ZCUT, SPLIT>(sx, sy, sz, m4, xq, yq, zq, E, pimax, dirz, cnt);
    else
        chunk_f64<MODE, NL, PA, ZCUT>(sx, sy, sz, m4, xq, yq, zq, E, pimax, dirz, cnt);
}

template <typename T, int MODE, int PA, int ZCUT>
__device__ __forceinline__ void chunk_dispatch_nl(const int nl, const T *sx, const T *sy, const T *sz, const int m4,
                                                  const T (&xq)[FAST_PRIM], const T (&yq)[FAST_PRIM],
                                                  const T (&zq)[FAST_PRIM], const T *E, const T pimax,
                                                  const bool dirz, int (&cnt)[FAST_LMAX])
{
    switch (nl) {
    case 1: chunk_T<T, MODE, 1, PA, ZCUT>(sx, sy, sz, m4, xq, yq, zq, E, pimax, dirz, cnt); break;
    case 2: chunk_T<T, MODE, 2, PA, ZCUT>(sx, sy, sz, m4, xq, yq, zq, E, pimax, dirz, cnt); break;
    case 3: chunk_T<T, MODE, 3, PA, ZCUT>(sx, sy, sz, m4, xq, yq, zq, E, pimax, dirz, cnt); break;
#if FAST_LMAX == 4
    default: chunk_T<T, MODE, 4, PA, ZCUT>(sx, sy, sz, m4, xq, yq, zq, E, pimax, dirz, cnt); break;
#else
    case 4: chunk_T<T, MODE, 4, PA, ZCUT>(sx, sy, sz, m4, xq, yq, zq, E, pimax, dirz, cnt); break;
    case 5: chunk_T<T, MODE, 5, PA, ZCUT>(sx, sy, sz, m4, xq, yq, zq, E, pimax, dirz, cnt); break;
    case 6: chunk_T<T, MODE, 6, PA, ZCUT>(sx, sy, sz, m4, xq, yq, zq, E, pimax, dirz, cnt); break;
    default: chunk_T<T, MODE, 8, PA, ZCUT>(sx, sy, sz, m4, xq, yq, zq, E, pimax, dirz, cnt); break;  // 7 runs as 8 (E[7] = +inf)
#endif
    }
}

// pa = primaries per lane actually in use in this tile (1..4): empty register slots are not evaluated
template <typename T, int MODE, int ZCUT>
__device__ __forceinline__ void chunk_dispatch(const int nl, const int pa, const T *sx, const T *sy, const T *sz,
                                               const int m4, const T (&xq)[FAST_PRIM], const T (&yq)[FAST_PRIM],
                                               const T (&zq)[FAST_PRIM], const T *E, const T pimax,
                                               const bool dirz, int (&cnt)[FAST_LMAX])
{
    switch (pa) {
    case 1: chunk_dispatch_nl<T, MODE, 1, ZCUT>(nl, sx, sy, sz, m4, xq, yq, zq, E, pimax, dirz, cnt); break;
    case 2: chunk_dispatch_nl<T, MODE, 2, ZCUT>(nl, sx, sy, sz, m4, xq, yq, zq, E, pimax, dirz, cnt); break;
    case 3: chunk_dispatch_nl<T, MODE, 3, ZCUT>(nl, sx, sy, sz, m4, xq, yq, zq, E, pimax, dirz, cnt); break;
    default: chunk_dispatch_nl<T, MODE, 4, ZCUT>(nl, sx, sy, sz, m4, xq, yq, zq, E, pimax, dirz, cnt); break;
    }
}

// ------------------------------------------------------------------------------------------------
// Staging of one chunk of secondaries (x, y, z runs of m4 elements) into the warp's buffer `bsel`.
//   TMA = true : three 1-D bulk copies issued by lane 0, completion on the buffer's mbarrier
//   TMA = false: 16-byte cp.async per lane, completion through cp.async groups
template <typename T, bool TMA>
__device__ __forceinline__ void stage_chunk(FastWarp<T> &W, const int bsel, const SetView<T> &B, const int first,
                                            const int m4, const int lane)
{
    if (TMA) {
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)m4 * sizeof(T);
            mbar_expect_tx(&W.mbar[bsel], 3 * bytes);
            tma_load_1d(W.buf[bsel][0], B.x + first, bytes, &W.mbar[bsel]);
            tma_load_1d(W.buf[bsel][1], B.y + first, bytes, &W.mbar[bsel]);
            tma_load_1d(W.buf[bsel][2], B.z + first, bytes, &W.mbar[bsel]);
        }
    } else {
        constexpr int EPV = 16 / (int)sizeof(T);
        for (int v = lane * EPV; v < m4; v += 32 * EPV) {
            cp_async16(&W.buf[bsel][0][v], B.x + first + v);
            cp_async16(&W.buf[bsel][1][v], B.y + first + v);
            cp_async16(&W.buf[bsel][2][v], B.z + first + v);
        }
        cp_async_commit();
    }
}

// adds the warp's signed 32-bit deltas to the block's 64-bit histogram and clears them
__device__ __forceinline__ void flush_warp_hist(int *wh, u64 *hist, const int nedges, const int lane)
{
    __syncwarp();
    for (int i = lane; i <= nedges; i += 32) {
        const int v = wh[i];
        if (v != 0) {
            atomicAdd(&hist[i], (u64)(long long)v);
            wh[i] = 0;
        }
    }
    __syncwarp();
}

template <typename T, int MODE, bool LIST, bool TMA>
__global__ void __launch_bounds__(FAST_WARPS * 32, sizeof(T) == 4 ? FAST_MINB_F32 : FAST_MINB_F64)
k_pairs_fast(const PairParams P, const SetView<T> A, const SetView<T> B)
{
    __shared__ __align__(16) FastShared<T> S;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nedges = P.nedges;
    for (int i = tid; i <= nedges; i += blockDim.x) S.hist[i] = 0ULL;
    for (int i = tid; i < nedges + FAST_LMAX; i += blockDim.x) {
        const T e = i < nedges ? ((const T *)P.edges)[i] : (sizeof(T) == 4 ? (T)CUDART_INF_F : (T)CUDART_INF);
        S.edges[i] = e;
        if (i < nedges) S.edges_d[i] = (double)e;
    }
    FastWarp<T> &W = S.w[wid];
    for (int i = lane; i < CFB_FAST_MAX_EDGES + 2; i += 32) W.wh[i] = 0;
    if (TMA && lane == 0) {
        mbar_init(&W.mbar[0], 1);
        mbar_init(&W.mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // Persistent warps: every warp pulls its next primary tile from a global counter (tiles differ in
    // work by the cell occupancies; one tile per warp left 7 % of the warp time waiting at the block's
    // final barrier).  Across ranks the work is sharded by primary cell (cfb_owns_cell).
    u64 my_eval = 0, my_jobs = 0, my_analytic = 0, my_levels = 0;
    unsigned wbound = 0;  // warp-uniform bound on the magnitude of any slot of W.wh
    uint32_t it = 0;  // chunks staged so far by this warp (buffer = it & 1, TMA parity = (it >> 1) & 1)

    for (;;) {
        long long gw = 0;
        if (lane == 0) {
            gw = (long long)atomicAdd(&P.counters[4], 1ULL);
            // interrupt (SIGINT / SIGTERM / SIGHUP caught by the host layer, cf. countpairs_impl.c.src:475-477): every 64th
            // fetch also reads the host's flag (a read over PCIe: doing it at every fetch cost config 1 a factor of four)
            // and, when it is set, pushes the tile counter past the end -- every later fetch of every warp then ends its loop
            if ((gw & 63) == 0 && P.abort && *P.abort) gw = (long long)atomicAdd(&P.counters[4], 1ULL << 40) + ((long long)1 << 40);
        }
        gw = __shfl_sync(0xffffffffu, gw, 0);
        // Work units: whole tiles, except that the last tiles of the list are handed out in pieces (groups of neighbour
        // rows).  A tile of config 5 is 12 ms of one warp's time; when the list runs out the warps stop at random points
        // of their last tile, and the kernel ends with its slowest warp: ~6 ms of idle SMs per launch -- 0.1 % of a 5-s
        // launch, but 1 % of the 640 ms an eighth of the work takes (tools/exp_shard.py: rank r of 8 emulated on one GPU
        // took 1.2-1.5 % longer than an eighth of the full kernel, whichever way the cells were dealt to the ranks).
        if (gw >= P.tail_first + (P.ntiles - P.tail_first) * P.tail_parts) break;
        int64_t tile = gw;
        int part = 0, parts = 1;
        if (gw >= P.tail_first) {
            const int64_t u = gw - P.tail_first;
            tile = P.tail_first + u / P.tail_parts;
            part = (int)(u % P.tail_parts);
            parts = P.tail_parts;
        }
        const int cellP = P.tile_cell[tile];
        if (!cfb_owns_cell(cellP, P.shard_rank, P.shard_n)) continue;  // another rank's cell (see cfb_owns_cell)
        const int toff = P.tile_off[tile];
        const int nP = A.count[cellP];
        const int startP = A.start[cellP];
        const int nv = min(CFB_TILE, nP - toff);  // valid primaries of this tile
        const int pa = (nv + 31) >> 5;            // primaries per lane in use
        const bool tsplit = pa == 4 && nv <= 112; // the fourth row fits a half-warp (see chunk_f32, SPLIT)
        const T nanv = sizeof(T) == 4 ? (T)CUDART_NAN_F : (T)CUDART_NAN;
        const T *pxg = A.x + startP + toff, *pyg = A.y + startP + toff, *pzg = A.z + startP + toff;

        int gx = 0, gy = 0, gz = 0, rax = 0, ray = 0, raz = 0;
        long long refA = 0;
        int nrow = 1, rowlen, wz = 1, rx = 0, ry = 0, rz = 0;
        int64_t list0 = 0;
        const int sub2 = LIST ? P.list_sub2 : 1;  // fine cells per reference cell of the RA/DEC lattice
        const int refP = LIST ? cellP / sub2 : 0;
        if (LIST) {
            list0 = P.list_off[refP];
            rowlen = (int)(P.list_off[refP + 1] - list0) * sub2;  // every fine cell of every listed reference cell
        } else {
            gz = cellP % P.g.ng[2];
            gy = (cellP / P.g.ng[2]) % P.g.ng[1];
            gx = cellP / (P.g.ng[2] * P.g.ng[1]);
            rax = gx / P.g.s[0];
            ray = gy / P.g.s[1];
            raz = gz / P.g.s[2];
            refA = ((long long)rax * P.g.n[1] + ray) * P.g.n[2] + raz;
            rx = P.g.reach[0];
            ry = P.g.reach[1];
            rz = P.g.reach[2];
            wz = 2 * rz + 1;
            nrow = 2 * rx + 1;             // one "row" of candidates per x offset
            rowlen = (2 * ry + 1) * wz;    // (y, z) offsets, 32 at a time across the lanes
        }
        // relative error bound of one rounded operation (with slack), in double
        const double eps = sizeof(T) == 4 ? 2.384185791015625e-07 /* 2^-22 */ : 4.440892098500626e-16 /* 2^-51 */;
        const T pimax = (T)P.pimax;
        const double v_self = MODE == CFB_THETA ? (sizeof(T) == 4 ? -16777216.0 : -1.0) : 0.0;

        int qn = 0;       // queued jobs (warp-uniform)
        int cur_code = 0;  // wrap code the shifted primaries xq/yq/zq currently hold
        T xq[FAST_PRIM], yq[FAST_PRIM], zq[FAST_PRIM];
#pragma unroll
        for (int p = 0; p < FAST_PRIM; p++) {
            const int i = lane + 32 * p;
            const bool ok = i < nv;
            xq[p] = ok ? pxg[i] : nanv;
            yq[p] = ok ? pyg[i] : nanv;
            zq[p] = ok ? pzg[i] : nanv;
        }

        // this unit's candidates (all of them for a whole tile).  Box lattices: tail_parts = the number of rows, a piece is
        // one row of neighbour cells; neighbour lists (DDtheta, one "row"): a piece is an eighth of the list's rounds
        int row_lo = 0, row_hi = nrow, base_lo = 0, base_hi = rowlen;
        if (parts > 1) {
            if (LIST) {
                const int nch = (rowlen + 31) >> 5;  // rounds of 32 candidates
                base_lo = (int)((long long)part * nch / parts) << 5;
                base_hi = min(rowlen, (int)((long long)(part + 1) * nch / parts) << 5);
            } else {
                row_lo = (int)((long long)part * nrow / parts);
                row_hi = (int)((long long)(part + 1) * nrow / parts);
            }
        }
        for (int row = row_lo; row < row_hi; row++) {
            for (int base = base_lo; base < base_hi; base += 32) {
                // ---------------- phase 1: one candidate per lane ----------------
                const int cand = base + lane;
                bool keep = cand < rowlen;
                int cellQ = -1, code = 0, kbase = 0, nl = 0, flags = 0;
                int j_start = 0, j_n = 0, j2_start = 0, j2_n = 0;  // same cell: diagonal + rectangle jobs
                unsigned npairs_w = 0;
                if (keep) {
                    double offd[3] = {0.0, 0.0, 0.0};
                    if (LIST) {
                        const int li = cand / sub2;
                        const int refQ = P.list_cells[list0 + li];
                        cellQ = refQ * sub2 + (cand - li * sub2);
                        // the host lists every unordered pair of reference cells once (and a cell with itself):
                        // inside one reference cell every unordered pair of fine cells is taken once
                        if (P.autocorr && refQ == refP && cellQ > cellP) keep = false;
                        if (!(P.autocorr && refQ == refP)) flags |= JOB_DIRZ;
                    } else {
                        const unsigned iy = fdiv((unsigned)cand, P.m_wz);
                        const int t[3] = {gx + row - rx, gy + (int)iy - ry, gz + (cand - (int)iy * wz) - rz};
                        const int ra[3] = {rax, ray, raz};
                        int q[3], rb[3];
#pragma unroll
                        for (int a = 0; a < 3; a++) {
                            const int ng = P.g.ng[a];
                            if (P.g.periodic[a]) {
                                if (t[a] < -ng || t[a] >= 2 * ng) keep = false;
                            } else if (t[a] < 0 || t[a] >= ng)
                                keep = false;
                            if (!keep) break;
                            // neighbour reference cell within +-refine of the primary's (gridlink_impl.c.src:499-518)
                            const int rt = (int)fdiv((unsigned)(t[a] + ng), P.m_s[a]) - P.g.n[a];
                            const int dref = rt - ra[a];
                            if (dref > P.g.refine[a] || dref < -P.g.refine[a]) keep = false;
                            if (t[a] < 0) {
                                q[a] = t[a] + ng;
                                rb[a] = rt + P.g.n[a];
                                code |= 1 << (2 * a);  // +wrap on the first particle (gridlink_impl.c.src:504)
                                offd[a] = P.wrap[a];
                            } else if (t[a] >= ng) {
                                q[a] = t[a] - ng;
                                rb[a] = rt - P.g.n[a];
                                code |= 2 << (2 * a);
                                offd[a] = -P.wrap[a];
                            } else {
                                q[a] = t[a];
                                rb[a] = rt;
                            }
                        }
                        if (keep) {
                            cellQ = (q[0] * P.g.ng[1] + q[1]) * P.g.ng[2] + q[2];
                            if (P.autocorr) {
                                // reference keeps icell2 <= icell (gridlink_impl.c.src:525); within one reference
                                // cell every unordered pair of fine cells is taken once
                                const long long refB = ((long long)rb[0] * P.g.n[1] + rb[1]) * P.g.n[2] + rb[2];
                                if (refB > refA || (refB == refA && cellQ < cellP)) keep = false;
                                if (refB != refA) flags |= JOB_DIRZ;
                                // a cell against its own periodic image: |d| >= L/2 > rmax, nothing to count
                                if (cellQ == cellP && code != 0) keep = false;
                            } else
                                flags |= JOB_DIRZ;
                        }
                    }
                    int nQ = 0;
                    if (keep) {
                        nQ = B.count[cellQ];
                        if (nQ == 0) keep = false;
                    }
                    if (keep) {
                        // ---- conservative interval of v over all pairs of the two cells ----
                        double dmin[3], dmax[3];
#pragma unroll
                        for (int a = 0; a < 3; a++) {
                            const double plo = (double)A.bounds[(int64_t)cellP * CFB_NB + 2 * a] + offd[a];
                            const double phi = (double)A.bounds[(int64_t)cellP * CFB_NB + 2 * a + 1] + offd[a];
                            const double qlo = (double)B.bounds[(int64_t)cellQ * CFB_NB + 2 * a];
                            const double qhi = (double)B.bounds[(int64_t)cellQ * CFB_NB + 2 * a + 1];
                            double lo = 0.0;
                            if (qlo > phi) lo = qlo - phi;
                            else if (plo > qhi) lo = plo - qhi;
                            const double hi = fmax(qhi - plo, phi - qlo);
                            // rounding of (first + wrap) and of the subtraction
                            const double del = eps * (fmax(fabs(plo), fabs(phi)) + hi);
                            dmin[a] = fmax(0.0, lo - del);
                            dmax[a] = hi + del;
                        }
                        double vlo, vhi;
                        if (MODE == CFB_WP) {
                            vlo = dmin[0] * dmin[0] + dmin[1] * dmin[1];
                            vhi = dmax[0] * dmax[0] + dmax[1] * dmax[1];
                        } else {
                            vlo = dmin[0] * dmin[0] + dmin[1] * dmin[1] + dmin[2] * dmin[2];
                            vhi = dmax[0] * dmax[0] + dmax[1] * dmax[1] + dmax[2] * dmax[2];
                        }
                        vlo *= (1.0 - 2.0 * eps);
                        vhi *= (1.0 + 2.0 * eps);
                        if (MODE == CFB_THETA) {
                            if (sizeof(T) == 4) {
                                vlo = 16777216.0 * (0.5 * vlo - 1.0) - 4.0;
                                vhi = 16777216.0 * (0.5 * vhi - 1.0) + 4.0;
                            } else {
                                vlo = (0.5 * vlo - 1.0) - 4.0 * eps;
                                vhi = (0.5 * vhi - 1.0) + 4.0 * eps;
                            }
                        }
                        bool zpartial = false;
                        if (MODE == CFB_WP) {
                            const double pm = P.pimax;
                            if (dmin[2] >= pm) keep = false;           // every |dz| >= pimax
                            else if (!(dmax[2] < pm)) zpartial = true;  // some pairs may fail the cut
                        }
                        if (vlo >= S.edges_d[nedges - 1] || vhi < S.edges_d[0]) keep = false;
                        if (keep) {
                            // klo = largest k with E[k] <= vlo, khi = smallest k with E[k] > vhi
                            int a = 0, b = nedges;  // first k with E[k] > vlo
                            while (a < b) {
                                const int m = (a + b) >> 1;
                                if (S.edges_d[m] <= vlo) a = m + 1; else b = m;
                            }
                            const int klo = a - 1;
                            b = nedges;  // first k with E[k] > vhi (>= the previous answer)
                            while (a < b) {
                                const int m = (a + b) >> 1;
                                if (S.edges_d[m] <= vhi) a = m + 1; else b = m;
                            }
                            const int khi = a;
                            kbase = klo + 1;
                            nl = khi - klo - 1;
                            const bool tri = P.autocorr && cellQ == cellP;
                            const int startQ = B.start[cellQ];
                            u64 npairs_an;  // pairs this candidate contributes if none is cut
                            if (tri) {
                                // same cell: the tile against itself + the secondaries after this tile
                                const int after = nP - (toff + CFB_TILE);
                                j_start = startQ + toff;
                                j_n = nv;
                                flags |= JOB_TRI;
                                if (after > 0) {
                                    j2_start = startQ + toff + CFB_TILE;
                                    j2_n = after;
                                }
                                npairs_an = (u64)nv * (u64)(nv - 1) / 2 + (u64)nv * (u64)(after > 0 ? after : 0);
                            } else {
                                j_start = startQ;
                                j_n = nQ;
                                npairs_an = (u64)nv * (u64)nQ;
                            }
                            if (zpartial) {
                                flags |= JOB_ZCUT;
                                if (khi < nedges) {  // the count below the top edge must be measured too
                                    nl += 1;
                                    flags |= JOB_TOPM;
                                }
                            } else if (npairs_an < (1ULL << 24)) {
                                atomicAdd(&W.wh[khi], (int)npairs_an);  // everything is below E[khi]
                            } else {
                                atomicAdd(&S.hist[khi], npairs_an);  // too big for the 32-bit warp histogram (huge theta cells)
                            }
                            if (!zpartial && npairs_an < (1ULL << 24)) npairs_w = (unsigned)npairs_an;  // booked into W.wh
                            if (nl == 0) {
                                keep = false;  // one bin for the whole block of pairs: nothing to evaluate
                                my_analytic += npairs_an;
                            } else {
                                my_eval += npairs_an;
                                my_levels += npairs_an * (u64)nl;
                            }
                        }
                    }
                }
                // ---------------- push the survivors ----------------
                {
                    const unsigned m1 = __ballot_sync(0xffffffffu, keep);
                    const unsigned m2 = __ballot_sync(0xffffffffu, keep && j2_n > 0);
                    const unsigned lt = (1u << lane) - 1u;
                    const int meta = code | (kbase << 6) | (nl << 13) | flags;
                    if (keep) {
                        FastJob jb;
                        jb.start = j_start;
                        jb.n = j_n;
                        jb.meta = meta;
                        W.q[qn + __popc(m1 & lt)] = jb;
                        if (j2_n > 0) {
                            jb.start = j2_start;
                            jb.n = j2_n;
                            jb.meta = meta & ~JOB_TRI;
                            W.q[qn + __popc(m1) + __popc(m2 & lt)] = jb;
                        }
                    }
                    qn += __popc(m1) + __popc(m2);
                    __syncwarp();
                }
                // overflow guard of the warp's 32-bit histogram: |slot| <= wbound, the pairs booked since the
                // last flush (<= 32 * 2^24 more per round here, <= 2^14 more per chunk pass below)
                wbound += __reduce_add_sync(0xffffffffu, npairs_w);
                if (wbound >= (1u << 30)) {
                    flush_warp_hist(W.wh, S.hist, nedges, lane);
                    wbound = 0;
                }
                // at most 2 * 32 new jobs per round: drain before another round could overflow the queue
                const bool last = (row == row_hi - 1) && (base + 32 >= base_hi);
                if (qn <= FAST_QCAP - 64 && !last) continue;

                // ---------------- phase 2: drain the queue ----------------
                my_jobs += qn;
                int e = 0, c0 = 0;  // job being computed, its chunk offset
                if (qn > 0) {
                    const FastJob jb = W.q[0];
                    stage_chunk<T, TMA>(W, it & 1, B, jb.start, (min(FAST_CH, jb.n) + 3) & ~3, lane);
                }
                while (e < qn) {
                    const FastJob jb = W.q[e];
                    const int m4 = (min(FAST_CH, jb.n - c0) + 3) & ~3;
                    // next chunk: same job or the next one
                    int e2 = e, c2 = c0 + FAST_CH;
                    if (c2 >= jb.n) {
                        e2 = e + 1;
                        c2 = 0;
                    }
                    const int bsel = it & 1;
                    if (e2 < qn) {
                        const FastJob jn = W.q[e2];
                        stage_chunk<T, TMA>(W, bsel ^ 1, B, jn.start + c2, (min(FAST_CH, jn.n - c2) + 3) & ~3, lane);
                        if (!TMA) cp_async_wait<1>();
                    } else if (!TMA)
                        cp_async_wait<0>();
                    if (TMA) mbar_wait(&W.mbar[bsel], (it >> 1) & 1);
                    else __syncwarp();
                    const T *sx = W.buf[bsel][0], *sy = W.buf[bsel][1], *sz = W.buf[bsel][2];

                    // first particle gets the wrap: xpos = x0 + off_xwrap (countpairs_kernels.c.src:79-81);
                    // shifted copies are kept in registers until a job with another wrap code comes up
                    // (then the primaries are read again: only tiles at the box faces ever do)
                    const int jcode = jb.meta & 63;
                    if (jcode != cur_code) {
                        cur_code = jcode;
                        const int cx = jcode & 3, cy = (jcode >> 2) & 3, cz = (jcode >> 4) & 3;
                        const T ox = cx == 1 ? (T)P.wrap[0] : -(T)P.wrap[0];
                        const T oy = cy == 1 ? (T)P.wrap[1] : -(T)P.wrap[1];
                        const T oz = cz == 1 ? (T)P.wrap[2] : -(T)P.wrap[2];
#pragma unroll
                        for (int p = 0; p < FAST_PRIM; p++) {
                            const int i = lane + 32 * p;
                            if (i < nv) {
                                const T xr = pxg[i], yr = pyg[i], zr = pzg[i];
                                xq[p] = cx ? xr + ox : xr;
                                yq[p] = cy ? yr + oy : yr;
                                zq[p] = cz ? zr + oz : zr;
                            }
                        }
                    }
                    const int jk = (jb.meta >> 6) & 127, jl = (jb.meta >> 13) & 127;
                    for (int l0 = 0; l0 < jl; l0 += FAST_LMAX) {
                        const int nlp = min(FAST_LMAX, jl - l0);
                        // the edge table is padded with +inf, and a measured top level (JOB_TOPM) may use
                        // its real edge: every unmasked value of the job lies below it
                        const T *Es = &S.edges[jk + l0];
                        int cnt[FAST_LMAX];
#pragma unroll
                        for (int l = 0; l < FAST_LMAX; l++) cnt[l] = 0;
                        if (MODE == CFB_WP && (jb.meta & JOB_ZCUT)) {
                            if (jb.meta & JOB_DIRZ)
                                chunk_dispatch<T, MODE, 2>(nlp, pa, sx, sy, sz, m4, xq, yq, zq, Es, pimax, true, cnt);
                            else
                                chunk_dispatch<T, MODE, 1>(nlp, pa, sx, sy, sz, m4, xq, yq, zq, Es, pimax, false, cnt);
                        } else
                        {
                            bool done = false;
                            if constexpr (sizeof(T) == 4 && MODE == CFB_DD && FAST_SPLIT) {
                                if (tsplit && nlp <= 3) {  // the three bodies that carry 95 % of the iterations
                                    if (nlp == 1) chunk_T<T, MODE, 1, 4, 0, true>(sx, sy, sz, m4, xq, yq, zq, Es, pimax, false, cnt);
                                    else if (nlp == 2) chunk_T<T, MODE, 2, 4, 0, true>(sx, sy, sz, m4, xq, yq, zq, Es, pimax, false, cnt);
                                    else chunk_T<T, MODE, 3, 4, 0, true>(sx, sy, sz, m4, xq, yq, zq, Es, pimax, false, cnt);
                                    done = true;
                                }
                            }
                            if (!done) chunk_dispatch<T, MODE, 0>(nlp, pa, sx, sy, sz, m4, xq, yq, zq, Es, pimax, false, cnt);
                        }
                        // ---- warp totals -> warp histogram: +C at the level's slot, -C one above ----
                        // a lane counts at most FAST_PRIM * FAST_CH = 512 pairs per level here and the warp 2^14:
                        // two levels share one 32-bit warp reduction
                        unsigned mine = 0;
#pragma unroll
                        for (int m = 0; m < FAST_LMAX / 2; m++) {
                            if (2 * m < nlp) {
                                const unsigned tot = __reduce_add_sync(0xffffffffu, (unsigned)cnt[2 * m] | ((unsigned)cnt[2 * m + 1] << 16));
                                if ((lane >> 1) == m) mine = (lane & 1) ? (tot >> 16) : (tot & 0xffffu);
                            }
                        }
                        const int k = jk + l0 + lane;
                        const bool top = (jb.meta & JOB_TOPM) && (l0 + lane == jl - 1);
                        int C = (int)mine;
                        if (lane < nlp) {
                            if (jb.meta & JOB_TRI)  // full square of the tile against itself: drop self pairs, halve
                                C = (C - (v_self < S.edges_d[k] ? nv : 0)) / 2;
                            W.wh[k] += C;  // distinct slots across the lanes
                        }
                        __syncwarp();
                        if (lane < nlp && !top) W.wh[k + 1] -= C;
                        __syncwarp();
                        wbound += (unsigned)(nv * m4);
                    }
                    if (wbound >= (1u << 30)) {
                        flush_warp_hist(W.wh, S.hist, nedges, lane);
                        wbound = 0;
                    }
                    __syncwarp();  // everyone is done with buf[bsel] before it is staged again
                    it++;
                    e = e2;
                    c0 = c2;
                }
                qn = 0;
            }
        }
    }
    // ---------------- merge ----------------
    flush_warp_hist(W.wh, S.hist, nedges, lane);
    __syncthreads();
    for (int i = 1 + tid; i < nedges; i += blockDim.x) {
        const u64 v = S.hist[i];
        if (v) atomicAdd(&P.npairs[i], v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        my_eval += __shfl_xor_sync(0xffffffffu, my_eval, o);
        my_jobs += __shfl_xor_sync(0xffffffffu, my_jobs, o);
        my_analytic += __shfl_xor_sync(0xffffffffu, my_analytic, o);
        my_levels += __shfl_xor_sync(0xffffffffu, my_levels, o);
    }
    if (lane == 0) {
        if (my_eval) atomicAdd(&P.counters[0], my_eval);
        if (my_jobs) atomicAdd(&P.counters[1], my_jobs / 32);  // my_jobs is warp-uniform: the shuffle sum counted it 32x
        if (my_analytic) atomicAdd(&P.counters[2], my_analytic);
        if (my_levels) atomicAdd(&P.counters[3], my_levels);
    }
}

// ------------------------------------------------------------------------------------------------
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

template <typename T, int MODE, bool LIST>
static int launch_fast(const PairParams &P, const ParticleSet &SA, const ParticleSet &SB, cudaStream_t st)
{
    static int use_tma = -1;  // CORRFUNC_B200_STAGE=tma selects the bulk-copy staging (default: cp.async)
    if (use_tma < 0) {
        const char *e = getenv("CORRFUNC_B200_STAGE");
        use_tma = (e && strcmp(e, "tma") == 0) ? 1 : 0;
    }
    const PairParams &Q = P;
    if (Q.ntiles <= 0) return 0;
    // persistent warps: as many blocks as the device keeps resident, never more than there are tiles
    static int resident[2][2] = {{0, 0}, {0, 0}};
    int &res = resident[sizeof(T) == 8][use_tma];
    if (res == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        CK(cudaGetDevice(&dev));
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (use_tma)
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pairs_fast<T, MODE, LIST, true>, FAST_WARPS * 32, 0));
        else
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pairs_fast<T, MODE, LIST, false>, FAST_WARPS * 32, 0));
        const char *cap = getenv("CORRFUNC_B200_FAST_BLOCKS_PER_SM");  // tuning knob: fewer resident warps
        if (cap && atoi(cap) > 0 && atoi(cap) < per_sm) per_sm = atoi(cap);
        res = sms * (per_sm > 0 ? per_sm : 1);
    }
    int64_t nblk = (Q.ntiles / (Q.shard_n > 1 ? Q.shard_n : 1) + FAST_WARPS - 1) / FAST_WARPS + 1;
    if (nblk > res) nblk = res;
    // the last tail_k tiles of every resident warp (of every rank) go out in pieces (see the kernel's work units)
    PairParams Qt = P;
    {
        static int tail_k = -1;
        if (tail_k < 0) {
            const char *e = getenv("CORRFUNC_B200_TAIL_TILES_PER_WARP");
            tail_k = (e && *e) ? atoi(e) : 2;
        }
        const int nrow = 2 * Qt.g.reach[0] + 1;
        Qt.tail_parts = tail_k <= 0 ? 1 : (LIST ? 8 : nrow);
        const int64_t tail_tiles = (int64_t)tail_k * res * FAST_WARPS * (Qt.shard_n > 1 ? Qt.shard_n : 1);
        Qt.tail_first = Qt.tail_parts > 1 ? (Qt.ntiles > tail_tiles ? Qt.ntiles - tail_tiles : 0) : Qt.ntiles;
    }
    {
        // blocks for the units, not the tiles: a small input whose tiles all go out row by row can use more warps
        const int64_t units = Qt.tail_first + (Qt.ntiles - Qt.tail_first) * Qt.tail_parts;
        nblk = (units / (Qt.shard_n > 1 ? Qt.shard_n : 1) + FAST_WARPS - 1) / FAST_WARPS + 1;
        if (nblk > res) nblk = res;
    }
    const PairParams &Q2 = Qt;
    if (use_tma)
        k_pairs_fast<T, MODE, LIST, true><<<(unsigned int)nblk, FAST_WARPS * 32, 0, st>>>(Q2, view_of<T>(SA), view_of<T>(SB));
    else
        k_pairs_fast<T, MODE, LIST, false><<<(unsigned int)nblk, FAST_WARPS * 32, 0, st>>>(Q2, view_of<T>(SA), view_of<T>(SB));
    cfb_ctx().launches++;
    CK(cudaGetLastError());
    return 0;
}

template <typename T>
static int launch_fast_T(const cfb_binning *bin, const PairParams &P, bool list_mode)
{
    Ctx &c = cfb_ctx();
    const ParticleSet &SA = c.set[0];
    const ParticleSet &SB = bin->autocorr ? c.set[0] : c.set[1];
    if (list_mode) return launch_fast<T, CFB_THETA, true>(P, SA, SB, c.stream);
    switch (bin->mode) {
    case CFB_DD:
    case CFB_XI: return launch_fast<T, CFB_DD, false>(P, SA, SB, c.stream);
    case CFB_WP: return launch_fast<T, CFB_WP, false>(P, SA, SB, c.stream);
    default: return cfb_fail("fast kernel: unsupported mode %d", bin->mode);
    }
}

int cfb_launch_pairs_fast(const cfb_binning *bin, const PairParams &P, int prec, bool list_mode)
{
    static_assert(FAST_PRIM * 32 == CFB_TILE, "a warp owns one tile");
    return prec == 4 ? launch_fast_T<float>(bin, P, list_mode) : launch_fast_T<double>(bin, P, list_mode);
}
