// pairs_sum.cu -- the per-pair-sum kernel: every statistic whose result needs something PER ACCEPTED PAIR -- a 2-D bin
// (DDrppi, DDsmu and their survey-geometry twins), an average separation (ravg / thetaavg) or a PAIR_PRODUCT weight.
//
// Replaces the per-cell-pair CPU kernels (theory/DDrppi/countpairs_rp_pi_kernels.c.src:24-285,
// theory/DDsmu/countpairs_s_mu_kernels.c.src:24-309, theory/DD/countpairs_kernels.c.src:25-279 with need_rpavg / weights,
// theory/wp/wp_kernels.c.src:23-305, mocks/DDtheta_mocks/countpairs_theta_mocks_kernels.c.src:872-1141,
// mocks/DDrppi_mocks/countpairs_rp_pi_mocks_kernels.c.src:200-300, mocks/DDsmu_mocks/countpairs_s_mu_mocks_kernels.c.src:196-290)
// and the cell-pair enumeration of generate_cell_pairs_DOUBLE (utils/gridlink_impl.c.src:439-625).
//
// Design:
//   * persistent warps; a warp owns one primary tile (up to 64 particles of one fine cell, 2 per lane in registers) and
//     pulls its next tile from a global counter;
//   * phase 1 (one candidate neighbour cell per lane): periodic index and wrap code, the reference's role filter, and --
//     from the two cells' particle bounding boxes, in double with explicit rounding margins -- a conservative interval of
//     the binned quantity.  It prunes the cell pair, and it leaves the few bin edges that can split its pairs: the bin
//     search of an accepted pair is a count over those "levels" (1-3 compares) instead of a search over all edges;
//   * phase 2: the neighbour's x|y|z|(w) runs are staged per warp in shared memory, double buffered, with 16-byte
//     cp.async.  The HOT LOOP computes the separation of every (primary, secondary) with the reference's arithmetic
//     (same subtraction order, same FMA association per statistic) and only the range test; accepted pairs are
//     WARP-COMPACTED (ballot + prefix popcount) into a per-warp ring in shared memory;
//   * the DRAIN takes 32 accepted pairs at a time with all lanes busy: level count -> separation bin, the 2-D bin index
//     evaluated in floating point exactly as the reference does (IEEE divide and square roots), sqrt for the average,
//     the weight product, and the histogram update.  Only 0.3 drains run per 32 separations, instead of a divergent
//     accept branch in nearly every iteration;
//   * histograms: per block in shared memory, 32-bit words updated with native ATOMS.ADD (a count word; sums as 64-bit
//     fixed point in two words with the carry taken from the returned old value: exact, order independent,
//     reproducible), laid out [slot][copy][word] and REPLICATED `copies` times with the copy chosen by the lane: with 8
//     copies lanes of different copy index never meet in a bank, and lanes of one drain that hit the same slot mostly do
//     not serialise (ncu on config 3: the conflicts of these atomics were half of all shared-memory wavefronts).
//
// Compiled with -fmad=false: an FMA appears exactly where the reference's AVX-512 kernels have one.
#include <math_constants.h>
#include <stdlib.h>

#include "cfb_internal.cuh"

#ifndef SUM_WARPS
#define SUM_WARPS 4
#endif
#define SUM_PA 2                // primaries per lane
#define SUM_TILE (32 * SUM_PA)  // primaries per tile; divides CFB_TILE
#ifndef SUM_CH
#define SUM_CH 32               // secondaries per staged chunk (64 measured 10 % slower on config 3: one resident block fewer)
#endif
#ifndef SUM_QCAP
#define SUM_QCAP 80             // job queue entries per warp
#endif
#ifndef SUM_MINB
#define SUM_MINB 5               // resident blocks per SM the kernel is compiled for (register budget): the kernel is latency bound, measured 25 % faster at 5 blocks than at 4
#endif
#ifndef SUM_QL
#define SUM_QL 32               // accepted pairs a lane can queue before the warp compacts and drains (16: 3 % slower on config 3)
#endif
#define SUM_MAX_EDGES 256
#define SUM_KSH 19              // separation-bin table: key = float bits >> 19 (exponent + 4 mantissa bits: 16 keys per octave)
#define SUM_MAX_KEYS 2048       // more keys than this (edges spanning > 128 octaves in the squared separation): binary search

typedef unsigned long long u64;

namespace {

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

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
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

// Shared-memory accumulation with the native 32-bit integer ATOMS.ADD only (64-bit, float and double shared atomics
// compile to compare-and-swap loops on sm_100a).  Measured with tools/ubench_atoms.cu: a full-warp ATOMS.ADD to spread
// addresses sustains 10-15 word updates per clock and SM when the result is not used and 5-10 as a carry chain; an ATOMS
// with 8 active lanes costs as much as one with 32 -- which is why the accepted pairs are compacted first -- and lanes
// that meet in one bank or one word serialise (ncu, config 3, one histogram copy: 6 wavefronts per ATOMS).
//   count: one word, +1 per pair, result unused (ATOMS.POPC.INC: lanes on one address are merged by the hardware);
//   sums : fixed point -- value * 2^k rounded to an integer q below 2^46 in magnitude (k from the bin's upper edge, or
//          from the largest weight product: 1.4e-14 of the bin's largest value per term, sums compared at 1e-10) -- held
//          in two words as a 64-bit two's-complement number: ATOMS.ADD of the low 32 bits returns the old value, which
//          gives the carry that is added with the high part.  q is formed by ONE fused multiply-add with the constant
//          1.5 * 2^52: the low 52 bits of the result are q in two's complement (no F2I.S64, no shifts).
// The low word wraps by design (every wrap is a carry).  The count word and the high word (|q >> 32| <= 2^14, + carry)
// must not wrap: every warp NORMALISES the block's histogram every `norm_drains` of its drains -- these words are
// exchanged with zero and what they held goes to the global histogram.  Between two normalisations one word takes at most
// warps x norm_drains x (32 / copies) adds; the launch chooses norm_drains so that this stays below 2^16, i.e. below
// 2^31 in magnitude.  Integer sums are exact and order independent.
#define SUM_NORM_DRAINS 512
#define SUM_MAGIC 6755399441055744.0  // 1.5 * 2^52
#define SUM_QBITS 46
// shared-window addresses (32 bit) and explicit .shared atomics: one address register, immediate word offsets
__device__ __forceinline__ void sh_red_add(const unsigned addr, const unsigned v)
{
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned sh_atom_add(const unsigned addr, const unsigned v)
{
    unsigned old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void add2(const unsigned addr, const double m)  // m = fma(value, scale, SUM_MAGIC)
{
    const unsigned lo = (unsigned)__double2loint(m);
    const int hi = __double2hiint(m) - 0x43380000;  // floor(q / 2^32)
    const unsigned old = sh_atom_add(addr, lo);
    sh_red_add(addr + 4, (unsigned)hi + (old > ~lo ? 1u : 0u));
}
// second half of add2: the high word + the carry of the low word's add (old = what the low word held before it)
__device__ __forceinline__ void add2_hi(const unsigned addr_hi, const double m, const unsigned old)
{
    const unsigned lo = (unsigned)__double2loint(m);
    const int hi = __double2hiint(m) - 0x43380000;  // floor(q / 2^32)
    sh_red_add(addr_hi, (unsigned)hi + (old > ~lo ? 1u : 0u));
}
// what the high word of one sum holds (exchanged with zero); last: also the low word (nobody adds any more)
__device__ __forceinline__ double take2(unsigned *w, const bool last)
{
    double v = 0.0;
    if (w[1]) v = (double)(int)atomicExch(w + 1, 0u) * 4294967296.0;
    if (last) v += (double)w[0];
    return v;
}
// 2^k such that |v| * 2^k < 2^SUM_QBITS for every |v| <= vmax
__device__ __forceinline__ double fixed_scale(const double vmax)
{
    int e = 0;
    if (vmax > 0.0 && vmax < 1.0e300) frexp(vmax, &e);  // vmax < 2^e
    return ldexp(1.0, SUM_QBITS - e);
}

struct SumJob {
    int start;  // first secondary (index into the sorted arrays, multiple of CFB_PAD)
    int n;      // secondaries
    int meta;   // bits 0-5 wrap code | 6-13 first level edge | 14-22 number of levels | flags
};
#define SJ_TRI (1 << 23)     // the tile against itself: only secondaries after the primary
#define SJ_DIRZ (1 << 24)    // primary and secondary lie in different reference cells (one-sided pi cut of wp / DDrppi)
#define SJ_NEEDLO (1 << 25)  // some pair of the two cells may lie below the first edge

// Per-warp shared memory.  NA = staged arrays (x, y, z and, with weights, w).
template <typename T, int NA>
struct SumWarp {
    alignas(16) T buf[2][NA][SUM_CH];
    // double runs: float copies of the current chunk's positions -- the hot loop is a conservative FLOAT filter there
    alignas(16) float fbuf[sizeof(T) == 8 ? 3 : 1][sizeof(T) == 8 ? SUM_CH : 4];
    // the tile's primaries with the current wrap applied (+ weights), gathered by the drain.  Primary `lane + 32 p` lives at
    // [2 lane + p]: a batch of the drain holds the two primaries of two or three neighbouring lanes, and at [lane + 32 p]
    // the two rows of one lane shared their banks (ncu: 3.6 wavefronts per gather against 1.9 ideal)
    alignas(16) T prim[NA][SUM_TILE];
    // accepted pairs, first per LANE (one byte each: secondary | primary row << 6; entry i of lane l at [i][l], so a push is
    // one predicated byte store and one add, no ballot, no prefix count), then -- when a lane's queue is full or the chunk
    // ends -- compacted into `ring` (lane << 8 | entry) by one warp scan over the queue lengths: the drain takes 32
    // consecutive ring entries, all lanes busy whatever the spread of the queue lengths was.
    unsigned char laneq[SUM_QL][32];
    unsigned short ring[32 * SUM_QL + 64];  // + 64: the last drain reads full rows (idle lanes decode to valid indices)
    SumJob q[SUM_QCAP];
};

// everything the drain needs that does not change within a kernel
template <typename T>
struct SumConst {
    const T *edges;        // shared: T[nedges]
    const double *scale;   // shared: double[nedges], fixed-point scale of the separation sum per bin
    unsigned *hist;        // shared histogram [slot][copy][word]: word 0 count, then (low, high) of the separation sum and
                           // of the weight sum; addressed through hword()
    unsigned hl;           // shared-window address of this lane's copy of slot 0
    const unsigned short *ktab;  // separation-bin table (see drain), nkeys entries: lowest | highest possible bin << 8
    int kmin, nkeys;       // first key, number of keys (0: no table)
    unsigned hstride_b;    // bytes from one slot to the next
    int copies_shift;      // log2(copies)
    int cl;                // this lane's copy
    int nedges;
    int nmu, nmu_p1_i;     // DDsmu: number of mu bins, + 1
    int norm_drains;
    bool fast_ok;          // the edges lie within the float range: the float estimates of the drain may be used
    T pimax, sqr_pimax, inv_dpi, inv_dmu, sqr_mumax, npi_p1, nmu_p1, e_lo, e_hi, sqr_max_sep;
    float inv_dmu_f, mu_tol_b;
    double w_scale;
    int fast_acos;
};
// first word of (slot, this lane's copy); NP words per (slot, copy).  NP is odd (1, 3 or 5), so the multiplication permutes
// the banks: two lanes meet in a bank only if (slot * copies + copy) agrees modulo 32 -- never for different copies of 8.
template <int NP, typename T>
__device__ __forceinline__ unsigned hword(const SumConst<T> &K, const int slot)
{
    return K.hl + (unsigned)slot * K.hstride_b;
}

// true when value v lies at or above edge e in the binning order (theta: edges are cosines, decreasing)
template <typename T, int MODE>
__device__ __forceinline__ bool at_or_above(const T v, const T e)
{
    return MODE == CFB_THETA ? (v <= e) : (v >= e);
}

// ------------------------------------------------------------------------------------------------
// One separation and its range test.  b = second payload value (|dz| for DDrppi, dz^2 for DDsmu).
template <typename T, int MODE, bool NEEDLO>
__device__ __forceinline__ bool eval_pair(const T x2, const T y2, const T z2, const T x1, const T y1, const T z1, const T tz,
                                          const SumConst<T> &K, const bool dirz, T &v, T &b)
{
    // second - (first + wrap) (countpairs_kernels.c.src:79-81,178-180)
    const T dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
    bool acc;
    b = 0;
    if (MODE == CFB_DD || MODE == CFB_XI) {
        v = fma_t<T>(dz, dz, fma_t<T>(dy, dy, dx * dx));  // countpairs_kernels.c.src:188-190
        acc = v < K.e_hi;
        if (NEEDLO) acc = acc && v >= K.e_lo;
    } else if (MODE == CFB_WP) {
        v = fma_t<T>(dy, dy, dx * dx);  // wp_kernels.c.src:197-198
        // two reference cells: survivors of the fast-forward over z1 <= zpos - pimax, then the SIGNED dz < pimax
        // (wp_kernels.c.src:139-142, 207-221); inside one reference cell j follows i in z order (dz >= 0)
        acc = dirz ? (z2 > tz && dz < K.pimax) : (dz > -K.pimax && dz < K.pimax);
        acc = acc && v < K.e_hi;
        if (NEEDLO) acc = acc && v >= K.e_lo;
    } else if (MODE == CFB_RPPI) {
        v = fma_t<T>(dy, dy, dx * dx);  // countpairs_rp_pi_kernels.c.src:196-197
        b = dz < 0 ? -dz : dz;
        // survivors of the fast-forward (:139-142), then |dz| < pimax (:196-207)
        acc = (dirz ? (z2 > tz) : (dz > -K.pimax)) && b < K.pimax;
        acc = acc && v < K.e_hi;
        if (NEEDLO) acc = acc && v >= K.e_lo;
    } else if (MODE == CFB_SMU) {
        b = dz * dz;
        v = fma_t<T>(dx, dx, fma_t<T>(dy, dy, b));  // countpairs_s_mu_kernels.c.src:212-214
        acc = v < K.e_hi;
        if (NEEDLO) acc = acc && v >= K.e_lo;
    } else if (MODE == CFB_THETA) {
        // cos(theta) = 1 - chord^2 / 2 (countpairs_theta_mocks_kernels.c.src:1062-1068); edges decrease
        const T chord2 = fma_t<T>(dz, dz, fma_t<T>(dy, dy, dx * dx));
        v = (T)1.0 - (T)0.5 * chord2;
        acc = v > K.e_hi;
        if (NEEDLO) acc = acc && v <= K.e_lo;
    } else if (MODE == CFB_RPPI_MOCKS) {
        v = fma_t<T>(dx, dx, fma_t<T>(dy, dy, dz * dz));  // sqr_sep; the rest of the pair runs in the drain
        acc = v < K.sqr_max_sep;
    } else {  // CFB_SMU_MOCKS
        v = fma_t<T>(dx, dx, fma_t<T>(dy, dy, dz * dz));
        acc = v < K.e_hi && v >= K.e_lo;
    }
    return acc;
}

__device__ __forceinline__ float rsqrt_apx(const float v)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v));
    return y;
}
__device__ __forceinline__ float sqrt_apx(const float v)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v));
    return y;
}
// sqrt for the average separation, rs ~ 1/sqrt(v) in float (relative error < 2^-21).  double: one Newton step in double,
// relative error < 6e-14 (the averages are sums of millions of terms compared at 1e-10; the reference's own sums depend on
// the thread schedule at the 1e-13 level).  float: IEEE.
__device__ __forceinline__ float sep_sqrt(const float v, const float) { return __fsqrt_rn(v); }
__device__ __forceinline__ double sep_sqrt(const double v, const float rs)
{
    const double y = (double)rs;
    const double r0 = v * y;
    const double e = __fma_rn(-r0, r0, v);
    return __fma_rn(e, 0.5 * y, r0);
}

// Separation bin of a value with MANY candidate edges (a pair of close cells spans most of a logarithmic binning, and
// those jobs hold a large share of all accepted pairs; ncu: the binary search was a quarter of the drain's instructions).
// The float image of the value, cut to its exponent and four mantissa bits, indexes a small table that holds the lowest
// and the highest bin a value with that key can have (built with one float ulp of slack on both sides, so the rounding of
// the conversion cannot leave the range): for any reasonable binning the two differ by at most one, and ONE exact compare
// against the edge between them decides.  The bin is the number of edges at or below the value, exactly as the level
// count / binary search give it.
// ------------------------------------------------------------------------------------------------
// The drain: n (<= 32 * SUM_NB) queued pairs, ring entries head .. head + n - 1, SUM_NB per lane.  Every pair is
// evaluated again with the statistic's own arithmetic and its exact range test (in double runs the hot loop was only a
// float filter).  Separation bin = the number of edges at or below the value: two compares against the job's level
// edges when the job has at most two levels; otherwise a FLOAT-KEY TABLE (a pair of close cells spans most of a
// logarithmic binning, and those jobs hold a large share of all accepted pairs -- ncu: the binary search was a quarter
// of the drain's instructions).  The float image of the value, cut to its exponent and four mantissa bits, indexes a
// table of the lowest and the highest bin a value with that key can have (built with one float ulp of slack on both
// sides, so the rounding of the conversion cannot leave the range): for any reasonable binning the two differ by at most
// one, and ONE exact compare against the edge between them decides (else a binary search between the two).
//
// DDsmu: the reference evaluates slot = (int)(sbin * (nmu + 1) + sqrt(dz^2 / s^2) * inv_dmu) with a true divide and an
// IEEE square root (countpairs_s_mu_kernels.c.src:228-277).  Here the mu bin is first located in FLOAT arithmetic with
// the approximate MUFU reciprocal square root and square root (the FP32 and XU pipes are idle in this kernel): if the
// estimate lies farther from an integer than its error bound -- 4e-6 (y + 1) covers the conversions, the two
// approximations and three multiplications; in float the rounding of the reference's floating-point index sum, half an
// ulp of the slot count, is added -- its integer part IS the reference's mu bin (or, at nmu and above, the pair fails
// the reference's dz^2 < s^2 mu_max^2).  Otherwise (about 1e-4 of the batches) the whole batch takes the reference's
// arithmetic.
#ifndef SUM_NB
#define SUM_NB 2  // batches of 32 pairs a drain call works on at once: the per-pair code is one long dependent chain
                  // (gather -> 6 DP -> conversions -> MUFU -> table -> compare -> Newton step -> atomics), two of them in
                  // flight per lane hide each other's latencies
#endif
template <typename T, int MODE, bool AVG, bool WGT, int NA>
__device__ __forceinline__ void drain(const int n, const int head, SumWarp<T, NA> &W, const int bsel, const SumConst<T> &K,
                                      const bool dirz, const T (&E)[4], const int kfirst, const int nl, const int lane)
{
    constexpr int NP = 1 + (AVG ? 2 : 0) + (WGT ? 2 : 0);
    constexpr int NB = SUM_NB;
    bool ok[NB];
    int pidx[NB], j[NB];
    T a[NB], b[NB], v[NB], f2[NB];
#pragma unroll
    for (int u = 0; u < NB; u++) {
        ok[u] = 32 * u + lane < n;
        const unsigned idx = W.ring[head + 32 * u + lane];
        pidx[u] = ((idx >> 7) & 62) | ((idx >> 6) & 1);  // 2 lane + p
        j[u] = idx & (SUM_CH - 1) & 63;
    }
#pragma unroll
    for (int u = 0; u < NB; u++) {
        const T x1 = W.prim[0][pidx[u]], y1 = W.prim[1][pidx[u]], z1 = W.prim[2][pidx[u]];
        const T x2 = W.buf[bsel][0][j[u]], y2 = W.buf[bsel][1][j[u]], z2 = W.buf[bsel][2][j[u]];
        // float runs: the hot loop applied this very test; double runs: it applied a float filter that passes a superset
        if (!eval_pair<T, MODE, true>(x2, y2, z2, x1, y1, z1, z1 - K.pimax, K, dirz, a[u], b[u])) ok[u] = false;
        v[u] = a[u];  // the quantity that is binned
        f2[u] = 0;    // second-dimension coordinate in bins (pi * inv_dpi or mu * inv_dmu), reference arithmetic
        if (MODE == CFB_RPPI) {
            f2[u] = b[u] * K.inv_dpi;
        } else if (MODE == CFB_RPPI_MOCKS || MODE == CFB_SMU_MOCKS) {
            // line of sight = pair midpoint (countpairs_rp_pi_mocks_kernels.c.src:200-290, countpairs_s_mu_mocks_kernels.c.src:196-290)
            const T dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
            const T parx = x2 + x1, pary = y2 + y1, parz = z2 + z1;
            const T term1 = parx * dx, term2 = pary * dy;
            const T s_dot_l = fma_t<T>(parz, dz, term1 + term2);
            const T sqr_s_dot_l = s_dot_l * s_dot_l;
            const T sqr_norm_l = fma_t<T>(parx, parx, fma_t<T>(pary, pary, parz * parz));
            if (MODE == CFB_RPPI_MOCKS) {
                // a = sqr_sep, already below sqr_max_sep
                if (!(sqr_s_dot_l < K.sqr_pimax * sqr_norm_l)) ok[u] = false;
                const T sqr_Dpar = divi_t<T>(sqr_s_dot_l, sqr_norm_l);
                const T sqr_Dperp = a[u] - sqr_Dpar;
                if (!(sqr_Dpar < K.sqr_pimax && sqr_Dperp < K.e_hi && sqr_Dperp >= K.e_lo)) ok[u] = false;
                v[u] = sqr_Dperp;
                f2[u] = sqrt_t<T>(sqr_Dpar) * K.inv_dpi;
            } else {
                // a = s2, already within [e_lo, e_hi)
                const T sqr_mu = divi_t<T>(sqr_s_dot_l, sqr_norm_l * a[u]);
                if (!(sqr_mu < K.sqr_mumax)) ok[u] = false;
                f2[u] = sqrt_t<T>(sqr_mu) * K.inv_dmu;
            }
        }
    }
    // separation bin: kfirst + the number of levels at or below the value (E beyond the job's levels: never reached)
    int kb[NB];
    float vf[NB];
    bool have_vf = false;
    if (nl <= 2) {
#pragma unroll
        for (int u = 0; u < NB; u++) {
            kb[u] = kfirst + (at_or_above<T, MODE>(v[u], E[0]) ? 1 : 0) + (at_or_above<T, MODE>(v[u], E[1]) ? 1 : 0);
            vf[u] = 0.0f;
        }
    } else if (MODE != CFB_THETA && K.nkeys > 0) {
        // float-key table (see drain): lowest and highest possible bin, one exact compare decides between neighbours
        have_vf = true;
        int lo[NB], hi[NB];
        bool wide = false;
#pragma unroll
        for (int u = 0; u < NB; u++) {
            vf[u] = (float)v[u];
            int key = (int)(__float_as_uint(vf[u]) >> SUM_KSH) - K.kmin;
            key = max(0, min(key, K.nkeys - 1));
            const unsigned lh = K.ktab[key];
            lo[u] = lh & 255, hi[u] = lh >> 8;
            wide = wide || hi[u] > lo[u] + 1;
        }
        if (__any_sync(0xffffffffu, wide)) {
#pragma unroll
            for (int u = 0; u < NB; u++) {
                int l = lo[u], h = hi[u];
                while (l < h) {  // first edge in [l, h) the value is not at or above
                    const int mid = (l + h) >> 1;
                    if (v[u] >= K.edges[mid]) l = mid + 1; else h = mid;
                }
                kb[u] = l;
            }
        } else {
#pragma unroll
            for (int u = 0; u < NB; u++) kb[u] = lo[u] + ((hi[u] > lo[u] && v[u] >= K.edges[lo[u]]) ? 1 : 0);
        }
    } else {
#pragma unroll
        for (int u = 0; u < NB; u++) {
            vf[u] = 0.0f;
            if (nl <= 4) {
                kb[u] = kfirst;
#pragma unroll
                for (int l = 0; l < 4; l++) kb[u] += at_or_above<T, MODE>(v[u], E[l]) ? 1 : 0;
            } else {
                int lo = kfirst, hi = kfirst + nl;  // first edge in [lo, hi) the value is not at or above
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (at_or_above<T, MODE>(v[u], K.edges[mid])) lo = mid + 1; else hi = mid;
                }
                kb[u] = lo;
            }
        }
    }
    int slot[NB];
    float rs[NB];        // ~ 1 / sqrt(v) in float, when already known
    bool have_rs = false;
#pragma unroll
    for (int u = 0; u < NB; u++) slot[u] = kb[u], rs[u] = 0.0f;
    if (MODE == CFB_RPPI || MODE == CFB_RPPI_MOCKS) {
        // rpbin*(npibin+1) + pi*inv_dpi evaluated in T, then truncated (countpairs_rp_pi_kernels.c.src:249-256)
#pragma unroll
        for (int u = 0; u < NB; u++) slot[u] = (int)((T)kb[u] * K.npi_p1 + f2[u]);
    } else if (MODE == CFB_SMU_MOCKS) {
#pragma unroll
        for (int u = 0; u < NB; u++) slot[u] = (int)((T)kb[u] * K.nmu_p1 + f2[u]);  // countpairs_s_mu_mocks_kernels.c.src
    } else if (MODE == CFB_SMU) {
        bool amb = false;
        int mb[NB];
#pragma unroll
        for (int u = 0; u < NB; u++) {
            mb[u] = 0;
            if (K.fast_ok) {
                const float af = have_vf ? vf[u] : (float)a[u], bf = (float)b[u];  // DDsmu bins a itself
                rs[u] = rsqrt_apx(af);
                const float y = sqrt_apx(bf) * rs[u] * K.inv_dmu_f;
                mb[u] = __float2int_rz(y);
                const float fr = y - (float)mb[u];
                const float tol = __fmaf_rn(4.0e-6f, y, K.mu_tol_b);  // mu_tol_b includes the 4e-6 * 1
                amb = amb || (ok[u] && !(fr > tol && fr < 1.0f - tol));
            } else
                amb = amb || ok[u];
        }
        have_rs = K.fast_ok;
        if (__any_sync(0xffffffffu, amb)) {
            // keep if dz^2 < s^2 mu_max^2; sqr_mu = dz^2 / s^2 (true divide, fast_divide_and_NR_steps == 0), mu = sqrt;
            // slot = (int)(sbin * (nmu + 1) + mu * inv_dmu) in T (countpairs_s_mu_kernels.c.src:216-277)
#pragma unroll
            for (int u = 0; u < NB; u++) {
                if (!(b[u] < a[u] * K.sqr_mumax)) ok[u] = false;
                const T f = sqrt_t<T>(divi_t<T>(b[u], a[u])) * K.inv_dmu;
                slot[u] = (int)((T)kb[u] * K.nmu_p1 + f);
            }
        } else {
#pragma unroll
            for (int u = 0; u < NB; u++) {
                if (mb[u] >= K.nmu) ok[u] = false;
                slot[u] = kb[u] * K.nmu_p1_i + mb[u];
            }
        }
    }
    // the sums' fixed-point images first, for all batches; then every low-word atomic of the call (their old values come
    // back independently -- issued one after the other's carry, the four round trips of a call were 7 % of all warp
    // samples), then the high words with the carries
    double msep[NB], mw[NB];
#pragma unroll
    for (int u = 0; u < NB; u++) {
        if (!ok[u]) slot[u] = 0, kb[u] = kfirst;  // idle lanes: any valid address
        msep[u] = 0.0, mw[u] = 0.0;
        if (AVG) {
            T sep;
            if (MODE == CFB_THETA) {
                const T cc = v[u] >= (T)1.0 ? (T)1.0 : v[u];
                const T th = K.fast_acos ? fast_acos_t<T>(cc) : (T)acos(cc);
                sep = (T)(th * (T)57.29577951308232087679815481410517);
            } else if (sizeof(T) == 8 && !K.fast_ok) {
                sep = sqrt_t<T>(v[u]);
            } else {
                if (sizeof(T) == 8 && !have_rs) rs[u] = rsqrt_apx(have_vf ? vf[u] : (float)v[u]);
                sep = sep_sqrt(v[u], rs[u]);
            }
            msep[u] = __fma_rn((double)sep, K.scale[kb[u]], SUM_MAGIC);
        }
        if (WGT) {
            const T w1 = W.prim[NA - 1][pidx[u]], w2 = W.buf[bsel][NA - 1][j[u]];
            mw[u] = __fma_rn((double)(T)(w1 * w2), K.w_scale, SUM_MAGIC);  // pair_product, weight_functions.h.src:71-91
        }
    }
    unsigned hw[NB], old_s[NB], old_w[NB];
#pragma unroll
    for (int u = 0; u < NB; u++) {
        hw[u] = hword<NP>(K, slot[u]);
        old_s[u] = 0u, old_w[u] = 0u;
#ifdef SUM_ABL_NOATOM  // ablation (timing experiments only, results wrong): no histogram update
        if (ok[u] && slot[u] == -12345) {
#else
        if (ok[u]) {
#endif
            sh_red_add(hw[u], 1u);
            if (AVG) old_s[u] = sh_atom_add(hw[u] + 4, (unsigned)__double2loint(msep[u]));
            if (WGT) old_w[u] = sh_atom_add(hw[u] + 4 + (AVG ? 8 : 0), (unsigned)__double2loint(mw[u]));
        }
    }
#pragma unroll
    for (int u = 0; u < NB; u++) {
#ifdef SUM_ABL_NOATOM
        if (ok[u] && slot[u] == -12345) {
#else
        if (ok[u]) {
#endif
            if (AVG) add2_hi(hw[u] + 8, msep[u], old_s[u]);
            if (WGT) add2_hi(hw[u] + 8 + (AVG ? 8 : 0), mw[u], old_w[u]);
        }
    }
}

// Moves what the wrapping-prone words of the block's histogram hold into the global histogram (any warp, any time).
// last: everything (end of the kernel, after a block barrier).
template <int MODE, bool AVG, bool WGT, typename T>
__device__ __forceinline__ void flush_hist(const PairParams &P, const SumConst<T> &K, const int ns, const int first,
                                           const int step, const bool last)
{
    constexpr int NP = 1 + (AVG ? 2 : 0) + (WGT ? 2 : 0);
    const int copies = 1 << K.copies_shift;
    for (int i = first; i < ns; i += step) {
        u64 cnt = 0;
        double ssep = 0.0, sw = 0.0;
        unsigned *h = K.hist + (unsigned)((i << K.copies_shift) * NP);
        for (int c = 0; c < copies; c++, h += NP) {
            if (h[0]) cnt += atomicExch(h, 0u);
            if (AVG) ssep += take2(h + 1, last);
            if (WGT) sw += take2(h + 1 + (AVG ? 2 : 0), last);
        }
        if (cnt) atomicAdd(&P.npairs[i], cnt);
        if (AVG && ssep != 0.0) {
            const int kb = (MODE == CFB_RPPI || MODE == CFB_RPPI_MOCKS)
                               ? i / (P.npibin + 1)
                               : ((MODE == CFB_SMU || MODE == CFB_SMU_MOCKS) ? i / (P.nmu_bins + 1) : i);
            atomicAdd(&P.sum_sep[i], ssep / K.scale[kb < K.nedges ? kb : K.nedges - 1]);
        }
        if (WGT && sw != 0.0) atomicAdd(&P.sum_w[i], sw / K.w_scale);
    }
}

// Float thresholds of the double runs' filter (see hot_loop): every pair the statistic accepts passes.
struct SumFilt {
    float hi, lo, pimax;
};
// bound on |v_float - v| for v = squared separation sqrt(e2)^2 computed in float from coordinates of magnitude <= cm:
// ed = 2^-21 cm covers the rounding of both coordinates and of the difference twice over; three more roundings in the sum
__device__ __forceinline__ double filt_margin(const double e2, const double ed)
{
    return 3.47 * ed * sqrt(fmax(e2, 0.0)) + 3.0 * ed * ed + 5.0e-7 * e2;
}
template <int MODE, bool NEEDLO>
__device__ __forceinline__ bool filter_pair(const float x2, const float y2, const float z2, const float x1, const float y1,
                                            const float z1, const SumFilt &F)
{
    const float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
    bool acc;
    if (MODE == CFB_WP || MODE == CFB_RPPI) {
        const float v = __fmaf_rn(dy, dy, dx * dx);
        acc = v < F.hi && fabsf(dz) < F.pimax;
        if (NEEDLO) acc = acc && v >= F.lo;
    } else {
        // DD / DDsmu: s^2; theta: chord^2 (1 - chord^2 / 2 > cos(theta_max)); mocks: s^2 (3-D)
        const float v = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
        acc = v < F.hi;
        if ((NEEDLO && MODE != CFB_RPPI_MOCKS) || MODE == CFB_SMU_MOCKS) acc = acc && v >= F.lo;
    }
    return acc;
}

// The hot loop: secondaries k .. m-1 of the staged chunk against the lane's PA primaries, until the chunk ends or some
// lane's queue may overflow in the next iteration.  It only decides WHICH pairs go on -- in float runs with the
// statistic's own arithmetic and range test (same subtraction order, same FMA association), in double runs with a
// conservative FLOAT filter (float copies of the positions, thresholds widened by a bound on the float error): the FP32
// pipe runs at twice the FP64 rate with half the latency, the loop needs neither the double primaries nor the level edges
// in registers, and the drain evaluates every queued pair exactly anyway.  Four independent separations per lane and
// iteration; a queued pair costs one predicated byte store and one add.  Returns the next secondary.
template <typename T, int MODE, int PA, bool TRI, bool NEEDLO>
__device__ __forceinline__ int hot_loop(int k, const int m, const float *sx, const float *sy, const float *sz,
                                        const float (&xh)[SUM_PA], const float (&yh)[SUM_PA], const float (&zh)[SUM_PA],
                                        const SumConst<T> &K, const SumFilt &F, const bool dirz, const int c0,
                                        const int lane, unsigned &qaddr, const unsigned qlimit)
{
    constexpr int NS = 4 / PA;  // secondaries per iteration
    float tz[PA];
#pragma unroll
    for (int p = 0; p < PA; p++) tz[p] = zh[p] - (float)K.pimax;  // target of the reference's fast-forward over z (wp, DDrppi)
    for (; k < m; k += NS) {
        float x2[NS], y2[NS], z2[NS];
        if constexpr (NS == 4) {
            const float4 X = *reinterpret_cast<const float4 *>(sx + k), Y = *reinterpret_cast<const float4 *>(sy + k),
                         Z = *reinterpret_cast<const float4 *>(sz + k);
            x2[0] = X.x, x2[1] = X.y, x2[2] = X.z, x2[3] = X.w;
            y2[0] = Y.x, y2[1] = Y.y, y2[2] = Y.z, y2[3] = Y.w;
            z2[0] = Z.x, z2[1] = Z.y, z2[2] = Z.z, z2[3] = Z.w;
        } else {
            const float2 X = *reinterpret_cast<const float2 *>(sx + k), Y = *reinterpret_cast<const float2 *>(sy + k),
                         Z = *reinterpret_cast<const float2 *>(sz + k);
            x2[0] = X.x, x2[1] = X.y, y2[0] = Y.x, y2[1] = Y.y, z2[0] = Z.x, z2[1] = Z.y;
        }
        bool acc[4];
#pragma unroll
        for (int s = 0; s < NS; s++) {
#pragma unroll
            for (int p = 0; p < PA; p++) {
                const int e = s * PA + p;
                if constexpr (sizeof(T) == 4) {
                    float v, b;
                    acc[e] = eval_pair<float, MODE, NEEDLO>(x2[s], y2[s], z2[s], xh[p], yh[p], zh[p], tz[p], K, dirz, v, b);
                } else
                    acc[e] = filter_pair<MODE, NEEDLO>(x2[s], y2[s], z2[s], xh[p], yh[p], zh[p], F);
                if (TRI) acc[e] = acc[e] && (c0 + k + s > lane + 32 * p);
            }
        }
#pragma unroll
        for (int s = 0; s < NS; s++) {
#pragma unroll
            for (int p = 0; p < PA; p++) {
                if (acc[s * PA + p]) {
                    // one address register (a shared-window address), one predicated store, one predicated add
                    asm volatile("st.shared.u8 [%0], %1;" ::"r"(qaddr), "r"((k + s) | (p << 6)) : "memory");
                    qaddr += 32;
                }
            }
        }
        if (__any_sync(0xffffffffu, qaddr > qlimit)) {
            k += NS;
            break;
        }
    }
    return k;
}

// The direct loop (count-only 1-D statistics and DDrppi, jobs with at most 8 levels): no queue at all -- every lane
// finishes its own pairs in the run's precision (level count, slot) and only the histogram update is predicated.
// Without per-pair sums that is fewer instructions than queueing and re-evaluating 0.3 accepted pairs per separation.
template <typename T, int MODE, int NA, int PA, bool TRI, bool NEEDLO, int DIRECT /* levels in registers */>
__device__ __forceinline__ void direct_loop(const int m, SumWarp<T, NA> &W, const int bsel, const T (&xq)[SUM_PA],
                                            const T (&yq)[SUM_PA], const T (&zq)[SUM_PA], const SumConst<T> &K,
                                            const bool dirz, const int c0, const int lane, const T (&E)[8],
                                            const int kfirst)
{
    const T *sx = W.buf[bsel][0], *sy = W.buf[bsel][1], *sz = W.buf[bsel][2];
    constexpr int NS = 4 / PA;  // secondaries per iteration
    T tz[PA];
#pragma unroll
    for (int p = 0; p < PA; p++) tz[p] = zq[p] - K.pimax;  // target of the reference's fast-forward over z (wp, DDrppi)
    for (int k = 0; k < m; k += NS) {
        T v[4], b[4];
        bool acc[4];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const T x2 = sx[k + s], y2 = sy[k + s], z2 = sz[k + s];
#pragma unroll
            for (int p = 0; p < PA; p++) {
                const int e = s * PA + p;
                acc[e] = eval_pair<T, MODE, NEEDLO>(x2, y2, z2, xq[p], yq[p], zq[p], tz[p], K, dirz, v[e], b[e]);
                if (TRI) acc[e] = acc[e] && (c0 + k + s > lane + 32 * p);
            }
        }
#pragma unroll
        for (int e = 0; e < 4; e++) {
            int kb = kfirst;
#pragma unroll
            for (int l = 0; l < DIRECT; l++) kb += at_or_above<T, MODE>(v[e], E[l]) ? 1 : 0;
            int slot = kb;
            // rpbin*(npibin+1) + |dz|*inv_dpi evaluated in T, then truncated (countpairs_rp_pi_kernels.c.src:249-256)
            if (MODE == CFB_RPPI) slot = (int)((T)kb * K.npi_p1 + b[e] * K.inv_dpi);
            if (acc[e]) sh_red_add(hword<1>(K, slot), 1u);
        }
    }
}

// Compacts the lanes' queues into the ring (one warp scan) and drains it, 32 pairs at a time.
template <typename T, int MODE, bool AVG, bool WGT, int NA>
__device__ __forceinline__ void flush_queues(SumWarp<T, NA> &W, const int bsel, const SumConst<T> &K, const PairParams &P,
                                             const int ns, const bool dirz, const int kfirst, const int nl, const int lane,
                                             const unsigned char *qbase, const unsigned qaddr0, unsigned &qaddr, int &drains)
{
    const int c = (int)(qaddr - qaddr0) >> 5;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    const int excl = incl - c;
    const int maxc = __reduce_max_sync(0xffffffffu, c);
    const unsigned lb = (unsigned)lane << 8;
    // fully unrolled over the queue depth, four entries per (warp-uniform) exit test: every load and store has an immediate
    // offset (the rolled loop spent 18 instructions per entry on index arithmetic, 6 % of the kernel's instructions)
    {
        unsigned short *const rdst = W.ring + excl;
#pragma unroll
        for (int i0 = 0; i0 < SUM_QL; i0 += 4) {
            if (i0 >= maxc) break;
            unsigned e[4];
#pragma unroll
            for (int u = 0; u < 4; u++) e[u] = qbase[32 * (i0 + u)];
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (i0 + u < c) rdst[i0 + u] = (unsigned short)(lb | e[u]);
        }
    }
    qaddr = qaddr0;
    __syncwarp();
    T E[4];  // the job's first level edges; beyond nl: an edge no value is at or above
#pragma unroll
    for (int l = 0; l < 4; l++) {
        const T none = MODE == CFB_THETA ? (sizeof(T) == 4 ? (T)-CUDART_INF_F : (T)-CUDART_INF)
                                         : (sizeof(T) == 4 ? (T)CUDART_INF_F : (T)CUDART_INF);
        E[l] = l < nl ? K.edges[min(kfirst + l, K.nedges - 1)] : none;
    }
#ifdef SUM_ABL_NODRAIN  // ablation (timing experiments only, results wrong): queue and compact, never drain
    if (total < 0)
#endif
    for (int head = 0; head < total; head += 32 * SUM_NB) {
        drain<T, MODE, AVG, WGT, NA>(min(32 * SUM_NB, total - head), head, W, bsel, K, dirz, E, kfirst, nl, lane);
    }
    drains += (total + 31) >> 5;  // at most SUM_QL per flush: the launch leaves that much slack in norm_drains
    if (drains >= K.norm_drains) {
        flush_hist<MODE, AVG, WGT, T>(P, K, ns, lane, 32, false);
        drains = 0;
    }
    __syncwarp();  // the ring and the queues are free again
}

template <typename T, int NA>
__device__ __forceinline__ void stage_chunk(SumWarp<T, NA> &W, const int bsel, const SetView<T> &B, const int first,
                                            const int m4, const int lane)
{
    constexpr int EPV = 16 / (int)sizeof(T);
    for (int v = lane * EPV; v < m4; v += 32 * EPV) {
        cp_async16(&W.buf[bsel][0][v], B.x + first + v);
        cp_async16(&W.buf[bsel][1][v], B.y + first + v);
        cp_async16(&W.buf[bsel][2][v], B.z + first + v);
        if (NA == 4) cp_async16(&W.buf[bsel][3][v], B.w + first + v);
    }
    cp_async_commit();
}

// ------------------------------------------------------------------------------------------------
template <typename T, int MODE, bool AVG, bool WGT, bool LIST>
__global__ void __launch_bounds__(SUM_WARPS * 32, SUM_MINB)
k_pairs_sum(const PairParams P, const SetView<T> A, const SetView<T> B)
{
    constexpr int NA = WGT ? 4 : 3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // dynamic smem: per-warp areas | edges (T) | edges as double, in search order | sep scale per edge | histograms
    SumWarp<T, NA> *warps = (SumWarp<T, NA> *)smem_raw;
    const int nedges = P.nedges;
    size_t off = (sizeof(SumWarp<T, NA>) * SUM_WARPS + 15) & ~(size_t)15;
    T *s_edges = (T *)(smem_raw + off);
    off += ((size_t)nedges * sizeof(T) + 15) & ~(size_t)15;
    double *s_edges_d = (double *)(smem_raw + off);
    off += (size_t)nedges * 8;
    double *s_scale = (double *)(smem_raw + off);
    off += (size_t)nedges * 8;
    unsigned short *s_ktab = (unsigned short *)(smem_raw + off);
    off += ((size_t)P.sum_nkeys * 2 + 15) & ~(size_t)15;
    unsigned *s_hist = (unsigned *)(smem_raw + off);
    const int ns = (int)P.nslots;
    const int cshift = P.sum_copies_shift;
    const int hwords = (ns << cshift) * (1 + (AVG ? 2 : 0) + (WGT ? 2 : 0));

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < nedges; i += blockDim.x) {
        const T e = ((const T *)P.edges)[i];
        s_edges[i] = e;
        s_edges_d[i] = MODE == CFB_THETA ? -(double)e : (double)e;
        // largest separation a pair of the bin below edge i can have: sqrt(edge) (edges are squared), or the angle in
        // degrees whose cosine the edge is
        double vmax;
        if (MODE == CFB_THETA) vmax = acos(fmin(1.0, fmax(-1.0, (double)e))) * 57.29577951308232087679815481410517 + 1e-9;
        else vmax = sqrt(fmax(0.0, (double)e));
        s_scale[i] = fixed_scale(vmax);
    }
    for (int i = tid; i < hwords; i += blockDim.x) s_hist[i] = 0u;
    SumWarp<T, NA> &W = warps[wid];

    SumConst<T> K;
    K.edges = s_edges;
    K.scale = s_scale;
    K.hist = s_hist;
    K.ktab = s_ktab;
    K.kmin = P.sum_kmin;
    K.nkeys = P.sum_nkeys;
    K.cl = lane & P.sum_cmask;
    K.hl = smem_u32(smem_raw) + P.sum_hist_off + (unsigned)K.cl * (4 * (1 + (AVG ? 2 : 0) + (WGT ? 2 : 0)));
    K.hstride_b = P.sum_hstride_b;
    K.norm_drains = P.sum_norm_drains;
    K.nmu = P.nmu_bins;
    K.nmu_p1_i = P.nmu_bins + 1;
    // float estimate of the mu bin (drain): 4e-6 for the estimate itself; in float also the rounding of the reference's
    // floating-point index sum sbin * (nmu + 1) + mu * inv_dmu (half an ulp at the slot count, taken twice over)
    K.mu_tol_b = 4.0e-6f + (sizeof(T) == 4 ? 2.5e-7f * (float)ns : 0.0f);
    K.copies_shift = cshift;
    K.nedges = nedges;
    K.pimax = (T)P.pimax;
    K.sqr_pimax = K.pimax * K.pimax;  // rppi mocks (countpairs_rp_pi_mocks_kernels.c.src:56-57)
    K.inv_dpi = (T)P.inv_dpi;
    K.inv_dmu = (T)P.inv_dmu;
    K.inv_dmu_f = P.sum_inv_dmu_f;
    K.sqr_mumax = (T)P.sqr_mumax;
    K.npi_p1 = (T)(P.npibin + 1);
    K.nmu_p1 = (T)(P.nmu_bins + 1);
    K.fast_acos = P.fast_acos;
    K.w_scale = 1.0;
    if (WGT) K.w_scale = fixed_scale(P.wmax[0] * P.wmax[1]);
    __syncthreads();
    if (MODE != CFB_THETA) {
        // separation-bin table (see drain): for every key the number of edges at or below the smallest / the largest value
        // whose float image can carry that key
        for (int k = tid; k < P.sum_nkeys; k += blockDim.x) {
            const double vlo = (double)__uint_as_float((unsigned)(k + P.sum_kmin) << SUM_KSH) * (1.0 - 2.4e-7);
            const double vhi = (double)__uint_as_float((unsigned)(k + P.sum_kmin + 1) << SUM_KSH) * (1.0 + 2.4e-7);
            int lo = 0, hi = 0;
            for (int i = 0; i < nedges; i++) {
                lo += s_edges_d[i] <= vlo ? 1 : 0;
                hi += s_edges_d[i] <= vhi ? 1 : 0;
            }
            if (hi > nedges - 1) hi = nedges - 1;  // keeps every index derived from a bin inside the tables
            if (lo > hi) lo = hi;
            s_ktab[k] = (unsigned short)(lo | (hi << 8));
        }
        __syncthreads();
    }
    K.e_lo = s_edges[0];
    K.e_hi = s_edges[nedges - 1];
    K.fast_ok = (double)K.e_lo >= 1.0e-30 && (double)K.e_hi <= 1.0e30;
    K.sqr_max_sep = K.e_hi + K.sqr_pimax;
    // double runs: thresholds of the hot loop's float filter, widened by the bound on the float error (filt_margin)
    SumFilt F;
    F.hi = CUDART_INF_F, F.lo = -CUDART_INF_F, F.pimax = CUDART_INF_F;
    if (sizeof(T) == 8) {
        const double ed = 4.76837158203125e-07 * P.coord_max;  // 2^-21 x the largest |coordinate| (periodic shift included)
        double hi2 = (double)K.e_hi, lo2 = (double)K.e_lo;
        if (MODE == CFB_THETA) hi2 = 2.0 * (1.0 - (double)K.e_hi) + 1e-15, lo2 = 2.0 * (1.0 - (double)K.e_lo) - 1e-15;
        if (MODE == CFB_RPPI_MOCKS) hi2 = (double)K.sqr_max_sep;
        F.hi = __double2float_ru(hi2 + filt_margin(hi2, ed));
        if (lo2 > 0.0 && sqrt(lo2) > 4.0 * ed) F.lo = __double2float_rd(lo2 - filt_margin(lo2, ed));
        F.pimax = __double2float_ru((double)K.pimax * (1.0 + 1e-6) + 2.0 * ed);
    }
    const double sqr_max_sep_d = (double)K.sqr_max_sep;

    u64 my_eval = 0, my_jobs = 0, my_levels = 0;
    int drains = 0;  // this warp's drains since it last normalised the block's histogram
    uint32_t it = 0;  // chunks staged so far by this warp (buffer = it & 1)
    constexpr int SPLIT = CFB_TILE / SUM_TILE;  // the gridlink tile table holds 128-particle tiles

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
        if (gw >= P.ntiles * SPLIT) break;
        const int64_t tile = gw / SPLIT;
        const int cellP = P.tile_cell[tile];
        if (!cfb_owns_cell(cellP, P.shard_rank, P.shard_n)) continue;  // another rank's cell
        const int toff = P.tile_off[tile] + (int)(gw % SPLIT) * SUM_TILE;
        const int nP = A.count[cellP];
        if (toff >= nP) continue;
        const int startP = A.start[cellP];
        const int nv = min(SUM_TILE, nP - toff);  // valid primaries of this tile
        const int pa = (nv + 31) >> 5;
        const T nanv = sizeof(T) == 4 ? (T)CUDART_NAN_F : (T)CUDART_NAN;
        const T *pxg = A.x + startP + toff, *pyg = A.y + startP + toff, *pzg = A.z + startP + toff;
        const T *pwg = WGT ? A.w + startP + toff : nullptr;

        int gx = 0, gy = 0, gz = 0, rax = 0, ray = 0, raz = 0;
        long long refA = 0;
        int nrow = 1, rowlen, wz = 1, rx = 0, ry = 0, rz = 0;
        int64_t list0 = 0;
        const int sub2 = LIST ? P.list_sub2 : 1;
        const int refP = LIST ? cellP / sub2 : 0;
        if (LIST) {
            list0 = P.list_off[refP];
            rowlen = (int)(P.list_off[refP + 1] - list0) * sub2;
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
            nrow = 2 * rx + 1;
            rowlen = (2 * ry + 1) * wz;
        }
        const double eps = sizeof(T) == 4 ? 2.384185791015625e-07 /* 2^-22 */ : 4.440892098500626e-16 /* 2^-51 */;

        int qn = 0;
        int cur_code = 0;
        T xq[SUM_PA], yq[SUM_PA], zq[SUM_PA];
        float xh[SUM_PA], yh[SUM_PA], zh[SUM_PA];  // what the hot loop works with: the primaries (float runs) / their float copies
        __syncwarp();  // the previous tile's drains are done with W.prim
#pragma unroll
        for (int p = 0; p < SUM_PA; p++) {
            const int i = lane + 32 * p;
            const bool ok = i < nv;
            xq[p] = ok ? pxg[i] : nanv;
            yq[p] = ok ? pyg[i] : nanv;
            zq[p] = ok ? pzg[i] : nanv;
            xh[p] = (float)xq[p], yh[p] = (float)yq[p], zh[p] = (float)zq[p];
            W.prim[0][2 * lane + p] = xq[p];
            W.prim[1][2 * lane + p] = yq[p];
            W.prim[2][2 * lane + p] = zq[p];
            if (WGT) W.prim[3][2 * lane + p] = ok ? pwg[i] : (T)0;
        }
        __syncwarp();

        for (int row = 0; row < nrow; row++) {
            for (int base = 0; base < rowlen; base += 32) {
                // ---------------- phase 1: one candidate per lane ----------------
                const int cand = base + lane;
                bool keep = cand < rowlen;
                int cellQ = -1, code = 0, kfirst = 1, nl = 0, flags = 0;
                int j_start = 0, j_n = 0, j2_start = 0, j2_n = 0;
                if (keep) {
                    double offd[3] = {0.0, 0.0, 0.0};
                    if (LIST) {
                        const int li = cand / sub2;
                        const int refQ = P.list_cells[list0 + li];
                        cellQ = refQ * sub2 + (cand - li * sub2);
                        // the host lists every unordered pair of reference cells once (and a cell with itself):
                        // inside one reference cell every unordered pair of fine cells is taken once
                        if (P.autocorr && refQ == refP && cellQ > cellP) keep = false;
                        if (!(P.autocorr && refQ == refP)) flags |= SJ_DIRZ;
                    } else {
                        const unsigned iy = fdiv((unsigned)cand, P.m_wz);
                        const int t[3] = {gx + row - rx, gy + (int)iy - ry, gz + (cand - (int)iy * wz) - rz};
                        const int ra[3] = {rax, ray, raz};
                        int q[3], rb[3];
#pragma unroll
                        for (int a = 0; a < 3; a++) {
                            const int ng = P.g.ng[a];
                            if (P.g.periodic[a]) {
                                // one image per side at most; further images are exact duplicates the reference
                                // suppresses (gridlink_utils.h.src:46-72)
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
                                // reference keeps icell2 <= icell (gridlink_impl.c.src:525); within one reference cell
                                // every unordered pair of fine cells is taken once
                                const long long refB = ((long long)rb[0] * P.g.n[1] + rb[1]) * P.g.n[2] + rb[2];
                                if (refB > refA || (refB == refA && cellQ < cellP)) keep = false;
                                if (refB != refA) flags |= SJ_DIRZ;
                                // a cell against its own periodic image: the reference pairs it (all pairs i, j with
                                // the wrap on i), here every |d| >= L/2 > rmax on that axis: nothing in range
                                if (cellQ == cellP && code != 0) keep = false;
                            } else
                                flags |= SJ_DIRZ;
                        }
                    }
                    int nQ = 0;
                    if (keep) {
                        nQ = B.count[cellQ];
                        if (nQ == 0) keep = false;
                    }
                    if (keep) {
                        // ---- conservative interval of the binned quantity over all pairs of the two cells ----
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
                        if (MODE == CFB_WP || MODE == CFB_RPPI) {
                            vlo = dmin[0] * dmin[0] + dmin[1] * dmin[1];
                            vhi = dmax[0] * dmax[0] + dmax[1] * dmax[1];
                            if (dmin[2] >= P.pimax) keep = false;  // every |dz| >= pimax
                        } else {
                            vlo = dmin[0] * dmin[0] + dmin[1] * dmin[1] + dmin[2] * dmin[2];
                            vhi = dmax[0] * dmax[0] + dmax[1] * dmax[1] + dmax[2] * dmax[2];
                        }
                        vlo *= (1.0 - 2.0 * eps);
                        vhi *= (1.0 + 2.0 * eps);
                        if (MODE == CFB_THETA) {  // -cos(theta) = chord^2 / 2 - 1, one more rounding near 1
                            vlo = (0.5 * vlo - 1.0) - 4.0 * eps;
                            vhi = (0.5 * vhi - 1.0) + 4.0 * eps;
                        }
                        if (MODE == CFB_RPPI_MOCKS) {
                            // the interval is that of the 3-D separation: it prunes, and it bounds rp^2 from above only
                            if (vlo >= sqr_max_sep_d) keep = false;
                            vlo = -1.0;
                        } else if (vlo >= s_edges_d[nedges - 1])
                            keep = false;
                        if (vhi < s_edges_d[0]) keep = false;
                        if (keep) {
                            // klo = largest k with E[k] <= vlo, khi = smallest k with E[k] > vhi
                            int a = 0, b = nedges;
                            while (a < b) {
                                const int m = (a + b) >> 1;
                                if (s_edges_d[m] <= vlo) a = m + 1; else b = m;
                            }
                            const int klo = a - 1;
                            b = nedges;
                            while (a < b) {
                                const int m = (a + b) >> 1;
                                if (s_edges_d[m] <= vhi) a = m + 1; else b = m;
                            }
                            const int khi = a;
                            // the pair's bin is kfirst + #{levels k in [kfirst, kfirst + nl) at or below its value}
                            kfirst = klo + 1 > 1 ? klo + 1 : 1;
                            const int klast = khi - 1 < nedges - 2 ? khi - 1 : nedges - 2;
                            nl = klast - kfirst + 1;
                            if (nl < 0) nl = 0;
                            if (klo < 0) flags |= SJ_NEEDLO;
                            const bool tri = P.autocorr && cellQ == cellP;
                            const int startQ = B.start[cellQ];
                            u64 npairs_an;
                            if (tri) {
                                // same cell: the tile against itself + the secondaries after this tile
                                const int after = nP - (toff + SUM_TILE);
                                j_start = startQ + toff;
                                j_n = nv;
                                flags |= SJ_TRI;
                                if (after > 0) {
                                    j2_start = startQ + toff + SUM_TILE;
                                    j2_n = after;
                                }
                                npairs_an = (u64)nv * (u64)(nv - 1) / 2 + (u64)nv * (u64)(after > 0 ? after : 0);
                            } else {
                                j_start = startQ;
                                j_n = nQ;
                                npairs_an = (u64)nv * (u64)nQ;
                            }
                            my_eval += npairs_an;
                            my_levels += npairs_an * (u64)nl;
                        }
                    }
                }
                // ---------------- push the survivors ----------------
                {
                    const unsigned m1 = __ballot_sync(0xffffffffu, keep);
                    const unsigned m2 = __ballot_sync(0xffffffffu, keep && j2_n > 0);
                    const unsigned lt = (1u << lane) - 1u;
                    const int meta = code | (kfirst << 6) | (nl << 14) | flags;
                    if (keep) {
                        SumJob jb;
                        jb.start = j_start;
                        jb.n = j_n;
                        jb.meta = meta;
                        W.q[qn + __popc(m1 & lt)] = jb;
                        if (j2_n > 0) {
                            jb.start = j2_start;
                            jb.n = j2_n;
                            jb.meta = meta & ~SJ_TRI;
                            W.q[qn + __popc(m1) + __popc(m2 & lt)] = jb;
                        }
                    }
                    qn += __popc(m1) + __popc(m2);
                    __syncwarp();
                }
                const bool last = (row == nrow - 1) && (base + 32 >= rowlen);
                if (qn <= SUM_QCAP - 64 && !last) continue;

                // ---------------- phase 2: drain the job queue ----------------
                my_jobs += qn;
                int e = 0, c0 = 0;
                if (qn > 0) {
                    const SumJob jb = W.q[0];
                    stage_chunk<T, NA>(W, it & 1, B, jb.start, (min(SUM_CH, jb.n) + 3) & ~3, lane);
                }
                while (e < qn) {
                    const SumJob jb = W.q[e];
                    const int m4 = (min(SUM_CH, jb.n - c0) + 3) & ~3;
                    int e2 = e, c2 = c0 + SUM_CH;
                    if (c2 >= jb.n) {
                        e2 = e + 1;
                        c2 = 0;
                    }
                    const int bsel = it & 1;
                    if (e2 < qn) {
                        const SumJob jn = W.q[e2];
                        stage_chunk<T, NA>(W, bsel ^ 1, B, jn.start + c2, (min(SUM_CH, jn.n - c2) + 3) & ~3, lane);
                        cp_async_wait<1>();
                    } else
                        cp_async_wait<0>();
                    __syncwarp();

                    // first particle gets the wrap: xpos = x0 + off_xwrap (countpairs_kernels.c.src:79-81)
                    const int jcode = jb.meta & 63;
                    if (jcode != cur_code) {
                        cur_code = jcode;
                        const int cx = jcode & 3, cy = (jcode >> 2) & 3, cz = (jcode >> 4) & 3;
                        const T ox = cx == 1 ? (T)P.wrap[0] : -(T)P.wrap[0];
                        const T oy = cy == 1 ? (T)P.wrap[1] : -(T)P.wrap[1];
                        const T oz = cz == 1 ? (T)P.wrap[2] : -(T)P.wrap[2];
#pragma unroll
                        for (int p = 0; p < SUM_PA; p++) {
                            const int i = lane + 32 * p;
                            if (i < nv) {
                                const T xr = pxg[i], yr = pyg[i], zr = pzg[i];
                                xq[p] = cx ? xr + ox : xr;
                                yq[p] = cy ? yr + oy : yr;
                                zq[p] = cz ? zr + oz : zr;
                                xh[p] = (float)xq[p], yh[p] = (float)yq[p], zh[p] = (float)zq[p];
                                W.prim[0][2 * lane + p] = xq[p];
                                W.prim[1][2 * lane + p] = yq[p];
                                W.prim[2][2 * lane + p] = zq[p];
                            }
                        }
                        __syncwarp();
                    }
                    const int jk = (jb.meta >> 6) & 255, jl = (jb.meta >> 14) & 511;
                    const bool dirz = (jb.meta & SJ_DIRZ) != 0;
                    const bool tri = (jb.meta & SJ_TRI) != 0, needlo = (jb.meta & SJ_NEEDLO) != 0;
                    constexpr bool DIRECT_OK = !AVG && !WGT && (MODE == CFB_DD || MODE == CFB_WP || MODE == CFB_RPPI || MODE == CFB_THETA);
                    if (DIRECT_OK && jl <= 8) {
                        T E[8];  // the job's level edges; beyond nl: an edge no value is at or above
#pragma unroll
                        for (int l = 0; l < 8; l++) {
                            const T none = MODE == CFB_THETA ? (sizeof(T) == 4 ? (T)-CUDART_INF_F : (T)-CUDART_INF)
                                                             : (sizeof(T) == 4 ? (T)CUDART_INF_F : (T)CUDART_INF);
                            E[l] = l < jl ? s_edges[min(jk + l, nedges - 1)] : none;
                        }
#define SUM_DIR(PA_, TRI_, LO_, NL_) \
    direct_loop<T, MODE, NA, PA_, TRI_, LO_, NL_>(m4, W, bsel, xq, yq, zq, K, dirz, c0, lane, E, jk)
#define SUM_DIR3(PA_, NL_)                           \
    do {                                             \
        if (tri) SUM_DIR(PA_, true, true, NL_);      \
        else if (needlo) SUM_DIR(PA_, false, true, NL_); \
        else SUM_DIR(PA_, false, false, NL_);        \
    } while (0)
                        if (jl <= 4) {
                            if (pa == 1) SUM_DIR3(1, DIRECT_OK ? 4 : 1);
                            else SUM_DIR3(2, DIRECT_OK ? 4 : 1);
                        } else {
                            if (pa == 1) SUM_DIR3(1, DIRECT_OK ? 8 : 1);
                            else SUM_DIR3(2, DIRECT_OK ? 8 : 1);
                        }
#undef SUM_DIR3
#undef SUM_DIR
                    } else {
                        const float *sx, *sy, *sz;
                        if constexpr (sizeof(T) == 8) {
                            // float copies of the chunk's positions for the filter
                            for (int v = lane; v < m4; v += 32) {
                                W.fbuf[0][v] = (float)W.buf[bsel][0][v];
                                W.fbuf[1][v] = (float)W.buf[bsel][1][v];
                                W.fbuf[2][v] = (float)W.buf[bsel][2][v];
                            }
                            __syncwarp();
                            sx = W.fbuf[0], sy = W.fbuf[1], sz = W.fbuf[2];
                        } else {
                            sx = (const float *)W.buf[bsel][0], sy = (const float *)W.buf[bsel][1], sz = (const float *)W.buf[bsel][2];
                        }
                        const unsigned char *const qbase = &W.laneq[0][lane];
                        const unsigned qaddr0 = smem_u32(qbase);
                        unsigned qaddr = qaddr0;
                        const unsigned qlimit = qaddr0 + 32 * (SUM_QL - 4);  // an iteration queues at most 4 pairs per lane
                        int k = 0;
                        for (;;) {
#define SUM_HOT(PA_, TRI_, LO_) \
    k = hot_loop<T, MODE, PA_, TRI_, LO_>(k, m4, sx, sy, sz, xh, yh, zh, K, F, dirz, c0, lane, qaddr, qlimit)
#define SUM_HOT3(PA_)                          \
    do {                                       \
        if (tri) SUM_HOT(PA_, true, true);     \
        else if (needlo) SUM_HOT(PA_, false, true); \
        else SUM_HOT(PA_, false, false);       \
    } while (0)
                            if (pa == 1) SUM_HOT3(1);
                            else SUM_HOT3(2);
#undef SUM_HOT3
#undef SUM_HOT
                            // some lane's queue is full or the chunk is done (the drain gathers from this chunk's buffer)
                            flush_queues<T, MODE, AVG, WGT, NA>(W, bsel, K, P, ns, dirz, jk, jl, lane, qbase, qaddr0, qaddr, drains);
                            if (k >= m4) break;
                        }
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
    __syncthreads();
    flush_hist<MODE, AVG, WGT, T>(P, K, ns, tid, blockDim.x, true);
    for (int o = 16; o > 0; o >>= 1) {
        my_eval += __shfl_xor_sync(0xffffffffu, my_eval, o);
        my_jobs += __shfl_xor_sync(0xffffffffu, my_jobs, o);
        my_levels += __shfl_xor_sync(0xffffffffu, my_levels, o);
    }
    if (lane == 0) {
        if (my_eval) atomicAdd(&P.counters[0], my_eval);
        if (my_jobs) atomicAdd(&P.counters[1], my_jobs / 32);  // my_jobs is warp-uniform: the shuffle sum counted it 32x
        if (my_levels) atomicAdd(&P.counters[3], my_levels);
    }
}

// ------------------------------------------------------------------------------------------------
template <typename T>
SetView<T> view_of(const ParticleSet &S)
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

template <typename T, bool WGT>
size_t sum_fixed_smem(int nedges, int nkeys)
{
    constexpr int NA = WGT ? 4 : 3;
    size_t off = (sizeof(SumWarp<T, NA>) * SUM_WARPS + 15) & ~(size_t)15;
    off += ((size_t)nedges * sizeof(T) + 15) & ~(size_t)15;
    off += (size_t)nedges * 16;
    off += ((size_t)nkeys * 2 + 15) & ~(size_t)15;
    return off;
}

template <typename T, int MODE, bool AVG, bool WGT, bool LIST>
int launch_inst(PairParams P, const ParticleSet &SA, const ParticleSet &SB, cudaStream_t st)
{
    auto kern = k_pairs_sum<T, MODE, AVG, WGT, LIST>;
    if (P.ntiles <= 0) return 0;
    int dev = 0, sms = 0, smem_sm = 0, smem_blk = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    CK(cudaDeviceGetAttribute(&smem_blk, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t fixed = sum_fixed_smem<T, WGT>(P.nedges, P.sum_nkeys);
    const size_t per_copy = (size_t)P.nslots * 4 * (1 + (AVG ? 2 : 0) + (WGT ? 2 : 0));
    // histogram copies: as many (1, 2, 4 or 8) as still leave SUM_MINB resident blocks per SM
    int want_blocks = SUM_MINB, max_shift = 3;
    if (const char *e = getenv("CORRFUNC_B200_SUM_BLOCKS")) want_blocks = atoi(e) > 0 ? atoi(e) : want_blocks;
    if (const char *e = getenv("CORRFUNC_B200_SUM_COPIES")) {
        max_shift = 0;
        while ((2 << max_shift) <= atoi(e) && max_shift < 5) max_shift++;
    }
    const size_t per_block_budget = (size_t)smem_sm / want_blocks - 1024;  // 1 KB per block is reserved by the runtime
    int shift = 0;
    while (shift < max_shift && fixed + (per_copy << (shift + 1)) <= per_block_budget &&
           fixed + (per_copy << (shift + 1)) <= (size_t)smem_blk)
        shift++;
    const size_t sm = fixed + (per_copy << shift);
    if (sm > (size_t)smem_blk) return -1;  // histogram too large for shared memory: the caller falls back
    P.sum_copies_shift = shift;
    P.sum_cmask = (1 << shift) - 1;
    P.sum_hstride_b = (unsigned)((4 * (1 + (AVG ? 2 : 0) + (WGT ? 2 : 0))) << shift);
    P.sum_hist_off = (unsigned)fixed;
    P.sum_inv_dmu_f = (float)P.inv_dmu;
    // a word takes at most warps x norm_drains x (32 / copies) adds between two normalisations: keep that below 2^16
    {
        const int lanes_per_copy = 32 >> (shift < 5 ? shift : 5);
        int nd = 65536 / (SUM_WARPS * (lanes_per_copy > 0 ? lanes_per_copy : 1)) - 1 - SUM_QL;
        P.sum_norm_drains = nd < SUM_NORM_DRAINS ? (nd > 1 ? nd : 1) : SUM_NORM_DRAINS;
    }
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SUM_WARPS * 32, sm));
    if (per_sm < 1) per_sm = 1;
    constexpr int SPLIT = CFB_TILE / SUM_TILE;
    int64_t nblk = (P.ntiles * SPLIT / (P.shard_n > 1 ? P.shard_n : 1) + SUM_WARPS - 1) / SUM_WARPS + 1;
    if (nblk > (int64_t)sms * per_sm) nblk = (int64_t)sms * per_sm;
    kern<<<(unsigned int)nblk, SUM_WARPS * 32, sm, st>>>(P, view_of<T>(SA), view_of<T>(SB));
    cfb_ctx().launches++;
    CK(cudaGetLastError());
    return 0;
}

template <typename T, int MODE, bool LIST>
int launch_mode(const cfb_binning *bin, const PairParams &P, const ParticleSet &SA, const ParticleSet &SB, cudaStream_t st)
{
    const bool a = bin->need_avg != 0, w = bin->need_weights != 0;
    if (a && w) return launch_inst<T, MODE, true, true, LIST>(P, SA, SB, st);
    if (a) return launch_inst<T, MODE, true, false, LIST>(P, SA, SB, st);
    if (w) return launch_inst<T, MODE, false, true, LIST>(P, SA, SB, st);
    return launch_inst<T, MODE, false, false, LIST>(P, SA, SB, st);
}

template <typename T>
int launch_T(const cfb_binning *bin, const PairParams &P, bool list_mode)
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

}  // namespace

// 0 = launched, -1 = not applicable (too many edges / histogram too large: use the generic kernel), 1 = error.
// P.wmax must be set when weights are on (cfb_weight_maxima).
int cfb_launch_pairs_sum(const cfb_binning *bin, const PairParams &P0, int prec, bool list_mode)
{
    PairParams P = P0;
    static_assert(CFB_TILE % SUM_TILE == 0, "sum tiles subdivide the gridlink tiles");
    static_assert(SUM_QL % 4 == 0 && SUM_PA == 2, "queue compaction is unrolled by four; the primaries of a lane are interleaved in pairs");
    static_assert(SUM_TILE <= 256 && SUM_CH <= 256, "a stack entry is (primary << 8 | secondary) in 16 bits");
    if (bin->nedges < 2 || bin->nedges > SUM_MAX_EDGES) return -1;
    for (int i = 0; i < bin->nedges; i++) {
        // the level search needs finite, strictly monotonic edges (increasing; theta: decreasing cosines)
        if (!isfinite(bin->edges[i])) return -1;
        if (i > 0 && (bin->mode == CFB_THETA ? !(bin->edges[i] < bin->edges[i - 1]) : !(bin->edges[i] > bin->edges[i - 1])))
            return -1;
    }
    if (bin->nslots >= (1 << 20)) return -1;
    // separation-bin table (see drain): keys of the float images of [first edge, last edge]
    P.sum_kmin = 0;
    P.sum_nkeys = 0;
    if (bin->mode != CFB_THETA && bin->edges[0] > 1.0e-30 && bin->edges[bin->nedges - 1] < 1.0e30) {
        const float flo = (float)(bin->edges[0] * (1.0 - 1e-6)), fhi = (float)(bin->edges[bin->nedges - 1] * (1.0 + 1e-6));
        unsigned blo, bhi;
        memcpy(&blo, &flo, 4);
        memcpy(&bhi, &fhi, 4);
        const int kmin = (int)(blo >> SUM_KSH), kmax = (int)(bhi >> SUM_KSH);
        if (kmax - kmin + 1 <= SUM_MAX_KEYS) {
            P.sum_kmin = kmin;
            P.sum_nkeys = kmax - kmin + 1;
        }
    }
    return prec == 4 ? launch_T<float>(bin, P, list_mode) : launch_T<double>(bin, P, list_mode);
}
