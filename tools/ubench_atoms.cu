// ubench_atoms.cu -- how fast can a B200 SM update a shared-memory histogram with SPREAD addresses?
// Decides the accumulation scheme of the per-pair-sum kernel (pairs_sum.cu).  Variants, all on a per-block histogram of
// S slots with pseudo-random slot per lane per step (a hot set of H slots):
//   0  ATOMS.ADD u32, result unused            (one word per update)
//   1  ATOMS.ADD u32, result used (carry chain of two dependent atomics, like the 96-bit fixed-point adds)
//   2  per-WARP private histogram, plain LDS/STS read-modify-write; intra-warp collisions resolved with an owner tag
//      (STS tag, LDS tag, winners update, losers retry)
//   3  like 2, collisions resolved with __match_any_sync
//   4  ATOMS.ADD u32 with only 8 of 32 lanes active (what the divergent round-1 kernel did)
// Reports lane-updates per clock per SM.   nvcc -arch=sm_100a -O3 -o tools/bin/ubench_atoms tools/ubench_atoms.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ unsigned rng(unsigned &s)
{
    s = s * 1664525u + 1013904223u;
    return s >> 8;
}

template <int VAR, int WORDS>
__global__ void k(const int S, const int H, const int steps, unsigned long long *out, long long *cyc)
{
    extern __shared__ unsigned sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    // variants 2, 3: per-warp private area of S * WORDS words + S tag words; others: one block histogram
    unsigned *hist = (VAR == 2 || VAR == 3) ? sm + (size_t)wid * S * (WORDS + 1) : sm;
    unsigned *tag = hist + (size_t)S * WORDS;
    const int tot = (VAR == 2 || VAR == 3) ? nw * S * (WORDS + 1) : S * WORDS;
    for (int i = threadIdx.x; i < tot; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    unsigned s = 12345u + threadIdx.x * 7919u + blockIdx.x * 104729u;
    const long long t0 = clock64();
    for (int it = 0; it < steps; it++) {
        const int slot = (int)(rng(s) % (unsigned)H);
        const unsigned val = (s >> 4) | 1u;
        if (VAR == 0) {
#pragma unroll
            for (int w = 0; w < WORDS; w++) atomicAdd(&hist[w * S + slot], val + w);
        } else if (VAR == 1) {
            unsigned carry = 0;
#pragma unroll
            for (int w = 0; w < WORDS; w++) {
                const unsigned add = val + carry;
                const unsigned old = atomicAdd(&hist[w * S + slot], add);
                carry = (old + add) < add ? 1u : 0u;
            }
        } else if (VAR == 4) {
            if ((lane & 3) == (it & 3)) {
#pragma unroll
                for (int w = 0; w < WORDS; w++) atomicAdd(&hist[w * S + slot], val + w);
            }
        } else if (VAR == 2) {
            bool todo = true;
            while (__any_sync(0xffffffffu, todo)) {
                if (todo) tag[slot] = lane;
                __syncwarp();
                const bool win = todo && tag[slot] == (unsigned)lane;
                if (win) {
#pragma unroll
                    for (int w = 0; w < WORDS; w++) hist[w * S + slot] += val + w;
                    todo = false;
                }
                __syncwarp();
            }
        } else if (VAR == 3) {
            const unsigned peers = __match_any_sync(0xffffffffu, slot);
            // serialise within a peer group: rank r updates in round r
            const int rank = __popc(peers & ((1u << lane) - 1u));
            const int rounds = __reduce_max_sync(0xffffffffu, (unsigned)rank) + 1;
            for (int r = 0; r < rounds; r++) {
                if (rank == r) {
#pragma unroll
                    for (int w = 0; w < WORDS; w++) hist[w * S + slot] += val + w;
                }
                __syncwarp();
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    unsigned long long acc = 0;
    for (int i = threadIdx.x; i < tot; i += blockDim.x) acc += sm[i];
    atomicAdd(out, acc);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int VAR, int WORDS>
static void run(const char *name, int S, int H, int warps, int blocks_per_sm)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int steps = 20000;
    const int nb = sms * blocks_per_sm;
    unsigned long long *out;
    long long *cyc;
    cudaMalloc(&out, 8);
    cudaMalloc(&cyc, nb * 8);
    cudaMemset(out, 0, 8);
    const size_t smem = (VAR == 2 || VAR == 3) ? (size_t)warps * S * (WORDS + 1) * 4 : (size_t)S * WORDS * 4;
    cudaFuncSetAttribute(k<VAR, WORDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<VAR, WORDS>, warps * 32, smem);
    if (occ < blocks_per_sm) {
        printf("%-28s S=%4d H=%4d words=%d warps/blk=%2d blk/SM=%d : does not fit (occ %d)\n", name, S, H, WORDS, warps, blocks_per_sm, occ);
        return;
    }
    k<VAR, WORDS><<<nb, warps * 32, smem>>>(S, H, 100, out, cyc);
    cudaDeviceSynchronize();
    k<VAR, WORDS><<<nb, warps * 32, smem>>>(S, H, steps, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%s: %s\n", name, cudaGetErrorString(e));
        exit(1);
    }
    long long *h = (long long *)malloc(nb * 8);
    cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < nb; i++) avg += (double)h[i];
    avg /= nb;
    const double lanes = (VAR == 4 ? 8.0 : 32.0);
    // updates of one slot (all its words) per clock per SM
    const double upd = (double)steps * lanes * warps * blocks_per_sm / avg;
    printf("%-28s S=%4d H=%4d words=%d warps/blk=%2d blk/SM=%d : %7.3f slot-updates/clk/SM  (%6.3f word-updates/clk/SM)\n", name, S, H,
           WORDS, warps, blocks_per_sm, upd, upd * WORDS);
    free(h);
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    for (int H : {462, 64, 8}) {
        for (int wpb : {4, 8, 16}) {
            run<0, 1>("ATOMS no-return", 512, H <= 512 ? H : 512, wpb, 2);
            run<0, 5>("ATOMS no-return", 512, H, wpb, 2);
            run<1, 5>("ATOMS carry chain", 512, H, wpb, 2);
            run<4, 5>("ATOMS 8 lanes", 512, H, wpb, 2);
            if (wpb <= 8) {
                run<2, 5>("private RMW, tag", 512, H, wpb, 1);
                run<3, 5>("private RMW, match", 512, H, wpb, 1);
            }
        }
    }
    return 0;
}
