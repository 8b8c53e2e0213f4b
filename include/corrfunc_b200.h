/* corrfunc_b200.h -- extensions of libcorrfunc_b200.so beyond the reference's C API.
 * None of these exist in the reference; they expose what a GPU build needs in addition:
 * multi-GPU work sharding (one process per GPU), per-call measurements, and tuning knobs.
 */
#ifndef CORRFUNC_B200_H
#define CORRFUNC_B200_H
#include <stdint.h>
#include "corrfunc_b200_device.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Multi-GPU: each process (rank) counts a disjoint share of the primary cells against a full replica
 * of the particles; the small per-bin histograms are then summed across ranks by `fn` (e.g. an NCCL
 * all-reduce issued through torch.distributed) BEFORE the host epilogue (x2, self pairs, averages,
 * xi/wp estimators) runs, so every rank returns the complete result.  Integer sums are exact, hence
 * npairs does not depend on the number of GPUs. fn returns 0 on success. */
typedef int (*corrfunc_b200_reduce_fn)(uint64_t *npairs, double *sum_sep, double *sum_w, int64_t nslots, void *user);
void corrfunc_b200_set_shard(int rank, int nranks);
void corrfunc_b200_set_reduce_hook(corrfunc_b200_reduce_fn fn, void *user);

/* Measurements of the most recent countpairs* call of this process. */
typedef struct {
    cfb_stats dev;        /* device-side timings and counters */
    double ms_host_total; /* wall time of the whole C call, host clock */
    double ms_upload;     /* wall time spent in cfb_upload (H2D) */
    int nmesh[3];         /* reference lattice */
    int refine[3];        /* bin refine factors after the reference's heuristics */
} corrfunc_b200_stats;
const corrfunc_b200_stats *corrfunc_b200_last_stats(void);

const char *corrfunc_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
