// context.cu -- process-wide device context, buffer pool, uploads, extents and the count_* drivers.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <thread>
#include <utility>
#include <vector>

#include "cfb_internal.cuh"

// One context per device.  Context 0 is the one every call starts on (CORRFUNC_B200_DEVICE or the current device); the
// others exist only while a pair count is sharded over several devices INSIDE one call (count_box_multi below), each
// driven by its own host thread whose thread-local `tl_ctx` points at it.
#define CFB_MAX_DEV 16
static Ctx g_ctxs[CFB_MAX_DEV];
static thread_local Ctx *tl_ctx = &g_ctxs[0];
#define g_ctx (*tl_ctx)
// persistent pinned host buffers handed to the host layer (cfb_host_scratch)
static void *g_host_scratch[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
static size_t g_host_scratch_cap[6] = {0, 0, 0, 0, 0, 0};
Ctx &cfb_ctx() { return *tl_ctx; }

int cfb_fail(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_ctx.err, sizeof(g_ctx.err), fmt, ap);
    va_end(ap);
    fprintf(stderr, "corrfunc_b200> %s\n", g_ctx.err);
    return 1;
}

extern "C" const char *cfb_last_error(void) { return g_ctx.err; }

int cfb_ensure(DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap && b.p) return 0;
    if (b.p) {
        cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;  // a little slack so slightly larger repeat calls do not realloc
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        want = bytes;
        e = cudaMalloc(&b.p, want);
    }
    if (e != cudaSuccess) return cfb_fail("cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    b.cap = want;
    return 0;
}

static int init_ctx(Ctx &c, int dev)
{
    if (c.ready) return 0;
    CK(cudaSetDevice(dev));
    c.dev = dev;
    CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    for (int i = 0; i < 8; i++) CK(cudaEventCreate(&c.ev[i]));
    c.pinned_cap = 1 << 20;
    CK(cudaMallocHost(&c.pinned, c.pinned_cap));
    c.ready = true;
    return 0;
}

extern "C" int cfb_init(void)
{
    Ctx &c = g_ctxs[0];
    if (c.ready) return 0;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return cfb_fail("no CUDA device available (%s); corrfunc_b200 has no CPU fallback",
                        e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    int dev = 0;
    const char *env = getenv("CORRFUNC_B200_DEVICE");
    if (env && *env)
        dev = atoi(env);
    else
        cudaGetDevice(&dev);
    if (dev < 0 || dev >= ndev) return cfb_fail("CORRFUNC_B200_DEVICE=%d out of range (have %d devices)", dev, ndev);
    return init_ctx(c, dev);
}

static void release_ctx(Ctx &c)
{
    if (!c.ready) return;
    cudaSetDevice(c.dev);
    cudaStreamSynchronize(c.stream);
    auto rel = [](DevBuf &b) {
        if (b.p) cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    };
    for (int s = 0; s < 2; s++) {
        ParticleSet &S = c.set[s];
        for (int i = 0; i < 6; i++) rel(S.rawbuf[i]);
        for (int i = 0; i < 4; i++) rel(S.sorted[i]);
        rel(S.cidx); rel(S.rank); rel(S.count); rel(S.start); rel(S.tstart); rel(S.bounds);
        rel(S.tile_cell); rel(S.tile_off);
        S = ParticleSet();
    }
    rel(c.scratch); rel(c.hist); rel(c.edges); rel(c.list_off); rel(c.list_cells); rel(c.ngrid_ra); rel(c.ra_off);
    rel(c.sort_rec); rel(c.sort_cid); rel(c.sort_cur);
    if (c.pinned) cudaFreeHost(c.pinned);
    c.pinned = nullptr;
    for (int i = 0; i < 8; i++) cudaEventDestroy(c.ev[i]);
    cudaStreamDestroy(c.stream);
    c.ready = false;
}

extern "C" void cfb_shutdown(void)
{
    if (!g_ctxs[0].ready) return;
    cudaSetDevice(g_ctxs[0].dev);
    cfb_spheres_release();
    for (int i = 0; i < 6; i++) {
        if (g_host_scratch[i]) cudaFreeHost(g_host_scratch[i]);
        g_host_scratch[i] = nullptr;
        g_host_scratch_cap[i] = 0;
    }
    for (int d = CFB_MAX_DEV - 1; d >= 0; d--) release_ctx(g_ctxs[d]);
    cudaSetDevice(g_ctxs[0].dev);
}

extern "C" void cfb_set_shard(int rank, int nranks)
{
    if (nranks < 1) nranks = 1;
    if (rank < 0 || rank >= nranks) rank = 0;
    g_ctx.shard_rank = rank;
    g_ctx.shard_n = nranks;
}
extern "C" void cfb_get_shard(int *rank, int *nranks)
{
    if (rank) *rank = g_ctx.shard_rank;
    if (nranks) *nranks = g_ctx.shard_n;
}
extern "C" void cfb_set_target_occupancy(int n) { g_ctx.target_occ = n; }
extern "C" void cfb_force_kernel(int kind) { g_ctx.force_kernel = kind; }

static bool is_device_ptr(const void *p)
{
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

extern "C" int cfb_is_device_ptr(const void *p) { return (p && cfb_init() == 0 && is_device_ptr(p)) ? 1 : 0; }
extern "C" int cfb_copy_to_host(void *dst, const void *src, size_t bytes)
{
    if (cfb_init()) return 1;
    CK(cudaSetDevice(g_ctxs[0].dev));
    CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

// Persistent pinned host buffers for arrays the host layer computes itself (DDtheta's unit vectors): written
// without page faults on repeated calls and uploaded at full PCIe rate.  Grow-only; freed by cfb_shutdown.
extern "C" void *cfb_host_scratch(int which, size_t bytes)
{
    if (which < 0 || which >= 6) return nullptr;
    if (cfb_init()) return nullptr;
    if (bytes <= g_host_scratch_cap[which]) return g_host_scratch[which];
    if (g_host_scratch[which]) cudaFreeHost(g_host_scratch[which]);
    g_host_scratch[which] = nullptr;
    g_host_scratch_cap[which] = 0;
    const size_t cap = bytes + bytes / 8 + 4096;
    if (cudaMallocHost(&g_host_scratch[which], cap) != cudaSuccess) {
        cudaGetLastError();
        g_host_scratch[which] = nullptr;
        return nullptr;
    }
    g_host_scratch_cap[which] = cap;
    return g_host_scratch[which];
}

// Catalogue cache: between cfb_set_catalog_cache(1) and (0) the caller promises that arrays passed under the same
// pointers hold the same values.  A set whose pointers, length and element size match what a slot already holds (either
// slot: the two are swapped when needed) is not copied again, and it keeps its sorted form while the lattice stays the
// same.  This is what the DD / DR / RR workflow wants (Corrfunc/utils.py:27-165: three counts over two catalogues):
// D and R cross PCIe once instead of twice each.
static bool g_cache_on = false;
static long long g_cache_hits = 0;
extern "C" void cfb_set_catalog_cache(int on)
{
    g_cache_on = on != 0;
    if (!g_cache_on)
        for (int d = 0; d < CFB_MAX_DEV; d++)
            for (int s = 0; s < 2; s++) g_ctxs[d].set[s].src_valid = false;
}
extern "C" long long cfb_catalog_cache_hits(void) { return g_cache_hits; }

static bool same_source(const ParticleSet &S, int prec, int64_t n, const void *const src[6])
{
    if (!S.src_valid || S.prec != prec || S.n != n) return false;
    for (int i = 0; i < 6; i++)
        if (S.src[i] != src[i]) return false;
    return true;
}

extern "C" int cfb_upload(int slot, int prec, int64_t n, const void *x, const void *y, const void *z, const void *w,
                          const void *ra, const void *dec)
{
    tl_ctx = &g_ctxs[0];
    if (cfb_init()) return 1;
    Ctx &c = g_ctx;
    CK(cudaSetDevice(c.dev));
    if (slot < 0 || slot > 1) return cfb_fail("bad particle slot %d", slot);
    if (prec != 4 && prec != 8) return cfb_fail("element size must be 4 or 8 (got %d)", prec);
    if (n < 0 || n >= (int64_t)2000000000) return cfb_fail("particle count %lld not supported", (long long)n);
    const void *src[6] = {x, y, z, w, ra, dec};
    if (g_cache_on && n > 0) {
        if (!same_source(c.set[slot], prec, n, src) && same_source(c.set[1 - slot], prec, n, src))
            std::swap(c.set[0], c.set[1]);
        if (same_source(c.set[slot], prec, n, src)) {
            g_cache_hits++;
            return 0;
        }
    }
    ParticleSet &S = c.set[slot];
    S.prec = prec;
    S.n = n;
    S.gridded = false;
    S.grid_sig_valid = false;
    S.src_valid = g_cache_on;
    for (int i = 0; i < 6; i++) S.src[i] = src[i];
    const size_t bytes = (size_t)n * prec;
    for (int i = 0; i < 6; i++) {
        S.raw[i] = nullptr;
        if (!src[i] || n == 0) continue;
        if (is_device_ptr(src[i])) {
            S.raw[i] = src[i];  // borrowed for the duration of the call
        } else {
            if (cfb_ensure(S.rawbuf[i], bytes)) return 1;
            CK(cudaMemcpyAsync(S.rawbuf[i].p, src[i], bytes, cudaMemcpyHostToDevice, c.stream));
            S.raw[i] = S.rawbuf[i].p;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// extents: per-block partial min/max, finished on the host (a few KB)
template <typename T>
__global__ void k_minmax(int64_t n, const T *a, const T *b, const T *cc, T *out)
{
    T lo[3], hi[3];
    const T big = sizeof(T) == 4 ? (T)3.402823466e+38F : (T)1.7976931348623157e+308;
    for (int k = 0; k < 3; k++) {
        lo[k] = big;
        hi[k] = -big;
    }
    const T *arr[3] = {a, b, cc};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        for (int k = 0; k < 3; k++) {
            if (!arr[k]) continue;
            const T v = arr[k][i];
            if (v < lo[k]) lo[k] = v;
            if (v > hi[k]) hi[k] = v;
        }
    }
    __shared__ T s[6][256];
    for (int k = 0; k < 3; k++) {
        s[k][threadIdx.x] = lo[k];
        s[3 + k][threadIdx.x] = hi[k];
    }
    __syncthreads();
    for (int st = blockDim.x / 2; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) {
            for (int k = 0; k < 3; k++) {
                const T l2 = s[k][threadIdx.x + st], h2 = s[3 + k][threadIdx.x + st];
                if (l2 < s[k][threadIdx.x]) s[k][threadIdx.x] = l2;
                if (h2 > s[3 + k][threadIdx.x]) s[3 + k][threadIdx.x] = h2;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int k = 0; k < 6; k++) out[(size_t)blockIdx.x * 6 + k] = s[k][0];
}

template <typename T>
static int extent_T(Ctx &c, ParticleSet &S, int which, double lohi[6])
{
    const int nb = 512;
    if (cfb_ensure(c.scratch, (size_t)nb * 6 * sizeof(T) + 4096)) return 1;
    const T *a = (const T *)S.raw[which ? 4 : 0], *b = (const T *)S.raw[which ? 5 : 1],
            *cc = which ? nullptr : (const T *)S.raw[2];
    k_minmax<T><<<nb, 256, 0, c.stream>>>(S.n, a, b, cc, (T *)c.scratch.p);
    c.launches++;
    CK(cudaGetLastError());
    T *h = (T *)c.pinned;
    CK(cudaMemcpyAsync(h, c.scratch.p, (size_t)nb * 6 * sizeof(T), cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    for (int blk = 0; blk < nb; blk++)
        for (int k = 0; k < 3; k++) {
            const double l = (double)h[blk * 6 + k], u = (double)h[blk * 6 + 3 + k];
            if (l < lohi[k]) lohi[k] = l;
            if (u > lohi[3 + k]) lohi[3 + k] = u;
        }
    return 0;
}

extern "C" int cfb_extent(int slot, int which, double lohi[6])
{
    Ctx &c = g_ctx;
    if (!c.ready) return cfb_fail("cfb_extent before cfb_upload");
    ParticleSet &S = c.set[slot];
    if (S.n == 0) return 0;
    return S.prec == 4 ? extent_T<float>(c, S, which, lohi) : extent_T<double>(c, S, which, lohi);
}

// ---------------------------------------------------------------------------------------------
// Fine lattice: every reference cell is split sub[] ways so that a fine cell holds about `target`
// particles (one warp tile = up to 128) and is as close to a cube as the integer splits allow.
static void choose_subdivision(const Ctx &c, const cfb_box_lattice *lat, int64_t nmax, int sub[3], int dflt)
{
    sub[0] = sub[1] = sub[2] = 1;
    const int target = c.target_occ > 0 ? c.target_occ : dflt;  // mostly 4 primaries per lane (97..128 per tile): measured best on config 5
    const double ncell = (double)lat->nmesh[0] * lat->nmesh[1] * lat->nmesh[2];
    const double occ = (double)nmax / ncell;
    if (occ <= 1.3 * target) return;
    double d[3];
    for (int k = 0; k < 3; k++) d[k] = lat->inv[k] > 0 ? 1.0 / lat->inv[k] : 0.0;
    if (!(d[0] > 0 && d[1] > 0 && d[2] > 0)) return;
    double best = 1e300;
    for (int a = 1; a <= 16; a++)
        for (int b = 1; b <= 16; b++)
            for (int e = 1; e <= 16; e++) {
                const double tot = ncell * a * b * e;
                if (tot > 48e6) continue;  // keep the fine lattice below ~48M cells
                const double o = occ / ((double)a * b * e);
                const double x = d[0] / a, y = d[1] / b, z = d[2] / e;
                const double mx = fmax(x, fmax(y, z)), mn = fmin(x, fmin(y, z));
                // over-full cells (second, nearly empty tile) cost more than slightly emptier ones
                const double score = (o > target ? 3.0 : 1.5) * fabs(log(o / target)) + log(mx / mn);
                if (score < best) {
                    best = score;
                    sub[0] = a;
                    sub[1] = b;
                    sub[2] = e;
                }
            }
}

static unsigned div_magic(unsigned d) { return d <= 1 ? 0u : (unsigned)((0x100000000ULL + d - 1) / d); }

static int alloc_hist(Ctx &c, int64_t nslots)
{
    const size_t bytes = (size_t)nslots * 8 * 3 + 64;
    if (cfb_ensure(c.hist, bytes)) return 1;
    CK(cudaMemsetAsync(c.hist.p, 0, bytes, c.stream));
    return 0;
}

static int fetch_hist(Ctx &c, int64_t nslots, const cfb_binning *bin, cfb_hist *out, cfb_stats *stats)
{
    const size_t bytes = (size_t)nslots * 8 * 3 + 64;
    void *h = c.pinned;
    bool tmp = false;
    if (bytes > c.pinned_cap) {
        h = malloc(bytes);
        tmp = true;
        if (!h) return cfb_fail("out of host memory");
    }
    CK(cudaMemcpyAsync(h, c.hist.p, bytes, cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    const unsigned long long *n = (const unsigned long long *)h;
    const double *s = (const double *)((char *)h + (size_t)nslots * 8);
    const double *w = (const double *)((char *)h + (size_t)nslots * 16);
    const unsigned long long *cnt = (const unsigned long long *)((char *)h + (size_t)nslots * 24);
    for (int64_t i = 0; i < nslots; i++) {
        out->npairs[i] = n[i];
        if (out->sum_sep) out->sum_sep[i] = bin->need_avg ? s[i] : 0.0;
        if (out->sum_w) out->sum_w[i] = bin->need_weights ? w[i] : 0.0;
    }
    if (stats) {
        stats->n_eval = cnt[0];
        stats->n_tilepairs = cnt[1];
        stats->n_analytic = cnt[2];
        stats->n_levelpairs = cnt[3];
    }
    if (tmp) free(h);
    return 0;
}

static int upload_edges(Ctx &c, const cfb_binning *bin, const double *edges)
{
    if (bin->nedges < 2 || bin->nedges > CFB_MAX_EDGES) return cfb_fail("number of bin edges %d not in [2,%d]", bin->nedges, CFB_MAX_EDGES);
    if (cfb_ensure(c.edges, (size_t)bin->nedges * 8)) return 1;
    // convert on the host into the run precision (values are already REAL-representable)
    if (bin->prec == 4) {
        float *h = (float *)c.pinned;
        for (int i = 0; i < bin->nedges; i++) h[i] = (float)edges[i];
        CK(cudaMemcpyAsync(c.edges.p, h, (size_t)bin->nedges * 4, cudaMemcpyHostToDevice, c.stream));
    } else {
        double *h = (double *)c.pinned;
        for (int i = 0; i < bin->nedges; i++) h[i] = edges[i];
        CK(cudaMemcpyAsync(c.edges.p, h, (size_t)bin->nedges * 8, cudaMemcpyHostToDevice, c.stream));
    }
    CK(cudaStreamSynchronize(c.stream));  // pinned staging area is reused below
    return 0;
}

// One int in mapped, portable pinned memory: the host layer's signal handler writes it, every device reads it.
static volatile int g_abort_fallback = 0;
static volatile int *g_abort_host = nullptr;
static const volatile int *g_abort_dev = nullptr;
extern "C" volatile int *cfb_abort_flag(void)
{
    if (g_abort_host) return g_abort_host;
    void *h = nullptr, *d = nullptr;
    if (cudaHostAlloc(&h, 64, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess &&
        cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) {
        memset(h, 0, 64);
        g_abort_host = (volatile int *)h;
        g_abort_dev = (const volatile int *)d;
    } else {
        cudaGetLastError();
        g_abort_host = &g_abort_fallback;  // no device: nothing will poll it
        g_abort_dev = nullptr;
    }
    return g_abort_host;
}

static void fill_common(PairParams &P, Ctx &c, const cfb_binning *bin, int64_t nslots)
{
    memset(&P, 0, sizeof(P));
    cfb_abort_flag();
    P.abort = g_abort_dev;
    P.mode = bin->mode;
    P.nedges = bin->nedges;
    P.npibin = bin->npibin;
    P.nmu_bins = bin->nmu_bins;
    P.autocorr = bin->autocorr;
    P.cross = bin->autocorr ? 0 : 1;
    P.nslots = nslots;
    P.pimax = bin->pimax;
    P.inv_dpi = bin->inv_dpi;
    P.sqr_mumax = bin->sqr_mumax;
    P.inv_dmu = bin->inv_dmu;
    P.fast_acos = bin->fast_acos;
    P.edges = c.edges.p;
    P.npairs = (unsigned long long *)c.hist.p;
    P.sum_sep = (double *)((char *)c.hist.p + (size_t)nslots * 8);
    P.sum_w = (double *)((char *)c.hist.p + (size_t)nslots * 16);
    P.counters = (unsigned long long *)((char *)c.hist.p + (size_t)nslots * 24);
    P.shard_rank = c.shard_rank;
    P.shard_n = c.shard_n;
}

// The fast kernel (pairs_fast.cu) handles the 1-D statistics without per-pair sums.  In float it
// needs the positions multiplied by 2^k such that the smallest positive (squared) edge is >= 2^24:
// then two distinct floats at or above that edge differ by >= 1 and one saturating subtract is an
// exact [v < edge] indicator.  Power-of-two scaling commutes with every rounding the reference does.
static bool fast_box_plan(const Ctx &c, const cfb_binning *bin, const cfb_box_lattice *lat, double *scale)
{
    *scale = 1.0;
    if (c.force_kernel == 0) return false;
    if (!(bin->mode == CFB_DD || bin->mode == CFB_XI || bin->mode == CFB_WP)) return false;
    if (bin->need_avg || bin->need_weights) return false;
    if (bin->nedges < 2 || bin->nedges > CFB_FAST_MAX_EDGES) return false;
    for (int i = 0; i < bin->nedges; i++) {
        if (!(bin->edges[i] >= 0.0) || !isfinite(bin->edges[i])) return false;
        if (i > 0 && !(bin->edges[i] > bin->edges[i - 1])) return false;
    }
    if (bin->prec == 8) return true;
    double emin = 0.0;
    for (int i = 0; i < bin->nedges; i++)
        if (bin->edges[i] > 0.0) {
            emin = bin->edges[i];
            break;
        }
    if (!(emin > 0.0)) return false;
    int e = 0;
    frexp(emin, &e);  // emin = m * 2^e, m in [0.5, 1)  ->  emin >= 2^(e-1)
    int k = (25 - e + 1) / 2;  // ceil((25 - e) / 2) for positive numerators
    if (25 - e <= 0) k = 0;
    double cmax = 1.0;
    for (int a = 0; a < 3; a++) {
        const double ext = lat->inv[a] > 0 ? lat->nmesh[a] / lat->inv[a] : 0.0;
        const double m = fabs(lat->lo[a]) + ext + fabs(lat->wrap[a]);
        if (m > cmax) cmax = m;
    }
    int ce = 0;
    frexp(cmax, &ce);
    if (ce + k > 55) return false;  // keep 3 * (2 cmax 2^k)^2 far below FLT_MAX
    *scale = ldexp(1.0, k);
    return true;
}

// One device's share of a box count (all of it when shard_n == 1): gridlink, pair kernel, histogram read-back.
static int count_box_one(const cfb_binning *bin, const cfb_box_lattice *lat, cfb_hist *out, cfb_stats *stats)
{
    Ctx &c = g_ctx;
    if (!c.ready) return cfb_fail("cfb_count_box before cfb_upload");
    CK(cudaSetDevice(c.dev));
    const int nsets = bin->autocorr ? 1 : 2;
    for (int s = 0; s < nsets; s++)
        if (c.set[s].prec != bin->prec) return cfb_fail("particle set %d precision mismatch", s);
    if (bin->need_weights)
        for (int s = 0; s < nsets; s++)
            if (c.set[s].n > 0 && !c.set[s].raw[3]) return cfb_fail("weights requested but set %d has none", s);
    c.launches = 0;
    cfb_stats st;
    memset(&st, 0, sizeof(st));
    CK(cudaEventRecord(c.ev[0], c.stream));

    int sub[3];
    int64_t nmax = c.set[0].n;
    if (nsets == 2 && c.set[1].n > nmax) nmax = c.set[1].n;
    double scale = 1.0;
    const bool fast = fast_box_plan(c, bin, lat, &scale);
    // fine cells: ~112 particles (4 primaries per lane) for the fast kernel; the per-pair-sum kernel works on 64-particle
    // tiles and gains more from tighter pruning than it loses to more jobs
    static int sum_occ = -1;
    if (sum_occ < 0) {
        const char *e = getenv("CORRFUNC_B200_SUM_OCC");
        sum_occ = (e && atoi(e) > 0) ? atoi(e) : 56;
    }
    // count-only DDrppi: the one-thread-per-primary generic kernel (pairs_generic.cu) on 112-particle cells is the fastest
    // of the three for it -- measured on config 2, double / float: 12.0 / 10.4 ms against 14.7 / 13.2 ms for the
    // per-pair-sum kernel (its range tests and 2-D slot arithmetic run for every lane anyway, so compaction buys nothing)
    const bool legacy_rppi = !fast && bin->mode == CFB_RPPI && !bin->need_avg && !bin->need_weights;
    c.prefer_legacy = legacy_rppi;
    choose_subdivision(c, lat, nmax, sub, (fast || legacy_rppi) ? 112 : sum_occ);
    for (int s = 0; s < nsets; s++)
        if (cfb_gridlink_box_set(c.set[s], lat, sub, scale)) return 1;
    CK(cudaEventRecord(c.ev[1], c.stream));

    const int64_t nslots = bin->nslots;
    if (alloc_hist(c, nslots)) return 1;
    double edges_v[CFB_FAST_MAX_EDGES];
    if (fast) {
        for (int i = 0; i < bin->nedges; i++) edges_v[i] = bin->edges[i] * scale * scale;
        if (upload_edges(c, bin, edges_v)) return 1;
    } else if (upload_edges(c, bin, bin->edges))
        return 1;
    PairParams P;
    fill_common(P, c, bin, nslots);
    P.pimax *= scale;
    for (int k = 0; k < 3; k++) {
        P.g.n[k] = lat->nmesh[k];
        P.g.s[k] = sub[k];
        P.g.ng[k] = lat->nmesh[k] * sub[k];
        P.g.refine[k] = lat->refine[k];
        P.g.reach[k] = sub[k] > 1 ? (lat->refine[k] + 1) * sub[k] - 1 : lat->refine[k];
        if (sub[k] > 1 && lat->inv[k] > 0) {
            // fine cells more than ceil(maxsep / fine size) + 1 away (two cells of slack) cannot hold a pair in range
            const double rmax = sqrt(bin->edges[bin->nedges - 1]);
            double sep = (k == 2 && (bin->mode == CFB_WP || bin->mode == CFB_RPPI)) ? bin->pimax : rmax;
            if (bin->mode == CFB_RPPI_MOCKS) sep = sqrt(bin->edges[bin->nedges - 1] + bin->pimax * bin->pimax);  // 3-D
            const double cells = ceil(sep * lat->inv[k] * sub[k] * (1.0 + 1e-6)) + 1.0;
            if (cells < P.g.reach[k]) P.g.reach[k] = (int)cells;
        }
        P.m_s[k] = div_magic((unsigned)sub[k]);
        P.g.periodic[k] = lat->periodic[k];
        P.wrap[k] = lat->wrap[k] * scale;
        P.max_sep[k] = lat->max_sep[k];
        {
            const double hi_k = lat->lo[k] + (lat->inv[k] > 0 ? (double)lat->nmesh[k] / lat->inv[k] : 0.0);
            const double cm = (fmax(fabs(lat->lo[k]), fabs(hi_k)) * (1.0 + 1e-6) + fabs(lat->wrap[k])) * scale;
            if (k == 0 || cm > P.coord_max) P.coord_max = cm;
        }
    }
    P.tile_cell = (const int *)c.set[0].tile_cell.p;
    P.tile_off = (const int *)c.set[0].tile_off.p;
    P.ntiles = c.set[0].ntiles;
    {
        const unsigned wy = 2 * P.g.reach[1] + 1, wz = 2 * P.g.reach[2] + 1;
        P.m_wz = div_magic(wz);
        // the magic-number divisions are exact while n * d < 2^32
        bool ok = (double)wy * wz * wz < 4.0e9;
        for (int k = 0; k < 3; k++) ok = ok && 3.0 * P.g.ng[k] * sub[k] < 4.0e9;
        if (fast && !ok) return cfb_fail("lattice too large for the fast kernel's index arithmetic");
    }
    if (c.set[0].n > 0 && c.set[nsets - 1].n > 0 && P.ntiles > 0) {
        if (fast ? cfb_launch_pairs_fast(bin, P, bin->prec, false) : cfb_launch_pairs_generic(bin, P, bin->prec, false))
            return 1;
    }
    st.kernel_kind = fast ? 1 : c.last_kind;
    CK(cudaEventRecord(c.ev[2], c.stream));
    if (fetch_hist(c, nslots, bin, out, &st)) return 1;
    float ms = 0;
    cudaEventElapsedTime(&ms, c.ev[0], c.ev[1]);
    st.ms_gridlink = ms;
    cudaEventElapsedTime(&ms, c.ev[1], c.ev[2]);
    st.ms_pairs = ms;
    cudaEventElapsedTime(&ms, c.ev[0], c.ev[2]);
    st.ms_total_device = ms;
    st.n_cells = c.set[0].ncells;
    st.n_tiles = c.set[0].ntiles;
    for (int k = 0; k < 3; k++) st.fine[k] = lat->nmesh[k] * sub[k];
    st.kernel_launches = c.launches;
    if (stats) *stats = st;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Several GPUs inside ONE call (SURVEY 8(e) "process model"; the reference shards its cell pairs over OpenMP threads inside
// the call, theory/DD/countpairs_impl.c.src:452-530).  The public signatures have no slot for a device count, so it comes
// from the environment: CORRFUNC_B200_NGPUS = n uses n devices; unset, every visible device is used once a particle set
// holds CFB_MULTI_MIN_N points (below that a second device costs more in replication than it saves).  One process per GPU
// (torchrun: cfb_set_shard with nranks > 1, or CORRFUNC_B200_DEVICE set) always stays on its own device.
// Device 0 of the call holds the uploaded arrays; the others receive them by peer copies over NVLink (900 GB/s per
// direction instead of another trip over PCIe), then every device grids its replica and counts the primary cells it
// owns (cfb_owns_cell), driven by its own host thread; the partial histograms (<= a few KB) are summed on the host.
#define CFB_MULTI_MIN_N 4000000
static int g_multi_devs[CFB_MAX_DEV];
static int g_multi_last = 1;  // devices the last box count ran on

static int plan_devices(const Ctx &c0, int64_t nmax, int *devs)
{
    devs[0] = c0.dev;
    if (c0.shard_n > 1) return 1;
    const char *en = getenv("CORRFUNC_B200_NGPUS");
    const char *ed = getenv("CORRFUNC_B200_DEVICE");
    int want;
    if (en && *en) want = atoi(en);
    else if (ed && *ed) want = 1;
    else want = nmax >= CFB_MULTI_MIN_N ? CFB_MAX_DEV : 1;
    if (want <= 1) return 1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    int n = 1;
    for (int d = 0; d < ndev && n < want && n < CFB_MAX_DEV; d++)
        if (d != c0.dev) devs[n++] = d;
    return n;
}

// copies set `slot` of context 0 (device memory there) into context `c` of another device
static int replicate_set(Ctx &c, const Ctx &c0, int slot)
{
    const ParticleSet &S0 = c0.set[slot];
    if (g_cache_on && S0.src_valid) {  // the replica of an earlier call of this workflow
        if (!same_source(c.set[slot], S0.prec, S0.n, S0.src) && same_source(c.set[1 - slot], S0.prec, S0.n, S0.src))
            std::swap(c.set[0], c.set[1]);
        if (same_source(c.set[slot], S0.prec, S0.n, S0.src)) return 0;
    }
    ParticleSet &S = c.set[slot];
    S.prec = S0.prec;
    S.n = S0.n;
    S.gridded = false;
    S.grid_sig_valid = false;
    S.src_valid = g_cache_on && S0.src_valid;
    for (int i = 0; i < 6; i++) S.src[i] = S0.src[i];
    const size_t bytes = (size_t)S0.n * S0.prec;
    for (int i = 0; i < 6; i++) {
        S.raw[i] = nullptr;
        if (i >= 4 || !S0.raw[i] || S0.n == 0) continue;  // x, y, z, w: what the box statistics use
        if (cfb_ensure(S.rawbuf[i], bytes)) return 1;
        CK(cudaMemcpyPeerAsync(S.rawbuf[i].p, c.dev, S0.raw[i], c0.dev, bytes, c.stream));
        S.raw[i] = S.rawbuf[i].p;
    }
    return 0;
}

static int count_box_multi(const cfb_binning *bin, const cfb_box_lattice *lat, cfb_hist *out, cfb_stats *stats, int nd,
                           const int *devs)
{
    Ctx &c0 = g_ctxs[0];
    const int nsets = bin->autocorr ? 1 : 2;
    const int64_t ns = bin->nslots;
    CK(cudaSetDevice(c0.dev));
    CK(cudaEventRecord(c0.ev[7], c0.stream));  // the uploads of this call are complete when this event is
    std::vector<uint64_t> np((size_t)nd * ns, 0);
    std::vector<double> ss((size_t)nd * ns, 0.0), sw((size_t)nd * ns, 0.0);
    std::vector<cfb_stats> st(nd);
    std::vector<int> rc(nd, 0);
    auto work = [&](int k) {
        tl_ctx = &g_ctxs[k];
        Ctx &c = g_ctxs[k];
        c.err[0] = 0;
        if (k > 0) {
            if (init_ctx(c, devs[k])) { rc[k] = 1; return; }
            if (cudaSetDevice(c.dev) != cudaSuccess) { rc[k] = 1; return; }
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, c.dev, c0.dev) == cudaSuccess && can) {
                if (cudaDeviceEnablePeerAccess(c0.dev, 0) != cudaSuccess) cudaGetLastError();  // already enabled: fine
            }
            if (cudaStreamWaitEvent(c.stream, c0.ev[7], 0) != cudaSuccess) { rc[k] = 1; return; }
            for (int s = 0; s < nsets; s++)
                if (replicate_set(c, c0, s)) { rc[k] = 1; return; }
            c.target_occ = c0.target_occ;
            c.force_kernel = c0.force_kernel;
        }
        c.shard_rank = k;
        c.shard_n = nd;
        cfb_hist h = {np.data() + (size_t)k * ns, out->sum_sep ? ss.data() + (size_t)k * ns : nullptr,
                      out->sum_w ? sw.data() + (size_t)k * ns : nullptr};
        rc[k] = count_box_one(bin, lat, &h, &st[k]);
        c.shard_rank = 0;
        c.shard_n = 1;
    };
    std::vector<std::thread> th;
    for (int k = 1; k < nd; k++) th.emplace_back(work, k);
    work(0);
    for (auto &t : th) t.join();
    tl_ctx = &g_ctxs[0];
    cudaSetDevice(c0.dev);
    for (int k = 0; k < nd; k++)
        if (rc[k]) {
            if (k > 0) snprintf(c0.err, sizeof(c0.err), "device %d: %s", devs[k], g_ctxs[k].err);
            return 1;
        }
    cfb_stats tot = st[0];
    for (int64_t i = 0; i < ns; i++) {
        uint64_t n = 0;
        double a = 0.0, b = 0.0;
        for (int k = 0; k < nd; k++) {
            n += np[(size_t)k * ns + i];
            a += ss[(size_t)k * ns + i];
            b += sw[(size_t)k * ns + i];
        }
        out->npairs[i] = n;
        if (out->sum_sep) out->sum_sep[i] = a;
        if (out->sum_w) out->sum_w[i] = b;
    }
    for (int k = 1; k < nd; k++) {  // times: the slowest device; work counters: the sum
        if (st[k].ms_gridlink > tot.ms_gridlink) tot.ms_gridlink = st[k].ms_gridlink;
        if (st[k].ms_pairs > tot.ms_pairs) tot.ms_pairs = st[k].ms_pairs;
        if (st[k].ms_total_device > tot.ms_total_device) tot.ms_total_device = st[k].ms_total_device;
        tot.n_eval += st[k].n_eval;
        tot.n_tilepairs += st[k].n_tilepairs;
        tot.n_analytic += st[k].n_analytic;
        tot.n_levelpairs += st[k].n_levelpairs;
        tot.kernel_launches += st[k].kernel_launches;
    }
    if (stats) *stats = tot;
    return 0;
}

extern "C" int cfb_count_box(const cfb_binning *bin, const cfb_box_lattice *lat, cfb_hist *out, cfb_stats *stats)
{
    tl_ctx = &g_ctxs[0];
    Ctx &c0 = g_ctxs[0];
    if (!c0.ready) return cfb_fail("cfb_count_box before cfb_upload");
    int64_t nmax = c0.set[0].n;
    if (!bin->autocorr && c0.set[1].n > nmax) nmax = c0.set[1].n;
    const int nd = plan_devices(c0, nmax, g_multi_devs);
    g_multi_last = nd;
    if (nd <= 1) return count_box_one(bin, lat, out, stats);
    return count_box_multi(bin, lat, out, stats, nd, g_multi_devs);
}

/* devices the last cfb_count_box ran on */
extern "C" int cfb_last_device_count(void) { return g_multi_last; }

extern "C" int cfb_theta_subdivision(int64_t nmax, int64_t ncells)
{
    // sub x sub fine cells per reference cell, about `target` particles each; the fine lattice stays below ~4M cells
    const Ctx &c = g_ctx;
    const int target = c.target_occ > 0 ? c.target_occ : 112;
    if (ncells <= 0 || nmax <= 0) return 1;
    const double occ = (double)nmax / (double)ncells;
    if (occ <= 1.5 * target) return 1;
    int s = (int)floor(sqrt(occ / target) + 0.5);
    if (s < 1) s = 1;
    while (s > 1 && (double)ncells * s * s > 4.0e6) s--;
    return s;
}

extern "C" int cfb_theta_gridlink(int slot, int prec, const cfb_theta_lattice *lat, int64_t ncells, int64_t *counts,
                                  double *ra_bounds, double *xyz_bounds)
{
    Ctx &c = g_ctx;
    if (!c.ready) return cfb_fail("cfb_theta_gridlink before cfb_upload");
    CK(cudaSetDevice(c.dev));
    ParticleSet &S = c.set[slot];
    if (S.prec != prec) return cfb_fail("particle set %d precision mismatch", slot);
    if (cfb_gridlink_theta_set(S, lat, ncells)) return 1;
    const int sub = lat->sub > 0 ? lat->sub : 1;
    const int s2 = sub * sub;
    c.theta_sub = sub;
    const int64_t nfine = ncells * s2;
    // bring the per-cell counts and bounds back for the host-side neighbour search, folded from the fine cells
    // into the reference cells the search works on
    int *hc = (int *)malloc((size_t)nfine * sizeof(int));
    void *hb = malloc((size_t)nfine * CFB_NB * prec);
    if (!hc || !hb) return cfb_fail("out of host memory");
    CK(cudaMemcpyAsync(hc, S.count.p, (size_t)nfine * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    CK(cudaMemcpyAsync(hb, S.bounds.p, (size_t)nfine * CFB_NB * prec, cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    auto bnd = [&](int64_t f, int k) -> double {
        return prec == 4 ? (double)((float *)hb)[f * CFB_NB + k] : ((double *)hb)[f * CFB_NB + k];
    };
    for (int64_t i = 0; i < ncells; i++) {
        int64_t n = 0;
        bool first = true;
        for (int q = 0; q < s2; q++) {
            const int64_t f = i * s2 + q;
            if (hc[f] == 0 && !(q == s2 - 1 && first)) continue;  // an empty reference cell reports its last sub-cell
            n += hc[f];
            for (int k = 0; k < 8; k++) {
                const double v = bnd(f, k);
                double *dst = k < 6 ? &xyz_bounds[i * 6 + k] : &ra_bounds[i * 2 + (k - 6)];
                if (first) *dst = v;
                else if (k & 1) *dst = v > *dst ? v : *dst;  // odd slots hold maxima
                else *dst = v < *dst ? v : *dst;
            }
            first = false;
        }
        counts[i] = n;
    }
    free(hc);
    free(hb);
    return 0;
}

extern "C" int cfb_count_theta(const cfb_binning *bin, int64_t ncells, const int64_t *ngb_offsets,
                               const int32_t *ngb_cells, cfb_hist *out, cfb_stats *stats)
{
    Ctx &c = g_ctx;
    if (!c.ready) return cfb_fail("cfb_count_theta before cfb_upload");
    CK(cudaSetDevice(c.dev));
    c.launches = 0;
    cfb_stats st;
    memset(&st, 0, sizeof(st));
    CK(cudaEventRecord(c.ev[0], c.stream));
    const int64_t nslots = bin->nslots;
    if (alloc_hist(c, nslots)) return 1;
    // fast kernel: works on v = -cos(theta) (increasing edges); in float on v * 2^24, where every
    // value the kernel can produce is an integer, so edges are rounded up to the next integer
    bool fast = c.force_kernel != 0 && !bin->need_avg && !bin->need_weights && bin->nedges >= 2 &&
                bin->nedges <= CFB_FAST_MAX_EDGES;
    double edges_v[CFB_FAST_MAX_EDGES];
    if (fast) {
        for (int i = 0; i < bin->nedges; i++) {
            edges_v[i] = bin->prec == 4 ? ceil(-bin->edges[i] * 16777216.0) : -bin->edges[i];
            if (!isfinite(edges_v[i]) || (i > 0 && !(edges_v[i] >= edges_v[i - 1]))) fast = false;
        }
    }
    if (upload_edges(c, bin, fast ? edges_v : bin->edges)) return 1;
    const int64_t nlist = ngb_offsets[ncells];
    if (cfb_ensure(c.list_off, (size_t)(ncells + 1) * 8)) return 1;
    if (cfb_ensure(c.list_cells, (size_t)(nlist > 0 ? nlist : 1) * 4)) return 1;
    CK(cudaMemcpyAsync(c.list_off.p, ngb_offsets, (size_t)(ncells + 1) * 8, cudaMemcpyHostToDevice, c.stream));
    if (nlist > 0) CK(cudaMemcpyAsync(c.list_cells.p, ngb_cells, (size_t)nlist * 4, cudaMemcpyHostToDevice, c.stream));
    PairParams P;
    fill_common(P, c, bin, nslots);
    P.list_off = (const int64_t *)c.list_off.p;
    P.list_cells = (const int32_t *)c.list_cells.p;
    P.list_sub2 = c.theta_sub * c.theta_sub;
    {
        // pruning radius of fine-cell pairs: the largest chord in range, from the last edge cos(thetamax)
        const double cmax = bin->edges[bin->nedges - 1];
        P.max_sep[0] = sqrt(fmax(0.0, 2.0 * (1.0 - cmax)));
    }
    P.coord_max = 1.0 + 1e-6;  // unit vectors
    P.tile_cell = (const int *)c.set[0].tile_cell.p;
    P.tile_off = (const int *)c.set[0].tile_off.p;
    P.ntiles = c.set[0].ntiles;
    if (P.ntiles > 0 && nlist > 0) {
        if (fast ? cfb_launch_pairs_fast(bin, P, bin->prec, true) : cfb_launch_pairs_generic(bin, P, bin->prec, true))
            return 1;
    }
    st.kernel_kind = fast ? 1 : c.last_kind;
    CK(cudaEventRecord(c.ev[2], c.stream));
    if (fetch_hist(c, nslots, bin, out, &st)) return 1;
    float ms = 0;
    cudaEventElapsedTime(&ms, c.ev[0], c.ev[2]);
    st.ms_pairs = ms;
    st.ms_total_device = ms;
    st.n_cells = ncells;
    st.n_tiles = c.set[0].ntiles;
    st.kernel_launches = c.launches;
    if (stats) *stats = st;
    return 0;
}
