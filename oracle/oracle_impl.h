/* oracle_impl.h -- TEST INFRASTRUCTURE ONLY (never linked into / called from the product path).
 *
 * Plain-C, scalar CPU restatement of the reference's gridded pair-counting path, included twice
 * by pairs_oracle.c with REAL = float and REAL = double (the reference does the same with its
 * `DOUBLE` sed templates, rules.mk:23-49).  Every routine cites the reference lines it follows.
 * Pinned against: the reference itself compiled here (oracle/_ref, see build_ref.sh), the
 * reference's golden file mocks/tests/Mr19_mock_wtheta.DD, and the data-free known-answer tests
 * of Corrfunc/tests/test_theory.py (tests/test_oracle_*.py).
 *
 * What is restated bit-for-bit: bin edges and their squares, lattice geometry (nmesh, bin size,
 * inverse, truncating cell index with the ix-- clamp), the cell-pair set with periodic wrap,
 * first/second roles, duplicate suppression and the autocorr icell2<=icell filter, the per-pair
 * arithmetic in the AVX-512 kernels' FMA association, the 2-D bin index evaluated in floating
 * point, and the epilogues.  What is NOT restated by default: the z-sorted early exits inside a cell pair
 * (pruning aids; in double they never change a count).
 *
 * LITERAL mode (oracle_set_literal_kernels(1)): for wp, DDrppi, DD and xi the cells are z-sorted and the AVX-512
 * kernels' control flow is followed chunk by chunk (16 float / 8 double lanes) -- the fast-forward over
 * secondaries with z1 <= zpos - pimax and the "some lane reached pimax -> last chunk" exit.  In float a
 * secondary that survives the fast-forward can still round to dz == -pimax exactly; the reference then
 *   wp     : counts it, because its mask is the SIGNED dz < pimax   (wp_kernels.c.src:196-207)
 *   DDrppi : takes |dz| first, sees |dz| >= pimax, and stops after this chunk -- every later secondary
 *            of this primary is dropped                             (countpairs_rp_pi_kernels.c.src:196-207)
 *   DD, xi : never visits a secondary outside the per-primary window |dz| < max_dz (see
 *            o_count_cellpair_literal_dd): at most a pair on the very edge of the last bin
 * Literal mode reproduces these, which is what makes the full-size float goldens bit-identical
 * (tests/test_cpu_oracle.py::test_*literal*).  For wp / DDrppi the per-primary minimum-separation shortcut
 * (:130-137) is not restated: it has never changed a count.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

#if REAL_IS_DOUBLE
#define FMA_R(a, b, c) fma((a), (b), (c))
#define SQRT_R(a) sqrt(a)
#define FABS_R(a) fabs(a)
#define ACOS_R(a) acos(a)
#define ASIN_R(a) asin(a)
#define COSD_R(X) cos((X) * ORC_PI_OVER_180)
#define SIND_R(X) sin((X) * ORC_PI_OVER_180)
#define MAXPOS_R DBL_MAX
#else
#define FMA_R(a, b, c) fmaf((a), (b), (c))
#define SQRT_R(a) sqrtf(a)
#define FABS_R(a) fabsf(a)
#define ACOS_R(a) acosf(a)
#define ASIN_R(a) asinf(a)
#define COSD_R(X) cosf((X) * ORC_PI_OVER_180)
#define SIND_R(X) sinf((X) * ORC_PI_OVER_180)
#define MAXPOS_R FLT_MAX
#endif

typedef struct {
    int64_t n;     /* particles in this cell */
    int64_t start; /* offset into the cell-sorted arrays */
    REAL xb[2], yb[2], zb[2];
    REAL rab[2], decb[2]; /* theta lattice only */
} FN(ocell);

/* utils/gridlink_utils.c.src:31-49 */
static int FN(o_get_binsize)(const REAL xdiff, const REAL xwrap, const REAL rmax, const int refine_factor,
                            const int max_ncells, REAL *xbinsize, int *nlattice)
{
    int nmesh = (int)(refine_factor * xdiff / rmax);
    nmesh = nmesh < 1 ? 1 : nmesh;
    if (xwrap > 0 && rmax >= xwrap / 2) {
        fprintf(stderr, "oracle> rmax=%f must be less than half of periodic boxsize=%f\n", (double)rmax, (double)xwrap);
        return EXIT_FAILURE;
    }
    if (nmesh > max_ncells) nmesh = max_ncells;
    if (nmesh < 2) nmesh = 2;
    *xbinsize = xdiff / nmesh;
    *nlattice = nmesh;
    return EXIT_SUCCESS;
}

typedef struct {
    int nmesh[3];
    int64_t totncells;
    FN(ocell) * cells;
    REAL *x, *y, *z, *w; /* cell-sorted copies */
} FN(olattice);

static void FN(o_free_lattice)(FN(olattice) * L)
{
    if (!L) return;
    free(L->cells);
    free(L->x);
    free(L->y);
    free(L->z);
    free(L->w);
    free(L);
}

/* utils/gridlink_impl.c.src:398-420 (sort_cell_in_z): ascending z inside every cell.  Ties keep the input
 * order (the reference's quicksort leaves them unspecified; they only matter on a chunk boundary). */
typedef struct {
    REAL z;
    int64_t i;
} FN(ozkey);
static int FN(o_zkey_cmp)(const void *a, const void *b)
{
    const FN(ozkey) *p = a, *q = b;
    if (p->z < q->z) return -1;
    if (p->z > q->z) return 1;
    return (p->i > q->i) - (p->i < q->i);
}
static void FN(o_sort_cells_in_z)(FN(olattice) * L)
{
    int64_t nmax = 1;
    for (int64_t c = 0; c < L->totncells; c++)
        if (L->cells[c].n > nmax) nmax = L->cells[c].n;
    FN(ozkey) *key = malloc(sizeof(*key) * (size_t)nmax);
    REAL *tmp = malloc(sizeof(REAL) * (size_t)nmax);
    for (int64_t c = 0; c < L->totncells; c++) {
        const int64_t n = L->cells[c].n, s0 = L->cells[c].start;
        if (n < 2) continue;
        for (int64_t i = 0; i < n; i++) {
            key[i].z = L->z[s0 + i];
            key[i].i = i;
        }
        qsort(key, (size_t)n, sizeof(*key), FN(o_zkey_cmp));
        REAL *arr[4] = {L->x, L->y, L->z, L->w};
        for (int a = 0; a < 4; a++) {
            if (!arr[a]) continue;
            for (int64_t i = 0; i < n; i++) tmp[i] = arr[a][s0 + key[i].i];
            memcpy(arr[a] + s0, tmp, sizeof(REAL) * (size_t)n);
        }
    }
    free(key);
    free(tmp);
}

/* utils/gridlink_impl.c.src:65-436 (copy_particles=1 branch; per-cell z sort omitted: pruning aid) */
static FN(olattice) * FN(o_gridlink)(const int64_t N, const REAL *X, const REAL *Y, const REAL *Z, const REAL *W,
                                      const REAL xmin, const REAL xmax, const REAL ymin, const REAL ymax,
                                      const REAL zmin, const REAL zmax, const REAL max_x, const REAL max_y,
                                      const REAL max_z, const REAL xwrap, const REAL ywrap, const REAL zwrap,
                                      const int rfx, const int rfy, const int rfz, const int max_cells)
{
    REAL xbin = 0, ybin = 0, zbin = 0;
    int nx, ny, nz;
    if (FN(o_get_binsize)(xmax - xmin, xwrap, max_x, rfx, max_cells, &xbin, &nx) ||
        FN(o_get_binsize)(ymax - ymin, ywrap, max_y, rfy, max_cells, &ybin, &ny) ||
        FN(o_get_binsize)(zmax - zmin, zwrap, max_z, rfz, max_cells, &zbin, &nz))
        return NULL;
    const int64_t tot = (int64_t)nx * ny * nz;
    FN(olattice) *L = calloc(1, sizeof(*L));
    L->nmesh[0] = nx;
    L->nmesh[1] = ny;
    L->nmesh[2] = nz;
    L->totncells = tot;
    L->cells = calloc(tot, sizeof(*L->cells));
    L->x = malloc(sizeof(REAL) * (N > 0 ? N : 1));
    L->y = malloc(sizeof(REAL) * (N > 0 ? N : 1));
    L->z = malloc(sizeof(REAL) * (N > 0 ? N : 1));
    L->w = W ? malloc(sizeof(REAL) * (N > 0 ? N : 1)) : NULL;
    int64_t *idx = malloc(sizeof(int64_t) * (N > 0 ? N : 1));
    /* gridlink_impl.c.src:156-158 */
    const REAL xinv = xbin > 0 ? 1.0 / xbin : 0.;
    const REAL yinv = ybin > 0 ? 1.0 / ybin : 0.;
    const REAL zinv = zbin > 0 ? 1.0 / zbin : 0.;
    int64_t oob = 0;
    for (int64_t i = 0; i < N; i++) { /* gridlink_impl.c.src:165-181 */
        int ix = (int)((X[i] - xmin) * xinv);
        int iy = (int)((Y[i] - ymin) * yinv);
        int iz = (int)((Z[i] - zmin) * zinv);
        if (ix > nx - 1) ix--;
        if (iy > ny - 1) iy--;
        if (iz > nz - 1) iz--;
        oob += ix < 0 || ix >= nx || iy < 0 || iy >= ny || iz < 0 || iz >= nz;
        idx[i] = (int64_t)ix * ny * nz + (int64_t)iy * nz + iz;
        if (oob == 0) L->cells[idx[i]].n++;
    }
    if (oob) {
        fprintf(stderr, "oracle> %" PRId64 " particles are out of bounds\n", oob);
        free(idx);
        FN(o_free_lattice)(L);
        return NULL;
    }
    int64_t off = 0;
    for (int64_t c = 0; c < tot; c++) {
        L->cells[c].start = off;
        off += L->cells[c].n;
        L->cells[c].n = 0;
        L->cells[c].xb[0] = L->cells[c].yb[0] = L->cells[c].zb[0] = MAXPOS_R;
        L->cells[c].xb[1] = L->cells[c].yb[1] = L->cells[c].zb[1] = -MAXPOS_R;
    }
    for (int64_t i = 0; i < N; i++) { /* gridlink_impl.c.src:289-315 */
        FN(ocell) *c = &L->cells[idx[i]];
        const int64_t p = c->start + c->n++;
        L->x[p] = X[i];
        L->y[p] = Y[i];
        L->z[p] = Z[i];
        if (W) L->w[p] = W[i];
        if (X[i] < c->xb[0]) c->xb[0] = X[i];
        if (Y[i] < c->yb[0]) c->yb[0] = Y[i];
        if (Z[i] < c->zb[0]) c->zb[0] = Z[i];
        if (X[i] > c->xb[1]) c->xb[1] = X[i];
        if (Y[i] > c->yb[1]) c->yb[1] = Y[i];
        if (Z[i] > c->zb[1]) c->zb[1] = Z[i];
    }
    free(idx);
    if (orc_literal_kernels) FN(o_sort_cells_in_z)(L);
    return L;
}

typedef struct {
    int64_t c1, c2;
    REAL xw, yw, zw;
    int same;
    /* struct cell_pair_DOUBLE's pruning fields (utils/cell_pair.h.src:20-36), used by the literal kernels only */
    REAL min_dx, min_dy, min_dz, closest_x1, closest_y1, closest_z1;
} FN(opair);

/* utils/gridlink_impl.c.src:439-625 */
static FN(opair) * FN(o_cell_pairs)(const FN(olattice) * L1, const FN(olattice) * L2, int64_t *npairs_out,
                                     const int rfx, const int rfy, const int rfz, const REAL xwrap,
                                     const REAL ywrap, const REAL zwrap, const REAL max_3D, const REAL max_2D,
                                     const REAL max_1D, const int enable_min_sep, const int autocorr, const int px,
                                     const int py, const int pz)
{
    const int nx = L1->nmesh[0], ny = L1->nmesh[1], nz = L1->nmesh[2];
    const int64_t tot = L1->totncells;
    const int64_t max_ngb = (int64_t)(2 * rfx + 1) * (2 * rfy + 1) * (2 * rfz + 1);
    FN(opair) *P = malloc(sizeof(*P) * (size_t)(tot * max_ngb > 0 ? tot * max_ngb : 1));
    int64_t np = 0;
    const int any_periodic = px || py || pz;
    const int check_dup = (any_periodic && (nx < 2 * rfx + 1 || ny < 2 * rfy + 1 || nz < 2 * rfz + 1)) ? 1 : 0;
    for (int64_t icell = 0; icell < tot; icell++) {
        const FN(ocell) *first = &L1->cells[icell];
        if (first->n == 0) continue;
        const int iz = icell % nz;
        const int ix = icell / ((int64_t)ny * nz);
        const int iy = (icell - iz - (int64_t)ix * nz * ny) / nz;
        int64_t nthis = 0;
        for (int iix = -rfx; iix <= rfx; iix++) {
            const int pix = (ix + iix + nx) % nx;
            const int iiix = px ? pix : ix + iix;
            if (iiix < 0 || iiix >= nx) continue;
            const REAL offx = ((ix + iix) >= 0) && ((ix + iix) < nx) ? 0.0 : ((ix + iix) < 0 ? xwrap : -xwrap);
            for (int iiy = -rfy; iiy <= rfy; iiy++) {
                const int piy = (iy + iiy + ny) % ny;
                const int iiiy = py ? piy : iy + iiy;
                if (iiiy < 0 || iiiy >= ny) continue;
                const REAL offy = ((iy + iiy) >= 0) && ((iy + iiy) < ny) ? 0.0 : ((iy + iiy) < 0 ? ywrap : -ywrap);
                for (int iiz = -rfz; iiz <= rfz; iiz++) {
                    const int piz = (iz + iiz + nz) % nz;
                    const int iiiz = pz ? piz : iz + iiz;
                    if (iiiz < 0 || iiiz >= nz) continue;
                    const REAL offz = ((iz + iiz) >= 0) && ((iz + iiz) < nz) ? 0.0 : ((iz + iiz) < 0 ? zwrap : -zwrap);
                    const int64_t icell2 = iiiz + (int64_t)nz * iiiy + (int64_t)nz * ny * iiix;
                    if ((autocorr == 1 && icell2 > icell) || L2->cells[icell2].n == 0) continue;
                    if (check_dup) { /* gridlink_utils.h.src:46-72 */
                        int dup = 0;
                        for (int64_t jj = 0; jj < nthis; jj++) {
                            const FN(opair) *q = &P[np - jj - 1];
                            if (q->c2 == icell2 && q->xw == offx && q->yw == offy && q->zw == offz) {
                                dup = 1;
                                break;
                            }
                        }
                        if (dup) continue;
                    }
                    const FN(ocell) *second = &L2->cells[icell2];
                    REAL cp_min[3] = {0, 0, 0}, cp_closest[3] = {0, 0, 0};
                    if (enable_min_sep) { /* gridlink_impl.c.src:538-589 */
                        const REAL x_low = first->xb[0] + offx, x_hi = first->xb[1] + offx;
                        const REAL y_low = first->yb[0] + offy, y_hi = first->yb[1] + offy;
                        const REAL z_low = first->zb[0] + offz, z_hi = first->zb[1] + offz;
                        const REAL first_x = iix < 0 ? x_low : x_hi;
                        const REAL second_x = iix < 0 ? second->xb[1] : second->xb[0];
                        const REAL min_dx = iix != 0 ? (first_x - second_x) : 0;
                        const REAL first_y = iiy < 0 ? y_low : y_hi;
                        const REAL second_y = iiy < 0 ? second->yb[1] : second->yb[0];
                        const REAL min_dy = iiy != 0 ? (first_y - second_y) : 0;
                        const REAL first_z = iiz < 0 ? z_low : z_hi;
                        const REAL second_z = iiz < 0 ? second->zb[1] : second->zb[0];
                        const REAL min_dz = iiz != 0 ? (first_z - second_z) : 0;
                        cp_min[0] = min_dx, cp_min[1] = min_dy, cp_min[2] = min_dz;
                        cp_closest[0] = iix < 0 ? x_low : (iix > 0 ? x_hi : 0); /* :543-545 */
                        cp_closest[1] = iiy < 0 ? y_low : (iiy > 0 ? y_hi : 0);
                        cp_closest[2] = iiz < 0 ? z_low : (iiz > 0 ? z_hi : 0);
                        if (max_3D > 0) {
                            const REAL s = min_dx * min_dx + min_dy * min_dy + min_dz * min_dz;
                            if (s >= max_3D * max_3D) continue;
                        }
                        if (max_2D > 0) {
                            const REAL s = min_dx * min_dx + min_dy * min_dy;
                            if (s >= max_2D * max_2D) continue;
                        }
                        if (max_1D > 0 && iiz != 0) {
                            const REAL s = min_dz * min_dz;
                            if (s >= max_1D * max_1D) continue;
                        }
                    }
                    P[np].c1 = icell;
                    P[np].c2 = icell2;
                    P[np].xw = offx;
                    P[np].yw = offy;
                    P[np].zw = offz;
                    P[np].same = (autocorr == 1 && icell2 == icell) ? 1 : 0;
                    P[np].min_dx = cp_min[0], P[np].min_dy = cp_min[1], P[np].min_dz = cp_min[2];
                    P[np].closest_x1 = cp_closest[0], P[np].closest_y1 = cp_closest[1], P[np].closest_z1 = cp_closest[2];
                    np++;
                    nthis++;
                }
            }
        }
    }
    *npairs_out = np;
    return P;
}

/* The per-pair arithmetic of the AVX-512 kernels, for one cell pair.
 *   DD / xi : theory/DD/countpairs_kernels.c.src:178-254   r2 = fma(dz,dz, fma(dy,dy, dx*dx))
 *   wp      : theory/wp/wp_kernels.c.src:186-262           r2 = fma(dy,dy, dx*dx), -pimax < dz < pimax
 *   DDrppi  : theory/DDrppi/countpairs_rp_pi_kernels.c.src:191-267
 *   DDsmu   : theory/DDsmu/countpairs_s_mu_kernels.c.src:207-290   s2 = fma(dx,dx, fma(dy,dy, dz*dz))
 *   DDtheta : mocks/DDtheta_mocks/countpairs_theta_mocks_kernels.c.src:1052-1128
 */
typedef struct {
    int mode;
    int nbin;             /* number of edges (= reference's nbin) */
    const REAL *edges;    /* rupp_sqr[] (theory) or costheta_upp[] (theta) */
    REAL pimax;           /* wp / rppi */
    int npibin;           /* rppi */
    REAL inv_dpi;         /* rppi */
    REAL sqr_mumax;       /* smu */
    REAL sqr_max_sep, sqr_pimax; /* rppi mocks: rp_pi_mocks_kernels:56-57 */
    int nmu;              /* smu */
    REAL inv_dmu;         /* smu */
    int need_avg, need_w; /* accumulate sums? */
    int fast_acos;
    uint64_t *npairs;
    double *avg;  /* reference accumulates in REAL; the oracle keeps double sums (order-robust) */
    double *wavg;
} FN(okern);

static inline REAL FN(o_fast_acos)(const REAL x)
{ /* utils/fast_acos.h:57-101 (degree-8 estimate) */
    const REAL xa = FABS_R(x);
    const REAL one = (REAL)1.0;
    REAL poly = (REAL) + 7.1796493341480527e-04;
    poly = (REAL)-4.1160981058965262e-03 + poly * xa;
    poly = (REAL) + 1.1272900916992512e-02 + poly * xa;
    poly = (REAL)-2.0949278766238422e-02 + poly * xa;
    poly = (REAL) + 3.2683762943179318e-02 + poly * xa;
    poly = (REAL)-5.0625279962389413e-02 + poly * xa;
    poly = (REAL) + 8.9034700107934128e-02 + poly * xa;
    poly = (REAL)-2.1460143648688035e-01 + poly * xa;
    poly = (REAL) + 1.5707963267948966 + poly * xa;
    poly = poly * SQRT_R(one - xa);
    return (x < 0) ? (REAL)(M_PI - poly) : poly;
}

static void FN(o_count_cellpair)(const FN(okern) * K, const int64_t N0, const REAL *x0, const REAL *y0,
                                 const REAL *z0, const REAL *w0, const int64_t N1, const REAL *x1, const REAL *y1,
                                 const REAL *z1, const REAL *w1, const int same, const REAL offx, const REAL offy,
                                 const REAL offz)
{
    const int nbin = K->nbin;
    const REAL *E = K->edges;
    for (int64_t i = 0; i < N0; i++) {
        const REAL xpos = x0[i] + offx, ypos = y0[i] + offy, zpos = z0[i] + offz;
        for (int64_t j = same ? i + 1 : 0; j < N1; j++) {
            const REAL dx = x1[j] - xpos, dy = y1[j] - ypos, dz = z1[j] - zpos;
            int64_t slot = -1;
            REAL sep = 0;
            switch (K->mode) {
            case ORC_DD:
            case ORC_XI: {
                const REAL r2 = FMA_R(dz, dz, FMA_R(dy, dy, dx * dx));
                if (!(r2 < E[nbin - 1] && r2 >= E[0])) continue;
                int k;
                for (k = nbin - 1; k >= 1; k--)
                    if (r2 >= E[k - 1]) break;
                slot = k;
                if (K->need_avg) sep = SQRT_R(r2);
            } break;
            case ORC_WP: {
                const REAL r2 = FMA_R(dy, dy, dx * dx);
                /* two cells: fast-forward over z1 <= zpos - pimax (:139-142), then the SIGNED mask dz < pimax
                 * (:207-221) -- a survivor whose dz rounds to exactly -pimax counts.  Same cell: j follows i in
                 * z order, dz >= 0, written symmetrically here because the cells are not z-sorted. */
                if (same ? !(dz > -K->pimax && dz < K->pimax) : !(z1[j] > zpos - K->pimax && dz < K->pimax)) continue;
                if (!(r2 < E[nbin - 1] && r2 >= E[0])) continue;
                int k;
                for (k = nbin - 1; k >= 1; k--)
                    if (r2 >= E[k - 1]) break;
                slot = k;
                if (K->need_avg) sep = SQRT_R(r2);
            } break;
            case ORC_RPPI: {
                const REAL r2 = FMA_R(dy, dy, dx * dx);
                /* two cells: fast-forward over z1 <= zpos - pimax (rp_pi_kernels:139-142), then |dz| < pimax */
                if (same ? !(dz > -K->pimax) : !(z1[j] > zpos - K->pimax)) continue;
                const REAL adz = FABS_R(dz);
                if (!(adz < K->pimax)) continue;
                if (!(r2 < E[nbin - 1] && r2 >= E[0])) continue;
                int k;
                for (k = nbin - 1; k >= 1; k--)
                    if (r2 >= E[k - 1]) break;
                /* finalbin evaluated in REAL floating point, then truncated (rp_pi_kernels:249-256) */
                const REAL pibin = adz * K->inv_dpi;
                const REAL lin = (REAL)k * (REAL)(K->npibin + 1);
                slot = (int64_t)(int)(lin + pibin);
                if (K->need_avg) sep = SQRT_R(r2);
            } break;
            case ORC_SMU: {
                const REAL sqr_dz = dz * dz;
                const REAL s2 = FMA_R(dx, dx, FMA_R(dy, dy, sqr_dz));
                const REAL max_sqr_dz = s2 * K->sqr_mumax;
                if (!(sqr_dz < max_sqr_dz)) continue;
                if (!(s2 < E[nbin - 1] && s2 >= E[0])) continue;
                const REAL sqr_mu = sqr_dz / s2; /* fast_divide_and_NR_steps == 0: true divide */
                const REAL mu = SQRT_R(sqr_mu);
                int k;
                for (k = nbin - 1; k >= 1; k--)
                    if (s2 >= E[k - 1]) break;
                const REAL mubin = mu * K->inv_dmu;
                const REAL lin = (REAL)k * (REAL)(K->nmu + 1);
                slot = (int64_t)(int)(lin + mubin);
                if (K->need_avg) sep = SQRT_R(s2);
            } break;
            case ORC_RPPI_MOCKS: { /* countpairs_rp_pi_mocks_kernels.c.src:200-300: line of sight = pair midpoint */
                const REAL parx = x1[j] + xpos, pary = y1[j] + ypos, parz = z1[j] + zpos;
                const REAL term1 = parx * dx, term2 = pary * dy;
                const REAL s_dot_l = FMA_R(parz, dz, term1 + term2);
                const REAL sqr_s_dot_l = s_dot_l * s_dot_l;
                const REAL sqr_sep = FMA_R(dx, dx, FMA_R(dy, dy, dz * dz));
                if (!(sqr_sep < K->sqr_max_sep)) continue;
                const REAL sqr_norm_l = FMA_R(parx, parx, FMA_R(pary, pary, parz * parz));
                if (!(sqr_s_dot_l < K->sqr_pimax * sqr_norm_l)) continue;
                const REAL sqr_Dpar = sqr_s_dot_l / sqr_norm_l; /* fast_divide_and_NR_steps == 0 */
                const REAL sqr_Dperp = sqr_sep - sqr_Dpar;
                if (!(sqr_Dpar < K->sqr_pimax && sqr_Dperp < E[nbin - 1] && sqr_Dperp >= E[0])) continue;
                const REAL Dpar = SQRT_R(sqr_Dpar);
                int k;
                for (k = nbin - 1; k >= 1; k--)
                    if (sqr_Dperp >= E[k - 1]) break;
                const REAL pibin = Dpar * K->inv_dpi;
                const REAL lin = (REAL)k * (REAL)(K->npibin + 1);
                slot = (int64_t)(int)(lin + pibin);
                if (K->need_avg) sep = SQRT_R(sqr_Dperp);
            } break;
            case ORC_SMU_MOCKS: { /* countpairs_s_mu_mocks_kernels.c.src:196-290 */
                const REAL parx = x1[j] + xpos, pary = y1[j] + ypos, parz = z1[j] + zpos;
                const REAL term1 = parx * dx, term2 = pary * dy;
                const REAL s_dot_l = FMA_R(parz, dz, term1 + term2);
                const REAL sqr_s_dot_l = s_dot_l * s_dot_l;
                const REAL sqr_s = FMA_R(dx, dx, FMA_R(dy, dy, dz * dz));
                if (!(sqr_s < E[nbin - 1])) continue;
                const REAL sqr_norm_l = FMA_R(parx, parx, FMA_R(pary, pary, parz * parz));
                const REAL sqr_mu = sqr_s_dot_l / (sqr_norm_l * sqr_s);
                const REAL mu = SQRT_R(sqr_mu);
                if (!(sqr_mu < K->sqr_mumax && sqr_s >= E[0])) continue;
                int k;
                for (k = nbin - 1; k >= 1; k--)
                    if (sqr_s >= E[k - 1]) break;
                const REAL mubin = mu * K->inv_dmu;
                const REAL lin = (REAL)k * (REAL)(K->nmu + 1);
                slot = (int64_t)(int)(lin + mubin);
                if (K->need_avg) sep = SQRT_R(sqr_s);
            } break;
            case ORC_THETA: {
                const REAL chord2 = FMA_R(dz, dz, FMA_R(dy, dy, dx * dx));
                const REAL costheta = (REAL)1.0 - (REAL)0.5 * chord2;
                /* costhetamax < costheta <= costhetamin ; E[] = cos(theta_upp[]) is decreasing */
                if (!(costheta > E[nbin - 1] && costheta <= E[0])) continue;
                int k;
                for (k = nbin - 1; k >= 1; k--)
                    if (costheta <= E[k - 1]) break;
                slot = k;
                if (K->need_avg) {
                    const REAL c = costheta >= (REAL)1.0 ? (REAL)1.0 : costheta; /* avx512_calls.h:333-338 */
                    const REAL th = K->fast_acos ? FN(o_fast_acos)(c) : ACOS_R(c);
                    sep = (REAL)(th * (REAL)ORC_INV_PI_OVER_180);
                }
            } break;
            default: continue;
            }
            K->npairs[slot]++;
            if (K->need_avg) K->avg[slot] += (double)sep;
            if (K->need_w) K->wavg[slot] += (double)(REAL)(w0[i] * w1[j]); /* weight_functions.h.src:71-73 */
        }
    }
}

/* LITERAL mode, wp and DDrppi: the control flow of wp_avx512_intrinsics (wp_kernels.c.src:76-262) and
 * countpairs_rp_pi_avx512_intrinsics (countpairs_rp_pi_kernels.c.src:81-267) over z-sorted cells. */
static void FN(o_count_cellpair_literal)(const FN(okern) * K, const int64_t N0, const REAL *x0, const REAL *y0,
                                         const REAL *z0, const REAL *w0, const int64_t N1, const REAL *x1,
                                         const REAL *y1, const REAL *z1, const REAL *w1, const int same,
                                         const REAL offx, const REAL offy, const REAL offz)
{
    const int nbin = K->nbin;
    const REAL *E = K->edges;
    const REAL pimax = K->pimax;
    const int NVEC = (int)(64 / sizeof(REAL));
    const int rppi = K->mode == ORC_RPPI;
    int64_t p = 0; /* the reference's z1 pointer: only ever moves forward */
    if (N1 == 0) return;
    for (int64_t i = 0; i < N0; i++) {
        const REAL xpos = x0[i] + offx, ypos = y0[i] + offy, zpos = z0[i] + offz;
        const REAL this_dz = z1[p] - zpos;
        if (this_dz >= pimax) continue;
        if (same) {
            p++;
        } else {
            const REAL target_z = zpos - pimax;
            while (p < N1 && z1[p] <= target_z) p++;
        }
        if (p == N1) break;
        for (int64_t j = p; j < N1; j += NVEC) {
            const int lanes = (N1 - j) >= NVEC ? NVEC : (int)(N1 - j);
            REAL dz[16];
            int any_gt_neg = 0, any_geq = 0, any_lt = 0;
            for (int l = 0; l < lanes; l++) {
                dz[l] = z1[j + l] - zpos;
                any_gt_neg |= dz[l] > -pimax;
            }
            if (!any_gt_neg) continue;
            for (int l = 0; l < lanes; l++) {
                if (rppi) dz[l] = FABS_R(dz[l]);
                any_geq |= dz[l] >= pimax;
                any_lt |= dz[l] < pimax;
            }
            const int64_t jbase = j;
            if (any_geq) j = N1; /* "do not break yet": this chunk is still processed */
            if (!any_lt) break;
            for (int l = 0; l < lanes; l++) {
                if (!(dz[l] < pimax)) continue;
                const int64_t jj = jbase + l;
                const REAL dx = x1[jj] - xpos, dy = y1[jj] - ypos;
                const REAL r2 = FMA_R(dy, dy, dx * dx);
                if (!(r2 < E[nbin - 1] && r2 >= E[0])) continue;
                int k;
                for (k = nbin - 1; k >= 1; k--)
                    if (r2 >= E[k - 1]) break;
                int64_t slot = k;
                if (rppi) {
                    const REAL pibin = dz[l] * K->inv_dpi;
                    const REAL lin = (REAL)k * (REAL)(K->npibin + 1);
                    slot = (int64_t)(int)(lin + pibin);
                }
                K->npairs[slot]++;
                if (K->need_avg) K->avg[slot] += (double)SQRT_R(r2);
                if (K->need_w) K->wavg[slot] += (double)(REAL)(w0[i] * w1[jj]);
            }
        }
    }
}

/* LITERAL mode, DD and xi: countpairs_avx512_intrinsics (theory/DD/countpairs_kernels.c.src:76-254) and
 * xi_avx512_intrinsics (theory/xi/xi_kernels.c.src:76-250) over z-sorted cells: the per-primary window
 * |dz| < max_dz = sqrt(rmax^2 - min_dx^2 - min_dy^2) built from the cell pair's bounding-box separations, the
 * fast-forward below the window and the "some lane reached max_dz -> last chunk" exit.  Inside a processed chunk
 * only r2 is tested.  In float, on coordinates much larger than rmax, the window can close a hair too early and a
 * pair with r2 < rmax^2 at the very edge of the last bin is never visited. */
static void FN(o_count_cellpair_literal_dd)(const FN(okern) * K, const FN(opair) * cp, const REAL rpmax, const int64_t N0,
                                            const REAL *x0, const REAL *y0, const REAL *z0, const REAL *w0,
                                            const int64_t N1, const REAL *x1, const REAL *y1, const REAL *z1,
                                            const REAL *w1)
{
    const int nbin = K->nbin;
    const REAL *E = K->edges;
    const REAL sqr_rpmax = E[nbin - 1], sqr_rpmin = E[0];
    const int NVEC = (int)(64 / sizeof(REAL));
    const REAL min_xdiff = cp->min_dx, min_ydiff = cp->min_dy, min_zdiff = cp->min_dz;
    /* DD squares the REAL rpmax (:77), xi takes the squared edge (xi_kernels:77) */
    const REAL max_all_dz = (K->mode == ORC_DD) ? SQRT_R(rpmax * rpmax - min_xdiff * min_xdiff - min_ydiff * min_ydiff)
                                                : SQRT_R(sqr_rpmax - min_xdiff * min_xdiff - min_ydiff * min_ydiff);
    int64_t p = 0;
    if (N1 == 0) return;
    for (int64_t i = 0; i < N0; i++) {
        const REAL xpos = x0[i] + cp->xw, ypos = y0[i] + cp->yw, zpos = z0[i] + cp->zw;
        REAL max_dz = max_all_dz;
        const REAL this_dz = z1[p] - zpos;
        if (this_dz >= max_all_dz) continue;
        if (cp->same) {
            p++;
        } else {
            const REAL min_dx = min_xdiff > 0 ? min_xdiff + FABS_R(xpos - cp->closest_x1) : min_xdiff;
            const REAL min_dy = min_ydiff > 0 ? min_ydiff + FABS_R(ypos - cp->closest_y1) : min_ydiff;
            const REAL min_dz = min_zdiff > 0 ? (this_dz > 0 ? this_dz : min_zdiff + FABS_R(zpos - cp->closest_z1)) : min_zdiff;
            const REAL sqr_min_sep_this_point = min_dx * min_dx + min_dy * min_dy + min_dz * min_dz;
            if (sqr_min_sep_this_point >= sqr_rpmax) continue;
            max_dz = SQRT_R(sqr_rpmax - min_dx * min_dx - min_dy * min_dy);
            const REAL target_z = zpos - max_all_dz;
            while (p < N1 && z1[p] <= target_z) p++;
        }
        if (p == N1) break;
        int64_t lp = p;
        const REAL target_z = zpos - max_dz;
        while (lp != N1 && z1[lp] <= target_z) lp++;
        for (int64_t j = lp; j < N1; j += NVEC) {
            const int lanes = (N1 - j) >= NVEC ? NVEC : (int)(N1 - j);
            const int64_t jbase = j;
            for (int l = 0; l < lanes; l++)
                if (z1[jbase + l] - zpos >= max_dz) j = N1; /* "do not break yet": this chunk is still processed */
            for (int l = 0; l < lanes; l++) {
                const int64_t jj = jbase + l;
                const REAL dx = x1[jj] - xpos, dy = y1[jj] - ypos, dz = z1[jj] - zpos;
                const REAL r2 = FMA_R(dz, dz, FMA_R(dy, dy, dx * dx));
                if (!(r2 < sqr_rpmax && r2 >= sqr_rpmin)) continue;
                int k;
                for (k = nbin - 1; k >= 1; k--)
                    if (r2 >= E[k - 1]) break;
                K->npairs[k]++;
                if (K->need_avg) K->avg[k] += (double)SQRT_R(r2);
                if (K->need_w) K->wavg[k] += (double)(REAL)(w0[i] * w1[jj]);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* theory entry point: DD / xi / wp / DDrppi / DDsmu                                            */
/* Follows theory/DD/countpairs_impl.c.src:136-707 and its four siblings.                       */
int FN(oracle_theory)(const int mode, const int64_t ND1, const REAL *X1, const REAL *Y1, const REAL *Z1,
                      const REAL *W1, const int64_t ND2, const REAL *X2, const REAL *Y2, const REAL *Z2,
                      const REAL *W2, int autocorr, const int nbin, const double *rupp, /* nbin edges */
                      const double pimax_in, const double mu_max_in, const int nmu_bins, int periodic,
                      const double boxsize_x, const double boxsize_y_in, const double boxsize_z_in,
                      const int *refine_in, const int binning_cust, int max_cells, const int enable_min_sep,
                      const int need_avg, const int need_w, uint64_t *npairs_out, double *avg_out,
                      double *wavg_out, double *cf_out, int *lattice_out /* nmesh[3], refine[3] */)
{
    int rf[3] = {refine_in[0], refine_in[1], refine_in[2]};
    if (max_cells == 0) max_cells = 100;
    for (int i = 0; i < 3; i++)
        if (rf[i] < 1) {
            rf[0] = 2;
            rf[1] = 2;
            rf[2] = 1;
            break;
        }
    const double rmin = rupp[0], rmax = rupp[nbin - 1];
    if (!(rmin >= 0.0 && rmax > 0.0 && rmin < rmax && nbin > 0)) return EXIT_FAILURE;
    REAL *esq = malloc(sizeof(REAL) * nbin);
    for (int i = 0; i < nbin; i++) esq[i] = rupp[i] * rupp[i]; /* double product rounded to REAL (DD impl:435-438) */

    const int is_box = (mode == ORC_XI || mode == ORC_WP);
    const int is_mocks = (mode == ORC_RPPI_MOCKS || mode == ORC_SMU_MOCKS);
    REAL *conv[6] = {NULL, NULL, NULL, NULL, NULL, NULL};
    if (is_mocks) {
        /* inputs are RA, DEC (degrees) and the comoving distance (is_comoving_dist = 1):
         * countpairs_rp_pi_mocks_impl.c.src:364-391 */
        if (!(rmin > 0.0)) { /* :304 -- the mocks statistics do not accept rmin = 0 */
            free(esq);
            return EXIT_FAILURE;
        }
        for (int sidx = 0; sidx < (autocorr ? 1 : 2); sidx++) {
            const int64_t n = sidx ? ND2 : ND1;
            const REAL *ra = sidx ? X2 : X1, *dec = sidx ? Y2 : Y1, *D = sidx ? Z2 : Z1;
            REAL *x = malloc(sizeof(REAL) * (n > 0 ? n : 1)), *y = malloc(sizeof(REAL) * (n > 0 ? n : 1)),
                 *z = malloc(sizeof(REAL) * (n > 0 ? n : 1));
            for (int64_t i = 0; i < n; i++) {
                x[i] = D[i] * COSD_R(dec[i]) * COSD_R(ra[i]);
                y[i] = D[i] * COSD_R(dec[i]) * SIND_R(ra[i]);
                z[i] = D[i] * SIND_R(dec[i]);
            }
            conv[3 * sidx] = x, conv[3 * sidx + 1] = y, conv[3 * sidx + 2] = z;
        }
        X1 = conv[0], Y1 = conv[1], Z1 = conv[2];
        if (!autocorr) X2 = conv[3], Y2 = conv[4], Z2 = conv[5];
        periodic = 0;
    }
    REAL xmin, xmax, ymin, ymax, zmin, zmax, xwrap, ywrap, zwrap;
    int px, py, pz;
    REAL max_x, max_y, max_z; /* gridlink cell sizes */
    REAL max3 = -1, max2 = -1, max1 = -1;
    REAL pimax = 0;
    int npibin = 0;
    REAL mu_max = 0;
    if (is_box) { /* xi impl:176-231, wp impl:191-260 */
        periodic = 1;
        autocorr = 1;
        xmin = ymin = zmin = 0.0;
        xmax = ymax = zmax = boxsize_x;
        xwrap = ywrap = zwrap = boxsize_x;
        px = py = pz = 1;
        if (mode == ORC_XI) {
            if (!binning_cust && rmax < 0.05 * boxsize_x) rf[0] = rf[1] = rf[2] = 1;
            max_x = max_y = max_z = rmax;
            max3 = rmax;
        } else {
            pimax = pimax_in;
            if (!binning_cust) {
                if (rmax < 0.05 * boxsize_x) rf[0] = rf[1] = 1;
                if (pimax_in < 0.05 * boxsize_x) rf[2] = 1;
            }
            max_x = max_y = rmax;
            max_z = pimax_in;
            max2 = rmax;
            max1 = pimax_in;
        }
    } else {
        xmin = ymin = zmin = MAXPOS_R;
        xmax = ymax = zmax = -MAXPOS_R;
        for (int s = 0; s < (autocorr ? 1 : 2); s++) { /* gridlink_utils.c.src:52-70 */
            const int64_t n = s ? ND2 : ND1;
            const REAL *x = s ? X2 : X1, *y = s ? Y2 : Y1, *z = s ? Z2 : Z1;
            for (int64_t i = 0; i < n; i++) {
                if (x[i] < xmin) xmin = x[i];
                if (y[i] < ymin) ymin = y[i];
                if (z[i] < zmin) zmin = z[i];
                if (x[i] > xmax) xmax = x[i];
                if (y[i] > ymax) ymax = y[i];
                if (z[i] > zmax) zmax = z[i];
            }
        }
        if (periodic && boxsize_x == -2.) return EXIT_FAILURE;
        const double bsy = boxsize_y_in == -2. ? boxsize_x : boxsize_y_in;
        const double bsz = boxsize_z_in == -2. ? boxsize_x : boxsize_z_in;
        px = periodic && boxsize_x >= 0;
        py = periodic && bsy >= 0;
        pz = periodic && bsz >= 0;
        xwrap = px ? (boxsize_x > 0 ? boxsize_x : (xmax - xmin)) : 0.;
        ywrap = py ? (bsy > 0 ? bsy : (ymax - ymin)) : 0.;
        zwrap = pz ? (bsz > 0 ? bsz : (zmax - zmin)) : 0.;
        if (mode == ORC_DD) { /* DD impl:252-282 */
            pimax = (REAL)rmax;
            if (!binning_cust) {
                if (rmax < 0.05 * xwrap) rf[0] = 1;
                if (rmax < 0.05 * ywrap) rf[1] = 1;
                if (pimax < 0.05 * zwrap) rf[2] = 1;
            }
            max_x = max_y = max_z = rmax;
            max3 = rmax;
        } else if (mode == ORC_RPPI) { /* rp_pi impl:188-290 */
            pimax = pimax_in;
            npibin = (int)pimax_in;
            if (!binning_cust) {
                if (rmax < 0.05 * xwrap) rf[0] = 1;
                if (rmax < 0.05 * ywrap) rf[1] = 1;
                if (pimax_in < 0.05 * zwrap) rf[2] = 1;
            }
            max_x = max_y = rmax;
            max_z = pimax_in;
            max2 = rmax;
            max1 = pimax_in;
        } else if (is_mocks) {
            REAL max_sep;
            if (mode == ORC_RPPI_MOCKS) { /* rp_pi_mocks_impl:307-308 */
                pimax = pimax_in;
                npibin = (int)pimax_in;
                const REAL sqr_max_sep = rmax * rmax + pimax * pimax;
                max_sep = SQRT_R(sqr_max_sep);
            } else { /* s_mu_mocks_impl:312-326, 424-432 */
                if (mu_max_in <= 0.0 || mu_max_in > 1.0 || nmu_bins < 1) return EXIT_FAILURE;
                mu_max = (REAL)mu_max_in;
                max_sep = rmax;
            }
            if (!binning_cust) { /* :413-423; DDsmu_mocks compares the double smax (s_mu_mocks_impl:424-432) */
                const double heur = (mode == ORC_SMU_MOCKS) ? rmax : (double)max_sep;
                if (heur < 0.05 * (xmax - xmin)) rf[0] = 1;
                if (heur < 0.05 * (ymax - ymin)) rf[1] = 1;
                if (heur < 0.05 * (zmax - zmin)) rf[2] = 1;
            }
            max_x = max_y = max_z = max_sep;
            max3 = max_sep;
        } else { /* ORC_SMU: s_mu impl:212-234, 305-345 */
            if (mu_max_in <= 0.0 || mu_max_in > 1.0 || nmu_bins < 1) return EXIT_FAILURE;
            mu_max = (REAL)mu_max_in;
            pimax = rmax * mu_max;
            max_x = max_y = rmax;
            max_z = pimax;
            max3 = rmax;
            max1 = pimax;
        }
    }

    FN(olattice) *L1 = FN(o_gridlink)(ND1, X1, Y1, Z1, need_w ? W1 : NULL, xmin, xmax, ymin, ymax, zmin, zmax, max_x,
                                      max_y, max_z, xwrap, ywrap, zwrap, rf[0], rf[1], rf[2], max_cells);
    if (!L1) return EXIT_FAILURE;
    if (mode != ORC_SMU) { /* boost: DD impl:298-332 (smu's boost multiplies by BOOST_BIN_REF=1: no-op) */
        double avg_np = ((double)ND1) / ((double)L1->nmesh[0] * L1->nmesh[1] * L1->nmesh[2]);
        if (mode == ORC_RPPI_MOCKS) { /* rp_pi_mocks_impl:442-444: the larger of the two sets (ND2 also when autocorr) */
            const double avg_np2 = ((double)ND2) / ((double)L1->nmesh[0] * L1->nmesh[1] * L1->nmesh[2]);
            if (avg_np2 > avg_np) avg_np = avg_np2;
        }
        const int max_nmesh = (int)fmax(L1->nmesh[0], fmax(L1->nmesh[1], L1->nmesh[2]));
        if ((max_nmesh <= 10 || avg_np >= 250) && max_nmesh < max_cells && !binning_cust) {
            FN(o_free_lattice)(L1);
            rf[0] += 1;
            rf[1] += 1;
            L1 = FN(o_gridlink)(ND1, X1, Y1, Z1, need_w ? W1 : NULL, xmin, xmax, ymin, ymax, zmin, zmax, max_x, max_y,
                                max_z, xwrap, ywrap, zwrap, rf[0], rf[1], rf[2], max_cells);
            if (!L1) return EXIT_FAILURE;
        }
    }
    FN(olattice) *L2 = L1;
    if (!autocorr) {
        L2 = FN(o_gridlink)(ND2, X2, Y2, Z2, need_w ? W2 : NULL, xmin, xmax, ymin, ymax, zmin, zmax, max_x, max_y,
                            max_z, xwrap, ywrap, zwrap, rf[0], rf[1], rf[2], max_cells);
        if (!L2) return EXIT_FAILURE;
    }
    if (lattice_out) {
        for (int i = 0; i < 3; i++) {
            lattice_out[i] = L1->nmesh[i];
            lattice_out[3 + i] = rf[i];
        }
    }
    int64_t ncp = 0;
    FN(opair) *CP = FN(o_cell_pairs)(L1, L2, &ncp, rf[0], rf[1], rf[2], xwrap, ywrap, zwrap, max3, max2, max1,
                                     enable_min_sep, autocorr, px, py, pz);

    int64_t nslots = nbin;
    FN(okern) K;
    memset(&K, 0, sizeof(K));
    K.mode = mode;
    K.nbin = nbin;
    K.edges = esq;
    K.need_avg = need_avg;
    K.need_w = need_w;
    K.pimax = pimax;
    if (mode == ORC_RPPI || mode == ORC_RPPI_MOCKS) { /* rp_pi_kernels:65-66, rp_pi_mocks_kernels:56-68 */
        K.npibin = npibin;
        const REAL dpi = pimax / npibin;
        K.inv_dpi = 1.0 / dpi;
        K.sqr_max_sep = esq[nbin - 1] + pimax * pimax;
        K.sqr_pimax = pimax * pimax;
        nslots = (int64_t)(npibin + 1) * (nbin + 1);
    } else if (mode == ORC_SMU || mode == ORC_SMU_MOCKS) { /* s_mu_kernels:65-68, s_mu_mocks_kernels:52-61 */
        K.nmu = nmu_bins;
        K.sqr_mumax = mu_max * mu_max;
        const REAL dmu = mu_max / (REAL)nmu_bins;
        K.inv_dmu = 1.0 / dmu;
        nslots = (int64_t)(nmu_bins + 1) * (nbin + 1);
    }

    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    uint64_t *tn = calloc((size_t)nthreads * nslots, sizeof(uint64_t));
    double *ta = calloc((size_t)nthreads * nslots, sizeof(double));
    double *tw = calloc((size_t)nthreads * nslots, sizeof(double));
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        FN(okern) k = K;
        k.npairs = tn + (size_t)tid * nslots;
        k.avg = ta + (size_t)tid * nslots;
        k.wavg = tw + (size_t)tid * nslots;
#ifdef _OPENMP
#pragma omp for schedule(dynamic)
#endif
        for (int64_t p = 0; p < ncp; p++) {
            const FN(ocell) *a = &L1->cells[CP[p].c1], *b = &L2->cells[CP[p].c2];
            if (orc_literal_kernels && (mode == ORC_DD || mode == ORC_XI)) {
                FN(o_count_cellpair_literal_dd)(&k, &CP[p], (REAL)rmax, a->n, L1->x + a->start, L1->y + a->start,
                                                L1->z + a->start, need_w ? L1->w + a->start : NULL, b->n,
                                                L2->x + b->start, L2->y + b->start, L2->z + b->start,
                                                need_w ? L2->w + b->start : NULL);
                continue;
            }
            (orc_literal_kernels && (mode == ORC_WP || mode == ORC_RPPI) ? FN(o_count_cellpair_literal)
                                                                          : FN(o_count_cellpair))(
                                 &k, a->n, L1->x + a->start, L1->y + a->start, L1->z + a->start,
                                 need_w ? L1->w + a->start : NULL, b->n, L2->x + b->start, L2->y + b->start,
                                 L2->z + b->start, need_w ? L2->w + b->start : NULL, CP[p].same, CP[p].xw, CP[p].yw,
                                 CP[p].zw);
        }
    }
    uint64_t *npairs = calloc(nslots, sizeof(uint64_t));
    double *avg = calloc(nslots, sizeof(double)), *wavg = calloc(nslots, sizeof(double));
    for (int t = 0; t < nthreads; t++)
        for (int64_t s = 0; s < nslots; s++) {
            npairs[s] += tn[(size_t)t * nslots + s];
            avg[s] += ta[(size_t)t * nslots + s];
            wavg[s] += tw[(size_t)t * nslots + s];
        }
    free(tn);
    free(ta);
    free(tw);
    free(CP);

    /* epilogue: DD impl:609-664 and siblings */
    if (autocorr) {
        for (int64_t s = 0; s < nslots; s++) {
            npairs[s] *= 2;
            avg[s] *= 2.0;
            wavg[s] *= 2.0;
        }
        if (rupp[0] <= 0.0) {
            const int64_t first = (mode == ORC_RPPI) ? (npibin + 1) : (mode == ORC_SMU ? (nmu_bins + 1) : 1);
            npairs[first] += ND1;
            if (need_w)
                for (int64_t j = 0; j < ND1; j++) wavg[1] += (double)(REAL)(W1[j] * W1[j]); /* always slot 1 */
        }
    }
    for (int64_t s = 0; s < nslots; s++)
        if (npairs[s] > 0) {
            avg[s] /= (double)npairs[s];
            wavg[s] /= (double)npairs[s];
        }
    for (int64_t s = 0; s < nslots; s++) {
        npairs_out[s] = npairs[s];
        avg_out[s] = need_avg ? avg[s] : 0.0;
        wavg_out[s] = need_w ? wavg[s] : 0.0;
    }
    if (is_box && cf_out) { /* xi impl:581-625, wp impl:619-664 -- arithmetic in REAL like the reference */
        REAL weightsum = (REAL)ND1, weight_sqr_sum = (REAL)ND1;
        if (need_w) {
            weightsum = 0;
            for (int64_t j = 0; j < ND1; j++) {
                weightsum += W1[j];
                weight_sqr_sum += W1[j] * W1[j];
            }
        }
        const REAL prefac = weightsum * (weightsum - weightsum / ND1) / (boxsize_x * boxsize_x * boxsize_x);
        REAL rlow = 0.0;
        const REAL twice_pimax = 2.0 * pimax_in;
        for (int i = 0; i < nbin; i++) {
            REAL weight0 = (REAL)npairs_out[i];
            if (need_w) weight0 *= wavg_out[i];
            const REAL vol = (mode == ORC_XI)
                                 ? (REAL)(4.0 / 3.0 * M_PI * (rupp[i] * rupp[i] * rupp[i] - rlow * rlow * rlow))
                                 : (REAL)(M_PI * (rupp[i] * rupp[i] - rlow * rlow) * twice_pimax);
            if (vol > 0.0) {
                REAL weightrandom = prefac * vol;
                if (rlow <= 0.) weightrandom += weight_sqr_sum;
                cf_out[i] = (mode == ORC_XI) ? (REAL)(weight0 / weightrandom - 1.0)
                                             : (REAL)((weight0 / weightrandom - 1) * twice_pimax);
            } else {
                cf_out[i] = (mode == ORC_XI) ? -2.0 : (REAL)(-2.0 * twice_pimax);
            }
            rlow = rupp[i];
        }
    }
    free(npairs);
    free(avg);
    free(wavg);
    free(esq);
    for (int i = 0; i < 6; i++) free(conv[i]);
    if (L2 != L1) FN(o_free_lattice)(L2);
    FN(o_free_lattice)(L1);
    return EXIT_SUCCESS;
}


/* ------------------------------------------------------------------------------------------ */
/* DDtheta: mocks/DDtheta_mocks/countpairs_theta_mocks_impl.c.src:456-1203 with the RA/DEC lattice */
/* of utils/gridlink_mocks_impl.c.src:1006-1650 (link_in_ra) / :552-1003 (DEC only) / one cell.   */
static REAL FN(o_min_sep_1d)(const REAL a[2], const REAL b[2])
{ /* find_closest_pos_DOUBLE, utils/gridlink_utils.c.src:91-113 */
    if (a[0] <= b[1] && b[0] <= a[1]) return 0;
    REAL m = FABS_R(a[0] - b[0]);
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) {
            const REAL d = FABS_R(a[i] - b[j]);
            if (d < m) m = d;
        }
    return m;
}

typedef struct {
    int64_t ncells;
    FN(ocell) * cells;
    REAL *x, *y, *z, *w;
} FN(othlat);

static FN(othlat) * FN(o_theta_gridlink)(const int64_t N, const REAL *RA, const REAL *DEC, const REAL *W,
                                          const int ngrid_dec, const int *ngrid_ra, const int64_t *ra_off,
                                          const int64_t ncells, const REAL dec_min, const REAL inv_dec_diff,
                                          const REAL ra_min, const REAL inv_ra_diff)
{
    FN(othlat) *L = calloc(1, sizeof(*L));
    L->ncells = ncells;
    L->cells = calloc(ncells, sizeof(*L->cells));
    L->x = malloc(sizeof(REAL) * N);
    L->y = malloc(sizeof(REAL) * N);
    L->z = malloc(sizeof(REAL) * N);
    L->w = W ? malloc(sizeof(REAL) * N) : NULL;
    int64_t *idx = malloc(sizeof(int64_t) * N);
    for (int64_t i = 0; i < N; i++) { /* gridlink_mocks_impl.c.src:1249-1263 */
        int idec = (int)(ngrid_dec * (DEC[i] - dec_min) * inv_dec_diff);
        if (idec >= ngrid_dec) idec--;
        int ira = (int)(ngrid_ra[idec] * (RA[i] - ra_min) * inv_ra_diff);
        if (ira >= ngrid_ra[idec]) ira--;
        if (idec < 0 || idec >= ngrid_dec || ira < 0 || ira >= ngrid_ra[idec]) {
            fprintf(stderr, "oracle> theta cell index out of range\n");
            return NULL;
        }
        idx[i] = ra_off[idec] + ira;
        L->cells[idx[i]].n++;
    }
    int64_t off = 0;
    for (int64_t c = 0; c < ncells; c++) {
        FN(ocell) *q = &L->cells[c];
        q->start = off;
        off += q->n;
        q->n = 0;
        q->xb[0] = q->yb[0] = q->zb[0] = q->rab[0] = q->decb[0] = MAXPOS_R;
        q->xb[1] = q->yb[1] = q->zb[1] = q->rab[1] = q->decb[1] = -MAXPOS_R;
    }
    for (int64_t i = 0; i < N; i++) {
        FN(ocell) *q = &L->cells[idx[i]];
        const int64_t p = q->start + q->n++;
        /* unit vectors with the host libm (countpairs_theta_mocks_impl.c.src:587-591) */
        const REAL X = COSD_R(DEC[i]) * COSD_R(RA[i]);
        const REAL Y = COSD_R(DEC[i]) * SIND_R(RA[i]);
        const REAL Z = SIND_R(DEC[i]);
        L->x[p] = X;
        L->y[p] = Y;
        L->z[p] = Z;
        if (W) L->w[p] = W[i];
        if (X < q->xb[0]) q->xb[0] = X;
        if (Y < q->yb[0]) q->yb[0] = Y;
        if (Z < q->zb[0]) q->zb[0] = Z;
        if (X > q->xb[1]) q->xb[1] = X;
        if (Y > q->yb[1]) q->yb[1] = Y;
        if (Z > q->zb[1]) q->zb[1] = Z;
        if (RA[i] < q->rab[0]) q->rab[0] = RA[i];
        if (RA[i] > q->rab[1]) q->rab[1] = RA[i];
    }
    free(idx);
    return L;
}

static void FN(o_free_thlat)(FN(othlat) * L)
{
    if (!L) return;
    free(L->cells);
    free(L->x);
    free(L->y);
    free(L->z);
    free(L->w);
    free(L);
}

/* RA1/DEC1 (and RA2/DEC2) must already be in [0,360] / [-90,90]; theta_upp[] are the nbin edges in
 * degrees (for REAL=float they must be float-representable, as setup_bins_float would produce). */
int FN(oracle_theta)(const int64_t ND1, const REAL *RA1, const REAL *DEC1, const REAL *W1, const int64_t ND2,
                     const REAL *RA2, const REAL *DEC2, const REAL *W2, const int autocorr, const int nbin,
                     const double *theta_upp, const int link_in_dec, const int link_in_ra, const int ra_refine,
                     const int dec_refine, int max_cells, const int enable_min_sep, const int need_avg,
                     const int need_w, const int fast_acos, uint64_t *npairs_out, double *avg_out,
                     double *wavg_out, int *lattice_out /* ngrid_dec, ncells, ncellpairs */)
{
    if (max_cells == 0) max_cells = 100;
    const REAL thetamax = theta_upp[nbin - 1];
    REAL *cosup = malloc(sizeof(REAL) * nbin);
    for (int i = 0; i < nbin; i++) {
        const REAL t = theta_upp[i];
        cosup[i] = COSD_R(t);
    }
    REAL ra_min = MAXPOS_R, dec_min = MAXPOS_R, ra_max = -MAXPOS_R, dec_max = -MAXPOS_R;
    for (int s = 0; s < (autocorr ? 1 : 2); s++) {
        const int64_t n = s ? ND2 : ND1;
        const REAL *ra = s ? RA2 : RA1, *dec = s ? DEC2 : DEC1;
        for (int64_t i = 0; i < n; i++) {
            if (ra[i] < ra_min) ra_min = ra[i];
            if (dec[i] < dec_min) dec_min = dec[i];
            if (ra[i] > ra_max) ra_max = ra[i];
            if (dec[i] > dec_max) dec_max = dec[i];
        }
    }
    const REAL dec_diff = dec_max - dec_min, ra_diff = ra_max - ra_min;
    int ngrid_dec = 1;
    if (link_in_dec || link_in_ra) { /* gridlink_mocks_impl.c.src:1061-1066 */
        const REAL this_ngrid_dec = (dec_diff / thetamax < 1) ? 1 : dec_diff / thetamax;
        const int this_ngrid_dec_int = ((int)this_ngrid_dec) * dec_refine;
        ngrid_dec = this_ngrid_dec_int > max_cells ? max_cells : this_ngrid_dec_int;
        ngrid_dec = ngrid_dec < 1 ? 1 : ngrid_dec;
    }
    int *ngrid_ra = malloc(sizeof(int) * ngrid_dec);
    int64_t *ra_off = malloc(sizeof(int64_t) * ngrid_dec);
    const REAL dec_binsize = dec_diff / ngrid_dec;
    const REAL sin_half_thetamax = SIND_R(0.5 * thetamax);
    const REAL max_phi_cell = ra_diff;
    int64_t ncells = 0;
    for (int idec = 0; idec < ngrid_dec; idec++) { /* gridlink_mocks_impl.c.src:1081-1147 */
        int nmesh_ra = 1;
        if (link_in_ra) {
            REAL this_min_dec, cos_min_dec;
            const REAL dec_lower = dec_min + idec * dec_binsize;
            const REAL dec_upper = dec_lower + dec_binsize;
            const REAL cos_dec_upper = COSD_R(dec_upper);
            const REAL cos_dec_lower = COSD_R(dec_lower);
            if (cos_dec_lower < cos_dec_upper) {
                this_min_dec = dec_lower;
                cos_min_dec = cos_dec_lower;
            } else {
                this_min_dec = dec_upper;
                cos_min_dec = cos_dec_upper;
            }
            REAL phi_cell = max_phi_cell;
            if ((90.0 - FABS_R(this_min_dec)) > 1.0) {
                const REAL _tmp = sin_half_thetamax / cos_min_dec;
                const REAL _tmp1 = _tmp < 0 ? 0 : (_tmp > 1.0 ? 1.0 : _tmp);
                phi_cell = 2.0 * ASIN_R(_tmp1) * ORC_INV_PI_OVER_180;
                if (phi_cell <= 0) phi_cell = max_phi_cell;
            }
            phi_cell = phi_cell > max_phi_cell ? max_phi_cell : phi_cell;
            const REAL this_nmesh_ra = (ra_diff / phi_cell < 1) ? 1 : ra_diff / phi_cell;
            const int this_nmesh_ra_int = ((int)this_nmesh_ra) * ra_refine;
            nmesh_ra = this_nmesh_ra_int > max_cells ? max_cells : this_nmesh_ra_int;
            if (nmesh_ra < 1) nmesh_ra = 1;
        }
        ngrid_ra[idec] = nmesh_ra;
        ra_off[idec] = ncells;
        ncells += nmesh_ra;
    }
    const REAL inv_dec_diff = dec_diff > 0 ? (REAL)(1.0 / dec_diff) : (REAL)0;
    const REAL inv_ra_diff = ra_diff > 0 ? (REAL)(1.0 / ra_diff) : (REAL)0;
    FN(othlat) *L1 = FN(o_theta_gridlink)(ND1, RA1, DEC1, need_w ? W1 : NULL, ngrid_dec, ngrid_ra, ra_off, ncells,
                                          dec_min, inv_dec_diff, ra_min, inv_ra_diff);
    FN(othlat) *L2 = L1;
    if (!autocorr)
        L2 = FN(o_theta_gridlink)(ND2, RA2, DEC2, need_w ? W2 : NULL, ngrid_dec, ngrid_ra, ra_off, ncells, dec_min,
                                  inv_dec_diff, ra_min, inv_ra_diff);
    if (!L1 || !L2) return EXIT_FAILURE;

    /* cell pairs: gridlink_mocks_impl.c.src:1481-1650 (RA+DEC) and :902-1003 (DEC only) */
    const REAL sqr_max_chord_sep = 2.0 * (1.0 - COSD_R(thetamax));
    size_t cap = (size_t)ncells * 8 + 64, np = 0;
    int64_t *pc1 = malloc(sizeof(int64_t) * cap), *pc2 = malloc(sizeof(int64_t) * cap);
    for (int idec = 0; idec < ngrid_dec; idec++) {
        for (int ira = 0; ira < ngrid_ra[idec]; ira++) {
            const int64_t icell = ra_off[idec] + ira;
            const FN(ocell) *first = &L1->cells[icell];
            if (first->n == 0) continue;
            const size_t first_np = np;
            for (int dr = -dec_refine; dr <= dec_refine; dr++) {
                const int this_dec = idec + dr;
                if (this_dec < 0 || this_dec >= ngrid_dec) continue;
                int lo_ra = 0, hi_ra = 0;
                if (link_in_ra) {
                    const int min_ra = (int)(ngrid_ra[this_dec] * (first->rab[0] - ra_min) * inv_ra_diff) - 1;
                    const int max_ra = (int)(ngrid_ra[this_dec] * (first->rab[1] - ra_min) * inv_ra_diff) + 1;
                    lo_ra = min_ra - ra_refine;
                    hi_ra = max_ra + ra_refine;
                }
                for (int iira = lo_ra; iira <= hi_ra; iira++) {
                    int this_ra = iira + ngrid_ra[this_dec];
                    while (this_ra < 0) this_ra += ngrid_ra[this_dec];
                    this_ra = this_ra % ngrid_ra[this_dec];
                    const int64_t icell2 = ra_off[this_dec] + this_ra;
                    const FN(ocell) *second = &L2->cells[icell2];
                    if (second->n == 0 || (autocorr == 1 && icell2 > icell)) continue;
                    int dup = 0;
                    for (size_t q = first_np; q < np; q++)
                        if (pc2[q] == icell2) {
                            dup = 1;
                            break;
                        }
                    if (dup) continue;
                    if (enable_min_sep) {
                        if (link_in_ra) {
                            const REAL mx = FN(o_min_sep_1d)(first->xb, second->xb);
                            const REAL my = FN(o_min_sep_1d)(first->yb, second->yb);
                            const REAL mz = FN(o_min_sep_1d)(first->zb, second->zb);
                            if (mx * mx + my * my + mz * mz >= sqr_max_chord_sep) continue;
                        } else if (dr != 0) {
                            const REAL fz = dr < 0 ? first->zb[0] : first->zb[1];
                            const REAL sz = dr < 0 ? second->zb[1] : second->zb[0];
                            const REAL mz = fz - sz;
                            if (mz * mz >= sqr_max_chord_sep) continue;
                        }
                    }
                    if (np + 1 > cap) {
                        cap *= 2;
                        pc1 = realloc(pc1, sizeof(int64_t) * cap);
                        pc2 = realloc(pc2, sizeof(int64_t) * cap);
                    }
                    pc1[np] = icell;
                    pc2[np] = icell2;
                    np++;
                }
            }
        }
    }
    if (lattice_out) {
        lattice_out[0] = ngrid_dec;
        lattice_out[1] = (int)ncells;
        lattice_out[2] = (int)np;
    }
    FN(okern) K;
    memset(&K, 0, sizeof(K));
    K.mode = ORC_THETA;
    K.nbin = nbin;
    K.edges = cosup;
    K.need_avg = need_avg;
    K.need_w = need_w;
    K.fast_acos = fast_acos;
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    uint64_t *tn = calloc((size_t)nthreads * nbin, sizeof(uint64_t));
    double *ta = calloc((size_t)nthreads * nbin, sizeof(double));
    double *tw = calloc((size_t)nthreads * nbin, sizeof(double));
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        FN(okern) k = K;
        k.npairs = tn + (size_t)tid * nbin;
        k.avg = ta + (size_t)tid * nbin;
        k.wavg = tw + (size_t)tid * nbin;
#ifdef _OPENMP
#pragma omp for schedule(dynamic)
#endif
        for (int64_t p = 0; p < (int64_t)np; p++) {
            const FN(ocell) *a = &L1->cells[pc1[p]], *b = &L2->cells[pc2[p]];
            FN(o_count_cellpair)(&k, a->n, L1->x + a->start, L1->y + a->start, L1->z + a->start,
                                 need_w ? L1->w + a->start : NULL, b->n, L2->x + b->start, L2->y + b->start,
                                 L2->z + b->start, need_w ? L2->w + b->start : NULL,
                                 (autocorr && pc1[p] == pc2[p]) ? 1 : 0, 0, 0, 0);
        }
    }
    for (int i = 0; i < nbin; i++) {
        npairs_out[i] = 0;
        avg_out[i] = 0;
        wavg_out[i] = 0;
        for (int t = 0; t < nthreads; t++) {
            npairs_out[i] += tn[(size_t)t * nbin + i];
            avg_out[i] += ta[(size_t)t * nbin + i];
            wavg_out[i] += tw[(size_t)t * nbin + i];
        }
    }
    free(tn);
    free(ta);
    free(tw);
    free(pc1);
    free(pc2);
    /* epilogue: countpairs_theta_mocks_impl.c.src:1125-1180 */
    if (autocorr) {
        for (int i = 0; i < nbin; i++) {
            npairs_out[i] *= 2;
            avg_out[i] *= 2.0;
            wavg_out[i] *= 2.0;
        }
        if (theta_upp[0] <= 0.0) {
            npairs_out[1] += ND1;
            if (need_w)
                for (int64_t j = 0; j < ND1; j++) wavg_out[1] += (double)(REAL)(W1[j] * W1[j]);
        }
    }
    for (int i = 1; i < nbin; i++)
        if (npairs_out[i] > 0) {
            avg_out[i] /= (double)npairs_out[i];
            wavg_out[i] /= (double)npairs_out[i];
        }
    if (!need_avg)
        for (int i = 0; i < nbin; i++) avg_out[i] = 0;
    if (!need_w)
        for (int i = 0; i < nbin; i++) wavg_out[i] = 0;
    if (L2 != L1) FN(o_free_thlat)(L2);
    FN(o_free_thlat)(L1);
    free(ngrid_ra);
    free(ra_off);
    free(cosup);
    return EXIT_SUCCESS;
}



/* ------------------------------------------------------------------------------------------ */
/* Counts-in-spheres on a survey catalogue: mocks/vpf_mocks/countspheres_mocks_impl.c.src:206-638 with the        */
/* AVX-512 kernel of vpf_mocks_kernels.c.src:20-100.  Centres are given (the reference reads them from its        */
/* centres file, or takes the randoms that pass o_vpf_neighbours below); the result does not depend on the        */
/* lattice, so every galaxy is tried against every centre.                                                        */
int FN(oracle_vpf_mocks)(const int64_t Ngal, const REAL *RA, const REAL *DEC, const REAL *D /* comoving */,
                         const int64_t nc, const REAL *xc, const REAL *yc, const REAL *zc /* shifted, as in the file */,
                         const double rmax_in, const int nbin, const int num_pN, double *pN_out /* [nbin][num_pN] */,
                         const double dmax_randoms /* largest distance among the randoms when they are in play, else 0 */,
                         double *rcube_out)
{
    if (!(rmax_in > 0.0) || nbin < 1 || nc < 1 || num_pN < 1) return EXIT_FAILURE;
    const REAL rmax = rmax_in;
    REAL *x = malloc(sizeof(REAL) * (Ngal > 0 ? Ngal : 1)), *y = malloc(sizeof(REAL) * (Ngal > 0 ? Ngal : 1)),
         *z = malloc(sizeof(REAL) * (Ngal > 0 ? Ngal : 1));
    REAL rcube = (REAL)dmax_randoms; /* :351-372: the randoms widen the bounding cube too */
    for (int64_t i = 0; i < Ngal; i++) { /* :338-349 */
        const REAL dc = D[i];
        if (dc > rcube) rcube = dc;
        x[i] = dc * COSD_R(DEC[i]) * COSD_R(RA[i]);
        y[i] = dc * COSD_R(DEC[i]) * SIND_R(RA[i]);
        z[i] = dc * SIND_R(DEC[i]);
    }
    rcube = rcube + 1.; /* :385-392: shift into [0, 2 rcube] */
    for (int64_t i = 0; i < Ngal; i++) {
        x[i] += rcube;
        y[i] += rcube;
        z[i] += rcube;
    }
    if (rcube_out) *rcube_out = (double)rcube;
    const REAL rstep = rmax / (REAL)nbin; /* vpf_mocks_kernels:38-46 */
    const REAL rmax_sqr = rmax * rmax;
    REAL *E = malloc(sizeof(REAL) * nbin);
    for (int k = 0; k < nbin; k++) E[k] = (k + 1) * rstep * rstep * (k + 1);
    double *pN = calloc((size_t)nbin * num_pN, sizeof(double));
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        uint64_t *counts = calloc((size_t)nbin, sizeof(uint64_t));
        double *mine = calloc((size_t)nbin * num_pN, sizeof(double));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (int64_t c = 0; c < nc; c++) {
            for (int k = 0; k < nbin; k++) counts[k] = 0;
            for (int64_t j = 0; j < Ngal; j++) {
                const REAL dx = xc[c] - x[j], dy = yc[c] - y[j], dz = zc[c] - z[j];
                const REAL r2 = FMA_R(dz, dz, FMA_R(dy, dy, dx * dx));
                if (!(r2 < rmax_sqr)) continue;
                /* :80-92 lane by lane: the bin with E[k-1] <= r2 < E[k]; whoever is left after k == 1 lands in bin 0
                 * (also an r2 at or above the last edge); with a single bin the loop never runs and nothing counts */
                int left = 1;
                for (int k = nbin - 1; k >= 1; k--)
                    if (r2 < E[k] && r2 >= E[k - 1]) {
                        counts[k]++;
                        left = 0;
                        break;
                    }
                if (left && nbin >= 2) counts[0]++;
            }
            for (int k = 1; k < nbin; k++) counts[k] += counts[k - 1]; /* impl:565-567 */
            for (int k = 0; k < nbin; k++)
                if (counts[k] < (uint64_t)num_pN) mine[(size_t)k * num_pN + counts[k]] += 1.0;
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        for (int64_t i = 0; i < (int64_t)nbin * num_pN; i++) pN[i] += mine[i];
        free(counts);
        free(mine);
    }
    const REAL inv_nc = ((REAL)1.0) / (REAL)nc; /* :616-621, arithmetic in REAL */
    for (int k = 0; k < nbin; k++)
        for (int i = 0; i < num_pN; i++) pN_out[(size_t)k * num_pN + i] = (double)(REAL)((REAL)pN[(size_t)k * num_pN + i] * inv_nc);
    free(pN);
    free(E);
    free(x);
    free(y);
    free(z);
    return EXIT_SUCCESS;
}

/* Counts-in-spheres in a simulation box: theory/vpf/countspheres_impl.c.src:138-479 with the AVX-512 kernel of
 * theory/vpf/vpf_kernels.c.src.  The centres are given (the reference draws them from gsl_rng_mt19937).  On a periodic
 * axis the reference shifts the centre by -+wrap for neighbour cells across the box edge; a particle that can count is
 * always met with its nearest image, which is what the brute force below takes. */
int FN(oracle_vpf_theory)(const int64_t np, const REAL *X, const REAL *Y, const REAL *Z, const int64_t nc, const REAL *xc,
                          const REAL *yc, const REAL *zc, const int periodic, const double wrapx, const double wrapy,
                          const double wrapz, const double rmax_in, const int nbin, const int num_pN, double *pN_out)
{
    if (!(rmax_in > 0.0) || nbin < 1 || nc < 1 || num_pN < 1) return EXIT_FAILURE;
    const REAL rmax = rmax_in;
    const REAL wrap[3] = {(REAL)wrapx, (REAL)wrapy, (REAL)wrapz};
    const REAL rstep = rmax / (REAL)nbin;
    const REAL rmax_sqr = rmax * rmax;
    REAL *E = malloc(sizeof(REAL) * nbin);
    for (int k = 0; k < nbin; k++) E[k] = (k + 1) * rstep * rstep * (k + 1);
    int64_t *pN = calloc((size_t)nbin * num_pN, sizeof(int64_t));
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        uint64_t *counts = calloc((size_t)nbin, sizeof(uint64_t));
        int64_t *mine = calloc((size_t)nbin * num_pN, sizeof(int64_t));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (int64_t c = 0; c < nc; c++) {
            const REAL C3[3] = {xc[c], yc[c], zc[c]};
            for (int k = 0; k < nbin; k++) counts[k] = 0;
            for (int64_t j = 0; j < np; j++) {
                const REAL P3[3] = {X[j], Y[j], Z[j]};
                REAL d[3];
                for (int a = 0; a < 3; a++) {
                    REAL cen = C3[a];
                    if (periodic) {
                        const REAL raw = P3[a] - C3[a], half = (REAL)0.5 * wrap[a];
                        if (raw > half) cen = C3[a] + wrap[a];
                        else if (raw < -half) cen = C3[a] - wrap[a];
                    }
                    d[a] = cen - P3[a];
                }
                const REAL r2 = FMA_R(d[2], d[2], FMA_R(d[1], d[1], d[0] * d[0]));
                if (!(r2 < rmax_sqr)) continue;
                int left = 1;
                for (int k = nbin - 1; k >= 1; k--)
                    if (r2 < E[k] && r2 >= E[k - 1]) {
                        counts[k]++;
                        left = 0;
                        break;
                    }
                if (left && nbin >= 2) counts[0]++;
            }
            for (int k = 1; k < nbin; k++) counts[k] += counts[k - 1];
            for (int k = 0; k < nbin; k++)
                if (counts[k] < (uint64_t)num_pN) mine[(size_t)k * num_pN + counts[k]]++;
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        for (int64_t i = 0; i < (int64_t)nbin * num_pN; i++) pN[i] += mine[i];
        free(counts);
        free(mine);
    }
    const REAL inv_nc = ((REAL)1.0) / (REAL)nc;
    for (int64_t i = 0; i < (int64_t)nbin * num_pN; i++) pN_out[i] = (double)(REAL)((int)pN[i] * inv_nc);
    free(pN);
    free(E);
    return EXIT_SUCCESS;
}

#undef FMA_R
#undef SQRT_R
#undef FABS_R
#undef ACOS_R
#undef ASIN_R
#undef COSD_R
#undef SIND_R
#undef MAXPOS_R
#undef CAT_
#undef CAT
#undef FN
