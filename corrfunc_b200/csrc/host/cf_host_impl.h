/* cf_host_impl.h -- precision-templated host logic (included by cf_host.c with REAL=float|double).
 *
 * cf_box_<sfx>   : countpairs / countpairs_rp_pi / countpairs_s_mu / countpairs_wp / countpairs_xi
 * cf_theta_<sfx> : countpairs_theta_mocks
 *
 * All REAL-typed arithmetic below mirrors the type and operation order of the reference's drivers so
 * that the lattice (nmesh, binsize, inverse), the squared edges and the 2-D bin scale factors come
 * out bit-identical; file:line citations point at the statement being mirrored.
 */
#define HCAT_(a, b) a##_##b
#define HCAT(a, b) HCAT_(a, b)
#define HFN(name) HCAT(name, SFX)

#if REAL_IS_DOUBLE
#define H_COSD(X) cos((X) * CF_PI_OVER_180)
#define H_SIND(X) sin((X) * CF_PI_OVER_180)
#define H_ASIN(X) asin(X)
#define H_FABS(X) fabs(X)
#define H_SQRT(X) sqrt(X)
#define H_MAXPOS DBL_MAX
#else
#define H_COSD(X) cosf((X) * CF_PI_OVER_180)
#define H_SIND(X) sinf((X) * CF_PI_OVER_180)
#define H_ASIN(X) asinf(X)
#define H_FABS(X) fabsf(X)
#define H_SQRT(X) sqrtf(X)
#define H_MAXPOS FLT_MAX
#endif

/* get_binsize_DOUBLE (utils/gridlink_utils.c.src:31-49) */
static int HFN(cf_binsize)(const REAL xdiff, const REAL xwrap, const REAL rmax, const int refine_factor,
                           const int max_ncells, REAL *xbinsize, int *nlattice)
{
    int nmesh = (int)(refine_factor * xdiff / rmax);
    nmesh = nmesh < 1 ? 1 : nmesh;
    if (xwrap > 0 && rmax >= xwrap / 2) {
        fprintf(stderr, "%s> ERROR: rmax=%f must be less than half of periodic boxize=%f to avoid double-counting particles\n",
                __FILE__, (double)rmax, (double)xwrap);
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
    REAL inv[3];
} HFN(cf_mesh);

/* the sizing part of gridlink_DOUBLE (utils/gridlink_impl.c.src:94-100,156-158) */
static int HFN(cf_mesh_for)(const REAL lo[3], const REAL hi[3], const REAL maxsize[3], const REAL wrap[3],
                            const int8_t refine[3], const int max_cells, HFN(cf_mesh) * M)
{
    int bad = 0;
    for (int a = 0; a < 3; a++) {
        REAL binsize = 0;
        bad |= HFN(cf_binsize)(hi[a] - lo[a], wrap[a], maxsize[a], refine[a], max_cells, &binsize, &M->nmesh[a]);
        M->inv[a] = binsize > 0 ? 1.0 / binsize : 0.;
    }
    if (bad) {
        fprintf(stderr, "Received an error status while sizing the lattice. Error\n");
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}

static const void *HFN(host_copy)(const void *p, const int64_t n, void **tofree)
{ /* the epilogue reads the weights on the host (self-pair term, weight sums of xi / wp).  Like the positions, the
     weight arrays may be device memory (cfb_upload borrows device pointers): those are copied back first. */
    *tofree = NULL;
    if (!p || n <= 0 || !cfb_is_device_ptr(p)) return p;
    void *h = malloc((size_t)n * sizeof(REAL));
    if (!h || cfb_copy_to_host(h, p, (size_t)n * sizeof(REAL))) {
        free(h);
        return NULL;
    }
    *tofree = h;
    return h;
}

static int HFN(cf_box)(const int mode, const int64_t ND1, void *X1, void *Y1, void *Z1, const int64_t ND2, void *X2,
                       void *Y2, void *Z2, const int numthreads, int autocorr, const char *binfile,
                       const double pimax_in, const double max_mu, const int nmu_bins, const double boxsize,
                       struct config_options *options, struct extra_options *extra, cf_box_out *out)
{
    (void)numthreads; /* host helper threads are not needed: the per-particle work is on the GPU */
    const double t_start = now_ms();
    if (options->float_type != sizeof(REAL)) {
        fprintf(stderr, "ERROR: In %s> Can only handle arrays of size=%zu. Got an array of size = %zu\n", __func__,
                sizeof(REAL), options->float_type);
        return EXIT_FAILURE;
    }
    struct extra_options dummy_extra;
    if (extra == NULL) { /* countpairs_impl.c.src:154-159 */
        dummy_extra = get_extra_options(NONE);
        extra = &dummy_extra;
    }
    const int need_weightavg = extra->weight_method != NONE;
    if (need_weightavg && extra->weight_method != PAIR_PRODUCT) {
        fprintf(stderr, "Error: unknown weight method %d\n", (int)extra->weight_method);
        return EXIT_FAILURE;
    }
    const int is_box = (mode == CFB_XI || mode == CFB_WP);
    /* survey geometry (mocks/DDrppi_mocks, mocks/DDsmu_mocks): X/Y/Z are the Cartesian positions HFN(cf_mocks) made */
    const int is_mocks = (mode == CFB_RPPI_MOCKS || mode == CFB_SMU_MOCKS);
    if (is_box) { /* countpairs_xi_impl.c.src:176-178, countpairs_wp_impl.c.src:191-193 */
        options->periodic = 1;
        options->autocorr = 1;
        autocorr = 1;
    }
    options->sort_on_z = 1;
    for (int i = 0; i < 3; i++) {
        if (options->bin_refine_factors[i] < 1) {
            fprintf(stderr, "Warning: bin refine factor along axis = %d *must* be >=1. Instead found bin refine factor =%d\n",
                    i, options->bin_refine_factors[i]);
            reset_bin_refine_factors(options);
            break;
        }
    }
    if (options->max_cells_per_dim == 0) {
        fprintf(stderr, "Warning: Max. cells per dimension is set to 0 - resetting to `NLATMAX' = %d\n", NLATMAX);
        options->max_cells_per_dim = NLATMAX;
    }
    if ((mode == CFB_SMU || is_mocks) && options->fast_divide_and_NR_steps >= MAX_FAST_DIVIDE_NR_STEPS) {
        options->fast_divide_and_NR_steps = 0;
    }
    if (!(autocorr == 0 || autocorr == 1)) {
        fprintf(stderr, "Error: Strange value of autocorr = %d. Expected to receive either 1 (auto-correlations) or 0 (cross-correlations)\n", autocorr);
        return EXIT_FAILURE;
    }

    /* ---- bins ---- */
    double *rupp = NULL, rpmin, rpmax;
    int nrpbin;
    if (cf_setup_bins(binfile, &rpmin, &rpmax, &nrpbin, &rupp, 0) != EXIT_SUCCESS) return EXIT_FAILURE;
    /* the mocks statistics insist on rmin > 0 (countpairs_rp_pi_mocks_impl.c.src:304, countpairs_s_mu_mocks_impl.c.src:304) */
    if (!((is_mocks ? rpmin > 0.0 : rpmin >= 0.0) && rpmax > 0.0 && rpmin < rpmax && nrpbin > 0)) {
        fprintf(stderr, "Error: Could not setup with R bins correctly. (rmin = %lf, rmax = %lf, with nbins = %d). "
                        "Expected non-zero rmin/rmax with rmax > rmin and nbins >=1 \n", rpmin, rpmax, nrpbin);
        free(rupp);
        return EXIT_FAILURE;
    }
    if (mode == CFB_SMU || mode == CFB_SMU_MOCKS) { /* countpairs_s_mu_impl.c.src:212-223, s_mu_mocks_impl:309-319 */
        if (max_mu <= 0.0 || max_mu > 1.0) {
            fprintf(stderr, "Error: max_mu (max. value for the cosine of the angle with line of sight) must be greater than 0 and at most 1).\n"
                            "The passed value is max_mu = %lf. Please change it to be > 0 and <= 1.0\n", max_mu);
            free(rupp);
            return EXIT_FAILURE;
        }
        if (nmu_bins < 1) {
            fprintf(stderr, "Error: Number of mu bins = %d must be at least 1\n", nmu_bins);
            free(rupp);
            return EXIT_FAILURE;
        }
    }
    double *rupp_sqr = malloc(sizeof(double) * (size_t)nrpbin);
    for (int i = 0; i < nrpbin; i++) {
        const REAL sq = rupp[i] * rupp[i]; /* double product rounded to REAL (countpairs_impl.c.src:435-438) */
        rupp_sqr[i] = (double)sq;
    }

    /* ---- particles to the device ---- */
    const double t_up0 = now_ms();
    const void *W1 = need_weightavg ? extra->weights0.weights[0] : NULL;
    const void *W2 = need_weightavg ? extra->weights1.weights[0] : NULL;
    if (need_weightavg && (W1 == NULL || (!autocorr && W2 == NULL))) {
        fprintf(stderr, "Error: weight method needs one weight array per particle set\n");
        free(rupp); free(rupp_sqr);
        return EXIT_FAILURE;
    }
    int status = cfb_upload(0, (int)sizeof(REAL), ND1, X1, Y1, Z1, W1, NULL, NULL);
    if (status == 0 && !autocorr) status = cfb_upload(1, (int)sizeof(REAL), ND2, X2, Y2, Z2, W2, NULL, NULL);
    if (status) {
        free(rupp); free(rupp_sqr);
        return EXIT_FAILURE;
    }
    const double t_up1 = now_ms();

    /* ---- extents, wrap, refine heuristics ---- */
    REAL lo[3], hi[3], wrap[3], maxsize[3];
    int periodic[3];
    REAL pimax = 0, mu_max = 0;
    int npibin = 0;
    double max_sep[3] = {-1.0, -1.0, -1.0};
    if (is_box) {
        for (int a = 0; a < 3; a++) { /* countpairs_xi_impl.c.src:222-224 */
            lo[a] = 0.0;
            hi[a] = boxsize;
            wrap[a] = boxsize;
            periodic[a] = 1;
        }
        if (mode == CFB_XI) {
            if (get_bin_refine_scheme(options) == BINNING_DFL && rpmax < 0.05 * boxsize)
                for (int i = 0; i < 3; i++) options->bin_refine_factors[i] = 1; /* xi_impl:213-219 */
            maxsize[0] = maxsize[1] = maxsize[2] = rpmax;
            max_sep[0] = (double)(REAL)rpmax;
        } else {
            if (get_bin_refine_scheme(options) == BINNING_DFL) { /* wp_impl:238-247 */
                if (rpmax < 0.05 * boxsize) options->bin_refine_factors[0] = options->bin_refine_factors[1] = 1;
                if (pimax_in < 0.05 * boxsize) options->bin_refine_factors[2] = 1;
            }
            pimax = pimax_in;
            maxsize[0] = maxsize[1] = rpmax;
            maxsize[2] = pimax_in;
            max_sep[1] = (double)(REAL)rpmax;
            max_sep[2] = (double)(REAL)pimax_in;
        }
    } else {
        double lohi[6] = {H_MAXPOS, H_MAXPOS, H_MAXPOS, -H_MAXPOS, -H_MAXPOS, -H_MAXPOS};
        status = cfb_extent(0, 0, lohi); /* get_max_min_DOUBLE, countpairs_impl.c.src:210-224 */
        if (status == 0 && !autocorr) status = cfb_extent(1, 0, lohi);
        if (status) {
            free(rupp); free(rupp_sqr);
            return EXIT_FAILURE;
        }
        for (int a = 0; a < 3; a++) {
            lo[a] = (REAL)lohi[a];
            hi[a] = (REAL)lohi[3 + a];
        }
        const int want_periodic = is_mocks ? 0 : options->periodic; /* a survey footprint never wraps */
        if (want_periodic && options->boxsize == BOXSIZE_NOTGIVEN) { /* countpairs_impl.c.src:231-235 */
            fprintf(stderr, "boxsize = %g must be specified with periodic wrap. Please specify a non-zero boxsize, or zero to detect the particle extent, or -1 to make a dimension non-periodic.\n",
                    options->boxsize);
            free(rupp); free(rupp_sqr);
            return EXIT_FAILURE;
        }
        const double bs[3] = {options->boxsize_x,
                              options->boxsize_y == BOXSIZE_NOTGIVEN ? options->boxsize : options->boxsize_y,
                              options->boxsize_z == BOXSIZE_NOTGIVEN ? options->boxsize : options->boxsize_z};
        for (int a = 0; a < 3; a++) { /* countpairs_impl.c.src:244-251 */
            periodic[a] = want_periodic && bs[a] >= 0;
            wrap[a] = periodic[a] ? (bs[a] > 0 ? bs[a] : (hi[a] - lo[a])) : 0.;
        }
        if (mode == CFB_DD) {
            pimax = (REAL)rpmax;
            if (get_bin_refine_scheme(options) == BINNING_DFL) { /* countpairs_impl.c.src:272-282 */
                if (rpmax < 0.05 * wrap[0]) options->bin_refine_factors[0] = 1;
                if (rpmax < 0.05 * wrap[1]) options->bin_refine_factors[1] = 1;
                if (pimax < 0.05 * wrap[2]) options->bin_refine_factors[2] = 1;
            }
            maxsize[0] = maxsize[1] = maxsize[2] = rpmax;
            max_sep[0] = (double)(REAL)rpmax;
        } else if (mode == CFB_RPPI) {
            pimax = pimax_in;
            npibin = (int)pimax_in; /* countpairs_rp_pi_impl.c.src:188 */
            if (npibin < 1) {
                fprintf(stderr, "Error: pimax = %lf must be at least 1 (the pi bins are 1 unit wide)\n", pimax_in);
                free(rupp); free(rupp_sqr);
                return EXIT_FAILURE;
            }
            if (get_bin_refine_scheme(options) == BINNING_DFL) { /* rp_pi_impl:277-287 */
                if (rpmax < 0.05 * wrap[0]) options->bin_refine_factors[0] = 1;
                if (rpmax < 0.05 * wrap[1]) options->bin_refine_factors[1] = 1;
                if (pimax_in < 0.05 * wrap[2]) options->bin_refine_factors[2] = 1;
            }
            maxsize[0] = maxsize[1] = rpmax;
            maxsize[2] = pimax_in;
            max_sep[1] = (double)(REAL)rpmax;
            max_sep[2] = (double)(REAL)pimax_in;
        } else if (is_mocks) {
            REAL max_sep_r;
            double heur; /* what the 0.05 * extent rule compares against */
            if (mode == CFB_RPPI_MOCKS) { /* countpairs_rp_pi_mocks_impl.c.src:281, 307-308 */
                pimax = pimax_in;
                npibin = (int)pimax;
                if (npibin < 1) {
                    fprintf(stderr, "Error: pimax = %lf must be at least 1 (the pi bins are 1 unit wide)\n", pimax_in);
                    free(rupp); free(rupp_sqr);
                    return EXIT_FAILURE;
                }
                const REAL sqr_max_sep = rpmax * rpmax + pimax * pimax;
                max_sep_r = H_SQRT(sqr_max_sep);
                heur = (double)max_sep_r;
            } else { /* countpairs_s_mu_mocks_impl.c.src:408, 424-432, 438 */
                mu_max = (REAL)max_mu;
                max_sep_r = (REAL)rpmax;
                heur = rpmax;
            }
            if (get_bin_refine_scheme(options) == BINNING_DFL) {
                for (int a = 0; a < 3; a++)
                    if (heur < 0.05 * (hi[a] - lo[a])) options->bin_refine_factors[a] = 1;
            }
            maxsize[0] = maxsize[1] = maxsize[2] = max_sep_r;
            max_sep[0] = (double)max_sep_r;
        } else { /* CFB_SMU: countpairs_s_mu_impl.c.src:231-234 */
            mu_max = (REAL)max_mu;
            pimax = rpmax * mu_max;
            maxsize[0] = maxsize[1] = rpmax;
            maxsize[2] = pimax;
            max_sep[0] = (double)(REAL)rpmax;
            max_sep[2] = (double)pimax;
        }
    }

    /* ---- lattice sizes (gridlink's sizing) and the boost heuristic ---- */
    HFN(cf_mesh) M;
    if (HFN(cf_mesh_for)(lo, hi, maxsize, wrap, options->bin_refine_factors, options->max_cells_per_dim, &M)) {
        free(rupp); free(rupp_sqr);
        return EXIT_FAILURE;
    }
    if (mode != CFB_SMU) { /* countpairs_impl.c.src:298-332; DDsmu's boost multiplies by BOOST_BIN_REF=1 (no-op) */
        double avg_np = ((double)ND1) / ((double)M.nmesh[0] * M.nmesh[1] * M.nmesh[2]);
        if (mode == CFB_RPPI_MOCKS) { /* the larger set, with ND2 as passed even when autocorr (rp_pi_mocks_impl:442-444) */
            const double avg_np2 = ((double)ND2) / ((double)M.nmesh[0] * M.nmesh[1] * M.nmesh[2]);
            if (avg_np2 > avg_np) avg_np = avg_np2;
        }
        const int max_nmesh = (int)fmax(M.nmesh[0], fmax(M.nmesh[1], M.nmesh[2]));
        if ((max_nmesh <= BOOST_CELL_THRESH || avg_np >= BOOST_NUMPART_THRESH) &&
            max_nmesh < options->max_cells_per_dim && get_bin_refine_scheme(options) == BINNING_DFL) {
            for (int i = 0; i < 2; i++) options->bin_refine_factors[i] += BOOST_BIN_REF;
            if (HFN(cf_mesh_for)(lo, hi, maxsize, wrap, options->bin_refine_factors, options->max_cells_per_dim, &M)) {
                free(rupp); free(rupp_sqr);
                return EXIT_FAILURE;
            }
        }
    }
    if (options->verbose)
        fprintf(stderr, "corrfunc_b200> reference lattice [nmesh_x, nmesh_y, nmesh_z] = %d,%d,%d refine = %d,%d,%d\n",
                M.nmesh[0], M.nmesh[1], M.nmesh[2], options->bin_refine_factors[0], options->bin_refine_factors[1],
                options->bin_refine_factors[2]);

    /* ---- device job ---- */
    cfb_binning B;
    memset(&B, 0, sizeof(B));
    B.mode = mode;
    B.prec = (int)sizeof(REAL);
    B.autocorr = autocorr;
    B.nedges = nrpbin;
    B.edges = rupp_sqr;
    B.pimax = (double)pimax;
    B.need_avg = options->need_avg_sep ? 1 : 0;
    B.need_weights = need_weightavg;
    int64_t nslots = nrpbin;
    int n2 = 0;
    if (mode == CFB_RPPI || mode == CFB_RPPI_MOCKS) { /* countpairs_rp_pi_kernels.c.src:65-66, rp_pi_mocks_kernels:64-65 */
        const REAL dpi = pimax / npibin;
        const REAL inv_dpi = 1.0 / dpi;
        B.npibin = npibin;
        B.inv_dpi = (double)inv_dpi;
        n2 = npibin;
        nslots = (int64_t)(npibin + 1) * (nrpbin + 1);
    } else if (mode == CFB_SMU || mode == CFB_SMU_MOCKS) { /* countpairs_s_mu_kernels.c.src:65-68, s_mu_mocks_kernels:52-61 */
        const REAL sqr_mumax = mu_max * mu_max;
        const REAL dmu = mu_max / (REAL)nmu_bins;
        const REAL inv_dmu = 1.0 / dmu;
        B.nmu_bins = nmu_bins;
        B.sqr_mumax = (double)sqr_mumax;
        B.inv_dmu = (double)inv_dmu;
        n2 = nmu_bins;
        nslots = (int64_t)(nmu_bins + 1) * (nrpbin + 1);
    }
    B.nslots = nslots;
    cfb_box_lattice L;
    memset(&L, 0, sizeof(L));
    for (int a = 0; a < 3; a++) {
        L.nmesh[a] = M.nmesh[a];
        L.refine[a] = options->bin_refine_factors[a];
        L.periodic[a] = periodic[a];
        L.lo[a] = (double)lo[a];
        L.inv[a] = (double)M.inv[a];
        L.wrap[a] = (double)wrap[a];
        L.max_sep[a] = options->enable_min_sep_opt ? max_sep[a] : -1.0;
    }
    uint64_t *npairs = calloc((size_t)nslots, sizeof(uint64_t));
    double *sum_sep = calloc((size_t)nslots, sizeof(double));
    double *sum_w = calloc((size_t)nslots, sizeof(double));
    if (!npairs || !sum_sep || !sum_w) {
        free(npairs); free(sum_sep); free(sum_w); free(rupp); free(rupp_sqr);
        return EXIT_FAILURE;
    }
    cfb_hist H = {npairs, sum_sep, sum_w};
    cfb_stats dst;
    memset(&dst, 0, sizeof(dst));
    {
        cf_sig_t prev_handlers[3];
        cf_signals_install(prev_handlers);
        if (ND1 > 0 && (autocorr || ND2 > 0)) status = cfb_count_box(&B, &L, &H, &dst);
        if (cf_signals_restore(prev_handlers)) status = 1; /* interrupted: EXIT_FAILURE (countpairs_impl.c.src:554-569) */
    }
    status = reduce_across_ranks(status, npairs, sum_sep, sum_w, nslots);
    if (status) {
        free(npairs); free(sum_sep); free(sum_w); free(rupp); free(rupp_sqr);
        return EXIT_FAILURE;
    }

    /* ---- epilogue (countpairs_impl.c.src:609-664 and siblings) ---- */
    if (autocorr == 1) {
        for (int64_t i = 0; i < nslots; i++) {
            npairs[i] *= 2;
            sum_sep[i] *= 2.0;
            sum_w[i] *= 2.0;
        }
        if (rupp[0] <= 0.0) {
            const int64_t first = (mode == CFB_RPPI) ? (npibin + 1) : (mode == CFB_SMU ? (nmu_bins + 1) : 1);
            npairs[first] += (uint64_t)ND1;
            if (need_weightavg) { /* always index 1, also in the 2-D layouts (rp_pi_impl:641, s_mu_impl:648) */
                void *tofree = NULL;
                const REAL *w = (const REAL *)HFN(host_copy)(W1, ND1, &tofree);
                if (!w) {
                    fprintf(stderr, "Error: could not read the weights back from the device\n");
                    free(npairs); free(sum_sep); free(sum_w); free(rupp); free(rupp_sqr);
                    return EXIT_FAILURE;
                }
                for (int64_t j = 0; j < ND1; j++) sum_w[1] += (double)(REAL)(w[j] * w[j]);
                free(tofree);
            }
        }
    }
    for (int64_t i = 0; i < nslots; i++) {
        if (npairs[i] > 0) {
            sum_sep[i] /= (double)npairs[i];
            sum_w[i] /= (double)npairs[i];
        }
        if (!options->need_avg_sep) sum_sep[i] = 0.0;
        if (!need_weightavg) sum_w[i] = 0.0;
    }

    out->nbin = nrpbin;
    out->n2 = n2;
    out->npairs = npairs;
    out->avg = sum_sep;
    out->wavg = sum_w;
    out->rupp = malloc(sizeof(double) * (size_t)nrpbin);
    for (int i = 0; i < nrpbin; i++) out->rupp[i] = rupp[i];
    out->cf = NULL;

    if (is_box) { /* xi: countpairs_xi_impl.c.src:581-625 ; wp: countpairs_wp_impl.c.src:619-664 */
        out->cf = calloc((size_t)nrpbin, sizeof(double));
        const int64_t ND = ND1;
        REAL weightsum = (REAL)ND, weight_sqr_sum = (REAL)ND;
        if (need_weightavg && extra->weight_method == PAIR_PRODUCT) {
            void *tofree = NULL;
            const REAL *weights = (const REAL *)HFN(host_copy)(W1, ND1, &tofree);
            weightsum = 0;
            for (int64_t j = 0; weights && j < ND; j++) {
                weightsum += weights[j];
                weight_sqr_sum += weights[j] * weights[j];
            }
            free(tofree);
        }
        const REAL prefac = weightsum * (weightsum - weightsum / ND) / (boxsize * boxsize * boxsize);
        REAL rlow = 0.0;
        const REAL twice_pimax = 2.0 * pimax_in;
        for (int i = 0; i < nrpbin; i++) {
            REAL weight0 = (REAL)npairs[i];
            if (need_weightavg && extra->weight_method == PAIR_PRODUCT) weight0 *= out->wavg[i];
            if (mode == CFB_XI) {
                const REAL vol = 4.0 / 3.0 * M_PI * (rupp[i] * rupp[i] * rupp[i] - rlow * rlow * rlow);
                if (vol > 0.0) {
                    REAL weightrandom = prefac * vol;
                    if (rlow <= 0.) weightrandom += weight_sqr_sum;
                    out->cf[i] = (weight0 / weightrandom - 1.0);
                } else
                    out->cf[i] = -2.0;
            } else {
                const REAL vol = M_PI * (rupp[i] * rupp[i] - rlow * rlow) * twice_pimax;
                if (vol > 0.0) {
                    REAL weightrandom = prefac * vol;
                    if (rlow <= 0.) weightrandom += weight_sqr_sum;
                    out->cf[i] = (weight0 / weightrandom - 1) * twice_pimax;
                } else
                    out->cf[i] = -2.0 * twice_pimax;
            }
            rlow = rupp[i];
        }
    }
    free(rupp);
    free(rupp_sqr);

    g_stats.dev = dst;
    g_stats.ms_upload = t_up1 - t_up0;
    for (int a = 0; a < 3; a++) {
        g_stats.nmesh[a] = M.nmesh[a];
        g_stats.refine[a] = options->bin_refine_factors[a];
    }
    reset_bin_refine_factors(options); /* countpairs_impl.c.src:697 */
    const double t_end = now_ms();
    g_stats.ms_host_total = t_end - t_start;
    if (options->c_api_timer) {
        /* seconds everywhere except wp, which reports nanoseconds (countpairs_wp_impl.c.src:672-676) */
        options->c_api_time = (mode == CFB_WP) ? (t_end - t_start) * 1.0e6 : (t_end - t_start) * 1.0e-3;
    }
    return EXIT_SUCCESS;
}

/* ========================================================================================== */
/* DDrppi_mocks / DDsmu_mocks: (RA, DEC, comoving distance) -> Cartesian on the host, then the box driver   */

/* cz -> comoving distance of one particle set (countpairs_rp_pi_mocks_impl.c.src:326-362).  czmax is the maximum over
 * BOTH sets of a cross-correlation: it only sets how far the table reaches (its entries do not depend on it). */
static int HFN(cf_cz_to_dist)(const int64_t N, const REAL *cz, const REAL czmax, const int cosmology, REAL *D)
{
    const REAL inv_speed_of_light = 1.0 / CF_SPEED_OF_LIGHT;
    const double zmax = czmax * inv_speed_of_light + 0.01;
    const int workspace_size = 10000;
    double *zc = calloc((size_t)workspace_size, sizeof(double)), *dc = calloc((size_t)workspace_size, sizeof(double));
    if (!zc || !dc) {
        free(zc); free(dc);
        return EXIT_FAILURE;
    }
    const int Nzdc = corrfunc_b200_cosmo_dist_table(zmax, workspace_size, zc, dc, cosmology);
    int64_t bad = -1;
    if (Nzdc >= 2) {
#if defined(_OPENMP)
#pragma omp parallel for schedule(static)
#endif
        for (int64_t i = 0; i < N; i++) {
            double y = 0.0;
            if (cf_interp_linear(zc, dc, Nzdc, cz[i] * inv_speed_of_light, &y)) {
#if defined(_OPENMP)
#pragma omp critical
#endif
                bad = i;
            }
            D[i] = y;
        }
    }
    free(zc);
    free(dc);
    if (Nzdc < 2) return EXIT_FAILURE;
    if (bad >= 0) {
        fprintf(stderr, "Error: cz[%" PRId64 "] = %g is outside the redshift range [1e-4, %g] of the distance table "
                        "(the reference's GSL interpolation aborts here)\n", bad, (double)cz[bad], zmax);
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}

/* check_ra_dec_cz_DOUBLE (mocks/DDrppi_mocks/countpairs_rp_pi_mocks_impl.c.src:43-110): RA in [-180,180] and DEC in
 * [0,180] are shifted, and redshifts passed as cz are scaled, IN PLACE like the reference does. */
static int HFN(cf_check_ra_dec_cz)(const int64_t N, REAL *phi, REAL *theta, REAL *cz, const int is_comoving_dist)
{
    if (N == 0) return EXIT_SUCCESS;
    if (phi == NULL || theta == NULL || cz == NULL) {
        fprintf(stderr, "Input arrays can not be NULL. Have RA = %p DEC = %p cz = %p\n", (void *)phi, (void *)theta, (void *)cz);
        return EXIT_FAILURE;
    }
    int fix_ra = 0, fix_dec = 0, fix_cz = 0;
    const REAL max_cz_threshold = 10.0; /* a maximum below this means redshifts were passed instead of cz */
    REAL max_cz = 0.0;
    for (int64_t i = 0; i < N; i++) {
        if (cz[i] > max_cz) max_cz = cz[i];
        if (phi[i] < 0.0) fix_ra = 1;
        if (theta[i] > 90.0) fix_dec = 1;
        if (theta[i] > 180) {
            fprintf(stderr, "theta[%" PRId64 "] = %lf should be less than 180 deg\n", i, (double)theta[i]);
            return EXIT_FAILURE;
        }
    }
    if (fix_ra) fprintf(stderr, "%s> Out of range values found for ra. Expected ra to be in the range [0.0,360.0]. Found ra values in [-180,180] -- fixing that\n", __func__);
    if (fix_dec) fprintf(stderr, "%s> Out of range values found for dec. Expected dec to be in the range [-90.0,90.0]. Found dec values in [0,180] -- fixing that\n", __func__);
    if ((max_cz < max_cz_threshold) && (is_comoving_dist == 0)) fix_cz = 1;
    if (fix_cz)
        fprintf(stderr, "%s> Out of range values found for cz. Expected input to be `cz' but found `z' instead. max_cz (found in input) = %g threshold = %g\n",
                __func__, (double)max_cz, (double)max_cz_threshold);
    if (fix_ra || fix_dec || fix_cz)
        for (int64_t i = 0; i < N; i++) {
            if (fix_ra) phi[i] += (REAL)180.0;
            if (fix_dec) theta[i] -= (REAL)90.0;
            if (fix_cz) cz[i] *= (REAL)CF_SPEED_OF_LIGHT; /* input was z -> convert to cz */
        }
    return EXIT_SUCCESS;
}

static int HFN(cf_mocks)(const int mode, const int64_t ND1, void *vra1, void *vdec1, void *vd1, const int64_t ND2,
                         void *vra2, void *vdec2, void *vd2, const int numthreads, const int autocorr,
                         const char *binfile, const double pimax, const double max_mu, const int nmu_bins,
                         const int cosmology, struct config_options *options, struct extra_options *extra,
                         cf_box_out *out)
{
    if (options->float_type != sizeof(REAL)) {
        fprintf(stderr, "ERROR: In %s> Can only handle arrays of size=%zu. Got an array of size = %zu\n", __func__,
                sizeof(REAL), options->float_type);
        return EXIT_FAILURE;
    }
    if (ND1 == 0 || (autocorr == 0 && ND2 == 0)) return EXIT_SUCCESS; /* rp_pi_mocks_impl:243-245: results untouched */
    REAL *ra[2] = {(REAL *)vra1, (REAL *)vra2}, *dec[2] = {(REAL *)vdec1, (REAL *)vdec2}, *D[2] = {(REAL *)vd1, (REAL *)vd2};
    const int64_t N[2] = {ND1, ND2};
    const int nsets = autocorr ? 1 : 2;
    for (int s = 0; s < nsets; s++)
        if (HFN(cf_check_ra_dec_cz)(N[s], ra[s], dec[s], D[s], options->is_comoving_dist)) return EXIT_FAILURE;
    if (!(cosmology == 1 || cosmology == 2)) { /* init_cosmology, utils/cosmology_params.c:21-54 */
        fprintf(stderr, "ERROR: In %s> Cosmology=%d not implemented\n", "init_cosmology", cosmology);
        return EXIT_FAILURE;
    }
    REAL *conv[2] = {NULL, NULL};
    if (options->is_comoving_dist == 0) { /* rp_pi_mocks_impl:326-362: cz -> comoving distance through the table */
        REAL czmax = 0.0;
        for (int s = 0; s < nsets; s++)
            for (int64_t i = 0; i < N[s]; i++)
                if (D[s][i] > czmax) czmax = D[s][i];
        for (int s = 0; s < nsets; s++) {
            conv[s] = malloc(sizeof(REAL) * (size_t)(N[s] > 0 ? N[s] : 1));
            if (!conv[s] || HFN(cf_cz_to_dist)(N[s], D[s], czmax, cosmology, conv[s])) {
                free(conv[0]); free(conv[1]);
                return EXIT_FAILURE;
            }
            D[s] = conv[s];
        }
    }
    REAL *xyz[2][3] = {{NULL, NULL, NULL}, {NULL, NULL, NULL}};
    for (int s = 0; s < nsets; s++) {
        for (int a = 0; a < 3; a++) {
            /* pinned, persistent across calls (no page faults, full-rate upload); owned by the device context */
            xyz[s][a] = cfb_host_scratch(3 * s + a, sizeof(REAL) * (size_t)(N[s] > 0 ? N[s] : 1));
            if (!xyz[s][a]) {
                fprintf(stderr, "Error: could not get host staging memory for %" PRId64 " positions: %s\n", N[s], cfb_last_error());
                free(conv[0]); free(conv[1]);
                return EXIT_FAILURE;
            }
        }
        REAL *X = xyz[s][0], *Y = xyz[s][1], *Z = xyz[s][2];
        const REAL *r = ra[s], *dc = dec[s], *dist = D[s];
        const int64_t n = N[s];
        /* rp_pi_mocks_impl:370-391, element by element with glibc trig: independent of the thread count */
#if defined(_OPENMP)
#pragma omp parallel for schedule(static)
#endif
        for (int64_t i = 0; i < n; i++) {
            X[i] = dist[i] * H_COSD(dc[i]) * H_COSD(r[i]);
            Y[i] = dist[i] * H_COSD(dc[i]) * H_SIND(r[i]);
            Z[i] = dist[i] * H_SIND(dc[i]);
        }
    }
    free(conv[0]);
    free(conv[1]);
    return HFN(cf_box)(mode, ND1, xyz[0][0], xyz[0][1], xyz[0][2], ND2, xyz[1][0], xyz[1][1], xyz[1][2], numthreads,
                       autocorr, binfile, pimax, max_mu, nmu_bins, 0.0, options, extra, out);
}

/* ========================================================================================== */
/* vpf_mocks: counts-in-spheres (mocks/vpf_mocks/countspheres_mocks_impl.c.src:206-638)                          */

static int HFN(cf_vpf_mocks)(const int64_t Ngal, void *vra, void *vdec, void *vcz, const int64_t Nran, void *vrra,
                             void *vrdec, void *vrcz, const int threshold_neighbors, const REAL rmax, const int nbin,
                             const int nc, const int num_pN, const char *centers_file, const int cosmology,
                             struct config_options *options, cf_vpf_out *out)
{
    const double t_start = now_ms();
    if (options->float_type != sizeof(REAL)) {
        fprintf(stderr, "ERROR: In %s> Can only handle arrays of size=%zu. Got an array of size = %zu\n", __func__,
                sizeof(REAL), options->float_type);
        return EXIT_FAILURE;
    }
    if (!(rmax > 0.0)) { fprintf(stderr, "rmax=%lf has to be positive", (double)rmax); return EXIT_FAILURE; }
    if (nbin < 1) { fprintf(stderr, "Number of bins=%d has to be at least 1", nbin); return EXIT_FAILURE; }
    if (nbin == 1) {
        /* the reference's vectorised shell loop never runs for a single shell and drops every particle it handles while its
         * scalar remainder loop counts them (vpf_mocks_kernels.c.src:63-92 vs :100-120): its answer depends on the cell
         * occupancies modulo the vector width.  Nothing meaningful to reproduce -- refuse instead of returning P0 = 1. */
        fprintf(stderr, "corrfunc_b200> counts-in-spheres with a single radial bin is not supported (use nbin >= 2)\n");
        return EXIT_FAILURE;
    }
    if (nc < 1) { fprintf(stderr, "Number of spheres=%d has to be at least 1", nc); return EXIT_FAILURE; }
    if (num_pN < 1) { fprintf(stderr, "Number of pN's=%d requested must be at least 1", num_pN); return EXIT_FAILURE; }
    if (!(cosmology == 1 || cosmology == 2)) {
        fprintf(stderr, "ERROR: In %s> Cosmology=%d not implemented\n", "init_cosmology", cosmology);
        return EXIT_FAILURE;
    }
    options->periodic = 0;
    const REAL *RA = (const REAL *)vra, *DEC = (const REAL *)vdec, *CZ = (const REAL *)vcz;
    const REAL *RRA = (const REAL *)vrra, *RDEC = (const REAL *)vrdec, *RCZ = (const REAL *)vrcz;

    /* ---- centres file: usable when its first radius covers rmax and it has at least nc lines (:256-283) ---- */
    int need_randoms = 1;
    FILE *fpcen = fopen(centers_file, "r");
    if (fpcen != NULL) {
        double rr = 0.0;
        if (fscanf(fpcen, "%*f %*f %*f %lf", &rr) != 1) {
            fprintf(stderr, "Could not read max. sphere radius from the centers file");
            fclose(fpcen);
            return EXIT_FAILURE;
        }
        if (rr >= rmax && count_data_lines(centers_file) >= nc) {
            need_randoms = 0;
            rewind(fpcen);
        } else {
            fclose(fpcen);
            fpcen = NULL;
        }
    }
    if (need_randoms) {
        fpcen = fopen(centers_file, "w"); /* the reference rewrites the file with the centres it places */
        if (fpcen == NULL) {
            fprintf(stderr, "Error: could not open centers file `%s' for writing\n", centers_file);
            return EXIT_FAILURE;
        }
    }

    int status = EXIT_FAILURE;
    REAL *gal[3] = {NULL, NULL, NULL}, *ran[3] = {NULL, NULL, NULL}, *Dg = NULL, *Dr = NULL;
    REAL *cen[3] = {NULL, NULL, NULL};
    uint32_t *counts = NULL, *ngb = NULL;
    uint64_t *pN = NULL;
    double *edges = NULL;
    const int64_t nr_used = need_randoms ? Nran : 0;
    for (int a = 0; a < 3; a++) {
        gal[a] = malloc(sizeof(REAL) * (size_t)(Ngal > 0 ? Ngal : 1));
        ran[a] = malloc(sizeof(REAL) * (size_t)(nr_used > 0 ? nr_used : 1));
        cen[a] = malloc(sizeof(REAL) * (size_t)nc);
    }
    Dg = malloc(sizeof(REAL) * (size_t)(Ngal > 0 ? Ngal : 1));
    Dr = malloc(sizeof(REAL) * (size_t)(nr_used > 0 ? nr_used : 1));
    counts = malloc(sizeof(uint32_t) * (size_t)nc * (size_t)nbin);
    pN = calloc((size_t)nbin * (size_t)num_pN, sizeof(uint64_t));
    edges = malloc(sizeof(double) * (size_t)nbin);
    if (!gal[0] || !gal[1] || !gal[2] || !ran[0] || !ran[1] || !ran[2] || !cen[0] || !cen[1] || !cen[2] || !Dg || !Dr ||
        !counts || !pN || !edges) {
        fprintf(stderr, "Error: In %s> out of memory\n", __func__);
        goto done;
    }

    /* ---- distances (:285-317), Cartesian positions and the shift into [0, 2 rcube] (:338-392) ---- */
    if (options->is_comoving_dist == 0) {
        REAL czmax = 0.0;
        for (int64_t i = 0; i < Ngal; i++) if (CZ[i] > czmax) czmax = CZ[i];
        for (int64_t i = 0; i < nr_used; i++) if (RCZ[i] > czmax) czmax = RCZ[i];
        if (HFN(cf_cz_to_dist)(Ngal, CZ, czmax, cosmology, Dg)) goto done;
        if (nr_used > 0 && HFN(cf_cz_to_dist)(nr_used, RCZ, czmax, cosmology, Dr)) goto done;
    } else {
        memcpy(Dg, CZ, sizeof(REAL) * (size_t)Ngal);
        if (nr_used > 0) memcpy(Dr, RCZ, sizeof(REAL) * (size_t)nr_used);
    }
    REAL rcube = 0.0;
    for (int64_t i = 0; i < Ngal; i++) {
        const REAL dc = Dg[i];
        if (dc > rcube) rcube = dc;
        gal[0][i] = dc * H_COSD(DEC[i]) * H_COSD(RA[i]);
        gal[1][i] = dc * H_COSD(DEC[i]) * H_SIND(RA[i]);
        gal[2][i] = dc * H_SIND(DEC[i]);
    }
    for (int64_t i = 0; i < nr_used; i++) {
        const REAL dc = Dr[i];
        if (dc > rcube) rcube = dc;
        ran[0][i] = dc * H_COSD(RDEC[i]) * H_COSD(RRA[i]);
        ran[1][i] = dc * H_COSD(RDEC[i]) * H_SIND(RRA[i]);
        ran[2][i] = dc * H_SIND(RDEC[i]);
    }
    rcube = rcube + 1.;
    for (int a = 0; a < 3; a++) {
        for (int64_t i = 0; i < Ngal; i++) gal[a][i] += rcube;
        for (int64_t i = 0; i < nr_used; i++) ran[a][i] += rcube;
    }
    rcube = 2.0 * rcube;
    const double cube_lo[3] = {0.0, 0.0, 0.0}, cube_ext[3] = {(double)rcube, (double)rcube, (double)rcube};
    const int no_wrap[3] = {0, 0, 0};

    if (cfb_upload(0, (int)sizeof(REAL), Ngal, gal[0], gal[1], gal[2], NULL, NULL, NULL)) goto done;
    if (need_randoms && cfb_upload(1, (int)sizeof(REAL), nr_used, ran[0], ran[1], ran[2], NULL, NULL, NULL)) goto done;

    /* ---- the centres (:470-505) ---- */
    const REAL rmax_sqr = rmax * rmax;
    int isucceed = 0;
    if (!need_randoms) {
        char buffer[10000];
        for (; isucceed < nc; isucceed++) {
            double rr = 0.0, c3[3];
            if (fgets(buffer, (int)sizeof(buffer), fpcen) == NULL) {
                fprintf(stderr, "ERROR: Could not read-in co-ordinates for the centers of the randoms spheres from file %s\n", centers_file);
                goto done;
            }
            /* the reference parses straight into REAL ("%f" / "%lf") */
            const int nitems = REAL_IS_DOUBLE ? sscanf(buffer, "%lf %lf %lf %lf", &c3[0], &c3[1], &c3[2], &rr) : 0;
            if (REAL_IS_DOUBLE) {
                if (nitems != 4) { fprintf(stderr, "ERROR in parsing centers file: buffer = `%s' \n", buffer); goto done; }
                for (int a = 0; a < 3; a++) cen[a][isucceed] = (REAL)c3[a];
            } else {
                float f3[3];
                if (sscanf(buffer, "%f %f %f %lf", &f3[0], &f3[1], &f3[2], &rr) != 4) {
                    fprintf(stderr, "ERROR in parsing centers file: buffer = `%s' \n", buffer);
                    goto done;
                }
                for (int a = 0; a < 3; a++) cen[a][isucceed] = (REAL)f3[a];
            }
            if (!(rr >= rmax)) { fprintf(stderr, "Rmax from the center file is >= rmax"); goto done; }
        }
    } else {
        /* randoms with more than threshold_neighbors randoms (itself included) within rmax, in input order, until nc
         * are placed (:478-490 with count_neighbors :140-204); the GPU counts a chunk of candidates at a time */
        const int64_t chunk = 1 << 16;
        ngb = malloc(sizeof(uint32_t) * (size_t)chunk);
        if (!ngb) goto done;
        int first = 1;
        for (int64_t base = 0; base < Nran && isucceed < nc; base += chunk) {
            const int64_t m = (Nran - base) < chunk ? (Nran - base) : chunk;
            if (cfb_count_spheres(1, (int)sizeof(REAL), cube_lo, cube_ext, no_wrap, cube_lo, first, m, ran[0] + base,
                                  ran[1] + base, ran[2] + base, (double)rmax, (double)rmax_sqr, 1, NULL, 0, ngb)) goto done;
            first = 0;
            for (int64_t i = 0; i < m && isucceed < nc; i++)
                if ((int64_t)ngb[i] > (int64_t)threshold_neighbors) {
                    for (int a = 0; a < 3; a++) cen[a][isucceed] = ran[a][base + i];
                    fprintf(fpcen, "%lf \t %lf \t %lf \t %lf\n", (double)ran[0][base + i], (double)ran[1][base + i],
                            (double)ran[2][base + i], (double)rmax);
                    isucceed++;
                }
        }
    }
    if (isucceed <= 0) {
        fprintf(stderr, "ERROR: Could not place even a single sphere within the volume. Please reduce the radius of the sphere (currently set to %lf)\n", (double)rmax);
        goto done;
    } else if (isucceed < nc) {
        fprintf(stderr, "WARNING: Could only place `%d' out of requested `%d' spheres. Increase the random-sample size might improve the situation\n", isucceed, nc);
    }

    /* ---- shell counts per centre on the GPU, then cumulative counts and pN on the host (:565-580) ---- */
    {
        const REAL rstep = rmax / (REAL)nbin; /* vpf_mocks_kernels.c.src:38-46 */
        for (int k = 0; k < nbin; k++) {
            const REAL e = (k + 1) * rstep * rstep * (k + 1);
            edges[k] = (double)e;
        }
        if (cfb_count_spheres(0, (int)sizeof(REAL), cube_lo, cube_ext, no_wrap, cube_lo, 1, isucceed, cen[0], cen[1], cen[2],
                              (double)rmax, (double)rmax_sqr, nbin, edges, 1, counts)) goto done;
        for (int64_t c = 0; c < isucceed; c++) {
            uint64_t cum = 0;
            for (int k = 0; k < nbin; k++) {
                cum += counts[c * nbin + k];
                if (cum < (uint64_t)num_pN) pN[(size_t)k * num_pN + cum]++;
            }
        }
    }
    out->nbin = nbin;
    out->num_pN = num_pN;
    out->pN = calloc((size_t)nbin, sizeof(double *));
    if (!out->pN) goto done;
    {
        const REAL inv_nc = ((REAL)1.0) / (REAL)isucceed; /* :616-621 */
        for (int k = 0; k < nbin; k++) {
            out->pN[k] = malloc(sizeof(double) * (size_t)num_pN);
            if (!out->pN[k]) goto done; /* rows allocated so far are released by the caller's free_results */
            for (int i = 0; i < num_pN; i++) out->pN[k][i] = (double)(REAL)((REAL)pN[(size_t)k * num_pN + i] * inv_nc);
        }
    }
    reset_bin_refine_factors(options);
    if (options->c_api_timer) options->c_api_time = (now_ms() - t_start) * 1.0e-3;
    status = EXIT_SUCCESS;
done:
    if (fpcen) fclose(fpcen);
    for (int a = 0; a < 3; a++) { free(gal[a]); free(ran[a]); free(cen[a]); }
    free(Dg); free(Dr); free(counts); free(ngb); free(pN); free(edges);
    return status;
}

/* ========================================================================================== */
/* theory vpf: counts-in-spheres in a simulation box (theory/vpf/countspheres_impl.c.src:138-479)                 */

static int HFN(cf_vpf_theory)(const int64_t np, void *vX, void *vY, void *vZ, const double rmax, const int nbin, const int nc,
                              const int num_pN, unsigned long seed, struct config_options *options, cf_vpf_out *out)
{
    const double t_start = now_ms();
    if (options->float_type != sizeof(REAL)) {
        fprintf(stderr, "ERROR: In %s> Can only handle arrays of size=%zu. Got an array of size = %zu\n", __func__,
                sizeof(REAL), options->float_type);
        return EXIT_FAILURE;
    }
    if (nbin == 1) {
        fprintf(stderr, "corrfunc_b200> counts-in-spheres with a single radial bin is not supported (use nbin >= 2)\n");
        return EXIT_FAILURE;
    }
    if (!(rmax > 0.0 && nbin > 0 && nc > 0 && num_pN > 0)) {
        fprintf(stderr, "Error: Invalid input parameters. Expected rmax > 0, number of bins > 0, number of random spheres > 0, number of pN's to calculate > 0.\n"
                        "Found rmax = %lf, nbin = %d nspheres = %d num_pN = %d\n", rmax, nbin, nc, num_pN);
        return EXIT_FAILURE;
    }
    if (np <= 0) {
        fprintf(stderr, "Error: In %s> no particles\n", __func__);
        return EXIT_FAILURE;
    }
    const REAL *P[3] = {(const REAL *)vX, (const REAL *)vY, (const REAL *)vZ};
    REAL lo[3], hi[3], wrap[3];
    for (int a = 0; a < 3; a++) { /* get_max_min_DOUBLE */
        lo[a] = H_MAXPOS;
        hi[a] = -H_MAXPOS;
        for (int64_t i = 0; i < np; i++) {
            if (P[a][i] < lo[a]) lo[a] = P[a][i];
            if (P[a][i] > hi[a]) hi[a] = P[a][i];
        }
    }
    if (options->periodic && options->boxsize == BOXSIZE_NOTGIVEN) {
        fprintf(stderr, "boxsize = %g must be specified with periodic wrap. Please specify a non-zero boxsize, or zero to detect the particle extent, or -1 to make a dimension non-periodic.\n",
                options->boxsize);
        return EXIT_FAILURE;
    }
    const double bs[3] = {options->boxsize_x, options->boxsize_y == BOXSIZE_NOTGIVEN ? options->boxsize : options->boxsize_y,
                          options->boxsize_z == BOXSIZE_NOTGIVEN ? options->boxsize : options->boxsize_z};
    int periodic[3];
    for (int a = 0; a < 3; a++) { /* :227-233 */
        periodic[a] = options->periodic && bs[a] >= 0;
        wrap[a] = (options->periodic && bs[a] > 0) ? bs[a] : (hi[a] - lo[a]);
    }

    int status = EXIT_FAILURE;
    REAL *cen[3] = {NULL, NULL, NULL};
    uint32_t *counts = NULL;
    uint64_t *pN = NULL;
    double *edges = NULL;
    for (int a = 0; a < 3; a++) cen[a] = malloc(sizeof(REAL) * (size_t)nc);
    counts = malloc(sizeof(uint32_t) * (size_t)nc * (size_t)nbin);
    pN = calloc((size_t)nbin * (size_t)num_pN, sizeof(uint64_t));
    edges = malloc(sizeof(double) * (size_t)nbin);
    if (!cen[0] || !cen[1] || !cen[2] || !counts || !pN || !edges) {
        fprintf(stderr, "Error: In %s> out of memory\n", __func__);
        goto done;
    }

    /* ---- the centres: three draws per trial; without periodic wrap a sphere that reaches past the data is redrawn
     * (:299-316).  Bounded so that an rmax larger than the box cannot spin forever. ---- */
    {
        cf_mt19937 rng;
        cf_mt_set(&rng, seed);
        int ic = 0;
        int64_t trials = 0;
        const int64_t max_trials = 1000000 + 100000 * (int64_t)nc;
        while (ic < nc) {
            if (++trials > max_trials) {
                fprintf(stderr, "Error: In %s> could not place %d spheres of radius %lf inside the particle extent\n", __func__, nc, rmax);
                goto done;
            }
            const REAL xc = wrap[0] * cf_mt_uniform(&rng) + lo[0];
            const REAL yc = wrap[1] * cf_mt_uniform(&rng) + lo[1];
            const REAL zc = wrap[2] * cf_mt_uniform(&rng) + lo[2];
            if (!options->periodic) {
                if ((xc - lo[0]) < rmax || (hi[0] - xc) < rmax || (yc - lo[1]) < rmax || (hi[1] - yc) < rmax ||
                    (zc - lo[2]) < rmax || (hi[2] - zc) < rmax)
                    continue;
            }
            cen[0][ic] = xc, cen[1][ic] = yc, cen[2][ic] = zc;
            ic++;
        }
    }

    if (cfb_upload(0, (int)sizeof(REAL), np, vX, vY, vZ, NULL, NULL, NULL)) goto done;
    {
        const REAL rmax_r = (REAL)rmax; /* the kernels take rmax as DOUBLE (vpf_kernels.c.src:20-40) */
        const REAL rstep = rmax_r / (REAL)nbin;
        const REAL rmax_sqr = rmax_r * rmax_r;
        for (int k = 0; k < nbin; k++) {
            const REAL e = (k + 1) * rstep * rstep * (k + 1);
            edges[k] = (double)e;
        }
        const double dlo[3] = {(double)lo[0], (double)lo[1], (double)lo[2]};
        const double dext[3] = {(double)(hi[0] - lo[0]), (double)(hi[1] - lo[1]), (double)(hi[2] - lo[2])};
        const double dwrap[3] = {(double)wrap[0], (double)wrap[1], (double)wrap[2]};
        const int per[3] = {options->periodic ? 1 : 0, options->periodic ? 1 : 0, options->periodic ? 1 : 0}; /* :332-380 */
        (void)periodic;
        if (cfb_count_spheres(0, (int)sizeof(REAL), dlo, dext, per, dwrap, 1, nc, cen[0], cen[1], cen[2], (double)rmax_r,
                              (double)rmax_sqr, nbin, edges, 1, counts)) goto done;
    }
    for (int64_t c = 0; c < nc; c++) { /* :401-416 */
        uint64_t cum = 0;
        for (int k = 0; k < nbin; k++) {
            cum += counts[c * nbin + k];
            if (cum < (uint64_t)num_pN) pN[(size_t)k * num_pN + cum]++;
        }
    }
    out->nbin = nbin;
    out->num_pN = num_pN;
    out->pN = calloc((size_t)nbin, sizeof(double *));
    if (!out->pN) goto done;
    {
        const REAL inv_nc = ((REAL)1.0) / (REAL)nc; /* :451-462: integer count times a REAL */
        for (int k = 0; k < nbin; k++) {
            out->pN[k] = malloc(sizeof(double) * (size_t)num_pN);
            if (!out->pN[k]) goto done;
            for (int i = 0; i < num_pN; i++) out->pN[k][i] = (double)(REAL)((int)pN[(size_t)k * num_pN + i] * inv_nc);
        }
    }
    reset_bin_refine_factors(options);
    if (options->c_api_timer) options->c_api_time = (now_ms() - t_start) * 1.0e-3;
    status = EXIT_SUCCESS;
done:
    for (int a = 0; a < 3; a++) free(cen[a]);
    free(counts); free(pN); free(edges);
    return status;
}

/* ========================================================================================== */
/* DDtheta                                                                                     */

/* find_closest_pos_DOUBLE (utils/gridlink_utils.c.src:91-113): min 1-D separation of two intervals */
static REAL HFN(cf_min_sep_1d)(const REAL a[2], const REAL b[2])
{
    if (a[0] <= b[1] && b[0] <= a[1]) return 0;
    REAL m = H_FABS(a[0] - b[0]);
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) {
            const REAL d = H_FABS(a[i] - b[j]);
            if (d < m) m = d;
        }
    return m;
}

/* check_ra_dec_DOUBLE (countpairs_theta_mocks_impl.c.src:42-84): shifts the caller's arrays in place */
static int HFN(cf_check_ra_dec)(const int64_t N, REAL *ra, REAL *dec)
{
    if (N == 0) return EXIT_SUCCESS;
    if (ra == NULL || dec == NULL) {
        fprintf(stderr, "Input arrays can not be NULL. Have RA = %p DEC = %p\n", (void *)ra, (void *)dec);
        return EXIT_FAILURE;
    }
    int fix_ra = 0, fix_dec = 0;
    for (int64_t i = 0; i < N; i++) {
        if (ra[i] < 0.0) fix_ra = 1;
        if (!(dec[i] <= 180.0)) {
            fprintf(stderr, "Declination should not be more than 180. Did you swap ra and dec?\n");
            return EXIT_FAILURE;
        }
        if (dec[i] > 90.0) fix_dec = 1;
    }
    if (fix_ra) fprintf(stderr, "DDtheta> Out of range values found for ra. Expected ra to be in the range [0.0,360.0]. Found ra values in [-180,180] -- fixing that\n");
    if (fix_dec) fprintf(stderr, "DDtheta> Out of range values found for dec. Expected dec to be in the range [-90.0,90.0]. Found dec values in [0,180] -- fixing that\n");
    if (fix_ra || fix_dec)
        for (int64_t i = 0; i < N; i++) {
            if (fix_ra) ra[i] += 180.0;
            if (fix_dec) dec[i] -= 90.0;
        }
    return EXIT_SUCCESS;
}

static int HFN(cf_theta)(const int64_t ND1, void *vra1, void *vdec1, const int64_t ND2, void *vra2, void *vdec2,
                         const int numthreads, const int autocorr, const char *binfile,
                         struct config_options *options, struct extra_options *extra, cf_box_out *out)
{
    const double t_start = now_ms();
    REAL *ra1 = vra1, *dec1 = vdec1, *ra2 = vra2, *dec2 = vdec2;
    if (options->float_type != sizeof(REAL)) return EXIT_FAILURE;
    struct extra_options dummy_extra;
    if (extra == NULL) {
        dummy_extra = get_extra_options(NONE);
        extra = &dummy_extra;
    }
    const int need_weightavg = extra->weight_method != NONE;
    options->sort_on_z = 1;
    options->autocorr = autocorr;
    if (HFN(cf_check_ra_dec)(ND1, ra1, dec1)) return EXIT_FAILURE;
    if (autocorr == 0 && HFN(cf_check_ra_dec)(ND2, ra2, dec2)) return EXIT_FAILURE;

    double *theta_upp = NULL, thetamin_d, thetamax_d;
    int nthetabin;
    /* setup_bins_DOUBLE: the float build parses the file with %f (utils/utils.c:154-196) */
    if (cf_setup_bins(binfile, &thetamin_d, &thetamax_d, &nthetabin, &theta_upp, !REAL_IS_DOUBLE)) return EXIT_FAILURE;
    const REAL thetamin = thetamin_d, thetamax = thetamax_d;
    if (!(thetamin >= 0.0 && thetamax > 0.0 && thetamin < thetamax && thetamax <= 180.0 && nthetabin >= 1)) {
        fprintf(stderr, "Error: Could not setup with theta bins correctly. (thetamin = %lf, thetamax = %lf, with nbins = %d). Expected non-zero rmin/rmax with thetamax > thetamin and nbins >=1 \n",
                (double)thetamin, (double)thetamax, nthetabin);
        free(theta_upp);
        return EXIT_FAILURE;
    }
    /* with OpenMP and !link_in_ra the reference raises the DEC refine factor to numthreads
       (countpairs_theta_mocks_impl.c.src:543-549); it only changes the lattice, never the result */
    if (options->link_in_ra == 0 && options->bin_refine_factors[1] < numthreads &&
        get_bin_refine_scheme(options) == BINNING_DFL)
        options->bin_refine_factors[1] = numthreads > 127 ? 127 : (int8_t)numthreads;
    if (options->link_in_ra && options->bin_refine_factors[0] < 1) reset_bin_refine_factors(options);
    if (options->link_in_dec && options->bin_refine_factors[1] < 1) reset_bin_refine_factors(options);
    if (options->max_cells_per_dim == 0) options->max_cells_per_dim = NLATMAX;

    double *costheta_upp = malloc(sizeof(double) * (size_t)nthetabin);
    for (int i = 0; i < nthetabin; i++) {
        const REAL t = theta_upp[i];
        const REAL c = H_COSD(t); /* countpairs_theta_mocks_impl.c.src:569-571 */
        costheta_upp[i] = (double)c;
    }

    /* RA/DEC -> unit vectors with the host libm, exactly as the reference (impl:576-608);
       CUDA's sin/cos are not bit-identical to glibc's, so this stays on the host */
    const int nsets = autocorr ? 1 : 2;
    REAL *XYZ[2][3] = {{NULL, NULL, NULL}, {NULL, NULL, NULL}};
    REAL ra_min = H_MAXPOS, dec_min = H_MAXPOS, ra_max = -H_MAXPOS, dec_max = -H_MAXPOS;
    int status = 0;
    for (int s = 0; s < nsets && !status; s++) {
        const int64_t N = s ? ND2 : ND1;
        const REAL *ra = s ? ra2 : ra1, *dec = s ? dec2 : dec1;
        for (int a = 0; a < 3; a++) {
            /* pinned, persistent across calls (no page faults, full-rate upload) */
            XYZ[s][a] = cfb_host_scratch(3 * s + a, sizeof(REAL) * (size_t)(N > 0 ? N : 1));
            if (!XYZ[s][a]) {
                fprintf(stderr, "Error: could not get host staging memory for %" PRId64 " unit vectors: %s\n", N, cfb_last_error());
                status = 1;
            }
        }
        if (status) break;
        REAL *X = XYZ[s][0], *Y = XYZ[s][1], *Z = XYZ[s][2];
#if defined(_OPENMP)
#pragma omp parallel for schedule(static)
#endif
        for (int64_t i = 0; i < N; i++) {
            X[i] = H_COSD(dec[i]) * H_COSD(ra[i]);
            Y[i] = H_COSD(dec[i]) * H_SIND(ra[i]);
            Z[i] = H_SIND(dec[i]);
        }
        /* get_max_min_ra_dec_DOUBLE (gridlink_utils.c.src:74-89); min/max are order independent */
#if defined(_OPENMP)
#pragma omp parallel for schedule(static) reduction(min : ra_min, dec_min) reduction(max : ra_max, dec_max)
#endif
        for (int64_t i = 0; i < N; i++) {
            if (ra[i] < ra_min) ra_min = ra[i];
            if (dec[i] < dec_min) dec_min = dec[i];
            if (ra[i] > ra_max) ra_max = ra[i];
            if (dec[i] > dec_max) dec_max = dec[i];
        }
    }
    int *ngrid_ra = NULL;
    int64_t *counts[2] = {NULL, NULL};
    double *rab[2] = {NULL, NULL}, *xyzb[2] = {NULL, NULL};
    int64_t *ngb_off = NULL;
    int32_t *ngb = NULL;
    uint64_t *npairs = NULL;
    double *sum_sep = NULL, *sum_w = NULL;
    cfb_stats dst;
    memset(&dst, 0, sizeof(dst));
    double t_up = 0;
    int ngrid_dec = 1;

    if (!status) {
        /* ---- lattice sizes: gridlink_mocks_theta_ra_dec_DOUBLE (gridlink_mocks_impl.c.src:1024-1148) ---- */
        const int ra_refine = options->bin_refine_factors[0], dec_refine = options->bin_refine_factors[1];
        const int max_size = options->max_cells_per_dim;
        const REAL dec_diff = dec_max - dec_min, ra_diff = ra_max - ra_min;
        if (options->link_in_dec || options->link_in_ra) {
            if (!(dec_diff > 0.0)) {
                fprintf(stderr, "All of the points can not be at the same declination. Declination difference = %lf must be non-zero\n", (double)dec_diff);
                status = 1;
            }
            if (options->link_in_ra && !(ra_diff > 0.0)) {
                fprintf(stderr, "All of the points can not be at the same RA. RA difference = %lf must be non-zero\n", (double)ra_diff);
                status = 1;
            }
        }
        if (!status) {
            if (options->link_in_dec || options->link_in_ra) {
                const REAL this_ngrid_dec = (dec_diff / thetamax < 1) ? 1 : dec_diff / thetamax;
                const int this_ngrid_dec_int = ((int)this_ngrid_dec) * dec_refine;
                ngrid_dec = this_ngrid_dec_int > max_size ? max_size : this_ngrid_dec_int;
                ngrid_dec = ngrid_dec < 1 ? 1 : ngrid_dec;
            }
            ngrid_ra = malloc(sizeof(int) * (size_t)ngrid_dec);
            const REAL dec_binsize = dec_diff / ngrid_dec;
            const REAL sin_half_thetamax = H_SIND(0.5 * thetamax);
            const REAL max_phi_cell = ra_diff;
            for (int idec = 0; idec < ngrid_dec; idec++) {
                int nmesh_ra = 1;
                if (options->link_in_ra) { /* gridlink_mocks_impl.c.src:1081-1147 */
                    REAL this_min_dec;
                    const REAL dec_lower = dec_min + idec * dec_binsize;
                    const REAL dec_upper = dec_lower + dec_binsize;
                    const REAL cos_dec_upper = H_COSD(dec_upper);
                    const REAL cos_dec_lower = H_COSD(dec_lower);
                    REAL cos_min_dec;
                    if (cos_dec_lower < cos_dec_upper) {
                        this_min_dec = dec_lower;
                        cos_min_dec = cos_dec_lower;
                    } else {
                        this_min_dec = dec_upper;
                        cos_min_dec = cos_dec_upper;
                    }
                    REAL phi_cell = max_phi_cell;
                    if ((90.0 - H_FABS(this_min_dec)) > 1.0) {
                        const REAL _tmp = sin_half_thetamax / cos_min_dec;
                        const REAL _tmp1 = _tmp < 0 ? 0 : (_tmp > 1.0 ? 1.0 : _tmp);
                        phi_cell = 2.0 * H_ASIN(_tmp1) * CF_INV_PI_OVER_180;
                        if (phi_cell <= 0) phi_cell = max_phi_cell;
                    }
                    if (!(phi_cell > 0)) {
                        fprintf(stderr, "Error: Encountered invalid binsize for RA bins for declination bin = %d\n", idec);
                        status = 1;
                        break;
                    }
                    phi_cell = phi_cell > max_phi_cell ? max_phi_cell : phi_cell;
                    const REAL this_nmesh_ra = (ra_diff / phi_cell < 1) ? 1 : ra_diff / phi_cell;
                    const int this_nmesh_ra_int = ((int)this_nmesh_ra) * ra_refine;
                    nmesh_ra = this_nmesh_ra_int > max_size ? max_size : this_nmesh_ra_int;
                    if (nmesh_ra < 1) nmesh_ra = 1;
                }
                ngrid_ra[idec] = nmesh_ra;
            }
        }
        int64_t ncells = 0;
        if (!status)
            for (int i = 0; i < ngrid_dec; i++) ncells += ngrid_ra[i];

        /* ---- upload + GPU gridlink ---- */
        cfb_theta_lattice TL;
        memset(&TL, 0, sizeof(TL));
        if (!status) {
            TL.ngrid_dec = ngrid_dec;
            TL.ngrid_ra = ngrid_ra;
            TL.dec_min = (double)dec_min;
            const REAL inv_dec_diff = (dec_diff > 0) ? (REAL)(1.0 / dec_diff) : (REAL)0;
            const REAL inv_ra_diff = (ra_diff > 0) ? (REAL)(1.0 / ra_diff) : (REAL)0;
            TL.inv_dec_diff = (double)inv_dec_diff;
            TL.ra_min = (double)ra_min;
            TL.ra_max = (double)ra_max;
            TL.inv_ra_diff = (double)inv_ra_diff;
            TL.ra_refine = ra_refine;
            TL.dec_refine = dec_refine;
            TL.sub = cfb_theta_subdivision(nsets == 2 && ND2 > ND1 ? ND2 : ND1, ncells);
            const double t0 = now_ms();
            for (int s = 0; s < nsets && !status; s++) {
                const int64_t N = s ? ND2 : ND1;
                const void *W = need_weightavg ? (s ? extra->weights1.weights[0] : extra->weights0.weights[0]) : NULL;
                if (need_weightavg && !W) {
                    fprintf(stderr, "Error: weight method needs one weight array per particle set\n");
                    status = 1;
                    break;
                }
                status = cfb_upload(s, (int)sizeof(REAL), N, XYZ[s][0], XYZ[s][1], XYZ[s][2], W, s ? ra2 : ra1,
                                    s ? dec2 : dec1);
                if (status) break;
                counts[s] = malloc(sizeof(int64_t) * (size_t)ncells);
                rab[s] = malloc(sizeof(double) * 2 * (size_t)ncells);
                xyzb[s] = malloc(sizeof(double) * 6 * (size_t)ncells);
                status = cfb_theta_gridlink(s, (int)sizeof(REAL), &TL, ncells, counts[s], rab[s], xyzb[s]);
            }
            t_up = now_ms() - t0;
        }

        /* ---- neighbour list: generate_cell_pairs_mocks_theta_ra_dec_DOUBLE (gridlink_mocks_impl.c.src:1481-1650)
                and, without RA linking, generate_cell_pairs_mocks_theta_dec_DOUBLE (:902-1003) ---- */
        if (!status) {
            const int s2 = autocorr ? 0 : 1;
            const REAL sqr_max_chord_sep = 2.0 * (1.0 - H_COSD(thetamax));
            const REAL inv_ra_diff = (ra_diff > 0) ? (REAL)(1.0 / ra_diff) : (REAL)0;
            int64_t *ra_off = malloc(sizeof(int64_t) * (size_t)ngrid_dec);
            int64_t o = 0;
            for (int i = 0; i < ngrid_dec; i++) {
                ra_off[i] = o;
                o += ngrid_ra[i];
            }
            size_t cap = (size_t)ncells * 8 + 64, nn = 0;
            ngb = malloc(sizeof(int32_t) * cap);
            ngb_off = malloc(sizeof(int64_t) * (size_t)(ncells + 1));
            for (int idec = 0; idec < ngrid_dec && !status; idec++) {
                for (int ira = 0; ira < ngrid_ra[idec]; ira++) {
                    const int64_t icell = ra_off[idec] + ira;
                    ngb_off[icell] = (int64_t)nn;
                    if (counts[0][icell] == 0) continue;
                    const size_t first_nn = nn;
                    for (int dr = -dec_refine; dr <= dec_refine; dr++) {
                        const int this_dec = idec + dr;
                        if (this_dec < 0 || this_dec >= ngrid_dec) continue;
                        int lo_ra, hi_ra;
                        if (options->link_in_ra) {
                            const REAL rb0 = rab[0][2 * icell], rb1 = rab[0][2 * icell + 1];
                            const int min_ra_this_dec = (int)(ngrid_ra[this_dec] * (rb0 - ra_min) * inv_ra_diff) - 1;
                            const int max_ra_this_dec = (int)(ngrid_ra[this_dec] * (rb1 - ra_min) * inv_ra_diff) + 1;
                            lo_ra = min_ra_this_dec - ra_refine;
                            hi_ra = max_ra_this_dec + ra_refine;
                        } else {
                            lo_ra = hi_ra = 0;
                        }
                        for (int iira = lo_ra; iira <= hi_ra; iira++) {
                            int this_ra = iira + ngrid_ra[this_dec];
                            while (this_ra < 0) this_ra += ngrid_ra[this_dec];
                            this_ra = this_ra % ngrid_ra[this_dec];
                            const int64_t icell2 = ra_off[this_dec] + this_ra;
                            if (counts[s2][icell2] == 0 || (autocorr == 1 && icell2 > icell)) continue;
                            int dup = 0; /* CHECK_AND_CONTINUE_FOR_DUPLICATE_NGB_CELLS */
                            for (size_t q = first_nn; q < nn; q++)
                                if (ngb[q] == (int32_t)icell2) {
                                    dup = 1;
                                    break;
                                }
                            if (dup) continue;
                            if (options->enable_min_sep_opt) {
                                REAL fb[3][2], sb[3][2];
                                for (int a = 0; a < 3; a++)
                                    for (int e = 0; e < 2; e++) {
                                        fb[a][e] = (REAL)xyzb[0][6 * icell + 2 * a + e];
                                        sb[a][e] = (REAL)xyzb[s2][6 * icell2 + 2 * a + e];
                                    }
                                if (options->link_in_ra) {
                                    const REAL min_dx = HFN(cf_min_sep_1d)(fb[0], sb[0]);
                                    const REAL min_dy = HFN(cf_min_sep_1d)(fb[1], sb[1]);
                                    const REAL min_dz = HFN(cf_min_sep_1d)(fb[2], sb[2]);
                                    const REAL sqr_min_sep_cells = min_dx * min_dx + min_dy * min_dy + min_dz * min_dz;
                                    if (sqr_min_sep_cells >= sqr_max_chord_sep) continue;
                                } else if (dr != 0) { /* dec-only linking prunes on z alone (:951-965) */
                                    const REAL first_z = dr < 0 ? fb[2][0] : fb[2][1];
                                    const REAL second_z = dr < 0 ? sb[2][1] : sb[2][0];
                                    const REAL min_dz = first_z - second_z;
                                    if (min_dz * min_dz >= sqr_max_chord_sep) continue;
                                }
                            }
                            if (nn + 1 > cap) {
                                cap *= 2;
                                int32_t *t = realloc(ngb, sizeof(int32_t) * cap);
                                if (!t) {
                                    status = 1;
                                    break;
                                }
                                ngb = t;
                            }
                            ngb[nn++] = (int32_t)icell2;
                        }
                    }
                }
            }
            ngb_off[ncells] = (int64_t)nn;
            free(ra_off);
        }

        /* ---- count ---- */
        if (!status) {
            cfb_binning B;
            memset(&B, 0, sizeof(B));
            B.mode = CFB_THETA;
            B.prec = (int)sizeof(REAL);
            B.autocorr = autocorr;
            B.nedges = nthetabin;
            B.edges = costheta_upp;
            B.need_avg = options->need_avg_sep ? 1 : 0;
            B.need_weights = need_weightavg;
            B.fast_acos = options->fast_acos;
            B.nslots = nthetabin;
            npairs = calloc((size_t)nthetabin, sizeof(uint64_t));
            sum_sep = calloc((size_t)nthetabin, sizeof(double));
            sum_w = calloc((size_t)nthetabin, sizeof(double));
            cfb_hist H = {npairs, sum_sep, sum_w};
            {
                cf_sig_t prev_handlers[3];
                cf_signals_install(prev_handlers);
                status = cfb_count_theta(&B, ncells, ngb_off, ngb, &H, &dst);
                if (cf_signals_restore(prev_handlers)) status = 1; /* interrupted: EXIT_FAILURE */
            }
            status = reduce_across_ranks(status, npairs, sum_sep, sum_w, nthetabin);
        }
    }
    for (int s = 0; s < 2; s++) {
        for (int a = 0; a < 3; a++) XYZ[s][a] = NULL; /* context-owned staging buffers */
        free(counts[s]); free(rab[s]); free(xyzb[s]);
    }
    free(ngrid_ra); free(ngb); free(ngb_off); free(costheta_upp);
    if (status) {
        free(npairs); free(sum_sep); free(sum_w); free(theta_upp);
        return EXIT_FAILURE;
    }

    /* ---- epilogue (countpairs_theta_mocks_impl.c.src:1125-1203) ---- */
    if (autocorr == 1) {
        for (int i = 0; i < nthetabin; i++) {
            npairs[i] *= 2;
            sum_sep[i] *= 2.0;
            sum_w[i] *= 2.0;
        }
        if (theta_upp[0] <= 0.0) {
            npairs[1] += (uint64_t)ND1;
            if (need_weightavg) {
                const REAL *w = (const REAL *)extra->weights0.weights[0];
                for (int64_t j = 0; j < ND1; j++) sum_w[1] += (double)(REAL)(w[j] * w[j]);
            }
        }
    }
    for (int i = 1; i < nthetabin; i++)
        if (npairs[i] > 0) {
            sum_sep[i] /= (double)npairs[i];
            sum_w[i] /= (double)npairs[i];
        }
    for (int i = 0; i < nthetabin; i++) {
        if (!options->need_avg_sep) sum_sep[i] = 0.0;
        if (!need_weightavg) sum_w[i] = 0.0;
    }
    out->nbin = nthetabin;
    out->n2 = 0;
    out->npairs = npairs;
    out->avg = sum_sep;
    out->wavg = sum_w;
    out->rupp = malloc(sizeof(double) * (size_t)nthetabin);
    for (int i = 0; i < nthetabin; i++) out->rupp[i] = theta_upp[i];
    out->cf = NULL;
    free(theta_upp);

    g_stats.dev = dst;
    g_stats.ms_upload = t_up;
    g_stats.nmesh[0] = ngrid_dec;
    g_stats.nmesh[1] = g_stats.nmesh[2] = 0;
    for (int a = 0; a < 3; a++) g_stats.refine[a] = options->bin_refine_factors[a];
    reset_bin_refine_factors(options);
    const double t_end = now_ms();
    g_stats.ms_host_total = t_end - t_start;
    if (options->c_api_timer) options->c_api_time = (t_end - t_start) * 1.0e-3;
    return EXIT_SUCCESS;
}

#undef H_COSD
#undef H_SIND
#undef H_ASIN
#undef H_FABS
#undef H_SQRT
#undef H_MAXPOS
#undef HCAT_
#undef HCAT
#undef HFN
