/*
 * r_shim.c -- the thin .Call layer between R and librfsi_b200.so (include/rfsi_b200.h).
 *
 * This is the binding a maintainer of the reference would add so that the scripts'
 *   near.obs(...)            simulation/simulation.R:191-196, temp_croatia/2_predictions.R:160-165
 *   predict(<ranger>, df)    simulation/simulation.R:214, temp_croatia/2_predictions.R:167,174-176
 * land in the CUDA engine instead of meteo/nabor and ranger.  R is not installed in the build
 * image, so this file is compile-checked against src/stub/Rinternals.h (tests/test_r_shim.py)
 * and exercised for real only on a machine with R:  R CMD SHLIB r_shim.c -lrfsi_b200.
 *
 * Rules kept here: R allocates and PROTECTs every output; native code only
 * fills REAL()/INTEGER() pointers; inputs are read-only, column-major; NA_real_/NA_integer_
 * pass through; on error every native temporary is released BEFORE Rf_error (no longjmp
 * across live device allocations: the engine frees its own before returning a status);
 * called from the main R thread only; the CUDA context is created lazily inside the first
 * engine call of each process, so forked foreach/%dopar% workers
 * (prcp_catalonia/5_prcp_prediction.R:243-244) each get their own as long as the parent has
 * not touched the engine before forking.
 */
#include <stdio.h>

#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#include "rfsi_b200.h"

static void check_rc(int rc) {
    if (rc < 0) Rf_error("rfsi_b200: %s", rfsi_last_error());
    if (rc == RFSI_W_TOO_FEW_OBS) Rf_warning("rfsi_b200: %s", rfsi_last_error());
}

/* near.obs: loc_xy (npix x 2 double), obs_xy (nobs x 2), obs_z (nobs), n_obs, rm_dupl
 * -> list(dist = npix x n, obs = npix x n, idx = npix x n integer) */
SEXP C_rfsi_near_obs(SEXP loc_xy, SEXP obs_xy, SEXP obs_z, SEXP n_obs, SEXP rm_dupl, SEXP device)
{
    const R_xlen_t npix = Rf_nrows(loc_xy), nobs = Rf_nrows(obs_xy);
    const int n = Rf_asInteger(n_obs);
    if (!Rf_isReal(loc_xy) || !Rf_isReal(obs_xy) || !Rf_isReal(obs_z)) Rf_error("near.obs: numeric inputs expected");
    if (Rf_ncols(loc_xy) != 2 || Rf_ncols(obs_xy) != 2 || XLENGTH(obs_z) != nobs) Rf_error("near.obs: bad shapes");
    SEXP dist = PROTECT(Rf_allocMatrix(REALSXP, (int)npix, n));
    SEXP obs = PROTECT(Rf_allocMatrix(REALSXP, (int)npix, n));
    SEXP idx = PROTECT(Rf_allocMatrix(INTSXP, (int)npix, n));
    const double* l = REAL(loc_xy);
    const double* o = REAL(obs_xy);
    const int rc = rfsi_near_obs_f64(l, l + npix, (int64_t)npix, o, o + nobs, REAL(obs_z), (int32_t)nobs, n,
                                     Rf_asLogical(rm_dupl) ? 1 : 0, REAL(dist), REAL(obs), INTEGER(idx),
                                     Rf_asInteger(device));
    check_rc(rc);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, dist);
    SET_VECTOR_ELT(out, 1, obs);
    SET_VECTOR_ELT(out, 2, idx);
    SEXP names = PROTECT(Rf_allocVector(STRSXP, 3));
    SET_STRING_ELT(names, 0, Rf_mkChar("dist"));
    SET_STRING_ELT(names, 1, Rf_mkChar("obs"));
    SET_STRING_ELT(names, 2, Rf_mkChar("idx"));
    Rf_setAttrib(out, R_NamesSymbol, names);
    UNPROTECT(5);
    return out;
}

/* near.obs as the data.frame meteo::near.obs returns: a list of 2 * n.obs numeric vectors named dist1..distN,
 * obs1..obsN with class "data.frame" and compact row names, every column filled in place by the engine
 * (rfsi_near_obs_cols_f64) -- no npix x n.obs matrix and no as.data.frame() copy on the R side. */
SEXP C_rfsi_near_obs_df(SEXP loc_xy, SEXP obs_xy, SEXP obs_z, SEXP n_obs, SEXP rm_dupl, SEXP device)
{
    const R_xlen_t npix = Rf_nrows(loc_xy), nobs = Rf_nrows(obs_xy);
    const int n = Rf_asInteger(n_obs);
    if (!Rf_isReal(loc_xy) || !Rf_isReal(obs_xy) || !Rf_isReal(obs_z)) Rf_error("near.obs: numeric inputs expected");
    if (Rf_ncols(loc_xy) != 2 || Rf_ncols(obs_xy) != 2 || XLENGTH(obs_z) != nobs || n < 1) Rf_error("near.obs: bad shapes");
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2 * (R_xlen_t)n));
    SEXP names = PROTECT(Rf_allocVector(STRSXP, 2 * (R_xlen_t)n));
    double** cols = (double**)R_alloc((size_t)2 * n, sizeof(double*));
    char buf[32];
    for (int j = 0; j < 2 * n; ++j) {
        SEXP col = Rf_allocVector(REALSXP, npix);
        SET_VECTOR_ELT(out, j, col);                  /* reachable from `out` at once: no PROTECT of its own needed */
        cols[j] = REAL(col);
        snprintf(buf, sizeof buf, "%s%d", j < n ? "dist" : "obs", (j < n ? j : j - n) + 1);
        SET_STRING_ELT(names, j, Rf_mkChar(buf));
    }
    const double* l = REAL(loc_xy);
    const double* o = REAL(obs_xy);
    const int rc = rfsi_near_obs_cols_f64(l, l + npix, (int64_t)npix, o, o + nobs, REAL(obs_z), (int32_t)nobs, n,
                                          Rf_asLogical(rm_dupl) ? 1 : 0, cols, cols + n, NULL, Rf_asInteger(device));
    check_rc(rc);
    Rf_setAttrib(out, R_NamesSymbol, names);
    SEXP rn = PROTECT(Rf_allocVector(INTSXP, 2));     /* compact row names c(NA_integer_, -npix) */
    INTEGER(rn)[0] = NA_INTEGER; INTEGER(rn)[1] = -(int)npix;
    Rf_setAttrib(out, R_RowNamesSymbol, rn);
    SEXP cls = PROTECT(Rf_mkString("data.frame"));
    Rf_setAttrib(out, R_ClassSymbol, cls);
    UNPROTECT(4);
    return out;
}

/* near.obs for every day at once (training tables, temp_croatia/1_models.R:1034-1047):
 * loc_xy NULL -> locations are the stations; obs_z nobs x ndays (column d = day d, NA = silent)
 * -> list(dist, obs) each a numeric array npix x n.obs x ndays */
SEXP C_rfsi_near_obs_days(SEXP loc_xy, SEXP obs_xy, SEXP obs_z, SEXP n_obs, SEXP device)
{
    const R_xlen_t nobs = Rf_nrows(obs_xy);
    const R_xlen_t npix = (loc_xy == R_NilValue) ? nobs : Rf_nrows(loc_xy);
    const int ndays = Rf_ncols(obs_z), n = Rf_asInteger(n_obs);
    SEXP dist = PROTECT(Rf_allocVector(REALSXP, npix * n * ndays));   /* [day][j][pixel] == npix x n x ndays */
    SEXP obs = PROTECT(Rf_allocVector(REALSXP, npix * n * ndays));
    const double* l = (loc_xy == R_NilValue) ? NULL : REAL(loc_xy);
    const double* o = REAL(obs_xy);
    check_rc(rfsi_near_obs_days_f64(l, l ? l + npix : NULL, (int64_t)npix, o, o + nobs, REAL(obs_z), (int32_t)nobs,
                                    ndays, n, 1, REAL(dist), REAL(obs), NULL, Rf_asInteger(device)));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, dist);
    SET_VECTOR_ELT(out, 1, obs);
    UNPROTECT(3);
    return out;
}

static void forest_finalizer(SEXP ptr) {
    rfsi_forest_t* f = (rfsi_forest_t*)R_ExternalPtrAddr(ptr);
    if (f) {
        rfsi_forest_destroy(f);
        R_ClearExternalPtr(ptr);
    }
}

/* rfsi_import_ranger(model) flattens the ranger object in R (R/import.R) and calls this:
 * node_offsets (double, T+1), left/right/split_var (integer), split_val (double),
 * rnv (double matrix max_nodes x T or NULL), nfeat -> external pointer */
SEXP C_rfsi_forest_new(SEXP node_offsets, SEXP left, SEXP right, SEXP split_var, SEXP split_val, SEXP rnv,
                       SEXP nfeat)
{
    const int T = (int)XLENGTH(node_offsets) - 1;
    int64_t* off = (int64_t*)R_alloc((size_t)T + 1, sizeof(int64_t));
    for (int t = 0; t <= T; ++t) off[t] = (int64_t)REAL(node_offsets)[t];
    int64_t* roff = NULL;
    const double* rv = NULL;
    if (rnv != R_NilValue) {
        const int64_t nr = Rf_nrows(rnv);
        if (Rf_ncols(rnv) != T) Rf_error("random.node.values must have num.trees columns");
        roff = (int64_t*)R_alloc((size_t)T, sizeof(int64_t));
        for (int t = 0; t < T; ++t) roff[t] = (int64_t)t * nr;   /* R matrices are column-major */
        rv = REAL(rnv);
    }
    rfsi_forest_t* f = NULL;
    const int rc = rfsi_forest_create(T, off, INTEGER(left), INTEGER(right), INTEGER(split_var), REAL(split_val),
                                      roff, rv, Rf_asInteger(nfeat), &f);
    check_rc(rc);
    SEXP ptr = PROTECT(R_MakeExternalPtr(f, Rf_install("rfsi_forest"), R_NilValue));
    R_RegisterCFinalizerEx(ptr, forest_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

static rfsi_forest_t* forest_of(SEXP h) {
    rfsi_forest_t* f = (rfsi_forest_t*)R_ExternalPtrAddr(h);
    if (!f) Rf_error("rfsi_b200: forest handle is NULL (saved/restored workspace? re-import the model)");
    return f;
}

/* predict: X npix x nfeat double matrix already in independent.variable.names order;
 * probs NULL -> numeric vector of means, else npix x length(probs) matrix of quantiles */
SEXP C_rfsi_predict(SEXP h, SEXP X, SEXP probs, SEXP device)
{
    rfsi_forest_t* f = forest_of(h);
    const R_xlen_t npix = Rf_nrows(X);
    if (!Rf_isReal(X) || Rf_ncols(X) != (int)rfsi_forest_info(f, 1)) Rf_error("predict: X must be npix x nfeat double");
    if (probs == R_NilValue) {
        SEXP out = PROTECT(Rf_allocVector(REALSXP, npix));
        check_rc(rfsi_forest_predict(f, REAL(X), (int64_t)npix, (int64_t)npix, REAL(out), Rf_asInteger(device)));
        UNPROTECT(1);
        return out;
    }
    const int q = (int)XLENGTH(probs);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)npix, q));
    check_rc(rfsi_forest_quantiles(f, REAL(X), (int64_t)npix, (int64_t)npix, REAL(probs), q, REAL(out),
                                   Rf_asInteger(device)));
    UNPROTECT(1);
    return out;
}

/* predict(type = "terminalNodes"): ranger's node ids, npix x num.trees integer matrix */
SEXP C_rfsi_terminal_nodes(SEXP h, SEXP X, SEXP device)
{
    rfsi_forest_t* f = forest_of(h);
    const R_xlen_t npix = Rf_nrows(X);
    if (!Rf_isReal(X) || Rf_ncols(X) != (int)rfsi_forest_info(f, 1)) Rf_error("predict: X must be npix x nfeat double");
    SEXP out = PROTECT(Rf_allocMatrix(INTSXP, (int)npix, (int)rfsi_forest_info(f, 0)));
    check_rc(rfsi_forest_terminal_nodes(f, REAL(X), (int64_t)npix, (int64_t)npix, INTEGER(out), Rf_asInteger(device)));
    UNPROTECT(1);
    return out;
}

/* fused near.obs -> predict for explicit locations over ndays:
 * loc_xy npix x 2; obs_xy nobs x 2; obs_z nobs x ndays (column d = day d, NA = missing);
 * cov: list of numeric vectors (npix, static) or npix x ndays matrices; feat_map integer codes;
 * probs NULL or numeric -> list(mean = npix x ndays, q = npix x ndays x Q) */
SEXP C_rfsi_pred_fused(SEXP h, SEXP loc_xy, SEXP obs_xy, SEXP obs_z, SEXP n_obs, SEXP cov, SEXP feat_map,
                       SEXP probs, SEXP device)
{
    rfsi_forest_t* f = forest_of(h);
    const R_xlen_t npix = Rf_nrows(loc_xy), nobs = Rf_nrows(obs_xy);
    const int ndays = Rf_ncols(obs_z), ncov = (int)XLENGTH(cov);
    rfsi_fused_args_t a;
    memset(&a, 0, sizeof a);
    a.struct_size = sizeof a;
    a.loc_x = REAL(loc_xy); a.loc_y = REAL(loc_xy) + npix; a.npix = (int64_t)npix; a.grid_ncol = 1;
    a.obs_x = REAL(obs_xy); a.obs_y = REAL(obs_xy) + nobs; a.nobs = (int32_t)nobs;
    a.obs_z = REAL(obs_z);            /* nobs x ndays column-major == [ndays][nobs] row-major */
    a.ndays = ndays; a.n_obs = Rf_asInteger(n_obs); a.rm_dupl = 1;
    const double** cols = (const double**)R_alloc((size_t)(ncov ? ncov : 1), sizeof(double*));
    int64_t* strides = (int64_t*)R_alloc((size_t)(ncov ? ncov : 1), sizeof(int64_t));
    for (int c = 0; c < ncov; ++c) {
        SEXP v = VECTOR_ELT(cov, c);
        cols[c] = REAL(v);
        strides[c] = (XLENGTH(v) == npix) ? 0 : (int64_t)npix;
    }
    a.ncov = ncov; a.cov = cols; a.cov_day_stride = strides; a.feat_map = INTEGER(feat_map);
    SEXP mean = PROTECT(Rf_allocMatrix(REALSXP, (int)npix, ndays));
    a.out_mean = REAL(mean);
    SEXP q = R_NilValue;
    int nprot = 1;
    if (probs != R_NilValue) {
        a.probs = REAL(probs); a.nprobs = (int32_t)XLENGTH(probs);
        q = PROTECT(Rf_allocVector(REALSXP, npix * ndays * a.nprobs)); ++nprot;
        a.out_q = REAL(q);
    }
    check_rc(rfsi_predict_fused(f, &a, Rf_asInteger(device)));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2)); ++nprot;
    SET_VECTOR_ELT(out, 0, mean);
    SET_VECTOR_ELT(out, 1, q);
    UNPROTECT(nprot);
    return out;
}

/* the map-making loop of prcp_catalonia/5_prcp_prediction.R:244-309 in one call:
 * cov is a list whose elements are numeric (as in C_rfsi_pred_fused) or raster descriptors
 *   list(values = getValues(r) (double or integer, cells row by row from the top-left, layers
 *        one after the other), dim = c(ncol, nrow, nlayers), geo = c(xmin, ymax, xres, yres),
 *        divisor = 10 or NULL)
 * sampled on the device at cov_loc (npix x 2, the grid in the rasters' CRS: `gg`) while
 * near.obs runs at loc_xy (`gg2`); fix_idx (k x 2, 1-based covariate indices target / other)
 * and fix_delta (k) are the repairs `ifelse(target > other, other - delta, target)`.
 * -> list(pred = round(mean*100), iqr = round((q_last - q_first)/2*100)) as integer npix x ndays
 * matrices, NA where the scripts write NAflag. */
SEXP C_rfsi_pred_maps(SEXP h, SEXP loc_xy, SEXP cov_loc, SEXP obs_xy, SEXP obs_z, SEXP n_obs, SEXP cov, SEXP feat_map,
                      SEXP probs, SEXP fix_idx, SEXP fix_delta, SEXP device)
{
    rfsi_forest_t* f = forest_of(h);
    const R_xlen_t npix = Rf_nrows(loc_xy), nobs = Rf_nrows(obs_xy);
    const int ndays = Rf_ncols(obs_z), ncov = (int)XLENGTH(cov);
    rfsi_fused_args_t a;
    memset(&a, 0, sizeof a);
    a.struct_size = sizeof a;
    a.loc_x = REAL(loc_xy); a.loc_y = REAL(loc_xy) + npix; a.npix = (int64_t)npix; a.grid_ncol = 1;
    if (cov_loc != R_NilValue) {
        if (Rf_nrows(cov_loc) != npix) Rf_error("pred.maps: cov.locations must have one row per location");
        a.cov_loc_x = REAL(cov_loc); a.cov_loc_y = REAL(cov_loc) + npix;
    }
    a.obs_x = REAL(obs_xy); a.obs_y = REAL(obs_xy) + nobs; a.nobs = (int32_t)nobs;
    a.obs_z = REAL(obs_z); a.ndays = ndays; a.n_obs = Rf_asInteger(n_obs); a.rm_dupl = 1;
    const size_t nc = (size_t)(ncov ? ncov : 1);
    const double** cols = (const double**)R_alloc(nc, sizeof(double*));
    int64_t* strides = (int64_t*)R_alloc(nc, sizeof(int64_t));
    rfsi_raster_t* rast = (rfsi_raster_t*)R_alloc(nc, sizeof(rfsi_raster_t));
    memset(rast, 0, nc * sizeof(rfsi_raster_t));
    for (int c = 0; c < ncov; ++c) {
        SEXP v = VECTOR_ELT(cov, c);
        cols[c] = NULL; strides[c] = 0;
        if (TYPEOF(v) == VECSXP) {          /* raster descriptor */
            SEXP val = VECTOR_ELT(v, 0), dim = VECTOR_ELT(v, 1), geo = VECTOR_ELT(v, 2), dv = VECTOR_ELT(v, 3);
            rfsi_raster_t* r = &rast[c];
            if (TYPEOF(val) == INTSXP) { r->data = INTEGER(val); r->dtype = RFSI_I32; r->nodata = (double)NA_INTEGER; }
            else { r->data = REAL(val); r->dtype = RFSI_F64; r->nodata = R_NaN; }   /* NA_real_ is a NaN: masked anyway */
            r->ncol = INTEGER(dim)[0]; r->nrow = INTEGER(dim)[1]; r->nlayers = INTEGER(dim)[2];
            r->x_ul = REAL(geo)[0]; r->y_ul = REAL(geo)[1]; r->dx = REAL(geo)[2]; r->dy = REAL(geo)[3];
            r->scale = 1.0; r->offset = 0.0;
            if (dv != R_NilValue) { r->scale = REAL(dv)[0]; r->dtype |= RFSI_RASTER_DIVIDE; }
        } else {
            cols[c] = REAL(v);
            strides[c] = (XLENGTH(v) == npix) ? 0 : (int64_t)npix;
        }
    }
    a.ncov = ncov; a.cov = cols; a.cov_day_stride = strides; a.cov_raster = rast; a.feat_map = INTEGER(feat_map);
    const int nfix = fix_idx == R_NilValue ? 0 : Rf_nrows(fix_idx);
    rfsi_cov_fix_t* fix = (rfsi_cov_fix_t*)R_alloc((size_t)(nfix ? nfix : 1), sizeof(rfsi_cov_fix_t));
    for (int k = 0; k < nfix; ++k) {
        fix[k].target = INTEGER(fix_idx)[k] - 1; fix[k].other = INTEGER(fix_idx)[k + nfix] - 1;
        fix[k].delta = REAL(fix_delta)[k];
    }
    a.cov_fix = nfix ? fix : NULL; a.ncov_fix = nfix;
    const size_t cells = (size_t)npix * (size_t)ndays;
    int16_t* m16 = (int16_t*)R_alloc(cells ? cells : 1, sizeof(int16_t));
    int16_t* q16 = NULL;
    a.out_mean_i16 = m16;
    double* mean = (double*)R_alloc(cells ? cells : 1, sizeof(double));   /* the Int16 epilogue rides on the mean */
    a.out_mean = mean;
    if (probs != R_NilValue) {
        a.probs = REAL(probs); a.nprobs = (int32_t)XLENGTH(probs);
        q16 = (int16_t*)R_alloc(cells ? cells : 1, sizeof(int16_t));
        a.out_iqr_i16 = q16;
    }
    /* device: one GPU, or an integer vector of GPUs (one band of pixels each) */
    if (XLENGTH(device) > 1) check_rc(rfsi_predict_fused_multi(f, &a, INTEGER(device), (int)XLENGTH(device)));
    else check_rc(rfsi_predict_fused(f, &a, Rf_asInteger(device)));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SEXP pm = PROTECT(Rf_allocMatrix(INTSXP, (int)npix, ndays));
    for (size_t i = 0; i < cells; ++i) INTEGER(pm)[i] = m16[i] == -32767 ? NA_INTEGER : (int)m16[i];
    SET_VECTOR_ELT(out, 0, pm);
    if (q16) {
        SEXP qm = PROTECT(Rf_allocMatrix(INTSXP, (int)npix, ndays));
        for (size_t i = 0; i < cells; ++i) INTEGER(qm)[i] = q16[i] == -32767 ? NA_INTEGER : (int)q16[i];
        SET_VECTOR_ELT(out, 1, qm);
        UNPROTECT(1);
    }
    UNPROTECT(2);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"C_rfsi_near_obs", (DL_FUNC)&C_rfsi_near_obs, 6},
    {"C_rfsi_near_obs_df", (DL_FUNC)&C_rfsi_near_obs_df, 6},
    {"C_rfsi_near_obs_days", (DL_FUNC)&C_rfsi_near_obs_days, 5},
    {"C_rfsi_forest_new", (DL_FUNC)&C_rfsi_forest_new, 7},
    {"C_rfsi_predict", (DL_FUNC)&C_rfsi_predict, 4},
    {"C_rfsi_terminal_nodes", (DL_FUNC)&C_rfsi_terminal_nodes, 3},
    {"C_rfsi_pred_fused", (DL_FUNC)&C_rfsi_pred_fused, 9},
    {"C_rfsi_pred_maps", (DL_FUNC)&C_rfsi_pred_maps, 12},
    {NULL, NULL, 0}};

void R_init_rfsib200(DllInfo* dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
