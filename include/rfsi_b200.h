/*
 * rfsi_b200.h -- C ABI of the B200-native RFSI prediction engine.
 *
 * The reference (AleksandarSekulic/RFSI) has no FFI of its own: its boundary for
 * this path is the R call shapes near.obs(...) and predict(<ranger>, ...) (and the
 * rfsi()/pred.rfsi() wrappers of README.md:93-143), which land in the CRAN
 * packages meteo/nabor and ranger.  Each entry point below names the reference
 * call it replaces; INTEGRATION.md shows the R-side .Call stub that binds it.
 *
 * Conventions
 *  - plain C, caller-owned buffers, no ownership transfer except the forest handle;
 *  - matrices are COLUMN-MAJOR (R's layout): column j of an npix x n matrix starts
 *    at base + j*ld;
 *  - missing values are R's: NA_real_ (NaN, payload 1954) and NA_integer_ (INT_MIN);
 *  - every function returns 0 (RFSI_OK), a positive warning code, or a negative
 *    RFSI_E_* error; rfsi_last_error() gives the thread-local message;
 *  - there is NO CPU fallback: without an sm_100 device calls fail with
 *    RFSI_E_NO_DEVICE;
 *  - fork(): CUDA is initialised lazily, by the first call that needs a device.  Forked workers
 *    (registerDoParallel + foreach %dopar%: simulation/simulation.R:113-114,
 *    prcp_catalonia/5_prcp_prediction.R:243-244) each get their own context as long as the PARENT
 *    has not made such a call before forking; rfsi_forest_create() never touches a device, so a
 *    model imported in the parent is usable in every child.  A child of a parent that did use a
 *    device gets RFSI_E_FORKED from every device call (a CUDA context does not survive fork);
 *  - rfsi_*      : HOST pointers in, HOST pointers out (copies are inside the call);
 *    rfsi_dev_*  : DEVICE pointers on `device`, enqueued on `stream` (a cudaStream_t
 *                  passed as void*; NULL = the legacy default stream), asynchronous.
 */
#ifndef RFSI_B200_H
#define RFSI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RFSI_B200_VERSION 100 /* 0.1.0 */

enum {
    RFSI_OK = 0,
    RFSI_W_TOO_FEW_OBS = 1, /* some outputs were NA-filled: fewer stations than n.obs(+1) */
    RFSI_E_CUDA = -1,
    RFSI_E_NO_DEVICE = -2,
    RFSI_E_ARG = -3,
    RFSI_E_NA_INPUT = -4, /* NA/NaN coordinate, or NA in a feature column the forest uses */
    RFSI_E_OOM = -5,
    RFSI_E_FOREST = -6, /* malformed forest (child out of range, cycle, bad varID) */
    RFSI_E_FORKED = -7  /* this process was fork()ed from one that had already used the engine on a GPU */
};

/* feature-map codes for rfsi_predict_fused: which model column is what */
#define RFSI_FEAT_COV(c) ((int32_t)(c))             /* external covariate column c      */
#define RFSI_FEAT_DIST(j) ((int32_t)(0x10000 + (j))) /* dist{j+1}, j = 0..n_obs-1        */
#define RFSI_FEAT_OBS(j) ((int32_t)(0x20000 + (j)))  /* obs{j+1},  j = 0..n_obs-1        */

typedef struct rfsi_forest rfsi_forest_t;

int rfsi_version(void);
const char* rfsi_last_error(void);
/* number of usable sm_100 devices (0 = none; negative = RFSI_E_*) */
int rfsi_device_count(void);

/* Pinned host memory for callers that want full-rate PCIe copies (optional). */
int rfsi_host_alloc(void** out, size_t bytes);
void rfsi_host_free(void* p);

/* ---- stage 1 ------------------------------------------------------------------
 * near.obs(locations, observations, zcol, n.obs): simulation/simulation.R:191-196,
 * simulation/sim_500.R:134-139, temp_croatia/2_predictions.R:160-165,
 * prcp_catalonia/5_prcp_prediction.R:278-283, temp_croatia/1_models.R:1040-1045.
 * k = n_obs+1 exact nearest observation points under the total order (d2, index),
 * zero-distance self match dropped when rm_dupl; out_dist/out_obs are npix x n_obs
 * column-major (ld = npix) = the dist1..distN / obs1..obsN columns; out_idx
 * (nullable) 1-based observation indices.  obs_z may hold NA (propagates to obs). */
int rfsi_near_obs_f64(const double* loc_x, const double* loc_y, int64_t npix,
                      const double* obs_x, const double* obs_y, const double* obs_z,
                      int32_t nobs, int32_t n_obs, int32_t rm_dupl,
                      double* out_dist, double* out_obs, int32_t* out_idx, int device);

/* The same call with one pointer PER COLUMN (dist_cols[j], obs_cols[j], idx_cols[j] (nullable array), each npix
 * long): what an R binding needs to return the data.frame near.obs() returns -- a list of 2 * n.obs numeric
 * vectors -- without going through an npix x n.obs matrix and as.data.frame() (a 100 MB copy at simulation.R's
 * 250 000 x 50 columns). */
int rfsi_near_obs_cols_f64(const double* loc_x, const double* loc_y, int64_t npix,
                           const double* obs_x, const double* obs_y, const double* obs_z,
                           int32_t nobs, int32_t n_obs, int32_t rm_dupl,
                           double* const* dist_cols, double* const* obs_cols, int32_t* const* idx_cols, int device);

/* fp32 mode (north_star: "1e-5 in fp32 mode"): float buffers in both directions through the same pinned, chunked
 * pipeline as the fp64 call (half the PCIe bytes and host copies: the call is 1.3-1.5x faster at simulation.R's
 * sizes, profiles/r02_i_fp32_mode.txt).  On the device the coordinates are widened exactly; with 128 stations or
 * more the cell-grid search forms d2 in fp32 first and drops on that value every station that provably cannot enter
 * or tie with the current list (|d2f - d2| < 2^-22 d2, bound applied with a 2^-20 margin); the survivors are
 * re-ranked by their exact fp64 d2.  Neighbour indices and their order are therefore bit-identical to the fp64 call
 * on the same float-representable coordinates; dist/obs are the fp64 results rounded to float (6e-8 relative).
 * Same status codes as rfsi_near_obs_f64 (RFSI_E_NA_INPUT for NA coordinates, RFSI_W_TOO_FEW_OBS + NA fill). */
int rfsi_near_obs_f32(const float* loc_x, const float* loc_y, int64_t npix,
                      const float* obs_x, const float* obs_y, const float* obs_z,
                      int32_t nobs, int32_t n_obs, int32_t rm_dupl,
                      float* out_dist, float* out_obs, int32_t* out_idx, int device);

int rfsi_dev_near_obs_f64(const double* loc_x, const double* loc_y, int64_t npix,
                          const double* obs_x, const double* obs_y, const double* obs_z,
                          int32_t nobs, int32_t n_obs, int32_t rm_dupl,
                          double* out_dist, double* out_obs, int32_t* out_idx,
                          int32_t* warn_flag /* device int, nullable, OR-ed into: bit 0 = rows with fewer
                                                neighbours than n.obs (NA-filled), bit 1 = locations with an NA
                                                coordinate (their rows are NA): the caller zeroes it */,
                          int device, void* stream);

/* near.obs for EVERY day of a station table in one call -- the per-day feature table, above
 * all the training tables built with
 *   foreach(t=time) near.obs(locations=day_df, observations=day_df, zcol, n.obs)
 * temp_croatia/1_models.R:1034-1047, prcp_catalonia/4_prcp_case_study_catalonia_RK.R:1699-1712.
 * obs_z is ndays x nobs (row d = day d, NA = no observation).  Per day only that day's
 * observed stations are searched (the scripts' !is.na filter).  loc_x == NULL: locations are
 * the stations themselves (npix = nobs; a row whose own station is silent that day is NA).
 * out_dist / out_obs / out_idx (nullable, 1-based ORIGINAL station index): [ndays][n_obs][npix]. */
int rfsi_near_obs_days_f64(const double* loc_x, const double* loc_y, int64_t npix,
                           const double* obs_x, const double* obs_y, const double* obs_z,
                           int32_t nobs, int32_t ndays, int32_t n_obs, int32_t rm_dupl,
                           double* out_dist, double* out_obs, int32_t* out_idx, int device);

/* ---- forest import --------------------------------------------------------------
 * The ranger model object saved at temp_croatia/1_models.R:1056-1063 and loaded at
 * temp_croatia/2_predictions.R:59 / prcp_catalonia/5_prcp_prediction.R:233, flattened
 * (rfsi_import_ranger): trees concatenated, tree t owning nodes
 * [node_offsets[t], node_offsets[t+1]); left/right = child.nodeIDs[[t]][[1]]/[[2]]
 * (tree-local, 0-based, both 0 at a terminal node); split_var = split.varIDs[[t]]
 * (0-based column of X, already re-based for old ranger objects); split_val =
 * split.values[[t]] (threshold, or the prediction at a terminal node);
 * random_node_values (nullable; needed for quantiles) with tree t's column starting
 * at rnv_offsets[t] and indexed by tree-local node id. */
int rfsi_forest_create(int32_t ntrees, const int64_t* node_offsets,
                       const int32_t* left, const int32_t* right,
                       const int32_t* split_var, const double* split_val,
                       const int64_t* rnv_offsets, const double* random_node_values,
                       int32_t nfeat, rfsi_forest_t** out);
void rfsi_forest_destroy(rfsi_forest_t* f);
/* push the device image of the forest to `device` now (otherwise done on first use) */
int rfsi_forest_upload(const rfsi_forest_t* f, int device);
/* info: 0 ntrees, 1 nfeat, 2 total nodes (device layout), 3 max depth, 4 has quantile
 * samples, 5 device bytes, 6 total internal nodes */
int64_t rfsi_forest_info(const rfsi_forest_t* f, int what);

/* ---- stage 2 ------------------------------------------------------------------
 * predict(<ranger>, df)$predictions: simulation/simulation.R:214,
 * temp_croatia/2_predictions.R:167, prcp_catalonia/5_prcp_prediction.R:287.
 * X is npix x nfeat column-major with leading dimension ldx, columns in the model's
 * independent.variable.names order.  NA/NaN in X -> RFSI_E_NA_INPUT (ranger errors). */
int rfsi_forest_predict(const rfsi_forest_t* f, const double* X, int64_t npix, int64_t ldx,
                        double* out, int device);

/* predict(<ranger quantreg=TRUE>, df, type="quantiles", quantiles=probs):
 * temp_croatia/2_predictions.R:174-176, prcp_catalonia/5_prcp_prediction.R:289-291.
 * out is npix x nprobs column-major.  Forests of up to 1024 trees (the reference grows 250);
 * more -> RFSI_E_ARG before any work is launched. */
int rfsi_forest_quantiles(const rfsi_forest_t* f, const double* X, int64_t npix, int64_t ldx,
                          const double* probs, int32_t nprobs, double* out, int device);

/* predict(type="terminalNodes"): npix x ntrees column-major tree-local node ids */
int rfsi_forest_terminal_nodes(const rfsi_forest_t* f, const double* X, int64_t npix,
                               int64_t ldx, int32_t* out, int device);

/* device-pointer forms; any of out_mean / out_q / out_terminal may be NULL.
 * scratch: device buffer of rfsi_dev_forest_scratch_bytes(f, npix, nprobs) bytes
 * (only needed when out_q != NULL). visits (nullable, device uint64) accumulates the
 * number of internal nodes visited (the roofline's algorithmic figure). */
size_t rfsi_dev_forest_scratch_bytes(const rfsi_forest_t* f, int64_t npix, int32_t nprobs);
int rfsi_dev_forest_predict(const rfsi_forest_t* f, const double* X, int64_t npix, int64_t ldx,
                            double* out_mean, const double* probs_host, int32_t nprobs,
                            double* out_q, int32_t* out_terminal, void* scratch,
                            unsigned long long* visits, int32_t* na_flag,
                            int device, void* stream);

/* ---- fused stage 1 -> stage 2 over pixels x days ----------------------------------
 * One call = what the mapping loops do per day (temp_croatia/2_predictions.R:73-183,
 * prcp_catalonia/5_prcp_prediction.R:244-309): filter that day's stations
 * (!is.na(value)), near.obs, cbind the covariates, predict mean and/or quantiles --
 * for every day of the batch, without materialising the feature table. */
/* ---- covariate rasters -------------------------------------------------------------
 * raster::extract(r, points) with the default method ("simple": the value of the cell a
 * point falls in), as the mapping scripts do for every covariate of every day before near.obs
 * (prcp_catalonia/5_prcp_prediction.R:253-276: tmax/tmin on the DEM grid in tenths of a
 * degree, IMERG on a 0.1-degree global grid).  North-up raster: layer l, row r (0 = top),
 * column c at data[(l*nrow + r)*ncol + c]; cell (r, c) spans x_ul + c*dx .. x_ul + (c+1)*dx
 * and y_ul - (r+1)*dy .. y_ul - r*dy.  value = raw*scale + offset -- or, with
 * RFSI_RASTER_DIVIDE or-ed into dtype, raw/scale + offset: the scripts' `extract(tmin, gg) / 10`
 * (5_prcp_prediction.R:254,256) is a division, and x/10 != x*0.1 in the last bit; raw == nodata
 * or a point outside the raster gives NA. */
enum { RFSI_F64 = 0, RFSI_F32 = 1, RFSI_I32 = 2, RFSI_I16 = 3, RFSI_U16 = 4, RFSI_U8 = 5 };
#define RFSI_RASTER_DIVIDE 0x100
typedef struct rfsi_raster {
    const void* data;
    int32_t dtype;   /* RFSI_F64 ... RFSI_U8 [| RFSI_RASTER_DIVIDE] */
    int32_t nlayers; /* 1 = static; otherwise one layer per day */
    int64_t ncol, nrow;
    double x_ul, y_ul, dx, dy; /* outer corner of the top-left cell; cell size (both > 0) */
    double nodata;             /* NaN = no nodata value */
    double scale, offset;
} rfsi_raster_t;

/* out[p] = extract(layer `layer` of r, location p); locations explicit (loc_x/loc_y) or, when
 * loc_x == NULL, the cells [pix_begin, pix_begin + npix) of the row-major grid
 * (grid_x0 + col*grid_dx, grid_y0 + row*grid_dy), as in rfsi_fused_args_t. */
int rfsi_raster_extract(const rfsi_raster_t* r, int32_t layer, const double* loc_x, const double* loc_y,
                        int64_t npix, double grid_x0, double grid_y0, double grid_dx, double grid_dy,
                        int64_t grid_ncol, int64_t pix_begin, double* out, int device);

/* A covariate repair applied after sampling and before cbind, in the order given:
 *   cov[target] <- ifelse(cov[target] > cov[other], cov[other] - delta, cov[target])
 * i.e. `df$tmin <- ifelse(df$tmin > df$tmax, df$tmax-1, df$tmin)` (5_prcp_prediction.R:258) is
 * {target = tmin, other = tmax, delta = 1}; NA where either side is NA (R's ifelse). */
typedef struct rfsi_cov_fix {
    int32_t target, other; /* covariate column indices, both static or both per-day */
    double delta;
} rfsi_cov_fix_t;

typedef struct rfsi_fused_args {
    size_t struct_size; /* = sizeof(rfsi_fused_args_t) */
    /* prediction locations: explicit (loc_x/loc_y, npix) or, when loc_x == NULL, the
     * raster cells [pix_begin, pix_begin+npix) of a row-major grid: cell p -> row
     * p / grid_ncol, col p % grid_ncol, x = grid_x0 + col*grid_dx, y = grid_y0 + row*grid_dy.
     * With explicit coordinates grid_ncol > 1 is an optional hint: the locations are raster
     * cells stored row by row, grid_ncol per row (lets the engine map threads to compact 2-D
     * tiles, 2.4x faster traversal; results do not depend on it; the host entry point detects
     * it by itself) */
    const double* loc_x;
    const double* loc_y;
    int64_t npix;
    double grid_x0, grid_y0, grid_dx, grid_dy;
    int64_t grid_ncol;
    int64_t pix_begin;
    /* stations: coordinates (day-invariant) and the ndays x nobs table of observed
     * values, row d = day d, NA/NaN = no observation that day */
    const double* obs_x;
    const double* obs_y;
    int32_t nobs;
    const double* obs_z;
    int32_t ndays;
    int32_t n_obs;
    int32_t rm_dupl;
    /* covariates: ncov columns; column c is cov[c][d*cov_day_stride[c] + p]
     * (cov_day_stride 0 = static over days, npix = one layer per day) */
    int32_t ncov;
    const double* const* cov;
    const int64_t* cov_day_stride;
    /* model column f is feat_map[f] (RFSI_FEAT_COV / _DIST / _OBS) */
    const int32_t* feat_map;
    /* outputs (each nullable): mean[d*npix + p]; q[(j*ndays + d)*npix + p] */
    const double* probs;
    int32_t nprobs;
    double* out_mean;
    double* out_q;
    /* product epilogue (nullable): round-half-even(mean*100) as Int16, NA -> -32767
     * (5_prcp_prediction.R:288,303); iqr = round((q[last]-q[first])/2*100) */
    int16_t* out_mean_i16;
    int16_t* out_iqr_i16;
    /* optional pixel validity mask (nullable): 0 = skip pixel, outputs NA / -32767 */
    const uint8_t* pix_mask;
    /* optional (nullable) array of ncov raster descriptors: where cov_raster[c].data != NULL
     * covariate c is sampled on the device from that raster at the prediction locations
     * (cov[c] is then ignored) -- host-buffer entry point only */
    const rfsi_raster_t* cov_raster;
    /* optional (nullable) list of ncov_fix covariate repairs -- host-buffer entry point only
     * (the device form takes the caller's covariate columns as read-only) */
    const rfsi_cov_fix_t* cov_fix;
    int32_t ncov_fix;
    /* optional (nullable, npix each): the coordinates at which cov_raster is sampled when they
     * are not the near.obs coordinates -- the scripts keep the grid twice, `gg` in the rasters'
     * lon/lat for extract() and `gg2 <- spTransform(gg, utm31)` in metres for near.obs
     * (5_prcp_prediction.R:42-46,254-283) -- host-buffer entry point only */
    const double* cov_loc_x;
    const double* cov_loc_y;
} rfsi_fused_args_t;

int rfsi_predict_fused(const rfsi_forest_t* f, const rfsi_fused_args_t* a, int device);
/* The same call spread over several GPUs of one box: the pixels are cut into ndevices contiguous
 * bands (whole raster rows where the locations are raster cells), one host thread per band drives
 * its device, forest and station tables are replicated, and every band's results are copied
 * straight into its slice of the caller's arrays -- no collective, no reduction (the work
 * shards trivially; this is what lets ONE R process use all GPUs of the box, where the scripts
 * use foreach %dopar% over days, 5_prcp_prediction.R:243-244).  devices may repeat.  Returns the
 * first error of any band, else the first warning. */
int rfsi_predict_fused_multi(const rfsi_forest_t* f, const rfsi_fused_args_t* a, const int* devices, int ndevices);
/* same with every pointer in `a` (except probs, feat_map, cov, cov_day_stride, which
 * stay host arrays; cov[c] are device pointers) already on `device`.  workspace:
 * rfsi_dev_fused_workspace_bytes() bytes of device memory.  counters (nullable):
 * device uint64[4] = {internal nodes visited, pixel-days done, fallback pixel-days,
 * too-few-obs pixel-days}. */
size_t rfsi_dev_fused_workspace_bytes(const rfsi_forest_t* f, const rfsi_fused_args_t* a);
int rfsi_dev_predict_fused(const rfsi_forest_t* f, const rfsi_fused_args_t* a, void* workspace,
                           unsigned long long* counters, int device, void* stream);

/* number of kernels this library has launched in this process (bench accounting) */
uint64_t rfsi_launch_count(void);

/* Per-kernel device timing for roofline accounting: while enabled, every launch is
 * bracketed by CUDA events on its own stream.  rfsi_profile_read sums (and clears) the
 * elapsed time of one kernel class: 0 kNN geometry pass, 1 forest (matrix input),
 * 2 fused near.obs->forest, 3 fused full-scan fallback, 4 quantile epilogue, 5 near.obs. */
int rfsi_profile_enable(int on);
int rfsi_profile_read(int kernel_class, double* total_ms, uint64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* RFSI_B200_H */
