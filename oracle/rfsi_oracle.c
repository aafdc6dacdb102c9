/*
 * rfsi_oracle.c -- CPU restatement of the RFSI prediction hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rfsi_b200/ (the product) may import,
 * link or execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * PINNING: stage 1 (near.obs) IS pinned against an output of the reference itself -- the 55 746
 * stored predictions of its nested 10-fold IDW cross-validation (temp_croatia/1_models.R:405-536,
 * temp_data/IDW_10f_pred.rda, computed in R by gstat), each a function of exactly near.obs's output
 * (the n nearest observed stations, their distances and values), are reproduced to 2e-14
 * (tests/test_oracle.py, tests/util.py::reference_idw_cv).  Stage 2 (ranger predict / quantiles) is
 * PARITY UNPINNED: the reference (AleksandarSekulic/RFSI) holds no saved forest, golden
 * vectors, known-answer tests or fixtures for it (SURVEY.md section 4 and
 * 8c), and its arithmetic lives in un-vendored, un-pinned R packages
 * (meteo::near.obs, nabor::knn / libnabo, ranger::predict.ranger, stats::quantile)
 * that cannot be run in this container (no R).  This file restates the published
 * semantics of those packages, anchored on the reference's own call sites:
 *
 *   near.obs(locations, observations, zcol, n.obs)
 *       simulation/simulation.R:191-196, simulation/sim_500.R:134-139,
 *       temp_croatia/2_predictions.R:160-165,
 *       prcp_catalonia/5_prcp_prediction.R:278-283,
 *       temp_croatia/1_models.R:1040-1045 (locations == observations)
 *   predict(<ranger>, df)$predictions
 *       simulation/simulation.R:214, temp_croatia/2_predictions.R:167,
 *       prcp_catalonia/5_prcp_prediction.R:287
 *   predict(<ranger>, df, type="quantiles", quantiles=c(.25,.5,.75))
 *       temp_croatia/2_predictions.R:174-176,
 *       prcp_catalonia/5_prcp_prediction.R:289-291
 *   iqr <- (q[,3]-q[,1])/2 ; round(iqr*100)
 *       temp_croatia/2_predictions.R:177-178,
 *       prcp_catalonia/5_prcp_prediction.R:292-296
 *
 * What pins it instead (tests/test_oracle.py): hand-computed cases, R
 * quantile() type-7 known answers, scipy cKDTree distances, sklearn tree
 * descent on fp32-representable inputs, and an independent numpy restatement
 * (oracle/rfsi_oracle_np.py) plus an independent kd-tree/threaded C++
 * restatement (oracle/cpu_baseline.cpp) that must agree bit for bit.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; no -ffast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_OK 0
#define ORACLE_E_ARG (-3)
#define ORACLE_W_TOO_FEW_OBS 1

/* R's NA_real_: a quiet NaN whose low word is 1954 (R arithmetic.c R_NaReal). */
static double r_na_real(void) {
    union { uint64_t u; double d; } v;
    v.u = 0x7FF00000000007A2ULL;
    return v.d;
}
#define R_NA_INTEGER INT32_MIN

double oracle_na_real(void) { return r_na_real(); }

/* ------------------------------------------------------------------------- */
/* Stage 1: near.obs                                                          */
/* ------------------------------------------------------------------------- */
/*
 * For every location the k = n_obs + 1 nearest observation points under the
 * total order (d2 ascending, then 0-based observation index ascending), d2 =
 * dx*dx + dy*dy evaluated in fp64 with the x term first and no FMA
 * contraction (libnabo compares on squared distances and takes the root at the
 * end).  If rm_dupl and the nearest has
 * d2 == 0 (the location IS an observation point: temp_croatia/1_models.R:1040
 * passes locations == observations, and simulation.R:122-125 draws stations
 * from the pixel centres) neighbour 1 is dropped and 2..n_obs+1 are kept,
 * otherwise 1..n_obs.  Fewer points than needed: remaining slots are NA_real_ /
 * NA_integer_ and the return value is ORACLE_W_TOO_FEW_OBS (A.1 item 6).
 *
 * Outputs are column-major npix x n_obs, i.e. column j (neighbour j+1) starts
 * at out[j*npix]: exactly the dist1..distN / obs1..obsN data.frame columns
 * (sim_500.R:146 relies on that order).  out_idx (nullable) is 1-based.
 */
int oracle_near_obs(const double* loc_x, const double* loc_y, int64_t npix,
                    const double* obs_x, const double* obs_y, const double* obs_z,
                    int32_t nobs, int32_t n_obs, int32_t rm_dupl,
                    double* out_dist, double* out_obs, int32_t* out_idx)
{
    if (npix < 0 || nobs < 0 || n_obs < 1) return ORACLE_E_ARG;
    const int k = n_obs + 1;
    double* bd = (double*)malloc(sizeof(double) * (size_t)k);
    int32_t* bi = (int32_t*)malloc(sizeof(int32_t) * (size_t)k);
    int rc = ORACLE_OK;
    for (int64_t p = 0; p < npix; ++p) {
        int cnt = 0;
        const double px = loc_x[p], py = loc_y[p];
        for (int32_t s = 0; s < nobs; ++s) {
            const double dx = px - obs_x[s];
            const double dy = py - obs_y[s];
            const double xx = dx * dx;
            const double yy = dy * dy;
            const double d2 = xx + yy;
            /* ascending-index scan + strict '<' == (d2, index) lexicographic */
            if (cnt == k && !(d2 < bd[k - 1])) continue;
            int j = (cnt < k) ? cnt : k - 1;
            while (j > 0 && bd[j - 1] > d2) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j; }
            bd[j] = d2; bi[j] = s;
            if (cnt < k) ++cnt;
        }
        const int first = (rm_dupl && cnt > 0 && bd[0] == 0.0) ? 1 : 0;
        for (int j = 0; j < n_obs; ++j) {
            const int src = j + first;
            const int64_t o = (int64_t)j * npix + p;
            if (src < cnt) {
                out_dist[o] = sqrt(bd[src]);
                out_obs[o] = obs_z[bi[src]];
                if (out_idx) out_idx[o] = bi[src] + 1;
            } else {
                out_dist[o] = r_na_real();
                out_obs[o] = r_na_real();
                if (out_idx) out_idx[o] = R_NA_INTEGER;
                rc = ORACLE_W_TOO_FEW_OBS;
            }
        }
    }
    free(bd); free(bi);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* Stage 2: predict.ranger, regression                                        */
/* ------------------------------------------------------------------------- */
/*
 * Forest in ranger's own layout, trees concatenated:
 * tree t owns nodes [node_offsets[t], node_offsets[t+1]); left/right are the
 * tree-local child ids (child.nodeIDs[[t]][[1]] / [[2]]), both 0 at a terminal
 * node; split_var the 0-based column of X; split_val the threshold, or the
 * prediction at a terminal node.
 */
static inline int32_t descend(const int32_t* left, const int32_t* right,
                              const int32_t* svar, const double* sval,
                              const double* X, int64_t ldx, int64_t row,
                              int64_t* visits)
{
    int32_t node = 0;
    while (left[node] != 0 || right[node] != 0) {
        const double v = X[(int64_t)svar[node] * ldx + row];
        node = (v <= sval[node]) ? left[node] : right[node];
        if (visits) ++*visits;
    }
    return node;
}

/* prediction[row] = (sum over trees in tree order of the leaf value) / ntrees
 * (Appendix A.2).  X is column-major with leading dimension ldx.
 * out_visits (nullable) receives the total number of internal nodes visited,
 * the figure bench.py's roofline uses. */
int oracle_forest_predict(int32_t ntrees, const int64_t* node_offsets,
                          const int32_t* left, const int32_t* right,
                          const int32_t* split_var, const double* split_val,
                          const double* X, int64_t npix, int64_t ldx,
                          double* out, int64_t* out_visits)
{
    if (ntrees < 1 || npix < 0 || ldx < npix) return ORACLE_E_ARG;
    int64_t visits = 0;
    for (int64_t r = 0; r < npix; ++r) {
        double sum = 0.0;
        for (int32_t t = 0; t < ntrees; ++t) {
            const int64_t o = node_offsets[t];
            const int32_t leaf = descend(left + o, right + o, split_var + o, split_val + o,
                                         X, ldx, r, out_visits ? &visits : NULL);
            sum += split_val[o + leaf];
        }
        out[r] = sum / (double)ntrees;
    }
    if (out_visits) *out_visits = visits;
    return ORACLE_OK;
}

/* terminal node ids, 0-based tree-local, column-major npix x ntrees
 * (predict(type="terminalNodes")). */
int oracle_forest_terminal_nodes(int32_t ntrees, const int64_t* node_offsets,
                                 const int32_t* left, const int32_t* right,
                                 const int32_t* split_var, const double* split_val,
                                 const double* X, int64_t npix, int64_t ldx,
                                 int32_t* out)
{
    if (ntrees < 1 || npix < 0 || ldx < npix) return ORACLE_E_ARG;
    for (int32_t t = 0; t < ntrees; ++t) {
        const int64_t o = node_offsets[t];
        for (int64_t r = 0; r < npix; ++r)
            out[(int64_t)t * npix + r] =
                descend(left + o, right + o, split_var + o, split_val + o, X, ldx, r, NULL);
    }
    return ORACLE_OK;
}

static int cmp_double(const void* a, const void* b) {
    const double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

/* R stats::quantile(x, probs, na.rm=TRUE), type 7, on m non-NA values
 * (Appendix A.3 item 3): index = 1 + (m-1)p; lo = floor, hi = ceiling,
 * h = index - lo; q = x[lo] unless (index > lo and x[hi] != x[lo]), in which
 * case q = (1-h)*x[lo] + h*x[hi] -- this expression, in this order. */
double oracle_quantile7_sorted(const double* xs, int32_t m, double p) {
    if (m <= 0) return r_na_real();
    const double index = 1.0 + (double)(m - 1) * p;
    const double lo = floor(index), hi = ceil(index);
    const double xlo = xs[(int32_t)lo - 1], xhi = xs[(int32_t)hi - 1];
    if (index > lo && xhi != xlo) {
        const double h = index - lo;
        return (1.0 - h) * xlo + h * xhi;
    }
    return xlo;
}

/* predict(type="quantiles"): per row gather random_node_values[leaf, tree]
 * over the trees, drop NA, sort, type-7 quantiles.  rnv_offsets[t] is where
 * tree t's column starts in random_node_values (t*nrow for the R matrix).
 * out is column-major npix x nprobs. */
int oracle_forest_quantiles(int32_t ntrees, const int64_t* node_offsets,
                            const int32_t* left, const int32_t* right,
                            const int32_t* split_var, const double* split_val,
                            const int64_t* rnv_offsets, const double* random_node_values,
                            const double* X, int64_t npix, int64_t ldx,
                            const double* probs, int32_t nprobs, double* out)
{
    if (ntrees < 1 || npix < 0 || ldx < npix || nprobs < 1) return ORACLE_E_ARG;
    double* v = (double*)malloc(sizeof(double) * (size_t)ntrees);
    for (int64_t r = 0; r < npix; ++r) {
        int32_t m = 0;
        for (int32_t t = 0; t < ntrees; ++t) {
            const int64_t o = node_offsets[t];
            const int32_t leaf = descend(left + o, right + o, split_var + o, split_val + o,
                                         X, ldx, r, NULL);
            const double s = random_node_values[rnv_offsets[t] + leaf];
            if (!isnan(s)) v[m++] = s;
        }
        qsort(v, (size_t)m, sizeof(double), cmp_double);
        for (int32_t q = 0; q < nprobs; ++q)
            out[(int64_t)q * npix + r] = oracle_quantile7_sorted(v, m, probs[q]);
    }
    free(v);
    return ORACLE_OK;
}

/* The mapping scripts' product step:
 * round(x*100) with R's round-half-even, Int16, NAflag -32767
 * (5_prcp_prediction.R:288,296-303; 2_predictions.R:172-173,178). */
int oracle_centi_int16(const double* x, int64_t n, int16_t* out) {
    for (int64_t i = 0; i < n; ++i) {
        const double v = x[i] * 100.0;
        if (isnan(v)) { out[i] = -32767; continue; }
        double r = nearbyint(v);           /* default rounding mode: half-even */
        if (r > 32767.0) r = 32767.0;
        if (r < -32767.0) r = -32767.0;
        out[i] = (int16_t)r;
    }
    return ORACLE_OK;
}
