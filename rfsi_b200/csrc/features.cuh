// features.cuh -- near.obs for every day of a station table in one launch: the per-day
// feature table itself (dist1..distN, obs1..obsN), for callers that want the columns rather
// than predictions -- above all the training tables the reference builds with
//   foreach(t = time) %dopar% near.obs(locations = day_df, observations = day_df, ...)
// (temp_croatia/1_models.R:1034-1047; prcp_catalonia/4_prcp_case_study_catalonia_RK.R:1699-1712),
// where locations are the stations themselves and every row's own observation is excluded.
//
// Same exact scheme as the fused predictor (fused.cuh): the day-invariant candidate lists of
// the geometry pass are compacted per day over that day's valid stations; a list that runs dry
// is redone in place by a full scan of the day's valid stations.
#pragma once
#include "common.cuh"
#include "knn.cuh"

namespace rfsi {

struct FeatureParams {
    LocSpec loc;
    long long npix;
    int ndays;
    const double* cand_d2;
    const int32_t* cand_id;
    long long ld_cand;
    int KC, exhaustive;
    const double2* st_xy;
    const double* obs_z;   // [ndays][S]
    int S, n_obs, rm_dupl;
    int loc_is_obs;        // locations are the stations: row (d, p) exists only if z[d][p0 + p] is observed
    long long loc_obs_offset;  // p0: station index of this chunk's first location
    double* out_dist;      // [ndays][n_obs][npix]
    double* out_obs;
    int32_t* out_idx;      // nullable, 1-based ORIGINAL station index
    int* warn;
    unsigned long long* fallbacks;
};

constexpr int kFeatMaxK = 66;   // n_obs <= 64 (+1 self, +1 slack)

__global__ void __launch_bounds__(128)
features_kernel(const __grid_constant__ FeatureParams a)
{
    const long long tile = blockIdx.x / a.ndays;
    const int d = (int)(blockIdx.x - tile * a.ndays);
    const long long p = tile * 128 + threadIdx.x;
    if (p >= a.npix) return;
    const double* zday = a.obs_z + (long long)d * a.S;
    const long long plane = (long long)a.n_obs * a.npix;
    double* od = a.out_dist + (long long)d * plane + p;
    double* oo = a.out_obs + (long long)d * plane + p;
    int32_t* oi = a.out_idx ? a.out_idx + (long long)d * plane + p : nullptr;

    int taken = 0;
    bool row_exists = true;
    if (a.loc_is_obs) { const double zs = zday[a.loc_obs_offset + p]; row_exists = (zs == zs); }
    if (row_exists) {
        bool first_valid = true;
        for (int c = 0; c < a.KC; ++c) {
            const int32_t id = a.cand_id[(long long)c * a.ld_cand + p];
            if (id == kIdxNone) break;
            const double z = __ldg(zday + id);
            if (z != z) continue;
            const double d2 = a.cand_d2[(long long)c * a.ld_cand + p];
            if (first_valid) {
                first_valid = false;
                if (a.rm_dupl && d2 == 0.0) continue;
            }
            od[(long long)taken * a.npix] = __dsqrt_rn(d2);
            oo[(long long)taken * a.npix] = z;
            if (oi) oi[(long long)taken * a.npix] = id + 1;
            if (++taken == a.n_obs) break;
        }
        if (taken < a.n_obs && !a.exhaustive) {
            // rare: more of this pixel's candidates are missing today than the list has spares.
            // Full scan of today's valid stations with a private (local-memory) list.
            if (a.fallbacks) atomicAdd(a.fallbacks, 1ULL);
            const int k = a.n_obs + 1;
            double bd[kFeatMaxK];
            int32_t bi[kFeatMaxK];
            int cnt = 0;
            double px, py;
            get_loc(a.loc, p, px, py);
            for (int s = 0; s < a.S; ++s) {
                const double z = __ldg(zday + s);
                if (z != z) continue;
                const double2 xy = __ldg(a.st_xy + s);
                const double d2 = dist2_rn(px, py, xy.x, xy.y);
                if (cnt == k && !(d2 < bd[k - 1])) continue;
                int j = cnt < k ? cnt : k - 1;
                while (j > 0 && bd[j - 1] > d2) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j; }
                bd[j] = d2; bi[j] = s;
                if (cnt < k) ++cnt;
            }
            const int first = (a.rm_dupl && cnt > 0 && bd[0] == 0.0) ? 1 : 0;
            taken = 0;
            for (int j = first; j < cnt && taken < a.n_obs; ++j, ++taken) {
                od[(long long)taken * a.npix] = __dsqrt_rn(bd[j]);
                oo[(long long)taken * a.npix] = __ldg(zday + bi[j]);
                if (oi) oi[(long long)taken * a.npix] = bi[j] + 1;
            }
        }
        if (taken < a.n_obs && a.warn) *a.warn = 1;
    }
    for (int j = taken; j < a.n_obs; ++j) {
        od[(long long)j * a.npix] = na_real();
        oo[(long long)j * a.npix] = na_real();
        if (oi) oi[(long long)j * a.npix] = kNaInteger;
    }
}

}  // namespace rfsi
