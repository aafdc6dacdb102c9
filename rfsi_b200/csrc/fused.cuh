// fused.cuh -- stage 1 -> stage 2 without materialising the feature table.
//
// One CTA = a tile of P pixels on one day: what one iteration of the mapping loops does
// (temp_croatia/2_predictions.R:122-125,160-168,174-178;
//  prcp_catalonia/5_prcp_prediction.R:246-251,278-296): keep that day's stations
// (!is.na(value)), near.obs, cbind the covariates, predict the mean and/or the quantiles.
//
// Station geometry does not change with the day, only validity does.  A geometry pass
// (knn_brute_kernel, MODE RAW) therefore stores, once per pixel, the KC = n_obs+1+m
// nearest stations of ALL stations in (d2, index) order; per day the nearest VALID
// stations are the first valid entries of that list -- exactly the list a per-day search
// returns, because the list is a prefix of the full order.  A pixel-day whose list runs
// dry before n_obs valid stations are found (possible only when more than m of its
// candidates are missing that day) is appended to a fallback list and redone by the
// FALLBACK instantiation below with a full scan of that day's valid stations.
#pragma once
#include "common.cuh"
#include "forest.cuh"
#include "knn.cuh"

namespace rfsi {

constexpr int kMaxNObs = 64;
constexpr int kMaxCov = 32;
// Rank caches of the rank images (the tile in shared memory holds ranks, not values): a rank search per feature and
// pixel-day is the prologue's whole cost (43 searches whose loads diverge across the warp in the distance columns,
// profiles/r02_t_*), and most of them repeat.
//   * cand_rank: candidate c of a pixel lands in distance column c - s, s = the candidates skipped before it (no
//     observation that day, or the pixel's own station).  The rank of its distance in the columns c, c-1, c-2 is computed
//     ONCE per pixel next to the geometry pass (cand_rank_kernel) and read back by every day of the batch; s > 2 searches.
//   * obs_rank: the rank of station s's observation of day d in observation column j depends on no pixel at all:
//     obs_rank_kernel fills [day][station][column] once per launch.
//     A station without an observation that day has kObsRankNone in all its columns (a shifted rank ends in zeros).
constexpr int kCandShifts = 3;
constexpr uint32_t kObsRankNone = 0xFFFFFFFFu;

struct FusedParams {
    LocSpec loc;                 // only the FALLBACK scan needs coordinates
    long long npix;              // pixels in this chunk
    int day0, ndays;             // days [day0, day0+ndays) of the obs table are in this launch
    // candidates (RAW lists of the geometry pass)
    const double* cand_d2;
    const int32_t* cand_id;
    long long ld_cand;
    int KC;
    const uint32_t* cand_rank;   // [KC][kCandShifts][ld_cand] shifted ranks of sqrt(cand_d2) in column c - s (nullable)
    const uint32_t* obs_rank;    // [ndays][S][n_obs] shifted ranks of obs_z in column j (nullable)
    int exhaustive;              // 1 when the candidate list holds every station
    // stations
    const double2* st_xy;
    const double* obs_z;         // [total days][S]
    int S;
    int n_obs, rm_dupl;
    // covariates
    int ncov;
    const double* cov[kMaxCov];
    long long cov_stride[kMaxCov];
    short cov_row[kMaxCov];      // model column of covariate c (-1 unused)
    short dist_row[kMaxNObs];    // model column of dist{j+1} (-1 unused)
    short obs_row[kMaxNObs];
    const unsigned char* pix_mask;
    // forest
    const Node* nodes;           // wide format (nullptr when the compact image is in use)
    QForest q;                   // compact format
    const int32_t* leaf_rid;
    int T, F;
    int t_begin, t_end;          // this launch walks trees [t_begin, t_end) (one L2-sized chunk)
    double* acc;                 // [ndays*npix] running tree-order sums between chunk launches
    // outputs; row index = (d - day0)*npix + p relative to the *_base pointers
    double* out_mean;
    int16_t* out_mean_i16;
    uint32_t* leaf_ids;          // [ndays*npix][ld_ids] leaf index per tree (quantile pass; nullable)
    long long ld_ids;
    int ids_word;                // sector image: word of the leaf block stored into leaf_ids (forest.cuh)
    unsigned char* row_skip;     // [ndays*npix]
    // fallback list
    int2* fb_list;
    unsigned int* fb_count;
    unsigned int fb_capacity;
    unsigned long long* counters;  // {visits, pixel-days, fallback pixel-days, too-few pixel-days}
    int* na_flag;
    const int32_t* perm;         // unordered explicit locations: Morton-order permutation (else nullptr)
    int tile2d;                  // raster form: 16x16 pixel tiles of 8x4 warp patches
    long long tiles_per_row;     // tile2d only
};

template <int P>
__device__ __forceinline__ long long fused_pixel_of_thread(const FusedParams& a, long long tile) {
    if (!a.tile2d) {
        const long long i = tile * P + threadIdx.x;
        return (a.perm != nullptr && i < a.npix) ? (long long)a.perm[i] : i;
    }
    // 16 x (P/16) pixel tile; a warp covers an 8 x 4 patch
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const long long tr = tile / a.tiles_per_row, tc = tile - tr * a.tiles_per_row;
    const long long first_row = a.loc.pix_begin / a.loc.ncol;
    const long long col = tc * 16 + (w & 1) * 8 + (l & 7);
    const long long row = first_row + tr * (P / 16) + (w >> 1) * 4 + (l >> 3);
    if (col >= a.loc.ncol) return -1;
    return row * a.loc.ncol + col - a.loc.pix_begin;  // may be < 0 or >= npix: caller checks
}

__device__ __forceinline__ int16_t centi_i16(double v) {
    const double c = __dmul_rn(v, 100.0);
    if (c != c) return (int16_t)-32767;
    return (int16_t)fmin(fmax(rint(c), -32767.0), 32767.0);
}

// feature row r of this thread's pixel: fp64 value (wide format) or its rank (compact format)
template <int P, bool Q>
__device__ __forceinline__ void put_feature(const FusedParams& a, unsigned char* smem_raw, int r, double v) {
    if (Q) (reinterpret_cast<uint32_t*>(smem_raw) + threadIdx.x)[r * P] = rank_of(a.q, r, v) << a.q.rank_shift;
    else (reinterpret_cast<double*>(smem_raw) + threadIdx.x)[r * P] = v;
}

template <int P, bool Q>
__host__ __device__ constexpr size_t fused_feature_bytes(int F) {
    return Q ? (((size_t)F * P * sizeof(uint32_t) + 15) & ~(size_t)15) : (size_t)(F + 1) * P * sizeof(double);
}

// shared memory of a CTA: feature tile | neighbour lists of the fallback scan
template <int P, bool Q>
__host__ __device__ constexpr size_t fused_smem_bytes(int F, int k, bool fallback) {
    return fused_feature_bytes<P, Q>(F) + (fallback ? (size_t)k * P * (sizeof(double) + sizeof(int32_t)) : 0);
}

// Resident CTAs per SM the register budget is cut for (rank images).  Measured on the bench forest (profiles/r02_b):
// 3 x 256 threads at 80 registers 129.6 M pixel-days/s, 4 x 256 at 64 registers 127.5 M (register moves and the
// re-materialised shared-memory base go away, 92 KB instead of 60 KB of L1), 2 x 256 at 96 registers 106.9 M; more
// than 8 trees in flight per thread is slower at any occupancy (J = 12: 122.6 M, J = 16: 111.4 M).
template <int P, int J>
__host__ __device__ constexpr int fused_min_blocks_q() {
#ifdef RFSI_FUSED_MINB   // experiment builds (tools/build_variant.sh)
    return (P * 64 * 4 <= 65536) ? (J >= 16 ? 2 : (J >= 12 ? (RFSI_FUSED_MINB < 3 ? RFSI_FUSED_MINB : 3) : RFSI_FUSED_MINB)) : 2;
#else
    return (P * 64 * 4 <= 65536) ? (J >= 16 ? 2 : 3) : 2;
#endif
}

// ranks of every candidate's distance in the distance columns it can land in: one thread per (candidate, pixel) of
// the chunk, pixels fastest (the lists are candidate-major), candidates beyond column n.obs - 1 + kCandShifts not launched
__global__ void __launch_bounds__(256)
cand_rank_kernel(const __grid_constant__ FusedParams a, uint32_t* __restrict__ out, int ncand)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.npix * ncand) return;
    const int c = (int)(i / a.npix);
    const long long p = i - (long long)c * a.npix;
    if (a.cand_id[(long long)c * a.ld_cand + p] == kIdxNone) return;
    const double dist = __dsqrt_rn(a.cand_d2[(long long)c * a.ld_cand + p]);
#pragma unroll
    for (int s = 0; s < kCandShifts; ++s) {
        const int j = c - s;
        if (j < 0 || j >= a.n_obs) continue;
        const int rd = a.dist_row[j];
        if (rd >= 0) out[(long long)(c * kCandShifts + s) * a.ld_cand + p] = rank_of(a.q, rd, dist) << a.q.rank_shift;
    }
}

// ranks of every station's observation of every day of the launch in every observation column
__global__ void __launch_bounds__(256)
obs_rank_kernel(const __grid_constant__ FusedParams a, uint32_t* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)a.ndays * a.S * a.n_obs;
    if (i >= n) return;
    const int j = (int)(i % a.n_obs);
    const long long ds = i / a.n_obs;                        // (day of the launch) * S + station
    const double z = a.obs_z[(long long)a.day0 * a.S + ds];
    const int ro = a.obs_row[j];
    out[i] = (z != z) ? kObsRankNone : (ro >= 0 ? rank_of(a.q, ro, z) << a.q.rank_shift : 0u);
}

template <int P, int J, bool COUNT, bool FALLBACK, bool Q>
__global__ void __launch_bounds__(P, Q ? fused_min_blocks_q<P, J>() : ((P * (20 + 6 * J) <= 32768) ? 2 : 1))
fused_kernel(const __grid_constant__ FusedParams a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* fs = reinterpret_cast<double*>(smem_raw) + threadIdx.x;

    long long p;
    int d;
    bool active;
    if (!FALLBACK) {
        const long long tile = blockIdx.x / a.ndays;
        d = a.day0 + (int)(blockIdx.x - tile * a.ndays);
        p = fused_pixel_of_thread<P>(a, tile);
        active = p >= 0 && p < a.npix;
        if (active && a.pix_mask && !a.pix_mask[p]) {
            const long long row = (long long)(d - a.day0) * a.npix + p;
            if (a.out_mean) a.out_mean[row] = na_real();
            if (a.out_mean_i16) a.out_mean_i16[row] = (int16_t)-32767;
            if (a.row_skip) a.row_skip[row] = 1;
            active = false;
        }
    } else {
        const unsigned int n = min(*a.fb_count, a.fb_capacity);
        const unsigned int e = blockIdx.x * P + threadIdx.x;
        active = e < n;
        p = 0; d = 0;
        if (active) { const int2 pd = a.fb_list[e]; p = pd.x; d = pd.y; }
        if (blockIdx.x * P >= n) return;  // whole CTA idle
    }
    const long long row = (long long)(d - a.day0) * a.npix + p;
    const double* zday = a.obs_z + (long long)d * a.S;

    int taken = 0;
    bool ok = active;
    if (!FALLBACK) {
        if (active) {
            bool first_valid = true;
            const bool ranked = Q && a.cand_rank != nullptr;   // the distances are only read where their ranks are not at hand
            for (int c = 0; c < a.KC; ++c) {
                const int32_t id = a.cand_id[(long long)c * a.ld_cand + p];
                if (id == kIdxNone) break;
                double z = 0.0;
                uint32_t zr = 0u;
                if (Q && a.obs_rank != nullptr) {   // the rank table says "no observation" itself: the value is not read at all
                    zr = __ldg(a.obs_rank + ((long long)(d - a.day0) * a.S + id) * a.n_obs + taken);
                    if (zr == kObsRankNone) continue;
                } else {
                    z = __ldg(zday + id);
                    if (z != z) continue;  // no observation at this station today
                }
                if (first_valid) {
                    first_valid = false;
                    if (a.rm_dupl && a.cand_d2[(long long)c * a.ld_cand + p] == 0.0) continue;  // the pixel IS this station
                }
                const int rd = a.dist_row[taken], ro = a.obs_row[taken];
                if (rd >= 0) {
                    if (ranked && c - taken < kCandShifts)
                        (reinterpret_cast<uint32_t*>(smem_raw) + threadIdx.x)[rd * P] =
                            a.cand_rank[(long long)(c * kCandShifts + (c - taken)) * a.ld_cand + p];
                    else
                        put_feature<P, Q>(a, smem_raw, rd, __dsqrt_rn(a.cand_d2[(long long)c * a.ld_cand + p]));
                }
                if (ro >= 0) {
                    if (Q && a.obs_rank != nullptr) (reinterpret_cast<uint32_t*>(smem_raw) + threadIdx.x)[ro * P] = zr;
                    else put_feature<P, Q>(a, smem_raw, ro, z);
                }
                if (++taken == a.n_obs) break;
            }
            if (taken < a.n_obs) {
                ok = false;
                if (a.exhaustive) {  // genuinely fewer valid stations than n.obs(+1)
                    if (a.counters && a.t_begin == 0) atomicAdd(&a.counters[3], 1ULL);
                    if (a.out_mean) a.out_mean[row] = na_real();
                    if (a.out_mean_i16) a.out_mean_i16[row] = (int16_t)-32767;
                    if (a.row_skip) a.row_skip[row] = 1;
                } else if (a.t_begin == 0) {  // queued once; the fallback kernel follows every chunk
                    const unsigned int slot = atomicAdd(a.fb_count, 1u);
                    if (slot < a.fb_capacity) a.fb_list[slot] = make_int2((int)p, d);
                    if (a.row_skip) a.row_skip[row] = 0;
                }
            }
        }
    } else {
        // full scan of today's valid stations; lists live after the feature rows
        const int k = a.n_obs + 1;
        unsigned char* lists = smem_raw + fused_feature_bytes<P, Q>(a.F);
        double* ld2 = reinterpret_cast<double*>(lists) + threadIdx.x;
        int32_t* lid = reinterpret_cast<int32_t*>(reinterpret_cast<double*>(lists) + (size_t)k * P) + threadIdx.x;
        if (active) {
            double px, py;
            get_loc(a.loc, p, px, py);
            const double inf = __longlong_as_double(0x7FF0000000000000LL);
            for (int j = 0; j < k; ++j) { ld2[j * P] = inf; lid[j * P] = kIdxNone; }
            double thr = inf;
            for (int s = 0; s < a.S; ++s) {
                const double z = __ldg(zday + s);
                if (z != z) continue;
                const double2 xy = __ldg(a.st_xy + s);
                const double d2 = dist2_rn(px, py, xy.x, xy.y);
                if (d2 < thr) {
                    list_insert<P>(ld2, lid, k, d2, s);
                    thr = ld2[(k - 1) * P];
                }
            }
            const int first = (a.rm_dupl && lid[0] != kIdxNone && ld2[0] == 0.0) ? 1 : 0;
            for (int j = 0; j < a.n_obs; ++j) {
                const int32_t id = lid[(j + first) * P];
                if (id == kIdxNone) break;
                const int rd = a.dist_row[j], ro = a.obs_row[j];
                if (rd >= 0) put_feature<P, Q>(a, smem_raw, rd, __dsqrt_rn(ld2[(j + first) * P]));
                if (ro >= 0) put_feature<P, Q>(a, smem_raw, ro, __ldg(zday + id));
                ++taken;
            }
            if (a.counters && a.t_begin == 0) atomicAdd(&a.counters[2], 1ULL);
            if (taken < a.n_obs) {
                ok = false;
                if (a.counters && a.t_begin == 0) atomicAdd(&a.counters[3], 1ULL);
                if (a.out_mean) a.out_mean[row] = na_real();
                if (a.out_mean_i16) a.out_mean_i16[row] = (int16_t)-32767;
                if (a.row_skip) a.row_skip[row] = 1;
            }
        }
    }

    if (ok) {
        bool has_na = false;
        for (int c = 0; c < a.ncov; ++c) {
            const int r = a.cov_row[c];
            if (r < 0) continue;
            const double v = a.cov[c][(long long)d * a.cov_stride[c] + p];
            has_na |= (v != v);
            if (v == v) put_feature<P, Q>(a, smem_raw, r, v);
        }
        if (has_na) {  // ranger refuses NA in a used column: report it, emit NA
            ok = false;
            if (a.na_flag && a.t_begin == 0) atomicOr(a.na_flag, 1);   // bit 1 belongs to the geometry pass (NA coordinates)
            if (a.out_mean) a.out_mean[row] = na_real();
            if (a.out_mean_i16) a.out_mean_i16[row] = (int16_t)-32767;
            if (a.row_skip) a.row_skip[row] = 1;
        }
    }

    unsigned nvisit = 0;
    if (ok) {
        ForestOut o;
        o.mean = nullptr;
        o.leaf_ids = a.leaf_ids;
        o.ld_ids = a.ld_ids;
        o.terminal = nullptr;
        o.leaf_rid = nullptr;
        o.ld_term = 0;
        o.ids_word = a.ids_word;
        o.row_skip = nullptr;
        if (!Q) fs[a.F * P] = __longlong_as_double(0xFFF0000000000000LL);  // sentinel row: -inf
        double sum = a.t_begin > 0 ? a.acc[row] : 0.0;
        if (Q && a.q.sect)
            sum = traverse_sect<P, J, COUNT>(a.q, a.t_begin, a.t_end, a.T, sum,
                                             reinterpret_cast<const uint32_t*>(smem_raw) + threadIdx.x, o, row, nvisit);
        else if (Q && a.q.blocks)
            sum = traverse_b<P, J, COUNT>(a.q, a.t_begin, a.t_end, a.T, sum,
                                          reinterpret_cast<const uint32_t*>(smem_raw) + threadIdx.x, o, row, nvisit);
        else if (Q)
            sum = traverse_q<P, J, COUNT>(a.q, a.t_begin, a.t_end, a.T, sum,
                                          reinterpret_cast<const uint32_t*>(smem_raw) + threadIdx.x, o, row, nvisit);
        else
            sum = traverse_all<P, J, COUNT>(a.nodes, a.t_begin, a.t_end, a.T, a.F, sum, fs, o, row, nvisit);
        if (a.t_end == a.T) {
            const double mean = sum / (double)a.T;
            if (a.out_mean) a.out_mean[row] = mean;
            if (a.out_mean_i16) a.out_mean_i16[row] = centi_i16(mean);
            if (a.row_skip) a.row_skip[row] = 0;
        } else {
            a.acc[row] = sum;
        }
    }
    if (COUNT && a.counters) {
        const unsigned long long w = warp_sum_u64(nvisit);
        const unsigned long long done = warp_sum_u64((ok && a.t_end == a.T) ? 1ULL : 0ULL);
        if ((threadIdx.x & 31) == 0) {
            if (w) atomicAdd(&a.counters[0], w);
            if (done) atomicAdd(&a.counters[1], done);
        }
    }
}


}  // namespace rfsi
