// knn.cuh -- stage 1: exact n-nearest-observation search (near.obs / nabor::knn).
//
// Replaces meteo::near.obs -> nabor::knn (libnabo kd-tree) as called at
// simulation/simulation.R:191-196, sim_500.R:134-139, temp_croatia/2_predictions.R:160-165,
// prcp_catalonia/5_prcp_prediction.R:278-283.
//
// B200 design: brute force, one pixel per thread.  Station coordinates stream
// through shared memory in tiles moved by the TMA engine (cp.async.bulk + mbarrier,
// double buffered) and are read as 128-bit broadcasts; each thread owns a sorted
// (d2, index) list held column-wise in shared memory (conflict-free) with the
// current k-th best d2 cached in a register, so the hot loop is 5 fp64 ops + one
// compare per (pixel, station) and insertions (rare after warm-up) stay off it.
// Total order (d2, station index) ascending; d2 without FMA contraction.
#pragma once
#include "common.cuh"

namespace rfsi {

struct LocSpec {
    const double* x;  // explicit coordinates (device), or nullptr for the raster form
    const double* y;
    double x0, y0, dx, dy;
    long long ncol;
    long long pix_begin;
};

__device__ __forceinline__ void get_loc(const LocSpec& L, long long p, double& px, double& py) {
    if (L.x != nullptr) {
        px = L.x[p];
        py = L.y[p];
    } else {
        const long long g = L.pix_begin + p;
        const long long r = g / L.ncol;
        const long long c = g - r * L.ncol;
        px = __dadd_rn(L.x0, __dmul_rn((double)c, L.dx));
        py = __dadd_rn(L.y0, __dmul_rn((double)r, L.dy));
    }
}

// ---- mbarrier / bulk-copy (TMA 1-D) primitives ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- per-thread sorted list in shared memory -------------------------------------
// entry j of this thread's list lives at ld2[j*P], lid[j*P] (ld2/lid already offset
// by threadIdx.x).  Precondition: (d2, id) sorts before the current last entry.
template <int P>
__device__ __forceinline__ void list_insert(double* ld2, int32_t* lid, int k, double d2, int32_t id) {
    int j = k - 1;
    while (j > 0) {
        const double pd = ld2[(j - 1) * P];
        const int32_t pi = lid[(j - 1) * P];
        if (pd < d2 || (pd == d2 && pi < id)) break;
        ld2[j * P] = pd;
        lid[j * P] = pi;
        --j;
    }
    ld2[j * P] = d2;
    lid[j * P] = id;
}

constexpr int kKnnTile = 1024;  // stations per shared-memory tile (16 KB)

template <int P>
__host__ __device__ constexpr size_t knn_smem_bytes(int k) {
    return 16 + 2 * (size_t)kKnnTile * sizeof(double2) + (size_t)k * P * (sizeof(double) + sizeof(int32_t));
}

enum KnnMode { KNN_RAW = 0, KNN_FINAL = 1 };

// MODE RAW  : write the k-entry lists as-is: out_d2[j*ld + p], out_id[j*ld + p]
//             (squared distances; empty slots +inf / kIdxNone) -- the day-invariant
//             candidate lists of the fused path.
// MODE FINAL: near.obs proper -- drop a zero-distance first neighbour when rm_dupl,
//             sqrt, gather the observed values: out_dist/out_obs[j*ld + p], j < k-1,
//             out_idx 1-based (nullable); short lists are NA-filled and *warn set.
template <int P, int MODE>
__global__ void __launch_bounds__(P)
knn_brute_kernel(LocSpec L, long long npix, const double2* __restrict__ st_xy, int S, int k,
                 const double* __restrict__ st_z, int rm_dupl,
                 double* __restrict__ out_a, double* __restrict__ out_b, int32_t* __restrict__ out_i,
                 long long ld, int* __restrict__ warn, int* __restrict__ na_flag)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    double2* tile0 = reinterpret_cast<double2*>(smem_raw + 16);
    double2* tile1 = tile0 + kKnnTile;
    double* ld2 = reinterpret_cast<double*>(tile1 + kKnnTile) + threadIdx.x;
    int32_t* lid = reinterpret_cast<int32_t*>(reinterpret_cast<double*>(tile1 + kKnnTile) + (size_t)k * P) +
                   threadIdx.x;

    const int tid = threadIdx.x;
    const long long p = (long long)blockIdx.x * P + tid;
    const bool in_range = p < npix;
    double px = 0.0, py = 0.0;
    if (in_range) get_loc(L, p, px, py);
    if (in_range && (px != px || py != py) && na_flag) atomicOr(na_flag, 2);   // bit 1: a location has an NA coordinate

    for (int j = 0; j < k; ++j) {
        ld2[j * P] = __longlong_as_double(0x7FF0000000000000LL);  // +inf
        lid[j * P] = kIdxNone;
    }
    double thr = __longlong_as_double(0x7FF0000000000000LL);

    const int ntiles = (S + kKnnTile - 1) / kKnnTile;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0 && ntiles > 0) {
        const int n0 = min(S, kKnnTile);
        mbar_expect_tx(&bar[0], (uint32_t)(n0 * sizeof(double2)));
        bulk_g2s(tile0, st_xy, (uint32_t)(n0 * sizeof(double2)), &bar[0]);
    }
    for (int t = 0; t < ntiles; ++t) {
        if (tid == 0 && t + 1 < ntiles) {
            const int base1 = (t + 1) * kKnnTile;
            const int n1 = min(S - base1, kKnnTile);
            uint64_t* b = &bar[(t + 1) & 1];
            mbar_expect_tx(b, (uint32_t)(n1 * sizeof(double2)));
            bulk_g2s(((t + 1) & 1) ? tile1 : tile0, st_xy + base1, (uint32_t)(n1 * sizeof(double2)), b);
        }
        mbar_wait(&bar[t & 1], (uint32_t)((t >> 1) & 1));
        const double2* tile = (t & 1) ? tile1 : tile0;
        const int base = t * kKnnTile;
        const int n = min(S - base, kKnnTile);
        if (in_range) {
#pragma unroll 4
            for (int i = 0; i < n; ++i) {
                const double2 s = tile[i];
                const double d2 = dist2_rn(px, py, s.x, s.y);
                // ascending station index + strict '<' == (d2, index) order
                if (d2 < thr) {
                    list_insert<P>(ld2, lid, k, d2, base + i);
                    thr = ld2[(k - 1) * P];
                }
            }
        }
        __syncthreads();  // everyone is done with this buffer before it is refilled
    }
    if (!in_range) return;

    if (MODE == KNN_RAW) {
        for (int j = 0; j < k; ++j) {
            out_a[(long long)j * ld + p] = ld2[j * P];
            out_i[(long long)j * ld + p] = lid[j * P];
        }
    } else {
        const int n_obs = k - 1;
        const int first = (rm_dupl && lid[0] != kIdxNone && ld2[0] == 0.0) ? 1 : 0;
        bool shortl = false;
        for (int j = 0; j < n_obs; ++j) {
            const int32_t id = lid[(j + first) * P];
            const long long o = (long long)j * ld + p;
            if (id != kIdxNone) {
                out_a[o] = __dsqrt_rn(ld2[(j + first) * P]);
                out_b[o] = st_z[id];
                if (out_i) out_i[o] = id + 1;
            } else {
                out_a[o] = na_real();
                out_b[o] = na_real();
                if (out_i) out_i[o] = kNaInteger;
                shortl = true;
            }
        }
        if (shortl && warn) atomicOr(warn, 1);                                   // bit 0: fewer neighbours than n.obs
    }
}

}  // namespace rfsi
