// forest.cuh -- stage 2: random-forest / quantile-forest forward pass.
//
// Replaces ranger::predict.ranger (regression, type="response" / "quantiles" /
// "terminalNodes") as called at simulation/simulation.R:214,
// temp_croatia/2_predictions.R:167,174-176 and
// prcp_catalonia/5_prcp_prediction.R:287-291.
//
// B200 design: one pixel per thread, a tile of P pixels per CTA.  The tile's
// feature vectors sit in shared memory column-major (feature f of pixel i at
// fs[f*P + i]: any mix of f across a warp is bank-conflict free), the forest is
// a flat array of 16-byte nodes fetched with one 128-bit read-only load per
// visit, siblings adjacent so both outcomes of a split share a 32-byte sector,
// roots of all trees contiguous.  Each thread walks J trees at once so J
// independent dependent-load chains are in flight per thread; spatially adjacent
// pixels in a warp take the same branches almost everywhere, so a warp's 32 loads
// collapse to a few sectors (served by L1/L2).  Leaf values are summed in tree
// order in fp64, which makes the mean bit-identical to the sequential CPU sum.
#pragma once
#include "common.cuh"
#include "sector.cuh"

namespace rfsi {

// device image of the compact format
struct QForest {
    const uint2* nodes;       // [n] {rank, delta<<12 | feat}
    const QRoot* roots;       // [T]
    const double* leaf_val;   // [n] prediction at leaf indices
    const double* uvals;      // sorted unique thresholds, features concatenated
    const int32_t* uoff;      // [F+1]
    const double* rpar;       // [F][4] bucket function of feature f: {lo/2, scale, buckets - 1, offset of its table in rtab}
    const int32_t* rtab;      // per feature [buckets+1] number of thresholds in buckets < b
    const uint32_t* blocks;   // blocked format (common.cuh): 32 words per block; nullptr = use `nodes`
    const uint32_t* broot;    // [T] root block of each tree (or a leaf marker)
    int rank_shift;           // ranks are stored in shared memory shifted left by this (8 when blocked / sector)
    const uint32_t* sect;     // sector format (sector.cuh): 8 words per block; nullptr = use `blocks` / `nodes`
    const uint32_t* sroot;    // [T] root block of each tree
    uint32_t sect_block_bytes;  // 32, and the byte stride between features of the rank tile (4 * pixels per CTA):
    uint32_t sect_tile_stride;  // run-time values on purpose, see SectorView
};

// number of thresholds of feature f strictly below x (x is never NaN here).  The bucket
// function is monotone, so every threshold in an earlier bucket is < x and every one in a later
// bucket is > x: only x's own bucket (a threshold or two: about one bucket per threshold) needs the binary search.
__device__ __forceinline__ uint32_t rank_of(const QForest& q, int f, double x) {
    const double* __restrict__ u = q.uvals + q.uoff[f];
    const double2 p0 = __ldg(reinterpret_cast<const double2*>(q.rpar) + 2 * f);
    const double2 p1 = __ldg(reinterpret_cast<const double2*>(q.rpar) + 2 * f + 1);
    const int b = rank_bucket(x, p0.x, p0.y, p1.x);
    const int32_t* __restrict__ tab = q.rtab + (long long)p1.y;
    int lo = __ldg(tab + b);
    int n = __ldg(tab + b + 1) - lo;
    while (n > 0) {
        const int half = n >> 1;
        if (__ldg(u + lo + half) < x) { lo += half + 1; n -= half + 1; } else { n = half; }
    }
    return (uint32_t)lo;
}

struct ForestOut {
    double* mean;             // [npix] (nullable)
    uint32_t* leaf_ids;       // [npix][ld_ids] pixel-major leaf (device-node) index per tree, for the quantile pass (nullable)
    long long ld_ids;         // row stride of leaf_ids in entries: T rounded up to 8, so whole groups are 32-byte aligned
    int32_t* terminal;        // [T][ld_term] tree-local terminal node ids (nullable)
    const int32_t* leaf_rid;  // per device-node ranger node id, needed if terminal
    long long ld_term;
    int ids_word;             // sector image: the word of the leaf block that goes into leaf_ids (leaf number, or
                              // the quantile key itself: the quantile kernel then has nothing left to gather)
    unsigned char* row_skip;  // [npix] matrix-input kernels: 1 = the row has NA in a used column (nullable)
};

// Quantile path: the leaves one group of J trees ended in, 4 bytes per tree (the sample
// random.node.values[leaf, tree] is gathered by the quantile kernel, where the T gathers of a
// row are independent loads of one warp instead of a dependent tail of this walk).  A whole
// group of 8 is one 32-byte sector, written by one 256-bit store; the scratch is read once by
// the next kernel and is larger than L2, so the stores are evict-first and do not push the
// forest image out of L2.
template <int J, class Idx>
__device__ __forceinline__ void store_leaf_ids(const ForestOut& o, long long base, int remaining, const Idx (&idx)[J]) {
    uint32_t* dst = o.leaf_ids + base;
    if (J % 8 == 0 && remaining >= J && (base & 7) == 0) {
#pragma unroll
        for (int j = 0; j < J; j += 8)
            asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + j),
                         "r"((uint32_t)idx[j]), "r"((uint32_t)idx[j + 1]), "r"((uint32_t)idx[j + 2]), "r"((uint32_t)idx[j + 3]),
                         "r"((uint32_t)idx[j + 4]), "r"((uint32_t)idx[j + 5]), "r"((uint32_t)idx[j + 6]), "r"((uint32_t)idx[j + 7])
                         : "memory");
    } else if (J % 4 == 0 && remaining >= J && (base & 3) == 0) {
#pragma unroll
        for (int j = 0; j < J; j += 4)
            __stcs(reinterpret_cast<uint4*>(dst + j),
                   make_uint4((uint32_t)idx[j], (uint32_t)idx[j + 1], (uint32_t)idx[j + 2], (uint32_t)idx[j + 3]));
    } else {
#pragma unroll
        for (int j = 0; j < J; ++j)
            if (j < remaining) dst[j] = (uint32_t)idx[j];
    }
}

// Walk trees [t_begin, t_end) for one pixel whose features are fs[f*P]; `sum` carries the
// tree-order partial sum of the trees before and the updated sum is returned.
//
// Leaves are absorbing: a leaf's `feat` names the sentinel feature row F (which holds -inf
// for every pixel) and its `child` is its own index, so "x > val ? child+1 : child" maps a
// leaf to itself.  The inner loop therefore has no per-chain leaf test and no divergent
// branch: J loads, J (LDS, compare, add), one min-reduction of the J feature ids to see
// whether any chain is still at an internal node.
template <int P, int J, bool COUNT>
__device__ __forceinline__ double traverse_all(const Node* __restrict__ nodes, int t_begin, int t_end, int T,
                                               int F, double sum, const double* __restrict__ fs,
                                               const ForestOut& o, long long row, unsigned& nvisit)
{
    for (int t0 = t_begin; t0 < t_end; t0 += J) {
        int idx[J];
        Node nd[J];
#pragma unroll
        for (int j = 0; j < J; ++j) idx[j] = min(t0 + j, t_end - 1);  // the root of tree t is node t
        int mn;
        do {
#pragma unroll
            for (int j = 0; j < J; ++j) nd[j] = load_node(nodes + idx[j]);
            mn = F;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const double x = fs[nd[j].feat * P];
                idx[j] = nd[j].child + ((x <= nd[j].val) ? 0 : 1);
                mn = min(mn, nd[j].feat);
                if (COUNT) nvisit += (nd[j].feat != F) ? 1u : 0u;
            }
        } while (mn != F);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            if (t0 + j < t_end) {
                sum += nd[j].val;
                if (o.terminal) o.terminal[(long long)(t0 + j) * o.ld_term + row] = o.leaf_rid[idx[j]];
            }
        }
        if (o.leaf_ids) store_leaf_ids<J>(o, row * o.ld_ids + t0, t_end - t0, idx);
    }
    return sum;
}

// Compact-format walk.  rs = this thread's column of the rank tile in shared memory
// (feature-major: feature f of pixel i at word f*P + i, so any mix of features across a
// warp is bank-conflict free; a pixel-major layout saved an instruction but cost 40 %
// extra shared-memory wavefronts when lanes read different features -- profiles/r01_e).
// Per visit: one 64-bit node load, shift + IMAD + LDS.32, one AND, and an add-with-carry pair that folds "rank(x) > rank(thr)" into the index update
// (add.cc of x's rank and the node's complemented rank sets the carry, addc consumes it) -- no predicates, no branches.

// nrank = ~rank is what the node stores: xr + ~rank carries out of 32 bits exactly when
// xr > rank, and addc folds that carry into idx + delta.
__device__ __forceinline__ uint32_t q_step(uint32_t idx, uint32_t delta, uint32_t nrank, uint32_t xr) {
    uint32_t out;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %3, %4;\n\t}"
        : "=r"(out) : "r"(xr), "r"(nrank), "r"(idx), "r"(delta));
    return out;  // idx + delta + (xr > rank)
}

template <int P, int J, bool COUNT>
__device__ __forceinline__ double traverse_q(const QForest& q, int t_begin, int t_end, int T, double sum,
                                             const uint32_t* __restrict__ rs, const ForestOut& o, long long row,
                                             unsigned& nvisit)
{
    const char* rsb = reinterpret_cast<const char*>(rs);
    for (int t0 = t_begin; t0 < t_end; t0 += J) {
        uint32_t idx[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int t = min(t0 + j, t_end - 1);
            const uint4 r = __ldg(reinterpret_cast<const uint4*>(q.roots + t));  // {rank, feat*4, child, -}
            const uint32_t xr = *reinterpret_cast<const uint32_t*>(rsb + r.y * P);
            idx[j] = q_step(r.z, 0u, r.x, xr);
            if (COUNT) nvisit += (r.x != ~kQLeafRank) ? 1u : 0u;
        }
        uint32_t any;
        do {
            uint2 w[J];
#pragma unroll
            for (int j = 0; j < J; ++j) w[j] = __ldg(q.nodes + idx[j]);
            any = 0;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const uint32_t xr = *reinterpret_cast<const uint32_t*>(rsb + (w[j].y >> kQDeltaBits) * P);
                idx[j] = q_step(idx[j], w[j].y & (kQMaxDelta - 1), w[j].x, xr);
                any |= w[j].y;
                if (COUNT) nvisit += (w[j].y != 0) ? 1u : 0u;
            }
        } while (any != 0);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            if (t0 + j < t_end) {
                sum += __ldg(q.leaf_val + idx[j]);
                if (o.terminal) o.terminal[(long long)(t0 + j) * o.ld_term + row] = o.leaf_rid[idx[j]];
            }
        }
        if (o.leaf_ids) store_leaf_ids<J>(o, row * o.ld_ids + t0, t_end - t0, idx);
    }
    return sum;
}

// Blocked-format walk (common.cuh).  rs holds the pixel's ranks << 8.  All J chains hop block by
// block in lockstep: four unrolled levels, then the exit word.  A finished chain (leaf marker in
// bb) idles on the all-pass-through block 0 until the slowest chain of the group is done.
template <int P, int J, bool COUNT>
__device__ __forceinline__ double traverse_b(const QForest& q, int t_begin, int t_end, int T, double sum,
                                             const uint32_t* __restrict__ rs, const ForestOut& o, long long row,
                                             unsigned& nvisit)
{
    const char* rsb = reinterpret_cast<const char*>(rs);
    for (int t0 = t_begin; t0 < t_end; t0 += J) {
        uint32_t bb[J];
#pragma unroll
        for (int j = 0; j < J; ++j) bb[j] = __ldg(q.broot + min(t0 + j, t_end - 1));
        uint32_t all_done;
        do {
            uint32_t h[J];
            const uint32_t* blk[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                blk[j] = q.blocks + (size_t)((bb[j] & kBLeafFlag) ? 0u : bb[j]) * 32u;
                h[j] = 1u;
            }
#pragma unroll
            for (int lev = 0; lev < 4; ++lev) {
                uint32_t w[J];
#pragma unroll
                for (int j = 0; j < J; ++j) w[j] = __ldg(blk[j] + h[j]);
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const uint32_t xs = *reinterpret_cast<const uint32_t*>(rsb + (~w[j] & 0xFFu) * (P * 4));
                    h[j] = q_step(h[j], h[j], w[j], xs);   // 2h + (xs > ~w)
                    if (COUNT) nvisit += (w[j] != kBPassWord && !(bb[j] & kBLeafFlag)) ? 1u : 0u;
                }
            }
            all_done = kBLeafFlag;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const uint32_t e = __ldg(blk[j] + h[j]);
                bb[j] = (bb[j] & kBLeafFlag) ? bb[j] : e;
                all_done &= bb[j];
            }
        } while (!all_done);
#pragma unroll
        for (int j = 0; j < J; ++j) bb[j] &= ~kBLeafFlag;   // leaf node indices
#pragma unroll
        for (int j = 0; j < J; ++j) {
            if (t0 + j < t_end) {
                sum += __ldg(q.leaf_val + bb[j]);
                if (o.terminal) o.terminal[(long long)(t0 + j) * o.ld_term + row] = o.leaf_rid[bb[j]];
            }
        }
        if (o.leaf_ids) store_leaf_ids<J>(o, row * o.ld_ids + t0, t_end - t0, bb);
    }
    return sum;
}

// Sector-format walk (sector.cuh): the leaf's value, number and quantile key sit in the leaf's own block.
template <int P, int J, bool COUNT>
__device__ __forceinline__ double traverse_sect(const QForest& q, int t_begin, int t_end, int T, double sum,
                                                const uint32_t* __restrict__ rs, const ForestOut& o, long long row,
                                                unsigned& nvisit)
{
    const SectorView v{q.sect, q.sroot, q.sect_block_bytes, q.sect_tile_stride};
    auto sink = [&](int t0, int remaining, const uint32_t (&lb)[J]) {
        if (o.terminal) {
#pragma unroll
            for (int j = 0; j < J; ++j)
                if (j < remaining)
                    o.terminal[(long long)(t0 + j) * o.ld_term + row] = o.leaf_rid[__ldg(q.sect + (size_t)lb[j] * 8u + kSWordLeafNo)];
        }
        if (o.leaf_ids) {
            uint32_t key[J];
#pragma unroll
            for (int j = 0; j < J; ++j) key[j] = __ldg(q.sect + (size_t)lb[j] * 8u + o.ids_word);
            store_leaf_ids<J>(o, row * o.ld_ids + t0, remaining, key);
        }
    };
    return traverse_s<P, J, COUNT>(v, t_begin, t_end, sum, rs, sink, nvisit);
}

// A row with NA/NaN in a column the forest uses (ranger: "Missing data in columns ..." is an error): nothing is
// walked; every output of the row is NA, the quantile pass is told to skip it, and the host reports RFSI_E_NA_INPUT.
__device__ __forceinline__ void forest_na_row(const ForestOut& o, long long p, int t_begin, int t_end, int* na_flag) {
    if (na_flag) *na_flag = 1;
    if (o.mean) o.mean[p] = na_real();
    if (o.row_skip) o.row_skip[p] = 1;
    if (o.terminal)
        for (int t = t_begin; t < t_end; ++t) o.terminal[(long long)t * o.ld_term + p] = kNaInteger;
}

template <int P>
__host__ __device__ constexpr size_t forest_q_smem_bytes(int F) {
    return (size_t)F * P * sizeof(uint32_t);
}

template <int P, int J, bool COUNT>
__global__ void __launch_bounds__(P, (P * 64 * 4 <= 65536) ? 4 : 2)
forest_xq_kernel(QForest q, int t_begin, int t_end, int T, int F, const double* __restrict__ X,
                 long long npix, long long ldx, ForestOut o, unsigned long long* __restrict__ visits,
                 int* __restrict__ na_flag)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* rs = reinterpret_cast<uint32_t*>(smem_raw) + threadIdx.x;
    const long long p = (long long)blockIdx.x * P + threadIdx.x;
    const bool in_range = p < npix;
    bool has_na = false;
    if (in_range) {
        for (int f = 0; f < F; ++f) {
            const double v = X[(long long)f * ldx + p];
            has_na |= (v != v);
            rs[f * P] = rank_of(q, f, v) << q.rank_shift;
        }
    }
    unsigned nvisit = 0;
    if (in_range) {
        if (has_na) {
            forest_na_row(o, p, t_begin, t_end, na_flag);
        } else {
            double sum = (t_begin > 0 && o.mean) ? o.mean[p] : 0.0;
            if (q.sect) sum = traverse_sect<P, J, COUNT>(q, t_begin, t_end, T, sum, rs, o, p, nvisit);
            else if (q.blocks) sum = traverse_b<P, J, COUNT>(q, t_begin, t_end, T, sum, rs, o, p, nvisit);
            else sum = traverse_q<P, J, COUNT>(q, t_begin, t_end, T, sum, rs, o, p, nvisit);
            if (o.mean) o.mean[p] = (t_end == T) ? sum / (double)T : sum;
            if (o.row_skip) o.row_skip[p] = 0;
        }
    }
    if (COUNT && visits) {
        const unsigned long long w = warp_sum_u64(nvisit);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(visits, w);
    }
}

template <int P>
__host__ __device__ constexpr size_t forest_smem_bytes(int F) {
    return (size_t)(F + 1) * P * sizeof(double);  // + the sentinel row of -inf
}

// Unfused form: features come from a column-major matrix X (predict(rf, df)).
// One launch walks trees [t_begin, t_end): the host issues one launch per L2-sized chunk of
// the forest so every CTA on the chip reads the same few tens of MB of nodes at a time; the
// running sum lives in o.mean between launches (still accumulated in tree order).
template <int P, int J, bool COUNT>
__global__ void __launch_bounds__(P, (P * (20 + 6 * J) <= 32768) ? 2 : 1)
forest_x_kernel(const Node* __restrict__ nodes, int t_begin, int t_end, int T, int F,
                const double* __restrict__ X, long long npix, long long ldx, ForestOut o,
                unsigned long long* __restrict__ visits, int* __restrict__ na_flag)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* fs = reinterpret_cast<double*>(smem_raw) + threadIdx.x;
    const long long p = (long long)blockIdx.x * P + threadIdx.x;
    const bool in_range = p < npix;
    bool has_na = false;
    if (in_range) {
        for (int f = 0; f < F; ++f) {
            const double v = X[(long long)f * ldx + p];  // coalesced across the warp
            fs[f * P] = v;
            has_na |= (v != v);
        }
        fs[F * P] = __longlong_as_double(0xFFF0000000000000LL);  // sentinel row: -inf
    }
    // each thread reads only its own column: no barrier needed
    unsigned nvisit = 0;
    if (in_range) {
        if (has_na) {
            forest_na_row(o, p, t_begin, t_end, na_flag);
        } else {
            double sum = (t_begin > 0 && o.mean) ? o.mean[p] : 0.0;
            sum = traverse_all<P, J, COUNT>(nodes, t_begin, t_end, T, F, sum, fs, o, p, nvisit);
            if (o.mean) o.mean[p] = (t_end == T) ? sum / (double)T : sum;
            if (o.row_skip) o.row_skip[p] = 0;
        }
    }
    if (COUNT && visits) {
        const unsigned long long w = warp_sum_u64(nvisit);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(visits, w);
    }
}

}  // namespace rfsi
