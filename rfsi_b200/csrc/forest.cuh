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

namespace rfsi {

struct ForestOut {
    double* mean;             // [npix] (nullable)
    double* samples;          // [npix][T] pixel-major leaf samples for the quantile pass (nullable)
    const double* leaf_sample;  // per device-node sample (random.node.values), needed if samples
    int32_t* terminal;        // [T][ld_term] tree-local terminal node ids (nullable)
    long long ld_term;
};

// Walk all T trees for one pixel whose features are fs[f*P].  Returns the tree-order
// sum of leaf predictions.  row = this pixel's output row.
template <int P, int J, bool COUNT>
__device__ __forceinline__ double traverse_all(const Node* __restrict__ nodes, int T,
                                               const double* __restrict__ fs, const ForestOut& o,
                                               long long row, unsigned& nvisit)
{
    double sum = 0.0;
    for (int t0 = 0; t0 < T; t0 += J) {
        int idx[J];
        Node nd[J];
        unsigned live = 0;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            idx[j] = t0 + j;  // the root of tree t is node t
            nd[j].val = 0.0; nd[j].child = 0; nd[j].feat = -1;
            if (t0 + j < T) live |= 1u << j;
        }
        while (live) {
#pragma unroll
            for (int j = 0; j < J; ++j)
                if (live & (1u << j)) nd[j] = load_node(nodes + idx[j]);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (live & (1u << j)) {
                    if (nd[j].feat >= 0) {
                        const double x = fs[nd[j].feat * P];
                        idx[j] = nd[j].child + ((x <= nd[j].val) ? 0 : 1);
                        if (COUNT) ++nvisit;
                    } else {
                        live &= ~(1u << j);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < J; ++j) {
            if (t0 + j < T) {
                sum += nd[j].val;
                if (o.samples) o.samples[row * T + t0 + j] = o.leaf_sample[idx[j]];
                if (o.terminal) o.terminal[(long long)(t0 + j) * o.ld_term + row] = nd[j].child;
            }
        }
    }
    return sum;
}

template <int P>
__host__ __device__ constexpr size_t forest_smem_bytes(int F) {
    return (size_t)F * P * sizeof(double);
}

// Unfused form: features come from a column-major matrix X (predict(rf, df)).
template <int P, int J, bool COUNT>
__global__ void __launch_bounds__(P)
forest_x_kernel(const Node* __restrict__ nodes, int T, int F, const double* __restrict__ X,
                long long npix, long long ldx, ForestOut o, unsigned long long* __restrict__ visits,
                int* __restrict__ na_flag)
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
    }
    // each thread reads only its own column: no barrier needed
    unsigned nvisit = 0;
    if (in_range) {
        if (has_na) {
            if (na_flag) *na_flag = 1;  // ranger: "Missing data in columns" -> error upstream
            if (o.mean) o.mean[p] = na_real();
        } else {
            const double sum = traverse_all<P, J, COUNT>(nodes, T, fs, o, p, nvisit);
            if (o.mean) o.mean[p] = sum / (double)T;
        }
    }
    if (COUNT && visits) {
        const unsigned long long w = warp_sum_u64(nvisit);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(visits, w);
    }
}

}  // namespace rfsi
