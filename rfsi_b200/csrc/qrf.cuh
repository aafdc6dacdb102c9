// qrf.cuh -- quantile-forest epilogue: per pixel, the T leaf samples
// (ranger's random.node.values[terminal node, tree]) -> R quantile(type 7, na.rm=TRUE).
//
// Replaces the R-level `apply(node.values, 1, quantile, quantiles, na.rm=TRUE)` inside
// predict.ranger(type="quantiles") as called at temp_croatia/2_predictions.R:174-176 and
// prcp_catalonia/5_prcp_prediction.R:289-291, and the script's own
// iqr <- (q[,3]-q[,1])/2 (2_predictions.R:177; 5_prcp_prediction.R:292)
// (SURVEY.md 8a rows a5/a6, Appendix A.3).
//
// B200 design: one warp per pixel.  The T samples are read coalesced (pixel-major
// scratch written by the traversal kernel), R = ceil(T/32) per lane in registers,
// sorted by a bitonic network (strides < R in registers, strides >= R by shuffle),
// then the needed order statistics are fetched by shuffle and interpolated with
// exactly R's expression (1-h)*x[lo] + h*x[hi].
#pragma once
#include "common.cuh"

namespace rfsi {

template <int R>
__device__ __forceinline__ void bitonic_sort_warp(double (&v)[R], int lane) {
    constexpr int N = 32 * R;
#pragma unroll
    for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= R) {
                const int lstride = stride / R;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int i = lane * R + r;
                    const double other = __shfl_xor_sync(0xffffffffu, v[r], lstride);
                    const bool up = (i & size) == 0;
                    const bool lower = (i & stride) == 0;
                    const double mn = fmin(v[r], other), mx = fmax(v[r], other);
                    v[r] = (up == lower) ? mn : mx;
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((r & stride) == 0) {
                        const int i = lane * R + r;
                        const bool up = (i & size) == 0;
                        const double a = v[r], b = v[r + stride];
                        const double mn = fmin(a, b), mx = fmax(a, b);
                        v[r] = up ? mn : mx;
                        v[r + stride] = up ? mx : mn;
                    }
                }
            }
        }
    }
}

template <int R>
__device__ __forceinline__ double fetch_sorted(const double (&v)[R], int e) {
    // element e of the sorted sequence lives in lane e / R, register e % R
    const int r = e % R;
    double mine = v[0];
#pragma unroll
    for (int q = 1; q < R; ++q) mine = (r == q) ? v[q] : mine;
    return __shfl_sync(0xffffffffu, mine, e / R);
}

// samples: [nrows][T] pixel-major; skip (nullable): rows whose outputs are NA.
// out_q[j*ld_q + row] for j < nprobs (nullable); out_iqr_i16[row] (nullable) =
// round-half-even((q[nprobs-1]-q[0])/2*100), NA -> -32767.
template <int R>
__global__ void __launch_bounds__(256)
qrf_quantile_kernel(const double* __restrict__ samples, const unsigned char* __restrict__ skip,
                    long long nrows, int T, const double* __restrict__ probs, int nprobs,
                    double* __restrict__ out_q, long long ld_q, int16_t* __restrict__ out_iqr_i16)
{
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const double inf = __longlong_as_double(0x7FF0000000000000LL);
    double v[R];
    int m = 0;
    const bool skipped = skip && skip[row];
    const double* src = samples + row * T;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        // sorting-network layout wants element i = lane*R + r; the load is coalesced
        // over i' = r*32 + lane instead -- any assignment works, it is a multiset
        const int t = r * 32 + lane;
        double x = inf;
        if (!skipped && t < T) {
            x = src[t];
            if (x != x) x = inf; else ++m;  // na.rm = TRUE
        }
        v[r] = x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
    bitonic_sort_warp<R>(v, lane);

    double qfirst = 0.0, qlast = 0.0;
    for (int j = 0; j < nprobs; ++j) {
        double q;
        if (m == 0) {
            q = na_real();
        } else {
            const double index = 1.0 + (double)(m - 1) * probs[j];
            const double lo = floor(index), hi = ceil(index);
            const double xlo = fetch_sorted<R>(v, (int)lo - 1);
            const double xhi = fetch_sorted<R>(v, (int)hi - 1);
            if (index > lo && xhi != xlo) {
                const double h = index - lo;
                q = __dadd_rn(__dmul_rn(1.0 - h, xlo), __dmul_rn(h, xhi));
            } else {
                q = xlo;
            }
        }
        if (lane == 0 && out_q) out_q[(long long)j * ld_q + row] = q;
        if (j == 0) qfirst = q;
        qlast = q;
    }
    if (lane == 0 && out_iqr_i16) {
        const double c = __dmul_rn(__dmul_rn(__dsub_rn(qlast, qfirst), 0.5), 100.0);
        int16_t r16 = -32767;
        if (c == c) {
            double rr = rint(c);
            rr = fmin(fmax(rr, -32767.0), 32767.0);
            r16 = (int16_t)rr;
        }
        out_iqr_i16[row] = r16;
    }
}

}  // namespace rfsi
