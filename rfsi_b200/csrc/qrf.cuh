// qrf.cuh -- quantile-forest epilogue: per pixel, the T leaf samples
// (ranger's random.node.values[terminal node, tree]) -> R quantile(type 7, na.rm=TRUE).
//
// Replaces the R-level `apply(node.values, 1, quantile, quantiles, na.rm=TRUE)` inside
// predict.ranger(type="quantiles") as called at temp_croatia/2_predictions.R:174-176 and
// prcp_catalonia/5_prcp_prediction.R:289-291, and the script's own
// iqr <- (q[,3]-q[,1])/2 (2_predictions.R:177; 5_prcp_prediction.R:292)
// (SURVEY.md 8a rows a5/a6, Appendix A.3).
//
// B200 design: one warp per pixel-day, samples in registers (R = ceil(T/32) per lane).
// Only 2 order statistics per quantile are needed, so instead of sorting (a 512-element
// bitonic network is ~10^4 instructions per row) the warp SELECTS them exactly:
//   1. min/max of the non-NA samples, then one histogram of 256 monotone buckets
//      b(x) = min(255, int((x - min) * 256/(max - min))) in shared memory (built once per row;
//      the maximum is pinned to bucket 255);
//   2. per wanted rank k: a warp scan of the histogram finds the bucket holding the k-th
//      smallest and how many samples precede it; the bucket's few members are compacted
//      into shared memory and ranked against each other (ties by position);
//   3. a bucket with more than 32 members (heavy ties / skew) is re-bucketed between its
//      own min and max until it is small or constant.
// b() is monotone non-decreasing (fp subtract, multiply by a positive scale and truncation
// all are), so bucket order == value order and the selected value IS the order statistic.
// Interpolation uses exactly R's expression (1-h)*x[lo] + h*x[hi].
#pragma once
#include "common.cuh"

namespace rfsi {

constexpr int kQrfWarps = 8;   // rows per CTA

// Monotone bucket of x in [lo, hi].  Halved operands keep hi - lo finite for any finite
// inputs; the maximum always lands in bucket 255 and the minimum in bucket 0, so a bucket
// that is re-bucketed between its own min and max always loses members (termination).
__device__ __forceinline__ int qrf_bucket(double x, double lo_half, double hi, double scale) {
    if (x == hi) return 255;
    const double t = (x * 0.5 - lo_half) * scale;   // x >= lo, so t >= 0; NaN never reaches here
    return (t >= 255.0) ? 255 : (int)t;
}

// Histogram of the members (bit r of `members` set <=> v[r] takes part) between lo and hi;
// returns the bucket of every member in b[], and this lane's 8-bin sum / exclusive prefix.
template <int R>
__device__ __forceinline__ void qrf_hist(const double (&v)[R], unsigned members, double lo, double hi,
                                         int* hist, int lane, unsigned char (&b)[R], int& lane_sum, int& lane_excl)
{
    const double lo_half = lo * 0.5;
    double scale = 256.0 / (hi * 0.5 - lo_half);
    if (!(scale < 1.0e300)) scale = 1.0e300;          // hi - lo tiny: keep the product finite and monotone
    reinterpret_cast<int4*>(hist)[lane * 2] = make_int4(0, 0, 0, 0);
    reinterpret_cast<int4*>(hist)[lane * 2 + 1] = make_int4(0, 0, 0, 0);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r) {
        b[r] = 0;
        if (members & (1u << r)) {
            b[r] = (unsigned char)qrf_bucket(v[r], lo_half, hi, scale);
            atomicAdd(&hist[b[r]], 1);
        }
    }
    __syncwarp();
    const int4 h0 = reinterpret_cast<const int4*>(hist)[lane * 2];
    const int4 h1 = reinterpret_cast<const int4*>(hist)[lane * 2 + 1];
    lane_sum = h0.x + h0.y + h0.z + h0.w + h1.x + h1.y + h1.z + h1.w;
    int incl = lane_sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    lane_excl = incl - lane_sum;
}

// The bucket that holds rank k (0-based among the histogram's members): returns the bucket,
// the number of members before it and its size.
__device__ __forceinline__ void qrf_locate(const int* hist, int lane, int lane_sum, int lane_excl, int k,
                                           int& bucket, int& before, int& count)
{
    const unsigned bal = __ballot_sync(0xffffffffu, lane_excl <= k && k < lane_excl + lane_sum);
    const int src = __ffs(bal) - 1;
    int B = 0, cb = 0, cnt = 0;
    if (lane == src) {
        int c = lane_excl;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int h = hist[lane * 8 + i];
            if (k >= c && k < c + h) { B = lane * 8 + i; cb = c; cnt = h; }
            c += h;
        }
    }
    bucket = __shfl_sync(0xffffffffu, B, src);
    before = __shfl_sync(0xffffffffu, cb, src);
    count = __shfl_sync(0xffffffffu, cnt, src);
}

// k-th smallest (0-based) of the valid samples, given the level-0 histogram.  *next receives the
// (k+1)-th smallest as well when it sits in the same final bucket (the usual case for the
// lo/hi pair of one quantile), and *has_next says so.
template <int R>
__device__ __forceinline__ double qrf_select(const double (&v)[R], unsigned valid, const unsigned char (&b0)[R],
                                             const int* hist0, int sum0, int excl0, int* hist1, double* mem,
                                             int lane, int k, double* next, bool* has_next)
{
    *has_next = false;
    int B, before, cnt;
    qrf_locate(hist0, lane, sum0, excl0, k, B, before, cnt);
    int kk = k - before;
    unsigned mm = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) mm |= ((valid >> r) & 1u & (b0[r] == B ? 1u : 0u)) << r;

    while (cnt > 32) {   // heavy bucket: re-bucket between its own min and max
        double lo = __longlong_as_double(0x7FF0000000000000LL), hi = __longlong_as_double(0xFFF0000000000000LL);
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (mm & (1u << r)) { lo = fmin(lo, v[r]); hi = fmax(hi, v[r]); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lo == hi) {            // all members equal
            if (kk + 1 < cnt) { *next = lo; *has_next = true; }
            return lo;
        }
        unsigned char b1[R];
        int s1, e1;
        qrf_hist<R>(v, mm, lo, hi, hist1, lane, b1, s1, e1);
        qrf_locate(hist1, lane, s1, e1, kk, B, before, cnt);
        kk -= before;
        unsigned m2 = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) m2 |= ((mm >> r) & 1u & (b1[r] == B ? 1u : 0u)) << r;
        mm = m2;
        __syncwarp();
    }

    // compact the bucket's members (<= 32) and rank them against each other
    int pos = 0;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool is = (mm >> r) & 1u;
        const unsigned bal = __ballot_sync(0xffffffffu, is);
        if (is) mem[pos + __popc(bal & lt)] = v[r];
        pos += __popc(bal);
    }
    __syncwarp();
    const double mine = (lane < cnt) ? mem[lane] : 0.0;
    int rk = 0;
    for (int j = 0; j < cnt; ++j) {
        const double other = mem[j];
        rk += (other < mine || (other == mine && j < lane)) ? 1 : 0;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, lane < cnt && rk == kk);
    const double out = __shfl_sync(0xffffffffu, mine, __ffs(hit) - 1);
    if (kk + 1 < cnt) {
        const unsigned hit2 = __ballot_sync(0xffffffffu, lane < cnt && rk == kk + 1);
        *next = __shfl_sync(0xffffffffu, mine, __ffs(hit2) - 1);
        *has_next = true;
    }
    __syncwarp();
    return out;
}

// samples: [nrows][T] pixel-major; skip (nullable): rows whose outputs are NA.
// out_q[j*ld_q + row] for j < nprobs (nullable); out_iqr_i16[row] (nullable) =
// round-half-even((q[nprobs-1]-q[0])/2*100), NA -> -32767.
template <int R>
__global__ void __launch_bounds__(32 * kQrfWarps)
qrf_quantile_kernel(const double* __restrict__ samples, const unsigned char* __restrict__ skip,
                    long long nrows, int T, const double* __restrict__ probs, int nprobs,
                    double* __restrict__ out_q, long long ld_q, int16_t* __restrict__ out_iqr_i16)
{
    __shared__ __align__(16) int s_hist0[kQrfWarps][256];
    __shared__ __align__(16) int s_hist1[kQrfWarps][256];
    __shared__ double s_mem[kQrfWarps][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.x * kQrfWarps + wib;
    if (row >= nrows) return;   // whole warp
    int* hist0 = s_hist0[wib];
    int* hist1 = s_hist1[wib];
    double* mem = s_mem[wib];

    double v[R];
    unsigned valid = 0;
    const bool skipped = skip && skip[row];
    const double* src = samples + row * T;
    double mn = __longlong_as_double(0x7FF0000000000000LL), mx = __longlong_as_double(0xFFF0000000000000LL);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int t = r * 32 + lane;   // coalesced
        double x = 0.0;
        if (!skipped && t < T) {
            x = src[t];
            if (x == x) { valid |= 1u << r; mn = fmin(mn, x); mx = fmax(mx, x); }   // na.rm = TRUE
        }
        v[r] = x;
    }
    int m = __popc(valid);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m += __shfl_xor_sync(0xffffffffu, m, o);
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }

    unsigned char b0[R];
    int sum0 = 0, excl0 = 0;
    if (m > 0 && mn < mx) qrf_hist<R>(v, valid, mn, mx, hist0, lane, b0, sum0, excl0);

    double qfirst = 0.0, qlast = 0.0;
    for (int j = 0; j < nprobs; ++j) {
        double q;
        if (m == 0) {
            q = na_real();
        } else if (!(mn < mx)) {
            q = mn;   // every sample equal
        } else {
            const double index = __dadd_rn(1.0, __dmul_rn((double)(m - 1), probs[j]));   // no FMA: R computes 1 + (n-1)*p in two roundings
            const double lo = floor(index), hi = ceil(index);
            double nxt = 0.0, unused = 0.0;
            bool has_nxt = false, unused_b = false;
            const double xlo = qrf_select<R>(v, valid, b0, hist0, sum0, excl0, hist1, mem, lane, (int)lo - 1, &nxt, &has_nxt);
            const double xhi = (hi == lo) ? xlo
                               : (has_nxt ? nxt
                                          : qrf_select<R>(v, valid, b0, hist0, sum0, excl0, hist1, mem, lane, (int)hi - 1,
                                                          &unused, &unused_b));
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
