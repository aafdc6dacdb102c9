// qrf.cuh -- quantile-forest epilogue: per pixel, the T leaf samples
// (ranger's random.node.values[terminal node, tree]) -> R quantile(type 7, na.rm=TRUE).
//
// Replaces the R-level `apply(node.values, 1, quantile, quantiles, na.rm=TRUE)` inside
// predict.ranger(type="quantiles") as called at temp_croatia/2_predictions.R:174-176 and
// prcp_catalonia/5_prcp_prediction.R:289-291, and the script's own
// iqr <- (q[,3]-q[,1])/2 (2_predictions.R:177; 5_prcp_prediction.R:292)
// (SURVEY.md 8a rows a5/a6, Appendix A.3).
//
// B200 design: one warp per pixel-day, R = ceil(T/32) samples per lane.  Only 2 order
// statistics per quantile are needed, so instead of sorting (a 512-element bitonic network is
// ~10^4 instructions per row) the warp SELECTS them exactly:
//   1. min/max of the non-NA samples, then one histogram of 256 monotone buckets
//      b(x) = min(255, int((x - min) * 256/(max - min))) in shared memory (built once per row;
//      the maximum is pinned to bucket 255);
//   2. per wanted rank k: a warp scan of the histogram finds the bucket holding the k-th
//      smallest and how many samples precede it; the bucket's few members are compacted
//      into shared memory and ranked against each other (ties by slot);
//   3. a bucket with more than 32 members (heavy ties / skew) is re-bucketed between its
//      own min and max until it is small or constant.
// b() is monotone non-decreasing (fp subtract, multiply by a positive scale and truncation
// all are), so bucket order == value order and the selected value IS the order statistic.
// Interpolation uses exactly R's expression (1-h)*x[lo] + h*x[hi].
//
// Memory pipeline: the forest walk left 4-byte leaf indices per (row, tree).  Each CTA owns a
// contiguous run of rows and its 8 warps walk it 8 adjacent rows at a time: adjacent pixels end
// in the same leaves, so most of the T gathers leaf_sample[id] of a row hit lines its
// neighbours just brought into L1 (a grid-strided assignment measured 4.6 % L1 hits and 7x the
// DRAM traffic).  While a warp selects in row i, the gathers of row i+8 are in flight into
// registers and the ids of row i+16 into shared memory (cp.async), so neither the DRAM read of
// the id scratch nor the gathers sit on the critical path.
#pragma once
#include "common.cuh"

namespace rfsi {

constexpr int kQrfWarps = 8;        // rows in flight per CTA
// resident CTAs per SM the register budget is cut for: two (16 warps, 2 x 66 KB of shared
// memory at T = 512, which leaves ~96 KB of L1 for the gathers); one for 1024-tree rows
#ifndef QRF_MIN_BLOCKS
#define QRF_MIN_BLOCKS 2
#endif
template <int R> constexpr int qrf_min_blocks() { return R >= 32 ? 1 : QRF_MIN_BLOCKS; }

// The kernel exists for two key types K.  double: the leaf table holds the samples themselves
// (8 B per leaf).  uint32_t: the table holds each sample's RANK in the forest-wide sorted
// dictionary of distinct sample values (4 B per leaf; the dictionary -- at most one entry per
// training row -- stays in L2); ranks order exactly like the values, so the selection runs on
// integers and only the two order statistics of a quantile are looked up as fp64.  Half the
// bytes per gather and twice the leaves per 32-byte sector for neighbouring pixels to share.

// shared memory of one warp, in bytes: samples of the current row, two histograms, the
// compaction buffer, the id buffer of the row after next
template <int R, class K>
__host__ __device__ constexpr size_t qrf_warp_smem() {
    return (size_t)R * 32 * sizeof(K) + 256 * 4 + 256 * 4 + 32 * sizeof(K) + (size_t)R * 32 * 4;
}
template <int R, class K>
__host__ __device__ constexpr size_t qrf_smem_bytes() { return kQrfWarps * qrf_warp_smem<R, K>(); }

template <class K>
struct QrfWarp {
    K* sv;           // [R*32] samples of the row, element r of lane l at r*32 + l
    int* hist0;      // [256]
    int* hist1;      // [256]
    K* mem;          // [32]
    uint32_t* ids;   // [R*32]
};

// Warp-wide min and max of non-NaN doubles without 64-bit shuffles: the IEEE bit pattern maps
// to an unsigned key of the same order, whose high word is reduced first (one REDUX) and the
// low word among the lanes that hold the winning high word second.
__device__ __forceinline__ unsigned long long qrf_key(double x) {
    const long long b = __double_as_longlong(x);
    return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000LL));
}
__device__ __forceinline__ double qrf_unkey(unsigned long long k) {
    const long long b = (long long)k;
    return __longlong_as_double(b ^ ((~b >> 63) | (long long)0x8000000000000000LL));
}
__device__ __forceinline__ double qrf_warp_min(double x) {
    const unsigned long long k = qrf_key(x);
    const unsigned hi = __reduce_min_sync(0xffffffffu, (unsigned)(k >> 32));
    const unsigned lo = __reduce_min_sync(0xffffffffu, (unsigned)(k >> 32) == hi ? (unsigned)k : 0xffffffffu);
    return qrf_unkey(((unsigned long long)hi << 32) | lo);
}
__device__ __forceinline__ double qrf_warp_max(double x) {
    const unsigned long long k = qrf_key(x);
    const unsigned hi = __reduce_max_sync(0xffffffffu, (unsigned)(k >> 32));
    const unsigned lo = __reduce_max_sync(0xffffffffu, (unsigned)(k >> 32) == hi ? (unsigned)k : 0u);
    return qrf_unkey(((unsigned long long)hi << 32) | lo);
}

// Key traits.  bucket() is monotone non-decreasing in x on [lo, hi]; the maximum always lands in
// bucket 255 and the minimum in bucket 0, so a bucket that is re-bucketed between its own min and
// max always loses members (termination).
template <class K> struct QrfKey;

template <> struct QrfKey<double> {
    struct Scale { double lo_half, scale; };
    static __device__ __forceinline__ double na() { return na_real(); }
    static __device__ __forceinline__ bool ok(double x) { return x == x; }                      // na.rm = TRUE
    static __device__ __forceinline__ double top() { return __longlong_as_double(0x7FF0000000000000LL); }
    static __device__ __forceinline__ double bottom() { return __longlong_as_double(0xFFF0000000000000LL); }
    static __device__ __forceinline__ double warp_min(double x) { return qrf_warp_min(x); }
    static __device__ __forceinline__ double warp_max(double x) { return qrf_warp_max(x); }
    // halved operands keep hi - lo finite for any finite inputs
    static __device__ __forceinline__ Scale make_scale(double lo, double hi) {
        Scale s{lo * 0.5, 0.0};
        s.scale = 256.0 / (hi * 0.5 - s.lo_half);
        if (!(s.scale < 1.0e300)) s.scale = 1.0e300;   // hi - lo tiny: keep the product finite and monotone
        return s;
    }
    static __device__ __forceinline__ int bucket(double x, double hi, const Scale& s) {
        const double t = (x * 0.5 - s.lo_half) * s.scale;   // x >= lo, so t >= 0; NaN never reaches here
        const int b = (t >= 255.0) ? 255 : (int)t;
        return (x == hi) ? 255 : b;
    }
    static __device__ __forceinline__ double value(double k, const double*) { return k; }
};

template <> struct QrfKey<uint32_t> {
    struct Scale { uint32_t lo; float scale; };
    static __device__ __forceinline__ uint32_t na() { return 0xFFFFFFFFu; }
    static __device__ __forceinline__ bool ok(uint32_t x) { return x != 0xFFFFFFFFu; }
    static __device__ __forceinline__ uint32_t top() { return 0xFFFFFFFFu; }
    static __device__ __forceinline__ uint32_t bottom() { return 0u; }
    static __device__ __forceinline__ uint32_t warp_min(uint32_t x) { return __reduce_min_sync(0xffffffffu, x); }
    static __device__ __forceinline__ uint32_t warp_max(uint32_t x) { return __reduce_max_sync(0xffffffffu, x); }
    static __device__ __forceinline__ Scale make_scale(uint32_t lo, uint32_t hi) {
        return Scale{lo, 256.0f / (float)(hi - lo)};     // only called with lo < hi
    }
    // integer -> float conversion, multiplication by a positive constant and truncation are all monotone
    static __device__ __forceinline__ int bucket(uint32_t x, uint32_t hi, const Scale& s) {
        const float t = (float)(x - s.lo) * s.scale;
        const int b = (t >= 255.0f) ? 255 : (int)t;
        return (x == hi) ? 255 : b;
    }
    static __device__ __forceinline__ double value(uint32_t k, const double* dict) { return __ldg(dict + k); }
};

// Histogram of the members (bit r of `members` set <=> sample r of this lane takes part)
// between lo and hi; bp packs the bucket of element r into byte r&3 of word r>>2; returns this
// lane's 8-bin sum and exclusive prefix.
template <int R, class K>
__device__ __forceinline__ void qrf_hist(const K* sv, unsigned members, K lo, K hi, int* hist, int lane,
                                         uint32_t (&bp)[(R + 3) / 4], int& lane_sum, int& lane_excl)
{
    const typename QrfKey<K>::Scale sc = QrfKey<K>::make_scale(lo, hi);
    reinterpret_cast<int4*>(hist)[lane * 2] = make_int4(0, 0, 0, 0);
    reinterpret_cast<int4*>(hist)[lane * 2 + 1] = make_int4(0, 0, 0, 0);
    __syncwarp();
#pragma unroll
    for (int w = 0; w < (R + 3) / 4; ++w) bp[w] = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool is = (members >> r) & 1u;
        const K x = is ? sv[r * 32 + lane] : lo;
        const int b = QrfKey<K>::bucket(x, hi, sc);
        if (is) atomicAdd(&hist[b], 1);
        bp[r >> 2] |= (uint32_t)b << (8 * (r & 3));
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
// the number of members before it and its size.  Every lane scans its own 8 bins without
// branching; the owning lane's answer travels in one packed shuffle (counts <= 1024).
__device__ __forceinline__ void qrf_locate(const int* hist, int lane, int lane_sum, int lane_excl, int k,
                                           int& bucket, int& before, int& count)
{
    const unsigned bal = __ballot_sync(0xffffffffu, lane_excl <= k && k < lane_excl + lane_sum);
    const int src = __ffs(bal) - 1;
    const int4 h0 = reinterpret_cast<const int4*>(hist)[lane * 2];
    const int4 h1 = reinterpret_cast<const int4*>(hist)[lane * 2 + 1];
    const int h[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
    uint32_t packed = 0;
    int c = lane_excl;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool in = (k >= c) && (k < c + h[i]);
        packed = in ? ((uint32_t)(lane * 8 + i) | ((uint32_t)c << 8) | ((uint32_t)h[i] << 19)) : packed;
        c += h[i];
    }
    packed = __shfl_sync(0xffffffffu, packed, src);
    bucket = (int)(packed & 0xFFu);
    before = (int)((packed >> 8) & 0x7FFu);
    count = (int)(packed >> 19);
}

// bit r set <=> byte r of the packed buckets equals B
template <int R>
__device__ __forceinline__ unsigned qrf_match(const uint32_t (&bp)[(R + 3) / 4], int B) {
    const uint32_t Bw = (uint32_t)B * 0x01010101u;
    unsigned m = 0;
#pragma unroll
    for (int w = 0; w < (R + 3) / 4; ++w) {
        const uint32_t eq = __vcmpeq4(bp[w], Bw) & 0x01010101u;   // 1 in the low bit of every equal byte
        m |= ((eq * 0x01020408u) >> 24 & 0xFu) << (4 * w);        // gather the four bits (no carries meet)
    }
    return m;
}

// k-th smallest (0-based) of the valid samples, given the level-0 histogram.  *next receives the
// (k+1)-th smallest as well when it sits in the same final bucket (the usual case for the
// lo/hi pair of one quantile), and *has_next says so.
template <int R, class K>
__device__ __forceinline__ K qrf_select(const QrfWarp<K>& s, unsigned valid, const uint32_t (&b0)[(R + 3) / 4],
                                        int sum0, int excl0, int lane, int k, K* next, bool* has_next)
{
    *has_next = false;
    int B, before, cnt;
    qrf_locate(s.hist0, lane, sum0, excl0, k, B, before, cnt);
    int kk = k - before;
    unsigned mm = qrf_match<R>(b0, B) & valid;

    while (cnt > 32) {   // heavy bucket: re-bucket between its own min and max
        K lo = QrfKey<K>::top(), hi = QrfKey<K>::bottom();
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (mm & (1u << r)) { const K x = s.sv[r * 32 + lane]; lo = x < lo ? x : lo; hi = x > hi ? x : hi; }
        lo = QrfKey<K>::warp_min(lo);
        hi = QrfKey<K>::warp_max(hi);
        if (lo == hi) {            // all members equal
            if (kk + 1 < cnt) { *next = lo; *has_next = true; }
            return lo;
        }
        uint32_t b1[(R + 3) / 4];
        int s1, e1;
        qrf_hist<R, K>(s.sv, mm, lo, hi, s.hist1, lane, b1, s1, e1);
        qrf_locate(s.hist1, lane, s1, e1, kk, B, before, cnt);
        kk -= before;
        mm &= qrf_match<R>(b1, B);
        __syncwarp();
    }

    // compact the bucket's members (<= 32) and rank them against each other (ties by slot).
    int pos = 0;
    const unsigned lt = (1u << lane) - 1u;
    unsigned any = __reduce_or_sync(0xffffffffu, mm);   // the few r at which some lane has a member
    while (any) {
        const int r = __ffs(any) - 1;
        any &= any - 1;
        const bool is = (mm >> r) & 1u;
        const unsigned bal = __ballot_sync(0xffffffffu, is);
        if (is) s.mem[pos + __popc(bal & lt)] = s.sv[r * 32 + lane];
        pos += __popc(bal);
    }
    __syncwarp();
    const K mine = (lane < cnt) ? s.mem[lane] : K(0);
    int rk = 0;
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
        const K other = s.mem[j];
        rk += (other < mine || (other == mine && j < lane)) ? 1 : 0;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, lane < cnt && rk == kk);
    const K out = __shfl_sync(0xffffffffu, mine, __ffs(hit) - 1);
    if (kk + 1 < cnt) {
        const unsigned hit2 = __ballot_sync(0xffffffffu, lane < cnt && rk == kk + 1);
        *next = __shfl_sync(0xffffffffu, mine, __ffs(hit2) - 1);
        *has_next = true;
    }
    __syncwarp();
    return out;
}

// One row, given its samples in s.sv and this lane's validity mask / min / max: the requested
// quantiles (R type 7) and the optional Int16 IQR product.  Whole warp.
template <int R, class K>
__device__ __forceinline__ void qrf_row_finish(const QrfWarp<K>& s, unsigned valid, K mn, K mx, int lane, long long row,
                                               const double* __restrict__ probs, int nprobs, const double* __restrict__ dict,
                                               double* __restrict__ out_q, long long ld_q, int16_t* __restrict__ out_iqr_i16)
{
    using KT = QrfKey<K>;
    const double nan = na_real();
    const int m = (int)__reduce_add_sync(0xffffffffu, (unsigned)__popc(valid));
    mn = KT::warp_min(mn);   // top / bottom when the row has no sample
    mx = KT::warp_max(mx);

    uint32_t b0[(R + 3) / 4];
    int sum0 = 0, excl0 = 0;
    if (m > 0 && mn < mx) qrf_hist<R, K>(s.sv, valid, mn, mx, s.hist0, lane, b0, sum0, excl0);

    double qfirst = 0.0, qlast = 0.0;
    for (int j = 0; j < nprobs; ++j) {
        double q;
        if (m == 0) {
            q = nan;
        } else if (!(mn < mx)) {
            q = KT::value(mn, dict);   // every sample equal
        } else {
            const double index = __dadd_rn(1.0, __dmul_rn((double)(m - 1), probs[j]));   // no FMA: R computes 1 + (n-1)*p in two roundings
            const double lo = floor(index), hi = ceil(index);
            K nxt = K(0), unused = K(0);
            bool has_nxt = false, unused_b = false;
            const K klo = qrf_select<R, K>(s, valid, b0, sum0, excl0, lane, (int)lo - 1, &nxt, &has_nxt);
            const K khi = (hi == lo) ? klo
                          : (has_nxt ? nxt
                                     : qrf_select<R, K>(s, valid, b0, sum0, excl0, lane, (int)hi - 1, &unused, &unused_b));
            const double xlo = KT::value(klo, dict);
            const double xhi = (khi == klo) ? xlo : KT::value(khi, dict);
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

__device__ __forceinline__ void qrf_cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}

// leaf_ids: [nrows][ld_ids] pixel-major leaf index per tree (written by the forest walk);
// leaf_key: the forest's per-leaf table (samples, or their dictionary ranks with NA = 0xFFFFFFFF);
// (When the walk can store the keys themselves -- sector image -- qrf_direct_kernel below is used instead.)
// dict: the sorted distinct sample values (rank form only); skip (nullable): rows whose outputs
// are NA (their ids were never written).
// out_q[j*ld_q + row] for j < nprobs (nullable); out_iqr_i16[row] (nullable) =
// round-half-even((q[nprobs-1]-q[0])/2*100), NA -> -32767.
template <int R, class K>
__global__ void __launch_bounds__(32 * kQrfWarps, qrf_min_blocks<R>())
qrf_quantile_kernel(const uint32_t* __restrict__ leaf_ids, long long ld_ids, const K* __restrict__ leaf_key,
                    const double* __restrict__ dict, const unsigned char* __restrict__ skip, long long nrows, int T,
                    const double* __restrict__ probs, int nprobs,
                    double* __restrict__ out_q, long long ld_q, int16_t* __restrict__ out_iqr_i16)
{
    extern __shared__ __align__(16) unsigned char qrf_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    using KT = QrfKey<K>;
    QrfWarp<K> s;
    {
        unsigned char* base = qrf_smem + (size_t)wib * qrf_warp_smem<R, K>();
        s.sv = reinterpret_cast<K*>(base);                    base += (size_t)R * 32 * sizeof(K);
        s.hist0 = reinterpret_cast<int*>(base);               base += 256 * 4;
        s.hist1 = reinterpret_cast<int*>(base);               base += 256 * 4;
        s.mem = reinterpret_cast<K*>(base);                   base += 32 * sizeof(K);
        s.ids = reinterpret_cast<uint32_t*>(base);
    }
    // this CTA's contiguous run of rows; its warps take them kQrfWarps at a time
    const long long per_cta = ((nrows + gridDim.x - 1) / gridDim.x + kQrfWarps - 1) / kQrfWarps * kQrfWarps;
    const long long row_end = min(nrows, (long long)(blockIdx.x + 1) * per_cta);
    const long long stride = kQrfWarps;
    long long row = (long long)blockIdx.x * per_cta + wib;
    if (row >= row_end) return;   // whole warp
    const int nchunk = (int)(ld_ids >> 2);   // 16-byte pieces of one id row
    const double nan = na_real();

    auto fetch_ids = [&](long long rw) {      // ids of row rw -> s.ids (asynchronous)
        if (rw < row_end) {
            const uint32_t* src = leaf_ids + rw * ld_ids;
            for (int c = lane; c < nchunk; c += 32) qrf_cp_async16(s.ids + 4 * c, src + 4 * c);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto ids_ready = [&]() {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
    };
    K vn[R];                                   // samples of the row about to be processed
    auto gather = [&](long long rw, bool skipped) {   // s.ids -> vn (loads stay in flight)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int t = r * 32 + lane;
            vn[r] = (rw < row_end && !skipped && t < T) ? __ldg(leaf_key + s.ids[t]) : KT::na();
        }
        __syncwarp();                          // every lane has read s.ids: it may be refilled
    };

    // prologue: ids(row) -> gathers(row) in flight, ids(row + stride) in flight
    bool skip_next = (skip && row + stride < row_end) ? skip[row + stride] != 0 : false;
    fetch_ids(row);
    ids_ready();
    gather(row, skip && skip[row]);
    fetch_ids(row + stride);

    for (; row < row_end; row += stride) {
        // ---- take this row's samples out of the registers ------------------------------
        unsigned valid = 0;
        K mn = KT::top(), mx = KT::bottom();
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const K x = vn[r];
            s.sv[r * 32 + lane] = x;
            if (KT::ok(x)) { valid |= 1u << r; mn = x < mn ? x : mn; mx = x > mx ? x : mx; }   // na.rm = TRUE
        }
        // ---- next row's gathers, the row after's ids ---------------------------------------
        ids_ready();                           // also publishes s.sv to the warp
        gather(row + stride, skip_next);
        fetch_ids(row + 2 * stride);
        skip_next = (skip && row + 2 * stride < row_end) ? skip[row + 2 * stride] != 0 : false;

        qrf_row_finish<R, K>(s, valid, mn, mx, lane, row, probs, nprobs, dict, out_q, ld_q, out_iqr_i16);
        __syncwarp();   // the row's shared memory is free again
    }
}


// ---- one row of integer keys: a counting sort by bucket, then a handful of keys per quantile (direct form) -----------
// Same selection idea as qrf_row_finish (monotone buckets between the row's min and max, exact ranking inside the few
// buckets that matter), restated for 4-byte keys so that a row costs well under half the instructions:
//   * NA keys (0xFFFFFFFF) are simply the largest keys: ranks below the number of valid samples never reach them, so
//     there is no validity mask; the row's tail beyond T is filled with NA by the caller;
//   * bucket(x) = min(255, umulhi(x - min, mult)), mult <= (2^40 - 1) / (max - min + 1): two integer instructions, monotone;
//   * one histogram (shared-memory atomics), one warp scan; the table becomes tab[b] = cursor | first rank << 16 and a second
//     round of atomics PLACES every key at srt[cursor++] -- the row is rewritten in place, grouped by bucket, buckets in
//     rank order (a counting sort; order inside a bucket is arbitrary).  Afterwards tab[b] = end rank | first rank << 16;
//   * rank k then sits somewhere in the bucket of srt[k]: one broadcast load of that key, its bucket, its table entry give
//     the bucket's rank interval [first, end); when k is its last rank, the bucket of srt[end] is appended (rank k+1 lives
//     there).  That contiguous WINDOW srt[first .. end2) -- a handful of keys -- goes one key per lane, and the member of
//     rank k is found by PEELING the window's distinct values from below with warp reductions (REDUX.MIN, a tie group per
//     step): no per-element marking, no compaction;
//   * a window with more than 32 members (ties: many trees returning the same training sample, zero-inflated data) is
//     peeled the same way with up to R keys per lane, at most 8 distinct values;
//   * lanes 0..2 finish three quantiles side by side (R's (1-h)*x[lo] + h*x[hi], two dictionary look-ups each).
// Returns false -- nothing written -- only when a heavy window does not resolve within 8 distinct values: the caller then
// runs qrf_row_finish on the (permuted) row, which re-buckets.  profiles/r02_q_*.
constexpr int kQrfGroup = 3;   // quantiles finished side by side

template <int R>
__device__ __forceinline__ bool qrf_row_fast(uint32_t* sv, int* tab, int lane, long long row,
                                             const double* __restrict__ probs, int nprobs, const double* __restrict__ dict,
                                             double* __restrict__ out_q, long long ld_q, int16_t* __restrict__ out_iqr_i16)
{
    static_assert(R % 2 == 0 || R == 1, "rows are read two elements at a time");
    constexpr uint32_t NA = 0xFFFFFFFFu;
    const double nan = na_real();
    uint32_t x[R];
    uint32_t mn = NA, mxp = 0u;
    unsigned nval = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) x[r] = sv[r * 32 + lane];
#pragma unroll
    for (int r = 0; r + 1 < R; r += 2) {
        const uint32_t y0 = x[r] + 1u, y1 = x[r + 1] + 1u;       // NA + 1 wraps to 0: the maximum of the valid keys, plus one
        mn = __vimin3_u32(mn, x[r], x[r + 1]);
        mxp = __vimax3_u32(mxp, y0, y1);
        nval += min(y0, 1u) + min(y1, 1u);
    }
    if (R & 1) {
        const uint32_t y0 = x[R - 1] + 1u;
        mn = min(mn, x[R - 1]); mxp = max(mxp, y0); nval += min(y0, 1u);
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    mxp = __reduce_max_sync(0xffffffffu, mxp);
    const int m = (int)__reduce_add_sync(0xffffffffu, nval);                          // na.rm = TRUE
    const uint32_t mx = mxp - 1u;

    double qfirst = nan, qlast = nan;
    if (m == 0 || mn == mx) {                       // no sample at all, or every sample equal
        const double q = (m == 0) ? nan : __ldg(dict + mn);
        if (out_q) for (int j = lane; j < nprobs; j += 32) out_q[(long long)j * ld_q + row] = q;
        qfirst = qlast = q;
    } else {
        // ---- histogram of 256 monotone buckets, scanned into first ranks ---------------------------
        const float fr = __uint2float_ru(mx - mn + 1u);                              // >= 2
        const uint32_t mult = __float2uint_rz(__fdividef(1099507433472.0f, fr));     // 2^40 (1 - 2^-18) / (range + 1), saturating
        auto bucket_of = [&](uint32_t key) { return min(__umulhi(key - mn, mult), 255u); };   // the clamp only ever acts on NA keys
        reinterpret_cast<int4*>(tab)[lane * 2] = make_int4(0, 0, 0, 0);
        reinterpret_cast<int4*>(tab)[lane * 2 + 1] = make_int4(0, 0, 0, 0);
        __syncwarp();
        int* slot[R];                                                                 // &tab[bucket of element r]
#pragma unroll
        for (int r = 0; r < R; ++r) {
            slot[r] = tab + bucket_of(x[r]);
            atomicAdd(slot[r], 1);
        }
        __syncwarp();
        {
            const int4 h0 = reinterpret_cast<const int4*>(tab)[lane * 2];
            const int4 h1 = reinterpret_cast<const int4*>(tab)[lane * 2 + 1];
            const int s1 = h0.x, s2 = s1 + h0.y, s3 = s2 + h0.z, s4 = s3 + h0.w;
            const int s5 = s4 + h1.x, s6 = s5 + h1.y, s7 = s6 + h1.z, s8 = s7 + h1.w;
            int incl = s8;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int e = incl - s8;                                                  // elements in the buckets of lower lanes
            // cursor (low half, starts at the bucket's first rank) | first rank << 16
            const int4 p0 = make_int4(e * 0x10001, (e + s1) * 0x10001, (e + s2) * 0x10001, (e + s3) * 0x10001);
            const int4 p1 = make_int4((e + s4) * 0x10001, (e + s5) * 0x10001, (e + s6) * 0x10001, (e + s7) * 0x10001);
            __syncwarp();
            reinterpret_cast<int4*>(tab)[lane * 2] = p0;
            reinterpret_cast<int4*>(tab)[lane * 2 + 1] = p1;
        }
        __syncwarp();
        // ---- place every key: the row in place, grouped by bucket (every lane holds its keys in registers).  A second round
        //      of atomics on the cursors; taking the arrival numbers from the histogram's atomics instead (returning form, the
        //      placement then a table load) measured slower: 6.04 vs 5.30 ms (profiles/r02_r_qrf_ab.txt) -- the counting round
        //      is ATOMS.POPC.INC, which merges lanes on the same bucket, the returning form serialises them ----------------
        int at[R];
#pragma unroll
        for (int r = 0; r < R; ++r) at[r] = atomicAdd(slot[r], 1);                    // all in flight before the first store
#pragma unroll
        for (int r = 0; r < R; ++r) sv[at[r] & 0xFFFF] = x[r];
        __syncwarp();                                                                 // tab[b] = end rank | first rank << 16

        for (int j0 = 0; j0 < nprobs; j0 += kQrfGroup) {
            // ---- lane j < 3: quantile j0 + j.  R: index = 1 + (n-1)*p in two roundings, lo = floor, hi = ceil --
            const int jq = min(j0 + min(lane, kQrfGroup - 1), nprobs - 1);
            const double index = __dadd_rn(1.0, __dmul_rn((double)(m - 1), probs[jq]));
            const double lo = floor(index);
            const int k_mine = (int)lo - 1;                                           // 0-based rank of x[lo]
            uint32_t klo_mine = 0, khi_mine = 0;
#pragma unroll
            for (int g = 0; g < kQrfGroup; ++g) {
                if (g > 0 && j0 + g >= nprobs) break;                                 // uniform
                const int k = __shfl_sync(0xffffffffu, k_mine, g);
                // the bucket of rank k and, when k is its last rank, the next non-empty one (rank k+1): all uniform
                const uint32_t iv = (uint32_t)tab[bucket_of(sv[k])];
                const int first = (int)(iv >> 16);
                int end = (int)(iv & 0xFFFFu);
                if (k + 1 == end && end < R * 32) end = (int)((uint32_t)tab[bucket_of(sv[end])] & 0xFFFFu);
                const int cnt = end - first;
                int a = k - first;                                                    // rank k inside the window
                uint32_t klo, khi;
                if (cnt <= 32) {
                    // peel the window's distinct values from below (a tie group at a time) until the group that covers
                    // rank a is reached: about |bucket| / 2 warp reductions, all results warp-uniform
                    uint32_t cur = (lane < cnt) ? sv[first + lane] : NA;
#pragma unroll 1
                    for (;;) {
                        const uint32_t v = __reduce_min_sync(0xffffffffu, cur);
                        const bool eq = cur == v;
                        const int n_eq = __popc(__ballot_sync(0xffffffffu, eq));      // v == NA: all 32 lanes, ends the loop
                        cur = eq ? NA : cur;
                        if (a < n_eq) {
                            klo = v;
                            khi = (a + 1 < n_eq) ? v : __reduce_min_sync(0xffffffffu, cur);   // NA when rank k+1 does not exist
                            break;
                        }
                        a -= n_eq;
                    }
                } else {
                    // heavy window: more members than lanes means ties.  The same peeling with up to R keys per lane: the
                    // smallest remaining value, its multiplicity, drop it; a handful of distinct values at most
                    uint32_t y[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) y[r] = (lane + 32 * r < cnt) ? sv[first + lane + 32 * r] : NA;
                    bool found = false;
                    klo = khi = NA;
#pragma unroll 1
                    for (int it = 0; it < 8 && !found; ++it) {
                        uint32_t v = NA;
#pragma unroll
                        for (int r = 0; r < R; ++r) v = min(v, y[r]);
                        v = __reduce_min_sync(0xffffffffu, v);
                        unsigned n_eq = 0;
#pragma unroll
                        for (int r = 0; r < R; ++r) { n_eq += y[r] == v ? 1u : 0u; y[r] = y[r] == v ? NA : y[r]; }
                        n_eq = __reduce_add_sync(0xffffffffu, n_eq);
                        if (a < (int)n_eq) {                                          // v == NA: every slot of every lane, ends the loop
                            klo = khi = v;
                            if (a + 1 >= (int)n_eq) {                                 // rank k+1 is the next distinct value
                                uint32_t nx = NA;
#pragma unroll
                                for (int r = 0; r < R; ++r) nx = min(nx, y[r]);
                                khi = __reduce_min_sync(0xffffffffu, nx);
                            }
                            found = true;
                        }
                        a -= (int)n_eq;
                    }
                    if (!found) return false;        // uniform: the caller re-buckets with qrf_row_finish
                }
                if (lane == g) { klo_mine = klo; khi_mine = khi; }
            }
            // ---- lanes 0..2: interpolate and store ---------------------------------------------------
            double q = nan;
            if (lane < kQrfGroup && j0 + lane < nprobs) {
                const double xlo = __ldg(dict + klo_mine);
                // rank k+1 only exists as a sample when index > lo (otherwise it may be an NA key: no look-up)
                const double xhi = (khi_mine == klo_mine || !(index > lo)) ? xlo : __ldg(dict + khi_mine);
                q = xlo;
                if (index > lo && xhi != xlo) {       // R: i <- which(index > lo & x[hi] != qs)
                    const double h = index - lo;
                    q = __dadd_rn(__dmul_rn(1.0 - h, xlo), __dmul_rn(h, xhi));
                }
                if (out_q) out_q[(long long)(j0 + lane) * ld_q + row] = q;
            }
            if (j0 == 0) qfirst = __shfl_sync(0xffffffffu, q, 0);
            if (j0 + kQrfGroup >= nprobs) qlast = __shfl_sync(0xffffffffu, q, nprobs - 1 - j0);
        }
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
    return true;
}

// ---- DIRECT form: the scratch rows already hold the keys ------------------------------------------
// With the sector forest image the walk stores each leaf's quantile key (the dictionary rank of ranger's
// random.node.values entry) instead of the leaf number, so nothing is gathered here: a row of T keys is T
// coalesced 4-byte words.  Each warp owns a ring of three row buffers in shared memory filled by cp.async two rows
// ahead; the row being selected is used IN PLACE as qrf's sample array (no staging through registers, which is
// what held the gather form at 121 registers and two CTAs per SM).
template <int R>
__host__ __device__ constexpr size_t qrf_direct_warp_smem() { return (size_t)3 * R * 32 * 4 + 256 * 4 + 256 * 4 + 32 * 8; }
template <int R>
__host__ __device__ constexpr size_t qrf_direct_smem_bytes() { return kQrfWarps * qrf_direct_warp_smem<R>(); }
template <int R> constexpr int qrf_direct_min_blocks() { return R >= 32 ? 1 : (R >= 16 ? 3 : 4); }

template <int R>
__global__ void __launch_bounds__(32 * kQrfWarps, qrf_direct_min_blocks<R>())
qrf_direct_kernel(const uint32_t* __restrict__ keys, long long ld_ids, const double* __restrict__ dict,
                  const unsigned char* __restrict__ skip, long long nrows, int T, const double* __restrict__ probs, int nprobs,
                  double* __restrict__ out_q, long long ld_q, int16_t* __restrict__ out_iqr_i16, int fast)
{
    extern __shared__ __align__(16) unsigned char qrf_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned char* base = qrf_smem + (size_t)wib * qrf_direct_warp_smem<R>();
    uint32_t* ring = reinterpret_cast<uint32_t*>(base);
    QrfWarp<uint32_t> s;
    s.hist0 = reinterpret_cast<int*>(base + (size_t)3 * R * 32 * 4);
    s.hist1 = s.hist0 + 256;
    s.mem = reinterpret_cast<uint32_t*>(s.hist1 + 256);
    s.ids = nullptr;
    s.sv = ring;
    // this CTA's contiguous run of rows; its warps take them kQrfWarps at a time
    const long long per_cta = ((nrows + gridDim.x - 1) / gridDim.x + kQrfWarps - 1) / kQrfWarps * kQrfWarps;
    const long long row_end = min(nrows, (long long)(blockIdx.x + 1) * per_cta);
    const long long stride = kQrfWarps;
    long long row = (long long)blockIdx.x * per_cta + wib;
    if (row >= row_end) return;   // whole warp
    const int nchunk = (int)(ld_ids >> 2);   // 16-byte pieces of one key row
    constexpr uint32_t NA = 0xFFFFFFFFu;

    // keys of row rw -> ring slot (asynchronous); always commits a group.  Rows whose outputs are NA are fetched too (their scratch
    // rows exist, whatever they hold): the copy then does not wait for the skip byte, which is read two rows ahead into a register
    auto skip_of = [&](long long rw) -> unsigned { return (skip && rw < row_end) ? (unsigned)skip[rw] : 0u; };
    auto fetch = [&](long long rw, int slot) {
        if (rw < row_end) {
            const uint32_t* src = keys + rw * ld_ids + 4 * lane;
            uint32_t* dst = ring + slot * (R * 32) + 4 * lane;
#pragma unroll
            for (int i = 0; i < (R + 3) / 4; ++i)     // R * 8 pieces of 16 bytes at most, 32 per round
                if (lane + 32 * i < nchunk) qrf_cp_async16(dst + 128 * i, src + 128 * i);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    unsigned sk0 = skip_of(row), sk1 = skip_of(row + stride);
    fetch(row, 0);
    fetch(row + stride, 1);
    int slot = 0;
    for (; row < row_end; row += stride) {
        int nslot = slot + 2; nslot -= nslot >= 3 ? 3 : 0;
        const unsigned sk2 = skip_of(row + 2 * stride);
        fetch(row + 2 * stride, nslot);                            // its previous tenant (row - stride) is finished
        asm volatile("cp.async.wait_group 2;" ::: "memory");       // this row's group has landed
        __syncwarp();
        s.sv = ring + slot * (R * 32);
        const bool skipped = sk0 != 0u;
        sk0 = sk1; sk1 = sk2;
        bool done = false;
        if (skipped) {                                             // NA row (masked pixel, NA covariate, too few stations)
            if (out_q) for (int j = lane; j < nprobs; j += 32) out_q[(long long)j * ld_q + row] = na_real();
            if (lane == 0 && out_iqr_i16) out_iqr_i16[row] = (int16_t)-32767;
            done = true;
        } else if (fast) {
            for (int t = T + lane; t < R * 32; t += 32) s.sv[t] = NA;   // the tail beyond the last tree: NA keys sort last
            __syncwarp();
            done = qrf_row_fast<R>(s.sv, s.hist0, lane, row, probs, nprobs, dict, out_q, ld_q, out_iqr_i16);
            __syncwarp();
        }
        if (!done) {                                               // heavy ties next to other values: re-bucketing selection
            unsigned valid = 0;
            uint32_t mn = NA, mx = 0u;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int t = r * 32 + lane;
                // after qrf_row_fast the row is permuted (valid keys may sit beyond T) and its tail is NA-filled: take every slot
                const uint32_t x = (fast || t < T) ? s.sv[t] : NA;
                if (x != NA) { valid |= 1u << r; mn = x < mn ? x : mn; mx = x > mx ? x : mx; }   // na.rm = TRUE
            }
            qrf_row_finish<R, uint32_t>(s, valid, mn, mx, lane, row, probs, nprobs, dict, out_q, ld_q, out_iqr_i16);
        }
        __syncwarp();   // the row's shared memory is free again
        slot = slot + 1 == 3 ? 0 : slot + 1;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");               // nothing may land after the warp is gone
}

}  // namespace rfsi
