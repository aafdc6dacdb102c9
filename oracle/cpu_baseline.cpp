// cpu_baseline.cpp -- the timed CPU arm: the reference's *algorithms* for the hot
// path (exact kd-tree kNN as nabor/libnabo does it, ranger-style pointer-chasing
// tree descent on std::thread workers, R-style type-7 quantiles), restated in C++.
//
// TEST / BENCH INFRASTRUCTURE ONLY (same rule as rfsi_oracle.c): only tests/,
// smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
// Pinning as for rfsi_oracle.c: the kd-tree near.obs reproduces the reference's stored IDW
// cross-validation predictions (tests/test_oracle.py); the forest half is PARITY UNPINNED (R and
// ranger cannot run here, no saved forest in the reference).  It is a second, independent restatement: tests require it to
// agree bit for bit with rfsi_oracle.c (brute force) on ties, self matches and
// short station lists.
//
// Reference call sites it stands in for (timed there as DT / PT):
//   near.obs  simulation/simulation.R:190-197   (DT)
//   predict   simulation/simulation.R:213-215   (PT)
//   predict(type="quantiles") temp_croatia/2_predictions.R:174-176
//
// Build: oracle/Makefile (g++ -O2 -ffp-contract=off -pthread).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <thread>
#include <vector>

namespace {

inline double na_real() {
    union { uint64_t u; double d; } v;
    v.u = 0x7FF00000000007A2ULL;
    return v.d;
}

// ---------------------------------------------------------------- kd-tree --
// Bucketed 2-D kd-tree (bucket 8, like libnabo's default), median split on the
// wider axis.  The search keeps the k best under the total order (d2, index)
// and prunes a far subtree only when its plane distance squared is strictly
// greater than the current worst d2, so equal-distance points with a smaller
// index are always seen: the result is the same function of the data as the
// brute-force scan in rfsi_oracle.c.
struct KdTree {
    struct Node { int32_t lo, hi, left, right; int32_t axis; double split; };
    std::vector<Node> nodes;
    std::vector<int32_t> perm;
    const double* x; const double* y;
    static constexpr int kBucket = 8;

    KdTree(const double* ox, const double* oy, int32_t n) : x(ox), y(oy) {
        perm.resize(n);
        std::iota(perm.begin(), perm.end(), 0);
        if (n > 0) build(0, n);
    }
    int32_t build(int32_t lo, int32_t hi) {
        const int32_t id = (int32_t)nodes.size();
        nodes.push_back({lo, hi, -1, -1, -1, 0.0});
        if (hi - lo <= kBucket) return id;
        double mnx = 1e300, mxx = -1e300, mny = 1e300, mxy = -1e300;
        for (int32_t i = lo; i < hi; ++i) {
            mnx = std::min(mnx, x[perm[i]]); mxx = std::max(mxx, x[perm[i]]);
            mny = std::min(mny, y[perm[i]]); mxy = std::max(mxy, y[perm[i]]);
        }
        const int axis = (mxx - mnx >= mxy - mny) ? 0 : 1;
        const double* c = axis == 0 ? x : y;
        const int32_t mid = (lo + hi) / 2;
        std::nth_element(perm.begin() + lo, perm.begin() + mid, perm.begin() + hi,
                         [&](int32_t a, int32_t b) { return c[a] < c[b] || (c[a] == c[b] && a < b); });
        const double split = c[perm[mid]];
        const int32_t l = build(lo, mid);
        const int32_t r = build(mid, hi);
        nodes[id].left = l; nodes[id].right = r; nodes[id].axis = axis; nodes[id].split = split;
        return id;
    }
    struct Best { double* d; int32_t* i; int k; int cnt; };
    static inline bool before(double d2a, int32_t ia, double d2b, int32_t ib) {
        return d2a < d2b || (d2a == d2b && ia < ib);
    }
    void offer(Best& b, double d2, int32_t s) const {
        if (b.cnt == b.k && !before(d2, s, b.d[b.k - 1], b.i[b.k - 1])) return;
        int j = b.cnt < b.k ? b.cnt : b.k - 1;
        while (j > 0 && before(d2, s, b.d[j - 1], b.i[j - 1])) { b.d[j] = b.d[j - 1]; b.i[j] = b.i[j - 1]; --j; }
        b.d[j] = d2; b.i[j] = s;
        if (b.cnt < b.k) ++b.cnt;
    }
    void search(int32_t id, double px, double py, Best& b) const {
        const Node& nd = nodes[id];
        if (nd.left < 0) {
            for (int32_t i = nd.lo; i < nd.hi; ++i) {
                const int32_t s = perm[i];
                const double dx = px - x[s], dy = py - y[s];
                const double xx = dx * dx, yy = dy * dy;
                offer(b, xx + yy, s);
            }
            return;
        }
        const double q = nd.axis == 0 ? px : py;
        const double diff = q - nd.split;
        const int32_t near = diff < 0 ? nd.left : nd.right;
        const int32_t far = diff < 0 ? nd.right : nd.left;
        search(near, px, py, b);
        // every point beyond the plane has |d_axis| >= |diff|, and rounding is
        // monotone, so diff*diff is a lower bound of its computed d2
        if (b.cnt < b.k || !(diff * diff > b.d[b.k - 1])) search(far, px, py, b);
    }
};

template <class F>
void parallel_rows(int64_t n, int nthreads, F&& f) {
    if (nthreads <= 1 || n < 1024) { f(0, n); return; }
    std::vector<std::thread> th;
    const int64_t chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        const int64_t a = t * chunk, b = std::min(n, a + chunk);
        if (a >= b) break;
        th.emplace_back([=, &f] { f(a, b); });
    }
    for (auto& t : th) t.join();
}

double quantile7_sorted(const double* xs, int32_t m, double p) {
    if (m <= 0) return na_real();
    const double index = 1.0 + (double)(m - 1) * p;
    const double lo = std::floor(index), hi = std::ceil(index);
    const double xlo = xs[(int32_t)lo - 1], xhi = xs[(int32_t)hi - 1];
    if (index > lo && xhi != xlo) {
        const double h = index - lo;
        return (1.0 - h) * xlo + h * xhi;
    }
    return xlo;
}

}  // namespace

extern "C" {

int cpu_hardware_threads(void) {
    const unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

// near.obs with a kd-tree; same contract as oracle_near_obs.
int cpu_near_obs_kdtree(const double* loc_x, const double* loc_y, int64_t npix,
                        const double* obs_x, const double* obs_y, const double* obs_z,
                        int32_t nobs, int32_t n_obs, int32_t rm_dupl,
                        double* out_dist, double* out_obs, int32_t* out_idx, int nthreads)
{
    if (npix < 0 || nobs < 0 || n_obs < 1) return -3;
    KdTree tree(obs_x, obs_y, nobs);
    const int k = n_obs + 1;
    int short_flag = 0;
    int* any_short = &short_flag;
    parallel_rows(npix, nthreads, [&](int64_t a, int64_t b) {
        std::vector<double> bd(k);
        std::vector<int32_t> bi(k);
        bool shortl = false;
        for (int64_t p = a; p < b; ++p) {
            KdTree::Best best{bd.data(), bi.data(), k, 0};
            if (nobs > 0) tree.search(0, loc_x[p], loc_y[p], best);
            const int first = (rm_dupl && best.cnt > 0 && bd[0] == 0.0) ? 1 : 0;
            for (int j = 0; j < n_obs; ++j) {
                const int src = j + first;
                const int64_t o = (int64_t)j * npix + p;
                if (src < best.cnt) {
                    out_dist[o] = std::sqrt(bd[src]);
                    out_obs[o] = obs_z[bi[src]];
                    if (out_idx) out_idx[o] = bi[src] + 1;
                } else {
                    out_dist[o] = na_real();
                    out_obs[o] = na_real();
                    if (out_idx) out_idx[o] = std::numeric_limits<int32_t>::min();
                    shortl = true;
                }
            }
        }
        if (shortl) __atomic_store_n(any_short, 1, __ATOMIC_RELAXED);
    });
    return short_flag ? 1 : 0;
}

// predict.ranger (regression): rows split across threads, trees walked in order
// inside a row block so each row's sum is accumulated in tree order.
int cpu_forest_predict(int32_t ntrees, const int64_t* node_offsets,
                       const int32_t* left, const int32_t* right,
                       const int32_t* split_var, const double* split_val,
                       const double* X, int64_t npix, int64_t ldx,
                       double* out, int nthreads)
{
    if (ntrees < 1 || npix < 0 || ldx < npix) return -3;
    parallel_rows(npix, nthreads, [&](int64_t a, int64_t b) {
        constexpr int64_t kBlock = 256;
        for (int64_t r0 = a; r0 < b; r0 += kBlock) {
            const int64_t r1 = std::min(b, r0 + kBlock);
            for (int64_t r = r0; r < r1; ++r) out[r] = 0.0;
            for (int32_t t = 0; t < ntrees; ++t) {
                const int64_t o = node_offsets[t];
                const int32_t* L = left + o; const int32_t* R = right + o;
                const int32_t* V = split_var + o; const double* S = split_val + o;
                for (int64_t r = r0; r < r1; ++r) {
                    int32_t node = 0;
                    while (L[node] != 0 || R[node] != 0)
                        node = (X[(int64_t)V[node] * ldx + r] <= S[node]) ? L[node] : R[node];
                    out[r] += S[node];
                }
            }
            for (int64_t r = r0; r < r1; ++r) out[r] /= (double)ntrees;
        }
    });
    return 0;
}

int cpu_forest_quantiles(int32_t ntrees, const int64_t* node_offsets,
                         const int32_t* left, const int32_t* right,
                         const int32_t* split_var, const double* split_val,
                         const int64_t* rnv_offsets, const double* rnv,
                         const double* X, int64_t npix, int64_t ldx,
                         const double* probs, int32_t nprobs, double* out, int nthreads)
{
    if (ntrees < 1 || npix < 0 || ldx < npix || nprobs < 1) return -3;
    parallel_rows(npix, nthreads, [&](int64_t a, int64_t b) {
        std::vector<double> v(ntrees);
        for (int64_t r = a; r < b; ++r) {
            int32_t m = 0;
            for (int32_t t = 0; t < ntrees; ++t) {
                const int64_t o = node_offsets[t];
                const int32_t* L = left + o; const int32_t* R = right + o;
                int32_t node = 0;
                while (L[node] != 0 || R[node] != 0)
                    node = (X[(int64_t)split_var[o + node] * ldx + r] <= split_val[o + node]) ? L[node] : R[node];
                const double s = rnv[rnv_offsets[t] + node];
                if (!std::isnan(s)) v[m++] = s;
            }
            std::sort(v.begin(), v.begin() + m);
            for (int32_t q = 0; q < nprobs; ++q)
                out[(int64_t)q * npix + r] = quantile7_sorted(v.data(), m, probs[q]);
        }
    });
    return 0;
}

}  // extern "C"
