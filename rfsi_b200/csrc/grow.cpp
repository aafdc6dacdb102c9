// grow.cpp -- host-side regression-forest grower (librfsi_grow.so).
//
// NOT on the accelerated path.  The reference trains with ranger::ranger()
// (simulation/simulation.R:208, temp_croatia/1_models.R:1056-1061) and training is out of
// scope for the CUDA engine.  This file exists because (a) the
// reference ships no trained forest (.MISSING_LARGE_BLOBS) and R is not installed here, so
// benchmarks need a way to get forests of realistic shape (500 trees x ~10^5 nodes for the
// scale-up config) in seconds, and (b) the rfsi() wrapper of README.md:93-110 needs a trainer
// behind it.  It grows CART regression trees the way ranger's defaults do -- bootstrap
// sample, mtry random candidate variables per node, variance-reduction split, x <= split
// goes left, min.node.size stop rule, leaf = in-bag mean, children appended in pairs.  Big nodes
// (>= kExactBelow in-bag rows) search split points on a per-variable grid of up to 255 quantile
// cut points (histogram CART: plenty of rows per bin there); smaller nodes sort their rows and
// try every midpoint between distinct values, exactly as ranger does.  (With the grid alone a
// node whose rows share a bin cannot be split on that variable: on the reference's Croatian
// data that cost 13 % RMSE, 1.58 vs 1.37 for exact CART -- tests/test_gpu_croatia_llocv.py.)
// Output is ranger's forest layout.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <thread>
#include <vector>

namespace {

struct Rng {  // splitmix64 / xorshift: deterministic per (seed, tree)
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ULL + 0xD1B54A32D192ED03ULL) { next(); next(); }
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    uint32_t below(uint32_t n) { return (uint32_t)((next() >> 32) * (uint64_t)n >> 32); }
};

struct Binned {
    int F = 0;
    int64_t n = 0;
    std::vector<std::vector<double>> cuts;  // per feature: ascending cut values (x <= cut -> bin <= i)
    std::vector<uint8_t> bins;              // [F][n]
};

Binned make_bins(const double* X, int64_t n, int F, int max_bins) {
    Binned b;
    b.F = F; b.n = n;
    b.cuts.resize(F);
    b.bins.resize((size_t)F * n);
    std::vector<double> col;
    for (int f = 0; f < F; ++f) {
        const double* x = X + (size_t)f * n;
        // cut candidates from a sorted subsample: midpoints between adjacent distinct values
        const int64_t ns = std::min<int64_t>(n, 20000);
        col.resize(ns);
        for (int64_t i = 0; i < ns; ++i) col[i] = x[(size_t)((double)i * n / ns)];
        std::sort(col.begin(), col.end());
        std::vector<double>& cuts = b.cuts[f];
        for (int q = 1; q < max_bins; ++q) {
            const int64_t i = (int64_t)((double)q * ns / max_bins);
            if (i <= 0 || i >= ns) continue;
            const double lo = col[i - 1], hi = col[i];
            if (!(hi > lo)) continue;
            double mid = (lo + hi) / 2.0;
            if (mid == hi) mid = lo;  // ranger's guard against a midpoint that rounds up
            if (cuts.empty() || mid > cuts.back()) cuts.push_back(mid);
        }
        uint8_t* bf = b.bins.data() + (size_t)f * n;
        for (int64_t i = 0; i < n; ++i)
            bf[i] = (uint8_t)(std::lower_bound(cuts.begin(), cuts.end(), x[i]) - cuts.begin());  // x <= cut[k] <=> bin <= k
    }
    return b;
}

constexpr int64_t kExactBelow = 4096;

struct Tree {
    std::vector<int32_t> left, right, var;
    std::vector<double> val;
    std::vector<double> sample;  // quantreg: one random response per terminal node
};

void grow_tree(const Binned& B, const double* X, const double* y, int mtry, int min_node_size, int max_depth,
               uint64_t seed, bool quantreg, Tree& tr)
{
    const int64_t n = B.n;
    const int F = B.F;
    Rng rng(seed);
    std::vector<int32_t> idx((size_t)n);
    for (int64_t i = 0; i < n; ++i) idx[i] = (int32_t)rng.below((uint32_t)n);  // bootstrap, with replacement
    struct Work { int64_t lo, hi; int depth; };
    std::vector<Work> work;
    tr.left.assign(1, 0); tr.right.assign(1, 0); tr.var.assign(1, 0); tr.val.assign(1, 0.0);
    work.push_back({0, n, 0});
    std::vector<int> feats(F);
    std::iota(feats.begin(), feats.end(), 0);
    double hs[256]; int32_t hc[256];
    std::vector<std::pair<double, double>> xy;   // (value, response) of a small node, one variable
    for (size_t node = 0; node < work.size(); ++node) {  // children are appended: index order == ranger's
        const Work w = work[node];
        const int64_t cnt = w.hi - w.lo;
        double sum = 0.0;
        for (int64_t i = w.lo; i < w.hi; ++i) sum += y[idx[i]];
        bool split = cnt > min_node_size && (max_depth <= 0 || w.depth < max_depth);
        int best_f = -1;
        double best_gain = 0.0, best_thr = 0.0;
        const bool exact = cnt < kExactBelow;
        if (split) {
            for (int m = 0; m < std::min(mtry, F); ++m) {  // mtry distinct candidate variables
                const int j = m + (int)rng.below((uint32_t)(F - m));
                std::swap(feats[m], feats[j]);
                const int f = feats[m];
                if (exact) {   // every midpoint between distinct sorted values
                    const double* xf = X + (size_t)f * n;
                    xy.resize((size_t)cnt);
                    for (int64_t i = 0; i < cnt; ++i) { const int32_t r = idx[w.lo + i]; xy[i] = {xf[r], y[r]}; }
                    std::sort(xy.begin(), xy.end(), [](const std::pair<double, double>& a, const std::pair<double, double>& b) { return a.first < b.first; });
                    double ls = 0.0;
                    for (int64_t i = 0; i + 1 < cnt; ++i) {
                        ls += xy[i].second;
                        if (!(xy[i].first < xy[i + 1].first)) continue;
                        const int64_t lc = i + 1, rc = cnt - lc;
                        const double rs = sum - ls;
                        const double gain = ls * ls / (double)lc + rs * rs / (double)rc;
                        if (gain > best_gain) {
                            best_gain = gain; best_f = f;
                            double mid = (xy[i].first + xy[i + 1].first) / 2.0;
                            if (mid == xy[i + 1].first) mid = xy[i].first;   // ranger's guard
                            best_thr = mid;
                        }
                    }
                    continue;
                }
                const int nb = (int)B.cuts[f].size();
                if (nb == 0) continue;
                const uint8_t* bf = B.bins.data() + (size_t)f * n;
                std::fill(hs, hs + nb + 1, 0.0);
                std::fill(hc, hc + nb + 1, 0);
                for (int64_t i = w.lo; i < w.hi; ++i) { const int32_t r = idx[i]; hs[bf[r]] += y[r]; ++hc[bf[r]]; }
                double ls = 0.0; int64_t lc = 0;
                for (int b = 0; b < nb; ++b) {
                    ls += hs[b]; lc += hc[b];
                    const int64_t rc = cnt - lc;
                    if (lc == 0 || rc == 0) continue;
                    const double rs = sum - ls;
                    const double gain = ls * ls / (double)lc + rs * rs / (double)rc;  // ranger's decrease
                    if (gain > best_gain) { best_gain = gain; best_f = f; best_thr = B.cuts[f][b]; }
                }
            }
            // no improvement over the unsplit node (sum^2/cnt) -> terminal
            if (best_f < 0 || !(best_gain > sum * sum / (double)cnt + 1e-12 * std::fabs(best_gain))) split = false;
        }
        if (!split) {
            tr.val[node] = sum / (double)cnt;
            continue;
        }
        const double* xb = X + (size_t)best_f * n;
        const int64_t mid = std::partition(idx.begin() + w.lo, idx.begin() + w.hi,
                                           [&](int32_t r) { return xb[r] <= best_thr; }) - idx.begin();
        if (mid == w.lo || mid == w.hi) { tr.val[node] = sum / (double)cnt; continue; }   // degenerate: keep a leaf
        const int32_t l = (int32_t)tr.left.size();
        tr.left[node] = l; tr.right[node] = l + 1; tr.var[node] = best_f; tr.val[node] = best_thr;
        for (int c = 0; c < 2; ++c) { tr.left.push_back(0); tr.right.push_back(0); tr.var.push_back(0); tr.val.push_back(0.0); }
        work.push_back({w.lo, mid, w.depth + 1});
        work.push_back({mid, w.hi, w.depth + 1});
    }
    if (quantreg) {
        // ranger's R code: every training row, in a random order, overwrites the cell of its
        // terminal node -> one random response per terminal node (NA where no row lands)
        tr.sample.assign(tr.left.size(), std::numeric_limits<double>::quiet_NaN());
        std::vector<int32_t> perm((size_t)n);
        std::iota(perm.begin(), perm.end(), 0);
        for (int64_t i = n - 1; i > 0; --i) std::swap(perm[i], perm[rng.below((uint32_t)(i + 1))]);
        for (int64_t i = 0; i < n; ++i) {
            const int32_t r = perm[i];
            int32_t nd = 0;
            while (tr.left[nd] != 0) nd = (X[(size_t)tr.var[nd] * n + r] <= tr.val[nd]) ? tr.left[nd] : tr.right[nd];
            tr.sample[nd] = y[r];
        }
    }
}

}  // namespace

extern "C" {

// X: n x F column-major.  Outputs are malloc'ed here and released with rfsi_grow_free.
int rfsi_grow_forest(const double* X, const double* y, int64_t n, int32_t F, int32_t ntrees, int32_t mtry,
                     int32_t min_node_size, int32_t max_depth, uint64_t seed, int32_t nthreads, int32_t quantreg,
                     int64_t** node_offsets, int32_t** left, int32_t** right, int32_t** split_var,
                     double** split_val, double** node_sample)
{
    if (!X || !y || n < 1 || n > 0x7fffffff || F < 1 || ntrees < 1) return -3;
    if (mtry < 1) mtry = std::max(1, (int)std::floor(std::sqrt((double)F)));
    if (min_node_size < 1) min_node_size = 5;
    if (nthreads < 1) nthreads = std::max(1u, std::thread::hardware_concurrency());
    const Binned B = make_bins(X, n, F, 256);
    std::vector<Tree> trees((size_t)ntrees);
    std::atomic<int> next{0};
    auto worker = [&] {
        for (int t; (t = next.fetch_add(1)) < ntrees;)
            grow_tree(B, X, y, mtry, min_node_size, max_depth, seed * 1000003ULL + (uint64_t)t, quantreg != 0, trees[t]);
    };
    std::vector<std::thread> th;
    for (int i = 0; i < std::min<int>(nthreads, ntrees); ++i) th.emplace_back(worker);
    for (auto& t : th) t.join();

    int64_t total = 0;
    for (auto& t : trees) total += (int64_t)t.left.size();
    *node_offsets = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ntrees + 1));
    *left = (int32_t*)malloc(sizeof(int32_t) * (size_t)total);
    *right = (int32_t*)malloc(sizeof(int32_t) * (size_t)total);
    *split_var = (int32_t*)malloc(sizeof(int32_t) * (size_t)total);
    *split_val = (double*)malloc(sizeof(double) * (size_t)total);
    if (node_sample) *node_sample = quantreg ? (double*)malloc(sizeof(double) * (size_t)total) : nullptr;
    int64_t o = 0;
    for (int t = 0; t < ntrees; ++t) {
        const Tree& tr = trees[t];
        const size_t m = tr.left.size();
        (*node_offsets)[t] = o;
        memcpy(*left + o, tr.left.data(), m * sizeof(int32_t));
        memcpy(*right + o, tr.right.data(), m * sizeof(int32_t));
        memcpy(*split_var + o, tr.var.data(), m * sizeof(int32_t));
        memcpy(*split_val + o, tr.val.data(), m * sizeof(double));
        if (quantreg && node_sample) memcpy(*node_sample + o, tr.sample.data(), m * sizeof(double));
        o += (int64_t)m;
    }
    (*node_offsets)[ntrees] = o;
    return 0;
}

void rfsi_grow_free(void* p) { free(p); }

}  // extern "C"
