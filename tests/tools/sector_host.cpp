// sector_host.cpp -- TEST infrastructure: the product's sector-image builder and walk
// (rfsi_b200/csrc/sector.cuh) compiled for the HOST, so that the layout and the very loop the
// GPU runs can be checked against the oracle in the CPU suite (tests/test_sector_image.py).
// Not part of librfsi_b200.so; nothing in the product calls this.
#define RFSI_SECTOR_HOST_CHECK 1
#include "../../rfsi_b200/csrc/sector.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace {

struct Wide { bool leaf; int32_t feat; uint32_t child; double val; int32_t rid; };

struct Image {
    int T = 0, F = 0;
    std::vector<uint32_t> words, root;
    std::vector<std::vector<double>> u;     // per feature: sorted distinct thresholds
    std::vector<int32_t> leaf_rid;          // leaf number -> ranger node id
    int64_t leaf_blocks = 0, unused = 0;
};

}  // namespace

// ranger layout in; returns 0 or a negative code
extern "C" int sh_build(int T, const int64_t* off, const int32_t* left, const int32_t* right, const int32_t* var,
                        const double* val, int F, void** out)
{
    Image* im = new Image;
    im->T = T; im->F = F;
    // wide layout as rfsi_forest_create lays it out: roots first, bodies in BFS order, siblings adjacent
    const int64_t roots = (T + 1) & ~int64_t(1);
    std::vector<Wide> nodes((size_t)roots, Wide{true, F, 0, 0.0, 0});
    for (int t = 0; t < T; ++t) {
        struct It { int32_t rid; uint32_t slot; };
        std::vector<It> q(1, It{0, (uint32_t)t});
        for (size_t h = 0; h < q.size(); ++h) {
            const It it = q[h];
            const int64_t g = off[t] + it.rid;
            if (left[g] == 0 && right[g] == 0) { nodes[it.slot] = Wide{true, F, it.slot, val[g], it.rid}; continue; }
            const uint32_t c = (uint32_t)nodes.size();
            nodes.push_back(Wide{}); nodes.push_back(Wide{});
            nodes[it.slot] = Wide{false, var[g], c, val[g], it.rid};
            q.push_back(It{left[g], c}); q.push_back(It{right[g], c + 1});
        }
    }
    im->u.resize((size_t)F);
    for (const Wide& n : nodes) if (!n.leaf) im->u[(size_t)n.feat].push_back(n.val);
    for (auto& u : im->u) { std::sort(u.begin(), u.end()); u.erase(std::unique(u.begin(), u.end()), u.end()); }
    for (auto& u : im->u) if (u.size() > rfsi::kSMaxRank) { delete im; return -1; }
    // leaf numbers: left to right within each tree
    std::vector<uint32_t> leaf_no(nodes.size(), 0u);
    for (int t = 0; t < T; ++t) {
        std::vector<uint32_t> st(1, (uint32_t)t);
        while (!st.empty()) {
            const uint32_t n = st.back(); st.pop_back();
            if (nodes[n].leaf) { leaf_no[n] = (uint32_t)im->leaf_rid.size(); im->leaf_rid.push_back(nodes[n].rid); }
            else { st.push_back(nodes[n].child + 1); st.push_back(nodes[n].child); }
        }
    }
    auto node_at = [&](uint32_t i) {
        const Wide& n = nodes[i];
        rfsi::SectorNode s{n.leaf, 0u, n.child, leaf_no[i], n.val};
        if (!n.leaf) {
            const auto& u = im->u[(size_t)n.feat];
            const uint32_t rank = (uint32_t)(std::lower_bound(u.begin(), u.end(), n.val) - u.begin());
            s.word = rfsi::sector_node_word(rank, (uint32_t)n.feat);
        }
        return s;
    };
    im->words.assign(8, rfsi::kSPassWord);
    im->words[0] = 0u;                                         // the idle block
    for (int t = 0; t < T; ++t) {
        std::vector<uint32_t> w;
        rfsi::sector_build_tree((uint32_t)t, node_at, w);
        const uint32_t base = (uint32_t)(im->words.size() / 8);
        rfsi::sector_relocate(w.data(), w.size() / 8, base);
        im->root.push_back(rfsi::sector_root_entry(w.data(), base));
        im->words.insert(im->words.end(), w.begin(), w.end());
    }
    for (size_t b = 1; b < im->words.size() / 8; ++b) {
        if (im->words[b * 8] & rfsi::kSLeafFlag) ++im->leaf_blocks;
        else if (im->words[b * 8] == 0u && im->words[b * 8 + 1] == 0u) ++im->unused;
    }
    *out = im;
    return 0;
}

extern "C" int64_t sh_stats(const void* h, int what) {
    const Image* im = static_cast<const Image*>(h);
    switch (what) {
        case 0: return (int64_t)im->words.size() / 8;
        case 1: return im->leaf_blocks;
        case 2: return im->unused;
        default: return -1;
    }
}

extern "C" void sh_free(void* h) { delete static_cast<Image*>(h); }

// quantile keys of the leaves (word 5), by leaf number
extern "C" void sh_set_keys(void* h, const uint32_t* key_of_leaf) {
    Image* im = static_cast<Image*>(h);
    for (size_t b = 1; b < im->words.size() / 8; ++b)
        if (im->words[b * 8] & rfsi::kSLeafFlag) im->words[b * 8 + rfsi::kSWordKey] = key_of_leaf[im->words[b * 8 + rfsi::kSWordLeafNo]];
}

template <int J>
static void walk(const Image* im, const double* X, int64_t npix, int64_t ldx, int t_begin, int t_end, double* sum,
                 int32_t* node, uint32_t* key, uint64_t* visits)
{
    rfsi::SectorView q{im->words.data(), im->root.data(), 32u, 4u};   // P = 1: one pixel's ranks, 4 bytes apart
    std::vector<uint32_t> rk((size_t)im->F);
    for (int64_t p = 0; p < npix; ++p) {
        for (int f = 0; f < im->F; ++f) {
            const auto& u = im->u[(size_t)f];
            rk[(size_t)f] = (uint32_t)(std::lower_bound(u.begin(), u.end(), X[(int64_t)f * ldx + p]) - u.begin()) << rfsi::kSRankShift;
        }
        unsigned nv = 0;
        auto sink = [&](int t0, int remaining, const uint32_t (&lb)[J]) {
            for (int j = 0; j < J && j < remaining; ++j) {
                const uint32_t* blk = im->words.data() + (size_t)lb[j] * 8;
                if (node) node[p * im->T + t0 + j] = im->leaf_rid[blk[rfsi::kSWordLeafNo]];
                if (key) key[p * im->T + t0 + j] = blk[rfsi::kSWordKey];
            }
        };
        sum[p] = rfsi::traverse_s<1, J, true>(q, t_begin, t_end, sum[p], rk.data(), sink, nv);
        if (visits) *visits += nv;
    }
}

// X: npix x F column-major; sum[p] is read (partial sum of the trees before t_begin) and updated
extern "C" int sh_walk(const void* h, const double* X, int64_t npix, int64_t ldx, int J, int t_begin, int t_end, double* sum,
                       int32_t* node, uint32_t* key, uint64_t* visits)
{
    const Image* im = static_cast<const Image*>(h);
    switch (J) {
        case 1: walk<1>(im, X, npix, ldx, t_begin, t_end, sum, node, key, visits); return 0;
        case 2: walk<2>(im, X, npix, ldx, t_begin, t_end, sum, node, key, visits); return 0;
        case 4: walk<4>(im, X, npix, ldx, t_begin, t_end, sum, node, key, visits); return 0;
        case 8: walk<8>(im, X, npix, ldx, t_begin, t_end, sum, node, key, visits); return 0;
        case 12: walk<12>(im, X, npix, ldx, t_begin, t_end, sum, node, key, visits); return 0;
        case 16: walk<16>(im, X, npix, ldx, t_begin, t_end, sum, node, key, visits); return 0;
        default: return -1;
    }
}
