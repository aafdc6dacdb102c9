// sector_image.cpp -- PROTOTYPE, not part of the product yet.
//
// A forest image for the traversal kernel in which one 32-byte SECTOR holds three tree levels and
// the children of a block are implicit:
//
//     word 0      first child block (the 8 children of a block are contiguous: child = w0 + exit)
//     words 1..7  heap of 7 nodes, children of slot h at 2h, 2h+1; node = (rank << 8) | feature
//     a LEAF is a block of its own: word 0 = 0x80000000 | leaf number, words 2..3 = the leaf value
//
// Why: profiles/r01_g shows the blocked 128-byte image (4 levels + 16 exit words per line) bound by
// long-scoreboard stalls with an L1 sector hit rate of 44.5 % -- a block visit touches three sectors
// one after the other (levels 1-3, level 4, the exit word), i.e. three serialised L2 round trips
// per four levels.  Here a block visit is ONE sector (one round trip, then two L1 hits) per three
// levels and no exit load at all: 0.33 instead of 0.75 sector misses per level, 1.0 instead of 1.25
// load requests per level, and the leaf value arrives with the leaf's own sector.
//
// This file builds the image from ranger's arrays and walks it on the host, so that the layout is
// proven (tests/test_sector_image_prototype.py: terminal nodes equal the oracle's) and its traffic
// can be counted before any kernel is written.  The device loop is the body of walk() with J
// chains in lock step, exactly like traverse_b (csrc/forest.cuh).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

constexpr uint32_t kLeaf = 0x80000000u;
constexpr uint32_t kPass = (0xFFFFFFu << 8);   // rank 2^24-1: nothing ranks above it, the walk goes left

struct Image {
    int T = 0, F = 0;
    std::vector<uint32_t> w;                    // 8 words per block
    std::vector<uint32_t> root;                 // block of each tree's root
    std::vector<std::vector<double>> uvals;     // per feature: sorted distinct thresholds
    std::vector<int32_t> leaf_rid;              // leaf number -> ranger's node id inside its tree
    int64_t leaf_blocks = 0, inner_blocks = 0, unreachable = 0;
};

uint32_t rank_of(const std::vector<double>& u, double x) {          // thresholds strictly below x
    return (uint32_t)(std::lower_bound(u.begin(), u.end(), x) - u.begin());
}

}  // namespace

extern "C" int sector_build(int ntrees, const int64_t* off, const int32_t* left, const int32_t* right, const int32_t* var,
                            const double* val, int nfeat, void** out) {
    Image* im = new Image;
    im->T = ntrees; im->F = nfeat;
    im->uvals.resize((size_t)nfeat);
    for (int t = 0; t < ntrees; ++t)
        for (int64_t i = off[t]; i < off[t + 1]; ++i)
            if (left[i] != 0 || right[i] != 0) im->uvals[(size_t)var[i]].push_back(val[i]);
    for (auto& u : im->uvals) { std::sort(u.begin(), u.end()); u.erase(std::unique(u.begin(), u.end()), u.end()); }
    for (auto& u : im->uvals) if (u.size() >= 0xFFFFFFu) { delete im; return -1; }
    std::vector<uint32_t>& W = im->w;
    W.assign(8, kPass);          // block 0: the idle block of finished chains -- all pass-through,
    W[0] = 0u;                   // and its first child is itself
    auto new_blocks = [&](int n) { const size_t b = W.size() / 8; W.resize(W.size() + 8 * (size_t)n, 0u); return (uint32_t)b; };
    struct Todo { int64_t node; uint32_t block; };
    for (int t = 0; t < ntrees; ++t) {
        const int64_t base = off[t];
        auto is_leaf = [&](int64_t n) { return left[base + n] == 0 && right[base + n] == 0; };
        // leaf numbers in order (left to right), like the product's blocked image
        std::vector<uint32_t> leaf_no((size_t)(off[t + 1] - base), 0u);
        {
            std::vector<int64_t> st(1, 0);
            while (!st.empty()) {
                const int64_t n = st.back(); st.pop_back();
                if (is_leaf(n)) { leaf_no[(size_t)n] = (uint32_t)im->leaf_rid.size(); im->leaf_rid.push_back((int32_t)n); }
                else { st.push_back(right[base + n]); st.push_back(left[base + n]); }
            }
        }
        auto make_leaf = [&](uint32_t blk, int64_t n) {
            W[(size_t)blk * 8] = kLeaf | leaf_no[(size_t)n];
            const double v = val[base + n];
            memcpy(&W[(size_t)blk * 8 + 2], &v, 8);
            ++im->leaf_blocks;
        };
        const uint32_t r = new_blocks(1);
        im->root.push_back(r);
        std::vector<Todo> todo;
        if (is_leaf(0)) make_leaf(r, 0); else todo.push_back({0, r});
        while (!todo.empty()) {
            const Todo cur = todo.back(); todo.pop_back();
            ++im->inner_blocks;
            const uint32_t kids = new_blocks(8);
            W[(size_t)cur.block * 8] = kids;
            // fill the heap below slot h with the subtree at tree node n (n < 0: below a leaf, pass-through)
            struct Item { int64_t node; int h; bool dead; };
            std::vector<Item> st(1, Item{cur.node, 1, false});
            while (!st.empty()) {
                const Item it = st.back(); st.pop_back();
                if (it.h >= 8) {                                     // an exit: child block kids + (h - 8)
                    const uint32_t child = kids + (uint32_t)(it.h - 8);
                    if (it.dead) { ++im->unreachable; continue; }    // right of a pass-through node: never visited
                    if (is_leaf(it.node)) make_leaf(child, it.node); else todo.push_back({it.node, child});
                    continue;
                }
                if (it.dead || is_leaf(it.node)) {                   // pad down to the exits: always left
                    W[(size_t)cur.block * 8 + it.h] = kPass;
                    st.push_back(Item{it.node, 2 * it.h + 1, true});
                    st.push_back(Item{it.node, 2 * it.h, it.dead});
                } else {
                    const int64_t g = base + it.node;
                    W[(size_t)cur.block * 8 + it.h] = (rank_of(im->uvals[(size_t)var[g]], val[g]) << 8) | (uint32_t)var[g];
                    st.push_back(Item{right[g], 2 * it.h + 1, false});
                    st.push_back(Item{left[g], 2 * it.h, false});
                }
            }
        }
    }
    *out = im;
    return 0;
}

// what: 0 blocks, 1 bytes, 2 leaf blocks, 3 inner blocks, 4 child slots that can never be reached
extern "C" int64_t sector_stats(const void* h, int what) {
    const Image* im = static_cast<const Image*>(h);
    switch (what) {
        case 0: return (int64_t)im->w.size() / 8;
        case 1: return (int64_t)im->w.size() * 4;
        case 2: return im->leaf_blocks;
        case 3: return im->inner_blocks;
        case 4: return im->unreachable;
        default: return -1;
    }
}

// X: npix x F column-major.  out_node[p*T + t] = ranger's terminal node id, out_sum[p] = sum of leaf values in
// tree order; sectors (nullable) += 32-byte sectors touched, levels (nullable) += tree levels descended.
extern "C" void sector_walk(const void* h, const double* X, int64_t npix, int64_t ldx, int32_t* out_node, double* out_sum,
                            int64_t* sectors, int64_t* levels) {
    const Image* im = static_cast<const Image*>(h);
    const uint32_t* W = im->w.data();
    std::vector<uint32_t> rk((size_t)im->F);
    int64_t ns = 0, nl = 0;
    for (int64_t p = 0; p < npix; ++p) {
        for (int f = 0; f < im->F; ++f) rk[(size_t)f] = rank_of(im->uvals[(size_t)f], X[(int64_t)f * ldx + p]) << 8;
        double sum = 0.0;
        for (int t = 0; t < im->T; ++t) {
            uint32_t b = im->root[(size_t)t];
            for (;;) {
                const uint32_t* blk = W + (size_t)b * 8;
                ++ns;
                const uint32_t w0 = blk[0];
                if (w0 & kLeaf) {
                    if (out_node) out_node[p * im->T + t] = im->leaf_rid[w0 & ~kLeaf];
                    double v; memcpy(&v, blk + 2, 8);
                    sum += v;
                    break;
                }
                uint32_t hh = 1;
                for (int lev = 0; lev < 3; ++lev) {
                    const uint32_t w = blk[hh];
                    hh = 2 * hh + (rk[w & 0xFFu] > (w & 0xFFFFFF00u) ? 1u : 0u);     // right iff rank(x) > rank(threshold)
                    nl += (w != kPass);
                }
                b = w0 + (hh - 8);
            }
        }
        if (out_sum) out_sum[p] = sum;
    }
    if (sectors) *sectors += ns;
    if (levels) *levels += nl;
}

extern "C" void sector_free(void* h) { delete static_cast<Image*>(h); }

// ---- the device loop (sector_walk.cuh), compiled for the host: J chains in lock step -----------------
#define RFSI_PROTO_HOST 1
#include "sector_walk.cuh"

template <int J>
static void lockstep(const Image* im, const double* X, int64_t npix, int64_t ldx, double* out_sum, int32_t* out_node) {
    rfsi_proto::SectorForest q{im->w.data(), im->root.data(), 0u};
    std::vector<uint32_t> rk((size_t)im->F), leaf((size_t)im->T);
    for (int64_t p = 0; p < npix; ++p) {
        for (int f = 0; f < im->F; ++f) rk[(size_t)f] = rank_of(im->uvals[(size_t)f], X[(int64_t)f * ldx + p]) << 8;
        out_sum[p] = rfsi_proto::traverse_s<1, J>(q, 0, im->T, 0.0, rk.data(), leaf.data());
        for (int t = 0; t < im->T; ++t) out_node[p * im->T + t] = im->leaf_rid[leaf[(size_t)t]];
    }
}

extern "C" int sector_walk_lockstep(const void* h, const double* X, int64_t npix, int64_t ldx, int J, double* out_sum,
                                    int32_t* out_node) {
    const Image* im = static_cast<const Image*>(h);
    switch (J) {
        case 1: lockstep<1>(im, X, npix, ldx, out_sum, out_node); return 0;
        case 3: lockstep<3>(im, X, npix, ldx, out_sum, out_node); return 0;
        case 4: lockstep<4>(im, X, npix, ldx, out_sum, out_node); return 0;
        case 8: lockstep<8>(im, X, npix, ldx, out_sum, out_node); return 0;
        default: return -1;
    }
}

// ---- warp-level traffic of both images for the SAME walks (no GPU needed) ----------------------------------
// X holds groups of 32 pixels (one warp: an 8 x 4 patch of the raster).  For every tree the 32 lanes walk in lock
// step, level by level; a warp-level load costs one request and as many sectors as the lanes' addresses cover.
//   blocked image (product, csrc/common.cuh): block = 4 levels rooted at depth 0, 4, 8 ...; heap slots 1-7 live
//     in sector 0 of the block's 128-byte line, slots 8-15 in sector 1, exit words 16-23 / 24-31 in sectors 2 / 3;
//     five loads per block visit (four levels + the exit word), then one 8-byte leaf-value gather per tree;
//   sector image (this file): block = 3 levels = one sector; three loads per block visit, the leaf is a block.
// out[0..5] = {requests, sector lookups} of the blocked image, the same of the sector image, then the sectors each
// touches for the first time within a block visit (what has to come from L2 when nothing is left in L1).
extern "C" void warp_traffic(int ntrees, const int64_t* off, const int32_t* left, const int32_t* right, const int32_t* var,
                             const double* val, const double* X, int64_t npix, int64_t ldx, int64_t* out) {
    int64_t req4 = 0, sec4 = 0, req3 = 0, sec3 = 0, miss4 = 0, miss3 = 0;
    std::vector<int64_t> key(32), seen;
    auto distinct = [&](int n) { std::sort(key.begin(), key.begin() + n); return (int64_t)(std::unique(key.begin(), key.begin() + n) - key.begin()); };
    for (int64_t g = 0; g + 32 <= npix; g += 32) {
        for (int t = 0; t < ntrees; ++t) {
            const int64_t base = off[t];
            // per lane: the path (node ids) to its leaf
            std::vector<std::vector<int32_t>> path(32);
            size_t longest = 0;
            for (int l = 0; l < 32; ++l) {
                int32_t n = 0;
                for (;;) {
                    path[l].push_back(n);
                    const int64_t q = base + n;
                    if (left[q] == 0 && right[q] == 0) break;
                    n = (X[(int64_t)var[q] * ldx + g + l] <= val[q]) ? left[q] : right[q];
                }
                longest = std::max(longest, path[l].size() - 1);      // internal nodes on the path
            }
            // a lane's position at depth d: its node there, or its leaf once it is past it (padding / idle block)
            auto at = [&](int l, size_t d) { return path[l][std::min(d, path[l].size() - 1)]; };
            // a lane visits the block rooted at depth d iff its node there is internal (else the previous exit was a leaf)
            auto visits = [&](int l, size_t d) { return d == 0 || d < path[l].size() - 1; };
            for (int levels : {4, 3}) {
                int64_t& req = levels == 4 ? req4 : req3;
                int64_t& sec = levels == 4 ? sec4 : sec3;
                int64_t& miss = levels == 4 ? miss4 : miss3;      // sectors touched for the first time in a block visit
                for (size_t d0 = 0; d0 < std::max<size_t>(longest, 1); d0 += (size_t)levels) {
                    // lanes whose leaf lies above this block's root have left the tree: they idle on a shared block
                    seen.clear();
                    for (int lev = 0; lev < levels; ++lev) {
                        int n = 0;
                        for (int l = 0; l < 32; ++l) {
                            if (!visits(l, d0)) { key[n++] = -1; continue; }                     // idle block
                            const int64_t root = at(l, d0);
                            // heap slot inside the block from the path bits below the block root
                            int h = 1;
                            for (int k = 0; k < lev; ++k) {
                                const size_t d = d0 + (size_t)k;
                                const bool right_turn = d + 1 < path[l].size() && path[l][d + 1] == right[base + path[l][d]] && !(left[base + path[l][d]] == 0 && right[base + path[l][d]] == 0);
                                h = 2 * h + (right_turn ? 1 : 0);
                            }
                            const int sector = levels == 4 ? (h >= 8 ? 1 : 0) : 0;
                            key[n++] = root * 4 + sector;
                        }
                        ++req; sec += distinct(n);
                        for (int i = 0; i < n; ++i) if (key[i] >= 0) seen.push_back(key[i]);
                    }
                    if (levels == 4) {                                  // the exit word
                        int n = 0;
                        for (int l = 0; l < 32; ++l) {
                            if (!visits(l, d0)) { key[n++] = -1; continue; }
                            int h = 1;
                            for (int k = 0; k < 4; ++k) {
                                const size_t d = d0 + (size_t)k;
                                const bool right_turn = d + 1 < path[l].size() && path[l][d + 1] == right[base + path[l][d]] && !(left[base + path[l][d]] == 0 && right[base + path[l][d]] == 0);
                                h = 2 * h + (right_turn ? 1 : 0);
                            }
                            key[n++] = (int64_t)at(l, d0) * 4 + (h >= 24 ? 3 : 2);
                        }
                        ++req; sec += distinct(n);
                        for (int i = 0; i < n; ++i) if (key[i] >= 0) seen.push_back(key[i]);
                    }
                    std::sort(seen.begin(), seen.end());
                    miss += (int64_t)(std::unique(seen.begin(), seen.end()) - seen.begin());
                }
                // the leaf: blocked image gathers leaf_val[leaf] (8 B, 4 leaves per sector, leaves numbered in order ~ node id / 2);
                // sector image reads the leaf's own block (one more block visit: 1 request)
                int n = 0;
                for (int l = 0; l < 32; ++l) key[n++] = levels == 4 ? (int64_t)path[l].back() / 8 : (int64_t)path[l].back();
                const int64_t dl = distinct(n);
                ++req; sec += dl; miss += dl;
            }
        }
    }
    out[0] = req4; out[1] = sec4; out[2] = req3; out[3] = sec3; out[4] = miss4; out[5] = miss3;
}
