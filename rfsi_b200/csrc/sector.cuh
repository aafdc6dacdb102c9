// sector.cuh -- the "sector" forest image and its walk (stage 2, default image).
//
// Same job as traverse_b / traverse_q in forest.cuh: the forward pass of
// ranger::predict.ranger (simulation/simulation.R:214, temp_croatia/2_predictions.R:167,174-176,
// prcp_catalonia/5_prcp_prediction.R:287-291), decisions made on threshold RANKS so that they are
// bit-identical to the fp64 compare (common.cuh).
//
// Layout.  One block = 8 words = ONE 32-byte sector = three tree levels:
//     word 0      index of the first of the block's 8 child blocks (children are contiguous and
//                 come after their parent: child = w0 + exit, exit = 0..7 = the three decisions as bits)
//     words 1..7  implicit heap of 7 nodes, children of slot h at 2h, 2h+1;
//                 node word = (~rank << 10) | leaf flags << 8 | feature (rank: 22 bits): the walk goes right iff
//                 (rank(x) << 10) + word carries out of 32 bits, i.e. iff rank(x) > rank -- the low ten bits
//                 cannot make up the difference because rank(x) << 10 ends in zeros; the feature byte is
//                 stored as it is, so the shared-memory address is one AND and one multiply-add.
//                 The two flag bits only mean something in the third level (slots 4..7): bit 9 = the block's
//                 exit to the LEFT of this node is a leaf block, bit 8 = the exit to the right is.  A chain
//                 therefore knows that it is finished when it takes its exit, WITHOUT visiting the leaf's block:
//                 one block visit less per tree (round 2, profiles/r02_s_*: the leaf visit was 1 of 8.2 visits
//                 of a chain and walked three pass-through levels for nothing).
//     a LEAF is a block of its own, touched once per tree by the loads of its payload:
//                 word 0 = 0x80000000 | its own block index (builder bookkeeping; the walk never reads it)
//                 word 3 = leaf number (left to right within the tree, trees one after the other)
//                 word 5 = quantile-forest key of the leaf (rank of ranger's random.node.values
//                          entry in the forest-wide sorted dictionary, 0xFFFFFFFF = NA)
//                 words 6..7 = the leaf's prediction (fp64)
//     block 0 is the idle block: 8 zero words (first child = itself, all nodes pass-through, no leaf flags).
// A leaf met before a block's third level is padded with pass-through nodes (rank field 0 = rank 0x3FFFFF:
// never exceeded) down to ONE exit, whose flag says "leaf"; the child slots to the right of padding are never
// visited and stay zero.  The root entry of a single-node tree carries the finished flag itself.
//
// Why (profiles/r01_g): the 128-byte blocked image needs three dependent sector fetches per four
// levels (levels 1-3 | level 4 | exit word) plus a leaf-value gather; here a block visit is one
// sector -- one L1 miss, then two L1 hits -- per three levels, no exit load, and the leaf's value,
// number and quantile key sit in the leaf's own sector.
//
// Walk state per chain: ONE register bb = block index, or 0x80000000 | leaf block once the
// chain is finished.  Per block visit:
//     idx = max((int)bb, 0)                 finished chains idle on block 0 (one shared sector)
//     {w0, n1} = 64-bit load                three levels: node load, shared-memory rank load, add.cc/addc
//     leaf = (n3 << (22 + last carry)) & 0x80000000
//     bb = umax(bb, (w0 | leaf) + exit)     children have larger indices than their parent, a leaf exit
//                                           carries the flag bit, the idle block yields 0: no selects
// and the lane leaves the loop as soon as every chain carries the flag (sign bit of AND_j bb_j).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#if defined(RFSI_SECTOR_HOST_CHECK)
// tests/tools/sector_host.cpp compiles the very same loop for the host (no CUDA headers)
#include <algorithm>
#define RFSI_SECT_DEV inline
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
template <class T> inline T __ldg(const T* p) { return *p; }
#else
#define RFSI_SECT_DEV __device__ __forceinline__
#endif

namespace rfsi {

constexpr uint32_t kSLeafFlag = 0x80000000u;
constexpr int kSRankShift = 10;                    // node word = (~rank << 10) | leaf flags << 8 | feature
constexpr uint32_t kSPassWord = 0x00000000u;       // rank 0x3FFFFF, feature 0: never carries
constexpr uint32_t kSMaxRank = 0x3FFFFEu;
constexpr uint32_t kSLeafLeft = 1u << 9, kSLeafRight = 1u << 8;   // third-level nodes: the exit on that side is a leaf block
constexpr int kSMaxFeat = 256;
constexpr int kSWordLeafNo = 3, kSWordKey = 5, kSWordValue = 6;

struct SectorView {
    const uint32_t* words;   // 8 per block
    const uint32_t* root;    // [T] block of each tree's root (sector_root_entry: flagged leaf block for a single-node tree)
    // P*4 (bytes between two features of the rank tile) as a RUN-TIME value: given the immediate, ptxas strength-reduces
    // feat*1024 + base into IMAD.SHL + LOP3 + IMAD.IADD; from the constant bank it stays one IMAD (the walk is bound by
    // integer issue slots: profiles/r02_sass_*.txt).  block_bytes (32) is only read by the host check: the device uses the
    // immediate (sector_block), which measured faster than the run-time value.
    uint32_t block_bytes;
    uint32_t tile_stride;
};

// h + h + carry(xs + w): one level.  xs = rank(x) << 10, w = (~rank << 10) | flags << 8 | feature.
RFSI_SECT_DEV uint32_t sector_step(uint32_t h, uint32_t w, uint32_t xs) {
#if defined(RFSI_SECTOR_HOST_CHECK)
    return 2u * h + (uint32_t)(((uint64_t)xs + (uint64_t)w) >> 32);
#else
    uint32_t out;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %3, %3;\n\t}" : "=r"(out) : "r"(xs), "r"(w), "r"(h));
    return out;
#endif
}

// The third level of a block visit: the exit e = 2h + carry (0..7) and, from the same carry, the flag of that exit moved
// to bit 31 -- the third-level node holds it at bit 9 (left exit, carry 0: shift by 22) or bit 8 (right exit: shift by 23).
RFSI_SECT_DEV uint32_t sector_last_step(uint32_t h, uint32_t w, uint32_t xs, uint32_t& leaf) {
#if defined(RFSI_SECTOR_HOST_CHECK)
    const uint32_t c = (uint32_t)(((uint64_t)xs + (uint64_t)w) >> 32);
    leaf = (w << (22u + c)) & 0x80000000u;
    return 2u * h + c;
#else
    uint32_t out, sh;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %2, %3;\n\taddc.u32 %0, %4, %4;\n\taddc.u32 %1, 22, 0;\n\t}"
        : "=r"(out), "=r"(sh) : "r"(xs), "r"(w), "r"(h));
    leaf = (w << sh) & 0x80000000u;
    return out;
#endif
}

RFSI_SECT_DEV uint32_t sector_umax(uint32_t a, uint32_t b) { return a > b ? a : b; }

inline uint32_t sector_node_word(uint32_t rank, uint32_t feat) { return ((~rank) << kSRankShift) | feat; }

// address of block max(bb, 0): one signed max, then index * 32 + base as LEA + LEA.HI.X on a base held in ordinary registers.
// Measured forms of this address (profiles/r02_ab_block_address.txt; SASS instructions per iteration of 8 chains):
//   run-time block size, base in uniform registers (what ptxas makes of a kernel parameter): IMAD.WIDE + IADD3 + IADD3.X, 231;
//   run-time block size, base in registers: IMAD.WIDE + IADD3 + IMAD.X, 230 (+1.9 %);  immediate 32, base in registers: 222 (+3.9 %).
// The base gets into ordinary registers by adding word 0 of the idle block to it -- a zero that only the memory knows, so
// the sum is a per-thread value as far as ptxas can tell (a `mov` through inline PTX is seen through).
RFSI_SECT_DEV const uint32_t* sector_base(const SectorView& q) {
#if defined(RFSI_SECTOR_HOST_CHECK)
    return q.words;
#else
    return q.words + __ldg(q.words);
#endif
}
RFSI_SECT_DEV const uint32_t* sector_block(const SectorView& q, const uint32_t* words, uint32_t bb) {
#if defined(RFSI_SECTOR_HOST_CHECK)
    return reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(words) +
                                             (size_t)((int32_t)bb > 0 ? bb : 0u) * q.block_bytes);
#else
    const uint32_t* out;
    asm("{\n\t.reg .s32 t;\n\tmax.s32 t, %1, 0;\n\tmad.wide.u32 %0, t, 32, %2;\n\t}" : "=l"(out) : "r"(bb), "l"(words));
    return out;
#endif
}

// Walk trees [t_begin, t_end) for the pixel whose pre-shifted ranks are rs[f * P] (feature-major tile in shared
// memory), J trees in lock step.  `sum` carries the tree-order partial sum of the trees before; the updated sum is
// returned.  After each group of J trees `sink(t0, remaining, leaf_block[J])` sees the leaf blocks the trees
// t0 .. t0+J-1 ended in (entries beyond `remaining` are 0).  nvisit += internal nodes visited if COUNT.
template <int P, int J, bool COUNT, class Sink>
RFSI_SECT_DEV double traverse_s(const SectorView& q, int t_begin, int t_end, double sum, const uint32_t* rs, Sink&& sink,
                                unsigned& nvisit)
{
    const char* rsb = reinterpret_cast<const char*>(rs);
    const uint32_t* words = sector_base(q);
    for (int t0 = t_begin; t0 < t_end; t0 += J) {
        uint32_t bb[J];
#pragma unroll
        for (int j = 0; j < J; ++j) bb[j] = (t0 + j < t_end) ? __ldg(q.root + t0 + j) : kSLeafFlag;   // no tree: finished, idles on block 0
        for (;;) {
            uint32_t fin = kSLeafFlag;
#pragma unroll
            for (int j = 0; j < J; ++j) fin &= bb[j];
            if (fin) break;                                                   // every chain has taken a leaf exit
            uint2 w01[J];
            const uint32_t* blk[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                blk[j] = sector_block(q, words, bb[j]);
                w01[j] = __ldg(reinterpret_cast<const uint2*>(blk[j]));     // w0 and the level-1 node: one request
            }
            uint32_t h[J];                                                    // the decisions so far, as bits
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const uint32_t xs = *reinterpret_cast<const uint32_t*>(rsb + (w01[j].y & 0xFFu) * q.tile_stride);
                h[j] = sector_step(0u, w01[j].y, xs);
                if (COUNT) nvisit += (w01[j].y >> kSRankShift) ? 1u : 0u;
            }
            {
                uint32_t w[J];
#pragma unroll
                for (int j = 0; j < J; ++j) w[j] = __ldg(blk[j] + 2 + h[j]);  // heap slot 2 or 3, same sector: L1 hit
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const uint32_t xs = *reinterpret_cast<const uint32_t*>(rsb + (w[j] & 0xFFu) * q.tile_stride);
                    h[j] = sector_step(h[j], w[j], xs);
                    if (COUNT) nvisit += (w[j] >> kSRankShift) ? 1u : 0u;
                }
            }
            {
                uint32_t w[J];
#pragma unroll
                for (int j = 0; j < J; ++j) w[j] = __ldg(blk[j] + 4 + h[j]);  // heap slot 4..7
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const uint32_t xs = *reinterpret_cast<const uint32_t*>(rsb + (w[j] & 0xFFu) * q.tile_stride);
                    uint32_t leaf;
                    h[j] = sector_last_step(h[j], w[j], xs, leaf);            // the exit, 0..7, and its leaf flag
                    if (COUNT) nvisit += (w[j] >> kSRankShift) ? 1u : 0u;
                    // implicit child, no exit load; a leaf exit finishes the chain without a visit to the leaf's block
                    bb[j] = sector_umax(bb[j], (w01[j].x | leaf) + h[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < J; ++j) bb[j] &= ~kSLeafFlag;
#pragma unroll
        for (int j = 0; j < J; ++j)
            if (t0 + j < t_end)                                                // the leaf's own sector: its only visit
                sum += __ldg(reinterpret_cast<const double*>(words + (size_t)bb[j] * 8u + kSWordValue));
        sink(t0, t_end - t0, bb);
    }
    return sum;
}


// ---- host: build the image ---------------------------------------------------------------------
// node(i) -> {is_leaf, feature, left child index (right = left + 1), complemented node word, leaf number, value}
// for the wide device layout of engine.cu (tree t's root is node t).  Blocks of a tree are contiguous, in BFS order
// (children after their parent); trees are built independently (`tree_begin`, `tree_end` let callers run them on
// several threads) into `w` with LOCAL indices starting at 0, and `sector_relocate` moves a tree to its place.
struct SectorNode {
    bool leaf;
    uint32_t word;      // sector_node_word(rank, feature)          (internal nodes)
    uint32_t child;     // index of the left child, right = child+1 (internal nodes)
    uint32_t leaf_no;   //                                          (leaves)
    double value;       //                                          (leaves)
};

template <class NodeAt, class Vec>
inline void sector_build_tree(uint32_t root_node, NodeAt&& node, Vec& w)
{
    struct Todo { uint32_t node, block; };
    struct Item { uint32_t node; int h; bool dead; };
    w.assign(8, 0u);
    auto make_leaf = [&](uint32_t blk, const SectorNode& n) {
        uint32_t* b = &w[(size_t)blk * 8];
        b[0] = kSLeafFlag | blk;
        b[kSWordLeafNo] = n.leaf_no;
        b[kSWordKey] = 0xFFFFFFFFu;
        uint64_t bits;
        static_assert(sizeof(double) == 8, "fp64");
        __builtin_memcpy(&bits, &n.value, 8);
        b[kSWordValue] = (uint32_t)bits;
        b[kSWordValue + 1] = (uint32_t)(bits >> 32);
    };
    const SectorNode r = node(root_node);
    if (r.leaf) { make_leaf(0, r); return; }
    std::vector<Todo> todo(1, Todo{root_node, 0u});          // FIFO of inner blocks still to fill: BFS order
    size_t head = 0;
    Item stack[64];
    while (head < todo.size()) {
        const Todo cur = todo[head++];
        const uint32_t kids = (uint32_t)(w.size() / 8);
        w.resize(w.size() + 64, 0u);
        w[(size_t)cur.block * 8] = kids;
        int sp = 0;
        stack[sp++] = Item{cur.node, 1, false};
        while (sp > 0) {
            const Item it = stack[--sp];
            if (it.h >= 8) {                                   // an exit: child block kids + (h - 8)
                if (it.dead) continue;                         // right of a pass-through node: never visited
                const uint32_t child = kids + (uint32_t)(it.h - 8);
                const SectorNode n = node(it.node);
                if (n.leaf) {
                    make_leaf(child, n);
                    w[(size_t)cur.block * 8 + (it.h >> 1)] |= (it.h & 1) ? kSLeafRight : kSLeafLeft;   // told by the third-level node
                } else {
                    todo.push_back(Todo{it.node, child});
                }
                continue;
            }
            const SectorNode n = it.dead ? SectorNode{true, 0, 0, 0, 0.0} : node(it.node);
            if (n.leaf) {                                      // pad down to one exit: always left
                w[(size_t)cur.block * 8 + it.h] = kSPassWord;
                stack[sp++] = Item{it.node, 2 * it.h + 1, true};
                stack[sp++] = Item{it.node, 2 * it.h, it.dead};
            } else {
                w[(size_t)cur.block * 8 + it.h] = n.word;
                stack[sp++] = Item{n.child + 1, 2 * it.h + 1, false};
                stack[sp++] = Item{n.child, 2 * it.h, false};
            }
        }
    }
}

// move a tree built with local indices to global block index `base` (in place)
inline void sector_relocate(uint32_t* w, size_t nblocks, uint32_t base) {
    for (size_t b = 0; b < nblocks; ++b) {
        uint32_t& w0 = w[b * 8];
        const bool unused = w0 == 0u && w[b * 8 + 1] == 0u;   // a child slot that is never visited
        if (!unused) w0 += base;                               // leaf flag is the top bit: the sum stays below it
    }
}

// root[] entry of a tree whose blocks start at `base` (tree_words = its first block, relocated or not): the block to
// start at, or -- a single-node tree -- the finished flag and the leaf block at once
inline uint32_t sector_root_entry(const uint32_t* tree_words, uint32_t base) { return base | (tree_words[0] & kSLeafFlag); }

}  // namespace rfsi
