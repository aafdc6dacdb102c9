// sector_walk.cuh -- PROTOTYPE device loop for the sector image (tools/prototypes/sector_image.cpp,
// DESIGN.md section 9 item 2).  Not included by the product; compile-checked for sm_100a by
// tools/prototypes/sector_walk_check.cu.  Same contract as traverse_b (csrc/forest.cuh): rs holds the
// pixel's ranks << 8 as rs[f * P] (feature-major tile in shared memory), J chains advance in lock
// step, leaf values are added in tree order.
#pragma once
#include <cstdint>
#include <cstddef>

#if defined(RFSI_PROTO_HOST)   // the same loop compiled for the host (sector_image.cpp: sector_walk_lockstep)
#include <algorithm>
#define __device__
#define __forceinline__ inline
#define __restrict__
struct uint2 { uint32_t x, y; };
template <class T> inline T __ldg(const T* p) { return *p; }
using std::min;
#endif

namespace rfsi_proto {

constexpr uint32_t kSLeaf = 0x80000000u;

struct SectorForest {
    const uint32_t* __restrict__ words;   // 8 per block: w0 = first child | leaf flag + number, w1..w7 = heap
    const uint32_t* __restrict__ root;    // block of each tree's root
    uint32_t idle;                        // an all-pass-through block whose w0 is its own index
};

// 2h + (xs > rank(w)): node word = (rank << 8) | feature, xs = rank(x) << 8
__device__ __forceinline__ uint32_t s_step(uint32_t h, uint32_t w, uint32_t xs) {
    return 2u * h + ((xs > (w & 0xFFFFFF00u)) ? 1u : 0u);
}

template <int P, int J>
__device__ __forceinline__ double traverse_s(const SectorForest& q, int t_begin, int t_end, double sum,
                                             const uint32_t* __restrict__ rs, uint32_t* leaf_out /* nullable: [t] */)
{
    const char* rsb = reinterpret_cast<const char*>(rs);
    for (int t0 = t_begin; t0 < t_end; t0 += J) {
        uint32_t bb[J], done[J];
#pragma unroll
        for (int j = 0; j < J; ++j) { bb[j] = __ldg(q.root + min(t0 + j, t_end - 1)); done[j] = 0u; }
        uint32_t live;
        do {
            uint2 w01[J];
            const uint32_t* blk[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                blk[j] = q.words + (size_t)(done[j] ? q.idle : bb[j]) * 8u;
                w01[j] = __ldg(reinterpret_cast<const uint2*>(blk[j]));            // w0 and the level-1 node: one request
            }
            uint32_t h[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const bool leaf = (w01[j].x & kSLeaf) != 0u;
                done[j] = done[j] | (leaf ? 1u : 0u);                              // a leaf block keeps its index in bb: value read below
                // words 2..3 of a leaf block are its VALUE, not nodes: finish this round on the idle block so that no
                // feature index is ever taken from them (it would address shared memory outside the rank tile)
                blk[j] = leaf ? q.words + (size_t)q.idle * 8u : blk[j];
                const uint32_t n1 = leaf ? 0xFFFFFF00u : w01[j].y;                   // (word 1 of a leaf block is unused: treat as pass-through)
                const uint32_t xs = *reinterpret_cast<const uint32_t*>(rsb + (n1 & 0xFFu) * (P * 4));
                h[j] = s_step(1u, n1, xs);
            }
#pragma unroll
            for (int lev = 1; lev < 3; ++lev) {
                uint32_t w[J];
#pragma unroll
                for (int j = 0; j < J; ++j) w[j] = __ldg(blk[j] + h[j]);                // same sector: L1 hits
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const uint32_t xs = *reinterpret_cast<const uint32_t*>(rsb + (w[j] & 0xFFu) * (P * 4));
                    h[j] = s_step(h[j], w[j], xs);
                }
            }
            live = 0u;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                bb[j] = done[j] ? bb[j] : (w01[j].x + (h[j] - 8u));                     // implicit child: no exit load
                live |= done[j] ^ 1u;
            }
        } while (live);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            if (t0 + j < t_end) {
                const uint32_t* blk = q.words + (size_t)bb[j] * 8u;
                sum += __ldg(reinterpret_cast<const double*>(blk + 2));                  // the leaf's own sector (already in L1)
                if (leaf_out) leaf_out[t0 + j] = __ldg(blk) & ~kSLeaf;
            }
        }
    }
    return sum;
}

}  // namespace rfsi_proto
