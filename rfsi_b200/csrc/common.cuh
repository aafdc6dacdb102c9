// common.cuh -- shared device/host definitions for the RFSI sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rfsi {

// One forest node, 16 bytes, loaded with a single 128-bit access.
//   internal: val = threshold, child = absolute index of the LEFT child (the right
//             child is child+1: siblings share one 32-byte sector), feat >= 0
//   leaf    : val = prediction, child = the leaf's OWN index, feat = nfeat (the sentinel
//             feature row, -inf for every pixel): x > val is never true there, so a leaf
//             maps to itself and the traversal loop needs no leaf test (forest.cuh)
// (ranger layout, SURVEY.md 8a row a7: split.values holds the threshold or, at a
// terminal node, the prediction.)
struct __align__(16) Node {
    double val;
    int32_t child;
    int32_t feat;
};
static_assert(sizeof(Node) == 16, "node record must be 16 bytes");

// Compact ("q") format: the same tree, same node indices, 8 bytes per node.  Thresholds are
// replaced by their rank among the sorted unique thresholds of their feature, features by the
// number of thresholds strictly below them, so x <= thr  <=>  rank(x) <= rank(thr): decisions
// are bit-identical to the fp64 compare.
//   w0 = ~(threshold rank), so "rank(x) > rank(thr)" is the carry of rank(x) + w0
//        (leaf: rank 0xFFFFFFFF -> w0 = 0, never carries)
//   w1 = (feat*4 << 20) | delta, delta = left child index - own index (leaf: w1 = 0, so a
//        leaf adds 0 to its own index: absorbing).  feat*4 is the byte offset of the feature
//        in the thread's rank vector in shared memory: one LEA.HI forms the LDS address.
// The root of each tree jumps from the root region into the tree body, too far for a 20-bit
// delta, so roots have their own record with an absolute child index.
struct __align__(16) QRoot {
    uint32_t rank;   // complemented, like w0
    uint32_t feat;
    uint32_t child;
    uint32_t pad;
};
constexpr uint32_t kQLeafRank = 0xFFFFFFFFu;
constexpr int kQDeltaBits = 20;
constexpr uint32_t kQMaxDelta = 1u << kQDeltaBits;
constexpr int kQMaxFeat = 1 << (32 - kQDeltaBits - 2);  // feat*4 must fit the top 12 bits

// Blocked ("b4") format, the default when it fits: the tree is cut into blocks of 4 levels,
// one 128-byte line each: 32 words, word 0 spare, words 1..15 an implicit heap of up to 15
// nodes (children of slot h at 2h and 2h+1), words 16..31 the 16 exits.
//   node word  = ~((rank << 8) | feat)   (feat < 256, rank < 2^24 - 1)
//   exit word  = index of the next block, or 0x80000000 | leaf node index
// A node goes right iff rank(x) > rank, i.e. iff (rank(x) << 8) + ~((rank << 8) | feat) carries
// (the feat byte can never make up the difference), so a level is one 32-bit node load, one
// 32-bit shared load of the pixel's pre-shifted rank, and add.cc / addc h = 2h + carry.  A
// leaf met before the fourth level is padded with pass-through nodes (rank 0xFFFFFF: never
// exceeded) down to its exits.  Per 4 levels: 5 loads of 4 bytes from ONE line instead of 4
// loads of 8 bytes from up to 4 lines.
constexpr uint32_t kBLeafFlag = 0x80000000u;
constexpr uint32_t kBPassWord = ~((0xFFFFFFu << 8) | 0u);   // complemented pass-through node
constexpr int kBMaxFeat = 256;
constexpr uint32_t kBMaxRank = 0xFFFFFEu;

// rank_of() narrows its binary search with a per-feature table of monotone buckets between the feature's smallest
// and largest threshold (forest.cuh).  The number of buckets is chosen per feature when the forest is created: about
// one bucket per distinct threshold (a power of two, at most kRankBucketsMax; RFSI_B200_RANK_BUCKETS lowers the
// cap).  Measured on the bench forest (7.7e5 thresholds per feature; profiles/r02_s_*, r02_t_*): the walk's prologue
// is 43 such searches per pixel-day whose loads diverge across a warp (neighbouring pixels sit hundreds of thresholds
// apart in the distance columns); 4096 buckets per feature (round 1, 8 search steps) -> 2^16 -> 2^18 -> 2^20 (about
// one step) is 130.2 -> 137.7 -> 140.2 -> 141.9 M pixel-days/s (with the leaf-flag image, itself +1 %).
constexpr int kRankBucketsMin = 64;
constexpr int kRankBucketsMax = 1 << 20;
// bucket of x: nbm1 = number of buckets - 1 (as a double, so the clamp is one compare)
__host__ __device__ __forceinline__ int rank_bucket(double x, double lo_half, double scale, double nbm1) {
#ifdef __CUDA_ARCH__
    const double t = __dmul_rn(__dsub_rn(__dmul_rn(x, 0.5), lo_half), scale);
#else
    const double t = (x * 0.5 - lo_half) * scale;   // host build: no FMA contraction on x86-64 SSE2
#endif
    return t >= nbm1 ? (int)nbm1 : (t > 0.0 ? (int)t : 0);
}

constexpr int32_t kIdxNone = 0x7fffffff;        // empty slot of a neighbour list
constexpr unsigned long long kNaRealBits = 0x7FF00000000007A2ULL;  // R NA_real_
constexpr int32_t kNaInteger = INT32_MIN;                          // R NA_integer_

__host__ __device__ __forceinline__ double na_real() {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)kNaRealBits);
#else
    union { unsigned long long u; double d; } v;
    v.u = kNaRealBits;
    return v.d;
#endif
}

#ifdef __CUDACC__
__device__ __forceinline__ Node load_node(const Node* __restrict__ p) {
    const int4 r = __ldg(reinterpret_cast<const int4*>(p));
    Node n;
    n.val = __hiloint2double(r.y, r.x);
    n.child = r.z;
    n.feat = r.w;
    return n;
}

// squared distance exactly as the CPU does it without FMA contraction:
// fl(fl(dx*dx) + fl(dy*dy)), x term first.
__device__ __forceinline__ double dist2_rn(double px, double py, double sx, double sy) {
    const double dx = __dsub_rn(px, sx);
    const double dy = __dsub_rn(py, sy);
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

}  // namespace rfsi
