// engine.cu -- host side of librfsi_b200.so: forest import, kernel launches, C ABI.
// Entry points are declared (with the reference call each one replaces) in
// include/rfsi_b200.h.  No CPU compute path exists in this library: without an
// sm_100 device every compute call returns RFSI_E_NO_DEVICE.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cerrno>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include <sys/mman.h>
#include <unistd.h>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/rfsi_b200.h"
#include "common.cuh"
#include "features.cuh"
#include "forest.cuh"
#include "fused.cuh"
#include "knn.cuh"
#include "knn_cells.cuh"
#include "qrf.cuh"
#include "raster.cuh"

using namespace rfsi;

// ------------------------------------------------------------------------------------
// errors, devices, options
// ------------------------------------------------------------------------------------
namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU_TRY(expr)                                                                         \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return fail(e__ == cudaErrorMemoryAllocation ? RFSI_E_OOM : RFSI_E_CUDA,         \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,   \
                        __LINE__);                                                           \
    } while (0)

#define RF_TRY(expr)              \
    do {                          \
        int rc__ = (expr);        \
        if (rc__ < 0) return rc__; \
    } while (0)

// fork-safety: the pid that first made a CUDA call through this library.  A forked child inherits the value and is
// refused (a CUDA context does not survive fork()) with a message that says what to do instead of the runtime's
// "initialization error" (prcp_catalonia/5_prcp_prediction.R:243-244 forks its day workers).
std::atomic<long> g_cuda_pid{0};
int check_fork() {
    const long me = (long)getpid();
    long seen = 0;
    if (g_cuda_pid.compare_exchange_strong(seen, me) || seen == me) return RFSI_OK;
    return fail(RFSI_E_FORKED,
                "this process (pid %ld) was forked from pid %ld after that process had used the GPU through rfsi_b200; a CUDA "
                "context does not survive fork(). Import models in the parent (rfsi_forest_create needs no device) and make "
                "the first device call inside each worker, or start workers with spawn / PSOCK clusters",
                me, seen);
}

int use_device(int device) {
    RF_TRY(check_fork());
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(RFSI_E_NO_DEVICE, "no CUDA device visible (%s); this library has no CPU path",
                    e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return fail(RFSI_E_NO_DEVICE, "device %d out of range [0,%d)", device, n);
    int major = 0;
    CU_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10)
        return fail(RFSI_E_NO_DEVICE, "device %d is sm_%d0, this build is sm_100a only", device, major);
    CU_TRY(cudaSetDevice(device));
    // temporaries come from the stream-ordered pool; keep freed blocks cached so repeated
    // calls do not pay cudaMalloc/cudaFree (hundreds of ms for multi-GB scratch) every time
    static std::mutex mu;
    static std::vector<char> tuned;
    std::lock_guard<std::mutex> lk(mu);
    if ((int)tuned.size() <= device) tuned.resize(device + 1, 0);
    if (!tuned[device]) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        tuned[device] = 1;
    }
    return RFSI_OK;
}

int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// Optional per-kernel device timing (bench.py's roofline): CUDA events recorded on the
// launching stream around each launch, summed when read.
enum ProfClass { PROF_KNN_RAW = 0, PROF_FOREST_X = 1, PROF_FUSED = 2, PROF_FALLBACK = 3, PROF_QRF = 4,
                 PROF_KNN_FINAL = 5, PROF_NCLASS = 6 };
std::atomic<int> g_prof_on{0};
std::mutex g_prof_mu;
struct ProfRec { int cls; cudaEvent_t a, b; };
std::vector<ProfRec> g_prof;

struct ProfScope {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t st;
    int cls;
    ProfScope(int c, cudaStream_t s) : st(s), cls(c) {
        if (!g_prof_on.load(std::memory_order_relaxed)) return;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { a = b = nullptr; return; }
        cudaEventRecord(a, st);
    }
    ~ProfScope() {
        if (!a) return;
        cudaEventRecord(b, st);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof.push_back({cls, a, b});
    }
};

// Device temporary from the stream-ordered pool, allocated/freed on the legacy stream.
// Callers host-synchronise their work streams before a DevBuf dies, and call
// pool_ready() after their allocations before touching them from other streams.
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFreeAsync(p, nullptr); }
    int alloc(size_t bytes) {
        if (p) { cudaFreeAsync(p, nullptr); p = nullptr; }
        CU_TRY(cudaMallocAsync(&p, bytes ? bytes : 1, nullptr));
        return RFSI_OK;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
inline int pool_ready() {
    CU_TRY(cudaStreamSynchronize(nullptr));
    return RFSI_OK;
}

struct Stream {
    cudaStream_t s = nullptr;
    ~Stream() { if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); } }  // declared after the DevBufs it uses
    int create() { CU_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); return RFSI_OK; }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// RFSI_B200_TRACE=1: host-side phase timings of the host-buffer entry points on stderr
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    Trace() : on(env_int("RFSI_B200_TRACE", 0) != 0), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[rfsi_b200 trace] %-28s %8.2f ms\n", what,
                std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};


// ---- host staging ------------------------------------------------------------------------------------
// The R-shaped entry points hand over PAGEABLE memory, and their outputs are freshly allocated R vectors whose
// pages have never been touched.  Measured on the B200 hosts (profiles/r02_a_host_copy_bench.txt, 100 MB): D2H into
// pinned memory 1.8 ms, cudaMemcpy into mapped pageable memory 5.4 ms, into FRESH pageable memory 46.5 ms (the
// driver's single staging thread takes every page fault); memcpy pinned -> fresh pages with 4 / 8 host threads
// 5.3 / 3.5 ms.  So results travel device -> pinned slot -> caller's array, the last hop on several host threads,
// each thread driving its own chunks through its own stream (chunk k, k + W, ... on worker k of W).
class PinnedPool {
public:
    static PinnedPool& get() { static PinnedPool p; return p; }
    void* acquire(size_t bytes) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (auto& b : bufs_)
                if (!b.busy && b.bytes >= bytes && b.pid == (long)getpid()) { b.busy = true; return b.p; }
        }
        void* p = nullptr;
        if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        std::lock_guard<std::mutex> lk(mu_);
        bufs_.push_back({p, bytes, true, (long)getpid()});
        return p;
    }
    void release(void* p) {
        std::lock_guard<std::mutex> lk(mu_);
        for (auto& b : bufs_) if (b.p == p) b.busy = false;
    }
private:
    struct Buf { void* p; size_t bytes; bool busy; long pid; };
    std::mutex mu_;
    std::vector<Buf> bufs_;
};
struct PinnedBuf {
    void* p = nullptr;
    ~PinnedBuf() { if (p) PinnedPool::get().release(p); }
    int alloc(size_t bytes) {
        p = PinnedPool::get().acquire(bytes);
        if (!p) return fail(RFSI_E_OOM, "pinned staging buffer of %zu bytes", bytes);
        return RFSI_OK;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
// Map the pages of a destination range in one system call before writing it (MADV_POPULATE_WRITE, Linux 5.14+; a
// no-op where it is unknown): an R vector or NumPy array that was just allocated has no pages yet, and 25 000
// single page faults from four threads contend on the process's mmap lock (DT 17-19 ms instead of 7 ms at
// simulation.R's size when the 50 columns are fresh 2 MB allocations, profiles/r02_f_*).
inline void prefault_for_write(void* p, size_t bytes) {
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
    static const long page = sysconf(_SC_PAGESIZE);
    static std::atomic<int> usable{1};
    if (!usable.load(std::memory_order_relaxed) || bytes < (size_t)(16 * page)) return;
    const uintptr_t a = ((uintptr_t)p + (uintptr_t)page - 1) & ~((uintptr_t)page - 1);
    const uintptr_t b = ((uintptr_t)p + bytes) & ~((uintptr_t)page - 1);
    if (b <= a) return;
    if (madvise(reinterpret_cast<void*>(a), b - a, MADV_POPULATE_WRITE) != 0 && errno == EINVAL) usable.store(0);
}

inline int host_workers(int64_t nchunks) {
    const int want = env_int("RFSI_B200_HOST_THREADS", 8);   // profiles/r02_p_dt_probe.md: 2 / 4 / 8 / 16 threads -> DT 21.9 / 7.7 / 6.4 / 6.4 ms
    const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    return (int)std::max<int64_t>(1, std::min<int64_t>(std::min(want, hw), nchunks));
}
// run fn(worker) on `workers` host threads (the caller's thread is worker 0); first error wins, messages carried over
template <class Fn>
int run_workers(int workers, int device, Fn&& fn) {
    std::vector<int> rc((size_t)workers, RFSI_OK);
    std::vector<std::string> msg((size_t)workers);
    std::vector<std::thread> th;
    for (int w = 1; w < workers; ++w)
        th.emplace_back([&, w] {
            if (cudaSetDevice(device) != cudaSuccess) { rc[(size_t)w] = RFSI_E_CUDA; msg[(size_t)w] = "cudaSetDevice failed in a staging thread"; return; }
            rc[(size_t)w] = fn(w);
            if (rc[(size_t)w] != RFSI_OK) msg[(size_t)w] = g_err;   // g_err is thread-local
        });
    rc[0] = fn(0);
    if (rc[0] != RFSI_OK) msg[0] = g_err;
    for (auto& t : th) t.join();
    for (int w = 0; w < workers; ++w)
        if (rc[(size_t)w] < 0) { g_err = msg[(size_t)w]; return rc[(size_t)w]; }
    for (int w = 0; w < workers; ++w)
        if (rc[(size_t)w] > 0) { g_err = msg[(size_t)w]; return rc[(size_t)w]; }
    return RFSI_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------
// forest: import ranger layout -> device layout
// ------------------------------------------------------------------------------------
struct rfsi_forest {
    int32_t ntrees = 0, nfeat = 0;
    std::vector<Node> nodes;          // device layout, host copy
    std::vector<double> leaf_sample;  // per device node (empty when no quantile samples)
    std::vector<int32_t> leaf_rid;    // per device node: ranger's tree-local node id (terminalNodes)
    int max_depth = 0;
    int64_t n_internal = 0;
    std::vector<int64_t> tree_nodes;  // reachable nodes per tree (device layout, excl. the root slot)
    // compact format (common.cuh): same indices as `nodes`
    bool compact_ok = false;
    std::vector<uint64_t> qnodes;     // (w1 << 32) | w0
    std::vector<QRoot> qroots;
    std::vector<double> leaf_val;
    std::vector<double> uvals;
    std::vector<int32_t> uoff;
    std::vector<double> rpar;         // [F][4] bucket function of rank_of(): lo/2, scale, buckets - 1, table offset
    std::vector<int32_t> rtab;        // per feature [buckets+1]
    // The two block images are derived from the compact one on first use (forest_image / rfsi_forest_info), under `mu`:
    // only the image a process actually walks is built (the sector image of the bench forest is 2.8 GB).
    // blocked format (common.cuh)
    mutable int blocked_state = 0;            // 0 = not tried, 1 = built, -1 = does not fit
    mutable std::vector<uint32_t> blocks;     // 32 words per block, block 0 = all pass-through
    mutable std::vector<uint32_t> broot;      // [T]
    // sector format (sector.cuh)
    mutable int sector_state = 0;
    mutable std::vector<uint32_t> sect;       // 8 words per block, block 0 = idle
    mutable std::vector<uint32_t> sroot;      // [T]
    mutable bool sect_keys = false;           // word 5 of the leaf blocks holds the quantile keys
    // leaf tables of the block images: leaves renumbered left to right within each tree, so the
    // leaves under any subtree are contiguous (neighbouring pixels end in neighbouring leaves)
    mutable bool leaf_tables = false;
    mutable std::vector<uint32_t> leaf_no;    // per device node
    mutable std::vector<double> leaf_val_c, leaf_sample_c;
    mutable std::vector<int32_t> leaf_rid_c;
    struct Image {
        Node* nodes = nullptr;
        double* leaf_sample = nullptr;     // per-leaf samples (fp64 form of the quantile kernel), or
        uint32_t* leaf_rank = nullptr;     // their ranks in sample_dict (default form, NA = 0xFFFFFFFF)
        double* sample_dict = nullptr;     // sorted distinct sample values
        int32_t* leaf_rid = nullptr;
        bool compact = false;
        bool blocked = false;
        bool sector = false;
        bool direct_keys = false;          // sector image whose leaf blocks carry the quantile keys (no leaf table on the device)
        QForest q{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 32u, 0u};
    };
    mutable std::mutex mu;
    mutable std::map<int, Image> images;
    // host copy of the rank form of the sample table (of whichever image layout is in use), built on first upload
    mutable std::vector<double> sample_dict_h;
    mutable std::vector<uint32_t> leaf_rank_h;
};

namespace {

// bytes per leaf of the quantile sample table: dictionary ranks (default) or the fp64 samples
inline size_t sample_entry_bytes() { return env_int("RFSI_B200_QRF_RANKS", 1) != 0 ? sizeof(uint32_t) : sizeof(double); }

// order-preserving integer key of a double's bit pattern (a total order: -0.0 < +0.0)
inline uint64_t sample_key(double x) {
    int64_t b;
    memcpy(&b, &x, 8);
    return (uint64_t)(b ^ ((b >> 63) | (int64_t)0x8000000000000000LL));
}

// dict = the distinct non-NA samples in ascending order; rank[i] = position of ls[i] in dict
// (0xFFFFFFFF for NA: ranger leaves cells of random.node.values unset)
void sample_ranks(const std::vector<double>& ls, std::vector<double>& dict, std::vector<uint32_t>& rank) {
    std::vector<uint64_t> keys;
    keys.reserve(ls.size());
    for (double x : ls) if (x == x) keys.push_back(sample_key(x));
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    dict.resize(keys.size());
    for (size_t i = 0; i < keys.size(); ++i) {
        const int64_t b = (int64_t)keys[i];
        const int64_t bits = b ^ ((~b >> 63) | (int64_t)0x8000000000000000LL);
        memcpy(&dict[i], &bits, 8);
    }
    rank.resize(ls.size());
    const size_t n = ls.size();
    const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&, t] {
            for (size_t i = n * t / nt; i < n * (t + 1) / nt; ++i) {
                const double x = ls[i];
                rank[i] = (x == x) ? (uint32_t)(std::lower_bound(keys.begin(), keys.end(), sample_key(x)) - keys.begin())
                                   : 0xFFFFFFFFu;
            }
        });
    for (auto& x : th) x.join();
}


// ---- block images, derived from the compact one on first use (f->mu held) ------------------------------
inline uint32_t rank_at(const rfsi_forest* f, const Node& nd) {
    const double* u0 = f->uvals.data() + f->uoff[nd.feat];
    const double* u1 = f->uvals.data() + f->uoff[nd.feat + 1];
    return (uint32_t)(std::lower_bound(u0, u1, nd.val) - u0);
}

// leaf numbers: in-order (left to right) within each tree, trees one after the other
void ensure_leaf_tables(const rfsi_forest* f) {
    if (f->leaf_tables) return;
    const int F = f->nfeat, T = f->ntrees;
    const size_t n = f->nodes.size();
    const bool has_s = !f->leaf_sample.empty();
    f->leaf_no.assign(n, 0u);
    f->leaf_val_c.clear(); f->leaf_sample_c.clear(); f->leaf_rid_c.clear();
    f->leaf_val_c.reserve(n / 2 + T); f->leaf_rid_c.reserve(n / 2 + T);
    if (has_s) f->leaf_sample_c.reserve(n / 2 + T);
    std::vector<uint32_t> stack;
    for (int t = 0; t < T; ++t) {
        stack.assign(1, (uint32_t)t);
        while (!stack.empty()) {
            const uint32_t nd = stack.back();
            stack.pop_back();
            const Node& n0 = f->nodes[nd];
            if (n0.feat >= F) {
                f->leaf_no[nd] = (uint32_t)f->leaf_val_c.size();
                f->leaf_val_c.push_back(n0.val);
                f->leaf_rid_c.push_back(f->leaf_rid[nd]);
                if (has_s) f->leaf_sample_c.push_back(f->leaf_sample[nd]);
            } else {
                stack.push_back((uint32_t)n0.child + 1);   // right, popped after the whole left subtree
                stack.push_back((uint32_t)n0.child);
            }
        }
    }
    f->leaf_tables = true;
}

// blocked image: 4 levels per 128-byte block (common.cuh)
bool ensure_blocked(const rfsi_forest* f) {
    if (f->blocked_state != 0) return f->blocked_state > 0;
    f->blocked_state = -1;
    const int F = f->nfeat, T = f->ntrees;
    if (!f->compact_ok || F > kBMaxFeat) return false;
    for (int j = 0; j < F; ++j)
        if ((size_t)(f->uoff[j + 1] - f->uoff[j]) > (size_t)kBMaxRank) return false;
    ensure_leaf_tables(f);
    const std::vector<uint32_t>& leaf_no = f->leaf_no;
    std::vector<uint32_t>& B = f->blocks;
    B.assign(32, kBPassWord);   // block 0: pass-through everywhere; its exits are never used
    f->broot.assign((size_t)T, 0u);
    struct Pending { uint32_t node, block; };
    std::vector<Pending> todo;
    bool fits = true;
    // fill the heap of block `blk` below slot `pos` with the subtree rooted at wide node `nd`
    auto place = [&](auto&& self, uint32_t blk, uint32_t nd, uint32_t pos) -> void {
        const Node& n0 = f->nodes[nd];
        const bool leaf = n0.feat >= F;
        if (pos >= 16) {   // an exit
            if (leaf) { B[(size_t)blk * 32 + pos] = kBLeafFlag | leaf_no[nd]; return; }
            const size_t nb = B.size() / 32;
            if (nb >= 0x7fffffffu) { fits = false; return; }
            B.resize(B.size() + 32, kBPassWord);
            B[(size_t)blk * 32 + pos] = (uint32_t)nb;
            todo.push_back({nd, (uint32_t)nb});
            return;
        }
        if (leaf) {        // pad down to the exits
            B[(size_t)blk * 32 + pos] = kBPassWord;
            self(self, blk, nd, 2 * pos);
            self(self, blk, nd, 2 * pos + 1);
            return;
        }
        B[(size_t)blk * 32 + pos] = ~((rank_at(f, n0) << 8) | (uint32_t)n0.feat);
        self(self, blk, (uint32_t)n0.child, 2 * pos);
        self(self, blk, (uint32_t)n0.child + 1, 2 * pos + 1);
    };
    for (int t = 0; t < T && fits; ++t) {
        if (f->nodes[t].feat >= F) { f->broot[t] = kBLeafFlag | leaf_no[t]; continue; }
        const size_t nb = B.size() / 32;
        B.resize(B.size() + 32, kBPassWord);
        f->broot[t] = (uint32_t)nb;
        todo.clear();
        todo.push_back({(uint32_t)t, (uint32_t)nb});
        for (size_t i = 0; i < todo.size() && fits; ++i) {   // blocks of a tree in BFS order
            const Pending pd = todo[i];
            place(place, pd.block, pd.node, 1);
        }
    }
    if (!fits) { B.clear(); B.shrink_to_fit(); f->broot.clear(); return false; }
    f->blocked_state = 1;
    return true;
}

// sector image: 3 levels per 32-byte block, implicit children, leaves are blocks (sector.cuh).  Trees are
// independent: built on all host cores with tree-local block indices, then moved to their place.
bool ensure_sector(const rfsi_forest* f) {
    if (f->sector_state != 0) return f->sector_state > 0;
    f->sector_state = -1;
    const int F = f->nfeat, T = f->ntrees;
    if (!f->compact_ok || F > kSMaxFeat) return false;
    for (int j = 0; j < F; ++j)
        if ((size_t)(f->uoff[j + 1] - f->uoff[j]) > (size_t)kSMaxRank) return false;
    ensure_leaf_tables(f);
    auto node_at = [&](uint32_t i) {
        const Node& n0 = f->nodes[i];
        SectorNode sn{n0.feat >= F, 0u, (uint32_t)n0.child, f->leaf_no[i], n0.val};
        if (!sn.leaf) sn.word = sector_node_word(rank_at(f, n0), (uint32_t)n0.feat);
        return sn;
    };
    std::vector<std::vector<uint32_t>> part((size_t)T);
    {
        std::atomic<int> next{0};
        auto work = [&] { for (int t; (t = next.fetch_add(1)) < T;) sector_build_tree((uint32_t)t, node_at, part[(size_t)t]); };
        const unsigned hw = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
        std::vector<std::thread> th;
        for (unsigned i = 0; i + 1 < hw && (int)i + 1 < T; ++i) th.emplace_back(work);
        work();
        for (auto& x : th) x.join();
    }
    size_t total = 1;
    for (const auto& v : part) total += v.size() / 8;
    if (total >= (size_t)0x7fffffff) return false;
    std::vector<uint32_t>& W = f->sect;
    W.clear();
    W.resize(total * 8);
    W[0] = 0u;                                          // block 0: idle, its first child is itself
    for (int i = 1; i < 8; ++i) W[(size_t)i] = kSPassWord;
    f->sroot.assign((size_t)T, 0u);
    size_t at = 1;
    for (int t = 0; t < T; ++t) {
        std::vector<uint32_t>& v = part[(size_t)t];
        sector_relocate(v.data(), v.size() / 8, (uint32_t)at);
        memcpy(W.data() + at * 8, v.data(), v.size() * sizeof(uint32_t));
        f->sroot[(size_t)t] = sector_root_entry(v.data(), (uint32_t)at);
        at += v.size() / 8;
        std::vector<uint32_t>().swap(v);
    }
    f->sect_keys = false;
    f->sector_state = 1;
    return true;
}

// RFSI_B200_COMPACT selects the forest image: 0 = 16-byte fp64 nodes, 1 = 8-byte rank nodes, 2 = 4-byte rank nodes in
// 128-byte blocks of 4 levels, 3 (default) = 32-byte sector blocks of 3 levels; a forest an image cannot hold falls
// back to the next lower one.  Returns the image in use (and builds it).
int forest_image_code(const rfsi_forest* f) {
    const int want = env_int("RFSI_B200_COMPACT", 3);
    if (!f->compact_ok || want <= 0) return 0;
    if (want >= 3 && ensure_sector(f)) return 3;
    if (want >= 2 && ensure_blocked(f)) return 2;
    return 1;
}

int forest_image(const rfsi_forest* f, int device, rfsi_forest::Image* out) {
    std::lock_guard<std::mutex> lk(f->mu);
    auto it = f->images.find(device);
    if (it != f->images.end()) { *out = it->second; return RFSI_OK; }
    rfsi_forest::Image im;
    const int code = forest_image_code(f);
    im.compact = code >= 1;
    im.blocked = code == 2;
    im.sector = code == 3;
    const bool renumbered = code >= 2;     // leaf tables indexed by the in-order leaf number
    auto up = [&](const void* src, size_t bytes, const void** dst) -> int {
        void* d = nullptr;
        CU_TRY(cudaMalloc(&d, bytes ? bytes : 8));
        CU_TRY(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
        *dst = d;
        return RFSI_OK;
    };
    // quantile samples first: the sector image carries their dictionary ranks inside its leaf blocks
    const bool ranks = !f->leaf_sample.empty() && env_int("RFSI_B200_QRF_RANKS", 1) != 0;
    if (!f->leaf_sample.empty()) {
        const std::vector<double>& ls = renumbered ? f->leaf_sample_c : f->leaf_sample;
        if (ranks && ls.size() < 0xFFFFFFFFull) {
            // 4-byte form: the distinct sample values (at most one per training row) sorted in
            // the total order of their bit patterns (-0.0 < +0.0 stay distinct entries, so the
            // values that come back are the very doubles ranger stored), and each leaf's rank
            // built once per forest (f->mu is held), shared by the images of all devices
            std::vector<double>& dict = f->sample_dict_h;
            std::vector<uint32_t>& rank = f->leaf_rank_h;
            if (rank.size() != ls.size()) sample_ranks(ls, dict, rank);
            if (im.sector) {
                if (!f->sect_keys) {       // key of every leaf into word 5 of its block
                    uint32_t* W = f->sect.data();
                    const size_t nb = f->sect.size() / 8;
                    for (size_t b = 1; b < nb; ++b)
                        if (W[b * 8] & kSLeafFlag) W[b * 8 + kSWordKey] = rank[W[b * 8 + kSWordLeafNo]];
                    f->sect_keys = true;
                }
                im.direct_keys = true;     // no per-leaf table on the device at all
            } else {
                CU_TRY(cudaMalloc(&im.leaf_rank, std::max<size_t>(rank.size(), 1) * sizeof(uint32_t)));
                CU_TRY(cudaMemcpy(im.leaf_rank, rank.data(), rank.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            }
            CU_TRY(cudaMalloc(&im.sample_dict, std::max<size_t>(dict.size(), 1) * sizeof(double)));
            CU_TRY(cudaMemcpy(im.sample_dict, dict.data(), dict.size() * sizeof(double), cudaMemcpyHostToDevice));
        } else {
            CU_TRY(cudaMalloc(&im.leaf_sample, ls.size() * sizeof(double)));
            CU_TRY(cudaMemcpy(im.leaf_sample, ls.data(), ls.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
    }
    if (im.compact) {
        if (im.sector) {
            RF_TRY(up(f->sect.data(), f->sect.size() * sizeof(uint32_t), reinterpret_cast<const void**>(&im.q.sect)));
            RF_TRY(up(f->sroot.data(), f->sroot.size() * sizeof(uint32_t), reinterpret_cast<const void**>(&im.q.sroot)));
            im.q.rank_shift = kSRankShift;
        } else if (im.blocked) {
            RF_TRY(up(f->blocks.data(), f->blocks.size() * sizeof(uint32_t), reinterpret_cast<const void**>(&im.q.blocks)));
            RF_TRY(up(f->broot.data(), f->broot.size() * sizeof(uint32_t), reinterpret_cast<const void**>(&im.q.broot)));
            im.q.rank_shift = 8;
            RF_TRY(up(f->leaf_val_c.data(), f->leaf_val_c.size() * sizeof(double), reinterpret_cast<const void**>(&im.q.leaf_val)));
        } else {
            RF_TRY(up(f->qnodes.data(), f->qnodes.size() * sizeof(uint64_t), reinterpret_cast<const void**>(&im.q.nodes)));
            RF_TRY(up(f->qroots.data(), f->qroots.size() * sizeof(QRoot), reinterpret_cast<const void**>(&im.q.roots)));
            RF_TRY(up(f->leaf_val.data(), f->leaf_val.size() * sizeof(double), reinterpret_cast<const void**>(&im.q.leaf_val)));
        }
        RF_TRY(up(f->uvals.data(), f->uvals.size() * sizeof(double), reinterpret_cast<const void**>(&im.q.uvals)));
        RF_TRY(up(f->uoff.data(), f->uoff.size() * sizeof(int32_t), reinterpret_cast<const void**>(&im.q.uoff)));
        RF_TRY(up(f->rpar.data(), f->rpar.size() * sizeof(double), reinterpret_cast<const void**>(&im.q.rpar)));
        RF_TRY(up(f->rtab.data(), f->rtab.size() * sizeof(int32_t), reinterpret_cast<const void**>(&im.q.rtab)));
    } else {
        CU_TRY(cudaMalloc(&im.nodes, f->nodes.size() * sizeof(Node)));
        CU_TRY(cudaMemcpy(im.nodes, f->nodes.data(), f->nodes.size() * sizeof(Node), cudaMemcpyHostToDevice));
    }
    f->images[device] = im;
    *out = im;
    return RFSI_OK;
}

// the ranger-id table is only needed by predict(type="terminalNodes"): uploaded on first use
int forest_leaf_rid(const rfsi_forest* f, int device, const int32_t** out) {
    std::lock_guard<std::mutex> lk(f->mu);
    rfsi_forest::Image& im = f->images[device];
    if (!im.leaf_rid) {
        const std::vector<int32_t>& lr = (im.blocked || im.sector) ? f->leaf_rid_c : f->leaf_rid;
        CU_TRY(cudaMalloc(&im.leaf_rid, lr.size() * sizeof(int32_t)));
        CU_TRY(cudaMemcpy(im.leaf_rid, lr.data(), lr.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    *out = im.leaf_rid;
    return RFSI_OK;
}

}  // namespace

namespace {

// Derive the compact image (common.cuh) from the wide one.  Exact: ranks preserve order.
void build_compact(rfsi_forest* f) {
    const int F = f->nfeat, T = f->ntrees;
    const size_t n = f->nodes.size();
    f->compact_ok = false;
    if (F > kQMaxFeat) return;
    const int64_t roots = (T + 1) & ~int64_t(1);
    // sorted unique thresholds per feature
    std::vector<std::vector<double>> u((size_t)F);
    {
        std::vector<size_t> cnt((size_t)F, 0);
        for (size_t i = 0; i < n; ++i) if (f->nodes[i].feat < F) ++cnt[f->nodes[i].feat];
        for (int j = 0; j < F; ++j) u[j].reserve(cnt[j]);
        for (size_t i = 0; i < n; ++i) if (f->nodes[i].feat < F) u[f->nodes[i].feat].push_back(f->nodes[i].val);
    }
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    {
        std::atomic<int> nextf{0};
        auto work = [&] {
            for (int j; (j = nextf.fetch_add(1)) < F;) {
                std::sort(u[j].begin(), u[j].end());
                u[j].erase(std::unique(u[j].begin(), u[j].end()), u[j].end());
            }
        };
        std::vector<std::thread> th;
        for (unsigned t = 0; t + 1 < hw; ++t) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
    }
    f->uoff.assign((size_t)F + 1, 0);
    for (int j = 0; j < F; ++j) {
        if (u[j].size() >= 0xFFFFFFF0ull || f->uoff[j] + (int64_t)u[j].size() > 0x7fffffff) return;
        f->uoff[j + 1] = f->uoff[j] + (int32_t)u[j].size();
    }
    f->uvals.resize((size_t)f->uoff[F]);
    for (int j = 0; j < F; ++j) std::copy(u[j].begin(), u[j].end(), f->uvals.begin() + f->uoff[j]);
    // bucket tables for rank_of(): counts of thresholds per monotone bucket, prefix-summed
    // about one bucket per distinct threshold (a power of two): the search inside a bucket is then a step or two
    const int cap = std::max(kRankBucketsMin, std::min(kRankBucketsMax, env_int("RFSI_B200_RANK_BUCKETS", kRankBucketsMax)));
    f->rpar.assign((size_t)F * 4, 0.0);
    std::vector<size_t> toff((size_t)F + 1, 0);
    std::vector<int> nbk((size_t)F, kRankBucketsMin);
    for (int j = 0; j < F; ++j) {
        while (nbk[j] < cap && (size_t)nbk[j] < u[j].size()) nbk[j] *= 2;
        toff[j + 1] = toff[j] + (size_t)nbk[j] + 1;
    }
    f->rtab.assign(toff[F], 0);
    for (int j = 0; j < F; ++j) {
        int32_t* tab = f->rtab.data() + toff[j];
        const int nb = nbk[j];
        f->rpar[4 * j + 2] = (double)(nb - 1); f->rpar[4 * j + 3] = (double)toff[j];
        if (u[j].empty()) continue;
        const double lo_half = u[j].front() * 0.5;
        double scale = (double)nb / (u[j].back() * 0.5 - lo_half);
        if (!(scale < 1.0e300)) scale = 1.0e300;   // one threshold, or a tiny range: still monotone
        f->rpar[4 * j] = lo_half; f->rpar[4 * j + 1] = scale;
        for (double v : u[j]) ++tab[rank_bucket(v, lo_half, scale, (double)(nb - 1)) + 1];
        for (int b = 0; b < nb; ++b) tab[b + 1] += tab[b];
    }

    f->qnodes.assign(n, (uint64_t)0);                // leaf: {w0 = ~0xFFFFFFFF = 0, w1 = 0}
    f->leaf_val.assign(n, 0.0);
    f->qroots.assign((size_t)T, QRoot{~kQLeafRank, 0u, 0u, 0u});
    std::atomic<bool> ok{true};
    auto rank_at = [&](const Node& nd) -> uint32_t {
        const std::vector<double>& v = u[nd.feat];
        return (uint32_t)(std::lower_bound(v.begin(), v.end(), nd.val) - v.begin());
    };
    auto fill = [&](size_t a, size_t b) {
        for (size_t i = a; i < b; ++i) {
            const Node& nd = f->nodes[i];
            if (nd.feat >= F) { f->leaf_val[i] = nd.val; continue; }          // leaf: stays absorbing
            if ((int64_t)i < roots) continue;                                  // internal root: QRoot below
            const int64_t delta = (int64_t)nd.child - (int64_t)i;
            if (delta <= 0 || delta >= (int64_t)kQMaxDelta) { ok = false; continue; }
            const uint32_t w1 = ((uint32_t)nd.feat * 4u << kQDeltaBits) | (uint32_t)delta;
            f->qnodes[i] = ((uint64_t)w1 << 32) | (uint32_t)~rank_at(nd);
        }
    };
    {
        std::vector<std::thread> th;
        const size_t chunk = (n + hw - 1) / hw;
        for (unsigned t = 0; t < hw; ++t) {
            const size_t a = t * chunk, b = std::min(n, a + chunk);
            if (a < b) th.emplace_back(fill, a, b);
        }
        for (auto& t : th) t.join();
    }
    for (int t = 0; t < T; ++t) {
        const Node& nd = f->nodes[t];
        if (nd.feat >= F) f->qroots[t] = QRoot{~kQLeafRank, 0u, (uint32_t)t, 0u};   // single-leaf tree
        else f->qroots[t] = QRoot{~rank_at(nd), (uint32_t)nd.feat * 4u, (uint32_t)nd.child, 0u};
    }
    if (!ok) { f->qnodes.clear(); f->qroots.clear(); f->leaf_val.clear(); return; }
    f->compact_ok = true;
}

}  // namespace

extern "C" int rfsi_forest_create(int32_t ntrees, const int64_t* node_offsets, const int32_t* left,
                                  const int32_t* right, const int32_t* split_var, const double* split_val,
                                  const int64_t* rnv_offsets, const double* rnv, int32_t nfeat,
                                  rfsi_forest_t** out)
{
    if (!out) return fail(RFSI_E_ARG, "out is NULL");
    *out = nullptr;
    if (ntrees < 1 || nfeat < 1 || !node_offsets || !left || !right || !split_var || !split_val)
        return fail(RFSI_E_ARG, "forest_create: ntrees/nfeat must be >= 1 and arrays non-NULL");
    if ((rnv == nullptr) != (rnv_offsets == nullptr))
        return fail(RFSI_E_ARG, "forest_create: random_node_values and rnv_offsets go together");
    if (nfeat > 32767) return fail(RFSI_E_ARG, "forest_create: nfeat %d > 32767 unsupported", nfeat);

    auto* f = new rfsi_forest();
    f->ntrees = ntrees;
    f->nfeat = nfeat;
    const int64_t roots = (ntrees + 1) & ~int64_t(1);  // roots first, padded so pairs start even
    int64_t total = roots;
    for (int32_t t = 0; t < ntrees; ++t) {
        const int64_t n = node_offsets[t + 1] - node_offsets[t];
        if (n < 1) { delete f; return fail(RFSI_E_FOREST, "tree %d has no nodes", t); }
        total += n;  // upper bound: unreachable nodes are dropped
    }
    if (total >= (int64_t)0x7fffffff) { delete f; return fail(RFSI_E_FOREST, "forest has >= 2^31 nodes"); }
    // unused slots (the root padding) are harmless self-looping leaves
    f->nodes.assign((size_t)total, Node{0.0, 0, nfeat});
    f->leaf_rid.assign((size_t)total, 0);
    if (rnv) f->leaf_sample.assign((size_t)total, na_real());

    struct Item { int32_t rid; int32_t slot; int32_t depth; };
    std::vector<Item> queue;
    int64_t next = roots;
    f->tree_nodes.assign((size_t)ntrees, 0);
    for (int32_t t = 0; t < ntrees; ++t) {
        const int64_t o = node_offsets[t];
        const int64_t n = node_offsets[t + 1] - o;
        const int64_t next_at_start = next;
        queue.clear();
        queue.push_back({0, t, 0});
        size_t head = 0;
        while (head < queue.size()) {
            const Item it = queue[head++];
            if ((int64_t)queue.size() > n) {
                delete f;
                return fail(RFSI_E_FOREST, "tree %d: a node is reachable twice (cycle or shared child)", t);
            }
            const int32_t l = left[o + it.rid], r = right[o + it.rid];
            f->max_depth = std::max(f->max_depth, it.depth);
            if (l == 0 && r == 0) {
                const double pred = split_val[o + it.rid];
                if (pred != pred) {
                    delete f;
                    return fail(RFSI_E_FOREST, "tree %d node %d: terminal node prediction is NaN", t, it.rid);
                }
                f->nodes[it.slot] = Node{pred, it.slot, nfeat};  // absorbing leaf (see common.cuh)
                f->leaf_rid[it.slot] = it.rid;
                if (rnv) f->leaf_sample[it.slot] = rnv[rnv_offsets[t] + it.rid];
            } else {
                if (l <= 0 || l >= n || r <= 0 || r >= n) {
                    delete f;
                    return fail(RFSI_E_FOREST, "tree %d node %d: child id out of range", t, it.rid);
                }
                const int32_t v = split_var[o + it.rid];
                if (v < 0 || v >= nfeat) {
                    delete f;
                    return fail(RFSI_E_FOREST, "tree %d node %d: split variable %d outside [0,%d)", t, it.rid, v, nfeat);
                }
                f->nodes[it.slot] = Node{split_val[o + it.rid], (int32_t)next, v};
                queue.push_back({l, (int32_t)next, it.depth + 1});
                queue.push_back({r, (int32_t)next + 1, it.depth + 1});
                next += 2;
                ++f->n_internal;
            }
        }
        f->tree_nodes[t] = next - next_at_start;
    }
    f->nodes.resize((size_t)next);
    f->leaf_rid.resize((size_t)next);
    if (rnv) f->leaf_sample.resize((size_t)next);
    build_compact(f);
    *out = f;
    return RFSI_OK;
}

extern "C" void rfsi_forest_destroy(rfsi_forest_t* f) {
    if (!f) return;
    const bool forked = g_cuda_pid.load() != 0 && g_cuda_pid.load() != (long)getpid();   // the parent's device images are not ours
    for (auto& kv : f->images) {
        if (forked) break;
        int cur = 0;
        cudaGetDevice(&cur);
        if (cudaSetDevice(kv.first) == cudaSuccess) {
            if (kv.second.nodes) cudaFree(kv.second.nodes);
            if (kv.second.leaf_sample) cudaFree(kv.second.leaf_sample);
            if (kv.second.leaf_rank) cudaFree(kv.second.leaf_rank);
            if (kv.second.sample_dict) cudaFree(kv.second.sample_dict);
            if (kv.second.leaf_rid) cudaFree(kv.second.leaf_rid);
            cudaFree(const_cast<uint2*>(kv.second.q.nodes)); cudaFree(const_cast<QRoot*>(kv.second.q.roots));
            cudaFree(const_cast<double*>(kv.second.q.leaf_val)); cudaFree(const_cast<double*>(kv.second.q.uvals));
            cudaFree(const_cast<int32_t*>(kv.second.q.uoff));
            cudaFree(const_cast<double*>(kv.second.q.rpar)); cudaFree(const_cast<int32_t*>(kv.second.q.rtab));
            cudaFree(const_cast<uint32_t*>(kv.second.q.blocks)); cudaFree(const_cast<uint32_t*>(kv.second.q.broot));
            cudaFree(const_cast<uint32_t*>(kv.second.q.sect)); cudaFree(const_cast<uint32_t*>(kv.second.q.sroot));
        }
        cudaSetDevice(cur);
    }
    delete f;
}

extern "C" int rfsi_forest_upload(const rfsi_forest_t* f, int device) {
    if (!f) return fail(RFSI_E_ARG, "forest is NULL");
    RF_TRY(use_device(device));
    rfsi_forest::Image im;
    return forest_image(f, device, &im);
}

extern "C" int64_t rfsi_forest_info(const rfsi_forest_t* f, int what) {
    if (!f) return -1;
    switch (what) {
        case 0: return f->ntrees;
        case 1: return f->nfeat;
        case 2: return (int64_t)f->nodes.size();
        case 3: return f->max_depth;
        case 4: return f->leaf_sample.empty() ? 0 : 1;
        case 5: case 7: case 9: {
            std::lock_guard<std::mutex> lk(f->mu);
            const int code = forest_image_code(f);
            if (what == 7) return code;
            if (what == 9) return code == 3 ? (int64_t)f->sect.size() / 8 : (code == 2 ? (int64_t)f->blocks.size() / 32 : 0);
            const int64_t tabs = (int64_t)(f->uvals.size() * 8 + f->rtab.size() * 4);
            const bool ranks = env_int("RFSI_B200_QRF_RANKS", 1) != 0;
            if (code == 3)   // values, numbers and quantile keys live in the leaf blocks; the fp64 form keeps its table
                return (int64_t)(f->sect.size() * 4) + tabs + (ranks ? 0 : (int64_t)f->leaf_sample_c.size() * 8);
            if (code == 2)
                return (int64_t)(f->blocks.size() * 4 + f->leaf_val_c.size() * 8) + tabs +
                       (int64_t)(f->leaf_sample_c.size() * sample_entry_bytes());
            if (code == 1)
                return (int64_t)(f->qnodes.size() * 8 + f->leaf_val.size() * 8 + f->qroots.size() * sizeof(QRoot)) + tabs +
                       (int64_t)(f->leaf_sample.size() * sample_entry_bytes());
            return (int64_t)(f->nodes.size() * sizeof(Node) + f->leaf_sample.size() * sample_entry_bytes());
        }
        case 6: return f->n_internal;
        case 8: return (int64_t)f->uvals.size();
        default: return -1;
    }
}

// ------------------------------------------------------------------------------------
// misc entry points
// ------------------------------------------------------------------------------------
extern "C" int rfsi_version(void) { return RFSI_B200_VERSION; }
extern "C" const char* rfsi_last_error(void) { return g_err.c_str(); }
extern "C" uint64_t rfsi_launch_count(void) { return g_launches.load(); }

extern "C" int rfsi_profile_enable(int on) {
    g_prof_on.store(on ? 1 : 0);
    return RFSI_OK;
}

extern "C" int rfsi_profile_read(int kernel_class, double* total_ms, uint64_t* launches) {
    if (kernel_class < 0 || kernel_class >= PROF_NCLASS) return fail(RFSI_E_ARG, "profile_read: class %d", kernel_class);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double ms = 0.0;
    uint64_t n = 0;
    std::vector<ProfRec> keep;
    for (auto& r : g_prof) {
        if (r.cls != kernel_class) { keep.push_back(r); continue; }
        float e = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&e, r.a, r.b) == cudaSuccess) { ms += e; ++n; }
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    g_prof.swap(keep);
    if (total_ms) *total_ms = ms;
    if (launches) *launches = n;
    return RFSI_OK;
}

extern "C" int rfsi_device_count(void) {
    RF_TRY(check_fork());
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

extern "C" int rfsi_host_alloc(void** out, size_t bytes) {
    if (!out) return fail(RFSI_E_ARG, "out is NULL");
    RF_TRY(check_fork());
    CU_TRY(cudaMallocHost(out, bytes ? bytes : 1));
    return RFSI_OK;
}
extern "C" void rfsi_host_free(void* p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------------------------
// launches
// ------------------------------------------------------------------------------------
namespace {

__global__ void interleave_xy_kernel(const double* __restrict__ x, const double* __restrict__ y, int n,
                                     double2* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_double2(x[i], y[i]);
}

constexpr int kKnnP = 128;

template <int MODE>
int launch_knn(const LocSpec& L, int64_t npix, const double2* st_xy, int S, int k, const double* st_z,
               int rm_dupl, double* out_a, double* out_b, int32_t* out_i, int64_t ld, int* warn, int* na_flag,
               cudaStream_t stream)
{
    if (npix == 0) return RFSI_OK;
    const size_t smem = knn_smem_bytes<kKnnP>(k);
    if (smem > 227 * 1024) return fail(RFSI_E_ARG, "n.obs = %d needs %zu B of shared memory (> 227 KB)", k - 1, smem);
    auto kern = knn_brute_kernel<kKnnP, MODE>;
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((npix + kKnnP - 1) / kKnnP);
    {
        ProfScope ps(MODE == KNN_RAW ? PROF_KNN_RAW : PROF_KNN_FINAL, stream);
        kern<<<grid, kKnnP, smem, stream>>>(L, npix, st_xy, S, k, st_z, rm_dupl, out_a, out_b, out_i, ld, warn, na_flag);
    }
    count_launch();
    CU_TRY(cudaGetLastError());
    return RFSI_OK;
}

// ---- cell-grid search for large station sets ---------------------------------------------
struct CellsWs {
    CellHeader* hdr = nullptr;
    int32_t* cell_start = nullptr;
    int32_t* cursor = nullptr;
    double2* sxy = nullptr;
    int32_t* sid = nullptr;
    float2* sxyf = nullptr;   // the sorted coordinates as floats (read by the fp32-mode prefilter only)
};

// Morton-order permutation of unordered explicit locations (knn_cells.cuh)
size_t loc_perm_bytes(int64_t npix) {
    if (npix <= 0) return 0;
    size_t temp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, (int)npix, 0, 42);
    return 256 + 2 * align_up((size_t)npix * 8, 256) + 2 * align_up((size_t)npix * 4, 256) + align_up(temp, 256);
}
// returns the permutation (inside `scratch`) or nullptr when npix does not fit an int
int build_loc_perm(const LocSpec& L, int64_t npix, void* scratch, const int32_t** perm_out, cudaStream_t st) {
    *perm_out = nullptr;
    if (npix <= 0 || npix > 0x7fffffff || !scratch) return RFSI_OK;
    char* p = reinterpret_cast<char*>(scratch);
    unsigned long long* bbox = reinterpret_cast<unsigned long long*>(p); p += 256;
    unsigned long long* k_in = reinterpret_cast<unsigned long long*>(p); p += align_up((size_t)npix * 8, 256);
    unsigned long long* k_out = reinterpret_cast<unsigned long long*>(p); p += align_up((size_t)npix * 8, 256);
    int32_t* v_in = reinterpret_cast<int32_t*>(p); p += align_up((size_t)npix * 4, 256);
    int32_t* v_out = reinterpret_cast<int32_t*>(p); p += align_up((size_t)npix * 4, 256);
    const unsigned long long init[4] = {~0ULL, 0ULL, ~0ULL, 0ULL};
    CU_TRY(cudaMemcpyAsync(bbox, init, sizeof init, cudaMemcpyHostToDevice, st));
    const unsigned g = (unsigned)std::min<int64_t>((npix + 255) / 256, 148 * 8);
    loc_bbox_kernel<<<g, 256, 0, st>>>(L, npix, bbox);
    loc_morton_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(L, npix, bbox, k_in, v_in);
    size_t temp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp, k_in, k_out, v_in, v_out, (int)npix, 0, 42, st);
    CU_TRY(cub::DeviceRadixSort::SortPairs(p, temp, k_in, k_out, v_in, v_out, (int)npix, 0, 42, st));
    g_launches.fetch_add(4, std::memory_order_relaxed);
    CU_TRY(cudaGetLastError());
    *perm_out = v_out;
    return RFSI_OK;
}

bool loc_rows(const LocSpec& L, int64_t npix) {   // thread -> pixel mapping can use compact 2-D tiles
    return L.ncol >= 16 && L.pix_begin % L.ncol == 0 && (!L.x || npix % L.ncol == 0);
}
int cells_G(int S) { return std::max(1, std::min(1024, (int)std::ceil(std::sqrt((double)std::max(S, 1) / 3.0)))); }
size_t cells_ws_bytes(int S) {
    const size_t G2 = (size_t)cells_G(S) * cells_G(S);
    return 256 + align_up((G2 + 1) * 4, 256) + align_up(G2 * 4, 256) + align_up((size_t)std::max(S, 1) * 16, 256) +
           align_up((size_t)std::max(S, 1) * 4, 256) + align_up((size_t)std::max(S, 1) * 8, 256);
}
CellsWs carve_cells(void* ws, int S) {
    const size_t G2 = (size_t)cells_G(S) * cells_G(S);
    char* p = reinterpret_cast<char*>(ws);
    CellsWs c;
    c.hdr = reinterpret_cast<CellHeader*>(p); p += 256;
    c.cell_start = reinterpret_cast<int32_t*>(p); p += align_up((G2 + 1) * 4, 256);
    c.cursor = reinterpret_cast<int32_t*>(p); p += align_up(G2 * 4, 256);
    c.sxy = reinterpret_cast<double2*>(p); p += align_up((size_t)std::max(S, 1) * 16, 256);
    c.sid = reinterpret_cast<int32_t*>(p); p += align_up((size_t)std::max(S, 1) * 4, 256);
    c.sxyf = reinterpret_cast<float2*>(p);
    return c;
}
int cells_min_S() { return env_int("RFSI_B200_CELLS_MIN", 128); }  // measured: the cell-grid search wins from ~100 stations up
int build_cells(const double2* xy, int S, const CellsWs& c, cudaStream_t st) {
    build_cells_kernel<<<1, 1024, 0, st>>>(xy, S, cells_G(S), c.hdr, c.cell_start, c.cursor, c.sxy, c.sid, c.sxyf);
    count_launch();
    CU_TRY(cudaGetLastError());
    return RFSI_OK;
}

// perm: Morton permutation for unordered explicit locations (build_loc_perm), or nullptr
// f32_mode: every coordinate is a float (rfsi_near_obs_f32): fp32 prefilter, exact fp64 re-rank (knn_cells.cuh)
template <int MODE>
int launch_knn_cells(const LocSpec& L, int64_t npix, const CellsWs& c, const int32_t* perm, int k, const double* st_z,
                     int rm_dupl, double* out_a, double* out_b, int32_t* out_i, int64_t ld, int* warn, int* na_flag,
                     cudaStream_t stream, bool f32_mode = false)
{
    if (npix == 0) return RFSI_OK;
    const size_t smem = knn_cells_smem_bytes<kKnnP>(k);
    if (smem > 227 * 1024) return fail(RFSI_E_ARG, "n.obs = %d needs %zu B of shared memory (> 227 KB)", k - 1, smem);
    auto kern = f32_mode ? knn_cells_kernel<kKnnP, MODE, true> : knn_cells_kernel<kKnnP, MODE, false>;
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // 2-D tiles need rows of L.ncol locations: the raster form, or explicit coordinates stored row by row
    const int tile2d = loc_rows(L, npix) ? 1 : 0;
    if (tile2d) perm = nullptr;
    long long tiles_per_row = 0, ntiles;
    if (tile2d) {
        tiles_per_row = (L.ncol + 15) / 16;
        const long long nrows = (npix + L.ncol - 1) / L.ncol;
        ntiles = tiles_per_row * ((nrows + (kKnnP / 16) - 1) / (kKnnP / 16));
    } else {
        ntiles = (npix + kKnnP - 1) / kKnnP;
    }
    {
        ProfScope ps(MODE == KNN_RAW ? PROF_KNN_RAW : PROF_KNN_FINAL, stream);
        kern<<<(unsigned)ntiles, kKnnP, smem, stream>>>(L, npix, c.hdr, c.cell_start, c.sxy, c.sxyf, c.sid, k, st_z, rm_dupl,
                                                         out_a, out_b, out_i, ld, warn, na_flag, tile2d, tiles_per_row,
                                                         perm);
    }
    count_launch();
    CU_TRY(cudaGetLastError());
    return RFSI_OK;
}

// trees in flight per thread: measured best 12 for the 16-byte format, 8 for the compact one
thread_local int g_default_J = 12;
int forest_J() {
    const int j = env_int("RFSI_B200_J", 0);
    return j > 0 ? j : g_default_J;
}

// Tree ranges whose nodes total about RFSI_B200_CHUNK_MB each: one launch per range keeps
// the part of the forest being walked resident in L2 (126 MB) while all CTAs sweep it.
std::vector<int> tree_chunks(const rfsi_forest* f) {
    static int mb = env_int("RFSI_B200_CHUNK_MB", 0);  // 0 = one launch (chunking measured slower)
    std::vector<int> b{0};
    if (mb <= 0) { b.push_back(f->ntrees); return b; }
    const int64_t cap = (int64_t)mb * (1 << 20) / (int64_t)sizeof(Node);
    int64_t cur = 0;
    for (int t = 0; t < f->ntrees; ++t) {
        if (cur > 0 && cur + f->tree_nodes[t] > cap) { b.push_back(t); cur = 0; }
        cur += f->tree_nodes[t];
    }
    b.push_back(f->ntrees);
    return b;
}

template <int P, int J>
int launch_forest_xq_pj(const rfsi_forest* f, const QForest& q_in, const double* X, int64_t npix, int64_t ldx,
                        const ForestOut& o, unsigned long long* visits, int* na_flag, cudaStream_t stream)
{
    const int T = f->ntrees, F = f->nfeat;
    QForest q = q_in;
    q.sect_tile_stride = (uint32_t)(P * sizeof(uint32_t));
    const size_t smem = forest_q_smem_bytes<P>(F);
    const unsigned grid = (unsigned)((npix + P - 1) / P);
    const std::vector<int> cb = tree_chunks(f);
    for (size_t c = 0; c + 1 < cb.size(); ++c) {
        ProfScope ps(PROF_FOREST_X, stream);
        if (visits) {
            auto kern = forest_xq_kernel<P, J, true>;
            CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, P, smem, stream>>>(q, cb[c], cb[c + 1], T, F, X, npix, ldx, o, visits, na_flag);
        } else {
            auto kern = forest_xq_kernel<P, J, false>;
            CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, P, smem, stream>>>(q, cb[c], cb[c + 1], T, F, X, npix, ldx, o, visits, na_flag);
        }
        count_launch();
        CU_TRY(cudaGetLastError());
    }
    return RFSI_OK;
}

template <int P, int J>
int launch_forest_x_pj(const rfsi_forest* f, const Node* nodes, const double* X, int64_t npix, int64_t ldx,
                       const ForestOut& o, unsigned long long* visits, int* na_flag, cudaStream_t stream)
{
    const int T = f->ntrees, F = f->nfeat;
    const size_t smem = forest_smem_bytes<P>(F);
    const unsigned grid = (unsigned)((npix + P - 1) / P);
    // the running sum lives in o.mean; without it (terminal nodes / samples only) chunks are independent
    const std::vector<int> cb = tree_chunks(f);
    for (size_t c = 0; c + 1 < cb.size(); ++c) {
        ProfScope ps(PROF_FOREST_X, stream);
        if (visits) {
            auto kern = forest_x_kernel<P, J, true>;
            CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, P, smem, stream>>>(nodes, cb[c], cb[c + 1], T, F, X, npix, ldx, o, visits, na_flag);
        } else {
            auto kern = forest_x_kernel<P, J, false>;
            CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, P, smem, stream>>>(nodes, cb[c], cb[c + 1], T, F, X, npix, ldx, o, visits, na_flag);
        }
        count_launch();
        CU_TRY(cudaGetLastError());
    }
    return RFSI_OK;
}

#define RFSI_DISPATCH_J(CALL_PREFIX, P, ...)                 \
    do {                                                     \
        const int J__ = forest_J();                          \
        if (J__ >= 16) return CALL_PREFIX<P, 16>(__VA_ARGS__); \
        if (J__ >= 12) return CALL_PREFIX<P, 12>(__VA_ARGS__); \
        if (J__ >= 8) return CALL_PREFIX<P, 8>(__VA_ARGS__);   \
        if (J__ >= 4) return CALL_PREFIX<P, 4>(__VA_ARGS__);   \
        return CALL_PREFIX<P, 2>(__VA_ARGS__);                 \
    } while (0)

int launch_forest_x(const rfsi_forest* f, const rfsi_forest::Image& im, const double* X, int64_t npix, int64_t ldx,
                    const ForestOut& o, unsigned long long* visits, int* na_flag, cudaStream_t stream)
{
    if (npix == 0) return RFSI_OK;
    const int F = f->nfeat;
    const Node* nodes = im.nodes;
    g_default_J = im.compact ? 8 : 12;
    if (im.compact) {
        if (forest_q_smem_bytes<256>(F) <= 110 * 1024)
            RFSI_DISPATCH_J(launch_forest_xq_pj, 256, f, im.q, X, npix, ldx, o, visits, na_flag, stream);
        if (forest_q_smem_bytes<128>(F) > 227 * 1024)
            return fail(RFSI_E_ARG, "forest with %d features exceeds the shared-memory feature tile", F);
        RFSI_DISPATCH_J(launch_forest_xq_pj, 128, f, im.q, X, npix, ldx, o, visits, na_flag, stream);
    }
    const bool big = forest_smem_bytes<256>(F) <= 100 * 1024;
    if (!big && forest_smem_bytes<128>(F) > 227 * 1024)
        return fail(RFSI_E_ARG, "forest with %d features exceeds the shared-memory feature tile (max %d)", F,
                    227 * 1024 / (128 * 8));
    if (big) RFSI_DISPATCH_J(launch_forest_x_pj, 256, f, nodes, X, npix, ldx, o, visits, na_flag, stream);
    RFSI_DISPATCH_J(launch_forest_x_pj, 128, f, nodes, X, npix, ldx, o, visits, na_flag, stream);
}

constexpr int kMaxProbs = 128;

// row stride (entries) of the leaf-id scratch: whole groups of 8 trees are 32-byte aligned
inline int64_t leaf_ids_ld(int T) { return ((int64_t)T + 7) & ~int64_t(7); }
inline size_t leaf_ids_bytes(int64_t rows, int T) { return sizeof(uint32_t) * (size_t)rows * (size_t)leaf_ids_ld(T); }

template <int R, class K>
int launch_qrf_r(const uint32_t* leaf_ids, const K* leaf_key, const double* dict, const unsigned char* skip, int64_t nrows,
                 int T, const double* probs_dev, int nprobs, double* out_q, int64_t ld_q, int16_t* out_iqr,
                 cudaStream_t stream)
{
    auto kern = qrf_quantile_kernel<R, K>;
    const size_t smem = qrf_smem_bytes<R, K>();
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0, per_sm = 0;
    CU_TRY(cudaGetDevice(&dev));
    CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * kQrfWarps, smem));
    // persistent warps: one resident wave, each warp strides over the rows
    const int64_t want = (nrows + kQrfWarps - 1) / kQrfWarps;
    const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)sms * std::max(per_sm, 1));
    kern<<<grid, 32 * kQrfWarps, smem, stream>>>(leaf_ids, leaf_ids_ld(T), leaf_key, dict, skip, nrows, T, probs_dev, nprobs,
                                                 out_q, ld_q, out_iqr);
    return RFSI_OK;
}

// direct form (qrf.cuh): the scratch rows hold the keys, nothing to gather
template <int R>
int launch_qrf_direct(const uint32_t* keys, const double* dict, const unsigned char* skip, int64_t nrows, int T,
                      const double* probs_dev, int nprobs, double* out_q, int64_t ld_q, int16_t* out_iqr, cudaStream_t stream)
{
    auto kern = qrf_direct_kernel<R>;
    const size_t smem = qrf_direct_smem_bytes<R>();
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0, per_sm = 0;
    CU_TRY(cudaGetDevice(&dev));
    CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * kQrfWarps, smem));
    const int64_t want = (nrows + kQrfWarps - 1) / kQrfWarps;   // persistent warps: one resident wave, a contiguous run of rows per CTA
    const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)sms * std::max(per_sm, 1));
    static const int fast = env_int("RFSI_B200_QRF_FAST", 1);   // 0: the re-bucketing selection for every row (A/B switch, tests)
    kern<<<grid, 32 * kQrfWarps, smem, stream>>>(keys, leaf_ids_ld(T), dict, skip, nrows, T, probs_dev, nprobs, out_q, ld_q, out_iqr,
                                                 fast);
    return RFSI_OK;
}

// the leaf table of an image in whichever form it was uploaded (ranks by default)
struct LeafSamples {
    const double* sample = nullptr;
    const uint32_t* rank = nullptr;
    const double* dict = nullptr;
    bool direct = false;      // the walk stored the keys themselves (sector image): nothing to gather
    explicit operator bool() const { return sample || rank || direct; }
};
inline LeafSamples leaf_samples(const rfsi_forest::Image& im) {
    return LeafSamples{im.leaf_sample, im.leaf_rank, im.sample_dict, im.direct_keys};
}
// the word of a sector leaf block the walk stores for the quantile pass
inline int leaf_ids_word(const rfsi_forest::Image& im) { return im.direct_keys ? kSWordKey : kSWordLeafNo; }

int launch_qrf(const uint32_t* leaf_ids, const LeafSamples& ls, const unsigned char* skip, int64_t nrows, int T,
               const double* probs_dev, int nprobs, double* out_q, int64_t ld_q, int16_t* out_iqr, cudaStream_t stream)
{
    if (nrows == 0) return RFSI_OK;
    const int R = (T + 31) / 32;
    ProfScope ps(PROF_QRF, stream);
#define RFSI_QRF_CASE(RR)                                                                                              \
    do {                                                                                                               \
        if (ls.direct) RF_TRY((launch_qrf_direct<RR>(leaf_ids, ls.dict, skip, nrows, T, probs_dev, nprobs, out_q, ld_q,   \
                                                     out_iqr, stream)));                                               \
        else if (ls.rank) RF_TRY((launch_qrf_r<RR, uint32_t>(leaf_ids, ls.rank, ls.dict, skip, nrows, T, probs_dev, nprobs, \
                                                             out_q, ld_q, out_iqr, stream)));                          \
        else RF_TRY((launch_qrf_r<RR, double>(leaf_ids, ls.sample, nullptr, skip, nrows, T, probs_dev, nprobs, out_q,   \
                                              ld_q, out_iqr, stream)));                                                \
    } while (0)
    if (R <= 1) RFSI_QRF_CASE(1);
    else if (R <= 2) RFSI_QRF_CASE(2);
    else if (R <= 4) RFSI_QRF_CASE(4);
    else if (R <= 8) RFSI_QRF_CASE(8);
    else if (R <= 16) RFSI_QRF_CASE(16);
    else if (R <= 32) RFSI_QRF_CASE(32);
    else return fail(RFSI_E_ARG, "quantile forests with more than 1024 trees are not supported (T=%d)", T);
#undef RFSI_QRF_CASE
    count_launch();
    CU_TRY(cudaGetLastError());
    return RFSI_OK;
}

int check_probs(const double* probs, int nprobs) {
    if (nprobs < 1 || nprobs > kMaxProbs || !probs) return fail(RFSI_E_ARG, "need 1..%d quantiles", kMaxProbs);
    for (int i = 0; i < nprobs; ++i)
        if (!(probs[i] >= 0.0 && probs[i] <= 1.0)) return fail(RFSI_E_ARG, "quantile %g outside [0,1]", probs[i]);
    return RFSI_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------
// stage 1 API
// ------------------------------------------------------------------------------------
extern "C" int rfsi_dev_near_obs_f64(const double* loc_x, const double* loc_y, int64_t npix,
                                     const double* obs_x, const double* obs_y, const double* obs_z,
                                     int32_t nobs, int32_t n_obs, int32_t rm_dupl, double* out_dist,
                                     double* out_obs, int32_t* out_idx, int32_t* warn_flag, int device,
                                     void* stream)
{
    if (npix < 0 || nobs < 0 || n_obs < 1) return fail(RFSI_E_ARG, "near_obs: bad sizes");
    RF_TRY(use_device(device));
    cudaStream_t st = (cudaStream_t)stream;
    // interleave the station coordinates so one 128-bit broadcast feeds a distance
    double2* xy = nullptr;
    void* cw = nullptr;
    void* pw = nullptr;
    const bool use_cells = nobs >= cells_min_S();
    LocSpec L{loc_x, loc_y, 0, 0, 0, 0, 1, 0};
    cudaError_t e = cudaMallocAsync(&xy, sizeof(double2) * (size_t)std::max(nobs, 1), st);
    if (e == cudaSuccess && use_cells) e = cudaMallocAsync(&cw, cells_ws_bytes(nobs), st);
    if (e == cudaSuccess && use_cells && npix > 0) e = cudaMallocAsync(&pw, loc_perm_bytes(npix), st);
    if (e != cudaSuccess) {
        if (xy) cudaFreeAsync(xy, st);
        if (cw) cudaFreeAsync(cw, st);
        return fail(e == cudaErrorMemoryAllocation ? RFSI_E_OOM : RFSI_E_CUDA, "near_obs: device scratch: %s",
                    cudaGetErrorString(e));
    }
    if (nobs > 0) {
        interleave_xy_kernel<<<(nobs + 255) / 256, 256, 0, st>>>(obs_x, obs_y, nobs, xy);
        count_launch();
    }
    int rc;
    if (use_cells) {
        const CellsWs c = carve_cells(cw, nobs);
        const int32_t* perm = nullptr;
        rc = build_cells(xy, nobs, c, st);
        if (rc >= 0) rc = build_loc_perm(L, npix, pw, &perm, st);
        if (rc >= 0)
            rc = launch_knn_cells<KNN_FINAL>(L, npix, c, perm, n_obs + 1, obs_z, rm_dupl, out_dist, out_obs, out_idx,
                                             npix, warn_flag, warn_flag, st);
        cudaFreeAsync(cw, st);
        if (pw) cudaFreeAsync(pw, st);
    } else {
        rc = launch_knn<KNN_FINAL>(L, npix, xy, nobs, n_obs + 1, obs_z, rm_dupl, out_dist, out_obs, out_idx, npix,
                                   warn_flag, warn_flag, st);
    }
    cudaFreeAsync(xy, st);   // stream-ordered: released after the kernels above have run
    return rc;
}

namespace {

__global__ void f32_to_f64_kernel(const float* __restrict__ in, double* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)in[i];
}
__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];   // NA_real_ stays a NaN
}

// near.obs with host buffers; column j of dist / obs / idx goes to dist_cols[j] / obs_cols[j] / idx_cols[j] (nullable array).
// T = double: the R call.  T = float: fp32 mode -- float buffers cross PCIe in both directions (half the bytes of the
// host-bound part of the call), coordinates are widened on the device (exactly), the search runs with the fp32 prefilter
// and the exact fp64 re-rank, results are rounded to float on the device.
template <class T>
int host_near_obs(const T* loc_x, const T* loc_y, int64_t npix, const T* obs_x, const T* obs_y,
                  const T* obs_z, int32_t nobs, int32_t n_obs, int32_t rm_dupl, T* const* dist_cols,
                  T* const* obs_cols, int32_t* const* idx_cols, int device)
{
    constexpr bool F32 = std::is_same<T, float>::value;
    static const bool prefilter = env_int("RFSI_B200_KNN_F32_PREFILTER", 1) != 0;   // A/B switch (profiles/r02_i)
    for (int32_t s = 0; s < nobs; ++s)
        if (obs_x[s] != obs_x[s] || obs_y[s] != obs_y[s])
            return fail(RFSI_E_NA_INPUT, "near_obs: observation %d has an NA coordinate", s + 1);
    RF_TRY(use_device(device));
    if (npix == 0) return RFSI_OK;
    Trace tr;

    std::vector<double2> xy((size_t)std::max(nobs, 1));
    std::vector<double> zz((size_t)std::max(nobs, 1));
    for (int32_t s = 0; s < nobs; ++s) { xy[s] = make_double2((double)obs_x[s], (double)obs_y[s]); zz[s] = (double)obs_z[s]; }
    DevBuf d_xy, d_z, d_flags, d_cells;
    RF_TRY(d_xy.alloc(sizeof(double2) * xy.size()));
    RF_TRY(d_z.alloc(sizeof(double) * (size_t)std::max(nobs, 1)));
    RF_TRY(d_flags.alloc(2 * sizeof(int)));
    CU_TRY(cudaMemcpy(d_xy.p, xy.data(), sizeof(double2) * xy.size(), cudaMemcpyHostToDevice));
    if (nobs > 0) CU_TRY(cudaMemcpy(d_z.p, zz.data(), sizeof(double) * nobs, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemset(d_flags.p, 0, 2 * sizeof(int)));
    const bool use_cells = nobs >= cells_min_S();
    CellsWs cells;
    if (use_cells) {
        RF_TRY(d_cells.alloc(cells_ws_bytes(nobs)));
        cells = carve_cells(d_cells.p, nobs);
        RF_TRY(build_cells(d_xy.as<double2>(), nobs, cells, nullptr));
        CU_TRY(cudaStreamSynchronize(nullptr));
    }
    // chunks of ~32 k locations: worker w takes chunks w, w + W, ...; each worker has its own stream, device
    // buffers and pinned slot, and scatters its finished chunk into the caller's columns itself
    const int64_t chunk = std::min<int64_t>(npix, std::max<int64_t>(4096, env_int("RFSI_B200_NEAR_CHUNK", 32768)));
    const int64_t nchunks = (npix + chunk - 1) / chunk;
    const int W = host_workers(nchunks);
    const size_t row_bytes = (size_t)n_obs * (2 * sizeof(T) + (idx_cols ? sizeof(int32_t) : 0));
    struct Lane {
        DevBuf lx, ly, dist, obs, idx, perm, fin, fout;   // fin / fout: the float side of fp32 mode
        PinnedBuf stage;
        Stream st;   // last: destroyed (synchronised) before the buffers above
    };
    std::vector<Lane> lane((size_t)W);
    for (int w = 0; w < W; ++w) {
        Lane& ln = lane[(size_t)w];
        RF_TRY(ln.st.create());
        RF_TRY(ln.lx.alloc(sizeof(double) * chunk));
        RF_TRY(ln.ly.alloc(sizeof(double) * chunk));
        RF_TRY(ln.dist.alloc(sizeof(double) * chunk * n_obs));
        RF_TRY(ln.obs.alloc(sizeof(double) * chunk * n_obs));
        if (idx_cols) RF_TRY(ln.idx.alloc(sizeof(int32_t) * chunk * n_obs));
        if (use_cells) RF_TRY(ln.perm.alloc(loc_perm_bytes(chunk)));
        if (F32) { RF_TRY(ln.fin.alloc(sizeof(float) * 2 * chunk)); RF_TRY(ln.fout.alloc(sizeof(float) * 2 * chunk * n_obs)); }
        RF_TRY(ln.stage.alloc(row_bytes * (size_t)chunk));
    }
    RF_TRY(pool_ready());
    tr.mark("near.obs setup");
    const int rc = run_workers(W, device, [&](int w) -> int {
        Lane& ln = lane[(size_t)w];
        cudaStream_t s = ln.st.s;
        for (int64_t c = w; c < nchunks; c += W) {
            const int64_t p0 = c * chunk, cur = std::min(chunk, npix - p0);
            if (F32) {
                CU_TRY(cudaMemcpyAsync(ln.fin.p, loc_x + p0, sizeof(float) * cur, cudaMemcpyHostToDevice, s));
                CU_TRY(cudaMemcpyAsync(ln.fin.as<float>() + cur, loc_y + p0, sizeof(float) * cur, cudaMemcpyHostToDevice, s));
                f32_to_f64_kernel<<<(unsigned)((cur + 255) / 256), 256, 0, s>>>(ln.fin.as<float>(), ln.lx.as<double>(), cur);
                f32_to_f64_kernel<<<(unsigned)((cur + 255) / 256), 256, 0, s>>>(ln.fin.as<float>() + cur, ln.ly.as<double>(), cur);
                count_launch(); count_launch();
            } else {
                CU_TRY(cudaMemcpyAsync(ln.lx.p, loc_x + p0, sizeof(double) * cur, cudaMemcpyHostToDevice, s));
                CU_TRY(cudaMemcpyAsync(ln.ly.p, loc_y + p0, sizeof(double) * cur, cudaMemcpyHostToDevice, s));
            }
            LocSpec L{ln.lx.as<double>(), ln.ly.as<double>(), 0, 0, 0, 0, 1, 0};
            const int32_t* perm = nullptr;
            if (use_cells) {
                RF_TRY(build_loc_perm(L, cur, ln.perm.p, &perm, s));
                RF_TRY(launch_knn_cells<KNN_FINAL>(L, cur, cells, perm, n_obs + 1, d_z.as<double>(), rm_dupl,
                                                   ln.dist.as<double>(), ln.obs.as<double>(),
                                                   idx_cols ? ln.idx.as<int32_t>() : nullptr, cur, d_flags.as<int>(),
                                                   d_flags.as<int>() + 1, s, F32 && prefilter));
            } else {
                RF_TRY(launch_knn<KNN_FINAL>(L, cur, d_xy.as<double2>(), nobs, n_obs + 1, d_z.as<double>(), rm_dupl,
                                             ln.dist.as<double>(), ln.obs.as<double>(),
                                             idx_cols ? ln.idx.as<int32_t>() : nullptr, cur, d_flags.as<int>(),
                                             d_flags.as<int>() + 1, s));
            }
            char* hs = ln.stage.as<char>();
            T* h_dist = reinterpret_cast<T*>(hs);
            T* h_obs = h_dist + (size_t)cur * n_obs;
            int32_t* h_idx = reinterpret_cast<int32_t*>(h_obs + (size_t)cur * n_obs);
            const void* src_dist = ln.dist.p;
            const void* src_obs = ln.obs.p;
            if (F32) {
                const long long ne = (long long)cur * n_obs;
                f64_to_f32_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, s>>>(ln.dist.as<double>(), ln.fout.as<float>(), ne);
                f64_to_f32_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, s>>>(ln.obs.as<double>(), ln.fout.as<float>() + ne, ne);
                count_launch(); count_launch();
                src_dist = ln.fout.p;
                src_obs = ln.fout.as<float>() + ne;
            }
            CU_TRY(cudaMemcpyAsync(h_dist, src_dist, sizeof(T) * cur * n_obs, cudaMemcpyDeviceToHost, s));
            CU_TRY(cudaMemcpyAsync(h_obs, src_obs, sizeof(T) * cur * n_obs, cudaMemcpyDeviceToHost, s));
            if (idx_cols) CU_TRY(cudaMemcpyAsync(h_idx, ln.idx.p, sizeof(int32_t) * cur * n_obs, cudaMemcpyDeviceToHost, s));
            CU_TRY(cudaStreamSynchronize(s));
            for (int j = 0; j < n_obs; ++j) {   // the caller's (possibly never touched) pages: this thread maps and fills them
                prefault_for_write(dist_cols[j] + p0, sizeof(T) * cur);
                prefault_for_write(obs_cols[j] + p0, sizeof(T) * cur);
                memcpy(dist_cols[j] + p0, h_dist + (size_t)j * cur, sizeof(T) * cur);
                memcpy(obs_cols[j] + p0, h_obs + (size_t)j * cur, sizeof(T) * cur);
                if (idx_cols) memcpy(idx_cols[j] + p0, h_idx + (size_t)j * cur, sizeof(int32_t) * cur);
            }
        }
        return RFSI_OK;
    });
    tr.mark("near.obs chunks (search + D2H + scatter)");
    RF_TRY(rc);
    int flags[2] = {0, 0};
    CU_TRY(cudaMemcpy(flags, d_flags.p, sizeof flags, cudaMemcpyDeviceToHost));
    if (flags[1]) return fail(RFSI_E_NA_INPUT, "near_obs: a location has an NA coordinate");
    if (flags[0]) {
        g_err = "near_obs: fewer observations than n.obs(+1); missing neighbours are NA";
        return RFSI_W_TOO_FEW_OBS;
    }
    return RFSI_OK;
}

int check_near_obs_args(int64_t npix, int32_t nobs, int32_t n_obs, const void* loc_x, const void* loc_y, const void* a,
                        const void* b, const void* ox, const void* oy, const void* oz) {
    if (npix < 0 || nobs < 0 || n_obs < 1 || n_obs > kMaxNObs)
        return fail(RFSI_E_ARG, "near_obs: npix/nobs must be >= 0, 1 <= n.obs <= %d", kMaxNObs);
    if (npix > 0 && (!loc_x || !loc_y || !a || !b)) return fail(RFSI_E_ARG, "near_obs: NULL buffer");
    if (nobs > 0 && (!ox || !oy || !oz)) return fail(RFSI_E_ARG, "near_obs: NULL observations");
    return RFSI_OK;
}

}  // namespace

extern "C" int rfsi_near_obs_f64(const double* loc_x, const double* loc_y, int64_t npix, const double* obs_x,
                                 const double* obs_y, const double* obs_z, int32_t nobs, int32_t n_obs,
                                 int32_t rm_dupl, double* out_dist, double* out_obs, int32_t* out_idx, int device)
{
    RF_TRY(check_near_obs_args(npix, nobs, n_obs, loc_x, loc_y, out_dist, out_obs, obs_x, obs_y, obs_z));
    std::vector<double*> dc((size_t)n_obs), oc((size_t)n_obs);
    std::vector<int32_t*> ic((size_t)n_obs);
    for (int j = 0; j < n_obs; ++j) {
        dc[(size_t)j] = out_dist + (size_t)j * npix;
        oc[(size_t)j] = out_obs + (size_t)j * npix;
        ic[(size_t)j] = out_idx ? out_idx + (size_t)j * npix : nullptr;
    }
    return host_near_obs<double>(loc_x, loc_y, npix, obs_x, obs_y, obs_z, nobs, n_obs, rm_dupl, dc.data(), oc.data(),
                                 out_idx ? ic.data() : nullptr, device);
}

extern "C" int rfsi_near_obs_cols_f64(const double* loc_x, const double* loc_y, int64_t npix, const double* obs_x,
                                      const double* obs_y, const double* obs_z, int32_t nobs, int32_t n_obs,
                                      int32_t rm_dupl, double* const* dist_cols, double* const* obs_cols,
                                      int32_t* const* idx_cols, int device)
{
    RF_TRY(check_near_obs_args(npix, nobs, n_obs, loc_x, loc_y, dist_cols, obs_cols, obs_x, obs_y, obs_z));
    for (int j = 0; j < n_obs && npix > 0; ++j)
        if (!dist_cols[j] || !obs_cols[j] || (idx_cols && !idx_cols[j]))
            return fail(RFSI_E_ARG, "near_obs: column %d of the output is NULL", j + 1);
    return host_near_obs<double>(loc_x, loc_y, npix, obs_x, obs_y, obs_z, nobs, n_obs, rm_dupl, dist_cols, obs_cols, idx_cols, device);
}

// fp32 mode of near.obs (north_star: "1e-5 in fp32 mode"): float buffers in and out through the same pinned, chunked
// pipeline as the fp64 call -- half the bytes over PCIe and through the host copies, which is where the call's time goes
// (profiles/r02_i_fp32_mode.txt) -- and fp32 distance arithmetic as a prefilter in the cell-grid search, every survivor
// re-ranked by its exact fp64 distance (knn_cells.cuh).  Neighbour indices are therefore bit-identical to the fp64 call on
// the same (float-representable) coordinates; dist/obs are its results rounded to float (6e-8 relative).
extern "C" int rfsi_near_obs_f32(const float* loc_x, const float* loc_y, int64_t npix, const float* obs_x,
                                 const float* obs_y, const float* obs_z, int32_t nobs, int32_t n_obs, int32_t rm_dupl,
                                 float* out_dist, float* out_obs, int32_t* out_idx, int device)
{
    RF_TRY(check_near_obs_args(npix, nobs, n_obs, loc_x, loc_y, out_dist, out_obs, obs_x, obs_y, obs_z));
    std::vector<float*> dc((size_t)n_obs), oc((size_t)n_obs);
    std::vector<int32_t*> ic((size_t)n_obs);
    for (int j = 0; j < n_obs; ++j) {
        dc[(size_t)j] = out_dist + (size_t)j * npix;
        oc[(size_t)j] = out_obs + (size_t)j * npix;
        ic[(size_t)j] = out_idx ? out_idx + (size_t)j * npix : nullptr;
    }
    return host_near_obs<float>(loc_x, loc_y, npix, obs_x, obs_y, obs_z, nobs, n_obs, rm_dupl, dc.data(), oc.data(),
                                out_idx ? ic.data() : nullptr, device);
}

// near.obs for all days of a station table (the per-day feature / training table)
extern "C" int rfsi_near_obs_days_f64(const double* loc_x, const double* loc_y, int64_t npix, const double* obs_x,
                                      const double* obs_y, const double* obs_z, int32_t nobs, int32_t ndays,
                                      int32_t n_obs, int32_t rm_dupl, double* out_dist, double* out_obs,
                                      int32_t* out_idx, int device)
{
    const bool loc_is_obs = (loc_x == nullptr);
    if (loc_is_obs) { if (loc_y) return fail(RFSI_E_ARG, "near_obs_days: loc_x/loc_y go together"); npix = nobs; }
    if (npix < 0 || nobs < 0 || ndays < 1 || n_obs < 1 || n_obs > 64)
        return fail(RFSI_E_ARG, "near_obs_days: need npix, nobs >= 0, ndays >= 1, 1 <= n.obs <= 64");
    if (npix > 0 && (!out_dist || !out_obs)) return fail(RFSI_E_ARG, "near_obs_days: NULL output");
    if (nobs > 0 && (!obs_x || !obs_y || !obs_z)) return fail(RFSI_E_ARG, "near_obs_days: NULL observations");
    for (int32_t s = 0; s < nobs; ++s)
        if (obs_x[s] != obs_x[s] || obs_y[s] != obs_y[s])
            return fail(RFSI_E_NA_INPUT, "near_obs_days: observation %d has an NA coordinate", s + 1);
    RF_TRY(use_device(device));
    if (npix == 0) return RFSI_OK;
    const int S = nobs;
    int worst_missing = 0;
    for (int d = 0; d < ndays; ++d) {
        int miss = 0;
        for (int s = 0; s < S; ++s) miss += std::isnan(obs_z[(size_t)d * S + s]) ? 1 : 0;
        worst_missing = std::max(worst_missing, miss);
    }
    const int k = n_obs + 1;
    const int KC = std::max(1, std::min(std::max(S, 1), k + std::min(worst_missing, 16)));
    std::vector<double2> xy((size_t)std::max(S, 1));
    for (int32_t s = 0; s < S; ++s) xy[s] = make_double2(obs_x[s], obs_y[s]);
    const int64_t chunk = std::min<int64_t>(npix, 1 << 18);
    const int db = (int)std::max<int64_t>(1, std::min<int64_t>(ndays, (int64_t(256) << 20) / (chunk * n_obs * 20)));
    DevBuf d_xy, d_z, d_lx, d_ly, d_cd2, d_cid, d_cells, d_dist, d_obs, d_idx, d_flags, d_perm;
    Stream st;   // declared after every buffer its work touches: destroyed (and synchronised) first on any return
    RF_TRY(st.create());
    RF_TRY(d_xy.alloc(sizeof(double2) * xy.size()));
    RF_TRY(d_z.alloc(sizeof(double) * (size_t)std::max(S, 1) * ndays));
    RF_TRY(d_lx.alloc(sizeof(double) * chunk)); RF_TRY(d_ly.alloc(sizeof(double) * chunk));
    RF_TRY(d_cd2.alloc(sizeof(double) * chunk * KC)); RF_TRY(d_cid.alloc(sizeof(int32_t) * chunk * KC));
    RF_TRY(d_dist.alloc(sizeof(double) * chunk * n_obs * db)); RF_TRY(d_obs.alloc(sizeof(double) * chunk * n_obs * db));
    if (out_idx) RF_TRY(d_idx.alloc(sizeof(int32_t) * chunk * n_obs * db));
    RF_TRY(d_flags.alloc(16));
    const bool use_cells = S >= cells_min_S();
    CellsWs cells;
    if (use_cells) {
        RF_TRY(d_cells.alloc(cells_ws_bytes(S)));
        cells = carve_cells(d_cells.p, S);
        RF_TRY(d_perm.alloc(loc_perm_bytes(chunk)));
    }
    CU_TRY(cudaMemcpy(d_xy.p, xy.data(), sizeof(double2) * xy.size(), cudaMemcpyHostToDevice));
    if (S > 0) CU_TRY(cudaMemcpy(d_z.p, obs_z, sizeof(double) * (size_t)S * ndays, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemset(d_flags.p, 0, 16));
    RF_TRY(pool_ready());
    cudaStream_t s = st.s;
    if (use_cells) RF_TRY(build_cells(d_xy.as<double2>(), S, cells, s));
    for (int64_t p0 = 0; p0 < npix; p0 += chunk) {
        const int64_t cur = std::min(chunk, npix - p0);
        CU_TRY(cudaMemcpyAsync(d_lx.p, (loc_is_obs ? obs_x : loc_x) + p0, sizeof(double) * cur, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(d_ly.p, (loc_is_obs ? obs_y : loc_y) + p0, sizeof(double) * cur, cudaMemcpyHostToDevice, s));
        LocSpec L{d_lx.as<double>(), d_ly.as<double>(), 0, 0, 0, 0, 1, 0};
        const int32_t* perm = nullptr;
        if (use_cells) RF_TRY(build_loc_perm(L, cur, d_perm.p, &perm, s));
        if (use_cells)
            RF_TRY(launch_knn_cells<KNN_RAW>(L, cur, cells, perm, KC, nullptr, 0, d_cd2.as<double>(), nullptr,
                                             d_cid.as<int32_t>(), cur, nullptr, d_flags.as<int>() + 1, s));
        else
            RF_TRY(launch_knn<KNN_RAW>(L, cur, d_xy.as<double2>(), S, KC, nullptr, 0, d_cd2.as<double>(), nullptr,
                                       d_cid.as<int32_t>(), cur, nullptr, d_flags.as<int>() + 1, s));
        for (int d0 = 0; d0 < ndays; d0 += db) {
            const int nd = std::min(db, ndays - d0);
            FeatureParams fp;
            memset(&fp, 0, sizeof fp);
            fp.loc = L; fp.npix = cur; fp.ndays = nd;
            fp.cand_d2 = d_cd2.as<double>(); fp.cand_id = d_cid.as<int32_t>(); fp.ld_cand = cur;
            fp.KC = KC; fp.exhaustive = KC >= S;
            fp.st_xy = d_xy.as<double2>();
            fp.obs_z = d_z.as<double>() + (size_t)d0 * S;
            fp.S = S; fp.n_obs = n_obs; fp.rm_dupl = rm_dupl;
            fp.loc_is_obs = loc_is_obs ? 1 : 0;
            fp.loc_obs_offset = p0;
            fp.out_dist = d_dist.as<double>(); fp.out_obs = d_obs.as<double>();
            fp.out_idx = out_idx ? d_idx.as<int32_t>() : nullptr;
            fp.warn = d_flags.as<int>(); fp.fallbacks = nullptr;
            const unsigned grid = (unsigned)(((cur + 127) / 128) * nd);
            features_kernel<<<grid, 128, 0, s>>>(fp);
            count_launch();
            CU_TRY(cudaGetLastError());
            // device [dl][j][cur] -> host [(d0+dl)][j][npix] at column p0
            const size_t rows = (size_t)nd * n_obs;
            CU_TRY(cudaMemcpy2DAsync(out_dist + (size_t)d0 * n_obs * npix + p0, sizeof(double) * npix, d_dist.p,
                                     sizeof(double) * cur, sizeof(double) * cur, rows, cudaMemcpyDeviceToHost, s));
            CU_TRY(cudaMemcpy2DAsync(out_obs + (size_t)d0 * n_obs * npix + p0, sizeof(double) * npix, d_obs.p,
                                     sizeof(double) * cur, sizeof(double) * cur, rows, cudaMemcpyDeviceToHost, s));
            if (out_idx)
                CU_TRY(cudaMemcpy2DAsync(out_idx + (size_t)d0 * n_obs * npix + p0, sizeof(int32_t) * npix, d_idx.p,
                                         sizeof(int32_t) * cur, sizeof(int32_t) * cur, rows, cudaMemcpyDeviceToHost, s));
            CU_TRY(cudaStreamSynchronize(s));   // the staging buffers are reused by the next batch
        }
    }
    int flags[2] = {0, 0};
    CU_TRY(cudaMemcpy(flags, d_flags.p, sizeof flags, cudaMemcpyDeviceToHost));
    if (flags[1]) return fail(RFSI_E_NA_INPUT, "near_obs_days: a location has an NA coordinate");
    if (flags[0]) {
        g_err = "near_obs_days: some rows have fewer valid observations than n.obs(+1); missing neighbours are NA";
        return RFSI_W_TOO_FEW_OBS;
    }
    return RFSI_OK;
}

// ------------------------------------------------------------------------------------
// stage 2 API
// ------------------------------------------------------------------------------------
extern "C" size_t rfsi_dev_forest_scratch_bytes(const rfsi_forest_t* f, int64_t npix, int32_t nprobs) {
    if (!f || npix <= 0 || nprobs <= 0) return 256;
    return align_up(leaf_ids_bytes(npix, f->ntrees), 256) + align_up(sizeof(double) * kMaxProbs, 256) +
           align_up((size_t)npix, 256);   // leaf ids | quantile levels | per-row skip bytes
}

extern "C" int rfsi_dev_forest_predict(const rfsi_forest_t* f, const double* X, int64_t npix, int64_t ldx,
                                       double* out_mean, const double* probs_host, int32_t nprobs, double* out_q,
                                       int32_t* out_terminal, void* scratch, unsigned long long* visits,
                                       int32_t* na_flag, int device, void* stream)
{
    if (!f) return fail(RFSI_E_ARG, "forest is NULL");
    if (npix < 0 || ldx < npix) return fail(RFSI_E_ARG, "predict: need 0 <= npix <= ldx");
    RF_TRY(use_device(device));
    rfsi_forest::Image im;
    RF_TRY(forest_image(f, device, &im));
    cudaStream_t st = (cudaStream_t)stream;
    ForestOut o{out_mean, nullptr, leaf_ids_ld(f->ntrees), out_terminal, nullptr, npix, leaf_ids_word(im), nullptr};
    if (out_terminal) RF_TRY(forest_leaf_rid(f, device, &o.leaf_rid));
    double* probs_dev = nullptr;
    if (out_q) {
        if (!leaf_samples(im)) return fail(RFSI_E_ARG, "quantiles need a forest imported with random.node.values (quantreg=TRUE)");
        if (f->ntrees > 1024) return fail(RFSI_E_ARG, "quantile forests with more than 1024 trees are not supported (T=%d)", f->ntrees);
        RF_TRY(check_probs(probs_host, nprobs));
        if (!scratch) return fail(RFSI_E_ARG, "quantiles need a scratch buffer");
        o.leaf_ids = reinterpret_cast<uint32_t*>(scratch);
        probs_dev = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) +
                                              align_up(leaf_ids_bytes(npix, f->ntrees), 256));
        // rows with NA in a used column are not walked: their leaf ids are never written and the quantile pass skips them
        o.row_skip = reinterpret_cast<unsigned char*>(probs_dev) + align_up(sizeof(double) * kMaxProbs, 256);
        CU_TRY(cudaMemcpyAsync(probs_dev, probs_host, sizeof(double) * nprobs, cudaMemcpyHostToDevice, st));
    }
    RF_TRY(launch_forest_x(f, im, X, npix, ldx, o, visits, na_flag, st));
    if (out_q) RF_TRY(launch_qrf(o.leaf_ids, leaf_samples(im), o.row_skip, npix, f->ntrees, probs_dev, nprobs, out_q, npix, nullptr, st));
    return RFSI_OK;
}

namespace {

// host-buffer driver shared by predict / quantiles / terminalNodes
int host_forest(const rfsi_forest_t* f, const double* X, int64_t npix, int64_t ldx, double* out_mean,
                const double* probs, int nprobs, double* out_q, int32_t* out_term, int device)
{
    if (!f) return fail(RFSI_E_ARG, "forest is NULL");
    if (npix < 0 || ldx < npix) return fail(RFSI_E_ARG, "predict: need 0 <= npix <= ldx");
    if (npix > 0 && !X) return fail(RFSI_E_ARG, "predict: X is NULL");
    if (out_q) RF_TRY(check_probs(probs, nprobs));
    RF_TRY(use_device(device));
    if (npix == 0) return RFSI_OK;
    rfsi_forest::Image im;
    RF_TRY(forest_image(f, device, &im));
    const int F = f->nfeat, T = f->ntrees;
    Trace tr;
    // Chunks of rows, worker w of W takes chunks w, w + W, ...: it gathers its rows of the caller's (pageable) columns
    // into a pinned slot, sends them in one copy, walks the forest on its own stream, brings the results back into the
    // pinned slot and scatters them into the caller's (possibly never touched) arrays -- see "host staging" above.
    int64_t chunk = std::max<int64_t>(4096, (int64_t(16) << 20) / (8 * (int64_t)F));          // ~16 MB of X per chunk
    if (out_q) chunk = std::min(chunk, std::max<int64_t>(4096, (int64_t(256) << 20) / (4 * (int64_t)T)));
    if (out_term) chunk = std::min(chunk, std::max<int64_t>(4096, (int64_t(64) << 20) / (4 * (int64_t)T)));
    if (env_int("RFSI_B200_PREDICT_CHUNK", 0) > 0) chunk = std::max(256, env_int("RFSI_B200_PREDICT_CHUNK", 0));   // tests
    chunk = std::min<int64_t>(npix, chunk / 256 * 256);
    const int64_t nchunks = (npix + chunk - 1) / chunk;
    const int W = host_workers(nchunks);
    DevBuf d_flag;
    RF_TRY(d_flag.alloc(sizeof(int)));
    CU_TRY(cudaMemset(d_flag.p, 0, sizeof(int)));
    const size_t out_row = (out_mean ? sizeof(double) : 0) + (out_q ? sizeof(double) * nprobs : 0) + (out_term ? sizeof(int32_t) * T : 0);
    struct Lane {
        DevBuf X, mean, q, term, scr;
        PinnedBuf in, out;
        Stream st;   // last: destroyed (synchronised) before the buffers above
    };
    std::vector<Lane> lane((size_t)W);
    for (int w = 0; w < W; ++w) {
        Lane& ln = lane[(size_t)w];
        RF_TRY(ln.st.create());
        RF_TRY(ln.X.alloc(sizeof(double) * chunk * F));
        if (out_mean) RF_TRY(ln.mean.alloc(sizeof(double) * chunk));
        if (out_q) {
            RF_TRY(ln.q.alloc(sizeof(double) * chunk * nprobs));
            RF_TRY(ln.scr.alloc(rfsi_dev_forest_scratch_bytes(f, chunk, nprobs)));
        }
        if (out_term) RF_TRY(ln.term.alloc(sizeof(int32_t) * chunk * T));
        RF_TRY(ln.in.alloc(sizeof(double) * (size_t)chunk * F));
        RF_TRY(ln.out.alloc(out_row * (size_t)chunk));
    }
    RF_TRY(pool_ready());
    tr.mark("predict setup");
    const int rc = run_workers(W, device, [&](int w) -> int {
        Lane& ln = lane[(size_t)w];
        cudaStream_t s = ln.st.s;
        for (int64_t c = w; c < nchunks; c += W) {
            const int64_t p0 = c * chunk, cur = std::min(chunk, npix - p0);
            double* hx = ln.in.as<double>();
            for (int j = 0; j < F; ++j) memcpy(hx + (size_t)j * cur, X + (size_t)j * ldx + p0, sizeof(double) * cur);
            CU_TRY(cudaMemcpyAsync(ln.X.p, hx, sizeof(double) * cur * F, cudaMemcpyHostToDevice, s));
            RF_TRY(rfsi_dev_forest_predict(f, ln.X.as<double>(), cur, cur, out_mean ? ln.mean.as<double>() : nullptr,
                                           probs, nprobs, out_q ? ln.q.as<double>() : nullptr,
                                           out_term ? ln.term.as<int32_t>() : nullptr, ln.scr.p, nullptr,
                                           d_flag.as<int>(), device, s));
            char* ho = ln.out.as<char>();
            double* h_mean = reinterpret_cast<double*>(ho);
            double* h_q = h_mean + (out_mean ? cur : 0);
            int32_t* h_term = reinterpret_cast<int32_t*>(h_q + (out_q ? (size_t)cur * nprobs : 0));
            if (out_mean) CU_TRY(cudaMemcpyAsync(h_mean, ln.mean.p, sizeof(double) * cur, cudaMemcpyDeviceToHost, s));
            if (out_q) CU_TRY(cudaMemcpyAsync(h_q, ln.q.p, sizeof(double) * cur * nprobs, cudaMemcpyDeviceToHost, s));
            if (out_term) CU_TRY(cudaMemcpyAsync(h_term, ln.term.p, sizeof(int32_t) * cur * T, cudaMemcpyDeviceToHost, s));
            CU_TRY(cudaStreamSynchronize(s));
            if (out_mean) { prefault_for_write(out_mean + p0, sizeof(double) * cur); memcpy(out_mean + p0, h_mean, sizeof(double) * cur); }
            if (out_q)
                for (int j = 0; j < nprobs; ++j) {
                    prefault_for_write(out_q + (size_t)j * npix + p0, sizeof(double) * cur);
                    memcpy(out_q + (size_t)j * npix + p0, h_q + (size_t)j * cur, sizeof(double) * cur);
                }
            if (out_term)
                for (int t = 0; t < T; ++t) {
                    prefault_for_write(out_term + (size_t)t * npix + p0, sizeof(int32_t) * cur);
                    memcpy(out_term + (size_t)t * npix + p0, h_term + (size_t)t * cur, sizeof(int32_t) * cur);
                }
        }
        return RFSI_OK;
    });
    tr.mark("predict chunks (gather + H2D + walk + D2H + scatter)");
    RF_TRY(rc);
    int flag = 0;
    CU_TRY(cudaMemcpy(&flag, d_flag.p, sizeof flag, cudaMemcpyDeviceToHost));
    if (flag) return fail(RFSI_E_NA_INPUT, "predict: missing data in a column the forest uses");
    return RFSI_OK;
}

}  // namespace

extern "C" int rfsi_forest_predict(const rfsi_forest_t* f, const double* X, int64_t npix, int64_t ldx,
                                   double* out, int device) {
    if (npix > 0 && !out) return fail(RFSI_E_ARG, "predict: out is NULL");
    return host_forest(f, X, npix, ldx, out, nullptr, 0, nullptr, nullptr, device);
}

extern "C" int rfsi_forest_quantiles(const rfsi_forest_t* f, const double* X, int64_t npix, int64_t ldx,
                                     const double* probs, int32_t nprobs, double* out, int device) {
    if (npix > 0 && !out) return fail(RFSI_E_ARG, "quantiles: out is NULL");
    if (f && f->leaf_sample.empty())
        return fail(RFSI_E_ARG, "quantiles need a forest imported with random.node.values (quantreg=TRUE)");
    return host_forest(f, X, npix, ldx, nullptr, probs, nprobs, out, nullptr, device);
}

extern "C" int rfsi_forest_terminal_nodes(const rfsi_forest_t* f, const double* X, int64_t npix, int64_t ldx,
                                          int32_t* out, int device) {
    if (npix > 0 && !out) return fail(RFSI_E_ARG, "terminal_nodes: out is NULL");
    return host_forest(f, X, npix, ldx, nullptr, nullptr, 0, nullptr, out, device);
}

// ------------------------------------------------------------------------------------
// covariate rasters
// ------------------------------------------------------------------------------------
namespace {

size_t raster_cell_bytes(int dtype) {
    switch (dtype & ~RFSI_RASTER_DIVIDE) { case RFSI_F64: return 8; case RFSI_F32: case RFSI_I32: return 4;
                     case RFSI_I16: case RFSI_U16: return 2; case RFSI_U8: return 1; default: return 0; }
}

int check_raster(const rfsi_raster_t* r, const char* who) {
    if (!r || !r->data) return fail(RFSI_E_ARG, "%s: raster or its data is NULL", who);
    if (raster_cell_bytes(r->dtype) == 0) return fail(RFSI_E_ARG, "%s: unknown raster dtype %d", who, r->dtype);
    if (r->ncol < 1 || r->nrow < 1 || r->nlayers < 1 || !(r->dx > 0) || !(r->dy > 0))
        return fail(RFSI_E_ARG, "%s: raster needs ncol, nrow, nlayers >= 1 and dx, dy > 0", who);
    if ((r->dtype & RFSI_RASTER_DIVIDE) && r->scale == 0.0)
        return fail(RFSI_E_ARG, "%s: RFSI_RASTER_DIVIDE with scale 0", who);
    return RFSI_OK;
}

RasterDev raster_dev(const rfsi_raster_t& r, const void* dev_data) {
    return RasterDev{dev_data, r.dtype, r.nlayers, r.ncol, r.nrow, r.x_ul, r.y_ul, r.dx, r.dy, r.nodata, r.scale, r.offset};
}

int launch_raster_extract(const RasterDev& r, int layer0, int nl, const LocSpec& L, int64_t npix, double* out,
                          int64_t ld, cudaStream_t st) {
    if (npix == 0) return RFSI_OK;
    raster_extract_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(r, layer0, nl, L, npix, out, ld);
    count_launch();
    CU_TRY(cudaGetLastError());
    return RFSI_OK;
}

}  // namespace

extern "C" int rfsi_raster_extract(const rfsi_raster_t* r, int32_t layer, const double* loc_x, const double* loc_y,
                                   int64_t npix, double grid_x0, double grid_y0, double grid_dx, double grid_dy,
                                   int64_t grid_ncol, int64_t pix_begin, double* out, int device)
{
    RF_TRY(check_raster(r, "raster_extract"));
    if (layer < 0 || layer >= r->nlayers) return fail(RFSI_E_ARG, "raster_extract: layer %d outside [0,%d)", layer, r->nlayers);
    if (npix < 0 || (npix > 0 && !out)) return fail(RFSI_E_ARG, "raster_extract: bad npix/out");
    if ((loc_x == nullptr) != (loc_y == nullptr)) return fail(RFSI_E_ARG, "raster_extract: loc_x/loc_y go together");
    if (!loc_x && grid_ncol < 1) return fail(RFSI_E_ARG, "raster_extract: raster form needs grid_ncol >= 1");
    RF_TRY(use_device(device));
    if (npix == 0) return RFSI_OK;
    const size_t layer_bytes = (size_t)r->ncol * r->nrow * raster_cell_bytes(r->dtype);
    DevBuf d_r, d_lx, d_ly, d_out;
    RF_TRY(d_r.alloc(layer_bytes));
    RF_TRY(d_out.alloc(sizeof(double) * npix));
    CU_TRY(cudaMemcpy(d_r.p, reinterpret_cast<const char*>(r->data) + layer_bytes * layer, layer_bytes, cudaMemcpyHostToDevice));
    LocSpec L{nullptr, nullptr, grid_x0, grid_y0, grid_dx, grid_dy, loc_x ? 1 : grid_ncol, loc_x ? 0 : pix_begin};
    if (loc_x) {
        RF_TRY(d_lx.alloc(sizeof(double) * npix)); RF_TRY(d_ly.alloc(sizeof(double) * npix));
        CU_TRY(cudaMemcpy(d_lx.p, loc_x, sizeof(double) * npix, cudaMemcpyHostToDevice));
        CU_TRY(cudaMemcpy(d_ly.p, loc_y, sizeof(double) * npix, cudaMemcpyHostToDevice));
        L.x = d_lx.as<double>(); L.y = d_ly.as<double>();
    }
    rfsi_raster_t one = *r;
    one.nlayers = 1;
    RF_TRY(launch_raster_extract(raster_dev(one, d_r.p), 0, 1, L, npix, d_out.as<double>(), npix, nullptr));
    CU_TRY(cudaMemcpy(out, d_out.p, sizeof(double) * npix, cudaMemcpyDeviceToHost));
    return RFSI_OK;
}

// ------------------------------------------------------------------------------------
// fused API
// ------------------------------------------------------------------------------------
namespace {

constexpr int kFusedP = 256;
constexpr int kFallbackP = 128;

struct FusedPlan {
    int KC = 0;
    int db = 1;               // days per launch
    int64_t rows = 0;         // npix * db
    int64_t npix_plan = 0;    // pixels the workspace was sized for
    bool want_q = false;
    bool cand_rank = false, obs_rank = false;   // rank caches of the rank images (fused.cuh)
    // workspace offsets
    size_t off_cand_rank = 0, off_obs_rank = 0;
    size_t off_cand_d2 = 0, off_cand_id = 0, off_xy = 0, off_cells = 0, off_perm = 0, off_fb = 0, off_acc = 0, off_misc = 0, off_samples = 0, off_skip = 0,
           off_probs = 0, total = 0;
};

int validate_fused(const rfsi_forest_t* f, const rfsi_fused_args_t* a) {
    if (!f || !a) return fail(RFSI_E_ARG, "fused: NULL forest or args");
    if (a->struct_size != sizeof(rfsi_fused_args_t))
        return fail(RFSI_E_ARG, "fused: struct_size %zu != %zu (header/library mismatch)", a->struct_size,
                    sizeof(rfsi_fused_args_t));
    if (a->npix < 0 || a->nobs < 0 || a->ndays < 1) return fail(RFSI_E_ARG, "fused: bad npix/nobs/ndays");
    if (a->n_obs < 1 || a->n_obs > kMaxNObs) return fail(RFSI_E_ARG, "fused: n.obs must be in 1..%d", kMaxNObs);
    if (a->ncov < 0 || a->ncov > kMaxCov) return fail(RFSI_E_ARG, "fused: at most %d covariates", kMaxCov);
    if (!a->feat_map) return fail(RFSI_E_ARG, "fused: feat_map is NULL");
    if (!a->loc_x && (a->grid_ncol < 1)) return fail(RFSI_E_ARG, "fused: raster form needs grid_ncol >= 1");
    if ((a->loc_x == nullptr) != (a->loc_y == nullptr)) return fail(RFSI_E_ARG, "fused: loc_x/loc_y go together");
    if (a->nobs > 0 && (!a->obs_x || !a->obs_y || !a->obs_z)) return fail(RFSI_E_ARG, "fused: NULL stations");
    if (a->npix > (int64_t)0x7ffffff0) return fail(RFSI_E_ARG, "fused: at most 2^31 pixels per call; band the raster");
    for (int fidx = 0; fidx < f->nfeat; ++fidx) {
        const int32_t c = a->feat_map[fidx];
        const bool ok = (c >= 0 && c < a->ncov) || (c >= 0x10000 && c < 0x10000 + a->n_obs) ||
                        (c >= 0x20000 && c < 0x20000 + a->n_obs);
        if (!ok) return fail(RFSI_E_ARG, "fused: feat_map[%d] = 0x%x is not a covariate/dist/obs code", fidx, c);
        for (int g = 0; g < fidx; ++g)   // each source feeds exactly one model column (names are unique in R)
            if (a->feat_map[g] == c)
                return fail(RFSI_E_ARG, "fused: feat_map[%d] and feat_map[%d] name the same source 0x%x", g, fidx, c);
    }
    if (a->cov_raster)
        for (int c = 0; c < a->ncov; ++c)
            if (a->cov_raster[c].data) {
                RF_TRY(check_raster(&a->cov_raster[c], "fused"));
                if (a->cov_raster[c].nlayers != 1 && a->cov_raster[c].nlayers < a->ndays)
                    return fail(RFSI_E_ARG, "fused: covariate raster %d has %d layers for %d days", c,
                                a->cov_raster[c].nlayers, a->ndays);
            }
    if (a->ncov > 0) {
        const bool any_plain = [&] {
            for (int c = 0; c < a->ncov; ++c)
                if (!(a->cov_raster && a->cov_raster[c].data)) return true;
            return false;
        }();
        if (any_plain && (!a->cov || !a->cov_day_stride))
            return fail(RFSI_E_ARG, "fused: ncov = %d but cov / cov_day_stride is NULL", a->ncov);
        for (int c = 0; c < a->ncov; ++c)
            if (!(a->cov_raster && a->cov_raster[c].data) && a->npix > 0 && !a->cov[c])
                return fail(RFSI_E_ARG, "fused: covariate %d has neither a column (cov[%d] is NULL) nor a raster", c, c);
    }
    if (a->npix > 0 && !a->out_mean && !a->out_q && !a->out_mean_i16 && !a->out_iqr_i16)
        return fail(RFSI_E_ARG, "fused: no output requested (out_mean, out_q, out_mean_i16, out_iqr_i16 are all NULL)");
    if ((a->cov_loc_x == nullptr) != (a->cov_loc_y == nullptr)) return fail(RFSI_E_ARG, "fused: cov_loc_x/cov_loc_y go together");
    if (a->ncov_fix < 0 || (a->ncov_fix > 0 && !a->cov_fix)) return fail(RFSI_E_ARG, "fused: bad cov_fix list");
    for (int k = 0; k < a->ncov_fix; ++k) {
        const rfsi_cov_fix_t& x = a->cov_fix[k];
        if (x.target < 0 || x.target >= a->ncov || x.other < 0 || x.other >= a->ncov || x.target == x.other)
            return fail(RFSI_E_ARG, "fused: cov_fix[%d] names covariates %d, %d of %d", k, x.target, x.other, a->ncov);
    }
    if ((a->out_q || a->out_iqr_i16)) {
        RF_TRY(check_probs(a->probs, a->nprobs));
        if (f->leaf_sample.empty())
            return fail(RFSI_E_ARG, "quantiles need a forest imported with random.node.values (quantreg=TRUE)");
        if (f->ntrees > 1024)
            return fail(RFSI_E_ARG, "quantile forests with more than 1024 trees are not supported (T=%d)", f->ntrees);
    }
    if (fused_smem_bytes<kFallbackP, false>(f->nfeat, a->n_obs + 1, true) > 227 * 1024 &&
        !(f->compact_ok && fused_smem_bytes<kFallbackP, true>(f->nfeat, a->n_obs + 1, true) <= 227 * 1024))
        return fail(RFSI_E_ARG, "fused: %d features exceed the shared-memory feature tile", f->nfeat);
    return RFSI_OK;
}

// bytes of quantile key scratch one fused launch may use (RFSI_B200_QRF_SCRATCH_MB; read once: the plan must not change
// between rfsi_dev_fused_workspace_bytes and the call)
int64_t qrf_scratch_budget() {
    static const int64_t mb = std::max(64, env_int("RFSI_B200_QRF_SCRATCH_MB", 8192));
    return mb << 20;
}

FusedPlan plan_fused(const rfsi_forest_t* f, const rfsi_fused_args_t* a, int kc_extra) {
    FusedPlan pl;
    const int k = a->n_obs + 1;
    pl.KC = std::max(1, std::min<int>(std::max(a->nobs, 1), k + std::max(0, kc_extra)));
    pl.want_q = a->out_q || a->out_iqr_i16;
    const int64_t npix = std::max<int64_t>(a->npix, 1);
    int db = std::min(a->ndays, 16);
    if (pl.want_q) {   // key scratch: 4 bytes per (pixel-day, tree); days per launch within the scratch budget
        const int64_t per_day = npix * leaf_ids_ld(f->ntrees) * 4;
        db = (int)std::max<int64_t>(1, std::min<int64_t>(db, qrf_scratch_budget() / per_day));
    }
    // keep the 1-D grid below 2^31 CTAs
    while (db > 1 && ((npix + kFusedP - 1) / kFusedP + 64) * db > (int64_t)0x7fffffff) --db;
    pl.db = db;
    pl.rows = npix * db;
    pl.npix_plan = npix;
    size_t o = 0;
    pl.off_cand_d2 = o; o += align_up(sizeof(double) * (size_t)pl.KC * npix, 256);
    pl.off_cand_id = o; o += align_up(sizeof(int32_t) * (size_t)pl.KC * npix, 256);
    pl.off_xy = o; o += align_up(sizeof(double2) * (size_t)std::max(a->nobs, 1), 256);
    pl.off_cells = o; o += align_up(cells_ws_bytes(a->nobs), 256);
    pl.off_perm = o; o += align_up(a->loc_x ? loc_perm_bytes(npix) : 0, 256);
    pl.off_fb = o; o += align_up(sizeof(int2) * (size_t)pl.rows, 256);
    pl.off_acc = o; o += align_up(sizeof(double) * (size_t)pl.rows, 256);
    pl.off_misc = o; o += 256;  // fb_count, na_flag
    pl.off_probs = o; o += align_up(sizeof(double) * kMaxProbs, 256);
    // rank caches (fused.cuh): worth their fill only when more than one day reads them / more pixels than stations
    if (f->compact_ok && env_int("RFSI_B200_RANK_CACHE", 1)) {
        const size_t ob = sizeof(uint32_t) * (size_t)db * (size_t)std::max(a->nobs, 1) * (size_t)a->n_obs;
        pl.cand_rank = a->ndays >= 2;
        pl.obs_rank = ob <= ((size_t)1 << 30) && (int64_t)a->nobs * 4 <= npix * (int64_t)db;
        if (pl.cand_rank) { pl.off_cand_rank = o; o += align_up(sizeof(uint32_t) * (size_t)pl.KC * kCandShifts * npix, 256); }
        if (pl.obs_rank) { pl.off_obs_rank = o; o += align_up(ob, 256); }
    }
    if (pl.want_q) {
        pl.off_samples = o; o += align_up(leaf_ids_bytes(pl.rows, f->ntrees), 256);
        pl.off_skip = o; o += align_up((size_t)pl.rows, 256);
    }
    pl.total = o;
    return pl;
}

int kc_extra_default() {
    static int v = env_int("RFSI_B200_KC_EXTRA", 11);
    return v;
}

template <int P, int J, bool FALLBACK, bool Q>
int launch_fused_pj(const FusedParams& fp_in, unsigned grid, int k, bool count, cudaStream_t st) {
    FusedParams fp = fp_in;
    fp.q.sect_tile_stride = (uint32_t)(P * sizeof(uint32_t));
    const size_t smem = fused_smem_bytes<P, Q>(fp.F, k, FALLBACK);
    ProfScope ps(FALLBACK ? PROF_FALLBACK : PROF_FUSED, st);
    const int carve = env_int("RFSI_B200_CARVEOUT", -1);   // experiment knob: shared-memory carveout in percent
    if (count) {
        auto kern = fused_kernel<P, J, true, FALLBACK, Q>;
        CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (carve >= 0) CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        kern<<<grid, P, smem, st>>>(fp);
    } else {
        auto kern = fused_kernel<P, J, false, FALLBACK, Q>;
        CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (carve >= 0) CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        kern<<<grid, P, smem, st>>>(fp);
    }
    count_launch();
    CU_TRY(cudaGetLastError());
    return RFSI_OK;
}

template <int P, bool FALLBACK, bool Q>
int launch_fused_p(const FusedParams& fp, unsigned grid, int k, bool count, cudaStream_t st) {
    const int J = forest_J();
    if (!FALLBACK) {
        if (J >= 16) return launch_fused_pj<P, 16, FALLBACK, Q>(fp, grid, k, count, st);
        if (J >= 12) return launch_fused_pj<P, 12, FALLBACK, Q>(fp, grid, k, count, st);
    }
    if (J >= 8) return launch_fused_pj<P, 8, FALLBACK, Q>(fp, grid, k, count, st);
    if (J >= 4) return launch_fused_pj<P, 4, FALLBACK, Q>(fp, grid, k, count, st);
    return launch_fused_pj<P, 2, FALLBACK, Q>(fp, grid, k, count, st);
}

// One pixel chunk, all days, everything already on the device.
int dev_fused(const rfsi_forest_t* f, const rfsi_fused_args_t* a, const FusedPlan& pl, void* workspace,
              unsigned long long* counters, int device, cudaStream_t st, bool reuse_candidates,
              bool count_visits)
{
    rfsi_forest::Image im;
    RF_TRY(forest_image(f, device, &im));
    if (a->npix == 0) return RFSI_OK;
    char* ws = reinterpret_cast<char*>(workspace);
    double* cand_d2 = reinterpret_cast<double*>(ws + pl.off_cand_d2);
    int32_t* cand_id = reinterpret_cast<int32_t*>(ws + pl.off_cand_id);
    double2* xy = reinterpret_cast<double2*>(ws + pl.off_xy);
    int2* fb = reinterpret_cast<int2*>(ws + pl.off_fb);
    double* acc = reinterpret_cast<double*>(ws + pl.off_acc);
    unsigned int* fb_count = reinterpret_cast<unsigned int*>(ws + pl.off_misc);
    int* na_flag = reinterpret_cast<int*>(ws + pl.off_misc + 16);
    double* probs_dev = reinterpret_cast<double*>(ws + pl.off_probs);
    uint32_t* leaf_ids = pl.want_q ? reinterpret_cast<uint32_t*>(ws + pl.off_samples) : nullptr;
    unsigned char* skip = pl.want_q ? reinterpret_cast<unsigned char*>(ws + pl.off_skip) : nullptr;

    // explicit coordinates: grid_ncol > 1 is the row length of a row-by-row ordering (a mapping hint only)
    LocSpec L{a->loc_x, a->loc_y, a->grid_x0, a->grid_y0, a->grid_dx, a->grid_dy,
              a->loc_x ? std::max<int64_t>(1, a->grid_ncol) : a->grid_ncol, a->loc_x ? 0 : a->pix_begin};
    // unordered explicit locations are visited along a Morton curve (kNN tiles and forest warps
    // stay spatially compact); the permutation lives in the workspace next to the candidates
    const bool want_perm = a->loc_x && !loc_rows(L, a->npix) && a->npix <= 0x7fffffff && env_int("RFSI_B200_MORTON", 1);
    const int32_t* perm = want_perm
        ? reinterpret_cast<const int32_t*>(ws + pl.off_perm + 256 + 2 * align_up((size_t)a->npix * 8, 256) +
                                           align_up((size_t)a->npix * 4, 256))
        : nullptr;
    if (!reuse_candidates) {
        if (a->nobs > 0) {
            interleave_xy_kernel<<<(a->nobs + 255) / 256, 256, 0, st>>>(a->obs_x, a->obs_y, a->nobs, xy);
            count_launch();
        }
        if (want_perm) {
            const int32_t* built = nullptr;
            RF_TRY(build_loc_perm(L, a->npix, ws + pl.off_perm, &built, st));
            if (built != perm) return fail(RFSI_E_CUDA, "fused: permutation layout mismatch");
        }
        // geometry pass: the KC nearest of ALL stations, once per pixel
        if (a->nobs >= cells_min_S()) {
            const CellsWs c = carve_cells(ws + pl.off_cells, a->nobs);
            RF_TRY(build_cells(xy, a->nobs, c, st));
            RF_TRY(launch_knn_cells<KNN_RAW>(L, a->npix, c, perm, pl.KC, nullptr, 0, cand_d2, nullptr, cand_id,
                                             a->npix, nullptr, na_flag, st));
        } else {
            RF_TRY(launch_knn<KNN_RAW>(L, a->npix, xy, a->nobs, pl.KC, nullptr, 0, cand_d2, nullptr, cand_id, a->npix,
                                       nullptr, na_flag, st));
        }
    }
    if (pl.want_q)
        CU_TRY(cudaMemcpyAsync(probs_dev, a->probs, sizeof(double) * a->nprobs, cudaMemcpyHostToDevice, st));

    FusedParams fp;
    memset(&fp, 0, sizeof fp);
    fp.loc = L;
    fp.npix = a->npix;
    fp.cand_d2 = cand_d2; fp.cand_id = cand_id; fp.ld_cand = a->npix; fp.KC = pl.KC;
    fp.exhaustive = pl.KC >= a->nobs;
    fp.st_xy = xy; fp.obs_z = a->obs_z; fp.S = a->nobs;
    fp.n_obs = a->n_obs; fp.rm_dupl = a->rm_dupl;
    fp.ncov = a->ncov;
    for (int c = 0; c < kMaxCov; ++c) fp.cov_row[c] = -1;
    for (int j = 0; j < kMaxNObs; ++j) { fp.dist_row[j] = -1; fp.obs_row[j] = -1; }
    for (int c = 0; c < a->ncov; ++c) { fp.cov[c] = a->cov[c]; fp.cov_stride[c] = a->cov_day_stride[c]; }
    for (int fi = 0; fi < f->nfeat; ++fi) {
        const int32_t code = a->feat_map[fi];
        if (code >= 0x20000) fp.obs_row[code - 0x20000] = (short)fi;
        else if (code >= 0x10000) fp.dist_row[code - 0x10000] = (short)fi;
        else fp.cov_row[code] = (short)fi;  // a covariate feeds one model column
    }
    fp.pix_mask = a->pix_mask;
    fp.nodes = im.nodes; fp.q = im.q; fp.leaf_rid = nullptr;
    fp.T = f->ntrees; fp.F = f->nfeat;
    fp.acc = acc;
    // rank caches: the distance ranks of the candidates once per pixel (they live as long as the candidate lists) ...
    uint32_t* cand_rank = (im.compact && pl.cand_rank) ? reinterpret_cast<uint32_t*>(ws + pl.off_cand_rank) : nullptr;
    uint32_t* obs_rank = (im.compact && pl.obs_rank && a->nobs > 0) ? reinterpret_cast<uint32_t*>(ws + pl.off_obs_rank) : nullptr;
    if (cand_rank && !reuse_candidates) {
        ProfScope ps(PROF_KNN_RAW, st);
        const int ncand = std::min(pl.KC, a->n_obs + kCandShifts - 1);   // later candidates have no column within the shifts
        const long long n = (long long)a->npix * ncand;
        cand_rank_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(fp, cand_rank, ncand);
        count_launch();
        CU_TRY(cudaGetLastError());
    }
    fp.cand_rank = cand_rank;
    fp.fb_list = fb; fp.fb_count = fb_count; fp.fb_capacity = (unsigned int)std::min<int64_t>(pl.rows, 0xffffffffLL);
    fp.counters = counters; fp.na_flag = na_flag;
    fp.tile2d = (env_int("RFSI_B200_TILE2D", 1) && loc_rows(L, a->npix)) ? 1 : 0;
    fp.perm = fp.tile2d ? nullptr : perm;

    const bool Q = im.compact;
    g_default_J = Q ? 8 : 12;
    const bool main256 = env_int("RFSI_B200_P", 256) >= 256 &&
                         (Q ? (fused_smem_bytes<kFusedP, true>(f->nfeat, 0, false) <= 110 * 1024)
                            : (fused_smem_bytes<kFusedP, false>(f->nfeat, 0, false) <= 100 * 1024));
    const int Pm = main256 ? kFusedP : kFallbackP;
    long long ntiles;
    if (fp.tile2d) {
        fp.tiles_per_row = (L.ncol + 15) / 16;
        const long long nrows = (a->npix + L.ncol - 1) / L.ncol;
        ntiles = fp.tiles_per_row * ((nrows + (Pm / 16) - 1) / (Pm / 16));
    } else {
        ntiles = (a->npix + Pm - 1) / Pm;
    }
    // node-visit counting (counters[0..1]) costs issue slots in the hot loop: only on request.
    // The too-few / fallback counters (counters[2..3]) are kept whenever `counters` is given.
    const bool count = counters != nullptr && count_visits;
    const std::vector<int> cb = tree_chunks(f);

    for (int d0 = 0; d0 < a->ndays; d0 += pl.db) {
        const int nd = std::min(pl.db, a->ndays - d0);
        const int64_t rows = a->npix * nd;
        fp.day0 = d0; fp.ndays = nd;
        const int64_t obase = (int64_t)d0 * a->npix;
        fp.out_mean = a->out_mean ? a->out_mean + obase : nullptr;
        fp.out_mean_i16 = a->out_mean_i16 ? a->out_mean_i16 + obase : nullptr;
        fp.leaf_ids = leaf_ids; fp.ld_ids = leaf_ids_ld(f->ntrees); fp.row_skip = skip;
        fp.ids_word = leaf_ids_word(im);
        CU_TRY(cudaMemsetAsync(fb_count, 0, sizeof(unsigned int), st));
        if (obs_rank) {   // ... and the observation ranks once per launch: [day][station][column]
            fp.obs_rank = nullptr;
            const long long n = (long long)nd * a->nobs * a->n_obs;
            obs_rank_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(fp, obs_rank);
            count_launch();
            CU_TRY(cudaGetLastError());
            fp.obs_rank = obs_rank;
        }
        const unsigned grid = (unsigned)(ntiles * nd);
        for (size_t c = 0; c + 1 < cb.size(); ++c) {  // one sweep of all tiles per L2-sized forest chunk
            fp.t_begin = cb[c]; fp.t_end = cb[c + 1];
            const int k = a->n_obs + 1;
            if (Q) {
                if (main256) RF_TRY((launch_fused_p<kFusedP, false, true>(fp, grid, k, count, st)));
                else RF_TRY((launch_fused_p<kFallbackP, false, true>(fp, grid, k, count, st)));
            } else {
                if (main256) RF_TRY((launch_fused_p<kFusedP, false, false>(fp, grid, k, count, st)));
                else RF_TRY((launch_fused_p<kFallbackP, false, false>(fp, grid, k, count, st)));
            }
            if (!fp.exhaustive) {
                const unsigned fgrid = (unsigned)((rows + kFallbackP - 1) / kFallbackP);
                if (Q) RF_TRY((launch_fused_p<kFallbackP, true, true>(fp, fgrid, k, count, st)));
                else RF_TRY((launch_fused_p<kFallbackP, true, false>(fp, fgrid, k, count, st)));
            }
        }
        if (pl.want_q) {
            double* oq = a->out_q ? a->out_q + obase : nullptr;
            // q[(j*ndays + d)*npix + p]: column stride between quantiles is ndays*npix
            RF_TRY(launch_qrf(leaf_ids, leaf_samples(im), skip, rows, f->ntrees, probs_dev, a->nprobs, oq, (int64_t)a->ndays * a->npix,
                              a->out_iqr_i16 ? a->out_iqr_i16 + obase : nullptr, st));
        }
    }
    return RFSI_OK;
}

}  // namespace

extern "C" size_t rfsi_dev_fused_workspace_bytes(const rfsi_forest_t* f, const rfsi_fused_args_t* a) {
    if (!f || !a || a->struct_size != sizeof(rfsi_fused_args_t)) return 0;
    return plan_fused(f, a, kc_extra_default()).total;
}

extern "C" int rfsi_dev_predict_fused(const rfsi_forest_t* f, const rfsi_fused_args_t* a, void* workspace,
                                      unsigned long long* counters, int device, void* stream)
{
    RF_TRY(validate_fused(f, a));
    if (a->cov_raster) return fail(RFSI_E_ARG, "fused (device form): sample rasters first; cov_raster is host-entry only");
    if (a->ncov_fix) return fail(RFSI_E_ARG, "fused (device form): covariate columns are read-only; cov_fix is host-entry only");
    if (a->cov_loc_x) return fail(RFSI_E_ARG, "fused (device form): cov_loc_x/cov_loc_y are host-entry only");
    if (!workspace) return fail(RFSI_E_ARG, "fused: workspace is NULL");
    RF_TRY(use_device(device));
    const FusedPlan pl = plan_fused(f, a, kc_extra_default());
    int* na_flag = reinterpret_cast<int*>(reinterpret_cast<char*>(workspace) + pl.off_misc + 16);
    CU_TRY(cudaMemsetAsync(na_flag, 0, sizeof(int), (cudaStream_t)stream));
    return dev_fused(f, a, pl, workspace, counters, device, (cudaStream_t)stream, false, true);
}

namespace {

int validate_fused_host(const rfsi_forest_t* f, const rfsi_fused_args_t* a) {
    RF_TRY(validate_fused(f, a));
    for (int32_t s = 0; s < a->nobs; ++s)
        if (a->obs_x[s] != a->obs_x[s] || a->obs_y[s] != a->obs_y[s])
            return fail(RFSI_E_NA_INPUT, "fused: station %d has an NA coordinate", s + 1);
    return RFSI_OK;
}

// Explicit coordinates that are really raster cells stored row by row: the row length (the
// caller's grid_ncol hint, else detected from the first two rows), or 0/1 when they are not.
int64_t explicit_row_len(const rfsi_fused_args_t* a) {
    int64_t row_len = a->loc_x ? a->grid_ncol : 0;
    if (a->loc_x && row_len <= 1) {
        int64_t i = 1;
        while (i < a->npix && a->loc_y[i] == a->loc_y[0]) ++i;
        if (i >= 16 && i < a->npix && a->npix % i == 0 && a->npix / i >= 4 && a->loc_y[i - 1] == a->loc_y[0] &&
            a->loc_y[2 * i - 1] == a->loc_y[i] && a->loc_x[i] == a->loc_x[0])
            row_len = i;
    }
    return row_len;
}

// The host-buffer pipeline for the pixels [pb, pe) of the call (the whole call: 0, npix; one
// band of it per device in rfsi_predict_fused_multi).  Outputs, covariates and masks are
// indexed by the call's global pixel index, so bands write disjoint slices of the same arrays.
int host_fused(const rfsi_forest_t* f, const rfsi_fused_args_t* a, int device, int64_t pb, int64_t pe)
{
    RF_TRY(use_device(device));
    if (pe <= pb) return RFSI_OK;
    const int64_t band = pe - pb;
    Trace tr;
    const int S = a->nobs, D = a->ndays;
    // candidates beyond n.obs+1: as many as the worst day has missing stations (capped)
    int worst_missing = 0;
    for (int d = 0; d < D; ++d) {
        int miss = 0;
        for (int s = 0; s < S; ++s) miss += std::isnan(a->obs_z[(size_t)d * S + s]) ? 1 : 0;
        worst_missing = std::max(worst_missing, miss);
    }
    const int kc_extra = std::min(worst_missing, std::max(kc_extra_default(), 0) + 5);

    // explicit coordinates that are really raster cells stored row by row (the usual R input:
    // coordinates(SpatialPixels) / a grid data.frame): find the row length so threads can be
    // mapped to compact 2-D tiles.  Only a mapping hint -- results do not depend on it.
    const int64_t row_len = explicit_row_len(a);

    // pixel chunking: bound per-chunk device memory
    const bool want_q = a->out_q || a->out_iqr_i16;
    int64_t chunk = std::max<int64_t>(4096, env_int("RFSI_B200_FUSED_CHUNK", 1 << 20));
    if (want_q) chunk = std::max<int64_t>(4096, std::min<int64_t>(chunk, qrf_scratch_budget() / (8 * (int64_t)f->ntrees)));
    chunk = std::min(chunk, band);
    // whole raster rows per chunk keep the 2-D tiles, whole tile rows (16 raster rows) keep every tile full: a chunk of
    // 104 rows would leave the lower half of its seventh tile row idle
    auto whole_rows = [&](int64_t ncol) {
        int64_t rows = chunk / ncol;
        if (rows >= kFusedP / 16) rows = rows / (kFusedP / 16) * (kFusedP / 16);
        return rows * ncol;
    };
    if (!a->loc_x && (a->pix_begin + pb) % a->grid_ncol == 0 && chunk < band && chunk >= a->grid_ncol) chunk = whole_rows(a->grid_ncol);
    if (a->loc_x && row_len > 1 && pb % row_len == 0 && chunk < band && chunk >= row_len) chunk = whole_rows(row_len);

    DevBuf d_ox, d_oy, d_oz, d_cnt;
    RF_TRY(d_ox.alloc(sizeof(double) * std::max(S, 1)));
    RF_TRY(d_oy.alloc(sizeof(double) * std::max(S, 1)));
    RF_TRY(d_oz.alloc(sizeof(double) * (size_t)std::max(S, 1) * D));
    if (S > 0) {
        CU_TRY(cudaMemcpy(d_ox.p, a->obs_x, sizeof(double) * S, cudaMemcpyHostToDevice));
        CU_TRY(cudaMemcpy(d_oy.p, a->obs_y, sizeof(double) * S, cudaMemcpyHostToDevice));
        CU_TRY(cudaMemcpy(d_oz.p, a->obs_z, sizeof(double) * (size_t)S * D, cudaMemcpyHostToDevice));
    }
    RF_TRY(d_cnt.alloc(4 * sizeof(unsigned long long)));
    CU_TRY(cudaMemset(d_cnt.p, 0, 4 * sizeof(unsigned long long)));

    // covariate rasters are uploaded whole, once, and sampled on the device per work item
    DevBuf d_rast[kMaxCov];
    bool is_rast[kMaxCov] = {false};
    for (int c = 0; c < a->ncov; ++c) {
        if (!(a->cov_raster && a->cov_raster[c].data)) continue;
        const rfsi_raster_t& r = a->cov_raster[c];
        const size_t bytes = (size_t)r.ncol * r.nrow * r.nlayers * raster_cell_bytes(r.dtype);
        RF_TRY(d_rast[c].alloc(bytes));
        CU_TRY(cudaMemcpy(d_rast[c].p, r.data, bytes, cudaMemcpyHostToDevice));
        is_rast[c] = true;
    }
    auto cov_dynamic = [&](int c) {
        return is_rast[c] ? a->cov_raster[c].nlayers != 1 : a->cov_day_stride[c] != 0;
    };

    for (int k = 0; k < a->ncov_fix; ++k)
        if (cov_dynamic(a->cov_fix[k].target) != cov_dynamic(a->cov_fix[k].other))
            return fail(RFSI_E_ARG, "fused: cov_fix[%d] pairs a static covariate with a per-day one", k);

    // per-slot device buffers: a work item is (pixel chunk, day batch)
    rfsi_fused_args_t proto = *a;
    proto.npix = chunk;
    FusedPlan pl = plan_fused(f, &proto, kc_extra);
    const int db = pl.db;
    struct Slot {
        DevBuf ws, lx, ly, clx, cly, mask, mean, q, mean16, iqr16;
        DevBuf cov[kMaxCov];
        int64_t cand_chunk = -1;
        Stream st;  // last member: destroyed (and synchronised) before the buffers above
    } slot[2];
    const int nslots = (band <= chunk && D <= db) ? 1 : 2;
    for (int s = 0; s < nslots; ++s) {
        RF_TRY(slot[s].st.create());
        RF_TRY(slot[s].ws.alloc(pl.total));
        if (a->loc_x) { RF_TRY(slot[s].lx.alloc(sizeof(double) * chunk)); RF_TRY(slot[s].ly.alloc(sizeof(double) * chunk)); }
        if (a->cov_loc_x) { RF_TRY(slot[s].clx.alloc(sizeof(double) * chunk)); RF_TRY(slot[s].cly.alloc(sizeof(double) * chunk)); }
        if (a->pix_mask) RF_TRY(slot[s].mask.alloc((size_t)chunk));
        if (a->out_mean) RF_TRY(slot[s].mean.alloc(sizeof(double) * chunk * db));
        if (a->out_q) RF_TRY(slot[s].q.alloc(sizeof(double) * chunk * db * a->nprobs));
        if (a->out_mean_i16) RF_TRY(slot[s].mean16.alloc(sizeof(int16_t) * chunk * db));
        if (a->out_iqr_i16) RF_TRY(slot[s].iqr16.alloc(sizeof(int16_t) * chunk * db));
        for (int c = 0; c < a->ncov; ++c)
            RF_TRY(slot[s].cov[c].alloc(sizeof(double) * chunk * (cov_dynamic(c) ? db : 1)));
        CU_TRY(cudaMemset(reinterpret_cast<char*>(slot[s].ws.p) + pl.off_misc, 0, 256));
    }

    RF_TRY(pool_ready());
    tr.mark("setup (alloc + stations)");
    int si = 0;
    for (int64_t p0 = pb; p0 < pe; p0 += chunk) {
        const int64_t cur = std::min(chunk, pe - p0);
        for (int d0 = 0; d0 < D; d0 += db, si = (si + 1) % nslots) {
            const int nd = std::min(db, D - d0);
            Slot& sl = slot[si];
            cudaStream_t s = sl.st.s;
            const bool reuse = sl.cand_chunk == p0;
            rfsi_fused_args_t b = *a;
            b.npix = cur;
            b.ndays = nd;
            b.obs_x = d_ox.as<double>(); b.obs_y = d_oy.as<double>();
            b.obs_z = d_oz.as<double>() + (size_t)d0 * S;
            if (a->loc_x) {
                if (!reuse) {
                    CU_TRY(cudaMemcpyAsync(sl.lx.p, a->loc_x + p0, sizeof(double) * cur, cudaMemcpyHostToDevice, s));
                    CU_TRY(cudaMemcpyAsync(sl.ly.p, a->loc_y + p0, sizeof(double) * cur, cudaMemcpyHostToDevice, s));
                }
                b.loc_x = sl.lx.as<double>(); b.loc_y = sl.ly.as<double>();
                b.grid_ncol = (row_len > 1 && cur % row_len == 0) ? row_len : 1;
            } else {
                b.pix_begin = a->pix_begin + p0;
            }
            if (a->pix_mask) {
                if (!reuse) CU_TRY(cudaMemcpyAsync(sl.mask.p, a->pix_mask + p0, (size_t)cur, cudaMemcpyHostToDevice, s));
                b.pix_mask = sl.mask.as<unsigned char>();
            }
            if (a->cov_loc_x && !reuse) {
                CU_TRY(cudaMemcpyAsync(sl.clx.p, a->cov_loc_x + p0, sizeof(double) * cur, cudaMemcpyHostToDevice, s));
                CU_TRY(cudaMemcpyAsync(sl.cly.p, a->cov_loc_y + p0, sizeof(double) * cur, cudaMemcpyHostToDevice, s));
            }
            const double* covp[kMaxCov];
            int64_t covs[kMaxCov];
            for (int c = 0; c < a->ncov; ++c) {
                if (is_rast[c]) {  // sample the raster at this chunk's locations, this batch's days
                    LocSpec Lc{b.loc_x, b.loc_y, a->grid_x0, a->grid_y0, a->grid_dx, a->grid_dy,
                               a->loc_x ? 1 : a->grid_ncol, a->loc_x ? 0 : b.pix_begin};
                    if (a->cov_loc_x) { Lc.x = sl.clx.as<double>(); Lc.y = sl.cly.as<double>(); }
                    const RasterDev rd = raster_dev(a->cov_raster[c], d_rast[c].p);
                    if (!cov_dynamic(c)) {
                        if (!reuse) RF_TRY(launch_raster_extract(rd, 0, 1, Lc, cur, sl.cov[c].as<double>(), cur, s));
                        covs[c] = 0;
                    } else {
                        RF_TRY(launch_raster_extract(rd, d0, nd, Lc, cur, sl.cov[c].as<double>(), cur, s));
                        covs[c] = cur;
                    }
                    covp[c] = sl.cov[c].as<double>();
                    continue;
                }
                const int64_t hs = a->cov_day_stride[c];
                if (hs == 0) {
                    if (!reuse)
                        CU_TRY(cudaMemcpyAsync(sl.cov[c].p, a->cov[c] + p0, sizeof(double) * cur, cudaMemcpyHostToDevice, s));
                    covs[c] = 0;
                } else {
                    CU_TRY(cudaMemcpy2DAsync(sl.cov[c].p, sizeof(double) * cur, a->cov[c] + (size_t)d0 * hs + p0,
                                             sizeof(double) * hs, sizeof(double) * cur, nd, cudaMemcpyHostToDevice, s));
                    covs[c] = cur;
                }
                covp[c] = sl.cov[c].as<double>();
            }
            for (int k = 0; k < a->ncov_fix; ++k) {  // repairs, on the staged columns (static ones once per chunk)
                const rfsi_cov_fix_t& x = a->cov_fix[k];
                const bool dyn = cov_dynamic(x.target);
                if (!dyn && reuse) continue;
                const long long n = (long long)cur * (dyn ? nd : 1);
                cov_fix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sl.cov[x.target].as<double>(),
                                                                            sl.cov[x.other].as<double>(), x.delta, n);
                count_launch();
                CU_TRY(cudaGetLastError());
            }
            b.cov = covp; b.cov_day_stride = covs; b.cov_raster = nullptr;
            b.cov_fix = nullptr; b.ncov_fix = 0;
            b.cov_loc_x = b.cov_loc_y = nullptr;
            b.out_mean = a->out_mean ? sl.mean.as<double>() : nullptr;
            b.out_q = a->out_q ? sl.q.as<double>() : nullptr;
            b.out_mean_i16 = a->out_mean_i16 ? sl.mean16.as<int16_t>() : nullptr;
            b.out_iqr_i16 = a->out_iqr_i16 ? sl.iqr16.as<int16_t>() : nullptr;
            FusedPlan bpl = pl;  // same offsets (sized for a full chunk); only row counts differ
            RF_TRY(dev_fused(f, &b, bpl, sl.ws.p, d_cnt.as<unsigned long long>(), device, s, reuse, false));
            sl.cand_chunk = p0;
            // results back: device [dloc*cur + p] -> host [(d0+dloc)*npix + p0 + p]
            if (a->out_mean)
                CU_TRY(cudaMemcpy2DAsync(a->out_mean + (size_t)d0 * a->npix + p0, sizeof(double) * a->npix, sl.mean.p,
                                         sizeof(double) * cur, sizeof(double) * cur, nd, cudaMemcpyDeviceToHost, s));
            if (a->out_mean_i16)
                CU_TRY(cudaMemcpy2DAsync(a->out_mean_i16 + (size_t)d0 * a->npix + p0, sizeof(int16_t) * a->npix,
                                         sl.mean16.p, sizeof(int16_t) * cur, sizeof(int16_t) * cur, nd,
                                         cudaMemcpyDeviceToHost, s));
            if (a->out_iqr_i16)
                CU_TRY(cudaMemcpy2DAsync(a->out_iqr_i16 + (size_t)d0 * a->npix + p0, sizeof(int16_t) * a->npix,
                                         sl.iqr16.p, sizeof(int16_t) * cur, sizeof(int16_t) * cur, nd,
                                         cudaMemcpyDeviceToHost, s));
            if (a->out_q)
                for (int j = 0; j < a->nprobs; ++j)
                    CU_TRY(cudaMemcpy2DAsync(a->out_q + ((size_t)j * D + d0) * a->npix + p0, sizeof(double) * a->npix,
                                             sl.q.as<double>() + (size_t)j * nd * cur, sizeof(double) * cur,
                                             sizeof(double) * cur, nd, cudaMemcpyDeviceToHost, s));
        }
    }
    tr.mark("enqueue all work items");
    for (int s = 0; s < nslots; ++s) CU_TRY(cudaStreamSynchronize(slot[s].st.s));
    tr.mark("stream sync");
    unsigned long long cnt[4];
    CU_TRY(cudaMemcpy(cnt, d_cnt.p, sizeof cnt, cudaMemcpyDeviceToHost));
    int na = 0;
    for (int s = 0; s < nslots; ++s) {
        int v = 0;
        CU_TRY(cudaMemcpy(&v, reinterpret_cast<char*>(slot[s].ws.p) + pl.off_misc + 16, sizeof v, cudaMemcpyDeviceToHost));
        na |= v;
    }
    if (na & 2) return fail(RFSI_E_NA_INPUT, "fused: a location has an NA coordinate (outputs NA there)");
    if (na) return fail(RFSI_E_NA_INPUT, "fused: missing data in a covariate column the forest uses (outputs NA there)");
    if (cnt[3]) {
        char buf[160];
        snprintf(buf, sizeof buf, "fused: %llu pixel-days had fewer valid stations than n.obs(+1); outputs NA there", cnt[3]);
        g_err = buf;
        return RFSI_W_TOO_FEW_OBS;
    }
    return RFSI_OK;
}

}  // namespace

extern "C" int rfsi_predict_fused(const rfsi_forest_t* f, const rfsi_fused_args_t* a, int device)
{
    RF_TRY(validate_fused_host(f, a));
    RF_TRY(use_device(device));
    return host_fused(f, a, device, 0, a->npix);
}

// north_star (c): the raster cut into one band of pixels per GPU, no collective on the compute
// path.  One host thread per band drives its device through host_fused() and its D2H copies land
// straight in the caller's arrays ("device-side gather to host" = ndevices independent streams).
extern "C" int rfsi_predict_fused_multi(const rfsi_forest_t* f, const rfsi_fused_args_t* a, const int* devices,
                                        int ndevices)
{
    RF_TRY(validate_fused_host(f, a));
    if (!devices || ndevices < 1 || ndevices > 64) return fail(RFSI_E_ARG, "fused_multi: needs 1..64 devices");
    for (int g = 0; g < ndevices; ++g) RF_TRY(use_device(devices[g]));
    if (ndevices == 1 || a->npix == 0) return host_fused(f, a, devices[0], 0, a->npix);
    // bands of whole raster rows where the locations are raster cells (keeps the 2-D pixel tiles)
    int64_t unit = 1;
    if (!a->loc_x && a->pix_begin % a->grid_ncol == 0) unit = a->grid_ncol;
    else if (a->loc_x && explicit_row_len(a) > 1) unit = explicit_row_len(a);
    const int64_t units = (a->npix + unit - 1) / unit;
    std::vector<int64_t> cut((size_t)ndevices + 1);
    for (int g = 0; g <= ndevices; ++g) cut[(size_t)g] = std::min<int64_t>(a->npix, units * g / ndevices * unit);
    cut[(size_t)ndevices] = a->npix;
    std::vector<int> rc((size_t)ndevices, RFSI_OK);
    std::vector<std::string> msg((size_t)ndevices);
    std::vector<std::thread> th;
    for (int g = 0; g < ndevices; ++g)
        th.emplace_back([&, g] {
            rc[(size_t)g] = host_fused(f, a, devices[g], cut[(size_t)g], cut[(size_t)g + 1]);
            if (rc[(size_t)g] != RFSI_OK) msg[(size_t)g] = g_err;   // g_err is thread-local
        });
    for (auto& t : th) t.join();
    int worst = RFSI_OK, who = -1;   // the first error, else the first warning
    for (int g = 0; g < ndevices; ++g)
        if (rc[(size_t)g] < 0 && worst >= 0) { worst = rc[(size_t)g]; who = g; }
        else if (rc[(size_t)g] > 0 && worst == RFSI_OK) { worst = rc[(size_t)g]; who = g; }
    if (who >= 0) g_err = msg[(size_t)who];
    return worst;
}
