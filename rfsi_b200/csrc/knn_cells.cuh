// knn_cells.cuh -- stage 1 for large station sets: the same exact (d2, index) search as
// knn_brute_kernel, but each tile of pixels only scans the stations of nearby grid cells.
//
// Stations are binned once into a uniform G x G cell grid (build_cells_kernel, counting
// sort).  A CTA owns a compact tile of pixels (16x8 raster patch, or 128 consecutive
// locations), finds the range of cells its pixels fall in, and scans rings of cells of
// growing Chebyshev radius r around that range.  After ring r every unvisited station is
// at least r*cell_size away from every pixel of the tile, so the tile may stop as soon as
// every pixel's k-th best distance is strictly below that bound: the result is the same
// function of the data as the full scan (ties included, because the bound is strict and
// list order is the explicit (d2, index) pair order).  When the rings cover the whole grid
// the scan was exhaustive.  Replaces nabor::knn's kd-tree for S >~ 10^3
// (simulation.R:191-196 at 5000 stations; the 10^4-station scale-up config).
#pragma once
#include "common.cuh"
#include "knn.cuh"

namespace rfsi {

// header written by build_cells_kernel and read by the search kernel
struct CellHeader {
    double x0, y0, cs, inv_cs, slack;
    int G, S;
};

__device__ __forceinline__ int cell_coord(double v, double v0, double inv_cs, int G) {
    const double t = (v - v0) * inv_cs;
    int c = (t <= 0.0) ? 0 : (t >= (double)G ? G - 1 : (int)t);
    return min(max(c, 0), G - 1);
}

// Single-CTA build: bbox -> grid -> counting sort.  counts: [G2max + 1] ints (workspace).
__global__ void __launch_bounds__(1024)
build_cells_kernel(const double2* __restrict__ xy, int S, int Gmax, CellHeader* __restrict__ hdr,
                   int32_t* __restrict__ cell_start, int32_t* __restrict__ cursor,
                   double2* __restrict__ sxy, int32_t* __restrict__ sid, float2* __restrict__ sxyf)
{
    __shared__ double red[4][32];
    __shared__ CellHeader h;
    __shared__ int scan_carry;
    const int tid = threadIdx.x, nt = blockDim.x;
    double mnx = 1e308, mxx = -1e308, mny = 1e308, mxy = -1e308;
    for (int i = tid; i < S; i += nt) {
        const double2 p = xy[i];
        mnx = fmin(mnx, p.x); mxx = fmax(mxx, p.x); mny = fmin(mny, p.y); mxy = fmax(mxy, p.y);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = mnx; red[1][tid >> 5] = mxx; red[2][tid >> 5] = mny; red[3][tid >> 5] = mxy; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < (nt >> 5); ++w) {
            red[0][0] = fmin(red[0][0], red[0][w]); red[1][0] = fmax(red[1][0], red[1][w]);
            red[2][0] = fmin(red[2][0], red[2][w]); red[3][0] = fmax(red[3][0], red[3][w]);
        }
        const double w = red[1][0] - red[0][0], hgt = red[3][0] - red[2][0];
        const double ext = fmax(fmax(w, hgt), 1e-300);
        // ~3 stations per cell on average
        int G = (int)ceil(sqrt((double)S / 3.0));
        G = max(1, min(G, Gmax));
        h.cs = ext / (double)G * (1.0 + 1e-9);
        h.inv_cs = 1.0 / h.cs;
        h.x0 = red[0][0]; h.y0 = red[2][0];
        h.G = G; h.S = S;
        const double mag = fmax(fmax(fabs(red[0][0]), fabs(red[1][0])), fmax(fabs(red[2][0]), fabs(red[3][0])));
        h.slack = 1e-9 * (mag + ext) + 1e-7 * h.cs;
        *hdr = h;
    }
    __syncthreads();
    const int G = h.G, G2 = G * G;
    for (int c = tid; c <= G2; c += nt) cell_start[c] = 0;
    __syncthreads();
    for (int i = tid; i < S; i += nt) {
        const double2 p = xy[i];
        const int c = cell_coord(p.y, h.y0, h.inv_cs, G) * G + cell_coord(p.x, h.x0, h.inv_cs, G);
        atomicAdd(&cell_start[c + 1], 1);
    }
    __syncthreads();
    // inclusive scan of cell_start[1..G2] in chunks of blockDim
    if (tid == 0) scan_carry = 0;
    __syncthreads();
    __shared__ int sbuf[1024];
    for (int base = 1; base <= G2; base += nt) {
        const int i = base + tid;
        int v = (i <= G2) ? cell_start[i] : 0;
        sbuf[tid] = v;
        __syncthreads();
        for (int o = 1; o < nt; o <<= 1) {
            const int t = (tid >= o) ? sbuf[tid - o] : 0;
            __syncthreads();
            sbuf[tid] += t;
            __syncthreads();
        }
        const int incl = sbuf[tid] + scan_carry;
        if (i <= G2) { cell_start[i] = incl; }
        __syncthreads();
        if (tid == nt - 1) scan_carry = incl;
        __syncthreads();
    }
    for (int c = tid; c < G2; c += nt) cursor[c] = cell_start[c];
    __syncthreads();
    for (int i = tid; i < S; i += nt) {
        const double2 p = xy[i];
        const int c = cell_coord(p.y, h.y0, h.inv_cs, G) * G + cell_coord(p.x, h.x0, h.inv_cs, G);
        const int slot = atomicAdd(&cursor[c], 1);
        sxy[slot] = p;
        sid[slot] = i;
        if (sxyf) sxyf[slot] = make_float2((float)p.x, (float)p.y);   // fp32 mode: the coordinates ARE floats (exact)
    }
}

// ---- explicit locations in arbitrary order: visit them along a Morton curve -----------------
// 128 consecutive locations of an unordered list can span the whole raster: the ring search
// would degenerate to a full scan and, worse, the 32 pixels of a warp would walk unrelated
// tree paths in the forest (measured 10x slower).  Unordered explicit locations are therefore
// visited in Morton (Z-curve) order of their coordinates: loc_bbox_kernel finds the bounding
// box, loc_morton_kernel builds 42-bit keys (21 bits per axis), a CUB radix sort gives the
// permutation.  Outputs still go to the original positions, so results do not change.
__device__ __forceinline__ unsigned long long dbl_ordered(double v) {   // monotone double -> uint64
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u & 0x8000000000000000ULL) ? ~u : (u | 0x8000000000000000ULL);
}
__device__ __forceinline__ double dbl_unordered(unsigned long long k) {
    const unsigned long long u = (k & 0x8000000000000000ULL) ? (k & 0x7FFFFFFFFFFFFFFFULL) : ~k;
    return __longlong_as_double((long long)u);
}

// bbox[0..3] = ordered keys of {min x, max x, min y, max y}; initialise to {~0, 0, ~0, 0}
__global__ void loc_bbox_kernel(LocSpec L, long long npix, unsigned long long* __restrict__ bbox)
{
    unsigned long long mnx = ~0ULL, mxx = 0ULL, mny = ~0ULL, mxy = 0ULL;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        double x, y;
        get_loc(L, p, x, y);
        if (x == x && y == y) {
            const unsigned long long kx = dbl_ordered(x), ky = dbl_ordered(y);
            mnx = min(mnx, kx); mxx = max(mxx, kx); mny = min(mny, ky); mxy = max(mxy, ky);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&bbox[0], mnx); atomicMax(&bbox[1], mxx); atomicMin(&bbox[2], mny); atomicMax(&bbox[3], mxy);
    }
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long v) {   // abc -> a0b0c
    v &= 0x1FFFFFULL;
    v = (v | (v << 16)) & 0x0000FFFF0000FFFFULL;
    v = (v | (v << 8)) & 0x00FF00FF00FF00FFULL;
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0FULL;
    v = (v | (v << 2)) & 0x3333333333333333ULL;
    v = (v | (v << 1)) & 0x5555555555555555ULL;
    return v;
}

__global__ void loc_morton_kernel(LocSpec L, long long npix, const unsigned long long* __restrict__ bbox,
                                  unsigned long long* __restrict__ keys, int32_t* __restrict__ vals)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const double x0 = dbl_unordered(bbox[0]), x1 = dbl_unordered(bbox[1]);
    const double y0 = dbl_unordered(bbox[2]), y1 = dbl_unordered(bbox[3]);
    const double ext = fmax(fmax(x1 - x0, y1 - y0), 1e-300);
    const double scale = 2097151.0 / ext;
    double x, y;
    get_loc(L, p, x, y);
    unsigned long long qx = 0, qy = 0;
    if (x == x && y == y) {
        qx = (unsigned long long)fmin(fmax((x - x0) * scale, 0.0), 2097151.0);
        qy = (unsigned long long)fmin(fmax((y - y0) * scale, 0.0), 2097151.0);
    }
    keys[p] = spread21(qx) | (spread21(qy) << 1);
    vals[p] = (int32_t)p;
}

template <int P>
__host__ __device__ constexpr size_t knn_cells_smem_bytes(int k) {
    return (size_t)k * P * (sizeof(double) + sizeof(int32_t));
}

// Pixel of this thread: raster form -> 16 x (P/16) tiles (same shape as the fused kernel's),
// explicit form -> consecutive locations.
template <int P>
__device__ __forceinline__ long long knn_pixel_of_thread(const LocSpec& L, long long npix, int tile2d,
                                                         long long tiles_per_row, const int32_t* __restrict__ perm) {
    if (!tile2d) {
        const long long i = (long long)blockIdx.x * P + threadIdx.x;
        return (perm != nullptr && i < npix) ? (long long)perm[i] : i;
    }
    const long long tile = blockIdx.x;
    const long long tr = tile / tiles_per_row, tc = tile - tr * tiles_per_row;
    const long long first_row = L.pix_begin / L.ncol;
    const long long col = tc * 16 + (threadIdx.x & 15);
    const long long row = first_row + tr * (P / 16) + (threadIdx.x >> 4);
    if (col >= L.ncol) return -1;
    return row * L.ncol + col - L.pix_begin;
}

// fp32 mode (PREF, rfsi_near_obs_f32; north_star's "1e-5 in fp32 mode"): every location and station coordinate is a
// float, so d2 is first formed in fp32 -- two subtractions, two products, one sum, each rounded once:
// d2f = d2 (1 + e), |e| < 2^-22 (+ 2^-148 when a product underflows) -- and a station is dropped on that value alone
// when d2f > lim = k-th best d2, rounded up to float, times (1 + 2^-20): then d2 >= d2f / (1 + 2^-22) > k-th best, the
// station can neither enter the list nor tie with its last entry.  Everything that survives is re-ranked by the exact
// fp64 d2 of the same coordinates, so neighbours and their order are bit-identical to the fp64 search; the distances
// that come out are the fp64 ones (rounded to float by the caller: 6e-8 relative).
__device__ __forceinline__ float knn_f32_limit(double thr) {
    return __fadd_ru(__fmul_ru(__double2float_ru(thr), 1.000001f), 1.0e-37f);
}

template <int P, int MODE, bool PREF = false>
__global__ void __launch_bounds__(P)
knn_cells_kernel(LocSpec L, long long npix, const CellHeader* __restrict__ hdr,
                 const int32_t* __restrict__ cell_start, const double2* __restrict__ sxy,
                 const float2* __restrict__ sxyf,
                 const int32_t* __restrict__ sid, int k, const double* __restrict__ st_z, int rm_dupl,
                 double* __restrict__ out_a, double* __restrict__ out_b, int32_t* __restrict__ out_i,
                 long long ld, int* __restrict__ warn, int* __restrict__ na_flag, int tile2d,
                 long long tiles_per_row, const int32_t* __restrict__ perm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* ld2 = reinterpret_cast<double*>(smem_raw) + threadIdx.x;
    int32_t* lid = reinterpret_cast<int32_t*>(reinterpret_cast<double*>(smem_raw) + (size_t)k * P) + threadIdx.x;
    __shared__ int s_box[4];   // cx0, cx1, cy0, cy1 of the tile
    __shared__ int s_more;

    const CellHeader h = *hdr;
    const int G = h.G;
    const long long p = knn_pixel_of_thread<P>(L, npix, tile2d, tiles_per_row, perm);
    const bool in_range = p >= 0 && p < npix;
    double px = 0.0, py = 0.0;
    if (in_range) get_loc(L, p, px, py);
    const bool bad = in_range && (px != px || py != py);
    if (bad && na_flag) atomicOr(na_flag, 2);   // bit 1: a location has an NA coordinate
    const bool act = in_range && !bad;

    const double inf = __longlong_as_double(0x7FF0000000000000LL);
    for (int j = 0; j < k; ++j) { ld2[j * P] = inf; lid[j * P] = kIdxNone; }
    double thr = inf;
    int32_t thr_id = kIdxNone;
    const float pxf = (float)px, pyf = (float)py;      // PREF: exact, the caller's coordinates are floats
    float lim = __int_as_float(0x7f800000);            // +inf: nothing is dropped before the list is full

    if (threadIdx.x == 0) { s_box[0] = G; s_box[1] = -1; s_box[2] = G; s_box[3] = -1; }
    __syncthreads();
    if (act) {
        const int cx = cell_coord(px, h.x0, h.inv_cs, G), cy = cell_coord(py, h.y0, h.inv_cs, G);
        atomicMin(&s_box[0], cx); atomicMax(&s_box[1], cx);
        atomicMin(&s_box[2], cy); atomicMax(&s_box[3], cy);
    }
    __syncthreads();
    const int cx0 = s_box[0], cx1 = s_box[1], cy0 = s_box[2], cy1 = s_box[3];
    // cx1 < cx0: no searchable pixel in this tile (uniform) -- nothing to scan, but the rows of pixels with NA
    // coordinates are still written below (empty lists: NA neighbours)
    for (int r = 0; cx1 >= cx0; ++r) {
        const int x_lo = max(cx0 - r, 0), x_hi = min(cx1 + r, G - 1);
        const int y_lo = max(cy0 - r, 0), y_hi = min(cy1 + r, G - 1);
        for (int y = y_lo; y <= y_hi; ++y) {
            const bool edge_row = (y == cy0 - r) || (y == cy1 + r);
            for (int x = x_lo; x <= x_hi; ++x) {
                // ring r only: cells not already covered by ring r-1
                if (r > 0 && !edge_row && x != cx0 - r && x != cx1 + r) continue;
                const int c = y * G + x;
                const int s_begin = __ldg(cell_start + c), s_end = __ldg(cell_start + c + 1);
                for (int s = s_begin; s < s_end; ++s) {
                    if (PREF) {
                        const float2 qf = __ldg(sxyf + s);
                        const float dxf = __fsub_rn(pxf, qf.x), dyf = __fsub_rn(pyf, qf.y);
                        const float d2f = __fadd_rn(__fmul_rn(dxf, dxf), __fmul_rn(dyf, dyf));
                        if (!act || d2f > lim) continue;                      // cannot enter the list, cannot tie
                        const double2 q = __ldg(sxy + s);
                        const int32_t id = __ldg(sid + s);
                        const double d2 = dist2_rn(px, py, q.x, q.y);         // exact re-rank of the survivors
                        if (d2 < thr || (d2 == thr && id < thr_id)) {
                            list_insert<P>(ld2, lid, k, d2, id);
                            thr = ld2[(k - 1) * P];
                            thr_id = lid[(k - 1) * P];
                            lim = knn_f32_limit(thr);
                        }
                        continue;
                    }
                    const double2 q = __ldg(sxy + s);
                    const int32_t id = __ldg(sid + s);
                    if (act) {
                        const double d2 = dist2_rn(px, py, q.x, q.y);
                        if (d2 < thr || (d2 == thr && id < thr_id)) {
                            list_insert<P>(ld2, lid, k, d2, id);
                            thr = ld2[(k - 1) * P];
                            thr_id = lid[(k - 1) * P];
                        }
                    }
                }
            }
        }
        // stop when the scan was exhaustive, or every pixel's k-th best is strictly inside the bound
        const bool whole_grid = (x_lo == 0 && y_lo == 0 && x_hi == G - 1 && y_hi == G - 1);
        const double bound = (double)r * h.cs - h.slack;
        const bool need_more = act && !whole_grid && !(bound > 0.0 && thr < bound * bound);
        __syncthreads();
        if (threadIdx.x == 0) s_more = 0;
        __syncthreads();
        if (need_more) s_more = 1;
        __syncthreads();
        if (!s_more || whole_grid) break;
    }
    if (!in_range) return;

    if (MODE == KNN_RAW) {
        for (int j = 0; j < k; ++j) {
            out_a[(long long)j * ld + p] = ld2[j * P];
            out_i[(long long)j * ld + p] = lid[j * P];
        }
    } else {
        const int n_obs = k - 1;
        const int first = (rm_dupl && lid[0] != kIdxNone && ld2[0] == 0.0) ? 1 : 0;
        bool shortl = false;
        for (int j = 0; j < n_obs; ++j) {
            const int32_t id = lid[(j + first) * P];
            const long long o = (long long)j * ld + p;
            if (id != kIdxNone) {
                out_a[o] = __dsqrt_rn(ld2[(j + first) * P]);
                out_b[o] = st_z[id];
                if (out_i) out_i[o] = id + 1;
            } else {
                out_a[o] = na_real();
                out_b[o] = na_real();
                if (out_i) out_i[o] = kNaInteger;
                shortl = true;
            }
        }
        if (shortl && warn) atomicOr(warn, 1);   // bit 0: fewer neighbours than n.obs
    }
}

}  // namespace rfsi
