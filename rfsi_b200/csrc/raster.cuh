// raster.cuh -- covariate sampling: raster::extract(r, points), method "simple" (the value
// of the cell a point falls in), the step right before near.obs in the mapping loops
// (prcp_catalonia/5_prcp_prediction.R:253-276).  One thread per
// location; rasters of any of six cell types are read in place (no host-side conversion to
// double), scaled (e.g. tenths of a degree -> degrees) and NA-masked on the fly.
#pragma once
#include "common.cuh"
#include "knn.cuh"

namespace rfsi {

struct RasterDev {
    const void* data;
    int dtype, nlayers;
    long long ncol, nrow;
    double x_ul, y_ul, dx, dy, nodata, scale, offset;
};

__device__ __forceinline__ double raster_raw(const RasterDev& r, long long i) {
    switch (r.dtype & 0xff) {
        case 0: return reinterpret_cast<const double*>(r.data)[i];
        case 1: return (double)reinterpret_cast<const float*>(r.data)[i];
        case 2: return (double)reinterpret_cast<const int32_t*>(r.data)[i];
        case 3: return (double)reinterpret_cast<const int16_t*>(r.data)[i];
        case 4: return (double)reinterpret_cast<const uint16_t*>(r.data)[i];
        default: return (double)reinterpret_cast<const uint8_t*>(r.data)[i];
    }
}

__device__ __forceinline__ double raster_value(const RasterDev& r, int layer, double x, double y) {
    // raster::cellFromXY: col = floor((x - xmin)/xres), the right/bottom outer edges belong to
    // the last column/row
    const double xmax = r.x_ul + (double)r.ncol * r.dx, ymin = r.y_ul - (double)r.nrow * r.dy;
    if (!(x >= r.x_ul && x <= xmax && y <= r.y_ul && y >= ymin)) return na_real();
    long long c = (long long)floor((x - r.x_ul) / r.dx);
    long long w = (long long)floor((r.y_ul - y) / r.dy);
    if (c >= r.ncol) c = r.ncol - 1;
    if (w >= r.nrow) w = r.nrow - 1;
    const double raw = raster_raw(r, ((long long)layer * r.nrow + w) * r.ncol + c);
    if (raw != raw || raw == r.nodata) return na_real();
    // two roundings, like the CPU (no FMA contraction); RFSI_RASTER_DIVIDE: R's `extract(r, pts) / 10`
    const double v = (r.dtype & 0x100) ? __ddiv_rn(raw, r.scale) : __dmul_rn(raw, r.scale);
    return __dadd_rn(v, r.offset);
}

// out[(l - layer0)*ld + p] for layers [layer0, layer0 + nl) (a static raster: nl = 1, layer0 = 0)
__global__ void raster_extract_kernel(RasterDev r, int layer0, int nl, LocSpec L, long long npix,
                                      double* __restrict__ out, long long ld)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    double x, y;
    get_loc(L, p, x, y);
    for (int l = 0; l < nl; ++l) out[(long long)l * ld + p] = raster_value(r, layer0 + l, x, y);
}

// df$t <- ifelse(df$t > df$o, df$o - delta, df$t) over n cells (5_prcp_prediction.R:258)
__global__ void cov_fix_kernel(double* __restrict__ t, const double* __restrict__ o, double delta, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = t[i], b = o[i];
    if (a != a || b != b) t[i] = na_real();          // ifelse(NA, ...) is NA
    else if (a > b) t[i] = __dsub_rn(b, delta);
}

}  // namespace rfsi
