// shape.cpp -- librfsi_io.so: polygon shapefiles and the pixel mask they define (include/rfsi_io.h).
// Host code; replaces readOGR(<border>.shp) and raster::mask(p, border) of the mapping scripts
// (temp_croatia/2_predictions.R:27-28,108-109).  Written against the ESRI Shapefile Technical
// Description (1998): 100-byte header, big-endian record headers, little-endian shape contents.
#include "../../include/rfsi_io.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

// the error slot lives in geotiff.cpp (one message per thread for the whole library)
extern "C" const char* rfsi_io_last_error(void);
namespace rfsi_io_detail { int fail(int code, const char* fmt, ...); }
using rfsi_io_detail::fail;

namespace {

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline int32_t le32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }
inline double le64(const uint8_t* p) { double v; memcpy(&v, p, 8); return v; }

struct Mapped {
    const uint8_t* p = nullptr;
    size_t n = 0;
    ~Mapped() { if (p) munmap(const_cast<uint8_t*>(p), n); }
};

int shp_read_impl(const char* path, double** xy_out, int64_t* npoints, int64_t** ring_out, int64_t* nrings) {
    if (!path || !xy_out || !npoints || !ring_out || !nrings) return fail(RFSI_IO_E_ARG, "shp_read_polygons: NULL argument");
    *xy_out = nullptr; *ring_out = nullptr; *npoints = 0; *nrings = 0;
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return fail(RFSI_IO_E_OPEN, "cannot open %s: %s", path, strerror(errno));
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size < 100) { ::close(fd); return fail(RFSI_IO_E_FORMAT, "%s: too short for a shapefile", path); }
    Mapped m;
    void* mp = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (mp == MAP_FAILED) return fail(RFSI_IO_E_OPEN, "cannot map %s: %s", path, strerror(errno));
    m.p = static_cast<const uint8_t*>(mp); m.n = (size_t)sb.st_size;
    if (be32(m.p) != 9994) return fail(RFSI_IO_E_FORMAT, "%s: file code %u is not a shapefile's 9994", path, be32(m.p));
    const int file_type = le32(m.p + 32);
    if (file_type != 5 && file_type != 15 && file_type != 25)
        return fail(RFSI_IO_E_UNSUPPORTED, "%s: shape type %d; only polygons (5, 15, 25) define a mask", path, file_type);
    std::vector<double> xy;
    std::vector<int64_t> rings(1, 0);
    size_t at = 100;
    while (at + 8 <= m.n) {
        const uint64_t bytes = (uint64_t)be32(m.p + at + 4) * 2;     // content length is in 16-bit words
        const size_t c = at + 8;
        if (bytes > m.n - c) return fail(RFSI_IO_E_FORMAT, "%s: record at byte %zu runs past the end of the file", path, at);
        at = c + bytes;
        if (bytes < 4) continue;
        const int type = le32(m.p + c);
        if (type == 0) continue;                                      // null shape
        if (type != 5 && type != 15 && type != 25) return fail(RFSI_IO_E_UNSUPPORTED, "%s: record of shape type %d", path, type);
        if (bytes < 44) return fail(RFSI_IO_E_FORMAT, "%s: polygon record of %llu bytes", path, (unsigned long long)bytes);
        const int64_t nparts = le32(m.p + c + 36), npts = le32(m.p + c + 40);
        if (nparts < 0 || npts < 0 || (uint64_t)44 + 4 * (uint64_t)nparts + 16 * (uint64_t)npts > bytes)
            return fail(RFSI_IO_E_FORMAT, "%s: polygon record with %lld parts / %lld points does not fit its %llu bytes", path,
                        (long long)nparts, (long long)npts, (unsigned long long)bytes);
        const uint8_t* parts = m.p + c + 44;
        const uint8_t* pts = parts + 4 * nparts;
        const int64_t base = (int64_t)xy.size() / 2;
        for (int64_t k = 0; k < nparts; ++k) {
            const int64_t a = le32(parts + 4 * k), b = (k + 1 < nparts) ? le32(parts + 4 * (k + 1)) : npts;
            if (a < 0 || b < a || b > npts) return fail(RFSI_IO_E_FORMAT, "%s: part %lld spans points %lld..%lld of %lld", path,
                                                        (long long)k, (long long)a, (long long)b, (long long)npts);
            if (k > 0) rings.push_back(base + a);
        }
        for (int64_t i = 0; i < npts; ++i) { xy.push_back(le64(pts + 16 * i)); xy.push_back(le64(pts + 16 * i + 8)); }
        rings.push_back(base + npts);
        if (nparts == 0 && npts > 0) { /* a record without a parts table: one ring */ }
    }
    double* xo = static_cast<double*>(malloc(std::max<size_t>(xy.size(), 1) * sizeof(double)));
    int64_t* ro = static_cast<int64_t*>(malloc(rings.size() * sizeof(int64_t)));
    if (!xo || !ro) { free(xo); free(ro); return fail(RFSI_IO_E_OPEN, "%s: out of memory", path); }
    if (!xy.empty()) memcpy(xo, xy.data(), xy.size() * sizeof(double));
    memcpy(ro, rings.data(), rings.size() * sizeof(int64_t));
    *xy_out = xo; *ring_out = ro; *npoints = (int64_t)xy.size() / 2; *nrings = (int64_t)rings.size() - 1;
    return RFSI_IO_OK;
}

int polygon_mask_impl(const double* xy, const int64_t* ring_start, int64_t nrings, double x_ul, double y_ul, double dx,
                      double dy, int64_t ncol, int64_t nrow, uint8_t* out) {
    if (!xy || !ring_start || !out || nrings < 0 || ncol < 1 || nrow < 1 || !(dx > 0) || !(dy > 0))
        return fail(RFSI_IO_E_ARG, "polygon_mask: needs rings, a grid with ncol, nrow >= 1 and dx, dy > 0, and an output");
    memset(out, 0, (size_t)ncol * (size_t)nrow);
    // scan-line fill on the rows of cell centres: every edge leaves one crossing in each row whose centre line
    // yc satisfies ylo <= yc < yhi (half-open, so a vertex on the line counts once)
    std::vector<std::vector<double>> cross((size_t)nrow);
    for (int64_t r = 0; r < nrings; ++r) {
        const int64_t a = ring_start[r], b = ring_start[r + 1];
        if (b - a < 3) continue;
        for (int64_t i = a; i < b; ++i) {
            const int64_t j = (i + 1 < b) ? i + 1 : a;                // closes the ring if the file did not repeat the first vertex
            double x1 = xy[2 * i], y1 = xy[2 * i + 1], x2 = xy[2 * j], y2 = xy[2 * j + 1];
            if (!(std::isfinite(x1) && std::isfinite(y1) && std::isfinite(x2) && std::isfinite(y2)))
                return fail(RFSI_IO_E_ARG, "polygon_mask: NaN or infinite vertex in ring %lld", (long long)r);
            if (y1 == y2) continue;
            if (y1 > y2) { std::swap(x1, x2); std::swap(y1, y2); }   // y1 < y2
            // rows whose centre yc = y_ul - (row + 0.5) dy lies in [y1, y2)
            double r_hi = std::floor((y_ul - y1) / dy - 0.5), r_lo = std::ceil((y_ul - y2) / dy - 0.5);
            if (r_hi < 0 || r_lo > (double)(nrow - 1)) continue;
            int64_t lo = (int64_t)std::max(r_lo, 0.0), hi = (int64_t)std::min(r_hi, (double)(nrow - 1));
            for (int64_t row = std::max<int64_t>(lo - 1, 0); row <= std::min<int64_t>(hi + 1, nrow - 1); ++row) {
                const double yc = y_ul - ((double)row + 0.5) * dy;
                if (yc >= y1 && yc < y2) cross[(size_t)row].push_back(x1 + (yc - y1) * (x2 - x1) / (y2 - y1));
            }
        }
    }
    for (int64_t row = 0; row < nrow; ++row) {
        std::vector<double>& c = cross[(size_t)row];
        if (c.size() < 2) continue;
        std::sort(c.begin(), c.end());
        uint8_t* o = out + (size_t)row * (size_t)ncol;
        for (size_t k = 0; k + 1 < c.size(); k += 2) {
            // cells whose centre x_ul + (col + 0.5) dx lies in [c[k], c[k+1])
            const double f0 = std::ceil((c[k] - x_ul) / dx - 0.5), f1 = std::ceil((c[k + 1] - x_ul) / dx - 0.5);
            if (!(f0 == f0 && f1 == f1)) continue;                          // overflowed crossing: no cells
            const int64_t c0 = (int64_t)std::min(std::max(f0, 0.0), (double)ncol);   // clamp as doubles, then convert
            const int64_t c1 = (int64_t)std::min(std::max(f1, 0.0), (double)ncol);
            for (int64_t col = c0; col < c1; ++col) o[col] = 1;
        }
    }
    return RFSI_IO_OK;
}

}  // namespace

extern "C" int rfsi_shp_read_polygons(const char* path, double** xy, int64_t* npoints, int64_t** ring_start, int64_t* nrings) {
    try {
        return shp_read_impl(path, xy, npoints, ring_start, nrings);
    } catch (const std::exception& e) {
        return fail(RFSI_IO_E_FORMAT, "shp_read_polygons: %s", e.what());
    }
}

extern "C" void rfsi_io_free(void* p) { free(p); }

extern "C" int rfsi_polygon_mask(const double* xy, const int64_t* ring_start, int64_t nrings, double x_ul, double y_ul,
                                 double dx, double dy, int64_t ncol, int64_t nrow, uint8_t* out) {
    try {
        return polygon_mask_impl(xy, ring_start, nrings, x_ul, y_ul, dx, dy, ncol, nrow, out);
    } catch (const std::exception& e) {
        return fail(RFSI_IO_E_ARG, "polygon_mask: %s", e.what());
    }
}
