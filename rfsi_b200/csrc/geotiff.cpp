// geotiff.cpp -- librfsi_io.so: GeoTIFF reader / writer for covariate and product rasters
// (include/rfsi_io.h).  Host code; replaces raster::raster(<file>) on the way in
// (prcp_catalonia/5_prcp_prediction.R:253-276, temp_croatia/dem.tif) and
// writeRaster(..., "GTiff", NAflag=-32767, datatype='INT2S') on the way out
// (5_prcp_prediction.R:303,309; temp_croatia/2_predictions.R:173,183).  Written against the
// TIFF 6.0 / BigTIFF / GeoTIFF 1.0 layouts; no GDAL or libtiff in this image.
#include "../../include/rfsi_io.h"

#include <zlib.h>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <thread>
#include <vector>

namespace {
thread_local std::string g_err;
}  // namespace

// shared with shape.cpp: one error slot per thread for the whole library
namespace rfsi_io_detail {
int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
}  // namespace rfsi_io_detail
using rfsi_io_detail::fail;

namespace {

// TIFF field types -> element size
const int kTypeSize[19] = {0, 1, 1, 2, 4, 8, 1, 1, 2, 4, 8, 4, 8, 4, 0, 0, 8, 8, 8};

enum : int {
    T_WIDTH = 256, T_HEIGHT = 257, T_BITS = 258, T_COMPRESSION = 259, T_PHOTOMETRIC = 262, T_FILLORDER = 266,
    T_STRIP_OFFSETS = 273, T_SPP = 277, T_ROWS_PER_STRIP = 278, T_STRIP_BYTES = 279, T_PLANAR = 284,
    T_PREDICTOR = 317, T_TILE_W = 322, T_TILE_H = 323, T_TILE_OFFSETS = 324, T_TILE_BYTES = 325,
    T_EXTRA_SAMPLES = 338, T_SAMPLE_FORMAT = 339, T_PIXEL_SCALE = 33550, T_TIEPOINT = 33922,
    T_TRANSFORM = 34264, T_GEOKEYS = 34735, T_GEODOUBLES = 34736, T_GEOASCII = 34737, T_GDAL_NODATA = 42113
};

struct Tag {
    int type = 0;
    uint64_t count = 0;
    std::vector<uint8_t> raw;  // host byte order
};

inline void swap_elems(uint8_t* p, size_t n, int size) {
    if (size <= 1) return;
    for (size_t i = 0; i < n; ++i, p += size) std::reverse(p, p + size);
}

}  // namespace

struct rfsi_tiff {
    std::string path;
    const uint8_t* base = nullptr;
    size_t size = 0;
    bool swap = false, big = false;
    std::map<int, Tag> tags;
    // derived
    int64_t W = 0, H = 0;
    int spp = 1, bits = 0, fmt = 1, comp = 1, pred = 1, planar = 1;
    bool tiled = false;
    int64_t tw = 0, th = 0, rps = 0;
    std::vector<uint64_t> off, cnt;
    rfsi_tiff_info_t info{};

    ~rfsi_tiff() {
        if (base) munmap(const_cast<uint8_t*>(base), size);
    }
    const Tag* tag(int id) const {
        auto it = tags.find(id);
        return it == tags.end() ? nullptr : &it->second;
    }
    bool ints(int id, std::vector<uint64_t>& out) const {
        const Tag* t = tag(id);
        if (!t) return false;
        out.resize(t->count);
        const uint8_t* p = t->raw.data();
        for (uint64_t i = 0; i < t->count; ++i) {
            switch (t->type) {
                case 1: case 7: out[i] = p[i]; break;
                case 6: out[i] = (uint64_t)(int64_t)(int8_t)p[i]; break;
                case 3: { uint16_t v; memcpy(&v, p + 2 * i, 2); out[i] = v; break; }
                case 8: { int16_t v; memcpy(&v, p + 2 * i, 2); out[i] = (uint64_t)(int64_t)v; break; }
                case 4: case 13: { uint32_t v; memcpy(&v, p + 4 * i, 4); out[i] = v; break; }
                case 9: { int32_t v; memcpy(&v, p + 4 * i, 4); out[i] = (uint64_t)(int64_t)v; break; }
                case 16: case 17: case 18: { uint64_t v; memcpy(&v, p + 8 * i, 8); out[i] = v; break; }
                default: return false;
            }
        }
        return true;
    }
    uint64_t int1(int id, uint64_t dflt) const {
        std::vector<uint64_t> v;
        return (ints(id, v) && !v.empty()) ? v[0] : dflt;
    }
    bool doubles(int id, std::vector<double>& out) const {
        const Tag* t = tag(id);
        if (!t) return false;
        out.resize(t->count);
        if (t->type == 12) memcpy(out.data(), t->raw.data(), 8 * t->count);
        else if (t->type == 11)
            for (uint64_t i = 0; i < t->count; ++i) { float v; memcpy(&v, t->raw.data() + 4 * i, 4); out[i] = v; }
        else {
            std::vector<uint64_t> iv;
            if (!ints(id, iv)) return false;
            for (uint64_t i = 0; i < t->count; ++i) out[i] = (double)(int64_t)iv[i];
        }
        return true;
    }
};

namespace {

// [at, at + n) lies inside a file of `size` bytes (no wrap-around: BigTIFF offsets are 64-bit)
inline bool in_file(uint64_t at, uint64_t n, uint64_t size) { return at <= size && n <= size - at; }

template <class T>
T rd(const rfsi_tiff& t, uint64_t at) {
    T v;
    memcpy(&v, t.base + at, sizeof(T));
    if (t.swap) std::reverse(reinterpret_cast<uint8_t*>(&v), reinterpret_cast<uint8_t*>(&v) + sizeof(T));
    return v;
}

int parse_ifd(rfsi_tiff& t) {
    if (t.size < 8) return fail(RFSI_IO_E_FORMAT, "%s: too short for a TIFF header", t.path.c_str());
    if (t.base[0] == 'I' && t.base[1] == 'I') t.swap = false;
    else if (t.base[0] == 'M' && t.base[1] == 'M') t.swap = true;
    else return fail(RFSI_IO_E_FORMAT, "%s: no II/MM byte-order mark", t.path.c_str());
    const uint16_t magic = rd<uint16_t>(t, 2);
    uint64_t ifd;
    if (magic == 42) { t.big = false; ifd = rd<uint32_t>(t, 4); }
    else if (magic == 43) {
        if (t.size < 16 || rd<uint16_t>(t, 4) != 8) return fail(RFSI_IO_E_FORMAT, "%s: bad BigTIFF header", t.path.c_str());
        t.big = true; ifd = rd<uint64_t>(t, 8);
    } else return fail(RFSI_IO_E_FORMAT, "%s: magic %u is neither TIFF (42) nor BigTIFF (43)", t.path.c_str(), magic);
    const uint64_t esz = t.big ? 20 : 12, csz = t.big ? 8 : 2, inl = t.big ? 8 : 4;
    if (!in_file(ifd, csz, t.size)) return fail(RFSI_IO_E_FORMAT, "%s: image directory beyond end of file", t.path.c_str());
    const uint64_t n = t.big ? rd<uint64_t>(t, ifd) : rd<uint16_t>(t, ifd);
    if (n > 4096 || !in_file(ifd + csz, n * esz, t.size))
        return fail(RFSI_IO_E_FORMAT, "%s: image directory with %llu entries does not fit the file", t.path.c_str(),
                    (unsigned long long)n);
    for (uint64_t e = 0; e < n; ++e) {
        const uint64_t at = ifd + csz + e * esz;
        const int id = rd<uint16_t>(t, at), type = rd<uint16_t>(t, at + 2);
        const uint64_t count = t.big ? rd<uint64_t>(t, at + 4) : rd<uint32_t>(t, at + 4);
        if (type < 1 || type > 18 || kTypeSize[type] == 0) continue;  // unknown field type: skip the entry
        if (count > (uint64_t(1) << 32)) return fail(RFSI_IO_E_FORMAT, "%s: tag %d has an absurd count", t.path.c_str(), id);
        const uint64_t bytes = count * (uint64_t)kTypeSize[type];
        uint64_t where = at + (t.big ? 12 : 8);
        if (bytes > inl) where = t.big ? rd<uint64_t>(t, where) : rd<uint32_t>(t, where);
        if (!in_file(where, bytes, t.size))
            return fail(RFSI_IO_E_FORMAT, "%s: tag %d points outside the file", t.path.c_str(), id);
        Tag tg;
        tg.type = type; tg.count = count;
        tg.raw.assign(t.base + where, t.base + where + bytes);
        if (t.swap) {
            const int es = (type == 5 || type == 10) ? 4 : kTypeSize[type];  // rationals: two 32-bit halves
            swap_elems(tg.raw.data(), bytes / es, es);
        }
        t.tags[id] = std::move(tg);
    }
    return RFSI_IO_OK;
}

int derive(rfsi_tiff& t) {
    const char* p = t.path.c_str();
    t.W = (int64_t)t.int1(T_WIDTH, 0); t.H = (int64_t)t.int1(T_HEIGHT, 0);
    if (t.W < 1 || t.H < 1) return fail(RFSI_IO_E_FORMAT, "%s: missing ImageWidth / ImageLength", p);
    if (t.W > (int64_t(1) << 31) || t.H > (int64_t(1) << 31) || (double)t.W * (double)t.H > 1.0e12)
        return fail(RFSI_IO_E_UNSUPPORTED, "%s: %lld x %lld cells is beyond what this reader takes", p, (long long)t.W, (long long)t.H);
    t.spp = (int)t.int1(T_SPP, 1);
    if (t.spp < 1 || t.spp > 65535) return fail(RFSI_IO_E_FORMAT, "%s: SamplesPerPixel %d", p, t.spp);
    std::vector<uint64_t> v;
    t.bits = 1;
    if (t.ints(T_BITS, v) && !v.empty()) {
        t.bits = (int)v[0];
        for (uint64_t b : v) if ((int)b != t.bits) return fail(RFSI_IO_E_UNSUPPORTED, "%s: bands of different bit depth", p);
    }
    t.fmt = 1;
    if (t.ints(T_SAMPLE_FORMAT, v) && !v.empty()) {
        t.fmt = (int)v[0];
        for (uint64_t b : v) if ((int)b != t.fmt) return fail(RFSI_IO_E_UNSUPPORTED, "%s: bands of different sample format", p);
    }
    if (t.fmt == 4) t.fmt = 1;  // "undefined" is read as unsigned, like GDAL
    int dt = -1;
    if (t.fmt == 3 && t.bits == 64) dt = RFSI_IO_F64;
    else if (t.fmt == 3 && t.bits == 32) dt = RFSI_IO_F32;
    else if (t.fmt == 2 && t.bits == 32) dt = RFSI_IO_I32;
    else if (t.fmt == 2 && t.bits == 16) dt = RFSI_IO_I16;
    else if (t.fmt == 1 && t.bits == 16) dt = RFSI_IO_U16;
    else if (t.fmt == 1 && t.bits == 8) dt = RFSI_IO_U8;
    else if (t.fmt == 1 && t.bits == 32) dt = RFSI_IO_U32;
    else if (t.fmt == 2 && t.bits == 8) dt = RFSI_IO_I8;
    if (dt < 0) return fail(RFSI_IO_E_UNSUPPORTED, "%s: %d-bit samples of format %d are not supported", p, t.bits, t.fmt);
    t.comp = (int)t.int1(T_COMPRESSION, 1);
    if (t.comp != 1 && t.comp != 5 && t.comp != 8 && t.comp != 32946 && t.comp != 32773)
        return fail(RFSI_IO_E_UNSUPPORTED, "%s: compression scheme %d is not supported", p, t.comp);
    t.pred = (int)t.int1(T_PREDICTOR, 1);
    if (t.pred < 1 || t.pred > 3 || (t.pred == 3 && t.fmt != 3))
        return fail(RFSI_IO_E_UNSUPPORTED, "%s: predictor %d is not supported for these samples", p, t.pred);
    if (t.int1(T_FILLORDER, 1) != 1) return fail(RFSI_IO_E_UNSUPPORTED, "%s: FillOrder 2 is not supported", p);
    t.planar = (int)t.int1(T_PLANAR, 1);
    if (t.planar != 1 && t.planar != 2) return fail(RFSI_IO_E_FORMAT, "%s: PlanarConfiguration %d", p, t.planar);
    t.tiled = t.tag(T_TILE_OFFSETS) != nullptr;
    uint64_t expect;
    if (t.tiled) {
        t.tw = (int64_t)t.int1(T_TILE_W, 0); t.th = (int64_t)t.int1(T_TILE_H, 0);
        if (t.tw < 1 || t.th < 1) return fail(RFSI_IO_E_FORMAT, "%s: tiled file without TileWidth / TileLength", p);
        if (t.tw > (1 << 20) || t.th > (1 << 20)) return fail(RFSI_IO_E_FORMAT, "%s: tile of %lld x %lld cells", p, (long long)t.tw, (long long)t.th);
        if (!t.ints(T_TILE_OFFSETS, t.off) || !t.ints(T_TILE_BYTES, t.cnt))
            return fail(RFSI_IO_E_FORMAT, "%s: tiled file without TileOffsets / TileByteCounts", p);
        expect = (uint64_t)((t.W + t.tw - 1) / t.tw) * (uint64_t)((t.H + t.th - 1) / t.th);
    } else {
        t.rps = (int64_t)std::min<uint64_t>(t.int1(T_ROWS_PER_STRIP, (uint64_t)t.H), (uint64_t)t.H);
        if (t.rps < 1) return fail(RFSI_IO_E_FORMAT, "%s: RowsPerStrip 0", p);
        if (!t.ints(T_STRIP_OFFSETS, t.off)) return fail(RFSI_IO_E_FORMAT, "%s: no StripOffsets", p);
        expect = (uint64_t)((t.H + t.rps - 1) / t.rps);
        if (!t.ints(T_STRIP_BYTES, t.cnt)) {
            // very old writers omit StripByteCounts: only recoverable for a single uncompressed strip
            if (t.comp == 1 && t.off.size() == 1) t.cnt.assign(1, t.size - t.off[0]);
            else return fail(RFSI_IO_E_FORMAT, "%s: no StripByteCounts", p);
        }
    }
    if (t.planar == 2) expect *= (uint64_t)t.spp;
    if (t.off.size() != expect || t.cnt.size() != expect)
        return fail(RFSI_IO_E_FORMAT, "%s: %zu chunk offsets / %zu byte counts, expected %llu", p, t.off.size(),
                    t.cnt.size(), (unsigned long long)expect);
    for (size_t i = 0; i < t.off.size(); ++i)
        if (!in_file(t.off[i], t.cnt[i], t.size)) return fail(RFSI_IO_E_FORMAT, "%s: chunk %zu lies outside the file", p, i);
    {   // one decoded chunk must be a sane allocation (rows x cols x samples x bytes)
        const double chunk = (double)(t.tiled ? t.th : t.rps) * (double)(t.tiled ? t.tw : t.W) * (t.planar == 1 ? t.spp : 1) * (t.bits / 8.0);
        if (chunk > 4.0e9) return fail(RFSI_IO_E_UNSUPPORTED, "%s: a single strip / tile of %.3g bytes", p, chunk);
    }

    rfsi_tiff_info_t& I = t.info;
    I.ncol = t.W; I.nrow = t.H; I.nlayers = t.spp; I.dtype = dt; I.compression = t.comp; I.predictor = t.pred;
    I.tiled = t.tiled; I.big = t.big;
    I.x_ul = 0; I.y_ul = (double)t.H; I.dx = 1; I.dy = 1;  // raster()'s extent for a file without georeference
    // georeference: ModelPixelScale + ModelTiepoint, or an axis-aligned ModelTransformation
    std::vector<double> sc, tp, mt;
    if (t.doubles(T_PIXEL_SCALE, sc) && sc.size() >= 2 && t.doubles(T_TIEPOINT, tp) && tp.size() >= 6) {
        if (!(sc[0] > 0) || !(sc[1] > 0)) return fail(RFSI_IO_E_UNSUPPORTED, "%s: non-positive pixel scale", p);
        I.dx = sc[0]; I.dy = sc[1];
        I.x_ul = tp[3] - tp[0] * sc[0];
        I.y_ul = tp[4] + tp[1] * sc[1];
        I.has_geo = 1;
    } else if (t.doubles(T_TRANSFORM, mt) && mt.size() >= 16) {
        if (mt[1] != 0 || mt[4] != 0 || !(mt[0] > 0) || !(mt[5] < 0))
            return fail(RFSI_IO_E_UNSUPPORTED, "%s: rotated or south-up rasters are not supported", p);
        I.dx = mt[0]; I.dy = -mt[5]; I.x_ul = mt[3]; I.y_ul = mt[7];
        I.has_geo = 1;
    }
    // GeoKeys: raster type (area / point) and the CRS code
    bool point = false;
    if (t.ints(T_GEOKEYS, v) && v.size() >= 4) {
        const uint64_t nk = v[3];
        for (uint64_t k = 1; k <= nk && 4 * k + 3 < v.size(); ++k) {
            const uint64_t key = v[4 * k], loc = v[4 * k + 1], val = v[4 * k + 3];
            if (loc != 0) continue;
            if (key == 1025) point = (val == 2);
            if (key == 3072) I.epsg = (int32_t)val;
            if (key == 2048 && I.epsg == 0) I.epsg = (int32_t)val;
        }
    }
    if (point && I.has_geo) { I.x_ul -= 0.5 * I.dx; I.y_ul += 0.5 * I.dy; }
    if (const Tag* nd = t.tag(T_GDAL_NODATA)) {
        std::string s(nd->raw.begin(), nd->raw.end());
        char* end = nullptr;
        const double val = strtod(s.c_str(), &end);  // accepts "nan", "-32767", "-3.4e+38"
        if (end != s.c_str()) { I.has_nodata = 1; I.nodata = val; }
    }
    return RFSI_IO_OK;
}

// ---- decompressors -------------------------------------------------------------------

// TIFF LZW (MSB-first codes, 9..12 bits, ClearCode 256, EndOfInformation 257, "early change").
// Returns the number of bytes produced (stops at dst_cap), or -1 on a damaged stream.
int64_t lzw_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t dst_cap) {
    struct Entry { uint16_t prefix; uint8_t last; uint8_t first; uint32_t len; };
    static thread_local std::vector<Entry> tab(4096 + 2);
    for (int i = 0; i < 256; ++i) tab[i] = Entry{0xFFFF, (uint8_t)i, (uint8_t)i, 1};
    uint32_t next = 258, width = 9;
    uint64_t acc = 0;
    int nbits = 0;
    size_t at = 0, out = 0;
    int prev = -1;
    while (out < dst_cap) {
        while (nbits < (int)width && at < n) { acc = (acc << 8) | src[at++]; nbits += 8; }
        if (nbits < (int)width) break;  // stream ended without EOI: tolerated
        const uint32_t code = (uint32_t)((acc >> (nbits - width)) & ((1u << width) - 1));
        nbits -= width;
        if (code == 257) break;
        if (code == 256) { next = 258; width = 9; prev = -1; continue; }
        uint32_t len;
        if (prev < 0) {
            if (code > 255) return -1;
            dst[out++] = (uint8_t)code;
            prev = (int)code;
            continue;
        }
        if (code < next) {
            len = tab[code].len;
            if (next < 4096) tab[next] = Entry{(uint16_t)prev, tab[code].first, tab[prev].first, tab[prev].len + 1};
        } else if (code == next && next < 4096) {  // KwKwK
            tab[next] = Entry{(uint16_t)prev, tab[prev].first, tab[prev].first, tab[prev].len + 1};
            len = tab[code].len;
        } else return -1;
        // write the string backwards
        const size_t room = dst_cap - out;
        uint32_t c = code, pos = len;
        while (pos > 0) {
            --pos;
            if (pos < room) dst[out + pos] = tab[c].last;
            c = tab[c].prefix;
        }
        out += std::min<size_t>(len, room);
        if (next < 4096) ++next;
        if (next + 1 >= (1u << width) && width < 12) ++width;  // early change
        prev = (int)code;
    }
    return (int64_t)out;
}

int64_t packbits_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    size_t at = 0, out = 0;
    while (at < n && out < cap) {
        const int8_t h = (int8_t)src[at++];
        if (h >= 0) {
            const size_t k = (size_t)h + 1;
            if (at + k > n) return -1;
            const size_t m = std::min(k, cap - out);
            memcpy(dst + out, src + at, m);
            at += k; out += m;
        } else if (h != -128) {
            if (at >= n) return -1;
            const size_t k = (size_t)(1 - (int)h), m = std::min(k, cap - out);
            memset(dst + out, src[at++], m);
            out += m;
        }
    }
    return (int64_t)out;
}

// one chunk (strip or tile): `rows` rows of `cols` pixels with `ns` interleaved samples
int decode_chunk(const rfsi_tiff& t, size_t ci, int64_t rows, int64_t cols, int ns, std::vector<uint8_t>& buf,
                 std::vector<uint8_t>& tmp) {
    const int B = t.bits / 8;
    const size_t want = (size_t)rows * cols * ns * B;
    buf.resize(want);
    const uint8_t* src = t.base + t.off[ci];
    const size_t n = t.cnt[ci];
    int64_t got;
    switch (t.comp) {
        case 1: got = (int64_t)std::min(n, want); memcpy(buf.data(), src, (size_t)got); break;
        case 5: got = lzw_decode(src, n, buf.data(), want); break;
        case 32773: got = packbits_decode(src, n, buf.data(), want); break;
        default: {
            uLongf dl = (uLongf)want;
            const int rc = uncompress(buf.data(), &dl, src, (uLong)n);
            got = (rc == Z_OK || rc == Z_BUF_ERROR) ? (int64_t)dl : -1;
        }
    }
    if (got < 0) return fail(RFSI_IO_E_FORMAT, "%s: chunk %zu: damaged compressed stream", t.path.c_str(), ci);
    if ((size_t)got < want) return fail(RFSI_IO_E_FORMAT, "%s: chunk %zu decodes to %lld bytes, expected %zu",
                                        t.path.c_str(), ci, (long long)got, want);
    const size_t rowb = (size_t)cols * ns * B;
    if (t.pred == 3) {  // floating-point predictor: byte-wise differences over byte planes, MSB plane first
        tmp.resize(rowb);
        const size_t ne = (size_t)cols * ns;
        for (int64_t r = 0; r < rows; ++r) {
            uint8_t* row = buf.data() + r * rowb;
            for (size_t i = (size_t)ns; i < rowb; ++i) row[i] = (uint8_t)(row[i] + row[i - ns]);
            memcpy(tmp.data(), row, rowb);
            for (size_t i = 0; i < ne; ++i)
                for (int b = 0; b < B; ++b) row[i * B + (B - 1 - b)] = tmp[(size_t)b * ne + i];  // little-endian host
        }
        return RFSI_IO_OK;
    }
    if (t.swap) swap_elems(buf.data(), want / B, B);
    if (t.pred == 2) {  // horizontal differencing per sample, modulo 2^bits
        for (int64_t r = 0; r < rows; ++r) {
            uint8_t* row = buf.data() + r * rowb;
            const size_t ne = (size_t)cols * ns;
            switch (B) {
                case 1: for (size_t i = ns; i < ne; ++i) row[i] = (uint8_t)(row[i] + row[i - ns]); break;
                case 2: { uint16_t* q = reinterpret_cast<uint16_t*>(row); for (size_t i = ns; i < ne; ++i) q[i] = (uint16_t)(q[i] + q[i - ns]); break; }
                case 4: { uint32_t* q = reinterpret_cast<uint32_t*>(row); for (size_t i = ns; i < ne; ++i) q[i] += q[i - ns]; break; }
                default: { uint64_t* q = reinterpret_cast<uint64_t*>(row); for (size_t i = ns; i < ne; ++i) q[i] += q[i - ns]; break; }
            }
        }
    }
    return RFSI_IO_OK;
}

// ---- compressors (writer) ---------------------------------------------------------------

void lzw_encode(const uint8_t* src, size_t n, std::vector<uint8_t>& out) {
    // open-addressed dictionary keyed by (prefix code << 8 | byte)
    constexpr int HS = 1 << 14;
    static thread_local std::vector<int32_t> hkey(HS), hval(HS);
    uint64_t acc = 0;
    int nbits = 0;
    uint32_t width = 9, next = 258;
    auto put = [&](uint32_t code) {
        acc = (acc << width) | code;
        nbits += (int)width;
        while (nbits >= 8) { out.push_back((uint8_t)(acc >> (nbits - 8))); nbits -= 8; }
    };
    auto reset = [&] { std::fill(hkey.begin(), hkey.end(), -1); next = 258; width = 9; };
    out.clear();
    out.reserve(n / 2 + 16);
    reset();
    put(256);
    if (n == 0) { put(257); if (nbits) out.push_back((uint8_t)(acc << (8 - nbits))); return; }
    int32_t cur = src[0];
    for (size_t i = 1; i < n; ++i) {
        const uint8_t c = src[i];
        const int32_t key = (cur << 8) | c;
        uint32_t h = ((uint32_t)key * 2654435761u) >> 18;
        bool found = false;
        while (hkey[h] >= 0) {
            if (hkey[h] == key) { cur = hval[h]; found = true; break; }
            h = (h + 1) & (HS - 1);
        }
        if (found) continue;
        put((uint32_t)cur);
        hkey[h] = key; hval[h] = (int32_t)next++;
        // the decoder adds its entry one code later, so the width switches one entry "early"
        if (next + 1 > (1u << width) && width < 12) ++width;
        if (next >= 4094) { put(256); reset(); }
        cur = c;
    }
    put((uint32_t)cur);
    // the decoder's table grows once more for this last code before it reads EOI
    ++next;
    if (next + 1 > (1u << width) && width < 12) ++width;
    put(257);
    if (nbits) out.push_back((uint8_t)(acc << (8 - nbits)));
}

int cell_bytes(int dtype) {
    switch (dtype) {
        case RFSI_IO_F64: return 8;
        case RFSI_IO_F32: case RFSI_IO_I32: case RFSI_IO_U32: return 4;
        case RFSI_IO_I16: case RFSI_IO_U16: return 2;
        case RFSI_IO_U8: case RFSI_IO_I8: return 1;
        default: return 0;
    }
}

int run_threads(int threads, size_t njobs, const std::function<int(size_t, int)>& job) {
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<size_t>(1, std::min<size_t>({(size_t)std::max(nt, 1), njobs, (size_t)64}));
    std::atomic<size_t> next{0};
    std::atomic<int> status{RFSI_IO_OK};
    std::vector<std::string> errs((size_t)nt);
    auto work = [&](int tid) {
        for (;;) {
            const size_t j = next.fetch_add(1);
            if (j >= njobs || status.load() != RFSI_IO_OK) return;
            int rc;
            try {
                rc = job(j, tid);
            } catch (const std::exception& e) {   // e.g. bad_alloc: never let it leave a worker thread
                rc = fail(RFSI_IO_E_FORMAT, "job %zu: %s", j, e.what());
            }
            if (rc != RFSI_IO_OK) { errs[(size_t)tid] = g_err; status.store(rc); return; }
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < nt; ++i) th.emplace_back(work, i);
        for (auto& x : th) x.join();
    }
    if (status.load() != RFSI_IO_OK)
        for (auto& e : errs) if (!e.empty()) { g_err = e; break; }
    return status.load();
}

}  // namespace

extern "C" const char* rfsi_io_last_error(void) { return g_err.c_str(); }

extern "C" int rfsi_tiff_open(const char* path, rfsi_tiff_t** out) {
    if (!out) return fail(RFSI_IO_E_ARG, "tiff_open: out is NULL");
    *out = nullptr;
    if (!path) return fail(RFSI_IO_E_ARG, "tiff_open: path is NULL");
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return fail(RFSI_IO_E_OPEN, "cannot open %s: %s", path, strerror(errno));
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size <= 0) { ::close(fd); return fail(RFSI_IO_E_OPEN, "%s: empty or unreadable", path); }
    void* m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (m == MAP_FAILED) return fail(RFSI_IO_E_OPEN, "cannot map %s: %s", path, strerror(errno));
    rfsi_tiff* t = new (std::nothrow) rfsi_tiff;
    if (!t) { munmap(m, (size_t)sb.st_size); return fail(RFSI_IO_E_OPEN, "%s: out of memory", path); }
    t->base = static_cast<const uint8_t*>(m); t->size = (size_t)sb.st_size;
    int rc;
    try {
        t->path = path;
        rc = parse_ifd(*t);
        if (rc == RFSI_IO_OK) rc = derive(*t);
    } catch (const std::exception& e) {   // no C++ exception may cross the C ABI
        rc = fail(RFSI_IO_E_FORMAT, "%s: %s while parsing", path, e.what());
    }
    if (rc != RFSI_IO_OK) { delete t; return rc; }
    *out = t;
    return RFSI_IO_OK;
}

extern "C" void rfsi_tiff_close(rfsi_tiff_t* t) { delete t; }

extern "C" int rfsi_tiff_get_info(const rfsi_tiff_t* t, rfsi_tiff_info_t* info) {
    if (!t || !info) return fail(RFSI_IO_E_ARG, "tiff_get_info: NULL argument");
    *info = t->info;
    return RFSI_IO_OK;
}

extern "C" int64_t rfsi_tiff_tag(const rfsi_tiff_t* t, int32_t tag, int32_t* field_type, void* out, size_t cap_bytes) {
    if (!t) return fail(RFSI_IO_E_ARG, "tiff_tag: NULL handle");
    const Tag* g = t->tag(tag);
    if (!g) { if (field_type) *field_type = 0; return 0; }
    if (field_type) *field_type = g->type;
    if (out && cap_bytes) memcpy(out, g->raw.data(), std::min(cap_bytes, g->raw.size()));
    return (int64_t)g->count;
}

namespace {
int tiff_read_impl(const rfsi_tiff_t* tp, void* out, size_t out_bytes, int threads);
int tiff_write_impl(const char* path, const rfsi_tiff_write_args_t* a);
}  // namespace

// no C++ exception may cross the C ABI
extern "C" int rfsi_tiff_read(const rfsi_tiff_t* tp, void* out, size_t out_bytes, int threads) {
    try {
        return tiff_read_impl(tp, out, out_bytes, threads);
    } catch (const std::exception& e) {
        return fail(RFSI_IO_E_FORMAT, "tiff_read: %s", e.what());
    }
}
extern "C" int rfsi_tiff_write(const char* path, const rfsi_tiff_write_args_t* a) {
    try {
        return tiff_write_impl(path, a);
    } catch (const std::exception& e) {
        return fail(RFSI_IO_E_ARG, "tiff_write: %s", e.what());
    }
}

namespace {
int tiff_read_impl(const rfsi_tiff_t* tp, void* out, size_t out_bytes, int threads) {
    if (!tp || !out) return fail(RFSI_IO_E_ARG, "tiff_read: NULL argument");
    const rfsi_tiff& t = *tp;
    const int B = t.bits / 8;
    const size_t plane = (size_t)t.W * t.H;
    if (out_bytes != plane * t.spp * B)
        return fail(RFSI_IO_E_ARG, "tiff_read: out_bytes %zu, the image needs %zu", out_bytes, plane * t.spp * B);
    uint8_t* dst = static_cast<uint8_t*>(out);
    const int ns = t.planar == 1 ? t.spp : 1;  // samples interleaved inside a chunk
    const int64_t cw = t.tiled ? t.tw : t.W, ch = t.tiled ? t.th : t.rps;
    const int64_t across = (t.W + cw - 1) / cw, down = (t.H + ch - 1) / ch;
    const size_t per_band = (size_t)across * down;
    std::vector<std::vector<uint8_t>> bufs, tmps;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, 64));
    bufs.resize((size_t)nt); tmps.resize((size_t)nt);
    return run_threads(nt, t.off.size(), [&](size_t ci, int tid) -> int {
        const size_t band0 = t.planar == 2 ? ci / per_band : 0, k = ci % per_band;
        const int64_t ty = (int64_t)(k / across), tx = (int64_t)(k % across);
        const int64_t r0 = ty * ch, c0 = tx * cw;
        // strips hold only the rows that exist; tiles are always full size
        const int64_t rows = t.tiled ? ch : std::min(ch, t.H - r0);
        const int rc = decode_chunk(t, ci, rows, cw, ns, bufs[(size_t)tid], tmps[(size_t)tid]);
        if (rc != RFSI_IO_OK) return rc;
        const uint8_t* src = bufs[(size_t)tid].data();
        const int64_t vr = std::min(rows, t.H - r0), vc = std::min(cw, t.W - c0);
        for (int s = 0; s < ns; ++s) {
            uint8_t* band = dst + (band0 + (size_t)s) * plane * B;
            for (int64_t r = 0; r < vr; ++r) {
                uint8_t* o = band + ((size_t)(r0 + r) * t.W + c0) * B;
                const uint8_t* in = src + ((size_t)r * cw * ns + s) * B;
                if (ns == 1) memcpy(o, in, (size_t)vc * B);
                else for (int64_t c = 0; c < vc; ++c) memcpy(o + c * B, in + (size_t)c * ns * B, B);
            }
        }
        return RFSI_IO_OK;
    });
}

}  // namespace

// ---- writer --------------------------------------------------------------------------

namespace {

struct OutTag {
    uint16_t id, type;
    uint64_t count;
    std::vector<uint8_t> bytes;
};

template <class T>
OutTag make_tag(uint16_t id, uint16_t type, const std::vector<T>& v) {
    OutTag o{id, type, v.size(), {}};
    o.bytes.resize(v.size() * sizeof(T));
    memcpy(o.bytes.data(), v.data(), o.bytes.size());
    return o;
}

bool write_all(FILE* f, const void* p, size_t n) { return n == 0 || fwrite(p, 1, n, f) == n; }

int tiff_write_impl(const char* path, const rfsi_tiff_write_args_t* a) {
    if (!path || !a) return fail(RFSI_IO_E_ARG, "tiff_write: NULL argument");
    if (a->struct_size != sizeof(rfsi_tiff_write_args_t))
        return fail(RFSI_IO_E_ARG, "tiff_write: struct_size %zu != %zu (header/library mismatch)", a->struct_size,
                    sizeof(rfsi_tiff_write_args_t));
    const int B = cell_bytes(a->dtype);
    if (!a->data || B == 0 || a->ncol < 1 || a->nrow < 1 || a->nlayers < 1 || a->nlayers > 65535)
        return fail(RFSI_IO_E_ARG, "tiff_write: needs data, a known dtype, ncol, nrow >= 1 and 1..65535 layers");
    if (!(a->dx > 0) || !(a->dy > 0)) return fail(RFSI_IO_E_ARG, "tiff_write: dx and dy must be > 0");
    if (a->compression != 1 && a->compression != 5 && a->compression != 8)
        return fail(RFSI_IO_E_ARG, "tiff_write: compression must be 1 (none), 5 (LZW) or 8 (Deflate)");
    const int64_t W = a->ncol, H = a->nrow;
    const int L = a->nlayers;
    const size_t rowb = (size_t)W * B;
    const int64_t rps = std::max<int64_t>(1, std::min<int64_t>(H, (int64_t)(262144 / rowb)));
    const int64_t spb = (H + rps - 1) / rps;  // strips per band
    const size_t nstrips = (size_t)spb * L;
    const uint8_t* src = static_cast<const uint8_t*>(a->data);
    auto strip_src = [&](size_t s, size_t& bytes) {
        const size_t band = s / spb;
        const int64_t r0 = (int64_t)(s % spb) * rps, rows = std::min(rps, H - r0);
        bytes = (size_t)rows * rowb;
        return src + (band * (size_t)H + (size_t)r0) * rowb;
    };
    // compress the strips (in parallel), or point at the caller's rows
    std::vector<std::vector<uint8_t>> packed;
    if (a->compression != 1) {
        packed.resize(nstrips);
        const int rc = run_threads(0, nstrips, [&](size_t s, int) -> int {
            size_t n;
            const uint8_t* p = strip_src(s, n);
            if (a->compression == 5) lzw_encode(p, n, packed[s]);
            else {
                uLongf dl = compressBound((uLong)n);
                packed[s].resize(dl);
                if (compress2(packed[s].data(), &dl, p, (uLong)n, 6) != Z_OK) return fail(RFSI_IO_E_ARG, "tiff_write: deflate failed");
                packed[s].resize(dl);
            }
            return RFSI_IO_OK;
        });
        if (rc != RFSI_IO_OK) return rc;
    }
    uint64_t data_bytes = 0;
    std::vector<uint64_t> counts(nstrips);
    for (size_t s = 0; s < nstrips; ++s) {
        size_t n;
        strip_src(s, n);
        counts[s] = a->compression == 1 ? n : packed[s].size();
        data_bytes += counts[s] + (counts[s] & 1);
    }
    // BigTIFF above 4 GB (RFSI_IO_FORCE_BIGTIFF=1 forces it, so that the tests can cover the layout)
    const char* force = getenv("RFSI_IO_FORCE_BIGTIFF");
    const bool big = data_bytes + nstrips * 16 + 4096 > 0xFFFF0000ull || (force && force[0] == '1');
    const uint64_t hdr = big ? 16 : 8;
    std::vector<uint64_t> offs(nstrips);
    uint64_t at = hdr;
    for (size_t s = 0; s < nstrips; ++s) { offs[s] = at; at += counts[s] + (counts[s] & 1); }

    // the directory
    std::vector<OutTag> tags;
    auto u16v = [](std::initializer_list<uint16_t> l) { return std::vector<uint16_t>(l); };
    auto offtag = [&](uint16_t id, const std::vector<uint64_t>& v) {
        if (big) return make_tag<uint64_t>(id, 16, v);
        std::vector<uint32_t> w(v.begin(), v.end());
        return make_tag<uint32_t>(id, 4, w);
    };
    tags.push_back(offtag(T_WIDTH, {(uint64_t)W}));
    tags.push_back(offtag(T_HEIGHT, {(uint64_t)H}));
    tags.push_back(make_tag<uint16_t>(T_BITS, 3, std::vector<uint16_t>((size_t)L, (uint16_t)(B * 8))));
    tags.push_back(make_tag<uint16_t>(T_COMPRESSION, 3, u16v({(uint16_t)a->compression})));
    tags.push_back(make_tag<uint16_t>(T_PHOTOMETRIC, 3, u16v({1})));
    tags.push_back(offtag(T_STRIP_OFFSETS, offs));
    tags.push_back(make_tag<uint16_t>(T_SPP, 3, u16v({(uint16_t)L})));
    tags.push_back(offtag(T_ROWS_PER_STRIP, {(uint64_t)rps}));
    tags.push_back(offtag(T_STRIP_BYTES, counts));
    tags.push_back(make_tag<uint16_t>(T_PLANAR, 3, u16v({(uint16_t)(L > 1 ? 2 : 1)})));
    if (L > 1) tags.push_back(make_tag<uint16_t>(T_EXTRA_SAMPLES, 3, std::vector<uint16_t>((size_t)L - 1, 0)));
    const uint16_t fmt = (a->dtype == RFSI_IO_F64 || a->dtype == RFSI_IO_F32) ? 3
                         : (a->dtype == RFSI_IO_I32 || a->dtype == RFSI_IO_I16 || a->dtype == RFSI_IO_I8) ? 2 : 1;
    tags.push_back(make_tag<uint16_t>(T_SAMPLE_FORMAT, 3, std::vector<uint16_t>((size_t)L, fmt)));
    tags.push_back(make_tag<double>(T_PIXEL_SCALE, 12, {a->dx, a->dy, 0.0}));
    tags.push_back(make_tag<double>(T_TIEPOINT, 12, {0.0, 0.0, 0.0, a->x_ul, a->y_ul, 0.0}));
    if (a->geokeys && a->ngeokeys >= 4) {
        tags.push_back(make_tag<uint16_t>(T_GEOKEYS, 3, std::vector<uint16_t>(a->geokeys, a->geokeys + a->ngeokeys)));
        if (a->geo_doubles && a->ngeo_doubles > 0)
            tags.push_back(make_tag<double>(T_GEODOUBLES, 12, std::vector<double>(a->geo_doubles, a->geo_doubles + a->ngeo_doubles)));
        if (a->geo_ascii && a->geo_ascii[0]) {
            const std::string s(a->geo_ascii);
            tags.push_back(make_tag<char>(T_GEOASCII, 2, std::vector<char>(s.c_str(), s.c_str() + s.size() + 1)));
        }
    } else {
        // GTRasterTypeGeoKey = RasterPixelIsArea, plus the CRS code when one was given
        std::vector<uint16_t> k = {1, 1, 0, 1, 1025, 0, 1, 1};
        if (a->epsg > 0 && a->epsg < 65536) {
            const bool geographic = a->epsg >= 4000 && a->epsg < 5000;
            k = {1, 1, 0, 3, 1024, 0, 1, (uint16_t)(geographic ? 2 : 1), 1025, 0, 1, 1,
                 (uint16_t)(geographic ? 2048 : 3072), 0, 1, (uint16_t)a->epsg};
        }
        tags.push_back(make_tag<uint16_t>(T_GEOKEYS, 3, k));
    }
    if (a->has_nodata) {
        char buf[64];
        if (a->nodata == std::floor(a->nodata) && std::fabs(a->nodata) < 1e15) snprintf(buf, sizeof buf, "%.0f", a->nodata);
        else snprintf(buf, sizeof buf, "%.17g", a->nodata);
        tags.push_back(make_tag<char>(T_GDAL_NODATA, 2, std::vector<char>(buf, buf + strlen(buf) + 1)));
    }
    std::sort(tags.begin(), tags.end(), [](const OutTag& x, const OutTag& y) { return x.id < y.id; });

    // out-of-line tag values follow the pixel data, then the directory
    const uint64_t inl = big ? 8 : 4;
    std::vector<uint64_t> where(tags.size(), 0);
    for (size_t i = 0; i < tags.size(); ++i)
        if (tags[i].bytes.size() > inl) { where[i] = at; at += tags[i].bytes.size() + (tags[i].bytes.size() & 1); }
    const uint64_t ifd = at;
    if (!big && ifd + 2 + tags.size() * 12 + 4 > 0xFFFFFFFFull) return fail(RFSI_IO_E_ARG, "tiff_write: size estimate failed");

    FILE* f = fopen(path, "wb");
    if (!f) return fail(RFSI_IO_E_OPEN, "cannot create %s: %s", path, strerror(errno));
    bool ok = true;
    const uint8_t zero = 0;
    if (big) {
        const uint8_t h[8] = {'I', 'I', 43, 0, 8, 0, 0, 0};
        ok = ok && write_all(f, h, 8) && write_all(f, &ifd, 8);
    } else {
        const uint8_t h[4] = {'I', 'I', 42, 0};
        const uint32_t o = (uint32_t)ifd;
        ok = ok && write_all(f, h, 4) && write_all(f, &o, 4);
    }
    for (size_t s = 0; s < nstrips && ok; ++s) {
        size_t n;
        const uint8_t* p = strip_src(s, n);
        ok = a->compression == 1 ? write_all(f, p, n) : write_all(f, packed[s].data(), packed[s].size());
        if (counts[s] & 1) ok = ok && write_all(f, &zero, 1);
    }
    for (size_t i = 0; i < tags.size() && ok; ++i)
        if (where[i]) {
            ok = write_all(f, tags[i].bytes.data(), tags[i].bytes.size());
            if (tags[i].bytes.size() & 1) ok = ok && write_all(f, &zero, 1);
        }
    if (ok) {
        if (big) { const uint64_t n = tags.size(); ok = write_all(f, &n, 8); }
        else { const uint16_t n = (uint16_t)tags.size(); ok = write_all(f, &n, 2); }
    }
    for (size_t i = 0; i < tags.size() && ok; ++i) {
        uint8_t e[20] = {0};
        memcpy(e, &tags[i].id, 2); memcpy(e + 2, &tags[i].type, 2);
        if (big) memcpy(e + 4, &tags[i].count, 8);
        else { const uint32_t c = (uint32_t)tags[i].count; memcpy(e + 4, &c, 4); }
        uint8_t* val = e + (big ? 12 : 8);
        if (where[i]) memcpy(val, &where[i], inl);
        else memcpy(val, tags[i].bytes.data(), tags[i].bytes.size());
        ok = write_all(f, e, big ? 20 : 12);
    }
    const uint64_t nextifd = 0;
    ok = ok && write_all(f, &nextifd, big ? 8 : 4);
    if (fclose(f) != 0) ok = false;
    if (!ok) return fail(RFSI_IO_E_OPEN, "short write to %s: %s", path, strerror(errno));
    return RFSI_IO_OK;
}

}  // namespace
