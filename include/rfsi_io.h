/*
 * rfsi_io.h -- C ABI of the host-side raster I/O that sits either side of the hot path.
 *
 * The reference's mapping loops read their covariates from GeoTIFF files with
 * raster::raster() (prcp_catalonia/5_prcp_prediction.R:253-276: ../min/<day>.tif,
 * ../max/<day>.tif, ../imerg/<name>.tif; temp_croatia/dem.tif) and write every product with
 *   writeRaster(p, <name>, "GTiff", NAflag=-32767, datatype='INT2S', overwrite=T)
 * (5_prcp_prediction.R:303,309; temp_croatia/2_predictions.R:173,183; simulation/sim_500.R:163).
 * GDAL is not part of this build; librfsi_io.so is a self-contained reader/writer for the
 * subset those files use (and what GDAL writes by default): classic TIFF and BigTIFF, either
 * byte order, strips or tiles, 8/16/32/64-bit integer and IEEE samples, compression none /
 * LZW / Deflate / PackBits, predictors 1-3, chunky or planar bands, and the GeoTIFF tags
 * ModelPixelScale / ModelTiepoint / GeoKeyDirectory / GDAL_NODATA.  SURVEY.md 8(f) rank 1.
 *
 * Host code only (no CUDA): the arrays it returns are what rfsi_raster_t (rfsi_b200.h) takes
 * in place, and what rfsi_predict_fused's Int16 products are written from.
 * Functions return 0 or a negative RFSI_IO_E_* code; rfsi_io_last_error() has the message.
 */
#ifndef RFSI_IO_H
#define RFSI_IO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    RFSI_IO_OK = 0,
    RFSI_IO_E_OPEN = -1,        /* cannot open / create / map the file */
    RFSI_IO_E_FORMAT = -2,      /* not a TIFF, or a damaged one */
    RFSI_IO_E_UNSUPPORTED = -3, /* valid TIFF outside the subset above (JPEG, 1-bit, rotated ...) */
    RFSI_IO_E_ARG = -4
};

/* cell types: 0..5 are rfsi_raster_t's RFSI_F64..RFSI_U8 (usable on the device as they are);
 * 6 and 7 are file-only types the caller widens before use */
enum { RFSI_IO_F64 = 0, RFSI_IO_F32 = 1, RFSI_IO_I32 = 2, RFSI_IO_I16 = 3, RFSI_IO_U16 = 4, RFSI_IO_U8 = 5,
       RFSI_IO_U32 = 6, RFSI_IO_I8 = 7 };

typedef struct rfsi_tiff rfsi_tiff_t;

typedef struct rfsi_tiff_info {
    int64_t ncol, nrow;
    int32_t nlayers;     /* SamplesPerPixel */
    int32_t dtype;       /* RFSI_IO_* */
    int32_t compression; /* TIFF tag 259: 1 none, 5 LZW, 8 / 32946 Deflate, 32773 PackBits */
    int32_t predictor;   /* TIFF tag 317 */
    int32_t tiled;
    int32_t big;         /* BigTIFF */
    int32_t has_geo;     /* pixel scale + tiepoint present */
    double x_ul, y_ul;   /* outer corner of the top-left cell (PixelIsPoint files are shifted by half a cell) */
    double dx, dy;       /* cell size, both > 0 (north-up) */
    int32_t has_nodata;
    double nodata;       /* GDAL_NODATA (tag 42113), what raster() reports as NAvalue */
    int32_t epsg;        /* ProjectedCSTypeGeoKey, else GeographicTypeGeoKey; 0 = none, 32767 = user-defined */
} rfsi_tiff_info_t;

const char* rfsi_io_last_error(void);

/* raster(<file>): parse the first image directory.  *out is NULL on failure. */
int rfsi_tiff_open(const char* path, rfsi_tiff_t** out);
void rfsi_tiff_close(rfsi_tiff_t* t);
int rfsi_tiff_get_info(const rfsi_tiff_t* t, rfsi_tiff_info_t* info);
/* all cells, decoded, in host byte order: out[(l*nrow + r)*ncol + c], row 0 = north.
 * out_bytes must be nlayers*nrow*ncol*sizeof(cell).  Strips/tiles are decoded on
 * `threads` host threads (0 = all cores). */
int rfsi_tiff_read(const rfsi_tiff_t* t, void* out, size_t out_bytes, int threads);
/* raw values of one tag, converted to host byte order (so a template's GeoKey tags
 * 34735 / 34736 / 34737 can be handed to rfsi_tiff_write).  Returns the element count
 * (0 = absent) and stores the TIFF field type; copies at most cap_bytes into out. */
int64_t rfsi_tiff_tag(const rfsi_tiff_t* t, int32_t tag, int32_t* field_type, void* out, size_t cap_bytes);

typedef struct rfsi_tiff_write_args {
    size_t struct_size; /* = sizeof(rfsi_tiff_write_args_t) */
    const void* data;   /* data[(l*nrow + r)*ncol + c] */
    int32_t dtype;      /* RFSI_IO_* (writeRaster's INT2S = RFSI_IO_I16) */
    int32_t nlayers;
    int64_t ncol, nrow;
    double x_ul, y_ul, dx, dy;
    int32_t has_nodata;
    double nodata;       /* NAflag */
    int32_t compression; /* 1 none, 5 LZW (what writeRaster's GTiff driver produced for the reference's files), 8 Deflate */
    int32_t epsg;        /* 0 = no CRS keys; 4000..4999 written as geographic, others as projected */
    /* optional verbatim GeoKey tags of a template file (override epsg): */
    const uint16_t* geokeys;   int32_t ngeokeys;     /* tag 34735, count of shorts */
    const double* geo_doubles; int32_t ngeo_doubles; /* tag 34736 */
    const char* geo_ascii;                           /* tag 34737, NUL-terminated */
} rfsi_tiff_write_args_t;

/* writeRaster(p, path, "GTiff", NAflag=, datatype=): little-endian, strips, planar bands;
 * switches to BigTIFF above 4 GB. */
int rfsi_tiff_write(const char* path, const rfsi_tiff_write_args_t* a);

/* ---- border polygons ------------------------------------------------------------------
 * The mapping scripts cut every product to the country border:
 *   border <- readOGR("borders/osm/union_of_selected_boundaries_AL2-AL2.shp"); border <- spTransform(border, CRS(utm33))
 *   p = crop(p, border); p = mask(p, border)            (temp_croatia/2_predictions.R:27-28,108-109,113-114,...)
 * rfsi_shp_read_polygons is the readOGR half for polygon shapefiles (ESRI shape types 5 / 15 / 25; Z and M values
 * are ignored): all rings of all records, concatenated.  *xy receives 2*npoints doubles (x0, y0, x1, y1, ...),
 * *ring_start receives nrings+1 offsets into the points; both are released with rfsi_io_free.
 * rfsi_polygon_mask is the mask() half: out[row*ncol + col] = 1 where the CENTRE of the cell lies inside the
 * polygons (even-odd rule over all rings, so holes and multi-part borders come out right), else 0 -- the
 * pix_mask argument of rfsi_predict_fused.  Vertices are used as given: project them first when the raster is
 * projected (the scripts' spTransform moves the vertices only, too). */
int rfsi_shp_read_polygons(const char* path, double** xy, int64_t* npoints, int64_t** ring_start, int64_t* nrings);
void rfsi_io_free(void* p);
int rfsi_polygon_mask(const double* xy, const int64_t* ring_start, int64_t nrings, double x_ul, double y_ul, double dx,
                      double dy, int64_t ncol, int64_t nrow, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif /* RFSI_IO_H */
