"""GeoTIFF in and out: raster(<file>) and writeRaster(..., "GTiff", NAflag=-32767, datatype='INT2S').

ctypes binding of librfsi_io.so (include/rfsi_io.h, csrc/geotiff.cpp).  The reference reads its
covariates from GeoTIFFs (prcp_catalonia/5_prcp_prediction.R:253-276; temp_croatia/dem.tif) and
writes every map as an Int16 GeoTIFF (5_prcp_prediction.R:303,309; temp_croatia/2_predictions.R:173,183);
GDAL is not available here, so the reader/writer is part of the build.  `read_geotiff` returns the
`Raster` that `raster_extract` / `predict_fused` sample on the device in the file's own cell type.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from pathlib import Path
from typing import Optional

import numpy as np

from .api import Raster

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "librfsi_io.so"

IO_DTYPES = {0: np.float64, 1: np.float32, 2: np.int32, 3: np.int16, 4: np.uint16, 5: np.uint8, 6: np.uint32, 7: np.int8}
_IO_CODES = {np.dtype(v).name: k for k, v in IO_DTYPES.items()}
COMPRESSION = {"none": 1, "lzw": 5, "deflate": 8}


class GeoTiffError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"rfsi_io error {code}: {message}")
        self.code = code


class _Info(C.Structure):
    _fields_ = [
        ("ncol", C.c_int64), ("nrow", C.c_int64), ("nlayers", C.c_int32), ("dtype", C.c_int32),
        ("compression", C.c_int32), ("predictor", C.c_int32), ("tiled", C.c_int32), ("big", C.c_int32),
        ("has_geo", C.c_int32), ("x_ul", C.c_double), ("y_ul", C.c_double), ("dx", C.c_double), ("dy", C.c_double),
        ("has_nodata", C.c_int32), ("nodata", C.c_double), ("epsg", C.c_int32),
    ]


class _WriteArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t), ("data", C.c_void_p), ("dtype", C.c_int32), ("nlayers", C.c_int32),
        ("ncol", C.c_int64), ("nrow", C.c_int64),
        ("x_ul", C.c_double), ("y_ul", C.c_double), ("dx", C.c_double), ("dy", C.c_double),
        ("has_nodata", C.c_int32), ("nodata", C.c_double), ("compression", C.c_int32), ("epsg", C.c_int32),
        ("geokeys", C.POINTER(C.c_uint16)), ("ngeokeys", C.c_int32),
        ("geo_doubles", C.POINTER(C.c_double)), ("ngeo_doubles", C.c_int32),
        ("geo_ascii", C.c_char_p),
    ]


_P = C.c_void_p
_SIGNATURES = [
    ("rfsi_io_last_error", C.c_char_p, []),
    ("rfsi_tiff_open", C.c_int, [C.c_char_p, C.POINTER(_P)]),
    ("rfsi_tiff_close", None, [_P]),
    ("rfsi_tiff_get_info", C.c_int, [_P, C.POINTER(_Info)]),
    ("rfsi_tiff_read", C.c_int, [_P, _P, C.c_size_t, C.c_int]),
    ("rfsi_tiff_tag", C.c_int64, [_P, C.c_int32, C.POINTER(C.c_int32), _P, C.c_size_t]),
    ("rfsi_tiff_write", C.c_int, [C.c_char_p, C.POINTER(_WriteArgs)]),
    ("rfsi_shp_read_polygons", C.c_int, [C.c_char_p, C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_int64),
                                         C.POINTER(C.POINTER(C.c_int64)), C.POINTER(C.c_int64)]),
    ("rfsi_io_free", None, [_P]),
    ("rfsi_polygon_mask", C.c_int, [_P, _P, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_int64, _P]),
]
EXPORTED_SYMBOLS = [s[0] for s in _SIGNATURES]
_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        path = Path(os.environ.get("RFSI_IO_LIB", LIB_PATH))
        if not path.exists():
            raise GeoTiffError(-1, f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(str(path))
        for name, restype, argtypes in _SIGNATURES:
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = restype, argtypes
        _lib = lib
    return _lib


def _check(rc: int) -> int:
    if rc < 0:
        raise GeoTiffError(rc, load().rfsi_io_last_error().decode("utf-8", "replace"))
    return rc


@dataclass
class GeoKeys:
    """A file's GeoKey tags (34735 / 34736 / 34737), kept verbatim so that products inherit the
    CRS of the raster they were predicted on (what writeRaster does with crs(p))."""
    directory: np.ndarray            # uint16
    doubles: Optional[np.ndarray]    # float64
    ascii: Optional[bytes]


@dataclass
class GeoTiffInfo:
    ncol: int
    nrow: int
    nlayers: int
    dtype: np.dtype
    compression: int
    predictor: int
    tiled: bool
    big: bool
    has_geo: bool
    x_ul: float
    y_ul: float
    dx: float
    dy: float
    nodata: Optional[float]
    epsg: int
    geokeys: Optional[GeoKeys] = None


class GeoTiff:
    """An open GeoTIFF (first image directory)."""

    def __init__(self, path):
        self._h = _P()
        self.path = str(path)
        _check(load().rfsi_tiff_open(os.fsencode(self.path), C.byref(self._h)))
        i = _Info()
        _check(load().rfsi_tiff_get_info(self._h, C.byref(i)))
        self.info = GeoTiffInfo(i.ncol, i.nrow, i.nlayers, np.dtype(IO_DTYPES[i.dtype]), i.compression, i.predictor,
                                bool(i.tiled), bool(i.big), bool(i.has_geo), i.x_ul, i.y_ul, i.dx, i.dy,
                                i.nodata if i.has_nodata else None, i.epsg)
        d = self.tag(34735)
        if d is not None:
            a = self.tag(34737)
            self.info.geokeys = GeoKeys(d.astype(np.uint16), self.tag(34736), None if a is None else bytes(a).rstrip(b"\0"))

    def tag(self, tag_id: int):
        """Raw values of one TIFF tag (numpy array; ASCII / BYTE tags as uint8), or None."""
        lib = load()
        ft = C.c_int32()
        n = _check(lib.rfsi_tiff_tag(self._h, tag_id, C.byref(ft), None, 0))
        if n == 0:
            return None
        np_t = {1: np.uint8, 2: np.uint8, 3: np.uint16, 4: np.uint32, 6: np.int8, 7: np.uint8, 8: np.int16, 9: np.int32,
                11: np.float32, 12: np.float64, 13: np.uint32, 16: np.uint64, 17: np.int64, 18: np.uint64}.get(ft.value)
        if np_t is None:    # rationals: pairs of 32-bit words
            np_t, n = (np.uint32 if ft.value == 5 else np.int32), 2 * n
        out = np.empty(n, dtype=np_t)
        lib.rfsi_tiff_tag(self._h, tag_id, None, out.ctypes.data_as(_P), out.nbytes)
        return out

    def read(self, threads: int = 0) -> np.ndarray:
        """All cells as (nrow, ncol), or (nlayers, nrow, ncol), in the file's cell type."""
        i = self.info
        out = np.empty((i.nlayers, i.nrow, i.ncol), dtype=i.dtype)
        _check(load().rfsi_tiff_read(self._h, out.ctypes.data_as(_P), out.nbytes, int(threads)))
        return out[0] if i.nlayers == 1 else out

    def close(self):
        if self._h:
            load().rfsi_tiff_close(self._h)
            self._h = _P()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_geotiff(path, scale: float = 1.0, offset: float = 0.0, divide: bool = False, threads: int = 0):
    """raster(path): -> (Raster, GeoTiffInfo).  The Raster keeps the file's cell type (the
    device kernels read Int16 / UInt16 / Int32 / Float32 in place) and its NAvalue; uint32 and
    int8 files are widened to float64 / int16.  scale/offset/divide describe the unit change the
    scripts apply right after extract() (`/ 10` for tmin/tmax, 5_prcp_prediction.R:254-256)."""
    with GeoTiff(path) as t:
        data, info = t.read(threads), t.info
    if data.dtype == np.uint32:
        data = data.astype(np.float64)
    elif data.dtype == np.int8:
        data = data.astype(np.int16)
    nodata = float("nan") if info.nodata is None else float(info.nodata)
    r = Raster(data, info.x_ul, info.y_ul, info.dx, info.dy, nodata=nodata, scale=scale, offset=offset, divide=divide)
    return r, info


def write_geotiff(path, data, x_ul: float, y_ul: float, dx: float, dy: float, nodata: Optional[float] = None,
                  epsg: int = 0, geokeys: Optional[GeoKeys] = None, compression: str = "lzw") -> None:
    """writeRaster(p, path, "GTiff", NAflag=nodata, datatype=<data.dtype>): data is (nrow, ncol)
    or (nlayers, nrow, ncol), row 0 = north.  Int16 data is the reference's INT2S."""
    a = np.ascontiguousarray(data)
    if a.ndim == 2:
        a = a[None]
    if a.ndim != 3:
        raise ValueError("write_geotiff: data must be (nrow, ncol) or (nlayers, nrow, ncol)")
    if a.dtype.name not in _IO_CODES:
        raise ValueError(f"write_geotiff: cell type {a.dtype} unsupported (use one of {sorted(_IO_CODES)})")
    w = _WriteArgs()
    w.struct_size = C.sizeof(_WriteArgs)
    w.data = a.ctypes.data
    w.dtype = _IO_CODES[a.dtype.name]
    w.nlayers, w.nrow, w.ncol = a.shape
    w.x_ul, w.y_ul, w.dx, w.dy = float(x_ul), float(y_ul), float(dx), float(dy)
    w.has_nodata = 0 if nodata is None else 1
    w.nodata = 0.0 if nodata is None else float(nodata)
    w.compression = COMPRESSION[compression]
    w.epsg = int(epsg)
    keep = []
    if geokeys is not None:
        d = np.ascontiguousarray(geokeys.directory, dtype=np.uint16)
        keep.append(d)
        w.geokeys = d.ctypes.data_as(C.POINTER(C.c_uint16))
        w.ngeokeys = d.size
        if geokeys.doubles is not None:
            g = np.ascontiguousarray(geokeys.doubles, dtype=np.float64)
            keep.append(g)
            w.geo_doubles = g.ctypes.data_as(C.POINTER(C.c_double))
            w.ngeo_doubles = g.size
        if geokeys.ascii:
            w.geo_ascii = bytes(geokeys.ascii)
    _check(load().rfsi_tiff_write(os.fsencode(str(path)), C.byref(w)))
    del keep


def write_product(path, values_i16: np.ndarray, like: GeoTiffInfo, compression: str = "lzw") -> None:
    """One product map as the scripts write it: Int16 centi-units, NAflag -32767, on the grid and
    CRS of `like` (the DEM the pixels came from)."""
    v = np.asarray(values_i16)
    if v.dtype != np.int16:
        raise ValueError("write_product: Int16 centi-units expected (predict_fused(..., centi_int16=True))")
    write_geotiff(path, v.reshape(like.nrow, like.ncol), like.x_ul, like.y_ul, like.dx, like.dy, nodata=-32767,
                  epsg=like.epsg, geokeys=like.geokeys, compression=compression)


@dataclass
class Polygons:
    """readOGR(<polygon shapefile>): all rings of all records.  xy (npoints, 2); ring r is
    xy[ring_start[r]:ring_start[r + 1]] (outer rings and holes alike: masks use the even-odd rule)."""
    xy: np.ndarray
    ring_start: np.ndarray

    def transformed(self, fn) -> "Polygons":
        """spTransform(border, crs): fn(x, y) -> (x', y') on the vertices (e.g. proj.lonlat_to_utm)."""
        x, y = fn(self.xy[:, 0], self.xy[:, 1])
        return Polygons(np.column_stack([x, y]), self.ring_start)


def read_polygons(path) -> Polygons:
    """readOGR("borders/...shp") for polygon shapefiles (temp_croatia/2_predictions.R:27)."""
    lib = load()
    xy = C.POINTER(C.c_double)()
    rs = C.POINTER(C.c_int64)()
    n, nr = C.c_int64(), C.c_int64()
    _check(lib.rfsi_shp_read_polygons(os.fsencode(str(path)), C.byref(xy), C.byref(n), C.byref(rs), C.byref(nr)))
    try:
        pts = np.ctypeslib.as_array(xy, shape=(max(n.value, 1) * 2,))[:n.value * 2].reshape(-1, 2).copy()
        start = np.ctypeslib.as_array(rs, shape=(nr.value + 1,)).copy()
    finally:
        lib.rfsi_io_free(xy)
        lib.rfsi_io_free(rs)
    return Polygons(pts, start)


def polygon_mask(poly: Polygons, x_ul: float, y_ul: float, dx: float, dy: float, ncol: int, nrow: int) -> np.ndarray:
    """mask(r, border) as a (nrow, ncol) uint8 array: 1 where the centre of the cell lies inside the
    polygons (temp_croatia/2_predictions.R:108-109); feed it to predict_fused(pix_mask=...)."""
    xy = np.ascontiguousarray(poly.xy, dtype=np.float64)
    rs = np.ascontiguousarray(poly.ring_start, dtype=np.int64)
    out = np.empty((int(nrow), int(ncol)), dtype=np.uint8)
    _check(load().rfsi_polygon_mask(xy.ctypes.data_as(_P), rs.ctypes.data_as(_P), len(rs) - 1, float(x_ul), float(y_ul),
                                    float(dx), float(dy), int(ncol), int(nrow), out.ctypes.data_as(_P)))
    return out
