"""The reference's map-making loop as one call: GeoTIFF covariates in, Int16 GeoTIFF products out.

prcp_catalonia/5_prcp_prediction.R:228-313 does, for every day i:

    obs_st  <- temp[, i, "prcp"];  obs_st <- obs_st[!is.na(obs_st$prcp),]   # that day's stations
    if (length(obs_st) < 2) return(NULL)                                     # skip the day
    df$tmin <- extract(raster("../min/<day>.tif"), gg) / 10                  # gg: DEM cells, lon/lat
    df$tmax <- extract(raster("../max/<day>.tif"), gg) / 10
    df$tmin <- ifelse(df$tmin > df$tmax, df$tmax-1, df$tmin)
    df$imerg <- extract(raster("../imerg/...<day>....tif"), gg)
    nearest_obs <- near.obs(locations = gg2, observations = obs_st, ...)     # gg2: gg in UTM31
    pre <- round(predict(rfsi_model, df)$predictions * 100)
    iqr <- round((q[,3] - q[,1]) / 2 * 100)                                  # quantiles .25/.5/.75
    writeRaster(..., "<day>",     "GTiff", NAflag=-32767, datatype='INT2S')
    writeRaster(..., "<day>_iqr", "GTiff", NAflag=-32767, datatype='INT2S')

(temp_croatia/2_predictions.R:156-183 is the same shape with in-memory covariates.)  Here the
files are decoded by librfsi_io.so, every day goes through ONE rfsi_predict_fused call (rasters
sampled, repaired, near.obs'd and predicted on the device; Int16 products come back), and the
products are written by librfsi_io.so.  Nothing in this module computes a prediction on the CPU.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Optional, Sequence, Union

import numpy as np

from . import api, geotiff, proj
from .forest import RangerForest


@dataclass
class PredictionGrid:
    """`gg` / `gg2`: the non-NA cells of a DEM, their centres in the raster's own CRS (what
    extract() is given) and in the CRS near.obs works in."""
    info: geotiff.GeoTiffInfo
    cell: np.ndarray       # (npix,) row-major cell index into the DEM grid
    xy_raster: np.ndarray  # (npix, 2) cell centres in the DEM's CRS
    xy: np.ndarray         # (npix, 2) coordinates for near.obs

    @property
    def npix(self) -> int:
        return int(self.cell.size)

    def to_raster_i16(self, values: np.ndarray) -> np.ndarray:
        """as(gg["pred"], "SpatialPixelsDataFrame") -> raster: values back on the full grid, NA elsewhere."""
        full = np.full(self.info.nrow * self.info.ncol, -32767, dtype=np.int16)
        full[self.cell] = values
        return full.reshape(self.info.nrow, self.info.ncol)


def grid_from_dem(dem_path, utm_zone: Optional[int] = None, border: Optional[geotiff.Polygons] = None) -> PredictionGrid:
    """gg <- as(readGDAL(dem), "SpatialPixelsDataFrame") (drops NA cells);
    gg2 <- spTransform(gg, utm<zone>) when utm_zone is given (5_prcp_prediction.R:38-46).
    border (in the DEM's CRS): additionally `crop(p, border); mask(p, border)` -- only cells whose
    centre lies inside the polygons are predicted (temp_croatia/2_predictions.R:108-109)."""
    r, info = geotiff.read_geotiff(dem_path)
    d = r.data if r.data.ndim == 2 else r.data[0]
    valid = ~np.isnan(d.astype(np.float64))
    if info.nodata is not None:
        valid &= d != d.dtype.type(info.nodata)
    if border is not None:
        valid &= geotiff.polygon_mask(border, info.x_ul, info.y_ul, info.dx, info.dy, info.ncol, info.nrow).astype(bool)
    cell = np.flatnonzero(valid.ravel())
    row, col = np.divmod(cell, info.ncol)
    x = info.x_ul + (col + 0.5) * info.dx
    y = info.y_ul - (row + 0.5) * info.dy
    xy_r = np.column_stack([x, y])
    xy = xy_r if utm_zone is None else np.column_stack(proj.lonlat_to_utm(x, y, utm_zone))
    return PredictionGrid(info, cell, xy_r, xy)


@dataclass
class CovariateFiles:
    """One covariate of the loop: a GeoTIFF per day (or one static file), plus the unit change
    the script applies right after extract() (`/ 10`: divide=True, scale=10)."""
    paths: Union[str, Path, Sequence[Union[str, Path]]]
    scale: float = 1.0
    offset: float = 0.0
    divide: bool = False

    def load(self, ndays: int) -> api.Raster:
        if isinstance(self.paths, (str, Path)):
            r, _ = geotiff.read_geotiff(self.paths, self.scale, self.offset, self.divide)
            return r
        if len(self.paths) != ndays:
            raise ValueError(f"{len(self.paths)} files for {ndays} days")
        layers, first = [], None
        for p in self.paths:
            r, _ = geotiff.read_geotiff(p, self.scale, self.offset, self.divide)
            if r.data.ndim != 2:
                raise ValueError(f"{p}: one band per daily file expected")
            if first is None:
                first = r
            elif (r.data.shape, r.data.dtype, r.x_ul, r.y_ul, r.dx, r.dy) != \
                    (first.data.shape, first.data.dtype, first.x_ul, first.y_ul, first.dx, first.dy) or \
                    not (r.nodata == first.nodata or (np.isnan(r.nodata) and np.isnan(first.nodata))):
                raise ValueError(f"{p}: grid, cell type or NAvalue differs from {self.paths[0]}")
            layers.append(r.data)
        return api.Raster(np.stack(layers), first.x_ul, first.y_ul, first.dx, first.dy, nodata=first.nodata,
                          scale=self.scale, offset=self.offset, divide=self.divide)


@dataclass
class MapResult:
    days: list                      # names of the days that were mapped (days with < min_stations are skipped)
    pred_i16: np.ndarray            # (ndays_done, npix) round(pred*100)
    iqr_i16: Optional[np.ndarray]   # (ndays_done, npix) round((q_hi - q_lo)/2*100)
    files: list = field(default_factory=list)


def predict_maps(model: RangerForest, grid: PredictionGrid, obs_xy, obs_z, day_names: Sequence[str], n_obs: int,
                 covariates: Dict[str, Union[CovariateFiles, np.ndarray]], cov_fix: Sequence = (),
                 quantiles: Optional[Sequence[float]] = (0.25, 0.5, 0.75), out_dir=None, min_stations: int = 2,
                 compression: str = "lzw", device: int = 0, devices: Optional[Sequence[int]] = None) -> MapResult:
    """The loop of 5_prcp_prediction.R:244-313 for all days at once.

    obs_xy (nobs, 2) in the near.obs CRS, obs_z (ndays, nobs) with NaN = no observation.
    covariates: name -> CovariateFiles (sampled on the device at the grid's raster coordinates)
    or an array (npix,) / (ndays, npix).  With out_dir, writes <day>.tif and <day>_iqr.tif.
    devices: several GPUs of the box, one band of pixels each (the scripts' foreach %dopar%).
    """
    obs_z = np.atleast_2d(np.asarray(obs_z, dtype=np.float64))
    ndays = obs_z.shape[0]
    if len(day_names) != ndays:
        raise ValueError("one name per day")
    use = np.flatnonzero((~np.isnan(obs_z)).sum(axis=1) >= min_stations)       # `if (length(obs_st) < 2) return(NULL)`
    cov = {}
    for name, c in covariates.items():
        if isinstance(c, CovariateFiles):
            r = c.load(ndays)
            if r.data.ndim == 3 and len(use) != ndays:
                r.data = np.ascontiguousarray(r.data[use])
            cov[name] = r
        else:
            v = np.asarray(c, dtype=np.float64)
            cov[name] = v[use] if v.ndim == 2 else v
    res = api.predict_fused(model, obs_xy=obs_xy, obs_z=obs_z[use], n_obs=n_obs, locations=grid.xy,
                            cov_locations=None if grid.xy is grid.xy_raster else grid.xy_raster, covariates=cov,
                            cov_fix=cov_fix, quantiles=quantiles, centi_int16=True, device=device, devices=devices)
    out = MapResult([day_names[i] for i in use], res["mean_i16"], res.get("iqr_i16"))
    if out_dir is not None:
        out_dir = Path(out_dir)
        out_dir.mkdir(parents=True, exist_ok=True)
        for k, name in enumerate(out.days):
            f = out_dir / f"{name}.tif"
            geotiff.write_product(f, grid.to_raster_i16(out.pred_i16[k]), grid.info, compression)
            out.files.append(f)
            if out.iqr_i16 is not None:
                f = out_dir / f"{name}_iqr.tif"
                geotiff.write_product(f, grid.to_raster_i16(out.iqr_i16[k]), grid.info, compression)
                out.files.append(f)
    return out
