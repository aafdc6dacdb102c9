"""Packs the reference's Catalonian precipitation inputs into one small fixture, so that the
map-making loop of prcp_catalonia/5_prcp_prediction.R:228-313 can be re-run on the GPU box
(where /root/reference does not exist) on the REAL station table and the REAL rasters:

  temp_data/stfdf_temp.rda   87 stations (lon/lat) x 1095 days: prcp, imerg, tmax, tmin
  dem_twi/dem_cat.tif        380 x 280 Float32 DEM, nodata -32767 (50 046 valid cells) = the grid
  min|max/2016010{1..4}.tif  daily tmin / tmax, Int16 tenths of a degree, on the DEM grid
  imerg/*2016010{1..4}*.tif  IMERG daily precipitation, UInt16, global 0.1 degree (cropped here
                             to the cells around Catalonia; the crop keeps whole cells)

The station table goes through rfsi_b200/rda.py, the rasters through rfsi_b200/geotiff.py (each
checked against PIL's libtiff decode below).  Run in the build container:
    python tests/golden/make_catalonia.py
"""
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent
REF = Path("/root/reference/prcp_catalonia")
DAYS = ["20160101", "20160102", "20160103", "20160104"]     # 5_prcp_prediction.R:31


def main():
    import sys
    sys.path.insert(0, str(OUT.parents[1]))
    from PIL import Image
    from rfsi_b200 import geotiff, rda

    st = rda.read_rda(REF / "temp_data" / "stfdf_temp.rda")["stfdf"]
    lonlat = np.asarray(st.attrs["sp"].attrs["coords"], dtype=np.float64)
    S = lonlat.shape[0]
    d = st.attrs["data"]
    tab = {k: np.asarray(d[k], dtype=np.float64).reshape(-1, S) for k in ("prcp", "imerg", "tmax", "tmin")}  # space runs fastest
    assert tab["prcp"].shape == (1095, 87)
    t0 = np.asarray(st.attrs["time"].attrs["index"], dtype=np.float64)
    assert t0[0] == 1451606400.0 and len(t0) == 1095        # 2016-01-01 UTC: day i of the table is DAYS[i]
    packed = {}
    for k, v in tab.items():                                # tenths are exact: store integers, rebuild with / 10
        t = np.round(v * 10)
        assert np.array_equal(np.where(np.isnan(v), 0, t / 10), np.where(np.isnan(v), 0, v)), k
        packed[k + "_x10"] = np.where(np.isnan(v), -32768, t).astype(np.int16 if np.nanmax(np.abs(t)) < 32000 else np.int32)

    def read(p):
        r, info = geotiff.read_geotiff(p)
        assert np.array_equal(r.data, np.array(Image.open(p)).astype(r.data.dtype)), p
        return r, info

    dem, di = read(REF / "dem_twi" / "dem_cat.tif")
    assert (di.ncol, di.nrow, di.nodata) == (380, 280, -32767.0) and int((dem.data != -32767).sum()) == 50046
    tmin = np.stack([read(REF / "min" / f"{x}.tif")[0].data for x in DAYS])
    tmax = np.stack([read(REF / "max" / f"{x}.tif")[0].data for x in DAYS])
    _, ti = read(REF / "min" / f"{DAYS[0]}.tif")
    im, ii = [], None
    for x in DAYS:
        r, ii = read(next((REF / "imerg").glob(f"*{x}*.tif")))
        im.append(r.data)
    c0, c1, r0, r1 = 1795, 1841, 465, 501                   # whole cells around lon 0.16..3.33, lat 40.5..42.9
    imerg = np.stack(im)[:, r0:r1, c0:c1]
    geo = lambda i, x_ul, y_ul: np.array([x_ul, y_ul, i.dx, i.dy])
    np.savez_compressed(
        OUT / "catalonia_maps.npz", lonlat=lonlat, days=np.array(DAYS), **packed,
        dem=dem.data, dem_geo=geo(di, di.x_ul, di.y_ul), dem_geokeys=di.geokeys.directory,
        dem_geodoubles=di.geokeys.doubles, dem_geoascii=np.frombuffer(di.geokeys.ascii, dtype=np.uint8),
        tmin=tmin, tmax=tmax, t_geo=geo(ti, ti.x_ul, ti.y_ul),
        imerg=imerg, imerg_geo=geo(ii, ii.x_ul + c0 * ii.dx, ii.y_ul - r0 * ii.dy))
    print("catalonia_maps.npz", (OUT / "catalonia_maps.npz").stat().st_size)


if __name__ == "__main__":
    main()
