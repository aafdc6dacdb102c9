"""Border polygons: readOGR(<polygon shapefile>) and raster::mask(p, border)
(temp_croatia/2_predictions.R:27-28,108-109) -- librfsi_io.so's rfsi_shp_read_polygons /
rfsi_polygon_mask against shapefiles assembled byte by byte here, a brute-force
point-in-polygon test, and the reference's own Catalonian border + DEM."""
import struct
from pathlib import Path

import numpy as np
import pytest

from rfsi_b200 import geotiff as gt

REF = Path("/root/reference/prcp_catalonia")


def _shp(records, shape_type=5):
    """records: list of polygons, each a list of rings (arrays (n, 2)).  Returns the .shp bytes."""
    body = b""
    for k, rings in enumerate(records, start=1):
        pts = np.concatenate(rings)
        parts = np.cumsum([0] + [len(r) for r in rings[:-1]]).astype("<i4")
        content = struct.pack("<i4d2i", shape_type, pts[:, 0].min(), pts[:, 1].min(), pts[:, 0].max(), pts[:, 1].max(),
                              len(rings), len(pts)) + parts.tobytes() + pts.astype("<f8").tobytes()
        if shape_type == 15:       # PolygonZ: z range + z values, m range + m values follow the points
            content += struct.pack("<2d", 0, 0) + np.zeros(len(pts)).tobytes() + struct.pack("<2d", 0, 0) + np.zeros(len(pts)).tobytes()
        body += struct.pack(">2i", k, len(content) // 2) + content
    head = struct.pack(">7i", 9994, 0, 0, 0, 0, 0, (100 + len(body)) // 2) + struct.pack("<2i", 1000, shape_type) + b"\0" * 64
    return head + body


def _inside(xy_rings, px, py):
    """even-odd rule by brute force (half-open in y, like a scan line through the cell centres)"""
    inside = np.zeros(px.shape, dtype=bool)
    for ring in xy_rings:
        x1, y1 = ring[:, 0], ring[:, 1]
        x2, y2 = np.roll(x1, -1), np.roll(y1, -1)
        for a, b, c, d in zip(x1, y1, x2, y2):
            if b == d:
                continue
            if b > d:
                a, b, c, d = c, d, a, b
            hit = (py >= b) & (py < d)
            xc = a + (py - b) * (c - a) / (d - b)
            inside ^= hit & (xc > px)
    return inside


def test_read_polygons_parts_holes_and_z(tmp_path):
    outer = np.array([[0, 0], [0, 10], [10, 10], [10, 0], [0, 0]], dtype=float)        # clockwise: outer ring
    hole = np.array([[3, 3], [7, 3], [7, 7], [3, 7], [3, 3]], dtype=float)
    island = np.array([[20, 0], [20, 5], [25, 5], [25, 0], [20, 0]], dtype=float)
    for st in (5, 15):
        (tmp_path / "b.shp").write_bytes(_shp([[outer, hole], [island]], st))
        poly = gt.read_polygons(tmp_path / "b.shp")
        assert poly.ring_start.tolist() == [0, 5, 10, 15]
        np.testing.assert_array_equal(poly.xy, np.concatenate([outer, hole, island]))
    m = gt.polygon_mask(poly, -1.0, 11.0, 1.0, 1.0, 28, 12)                       # cell centres at k + 0.5
    assert m.dtype == np.uint8 and m.shape == (12, 28)
    exp = np.zeros((12, 28), dtype=np.uint8)
    exp[1:11, 1:11] = 1; exp[4:8, 4:8] = 0; exp[6:11, 21:26] = 1                  # row 0 is y in (10, 11)
    np.testing.assert_array_equal(m, exp)
    moved = poly.transformed(lambda x, y: (x + 100.0, y))
    assert gt.polygon_mask(moved, -1.0, 11.0, 1.0, 1.0, 28, 12).sum() == 0
    assert gt.polygon_mask(moved, 99.0, 11.0, 1.0, 1.0, 28, 12).sum() == exp.sum()


def test_mask_matches_brute_force_on_random_polygons():
    rng = np.random.default_rng(0)
    for trial in range(30):
        rings = []
        for _ in range(int(rng.integers(1, 4))):
            n = int(rng.integers(3, 40))
            ang = np.sort(rng.uniform(0, 2 * np.pi, n))
            rad = rng.uniform(2, 12, n)
            c = rng.uniform(8, 22, 2)
            ring = np.column_stack([c[0] + rad * np.cos(ang), c[1] + rad * np.sin(ang)])
            if trial % 3 == 0:
                ring = np.round(ring * 2) / 2          # vertices ON cell-centre lines and cell edges
            rings.append(ring)
        poly = gt.Polygons(np.concatenate(rings), np.cumsum([0] + [len(r) for r in rings]))
        ncol, nrow, dx, dy, x0, y0 = 37, 41, 1.0, 0.75, -2.5, 33.0
        got = gt.polygon_mask(poly, x0, y0, dx, dy, ncol, nrow).astype(bool)
        cx = x0 + (np.arange(ncol) + 0.5) * dx
        cy = y0 - (np.arange(nrow) + 0.5) * dy
        px, py = np.meshgrid(cx, cy)
        exp = _inside(rings, px, py)
        # a cell centre exactly on an edge may fall either way; everything else must agree
        diff = got != exp
        if diff.any():
            on_edge = np.zeros_like(diff)
            for ring in rings:
                a, b = ring, np.roll(ring, -1, axis=0)
                for (x1, y1), (x2, y2) in zip(a, b):
                    cr = (px - x1) * (y2 - y1) - (py - y1) * (x2 - x1)
                    on_edge |= np.abs(cr) < 1e-9
            assert not (diff & ~on_edge).any(), trial
        assert got.sum() > 0


def test_errors_and_damaged_files(tmp_path):
    with pytest.raises(gt.GeoTiffError) as e:
        gt.read_polygons(tmp_path / "none.shp")
    assert e.value.code == -1
    (tmp_path / "junk.shp").write_bytes(b"x" * 200)
    with pytest.raises(gt.GeoTiffError) as e:
        gt.read_polygons(tmp_path / "junk.shp")
    assert e.value.code == -2
    sq = np.array([[0, 0], [0, 1], [1, 1], [1, 0], [0, 0]], dtype=float)
    (tmp_path / "pts.shp").write_bytes(_shp([[sq]], 5).replace(struct.pack("<2i", 1000, 5), struct.pack("<2i", 1000, 1), 1))
    with pytest.raises(gt.GeoTiffError) as e:
        gt.read_polygons(tmp_path / "pts.shp")                     # a point shapefile defines no mask
    assert e.value.code == -3
    good = _shp([[sq * 5, sq + 1.5], [sq + 20]])
    rng = np.random.default_rng(5)
    ok = bad = 0
    for _ in range(400):                                           # damaged files are refused, never a crash
        v = bytearray(good)
        for _ in range(int(rng.integers(1, 5))):
            v[int(rng.integers(24, len(v)))] = int(rng.integers(0, 256))
        if rng.random() < 0.2:
            v = v[:int(rng.integers(1, len(v)))]
        (tmp_path / "m.shp").write_bytes(bytes(v))
        try:
            p = gt.read_polygons(tmp_path / "m.shp")
            if np.isfinite(p.xy).all():
                gt.polygon_mask(p, -5.0, 30.0, 1.0, 1.0, 40, 40)
            ok += 1
        except gt.GeoTiffError:
            bad += 1
    assert ok + bad == 400 and bad > 20
    with pytest.raises(gt.GeoTiffError):
        gt.polygon_mask(gt.Polygons(np.array([[0.0, 0], [np.nan, 1], [1, 1]]), np.array([0, 3])), 0, 1, 1, 1, 2, 2)
    with pytest.raises(gt.GeoTiffError):
        gt.polygon_mask(gt.Polygons(sq, np.array([0, 5])), 0, 1, 0.0, 1, 2, 2)
    # absurd but finite coordinates (what a damaged file decodes to) rasterise to something, without overflow
    wild = gt.Polygons(np.array([[-1e300, -1e300], [1e300, -1e300], [1e300, 1e300], [-1e300, 1e300]]), np.array([0, 4]))
    assert gt.polygon_mask(wild, 0.0, 8.0, 1.0, 1.0, 8, 8).sum() == 64
    far = gt.Polygons(np.array([[1e300, 0.0], [1e301, 0.0], [1e301, 8.0], [1e300, 8.0]]), np.array([0, 4]))
    assert gt.polygon_mask(far, 0.0, 8.0, 1.0, 1.0, 8, 8).sum() == 0
    with pytest.raises(gt.GeoTiffError):
        gt.polygon_mask(gt.Polygons(np.array([[0.0, 0], [np.inf, 1], [1, 1]]), np.array([0, 3])), 0, 1, 1, 1, 2, 2)


@pytest.mark.skipif(not REF.exists(), reason="reference data not on this box")
def test_catalonian_border_reproduces_the_dem_mask():
    """The reference's DEM (dem_twi/dem_cat.tif, 50 046 valid cells) is what mask(dem, border) leaves: every valid
    cell lies inside catalonia_border/*.shp, and the border adds only a few coastal cells the DEM has no data for."""
    poly = gt.read_polygons(REF / "catalonia_border" / "union_of_selected_boundaries_AL4-AL4.shp")
    assert poly.xy.shape == (133062, 2) and len(poly.ring_start) - 1 == 20
    r, info = gt.read_geotiff(REF / "dem_twi" / "dem_cat.tif")
    m = gt.polygon_mask(poly, info.x_ul, info.y_ul, info.dx, info.dy, info.ncol, info.nrow).astype(bool)
    valid = r.data != -32767
    assert (valid & ~m).sum() == 0 and (m & ~valid).sum() == 39 and m.sum() == 50085
    from rfsi_b200 import mapping
    assert mapping.grid_from_dem(REF / "dem_twi" / "dem_cat.tif", border=poly).npix == 50046
