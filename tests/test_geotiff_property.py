"""Property tests (hypothesis) for the GeoTIFF codec: whatever goes in comes back, for every cell
type, layer count and compression, on data chosen to stress the LZW tables (long runs, short
periods, incompressible noise) -- and libtiff (PIL) agrees on the single-band files."""
import io

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from rfsi_b200 import geotiff as gt

DTYPES = ["float64", "float32", "int32", "int16", "uint16", "uint8", "uint32", "int8"]


def _cells(kind, dtype, shape, seed):
    rng = np.random.default_rng(seed)
    n = int(np.prod(shape))
    if kind == "noise":
        raw = rng.integers(0, 256, size=n * np.dtype(dtype).itemsize, dtype=np.uint8)
        a = raw.view(dtype)[:n]
        if np.dtype(dtype).kind == "f":
            a = np.where(np.isfinite(a), a, 0).astype(dtype)
    elif kind == "runs":
        a = np.repeat(rng.integers(0, 100, size=n // 7 + 1), 7)[:n].astype(dtype)
    elif kind == "period":
        a = np.tile(rng.integers(0, 100, size=int(rng.integers(1, 9))), n)[:n].astype(dtype)
    else:
        a = np.full(n, int(rng.integers(0, 100))).astype(dtype)
    return np.ascontiguousarray(a).reshape(shape)


@settings(max_examples=120, deadline=None)
@given(dtype=st.sampled_from(DTYPES), layers=st.integers(1, 3), nrow=st.integers(1, 70), ncol=st.integers(1, 90),
       kind=st.sampled_from(["noise", "runs", "period", "flat"]), comp=st.sampled_from(["none", "lzw", "deflate"]),
       seed=st.integers(0, 1000), nodata=st.sampled_from([None, -32767, 0, 255]), big=st.booleans())
def test_round_trip(tmp_path_factory, monkeypatch_session, dtype, layers, nrow, ncol, kind, comp, seed, nodata, big):
    monkeypatch_session.setenv("RFSI_IO_FORCE_BIGTIFF", "1" if big else "0")
    a = _cells(kind, dtype, (layers, nrow, ncol), seed)
    p = tmp_path_factory.mktemp("rt") / "a.tif"
    gt.write_geotiff(p, a if layers > 1 else a[0], 12.5, -3.25, 0.5, 2.0, nodata=nodata, epsg=32633, compression=comp)
    with gt.GeoTiff(p) as t:
        got, i = t.read(), t.info
    np.testing.assert_array_equal(got, a if layers > 1 else a[0])
    assert (i.x_ul, i.y_ul, i.dx, i.dy, i.epsg, i.big) == (12.5, -3.25, 0.5, 2.0, 32633, big)
    assert i.nodata == (None if nodata is None else float(nodata))
    if layers == 1 and not big and dtype in ("uint8", "uint16", "int32", "float32"):
        Image = pytest.importorskip("PIL.Image")
        np.testing.assert_array_equal(np.array(Image.open(io.BytesIO(p.read_bytes()))).astype(dtype), a[0])


@pytest.fixture(scope="session")
def monkeypatch_session():
    mp = pytest.MonkeyPatch()
    yield mp
    mp.undo()
