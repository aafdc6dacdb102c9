"""librfsi_io.so: the GeoTIFF reader (raster(<file>), prcp_catalonia/5_prcp_prediction.R:253-276)
and writer (writeRaster(..., "GTiff", NAflag=-32767, datatype='INT2S'), 5_prcp_prediction.R:303).

Checked against (a) libtiff through PIL, in both directions, (b) files assembled byte by byte
here for the layouts PIL does not write (tiles, big-endian, BigTIFF, planar bands, predictors),
(c) the reference's own 15 GeoTIFFs when /root/reference is present, through digests minted by
tests/golden/make_geotiff_golden.py (PIL's decode of each file)."""
import hashlib
import json
import re
import struct
import zlib
from pathlib import Path

import numpy as np
import pytest

from rfsi_b200 import geotiff as gt

ROOT = Path(__file__).resolve().parents[1]
GOLD = Path(__file__).parent / "golden"
DTYPES = ["float64", "float32", "int32", "int16", "uint16", "uint8", "uint32", "int8"]


def test_io_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "rfsi_io.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rfsi_[a-z0-9_]+)\s*\(", header))
    lib = gt.load()
    assert declared == set(gt.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)


def test_struct_layouts_match_header(tmp_path):
    import ctypes as C
    import subprocess
    src = ('#include "rfsi_io.h"\n#include <stdio.h>\nint main(){printf("%zu %zu\\n", sizeof(rfsi_tiff_info_t), '
           'sizeof(rfsi_tiff_write_args_t));return 0;}\n')
    (tmp_path / "s.c").write_text(src)
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(tmp_path / "s.c"), "-o", str(tmp_path / "s")], check=True)
    a, b = subprocess.run([str(tmp_path / "s")], check=True, capture_output=True, text=True).stdout.split()
    assert int(a) == C.sizeof(gt._Info) and int(b) == C.sizeof(gt._WriteArgs)


def _data(dtype, shape, seed=0):
    rng = np.random.default_rng(seed)
    if np.dtype(dtype).kind == "f":
        return (rng.standard_normal(shape) * 100).astype(dtype)
    i = np.iinfo(dtype)
    smooth = np.cumsum(rng.integers(-3, 4, size=shape), axis=-1)            # compressible, with runs
    return np.clip(smooth + int(i.min // 2 + i.max // 2) // 1000, i.min, i.max).astype(dtype)


@pytest.mark.parametrize("compression", ["none", "lzw", "deflate"])
@pytest.mark.parametrize("dtype", DTYPES)
def test_write_read_round_trip(tmp_path, dtype, compression):
    for shape in [(37, 53), (3, 20, 31), (1, 1), (700, 9)]:
        a = _data(dtype, shape, seed=len(shape))
        p = tmp_path / f"{dtype}_{compression}_{len(shape)}_{shape[-1]}.tif"
        gt.write_geotiff(p, a, x_ul=359641.0, y_ul=5156149.0, dx=1000.0, dy=250.0, nodata=-32767, epsg=32633,
                         compression=compression)
        with gt.GeoTiff(p) as t:
            i = t.info
            np.testing.assert_array_equal(t.read(), a)
            np.testing.assert_array_equal(t.read(threads=1), a)
        assert (i.ncol, i.nrow, i.nlayers, i.dtype) == (shape[-1], shape[-2], 1 if len(shape) == 2 else shape[0], np.dtype(dtype))
        assert (i.x_ul, i.y_ul, i.dx, i.dy, i.nodata, i.epsg, i.has_geo) == (359641.0, 5156149.0, 1000.0, 250.0, -32767.0, 32633, True)
        assert i.compression == gt.COMPRESSION[compression]


def test_bigtiff_writer_layout(tmp_path, monkeypatch):
    """The writer switches to BigTIFF above 4 GB; the layout is forced here and must read back
    the same (our reader) with the same tags."""
    monkeypatch.setenv("RFSI_IO_FORCE_BIGTIFF", "1")
    for dtype, comp in (("int16", "lzw"), ("float64", "none"), ("uint8", "deflate")):
        a = _data(dtype, (2, 301, 97), seed=9)
        p = tmp_path / f"big_{dtype}.tif"
        gt.write_geotiff(p, a, 10.0, 20.0, 0.5, 0.25, nodata=-32767, epsg=4326, compression=comp)
        assert p.read_bytes()[:4] == b"II\x2b\x00"
        with gt.GeoTiff(p) as t:
            assert t.info.big and (t.info.x_ul, t.info.y_ul, t.info.dx, t.info.dy, t.info.nodata, t.info.epsg) == \
                (10.0, 20.0, 0.5, 0.25, -32767.0, 4326)
            np.testing.assert_array_equal(t.read(), a)


def test_lzw_codec_corner_cases(tmp_path):
    """Streams that cross every code-width switch, fill the table (ClearCode), and hit the
    KwKwK case; each is also decoded by libtiff (PIL) and libtiff's encoding by us."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(1)
    cases = {
        "random": rng.integers(0, 256, size=(300, 400), dtype=np.uint8),       # incompressible: table overflows repeatedly
        "constant": np.full((200, 500), 7, dtype=np.uint8),                    # aaaa...: KwKwK on every code
        "period2": np.tile(np.array([1, 2], dtype=np.uint8), (100, 300)),
        "ramp": (np.arange(256 * 64) % 251).astype(np.uint8).reshape(64, 256),
        "tiny": np.array([[5]], dtype=np.uint8),
        "i16": rng.integers(-3000, 3000, size=(280, 380)).astype(np.int16),
    }
    for w in (509, 510, 511, 512, 513, 1021, 1022, 1023, 1024, 2047, 2048, 4093, 4094, 4095, 4096):
        cases[f"distinct{w}"] = (rng.permutation(70000)[:w * 3] % 256).astype(np.uint8).reshape(3, w)
    for name, a in cases.items():
        p = tmp_path / f"{name}.tif"
        gt.write_geotiff(p, a, 0.0, float(a.shape[0]), 1.0, 1.0, compression="lzw")
        r, info = gt.read_geotiff(p)
        np.testing.assert_array_equal(r.data, a, err_msg=name)
        np.testing.assert_array_equal(np.array(Image.open(p)).astype(a.dtype), a, err_msg=f"libtiff decode of {name}")
        q = tmp_path / f"{name}_pil.tif"
        Image.fromarray(a).save(q, compression="tiff_lzw")
        np.testing.assert_array_equal(gt.read_geotiff(q)[0].data, a, err_msg=f"libtiff-encoded {name}")


def test_reads_what_libtiff_writes(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    a16 = _data("uint16", (211, 97), 3)
    f32 = _data("float32", (64, 130), 4)
    for comp in ("raw", "tiff_lzw", "tiff_adobe_deflate", "tiff_deflate", "packbits"):
        for a in (a16, f32, _data("uint8", (90, 90), 5), _data("int32", (33, 65), 6)):
            p = tmp_path / f"{comp}_{a.dtype}.tif"
            Image.fromarray(a).save(p, compression=comp)
            np.testing.assert_array_equal(gt.read_geotiff(p)[0].data, a)
    # horizontal differencing (what gdal_translate -co PREDICTOR=2 produces)
    for a in (a16, _data("uint8", (50, 300), 7)):
        p = tmp_path / f"pred2_{a.dtype}.tif"
        Image.fromarray(a).save(p, compression="tiff_lzw", tiffinfo={317: 2})
        with gt.GeoTiff(p) as t:
            assert t.info.predictor == 2
            np.testing.assert_array_equal(t.read(), a)


# ---- files assembled by hand: layouts PIL does not produce --------------------------------

def _build_tiff(a, *, big=False, be=False, tile=None, planar=2, compression=1, predictor=1, chunky=False, extra_tags=()):
    """a: (bands, H, W).  Returns the bytes of a TIFF with the requested layout."""
    E = ">" if be else "<"
    bands, H, W = a.shape
    B = a.dtype.itemsize
    fmt = {"u": 1, "i": 2, "f": 3}[a.dtype.kind]
    src = a.astype(a.dtype.newbyteorder(E))

    def encode(block):                                   # block: (rows, cols, ns) in file byte order
        rows, cols, ns = block.shape
        raw = block
        if predictor == 2:
            nat = np.ascontiguousarray(block.astype(block.dtype.newbyteorder("="))).view(f"u{B}")   # differences of the bit patterns
            d = nat.copy()
            d[:, 1:, :] = nat[:, 1:, :] - nat[:, :-1, :]
            raw = d.astype(np.dtype(f"u{B}").newbyteorder(E))
        by = np.ascontiguousarray(raw).tobytes()
        if predictor == 3:
            be_bytes = np.ascontiguousarray(block.astype(block.dtype.newbyteorder(">"))).view(np.uint8).reshape(rows, cols * ns, B)
            planes = np.ascontiguousarray(be_bytes.transpose(0, 2, 1)).reshape(rows, B * cols * ns)   # MSB plane first
            d = planes.copy()
            d[:, ns:] = planes[:, ns:] - planes[:, :-ns]
            by = d.tobytes()
        return zlib.compress(by) if compression == 8 else by

    chunks = []
    groups = [src.transpose(1, 2, 0)] if chunky else [src[b][:, :, None] for b in range(bands)]
    for g in groups:
        if tile:
            th, tw = tile
            for r0 in range(0, H, th):
                for c0 in range(0, W, tw):
                    blk = np.zeros((th, tw, g.shape[2]), dtype=g.dtype)
                    part = g[r0:r0 + th, c0:c0 + tw]
                    blk[:part.shape[0], :part.shape[1]] = part
                    chunks.append(encode(blk))
        else:
            rps = 7
            for r0 in range(0, H, rps):
                chunks.append(encode(g[r0:r0 + rps]))
    hdr = 16 if big else 8
    offs, at = [], hdr
    body = b""
    for c in chunks:
        offs.append(at)
        body += c + (b"\0" if len(c) & 1 else b"")
        at += len(c) + (len(c) & 1)
    LONG = (16, "Q") if big else (4, "I")
    tags = [(256, 3, [W]), (257, 3, [H]), (258, 3, [B * 8] * bands), (259, 3, [compression]), (262, 3, [1]),
            (277, 3, [bands]), (284, 3, [1 if chunky else planar]), (317, 3, [predictor]), (339, 3, [fmt] * bands)]
    if tile:
        tags += [(322, 3, [tile[1]]), (323, 3, [tile[0]]), (324, LONG[0], offs), (325, LONG[0], [len(c) for c in chunks])]
    else:
        tags += [(278, 3, [7]), (273, LONG[0], offs), (279, LONG[0], [len(c) for c in chunks])]
    tags += list(extra_tags)
    tags.sort(key=lambda t: t[0])
    code = {1: "B", 2: "B", 3: "H", 4: "I", 12: "d", 16: "Q"}
    inl = 8 if big else 4
    blobs, entries = b"", []
    for tid, typ, vals in tags:
        raw = struct.pack(E + code[typ] * len(vals), *vals)
        if len(raw) <= inl:
            val = raw.ljust(inl, b"\0")
        else:
            val = struct.pack(E + ("Q" if big else "I"), at + len(blobs))
            blobs += raw + (b"\0" if len(raw) & 1 else b"")
        entries.append(struct.pack(E + "HH" + ("Q" if big else "I"), tid, typ, len(vals)) + val)
    ifd = at + len(blobs)
    head = (b"MM" if be else b"II") + (struct.pack(E + "HHHQ", 43, 8, 0, ifd) if big else struct.pack(E + "HI", 42, ifd))
    tail = struct.pack(E + ("Q" if big else "H"), len(entries)) + b"".join(entries) + struct.pack(E + ("Q" if big else "I"), 0)
    return head + body + blobs + tail


@pytest.mark.parametrize("layout", [
    dict(), dict(be=True), dict(big=True), dict(big=True, be=True), dict(tile=(16, 32)), dict(tile=(16, 16), be=True, compression=8),
    dict(chunky=True), dict(chunky=True, tile=(32, 16), compression=8), dict(compression=8, predictor=2),
    dict(chunky=True, compression=8, predictor=2, be=True), dict(tile=(16, 16), predictor=2, big=True),
])
@pytest.mark.parametrize("dtype", ["int16", "uint8", "int32", "float32", "float64"])
def test_hand_built_layouts(tmp_path, layout, dtype):
    a = _data(dtype, (3, 45, 70), seed=11)
    p = tmp_path / "t.tif"
    p.write_bytes(_build_tiff(a, **layout))
    with gt.GeoTiff(p) as t:
        assert (t.info.ncol, t.info.nrow, t.info.nlayers) == (70, 45, 3)
        assert t.info.big == bool(layout.get("big")) and t.info.tiled == ("tile" in layout)
        np.testing.assert_array_equal(t.read(), a)


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("layout", [dict(), dict(chunky=True), dict(tile=(16, 16), be=True)])
def test_floating_point_predictor(tmp_path, dtype, layout):
    a = _data(dtype, (2, 30, 41), seed=5)
    p = tmp_path / "p3.tif"
    p.write_bytes(_build_tiff(a, compression=8, predictor=3, **layout))
    np.testing.assert_array_equal(gt.read_geotiff(p)[0].data, a)


def test_georeference_variants(tmp_path):
    a = _data("int16", (1, 10, 12))
    scale, tie = (33550, 12, [0.5, 0.25, 0.0]), (33922, 12, [0.0, 0.0, 0.0, 100.0, 50.0, 0.0])
    p = tmp_path / "g.tif"
    # PixelIsPoint: the tie point is the centre of cell (0,0); the extent moves out by half a cell
    keys = (34735, 3, [1, 1, 0, 3, 1024, 0, 1, 2, 1025, 0, 1, 2, 2048, 0, 1, 4326])
    p.write_bytes(_build_tiff(a, extra_tags=[scale, tie, keys, (42113, 2, list(b"-32767\0"))]))
    i = gt.GeoTiff(p).info
    assert (i.x_ul, i.y_ul, i.dx, i.dy, i.epsg, i.nodata) == (99.75, 50.125, 0.5, 0.25, 4326, -32767.0)
    # tie point given for an interior cell
    p.write_bytes(_build_tiff(a, extra_tags=[scale, (33922, 12, [2.0, 4.0, 0.0, 101.0, 49.0, 0.0])]))
    i = gt.GeoTiff(p).info
    assert (i.x_ul, i.y_ul, i.nodata, i.epsg) == (100.0, 50.0, None, 0)
    # axis-aligned ModelTransformation instead of scale + tie point; rotation is refused
    m = [0.5, 0, 0, 100.0, 0, -0.25, 0, 50.0, 0, 0, 0, 0, 0, 0, 0, 1]
    p.write_bytes(_build_tiff(a, extra_tags=[(34264, 12, m)]))
    i = gt.GeoTiff(p).info
    assert (i.x_ul, i.y_ul, i.dx, i.dy) == (100.0, 50.0, 0.5, 0.25)
    m[1] = 0.1
    p.write_bytes(_build_tiff(a, extra_tags=[(34264, 12, m)]))
    with pytest.raises(gt.GeoTiffError) as e:
        gt.GeoTiff(p)
    assert e.value.code == -3
    # no georeference at all: raster()'s default extent 0..ncol, 0..nrow
    p.write_bytes(_build_tiff(a))
    i = gt.GeoTiff(p).info
    assert (i.has_geo, i.x_ul, i.y_ul, i.dx, i.dy) == (False, 0.0, 10.0, 1.0, 1.0)


def test_products_inherit_grid_and_crs_of_the_dem(tmp_path):
    dem = np.where(np.arange(280 * 380).reshape(280, 380) % 7 == 0, -32767, 100).astype(np.float32)
    keys = gt.GeoKeys(np.array([1, 1, 0, 7, 1024, 0, 1, 1, 1025, 0, 1, 1, 1026, 34737, 33, 0, 2049, 34737, 7, 33, 2054, 0,
                                1, 9102, 2062, 34736, 1, 0, 3072, 0, 1, 32767], dtype=np.uint16),
                      np.array([0.0]), b"UTM Zone 33, Northern Hemisphere|WGS 84|")
    gt.write_geotiff(tmp_path / "dem.tif", dem, 0.158, 42.858, 0.0083333333, 0.0083333333, nodata=-32767, geokeys=keys)
    with gt.GeoTiff(tmp_path / "dem.tif") as t:
        like = t.info
    np.testing.assert_array_equal(like.geokeys.directory, keys.directory)
    assert like.geokeys.ascii == keys.ascii and like.epsg == 32767
    pred = np.arange(280 * 380, dtype=np.int64).astype(np.int16)
    gt.write_product(tmp_path / "20160101.tif", pred, like)
    r, info = gt.read_geotiff(tmp_path / "20160101.tif")
    np.testing.assert_array_equal(r.data.ravel(), pred)
    assert info.dtype == np.int16 and info.nodata == -32767.0 and info.compression == 5        # INT2S, NAflag, LZW
    assert (info.x_ul, info.y_ul, info.dx, info.dy) == (like.x_ul, like.y_ul, like.dx, like.dy)
    np.testing.assert_array_equal(info.geokeys.directory, keys.directory)
    with pytest.raises(ValueError):
        gt.write_product(tmp_path / "x.tif", pred.astype(np.float64), like)


def test_errors(tmp_path):
    with pytest.raises(gt.GeoTiffError) as e:
        gt.GeoTiff(tmp_path / "missing.tif")
    assert e.value.code == -1
    (tmp_path / "junk.tif").write_bytes(b"not a tiff at all")
    with pytest.raises(gt.GeoTiffError) as e:
        gt.GeoTiff(tmp_path / "junk.tif")
    assert e.value.code == -2
    good = _build_tiff(_data("int16", (1, 40, 40)), compression=8)
    (tmp_path / "cut.tif").write_bytes(good[:len(good) // 2])               # directory gone
    with pytest.raises(gt.GeoTiffError):
        gt.GeoTiff(tmp_path / "cut.tif")
    bad = bytearray(good)
    bad[40:60] = b"\xff" * 20                                                 # damaged deflate stream
    (tmp_path / "bad.tif").write_bytes(bytes(bad))
    with gt.GeoTiff(tmp_path / "bad.tif") as t, pytest.raises(gt.GeoTiffError) as e:
        t.read()
    assert e.value.code == -2
    jpeg = _build_tiff(_data("uint8", (1, 8, 8)), compression=1).replace(struct.pack("<HHIHH", 259, 3, 1, 1, 0),
                                                                       struct.pack("<HHIHH", 259, 3, 1, 7, 0))
    (tmp_path / "jpeg.tif").write_bytes(jpeg)
    with pytest.raises(gt.GeoTiffError) as e:
        gt.GeoTiff(tmp_path / "jpeg.tif")
    assert e.value.code == -3
    with pytest.raises(ValueError):
        gt.write_geotiff(tmp_path / "c.tif", np.zeros((2, 2), dtype=np.complex64), 0, 2, 1, 1)
    with pytest.raises(gt.GeoTiffError) as e:
        gt.write_geotiff(tmp_path / "no_such_dir" / "c.tif", np.zeros((2, 2)), 0, 2, 1, 1)
    assert e.value.code == -1


def test_reader_survives_mutated_files(tmp_path):
    """A reader of files from disk must reject damage, not crash on it (tools/fuzz_geotiff.cpp is the
    long-running ASan version of this; it found a 64-bit offset wrap-around in BigTIFF chunk checks)."""
    rng = np.random.default_rng(7)
    a = _data("int16", (2, 24, 40), 1)
    seeds = [_build_tiff(a, **l) for l in (dict(), dict(be=True), dict(big=True), dict(tile=(16, 16)),
                                           dict(chunky=True, compression=8, predictor=2),
                                           dict(tile=(16, 16), compression=8, big=True, be=True))]
    gt.write_geotiff(tmp_path / "lzw.tif", _data("int16", (50, 60), 3), 0, 50, 1, 1, nodata=-32767, epsg=4326)
    seeds.append((tmp_path / "lzw.tif").read_bytes())
    # the wrap-around case itself: a BigTIFF whose strip offset + byte count overflows 64 bits
    big = bytearray(_build_tiff(_data("uint8", (1, 7, 8)), big=True))
    ifd = struct.unpack_from("<Q", big, 8)[0]
    n = struct.unpack_from("<Q", big, ifd)[0]
    for e in range(n):
        at = ifd + 8 + 20 * e
        if struct.unpack_from("<H", big, at)[0] == 273:
            struct.pack_into("<Q", big, at + 12, 0xFFFFFFFFFFFFFFF0)
    (tmp_path / "wrap.tif").write_bytes(bytes(big))
    with pytest.raises(gt.GeoTiffError):
        gt.GeoTiff(tmp_path / "wrap.tif")
    ok = bad = 0
    p = tmp_path / "m.tif"
    for base in seeds:
        for _ in range(150):
            v = bytearray(base)
            for _ in range(int(rng.integers(1, 6))):
                pos = int(rng.integers(0, len(v)))
                if rng.random() < 0.5:      # aim at the directory
                    ifd = struct.unpack_from(">I" if v[:2] == b"MM" else "<I", v, 4)[0]
                    if ifd < len(v):
                        pos = min(len(v) - 1, ifd + int(rng.integers(0, 300)))
                v[pos] = int(rng.integers(0, 256))
            if rng.random() < 0.1:
                v = v[:int(rng.integers(1, len(v)))]
            p.write_bytes(bytes(v))
            try:
                with gt.GeoTiff(p) as t:
                    if t.info.ncol * t.info.nrow * t.info.nlayers < 10_000_000:
                        t.read(threads=2)
                ok += 1
            except gt.GeoTiffError:
                bad += 1
    assert ok + bad == 150 * len(seeds) and bad > 100


@pytest.mark.skipif(not Path("/root/reference/prcp_catalonia").exists(), reason="reference data not on this box")
def test_reference_geotiffs_decode_to_the_golden_digests():
    gold = json.loads((GOLD / "geotiff_reference.json").read_text())
    assert len(gold) == 15
    for rel, g in gold.items():
        r, info = gt.read_geotiff(Path("/root/reference") / rel)
        assert hashlib.sha256(np.ascontiguousarray(r.data).tobytes()).hexdigest() == g["sha256"], rel
        assert [info.ncol, info.nrow, str(info.dtype), info.nodata, info.epsg] == g["meta"], rel
        assert [info.x_ul, info.y_ul, info.dx, info.dy] == g["geo"], rel
