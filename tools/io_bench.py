#!/usr/bin/env python
"""Host-side raster I/O timings (CPU only): librfsi_io.so against libtiff (through PIL) on the
reference's own GeoTIFFs and on a product-sized Int16 raster.  python tools/io_bench.py [threads]"""
import glob
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from rfsi_b200 import geotiff as gt

try:
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
except Exception:
    Image = None


def best(fn, reps=5):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return min(ts)


threads = int(sys.argv[1]) if len(sys.argv) > 1 else 0
print(f"host cores: {os.cpu_count()}, decode threads: {threads or 'all'}")
ref = sorted(glob.glob("/root/reference/**/*.tif", recursive=True))
for p in [f for f in ref if "20160101" in f or f.endswith("dem.tif") or f.endswith("dem_cat.tif")]:
    with gt.GeoTiff(p) as t:
        i = t.info
        mb = i.ncol * i.nrow * i.nlayers * i.dtype.itemsize / 1e6
        ours = best(lambda: t.read(threads))
    line = f"read  {Path(p).name[:44]:44s} {i.ncol}x{i.nrow} {str(i.dtype):8s} LZW: {ours * 1e3:7.2f} ms ({mb / ours:7.0f} MB/s decoded)"
    if Image is not None:
        pil = best(lambda: np.array(Image.open(p)))
        line += f" | libtiff {pil * 1e3:7.2f} ms"
    print(line)
# a product the size of one C5 band: 10^4 x 2000 Int16 (what predict_maps writes per day and band)
rng = np.random.default_rng(0)
prod = (np.cumsum(rng.integers(-3, 4, size=(2000, 10000)), axis=1) + 1500).astype(np.int16)
prod[rng.random(prod.shape) < 0.3] = -32767
with tempfile.TemporaryDirectory() as td:
    for comp in ("none", "lzw", "deflate"):
        f = Path(td) / f"p_{comp}.tif"
        w = best(lambda: gt.write_geotiff(f, prod, 0.0, 2000.0, 1.0, 1.0, nodata=-32767, epsg=32633, compression=comp), 3)
        r = best(lambda: gt.read_geotiff(f, threads=threads), 3)
        print(f"write 10000x2000 int16 product, {comp:8s}: {w * 1e3:7.1f} ms ({prod.nbytes / 1e6 / w:6.0f} MB/s), "
              f"file {f.stat().st_size / 1e6:6.1f} MB, read back {r * 1e3:7.1f} ms")
