"""Digests of the reference's 15 GeoTIFFs as decoded by libtiff (through PIL), plus the
georeference PIL reports in the raw tags, for tests/test_geotiff.py to hold librfsi_io.so to.
Run in the build container (needs /root/reference):  python tests/golden/make_geotiff_golden.py"""
import hashlib
import json
from pathlib import Path

import numpy as np
from PIL import Image

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "geotiff_reference.json"


def main():
    gold = {}
    for p in sorted(REF.rglob("*.tif")):
        im = Image.open(p)
        tags = im.tag_v2
        bits, fmt = tags[258][0], tags.get(339, (1,))[0]
        dt = np.dtype({(16, 1): "uint16", (16, 2): "int16", (32, 2): "int32", (32, 3): "float32", (8, 1): "uint8"}[(bits, fmt)])
        a = np.array(im).astype(dt)                      # PIL widens int16 to int32: values are what matter
        sx, sy = tags[33550][0], tags[33550][1]
        tie = tags[33922]
        keys = tags[34735]
        kv = {keys[4 * k]: keys[4 * k + 3] for k in range(1, keys[3] + 1) if keys[4 * k + 1] == 0}
        assert kv.get(1025, 1) == 1                      # all PixelIsArea
        nod = tags.get(42113)
        gold[str(p.relative_to(REF))] = {
            "sha256": hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest(),
            "meta": [im.size[0], im.size[1], str(dt), None if nod is None else float(nod), int(kv.get(3072, kv.get(2048, 0)))],
            "geo": [tie[3] - tie[0] * sx, tie[4] + tie[1] * sy, sx, sy],
        }
    OUT.write_text(json.dumps(gold, indent=1, sort_keys=True) + "\n")
    print(len(gold), "files ->", OUT)


if __name__ == "__main__":
    main()
