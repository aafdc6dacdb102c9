#!/usr/bin/env python
"""Per-kernel device time of one `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised:
read the SHARES, not the absolutes).  python tools/launch_summary.py gpurun_out/launches.csv "<command that ran>" """
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
t, n = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[iu], 1e-6)
    t[r[ik][:80]] += v
    n[r[ik][:80]] += 1
tot = sum(t.values())
print(f"# per-kernel device time of `{sys.argv[2] if len(sys.argv) > 2 else '?'}` under ncu (cold-cache, serialised: shares, not absolutes)")
for k, v in t.most_common(14):
    print(f"{v:10.2f} ms  {100 * v / tot:4.1f} %  x {n[k]:3d}  {k}")
