#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of the fused kernel at the bench shape: the ncu-derived
fields of bench.py's roofline record, stamped with the hash of the kernel sources they were measured on (bench.py
compares it with the sources it runs and flags `ncu_stale`).
python tools/traffic_from_ncu.py gpurun_out/fused.ncu-rep <pixel-days per launch> <trees> <summary file under profiles/>"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

rep, units, trees, summary = sys.argv[1], float(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))


def val(k):
    x = float(d[k].replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u[k], 1.0)
    return x * scale


warp_trees = units / 32 * trees
rec = {
    "source_sha": bench.source_sha(),
    "captured": summary,
    "kernel": d.get("Kernel Name", "")[:80],
    "pixel_days_per_launch": units,
    "launch_ms_under_ncu": val("gpu__time_duration.sum") / (1e6 if u["gpu__time_duration.sum"] in ("ns", "nsecond") else 1.0),
    "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
    "dram_bytes_per_pixel_day": (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / units,
    "issue_active_frac": val("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100,
    "l1_sector_hit_frac": val("l1tex__t_sector_hit_rate.pct") / 100,
    "l1tex_data_pipe_frac": val("l1tex__throughput.avg.pct_of_peak_sustained_elapsed") / 100,
    "l2_sector_hit_frac": val("lts__t_sector_hit_rate.pct") / 100,
    "global_load_requests_per_warp_tree": val("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum") / warp_trees,
    "l1_sectors_per_request": val("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum") / val("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"),
    "instructions_per_warp_tree": val("smsp__inst_executed.sum") / warp_trees,
}
p = ROOT / "profiles" / "traffic.json"
allrec = json.loads(p.read_text()) if p.exists() else {}
allrec["fused_kernel"] = rec
p.write_text(json.dumps(allrec, indent=1) + "\n")
print(json.dumps(rec, indent=1))
