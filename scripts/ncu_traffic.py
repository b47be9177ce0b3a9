#!/usr/bin/env python
"""DRAM traffic per launch of every kernel in `ncu --set full` captures, as JSON for bench.py's
roofline.traffic:  ncu_traffic.py workload=file.ncu-rep [workload=file.ncu-rep ...] > profiles/rNN_traffic.json
(dram__bytes_read.sum + dram__bytes_write.sum, averaged over the captured launches of each kernel)"""
import csv, io, json, subprocess, sys, collections

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
        "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}     # times in microseconds
SHORT = {"k_density": "density", "k_force": "force", "k_advect_bin": "advect_bin", "k_scan": "scan",
         "k_scatter_ids": "scatter_ids", "k_reorder": "reorder"}


def load(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        i = idx[name]
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    acc = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        key = next((v for k, v in SHORT.items() if k in name), name.split("(")[0])
        acc[key]["dram_read_bytes"].append(val(r, "dram__bytes_read.sum"))
        acc[key]["dram_write_bytes"].append(val(r, "dram__bytes_write.sum"))
        acc[key]["ncu_duration_us"].append(val(r, "gpu__time_duration.sum"))
        acc[key]["issue_active_pct"].append(val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"))
        acc[key]["warp_instructions"].append(val(r, "smsp__inst_executed.sum"))
    out = {}
    for k, d in acc.items():
        out[k] = {m: round(sum(v) / len(v), 3) for m, v in d.items()}
        out[k]["traffic_bytes"] = round(out[k]["dram_read_bytes"] + out[k]["dram_write_bytes"], 1)
        out[k]["launches_captured"] = len(d["ncu_duration_us"])
    return out


res = {}
for a in sys.argv[1:]:
    wl, path = a.split("=", 1)
    res.setdefault(wl, {}).update(load(path))
json.dump(res, sys.stdout, indent=1)
print()
