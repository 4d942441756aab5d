#!/usr/bin/env python3
"""Builds profiles/<tag>_sptrsv_traffic.json from an ncu CSV captured with
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:"k_(fwd|bwd|spmv)" --csv --log-file <csv> python tools/gpu_big.py --profile <grid>
(one solve = one forward+backward sweep + one residual SpMV).  bench.py reads the result as `roofline.traffic`."""
import collections, csv, json, sys
src, out, grid = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = list(csv.DictReader([l for l in open(src) if not l.startswith("==")]))
per = collections.OrderedDict()
ids = collections.defaultdict(set)
for r in rows:
    name = r["Kernel Name"].split("(")[0].replace("b200::", "")
    k = per.setdefault(name, {"launches": 0, "time_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
    ids[name].add(r["ID"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        k["time_us"] += v / 1e3 if unit == "ns" else v * (1e3 if unit == "ms" else 1.0)
    else:
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        k["dram_read_bytes" if "read" in m else "dram_write_bytes"] += v * scale
for name in per:  # per LAUNCH (the capture may hold several solves)
    nl = per[name]["launches"] = len(ids[name])
    for key in ("time_us", "dram_read_bytes", "dram_write_bytes"):
        per[name][key] /= nl
sweep = sum(v["dram_read_bytes"] + v["dram_write_bytes"] for k, v in per.items() if "spmv" not in k)
spmv = sum(v["dram_read_bytes"] + v["dram_write_bytes"] for k, v in per.items() if "spmv" in k)
json.dump({"workload": "5-point 2D Laplacian %dx%d, one forward+backward SpTRSV sweep (no refinement step)" % (grid, grid), "grid": grid,
           "source": src if src.startswith("profiles/") else "profiles/" + src.split("/")[-1], "kernels": per,
           "sptrsv_sweep_traffic_bytes": sweep, "spmv_traffic_bytes": spmv}, open(out, "w"), indent=1)
print(out, "sweep traffic %.1f MB" % (sweep / 1e6), "time %.1f us" % sum(v["time_us"] for k, v in per.items() if "spmv" not in k))
