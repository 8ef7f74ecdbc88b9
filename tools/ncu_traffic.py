"""ncu CSV (metrics dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum; one row per kernel launch and metric) ->
profiles/r02_traffic.json: dram bytes per launch of `raster_kernel` and of one `b2s_step` (the sum over every other kernel of the
profiled window, i.e. the 31 nodes of the control-step graph), plus a markdown table per kernel.  bench.py reads the json for
`roofline.traffic`."""
import collections
import csv
import json
import os
import sys

path = sys.argv[1]
out_json = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_traffic.json")
rows = [r for r in csv.reader(open(path, errors="ignore")) if r]
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
hdr = rows[hdr_i]
ik, im, iv, iu, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
launch = collections.OrderedDict()
for r in rows[hdr_i + 1:]:
    if len(r) <= iv:
        continue
    key = r[iid]
    full = r[ik]
    short = full[5:] if full.startswith("void ") else full
    short = short.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    short = short.split("<")[0].split("(")[0].split("::")[-1].strip()
    d = launch.setdefault(key, dict(name=short, full=full))
    try:
        d[r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1)
    except ValueError:
        pass
per = collections.OrderedDict()
for d in launch.values():
    p = per.setdefault(d["name"], dict(launches=0, bytes=0.0, ms=0.0))
    p["launches"] += 1
    p["bytes"] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    p["ms"] += d.get("gpu__time_duration.sum", 0)
step = sum(p["bytes"] for n, p in per.items() if n != "raster_kernel")
res = {"b2s_step": step, "raster_kernel": per.get("raster_kernel", {}).get("bytes", 0) / max(per.get("raster_kernel", {}).get("launches", 1), 1),
       "source": os.path.basename(path), "how": "ncu --profile-from-start off --cache-control none --clock-control none, one control step after warm-up",
       "per_kernel": {n: p for n, p in per.items()}}
json.dump(res, open(out_json, "w"), indent=1)
print("| kernel | launches | dram MB (read+write, all launches) | ms (sum, under ncu) |\n|---|---|---|---|")
for n, p in per.items():
    print(f"| {n} | {p['launches']} | {p['bytes'] / 1e6:.2f} | {p['ms']:.4f} |")
print(f"\nb2s_step total {step / 1e6:.1f} MB per control step; raster_kernel {res['raster_kernel'] / 1e6:.1f} MB per launch")
