"""Per-kernel table (launches, device time, DRAM bytes) from an ncu csv of ONE step
(ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv), and the
entry of profiles/r02_dram_bytes.json that bench.py reads for roofline.traffic.
usage: python tools/ncu_traffic.py <launches.csv> <workload key> <batch> [--update]"""
import collections
import csv
import json
import os
import re
import sys

path, key, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = list(csv.DictReader([l for l in open(path) if l.startswith('"')]))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0,
        "nsecond": 1e-3, "msecond": 1e3}
agg = collections.defaultdict(lambda: {"launches": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
seen = set()
for r in rows:
    kn = r["Kernel Name"]
    m = re.search(r"gemm_nt_kernel<(?:gpb::)?(\w+), *(?:\(int\))?(\d+), *(?:\(int\))?(\d+), *(?:\(int\))?(\d+)", kn)
    name = ("gemm_nt_kernel<%s,%sx%s,%s>" % (m.group(1), m.group(2), m.group(3), "tma" if m.group(4) != "0" else "cp.async")
            if m else re.sub(r"\(.*", "", kn).replace("void ", ""))
    v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    a = agg[name]
    mn = r["Metric Name"]
    if mn == "gpu__time_duration.sum":
        a["us"] += v
        a["launches"] += 1
    elif mn == "dram__bytes_read.sum":
        a["rd"] += v
    elif mn == "dram__bytes_write.sum":
        a["wr"] += v
tot_us = sum(a["us"] for a in agg.values())
print(f"# {path}: {sum(a['launches'] for a in agg.values())} launches, {tot_us / 1e3:.2f} ms summed device time "
      "(under ncu: serialised, cold caches -- compare shares)")
print(f"{'kernel':44s} {'launches':>8s} {'total us':>11s} {'share':>7s} {'DRAM rd GB':>11s} {'DRAM wr GB':>11s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    print(f"{k:44s} {a['launches']:8d} {a['us']:11.1f} {100 * a['us'] / tot_us:6.1f}% {a['rd'] / 1e9:11.3f} {a['wr'] / 1e9:11.3f}")
gemm = {k: a for k, a in agg.items() if k.startswith("gemm_nt_kernel")}
entry = {"batch": batch, "source": os.path.basename(path),
         "gemm_launches_per_step": sum(a["launches"] for a in gemm.values()),
         "gemm_dram_bytes_per_step": sum(a["rd"] + a["wr"] for a in gemm.values()),
         "all_dram_bytes_per_step": sum(a["rd"] + a["wr"] for a in agg.values()),
         "gemm_time_share_under_ncu": sum(a["us"] for a in gemm.values()) / tot_us}
print(json.dumps({key: entry}))
if "--update" in sys.argv:
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_dram_bytes.json")
    try:
        table = json.load(open(out))
    except (OSError, ValueError):
        table = {}
    table[key] = entry
    json.dump(table, open(out, "w"), indent=1, sort_keys=True)
