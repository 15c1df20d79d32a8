"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    kn = r["Kernel Name"]
    m = re.search(r"gemm_nt_kernel<(?:gpb::)?(\w+), *(?:\(int\))?(\d+), *(?:\(int\))?(\d+), *(?:\(int\))?(\d+)", kn)
    if m:
        name = "gemm_nt_kernel<%s,%sx%s,%s>" % (m.group(1), m.group(2), m.group(3), "tma" if m.group(4) == "1" else "cp.async")
    else:
        name = re.sub(r"\(.*", "", kn).replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += float(r["Metric Value"]) / 1e3
tot = sum(v[1] for v in agg.values())
print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms summed device time "
      "(cold-cache, serialised: compare shares)")
print(f"{'kernel':42s} {'launches':>8s} {'total us':>12s} {'share':>7s} {'avg us':>10s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:42s} {v[0]:8d} {v[1]:12.1f} {100 * v[1] / tot:6.1f}% {v[1] / v[0]:10.1f}")
