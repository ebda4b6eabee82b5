"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:

    python tools/launch_summary.py gpurun_out/launches.csv [launches_per_step] > profiles/rN_launches_..._summary.txt
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if r is not hdr and len(r) == len(hdr) and r[ix["Metric Name"]] == "gpu__time_duration.sum"]
tot, cnt = collections.Counter(), collections.Counter()
for r in data:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("nvs::", "")
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
    tot[name] += us
    cnt[name] += 1
n = len(data)
steps = n / per_step if per_step else 1
print(f"# {path}: {n} launches" + (f" = {steps:.1f} steps of {per_step} launches" if per_step else ""))
all_us = sum(tot.values())
print(f"# per step: {all_us / steps / 1e3:.3f} ms (cold-cache, serialised by the profiler: shares, not absolute times)")
for k, v in tot.most_common():
    print(f"{v / steps:10.1f} us {100 * v / all_us:5.1f}%  x{cnt[k] / steps:<5.1f} {k}")
