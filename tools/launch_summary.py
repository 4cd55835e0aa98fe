"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name.
python tools/launch_summary.py gpurun_out/final_launches.csv profiles/<tag>_launches_summary.csv"""
import csv
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.reader(lines))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("spr::(anonymous namespace)::", "").replace("spr::", "")
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}[r[ui]]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(dst, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "total_us", "share_pct"])
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, n, f"{t:.1f}", f"{100 * t / tot:.2f}"])
print(f"{dst}: {len(agg)} kernels, {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms")
