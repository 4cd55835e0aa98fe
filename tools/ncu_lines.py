"""Per-source-line instruction and stall-sample shares of one profiled launch (read here from an .ncu-rep).
python tools/ncu_lines.py gpurun_out/x.ncu-rep <launch index> [top]"""
import csv
import subprocess
import sys

rep, launch, top = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[start]
ci = {h: i for i, h in enumerate(hdr)}
agg, ti, ts = {}, 0, 0
for r in rows[start + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        ln, ie, ss = int(r[0]), int(r[ci["Instructions Executed"]] or 0), int(r[ci["# Samples"]] or 0)
    except ValueError:
        continue
    a = agg.setdefault(ln, [r[1], 0, 0])
    a[1] += ie
    a[2] += ss
    ti += ie
    ts += ss
print(f"warp instructions {ti}, samples {ts}")
for ln, (src, ie, ss) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{ln:4d} {100 * ie / ti:5.1f}%i {100 * ss / max(ts, 1):5.1f}%s  {src.strip()[:120]}")
