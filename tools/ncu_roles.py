"""Warp-instruction and stall-sample shares of one profiled launch, per source FILE and (for the kernel's own file) per
warp role, from an .ncu-rep with imported sources:  python tools/ncu_roles.py rep.ncu-rep path/to/kernel.cu [launch]"""
import csv
import re
import subprocess
import sys

rep, kernel_src = sys.argv[1], sys.argv[2]
launch = sys.argv[3] if len(sys.argv) > 3 else "0"
src = open(kernel_src).read().splitlines()
marks = [(i, m.group(1)) for i, l in enumerate(src, 1) for m in [re.search(r"// =+ (.*?) =+", l)] if m]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", launch,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
agg, lines = {}, {}
for hi, h in enumerate(heads):
    end = heads[hi + 1] if hi + 1 < len(heads) else len(rows)
    fname = ""
    for r in rows[max(0, h - 6):h]:
        if r and r[0] in ("File Path", "File Name", "Source File"):
            fname = r[1]
    hdr = rows[h]
    ci = {k: i for i, k in enumerate(hdr)}
    cur_line, cur_txt = None, ""
    for r in rows[h + 1:end]:
        if len(r) < len(hdr):
            continue
        if r[0].strip().isdigit():
            cur_line, cur_txt = int(r[0]), r[1]
            try:
                ie, ss = int(r[ci["Instructions Executed"]] or 0), int(r[ci["# Samples"]] or 0)
            except ValueError:
                continue
            own = kernel_src.split("/")[-1] in fname or not fname
            role = "?"
            if own and hi == 0:
                role = "preamble"
                for i, name in marks:
                    if cur_line >= i:
                        role = name
            key = (hi, fname.split("/")[-1] or f"section{hi}", role)
            a = agg.setdefault(key, [0, 0])
            a[0] += ie
            a[1] += ss
            lines[(hi, cur_line, cur_txt.strip()[:70])] = (ie, ss)
ti = sum(a[0] for a in agg.values())
ts = sum(a[1] for a in agg.values())
print(f"warp instructions {ti}, samples {ts}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"section {k[0]} {k[1]:28s} {k[2]:28s} {100 * a[0] / ti:5.1f}%i {100 * a[1] / max(ts, 1):5.1f}%s")
print("top lines:")
for k, (ie, ss) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"  s{k[0]} {k[1]:4d} {100 * ie / ti:5.1f}%i {100 * ss / max(ts, 1):5.1f}%s  {k[2]}")
