"""profiles/kpconv_traffic.json (read by bench.py for roofline.traffic) from an ncu --set full capture of every KPConv
launch of one bench step.   python tools/make_traffic.py gpurun_out/kpconv_step.ncu-rep profiles/<tag>_kpconv_step_ncu_summary.csv"""
import csv
import json
import os
import subprocess
import sys

rep, summary = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[0], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
unit = {h: rows[1][i] for i, h in enumerate(hdr)}


def to_bytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


rd = [to_bytes(r[col["dram__bytes_read.sum"]], unit["dram__bytes_read.sum"]) for r in data]
wr = [to_bytes(r[col["dram__bytes_write.sum"]], unit["dram__bytes_write.sum"]) for r in data]
names = [r[col["Kernel Name"]] for r in data]
n = len(data)
out = {
    "kernel": "every KPConv launch of one bench step (k_kpconv_cin1 stem + k_kpconv_tc layers), 4-stage, 32 pairs",
    "launches": n,
    "dram_bytes_per_launch": (sum(rd) + sum(wr)) / n,
    "dram_bytes_per_step": sum(rd) + sum(wr),
    "dram_read_per_step": sum(rd),
    "dram_write_per_step": sum(wr),
    "per_launch": [{"kernel": nm[:60], "dram_read": a, "dram_write": b} for nm, a, b in zip(names, rd, wr)],
    "source": f"{os.path.basename(summary)} (ncu --set full --clock-control none, tools/ncu_kpconv.sh)",
}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump(out, open(os.path.join(root, "profiles", "kpconv_traffic.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "per_launch"}, indent=1))
subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), rep, summary], check=True)
