"""Timestamp trace of the gather kernel's hand-overs (needs libspr built with -DSPR_G_TRACE, see tools/r2_trace.sh)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from superpoints_registration_b200 import _lib, ops
from superpoints_registration_b200.kernel_points import load_kernels
c = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda:0"
rng = np.random.default_rng(0)
n, H, r = 148 * 64 * 4, 40, 0.05
lens = np.array([n], np.int32)
pts = rng.uniform(0, 1, size=(n, 3)).astype(np.float32); pts[:, 2] *= 0.12
tp, tl = torch.from_numpy(pts).to(dev), torch.from_numpy(lens).to(dev)
grid = ops.CellGrid(tp, tl, r)
idx, mc = grid.query(tp, tl, H, index_dtype=torch.int32)
print("valid/row", float((idx < n).sum()) / n)
x = torch.from_numpy(rng.normal(size=(n, c)).astype(np.float32)).to(dev)
prep = ops.instance_norm_lrelu_ex(x, tl, slope=0.1, want_f32=False, kpconv_points=tp)["kpconv"]
w = torch.from_numpy((rng.normal(size=(15, c, c)) / np.sqrt(15 * c)).astype(np.float32)).to(dev)
kp = torch.from_numpy(load_kernels(r, 15)).to(dev)
for _ in range(2):
    out = ops.kpconv_forward_prepared(tp, idx, prep, w, kp, r * 0.8, generation=2)
torch.cuda.synchronize()
L = ctypes.CDLL(_lib.LIB_PATH)
buf = np.zeros(8 * 4096, np.int64)
rc = L.spr_kpconv_g_trace(buf.ctypes.data_as(ctypes.c_void_p))
t = buf.reshape(8, 4096)
t0 = t[0, 0]
names = ["P0 before slot wait", "P0 slot free", "P0 arrived", "MMA full seen", "MMA d1free seen", "MMA committed", "R0 d1full seen", "R0 ld done"]
nsl = 8 if c == 32 else 6
print("first 6 fills of producer slot 0 (clocks since start):")
for u in range(1, 7):
    q = u * nsl   # global query index in CTA order handled by slot 0 at use u (pass structure ignored for C=32)
    print(f" use {u}: P wait-start {t[0,u]-t0:7d} free {t[1,u]-t0:7d} arrive {t[2,u]-t0:7d} | MMA(q={q}) full {t[3,q]-t0:7d} d1free {t[4,q]-t0:7d} commit {t[5,q]-t0:7d}")
d = np.diff(t[5, :400])
print("MMA commit-to-commit interval: median", np.median(d), "mean", d.mean())
print("MMA wait for full  (full seen - previous commit): median", np.median(t[3, 1:400] - t[5, 0:399]))
print("MMA wait for d1free: median", np.median(t[4, 1:400] - t[3, 1:400]))
print("P0: slot wait median", np.median(t[1, 1:50] - t[0, 1:50]), " fill (free->arrive) median", np.median(t[2, 1:50] - t[1, 1:50]), " period median", np.median(np.diff(t[2, 1:50])))
print("P0 arrive -> MMA full seen (copy latency + queue): median", np.median([t[3, u * nsl] - t[2, u] for u in range(1, 40)]))
print("R0: ld latency median", np.median(t[7, 1:100] - t[6, 1:100]), " pair period median", np.median(np.diff(t[6, 1:100])))
