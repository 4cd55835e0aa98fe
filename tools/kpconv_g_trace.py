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
prep = ops.instance_norm_lrelu_ex(x, tl, slope=0.1, want_f32=False, kpconv_points=tp, kpconv_planar=True)["kpconv"]
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
# CTA 0, slot pair 0 (needs NSLOT = 8: C = 32, or H <= 32): fill f of the pair is produced by warp half f % 2 (traced: slot 0,
# half 0 -> even fills, item f / 2), multiplied by issuer 0 (its event 2 f; 2 f + 1 is pair 2) and read back by group 0
print("fill:  P wait-start  slot-free  arrived | MMA full-seen d1free-seen committed | R d1full-seen ld-done   (clocks)")
for f in range(2, 26, 2):
    i = f // 2
    print(f" {f:3d}: {t[0,i]-t0:9d} {t[1,i]-t0:9d} {t[2,i]-t0:9d} | {t[3,2*f]-t0:9d} {t[4,2*f]-t0:9d} {t[5,2*f]-t0:9d} |"
          f" {t[6,f]-t0:9d} {t[7,f]-t0:9d}")
F = np.arange(4, 200, 2)
I = F // 2
med = lambda a: float(np.median(a))
print("pair-0 fill period (R d1full seen, per fill):", med(np.diff(t[6, 4:200])))
print("P: prepare (previous arrive -> wait-start)", med(t[0, I] - t[2, I - 1]), " slot wait", med(t[1, I] - t[0, I]),
      " copy issue (free -> arrived)", med(t[2, I] - t[1, I]), " own period", med(np.diff(t[2, I])))
print("P arrived -> MMA full seen (copy latency + other slot + queue)", med(t[3, 2 * F] - t[2, I]))
print("MMA: full -> d1free seen", med(t[4, 2 * F] - t[3, 2 * F]), " issue (d1free -> committed)", med(t[5, 2 * F] - t[4, 2 * F]),
      " issuer event period", med(np.diff(t[5, 8:400])))
print("MMA committed -> R d1full seen", med(t[6, F] - t[5, 2 * F]), " R ld", med(t[7, F] - t[6, F]))
print("R ld-done -> next fill's P slot-free (other warp, not traced); R ld-done(f-1) -> P slot-free(f)", med(t[1, I] - t[7, F - 1]))
