"""Per-kernel device time of one forward pass (torch.profiler / CUPTI), aggregated by kernel name.
Development aid: python tools/kernel_times.py [--pairs 8]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

import superpoints_registration_b200 as spr
from superpoints_registration_b200 import synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=8)
ap.add_argument("--points", type=int, default=20000)
ap.add_argument("--top", type=int, default=32)
ap.add_argument("--arch", default="4stage", choices=["3stage", "4stage"])
ap.add_argument("--kind", default="3dmatch", choices=["3dmatch", "kitti", "modelnet"])
args = ap.parse_args()
dev = "cuda:0"
cfg = spr.threedmatch_config() if args.arch == "3stage" else spr.threedmatch_4stage_config()
gen_kw = dict(n_points=args.points)
if args.kind == "kitti":
    cfg, gen_kw = spr.kitti_config(first_subsampling_dl=0.3), dict(n_points=30000, voxel=0.3)
elif args.kind == "modelnet":
    cfg, gen_kw = spr.modelnet_config(), {}
torch.manual_seed(0); np.random.seed(0)
model = spr.RegTR(cfg).to(dev).eval()
model.return_attn = False
data = synthetic.make_batch(args.kind, args.pairs, seed=2, **gen_kw)
batch = {"src_xyz": [torch.from_numpy(c).to(dev) for c in data["src_xyz"]],
         "tgt_xyz": [torch.from_numpy(c).to(dev) for c in data["tgt_xyz"]]}
for _ in range(3):
    model(dict(batch))
torch.cuda.synchronize()
N = 3
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(N):
        model(dict(batch))
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / N, e.count / N) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total device time per forward: {tot / 1e3:.3f} ms over {sum(r[2] for r in rows):.0f} kernels")
for k, t, c in rows[:args.top]:
    print(f"{t / 1e3:8.3f} ms {100 * t / tot:5.1f}%  n={c:5.0f}  {k[:100]}")

# ---- per-launch detail of the normalisation kernels against their byte counts (HBM-bound by design) ----
if os.environ.get("SPR_IN_DETAIL"):
    from superpoints_registration_b200 import ops
    calls = []
    raw = ops.instance_norm_lrelu_ex

    def rec(x, lengths, *a, **kw):
        out = raw(x, lengths, *a, **kw)
        calls.append((x.shape[0], x.shape[1], kw.get("residual") is not None, out.get("f32") is not None,
                      out.get("image") is not None, out.get("kpconv") is not None))
        return out

    ops.instance_norm_lrelu_ex = rec
    import superpoints_registration_b200.kpconv_blocks as kb
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model(dict(batch))
        torch.cuda.synchronize()
    ev = sorted([e for e in prof.events() if e.device_type.name == "CUDA"], key=lambda e: e.time_range.start)
    part = [e.device_time for e in ev if "k_in_partial" in e.name]
    appl = [e.device_time for e in ev if "k_in_apply_ex" in e.name]
    print(f"{len(calls)} instance-norm calls, {len(part)} partial, {len(appl)} apply launches")
    for (n, c, res, f32, img, kp), tp, ta in zip(calls, part, appl):
        rd = n * c * 4 * (2 if res else 1)
        wr = n * c * 4 * (int(f32) + int(img) + int(kp)) + (n * 16 if kp else 0)
        print(f"n={n:7d} c={c:4d} res={int(res)} f32={int(f32)} img={int(img)} kp={int(kp)}  partial {tp:6.1f} us "
              f"{n * c * 4 / tp / 1e3:7.0f} GB/s   apply {ta:6.1f} us {(rd + wr) / ta / 1e3:7.0f} GB/s")

# ---- per-launch detail of the tensor-core GEMMs against their operand / result bytes ----
if os.environ.get("SPR_GEMM_DETAIL"):
    from superpoints_registration_b200 import ops
    calls = []
    raw_gemm = ops.gemm_tc

    def rec_gemm(a_img, wi, bias, T, mode=ops.OUT_F32, residual=None, **kw):
        calls.append((int(T), wi.N, wi.K, int(mode), residual is not None, kw.get("stats16") is not None))
        return raw_gemm(a_img, wi, bias, T, mode, residual=residual, **kw)

    ops.gemm_tc = rec_gemm
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model(dict(batch))
        torch.cuda.synchronize()
    ev = sorted([e for e in prof.events() if e.device_type.name == "CUDA" and "k_gemm_tc" in e.name],
                key=lambda e: e.time_range.start)
    print(f"{len(calls)} gemm calls, {len(ev)} launches")
    agg = {}
    for (T, N, K, mode, res, st), e in zip(calls, ev):
        kp = (K + 63) // 64 * 64
        rd = T * kp * 4 + N * kp * 4 + (T * N * 4 if res else 0)
        wr = T * N * 4 + (T // 16 * N * 8 if st else 0)
        key = (T, N, K, mode, res, st)
        a = agg.setdefault(key, [0, 0.0, rd + wr, 2.0 * T * N * K])
        a[0] += 1
        a[1] += e.device_time
    for key, (cnt, t, by, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        T, N, K, mode, res, st = key
        print(f"T={T:8d} N={N:5d} K={K:5d} mode={mode} res={int(res)} stats={int(st)} x{cnt:2d}  {t:8.1f} us total "
              f"{t / cnt:7.1f} us each  {by / (t / cnt) / 1e3:6.0f} GB/s  {fl / (t / cnt) / 1e6:6.1f} TFLOP/s")

# ---- every launch of one kernel, in launch order:  SPR_LAUNCHES=k_radius_query ----
if os.environ.get("SPR_LAUNCHES"):
    pat = os.environ["SPR_LAUNCHES"]
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model(dict(batch))
        torch.cuda.synchronize()
    ev = sorted([e for e in prof.events() if e.device_type.name == "CUDA" and pat in e.name],
                key=lambda e: e.time_range.start)
    print(pat, [round(e.device_time, 1) for e in ev])

# ---- idle time of the GPU between consecutive kernels of one forward:  SPR_GAPS=1 ----
if os.environ.get("SPR_GAPS"):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model(dict(batch))
        torch.cuda.synchronize()
    ev = sorted([e for e in prof.events() if e.device_type.name == "CUDA"], key=lambda e: e.time_range.start)
    gaps = []
    for a, b in zip(ev[:-1], ev[1:]):
        gaps.append((b.time_range.start - a.time_range.end, a.name[:50], b.name[:50]))
    span = ev[-1].time_range.end - ev[0].time_range.start
    busy = sum(e.device_time for e in ev)
    print(f"span {span / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms, idle {sum(g[0] for g in gaps) / 1e3:.2f} ms over {len(gaps)} gaps")
    big = sorted(gaps, key=lambda g: -g[0])[:14]
    for g in big:
        print(f"{g[0]:8.1f} us  after {g[1]}  before {g[2]}")
    small = [g[0] for g in gaps if g[0] < 10]
    print(f"{len(small)} gaps under 10 us, mean {sum(small) / max(len(small), 1):.2f} us, total {sum(small) / 1e3:.2f} ms")
