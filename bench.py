#!/usr/bin/env python
"""bench.py -- 3DMatch-shape pairs/sec of the registration hot path on N B200s, plus the KPConv roofline.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA hot path)
    python bench.py --impl reference --steps K --warmup W     # the reference algorithm on the host CPU cores

A step is one pass of the whole hot path (pyramid -> KPConv encoder -> transformer -> superpoint matching
-> pose) over one batch of `--pairs` synthetic 3DMatch-shape pairs per GPU.  Pairs are independent, so ranks
shard them with no data-path collective; the only communication is the final all_gather of the poses.
Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "3DMatch-shape pairs/sec"
UNIT = "pairs/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=32, help="pairs per GPU per step")
    ap.add_argument("--points", type=int, default=20000, help="nominal points per fragment")
    ap.add_argument("--arch", default="4stage", choices=["3stage", "4stage"],
                    help="4stage = the configuration BASELINE.json names; 3stage = the shipped yaml")
    ap.add_argument("--no-alt", action="store_true", help="skip the short run of the other architecture")
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-pairs", type=int, default=8, help="pairs in the bounded CPU sample (about 5 s of host work)")
    return ap.parse_args()


def make_cfg(arch):
    import superpoints_registration_b200 as spr
    return spr.threedmatch_config() if arch == "3stage" else spr.threedmatch_4stage_config()


def workload_name(args):
    stages = ("3-stage (shipped conf/qk_regtr_full_3dmatch.yaml:56-63)" if args.arch == "3stage"
              else "4-stage (BASELINE.json configs[2]; conf/qk_regtr_full_3dmatch.yaml:64-74)")
    return (f"3DMatch-shape fragments (~{args.points // 1000}k pts, voxel 0.025 m, {stages} KPConv), "
            f"full forward + pose, Sinkhorn x3")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------------
# the reference algorithm on the host CPU (oracle) -- cpu_baseline leg and --impl reference
# --------------------------------------------------------------------------------------------------------

def cpu_forward_timing(args, n_pairs, repeats=1):
    """Time the CPU restatement of the reference path on `n_pairs` pairs of the bench workload."""
    import torch

    import oracle
    from oracle import pipeline
    import superpoints_registration_b200 as spr
    from superpoints_registration_b200 import synthetic

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = make_cfg(args.arch)
    torch.manual_seed(0)
    np.random.seed(0)
    model = spr.RegTR(cfg)  # host-side module only used as a container of reference-named random-init weights
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    data = synthetic.make_batch("3dmatch", n_pairs, seed=args.seed, n_points=args.points)
    backend = "reference" if oracle.have_ref() else "port"
    best, stages = None, {}
    for _ in range(repeats):
        t = {}
        t0 = time.perf_counter()
        pipeline.forward(sd, cfg, data["src_xyz"], data["tgt_xyz"], backend=backend, timings=t)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, stages = dt, t
    kind = "port"
    sample = (f"{n_pairs} pair(s) of the bench workload through oracle/pipeline.py:forward (torch-CPU fp32 restatement "
              f"of the reference network, {cores} torch threads; preprocessing by "
              f"{'the reference C++ core (oracle/_ref), single-threaded as in the reference' if backend == 'reference' else 'the C restatement'}); "
              f"stage seconds: " + ", ".join(f"{k}={v:.2f}" for k, v in stages.items()))
    return n_pairs / best, best, cores, kind, sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # warm-up once (imports, thread pools), then K bounded steps
    n = max(1, args.cpu_pairs)
    for _ in range(min(args.warmup, 1)):
        cpu_forward_timing(args, n)
    times = []
    info = None
    for _ in range(max(1, args.steps)):
        v, dt, cores, kind, sample = cpu_forward_timing(args, n)
        times.append(dt)
        info = (cores, kind, sample)
    total = sum(times)
    value = n * len(times) / total
    cores, kind, sample = info
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "pairs_per_step": n, "host": "cpu"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------------------------

class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for p in self.samples:
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for name, val in zip(names, p[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------

def kpconv_algorithmic_bytes(nq, H, cin, cout, K=15):
    """gather-expanded bytes of one KPConv layer, 4-byte indices."""
    return nq * H * (4 * cin + 12 + 4) + nq * (12 + 4 * cout) + 4 * K * cin * cout + 12 * K


def run_ours(args):
    import torch
    import torch.distributed as dist

    import superpoints_registration_b200 as spr
    from superpoints_registration_b200 import _lib, ops, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun exactly as the driver would
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29577", os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: whatever NCCL writes while the communicator comes up (its version
        # banner at NCCL_DEBUG=WARN/VERSION/INFO) is sent to stderr by pointing fd 1 at fd 2 for that moment
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # later NCCL log lines (NCCL_DEBUG=INFO) too
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)  # the first collective creates the communicator
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    cfg = make_cfg(args.arch)
    torch.manual_seed(0)
    np.random.seed(0)
    model = spr.RegTR(cfg).to(dev).eval()
    model.return_attn = False  # outputs['attn'] (N x M per pair) is training/analysis output; pose path needs corr only
    B = args.pairs
    data = synthetic.make_batch("3dmatch", B, seed=args.seed * 100 + rank, n_points=args.points)
    host_src = [torch.from_numpy(c).pin_memory() for c in data["src_xyz"]]
    host_tgt = [torch.from_numpy(c).pin_memory() for c in data["tgt_xyz"]]
    dev_batch = {"src_xyz": [c.to(dev) for c in host_src], "tgt_xyz": [c.to(dev) for c in host_tgt]}
    h2d_bytes = sum(c.numel() * 4 for c in host_src + host_tgt)
    # end-to-end staging: all fragments of the step in ONE pinned buffer -> one host-to-device copy per step
    cloud_lens = [c.shape[0] for c in host_src + host_tgt]
    host_all = torch.cat(host_src + host_tgt, dim=0).pin_memory()
    gathered = torch.empty((world * B, 3, 4), device=dev) if world > 1 else None
    host_pose = torch.empty((B, 3, 4)).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step(batch):
        out = model(dict(batch))
        pose = out["pose"]
        if world > 1:
            dist.all_gather_into_tensor(gathered, pose.contiguous())
        return pose

    # ---- KPConv instrumentation (CUDA events on the launching stream) ----
    records = []
    raw_kpconv = ops.kpconv_forward
    recording = {"on": False}

    def timed_kpconv(q_pts, s_pts, neighb_inds, x, weights, kernel_points, extent, mode=0):
        if not recording["on"]:
            return raw_kpconv(q_pts, s_pts, neighb_inds, x, weights, kernel_points, extent, mode)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = raw_kpconv(q_pts, s_pts, neighb_inds, x, weights, kernel_points, extent, mode)
        e1.record()
        records.append((e0, e1, q_pts.shape[0], neighb_inds.shape[1], weights.shape[1], weights.shape[2]))
        return y

    ops.kpconv_forward = timed_kpconv

    # the encoder blocks feed the tensor-core KPConv with operands written by the preceding normalisation kernel
    raw_prepared = ops.kpconv_forward_prepared

    def timed_prepared(q_pts, neighb_inds, feats, weights, kernel_points, extent, **kw):
        if not recording["on"]:
            return raw_prepared(q_pts, neighb_inds, feats, weights, kernel_points, extent, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = raw_prepared(q_pts, neighb_inds, feats, weights, kernel_points, extent, **kw)
        e1.record()
        records.append((e0, e1, q_pts.shape[0], neighb_inds.shape[1], weights.shape[1], weights.shape[2]))
        return y

    ops.kpconv_forward_prepared = timed_prepared

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(dev_batch)
    barrier()

    # ---- timed region 1: inputs resident in HBM ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    recording["on"] = True
    evs = []
    barrier()
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the timed interval)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(dev_batch)
        e1.record()
        evs.append((e0, e1))
    barrier()
    recording["on"] = False
    launches = _lib.launch_count() - launches0
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 2: end to end through the public API with HOST buffers ----
    e2e_evs = []
    barrier()
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        clouds = torch.split(host_all.to(dev, non_blocking=True), cloud_lens)
        batch = {"src_xyz": list(clouds[:B]), "tgt_xyz": list(clouds[B:])}
        pose = step(batch)
        host_pose.copy_(pose, non_blocking=True)
        e1.record()
        e2e_evs.append((e0, e1))
    barrier()
    e2e_ms = sum(a.elapsed_time(b) for a, b in e2e_evs)

    t = torch.tensor([total_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        pairs_total = world * B * args.steps
        value = pairs_total / (total_ms * 1e-3)
        e2e_value = pairs_total / (e2e_ms * 1e-3)
        peak, peak_src = peaks()
        k_bytes = sum(kpconv_algorithmic_bytes(nq, H, ci, co) for (_, _, nq, H, ci, co) in records)
        k_ms = sum(a.elapsed_time(b) for (a, b, *_r) in records)
        k_flops = sum(2.0 * nq * 15 * ci * co + 2.0 * nq * 15 * H * ci for (_, _, nq, H, ci, co) in records)
        achieved = k_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "kpconv_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        per_layer = {}
        for (a, b, nq, H, ci, co) in records:
            key = f"Nq={nq},H={H},C={ci}->{co}"
            d = per_layer.setdefault(key, {"ms": 0.0, "bytes": 0, "n": 0})
            d["ms"] += a.elapsed_time(b)
            d["bytes"] += kpconv_algorithmic_bytes(nq, H, ci, co)
            d["n"] += 1
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "pairs_per_gpu_per_step": B, "parallelism": f"pairs sharded x{world}",
                       "points_per_step": int(sum(c.shape[0] for c in host_src + host_tgt)),
                       "l2": "flushed between timed iterations (256 MiB write)", "seed": args.seed},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": k_bytes / max(len(records), 1),
                         "kernel_ms_per_launch": k_ms / max(len(records), 1), "kernel": "k_kpconv_tc (operands pre-split by the preceding norm kernel) + k_kpconv_cin1 stem: all KPConv layers of the step",
                         "peak_source": peak_src, "algorithmic_bytes_per_step": k_bytes / max(args.steps, 1),
                         "kpconv_ms_per_step": k_ms / max(args.steps, 1),
                         "kpconv_tflops_fp32": k_flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0,
                         "per_layer_gbs": {k: v["bytes"] / (v["ms"] * 1e-3) / 1e9 for k, v in per_layer.items()}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": int(host_pose.numel() * 4), "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_alt:
            # the other architecture of the same yaml, short run, same timing rules (reported beside, not the headline)
            other = "3stage" if args.arch == "4stage" else "4stage"
            torch.manual_seed(0)
            alt_model = spr.RegTR(make_cfg(other)).to(dev).eval()
            alt_model.return_attn = False
            for _ in range(3):
                alt_model(dict(dev_batch))
            torch.cuda.synchronize()
            alt_ms = 0.0
            alt_steps = max(3, args.steps // 2)
            for _ in range(alt_steps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                alt_model(dict(dev_batch))
                e1.record()
                torch.cuda.synchronize()
                alt_ms += e0.elapsed_time(e1)
            line["alt"] = {"workload": workload_name(argparse.Namespace(**{**vars(args), "arch": other})),
                           "value": B * alt_steps / (alt_ms * 1e-3), "unit": UNIT, "ms_per_step": alt_ms / alt_steps,
                           "steps": alt_steps}
            del alt_model
        if world == 1 and not args.no_cpu_baseline:
            v, dt, cores, kind, sample = cpu_forward_timing(args, max(1, args.cpu_pairs), repeats=2)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
