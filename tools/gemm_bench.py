"""Micro-benchmark of the tcgen05 GEMM (ops.gemm_tc) and the attention kernel on the cross-encoder's shapes.
    python tools/gemm_bench.py [--tokens 20000] [--reps 5] [--only gemm|attn]"""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from superpoints_registration_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--tokens", type=int, default=20032)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--only", default="")
args = ap.parse_args()
dev = "cuda:0"
T = args.tokens
torch.manual_seed(0)


def timeit(fn):
    best = 1e9
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


if args.only in ("", "gemm"):
    x = torch.randn(T, 256, device=dev)
    img256 = ops.gemm_prepare_input(x)
    img1024 = ops.gemm_prepare_input(torch.randn(T, 1024, device=dev))
    for name, N, K, mode, img in (("qkv   256->768  planes", 768, 256, ops.OUT_PLANES, img256),
                                  ("oproj 256->256  f32+res", 256, 256, ops.OUT_F32, img256),
                                  ("ffn1  256->1024 image", 1024, 256, ops.OUT_AIMG, img256),
                                  ("ffn2  1024->256 f32+res", 256, 1024, ops.OUT_F32, img1024)):
        w = torch.randn(N, K, device=dev) / math.sqrt(K)
        b = torch.randn(N, device=dev)
        wi = ops.weight_image(w)
        res = torch.randn(T, N, device=dev) if mode == ops.OUT_F32 else None
        out = torch.empty((T, N), device=dev) if mode == ops.OUT_F32 else None
        ms = timeit(lambda: ops.gemm_tc(img, wi, b, T, mode, residual=res, out=out))
        print(f"gemm {name}: {ms * 1e3:8.1f} us   {2.0 * T * N * K / ms / 1e9:8.1f} TFLOP/s (fp32-equivalent)")
if args.only in ("", "attn"):
    n_clouds, d, nh = 16, 256, 8
    lens = [T // n_clouds] * n_clouds
    offs = [i * lens[0] for i in range(n_clouds)]
    qkv = torch.randn(sum(lens), 768, device=dev)
    hi, lo = ops.split_f16(qkv, n_scaled=256, scale=math.log2(math.e) / math.sqrt(32))
    tiles = ops.attention_tiles(offs, lens, offs, lens, dev)
    img = ops.gemm_a_image(sum(lens), 256, dev)
    ms = timeit(lambda: ops.attention_varlen(hi, lo, tiles, nh, 0, d, 2 * d, d, out_image=img, image_scale=16.0))
    macs = sum(n * n for n in lens) * d * 2
    print(f"attention {n_clouds} x {lens[0]} tokens: {ms * 1e3:8.1f} us   {2.0 * macs / ms / 1e9:8.1f} TFLOP/s (fp32-equivalent)")
