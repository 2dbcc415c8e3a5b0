#!/usr/bin/env python
"""Fused GEGLU kernels alone: device time and HBM GB/s at the SDXL feed-forward shapes (CUDA events around graph replays;
working set 250-420 MB > L2).   python tools/bench_geglu.py [--once N]   (--once: a few plain launches, for ncu)"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pairwise_sample_optimization_b200 import feed_forward  # noqa: E402
from tools.bench_lora_gemm import timed  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--once", type=int, default=0)
args = ap.parse_args()
dev = torch.device("cuda", 0)
for (M, I) in ((8192, 5120), (32768, 2560)):
    proj = torch.randn(M, 2 * I, device=dev).bfloat16().requires_grad_(True)
    dout = torch.randn(M, I, device=dev).bfloat16()
    if args.once:
        for _ in range(args.once):
            feed_forward.geglu(proj).backward(dout)
        torch.cuda.synchronize()
        continue
    with torch.no_grad():
        t_f = timed(lambda: feed_forward.geglu(proj), per_graph=1)
    t_b = timed(lambda: torch.autograd.grad(feed_forward.geglu(proj), proj, dout), per_graph=1) - t_f  # fwd + bwd, minus fwd

    def stock():
        h, g = proj.chunk(2, dim=-1)
        return h * F.gelu(g)
    with torch.no_grad():
        t_sf = timed(stock, per_graph=1)
    t_sb = timed(lambda: torch.autograd.grad(stock(), proj, dout), per_graph=1) - t_sf
    bf, bb = 3 * M * I * 2, 5 * M * I * 2
    print(f"M={M} I={I}: fwd {t_f:.1f} us ({bf / t_f / 1e3:.0f} GB/s), bwd {t_b:.1f} us ({bb / t_b / 1e3:.0f} GB/s); "
          f"stock torch fwd {t_sf:.1f} us, bwd {t_sb:.1f} us", flush=True)
print("ok")
