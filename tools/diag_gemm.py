#!/usr/bin/env python
"""Timing experiments for the tcgen05 GEMM: which stage bounds a tile (loads, MMAs, epilogue)?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pairwise_sample_optimization_b200 import gemm
from tools.bench_lora_gemm import timed

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
for (M, K, N) in [(128, 81920, 256), (8192, 1280, 1280), (8192, 5120, 1280), (32768, 640, 640), (18944, 1280, 1280)]:
    x, w = rn(M, K), rn(N, K)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    nkb = (K + 63) // 64
    tiles_per_cta = -(-((M + 127) // 128) * ((N + 255) // 256) // 148)
    for stages in (0,):
        line = f"M={M} K={K} N={N} stages={stages}:"
        for diag, name in ((0, "full"), (1, "noMMA"), (2, "noStore"), (7, "Aonly"), (15, "noloads"), (14, "MMAonly")):
            us = timed(lambda: gemm.lora_gemm(x, w, out=out, tune_bn=256, diag=diag | (stages << 8)), per_graph=2)
            line += f" {name} {us:.1f}us ({us * 1e3 * 1.965 / (nkb * tiles_per_cta):.0f} cyc/kblock)"
        print(line, flush=True)
    us = timed(lambda: gemm.lora_gemm(x, w, out=out)); ut = timed(lambda: torch.matmul(x, w.t(), out=out))
    print(f"   ours {us:.1f}us = {2.0 * M * K * N / us / 1e6:.0f} TF/s | torch.matmul {ut:.1f}us = {2.0 * M * K * N / ut / 1e6:.0f} TF/s", flush=True)
