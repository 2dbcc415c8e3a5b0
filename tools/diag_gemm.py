#!/usr/bin/env python
"""cta_group::1 vs cta_group::2 GEMM kernels vs torch.matmul (cuBLAS) on large shapes; diag bits skip stages for timing
experiments (1 MMAs, 2 stores, 4 B loads, 8 A loads; 0x10000 force the CTA-pair kernel, 0x20000 forbid it)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pairwise_sample_optimization_b200 import gemm
from tools.bench_lora_gemm import timed

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
print(f"{'M':>6} {'K':>5} {'N':>5} | {'1-SM us':>8} {'TF/s':>6} | {'2-SM us':>8} {'TF/s':>6} | {'auto us':>8} | {'cuBLAS us':>9} {'TF/s':>6}")
for (M, K, N) in [(2048, 1280, 1280), (4096, 1280, 1280), (8192, 640, 640), (8192, 1280, 1280), (8192, 5120, 1280), (32768, 640, 640),
                  (18944, 1280, 1280), (18944, 1280, 256), (37888, 1280, 1280), (16384, 2560, 2560), (8192, 8192, 8192)]:
    x, w = rn(M, K), rn(N, K)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * M * K * N
    t1 = timed(lambda: gemm.lora_gemm(x, w, out=out, diag=0x20000))
    t2 = timed(lambda: gemm.lora_gemm(x, w, out=out, diag=0x10000))
    ta = timed(lambda: gemm.lora_gemm(x, w, out=out))
    tt = timed(lambda: torch.matmul(x, w.t(), out=out))
    print(f"{M:6d} {K:5d} {N:5d} | {t1:8.1f} {fl / t1 / 1e6:6.0f} | {t2:8.1f} {fl / t2 / 1e6:6.0f} | {ta:8.1f} | {tt:9.1f} {fl / tt / 1e6:6.0f}", flush=True)
