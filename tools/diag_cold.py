#!/usr/bin/env python
"""Why do the skinny LoRA launches take 24 us inside the step and 13 us in a tight loop?  Times t = s x A^T (M = 8192, K = 1280,
r = 64) between its own events inside a captured graph with (a) nothing in between, (b) the operands evicted from L2 by a
512 MB fill before every launch, (c) other kernels (LayerNorm, matmul, softmax: different code) in between, (d) both."""
import ctypes as C
import os
import statistics
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pairwise_sample_optimization_b200 import _lib  # noqa: E402
from tools.diag_lora_parts import gemm  # noqa: E402

dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=gen) * sc).bfloat16()
M, K, N, r = 8192, 1280, 1280, 64
x, w, A = rn(M, K), rn(N, K, sc=K ** -0.5), rn(r, K, sc=1 / r)
t, tt, y = rn(M, r), rn(r, M), torch.empty(M, N, device=dev, dtype=torch.bfloat16)
skinny = gemm(dev, a1=x, lda1=K, b1=A, ldb1=K, M=M, N=r, K1=K, d=t, ldd=r, dt=tt, lddt=M)
main = gemm(dev, a1=x, lda1=K, b1=w, ldb1=K, M=M, N=N, K1=K, d=y, ldd=N)
big = torch.empty(256 << 20, dtype=torch.bfloat16, device=dev)
ln_in, mm_a, mm_b = rn(4096, 1280), rn(2048, 2048), rn(2048, 2048)


def other():
    F.layer_norm(ln_in, (1280,))
    torch.softmax(mm_a.float(), -1)
    mm_a @ mm_b


def run(name, fn, between):
    evs = []
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):
            between(); fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(10):
            between()
            a, b = torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True)
            a.record(); fn(); b.record()
            evs.append((a, b))
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    print(f"{statistics.median(a.elapsed_time(b) for a, b in evs) * 1e3:7.1f} us  {name}", flush=True)


small = torch.zeros(1024, device=dev)
for kname, fn in (("t = s x A^T (64 CTAs)", skinny),):
    run(kname + ": one tiny elementwise kernel in between", fn, lambda: small.add_(1.0))
    run(kname + ": LayerNorm only in between", fn, lambda: F.layer_norm(ln_in, (1280,)))
    run(kname + ": cuBLAS matmul only in between", fn, lambda: mm_a @ mm_b)
    run(kname + ": softmax(fp32) only in between", fn, lambda: torch.softmax(mm_a.float(), -1))
    run(kname + ": the OTHER gemm kernel of ours (pair kernel) in between", fn, main)
for kname, fn in (("t = s x A^T (64 CTAs)", skinny), ("y = x W^T (pair kernel)", main)):
    run(kname + ": back to back", fn, lambda: None)
    run(kname + ": operands evicted from L2 (512 MB fill) before each launch", fn, lambda: big.fill_(1.0))
    run(kname + ": other kernels in between (LayerNorm, softmax, cuBLAS)", fn, other)
    run(kname + ": both", fn, lambda: (big.fill_(1.0), other()))
