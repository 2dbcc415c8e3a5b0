"""Per-CTA timeline of a CTA-pair GEMM launch (diag bit 0x40000): where the fixed cost of a launch sits -- prologue, first
operands, first / last accumulator, drain, exit -- warm (same launch back to back) and cold (other kernels in between).
Diagnostics: the stamped launch itself is a few hundred ns slower than an unstamped one."""
import ctypes as C
import os, statistics, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pairwise_sample_optimization_b200 import _lib, gemm
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
ln_in, mm_a, mm_b = rn(8192, 1280), rn(2048, 2048), rn(2048, 2048)
NAMES = ["entry", "prologue done", "first operands landed", "first accumulator complete", "first tile drained",
         "last accumulator complete", "last tile drained", "exit"]


def other():
    F.layer_norm(ln_in, (1280,))
    torch.softmax(mm_a.float(), -1)
    mm_a @ mm_b


def timeline(fn, between, label):
    for _ in range(3):
        between(); fn(0)
    between()
    torch.cuda.synchronize()
    fn(0x40000)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 8192)()
    n = _lib.lib().psob200_lora_gemm_timeline(buf, 8192)
    assert n == 8192, n
    t = torch.tensor(list(buf), dtype=torch.int64).view(512, 8, 2)
    live = t[:, 0, 0] > 0
    ns = t[live][:, :, 0].double()
    t0 = ns[:, 0].min()
    print(f"  {label}: {int(live.sum())} CTAs; launch skew (entry max - min) {float(ns[:, 0].max() - t0) / 1e3:.2f} us; "
          f"last exit - first entry {float(ns[:, 7].max() - t0) / 1e3:.2f} us")
    for s in range(1, 8):
        v = ns[:, s]
        ok = v > 0
        if not bool(ok.any()):
            continue
        rel = (v[ok] - t0) / 1e3
        own = (v[ok] - ns[ok][:, 0]) / 1e3
        print(f"      {NAMES[s]:28s} since first entry: median {float(rel.median()):6.2f}  max {float(rel.max()):6.2f} us"
              f"   | since own entry: median {float(own.median()):6.2f} us")


import sys as _sys
SHAPES = [tuple(int(v) for v in a.split(',')) for a in _sys.argv[1:]] or [(8192, 1280, 1280), (2048, 1280, 1280)]
for (M, K, N) in SHAPES:
    x, w = rn(M, K), rn(N, K)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    print(f"M={M} K={K} N={N}")
    fn = lambda d: gemm.lora_gemm(x, w, out=out, diag=d | 0x10000)
    timeline(fn, lambda: None, "warm (back to back)")
    timeline(fn, other, "cold (LayerNorm, softmax, cuBLAS matmul in between)")
    for d, name in ((2, "no stores"), (12, "no loads"), (1, "no MMAs")):
        timeline(lambda dd: gemm.lora_gemm(x, w, out=out, diag=dd | d | 0x10000), lambda: None, f"warm, {name}")
