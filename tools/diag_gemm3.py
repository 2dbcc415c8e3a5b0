"""Stage isolation of the CTA-pair GEMM kernel at the in-step shapes (diag bits: 1 no MMAs, 2 no stores, 12 no operand loads)
against cuBLAS, plus a tile-width sweep.  CUDA events around graph replays."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pairwise_sample_optimization_b200 import gemm
from tools.bench_lora_gemm import timed
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
for (M, K, N) in [(8192, 1280, 1280), (8192, 1280, 3840), (18944, 1280, 1280), (2048, 1280, 1280)]:
    x, w = rn(M, K), rn(N, K)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * M * K * N
    line = f"M={M} K={K} N={N}:"
    for diag, name in ((0, "full"), (2, "noStore"), (1, "noMMA"), (12, "noLoads"), (13, "handshake only"), (15, "nothing")):
        us = timed(lambda: gemm.lora_gemm(x, w, out=out, diag=diag | 0x10000), per_graph=4)
        line += f"  {name} {us:.1f}us"
    tt = timed(lambda: torch.matmul(x, w.t(), out=out), per_graph=4)
    print(line + f"  | cuBLAS {tt:.1f}us ({fl / tt / 1e6:.0f} TF/s)", flush=True)
    line = "   bn sweep:"
    for bn in (128, 160, 192, 224, 256):
        us = timed(lambda: gemm.lora_gemm(x, w, out=out, diag=0x10000, tune_bn=bn), per_graph=4)
        line += f"  {bn}: {us:.1f}us ({fl / us / 1e6:.0f})"
    print(line, flush=True)
