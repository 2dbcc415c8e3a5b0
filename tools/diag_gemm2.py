import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pairwise_sample_optimization_b200 import gemm
from tools.bench_lora_gemm import timed
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
e = torch.empty(1, device=dev)
print(f"empty torch kernel in graph: {timed(lambda: e.add_(1.0)):.2f} us")
for (M, K, N) in [(18944, 64, 32), (18944, 64, 256), (18944, 1024, 256), (18944, 4096, 256), (128, 64, 256), (128, 4096, 256), (2 * 18944, 64, 256), (4 * 18944, 64, 256)]:
    x, w = rn(M, K), rn(N, K)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    line = f"M={M} K={K} N={N}:"
    for diag, name in ((0, "full"), (2, "noStore"), (15, "noloads")):
        us = timed(lambda: gemm.lora_gemm(x, w, out=out, diag=diag), per_graph=4)
        line += f" {name} {us:.1f}us"
    print(line, flush=True)
