"""A few launches of the CTA-pair GEMM kernel (for ncu): python tools/once_gemm2.py [M K N]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pairwise_sample_optimization_b200 import gemm
M, K, N = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (18944, 1280, 1280)
x = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    gemm.lora_gemm(x, w, out=out, diag=0x10000)
torch.cuda.synchronize(); print("ok")
