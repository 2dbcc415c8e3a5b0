"""A few forward launches of a cross-attention k / v bank (for ncu): 120 stacked projections of the prompt embeddings, the in-step
shape of dmd128 (M = 8 x 77 = 616, K = 2048, N = 1280, r = 64): python tools/once_bank.py [G M K N r]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pairwise_sample_optimization_b200 import lora  # noqa: E402

G, M, K, N, r = (int(v) for v in sys.argv[1:6]) if len(sys.argv) >= 6 else (120, 616, 2048, 1280, 64)
torch.manual_seed(0)
layers = []
for _ in range(G):
    lay = lora.LoRALinear(torch.nn.Linear(K, N, bias=False, device="cuda", dtype=torch.bfloat16), r, r)
    with torch.no_grad():
        lay.lora_B["default"].weight.normal_(std=0.02)
    layers.append(lay)
group = lora.LoRAProjectionGroup(layers)
x = torch.randn(M, K, device="cuda").bfloat16()
for _ in range(3):  # launches 0-2 with adapters (t tiles + y tiles), 3-5 the frozen pass
    group(x)
for lay in layers:
    lay.enable_adapters(False)
with torch.no_grad():
    for _ in range(3):
        group(x)
torch.cuda.synchronize()
print("ok")
