"""A few launches of the stacked LoRA projection group (for ncu): the fused forward (t tiles + y tiles, one launch), the fused
input-gradient launch (u tiles + dx tiles) and the weight-gradient launch, at the in-step shape of dmd128:

    python tools/once_group.py [G M K N r]        # default 3 8192 1280 1280 64
    ncu --set full --clock-control none --import-source on -k regex:lora_gemm -s 3 -c 3 -o gpurun_out/<name> python tools/once_group.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pairwise_sample_optimization_b200 import lora  # noqa: E402

G, M, K, N, r = (int(v) for v in sys.argv[1:6]) if len(sys.argv) >= 6 else (3, 8192, 1280, 1280, 64)
torch.manual_seed(0)
layers = []
for _ in range(G):
    lay = lora.LoRALinear(torch.nn.Linear(K, N, bias=False, device="cuda", dtype=torch.bfloat16), r, r)
    with torch.no_grad():
        lay.lora_B["default"].weight.normal_(std=0.02)
    layers.append(lay)
group = lora.LoRAProjectionGroup(layers)
x = torch.randn(M, K, device="cuda").bfloat16().requires_grad_(True)
dys = [torch.randn(M, N, device="cuda").bfloat16() for _ in range(G)]
for _ in range(2):  # launches 0-2 warm up, 3-5 are the ones to capture (forward, input gradient, weight gradients)
    ys = group(x)
    ys = ys if isinstance(ys, tuple) else (ys,)
    torch.autograd.backward(list(ys), dys)
torch.cuda.synchronize()
print("ok")
