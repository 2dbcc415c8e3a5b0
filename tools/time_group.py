"""Forward / input-gradient time of a stacked LoRA projection group (fused launches) and of the frozen pass, CUDA events around
graph replays; for A/B runs under PSOB200_GEMM_DIAG / PSOB200_FUSED_BN (read once per process).

    python tools/time_group.py "1,8192,1280,1280,64;3,8192,1280,1280,64"
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pairwise_sample_optimization_b200 import lora
from tools.bench_lora_gemm import timed

shapes = sys.argv[1] if len(sys.argv) > 1 else "1,8192,1280,1280,64;3,8192,1280,1280,64;1,32768,640,640,64;1,2048,1280,1280,8;3,2048,1280,1280,8"
tag = f"diag={os.environ.get('PSOB200_GEMM_DIAG', '0')} fused_bn={os.environ.get('PSOB200_FUSED_BN', '-')}"
lora.set_wgrad_stream(False)
out = []
for shp in shapes.split(";"):
    G, M, K, N, r = (int(v) for v in shp.split(","))
    torch.manual_seed(0)
    layers = []
    for _ in range(G):
        lay = lora.LoRALinear(torch.nn.Linear(K, N, bias=False, device="cuda", dtype=torch.bfloat16), r, r)
        with torch.no_grad():
            lay.lora_B["default"].weight.normal_(std=0.02)
        layers.append(lay)
    group = lora.LoRAProjectionGroup(layers) if G > 1 else None
    call = (lambda x: group(x)) if G > 1 else (lambda x: (layers[0](x),))
    x = torch.randn(M, K, device="cuda").bfloat16().requires_grad_(True)
    dys = [torch.randn(M, N, device="cuda").bfloat16() for _ in range(G)]

    def fwd():
        with torch.no_grad():
            call(x)

    def fwd_bwd():
        ys = call(x)
        torch.autograd.backward(list(ys), dys)
        x.grad = None

    f = timed(fwd)
    fb = timed(fwd_bwd)
    for lay in layers: lay.enable_adapters(False)
    fr = timed(fwd)
    for lay in layers: lay.enable_adapters(True)
    out.append(f"G={G} M={M} K={K} N={N} r={r}: fwd {f:.1f} fwd+bwd {fb:.1f} frozen {fr:.1f}")
print(tag + " | " + " | ".join(out), flush=True)
