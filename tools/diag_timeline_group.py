"""Timeline of the FUSED projection launches (t tiles + y tiles; u tiles + dx tiles) of a stacked group against the frozen pass:
run with PSOB200_GEMM_DIAG=0x40000 so that every CTA-pair launch records its per-CTA stamps (tools/diag_timeline.py has the
slot meanings).  Prints, per launch kind, when each CTA finished and how the finish times spread."""
import ctypes as C
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pairwise_sample_optimization_b200 import _lib, lora

assert int(os.environ.get("PSOB200_GEMM_DIAG", "0"), 0) & 0x40000, "run with PSOB200_GEMM_DIAG=0x40000"


def read(label):
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 8192)()
    assert _lib.lib().psob200_lora_gemm_timeline(buf, 8192) == 8192
    t = torch.tensor(list(buf), dtype=torch.int64).view(512, 8, 2)[:148, :, 0].double()
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    names = ["entry", "prologue", "1st operands", "1st acc", "1st drained", "last acc", "last drained", "exit"]
    line = f"  {label}: total {float(rel[:, 7].max()):.2f} us |"
    for s in (1, 2, 3, 4, 5, 6, 7):
        v = rel[:, s][t[:, s] > 0]
        line += f" {names[s]} {float(v.median()):.1f}/{float(v.max()):.1f}"
    print(line)
    if os.environ.get("TIMELINE_PER_CTA"):
        for c in list(range(0, 12, 2)) + list(range(60, 72, 2)) + list(range(84, 92, 2)) + list(range(140, 148, 2)):
            print(f"      CTA {c:3d}: " + " ".join(f"{names[s_]} {float(rel[c, s_]):.1f}" for s_ in range(1, 8)))
    ex = rel[:, 7].sort().values
    print("      exit times of the CTAs (us), every 8th: " + " ".join(f"{float(v):.1f}" for v in ex[::8]))


for (G, M, K, N, r) in [(1, 8192, 1280, 1280, 64), (3, 8192, 1280, 1280, 64)]:
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
    print(f"G={G} M={M} K={K} N={N} r={r}")
    lora.set_wgrad_stream(False)
    for it in range(3):
        ys = call(x)
        if it == 2: read("forward  (t tiles + y tiles)")
        torch.autograd.backward(list(ys), dys)   # input gradient, then the weight-gradient launch (1-SM kernel: no stamps)
        if it == 2: read("backward (u tiles + dx tiles)")
    with torch.no_grad():
        for lay in layers: lay.enable_adapters(False)
        ys = call(x)
        read("frozen forward")
        for lay in layers: lay.enable_adapters(True)
