"""Cold-start cost of a main LoRA GEMM launch inside a step: the CTA-pair kernel and cuBLAS at the in-step shapes, each timed
between its own events inside a captured graph with (a) nothing but a tiny kernel in between, (b) other kernels (LayerNorm,
softmax, a different matmul: evicts code and tensor maps) in between, (c) 64 different weight matrices in rotation (weights come
from HBM, as in the step), (d) both."""
import os, statistics, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pairwise_sample_optimization_b200 import gemm
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
ln_in, mm_a, mm_b = rn(8192, 1280), rn(2048, 2048), rn(2048, 2048)
small = torch.zeros(1024, device=dev)


def other():
    F.layer_norm(ln_in, (1280,))
    torch.softmax(mm_a.float(), -1)
    mm_a @ mm_b


def run(fn, between, n=16):
    evs = []
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for i in range(2):
            between(); fn(i)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=side):
        for i in range(n):
            between()
            a, b = torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True)
            a.record(); fn(i); b.record()
            evs.append((a, b))
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in evs) * 1e3


for (M, K, N) in [(8192, 1280, 1280), (2048, 1280, 1280), (8192, 1280, 3840)]:
    x = rn(M, K)
    ws = [rn(N, K) for _ in range(64)]   # 64 x 3.3 MB (10 MB): more than L2 holds next to the activations
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for name, call in (("pair kernel", lambda w: gemm.lora_gemm(x, w, out=out)), ("cuBLAS     ", lambda w: torch.matmul(x, w.t(), out=out))):
        same, rot = (lambda i: call(ws[0])), (lambda i: call(ws[(7 * i) % 64]))
        r = [run(same, lambda: small.add_(1.0)), run(same, other), run(rot, lambda: small.add_(1.0)), run(rot, other)]
        print(f"M={M} K={K} N={N} {name}: warm {r[0]:.1f}us | other kernels between {r[1]:.1f} | weights in rotation {r[2]:.1f} | both {r[3]:.1f}",
              flush=True)
