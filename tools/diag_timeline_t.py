"""Timeline of the rank-r side product t = x A^T as a CTA-pair launch of its own: row-major output only / with the transposed copy."""
import ctypes as C
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pairwise_sample_optimization_b200 import _lib, gemm
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
M, K, r = 8192, 1280, 64
x, A = rn(M, K), rn(r, K)
t, tt = torch.empty(M, r, device=dev, dtype=torch.bfloat16), torch.empty(r, M, device=dev, dtype=torch.bfloat16)
names = ["entry", "prologue", "1st operands", "1st acc", "1st drained", "last acc", "last drained", "exit"]
for label, kw in (("d only", dict(out=t)), ("d + dt", dict(out=t, out_t=tt, want_out_t=True)), ("dt only", dict(out_t=tt, want_out=False, want_out_t=True))):
    for extra, en in ((0, ""), (2, " (no stores)")):
        for _ in range(3):
            gemm.lora_gemm(x, A, diag=0x10000 | extra, **kw)
        torch.cuda.synchronize()
        gemm.lora_gemm(x, A, diag=0x50000 | extra, **kw)
        torch.cuda.synchronize()
        buf = (C.c_ulonglong * 8192)()
        assert _lib.lib().psob200_lora_gemm_timeline(buf, 8192) == 8192
        tl = torch.tensor(list(buf), dtype=torch.int64).view(512, 8, 2)[:64, :, 0].double()
        rel = (tl - tl[:, 0:1]) / 1e3
        print(f"{label}{en}: " + " ".join(f"{names[s]} {float(rel[:, s].median()):.1f}" for s in range(1, 8)), flush=True)
