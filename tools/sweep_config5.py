#!/usr/bin/env python
"""BASELINE config 5: PSO loss/grad + LoRA GEMM microbench sweep on one GPU (batch 1-256 pairs, latents 64^2-128^2,
rank 4-128).  Writes a markdown table (stdout) and optionally JSON.  Device times: CUDA events around CUDA-graph replays."""
import argparse, json, os, sys, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import pairwise_sample_optimization_b200 as pso
from pairwise_sample_optimization_b200 import gemm

ap = argparse.ArgumentParser(); ap.add_argument("--json", default=""); args = ap.parse_args()
dev = torch.device("cuda", 0)
peaks = bench.measured_peaks()
g = torch.Generator(device=dev).manual_seed(0)
betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
sched = types.SimpleNamespace(alphas_cumprod=torch.cumprod(1.0 - betas, dim=0).to(dev))
rows = {"loss": [], "gemm": []}
print("### fused loss+grad kernel (bf16 storage, DMD2 schedule)\n")
print("| pairs | latent | time (us) | GB/s (10 N bytes per pair) | of measured HBM |\n|---|---|---|---|---|")
for hw in (64, 128):
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 256) + ((1024,) if hw == 64 else ()):
        mk = lambda: torch.randn(B, 4, hw, hw, device=dev, generator=g).bfloat16()
        x0, x1, r0, r1, n0, n1 = mk(), mk(), mk(), mk(), mk(), mk()
        p0 = (r0.float() + 0.02 * torch.randn_like(r0, dtype=torch.float32)).bfloat16()
        p1 = (r1.float() + 0.02 * torch.randn_like(r1, dtype=torch.float32)).bfloat16()
        ts = torch.tensor([999, 749, 499], device=dev)[torch.randint(0, 3, (B,), device=dev)]
        h = torch.tensor([[-1.0, 1.0]], device=dev).repeat(B, 1)
        def call():
            with torch.no_grad():
                pso.pso_pair_loss(p0, p1, r0, r1, x0, x1, n0, n1, ts, ts, h, scheduler=sched, kind="dmd", step_ratio=250)
        us = bench.graph_timed(call, 20, per_graph=4) * 1e3
        alg = 10 * 4 * hw * hw * 2 * B
        gbs = alg / us / 1e3
        rows["loss"].append({"pairs": B, "hw": hw, "us": us, "gbs": gbs})
        print(f"| {B} | 4x{hw}x{hw} | {us:.1f} | {gbs:.0f} | {100 * gbs / peaks['hbm']:.1f} % |", flush=True)
print("\n### fused base+LoRA projection, forward (t = x A^T ; y = x W^T + b + t B^T), bf16\n")
print("| M | K | N | r | 2 launches (us) | TF/s | of burst peak | torch 3 GEMMs + scale + add (us) |\n|---|---|---|---|---|---|---|---|")
rn = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).bfloat16()
for (M, K, N) in [(2048, 1280, 1280), (8192, 640, 640), (8192, 1280, 1280), (32768, 640, 640), (18944, 1280, 1280)]:
    for r in (4, 8, 16, 32, 64, 128):
        x, w, b = rn(M, K), rn(N, K, sc=K ** -0.5), rn(N)
        A, Bm = rn(r, K, sc=1 / r), rn(N, max(r, 8), sc=0.05)[:, :r]
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        def fused():
            T, _ = gemm.lora_gemm(x, A, pdl=1)
            gemm.lora_gemm(x, w, T, Bm, bias=b, out=out, pdl=2)
        def ref():
            return torch.nn.functional.linear(x, w, b) + torch.nn.functional.linear(torch.nn.functional.linear(x, A), Bm) * 1.0
        us, ut = bench.graph_timed(fused, 15) * 1e3, bench.graph_timed(ref, 15) * 1e3
        fl = 2.0 * M * K * N + 2.0 * M * r * (K + N)
        rows["gemm"].append({"M": M, "K": K, "N": N, "r": r, "us": us, "tflops": fl / us / 1e6, "torch_us": ut})
        print(f"| {M} | {K} | {N} | {r} | {us:.1f} | {fl / us / 1e6:.0f} | {100 * fl / us / 1e6 / peaks['tf_burst']:.1f} % | {ut:.1f} |", flush=True)
if args.json:
    json.dump(rows, open(args.json, "w"), indent=1)
