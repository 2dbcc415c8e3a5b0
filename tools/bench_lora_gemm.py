#!/usr/bin/env python
"""LoRA projection GEMM microbench on one GPU (BASELINE config 5: rank 4-128, SDXL attention shapes).

    python tools/bench_lora_gemm.py                    # table: tcgen05 path vs torch (cuBLAS 3 GEMMs + scale + add)
    python tools/bench_lora_gemm.py --once 3 --shape 8192,1280,1280 --rank 64    # a few launches, for ncu

Times are device times (CUDA events around CUDA-graph replays, operands larger than L2 rotated between replays is
NOT needed here: the weight is meant to stay L2-resident, activations stream)."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pairwise_sample_optimization_b200 import gemm, lora  # noqa: E402

# (M per sample, K, N, count in the SDXL UNet) -- SURVEY.md section 8a row a9
SDXL_64 = [(1024, 640, 640), (256, 1280, 1280), (77, 2048, 640), (77, 2048, 1280)]
SDXL_128 = [(4096, 640, 640), (1024, 1280, 1280), (77, 2048, 640), (77, 2048, 1280)]


def timed(fn, reps=20, per_graph=4):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for _ in range(per_graph):
            fn()
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        graph.replay()
        b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in evs) * 1e3 / per_graph  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--once", type=int, default=0)
    ap.add_argument("--shape", default="")
    ap.add_argument("--rank", type=int, default=64)
    ap.add_argument("--batch", type=int, default=8, help="samples batched into M (win+lose of B/2 pairs)")
    ap.add_argument("--bn", type=int, default=0)
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    dt = torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s, scale=1.0: (torch.randn(*s, device=dev, generator=g) * scale).to(dt)

    def setup(M, K, N, r):
        x, w, b = rn(M, K), rn(N, K, scale=K ** -0.5), rn(N)
        A, Bm = rn(r, K, scale=1.0 / r), rn(N, max(r, 8), scale=0.05)[:, :r]
        return x, w, b, A, Bm

    if args.once:
        M, K, N = (int(v) for v in args.shape.split(","))
        x, w, b, A, Bm = setup(M, K, N, args.rank)
        for _ in range(args.once):
            T, _ = gemm.lora_gemm(x, A, alpha=1.0)
            gemm.lora_gemm(x, w, T, Bm, bias=b, tune_bn=args.bn)
        torch.cuda.synchronize()
        print("ok")
        return

    rows = []
    print(f"{'M':>6} {'K':>5} {'N':>5} {'r':>4} | {'fused fwd us':>12} {'TF/s':>7} | {'main us':>8} {'TF/s':>7} | "
          f"{'torch us':>9} {'TF/s':>7} | {'f+b us':>8} {'TF/s':>7} | {'torch f+b':>9}")
    shapes = [(m * args.batch, k, n) for (m, k, n) in SDXL_64 + SDXL_128[:2]]
    if args.shape:
        shapes = [tuple(int(v) for v in args.shape.split(","))]
    for (M, K, N) in shapes:
        for r in (8, 64) if not args.shape else (args.rank,):
            x, w, b, A, Bm = setup(M, K, N, r)
            dy = rn(M, N)
            flops_f = 2.0 * M * K * N + 2.0 * M * r * (K + N)
            flops_b = flops_f + 2.0 * M * K * N + 2.0 * M * r * (K + N) + 2.0 * M * r * (K + N)  # fwd + dX (+U) + dA, dB

            def fused():
                T, _ = gemm.lora_gemm(x, A, alpha=1.0, pdl=1)
                return gemm.lora_gemm(x, w, T, Bm, bias=b, tune_bn=args.bn, pdl=2)

            def main_only(T=gemm.lora_gemm(x, A)[0]):
                return gemm.lora_gemm(x, w, T, Bm, bias=b, tune_bn=args.bn)

            def torch_ref():
                return torch.nn.functional.linear(x, w, b) + torch.nn.functional.linear(torch.nn.functional.linear(x, A), Bm) * 1.0

            lay = lora.LoRALinear(torch.nn.Linear(K, N, bias=True, device=dev, dtype=dt), r, r, lora_dtype=torch.float32)
            xg = x.clone().requires_grad_(True)

            def bwd():  # forward + backward through the public module (dX, dA, dB)
                lay(xg).backward(dy)

            xa = x.clone().requires_grad_(True)
            Ap, Bp = A.clone().requires_grad_(True), Bm.clone().contiguous().requires_grad_(True)

            def torch_bwd():
                yt = torch.nn.functional.linear(xa, w, b) + torch.nn.functional.linear(torch.nn.functional.linear(xa, Ap), Bp)
                yt.backward(dy)

            t_f, t_m, t_t = timed(fused), timed(main_only), timed(torch_ref)
            t_b = timed(bwd, per_graph=2)
            t_tb = timed(torch_bwd, per_graph=2)
            row = {"M": M, "K": K, "N": N, "r": r, "fused_fwd_us": t_f, "main_us": t_m, "torch_fwd_us": t_t, "bwd_us": t_b,
                   "torch_bwd_us": t_tb, "fused_fwd_tflops": flops_f / t_f / 1e6, "main_tflops": flops_f / t_m / 1e6,
                   "torch_fwd_tflops": flops_f / t_t / 1e6, "bwd_tflops": flops_b / t_b / 1e6}
            rows.append(row)
            print(f"{M:6d} {K:5d} {N:5d} {r:4d} | {t_f:12.1f} {row['fused_fwd_tflops']:7.0f} | {t_m:8.1f} {row['main_tflops']:7.0f} | "
                  f"{t_t:9.1f} {row['torch_fwd_tflops']:7.0f} | {t_b:8.1f} {row['bwd_tflops']:7.0f} | {t_tb:9.1f}", flush=True)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
