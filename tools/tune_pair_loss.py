#!/usr/bin/env python
"""Sweep launch geometries of the fused loss+grad kernel (and the sampler-update kernel) on one GPU.

    python tools/tune_pair_loss.py                 # sweep table (CUDA events, inputs larger than L2)
    python tools/tune_pair_loss.py --once 5        # 5 launches of the default geometry (for ncu)
"""
from __future__ import annotations

import argparse
import os
import statistics
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import pairwise_sample_optimization_b200 as pso  # noqa: E402
from pairwise_sample_optimization_b200 import _lib, runtime, step_ops  # noqa: E402


def timed(fn, reps, per_graph=5):
    """Median device time of one call of ``fn`` in microseconds.  The Python/ctypes cost of a call (tens of
    microseconds) exceeds these kernels' run time, so ``per_graph`` calls are captured in a CUDA graph and the
    graph is replayed between CUDA events: the host then runs ahead of the GPU and only device time is seen."""
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
    ap.add_argument("--pairs", type=int, default=256)
    ap.add_argument("--hw", type=int, default=128)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    dt = {"bf16": torch.bfloat16, "fp32": torch.float32, "fp16": torch.float16}[args.dtype]
    B, n = args.pairs, 4 * args.hw * args.hw
    g = torch.Generator(device=dev).manual_seed(0)
    mk = lambda: torch.randn(B, 4, args.hw, args.hw, device=dev, generator=g).to(dt)
    x0, x1, r0, r1, n0, n1 = mk(), mk(), mk(), mk(), mk(), mk()
    p0 = (r0.float() + 0.02 * torch.randn_like(r0, dtype=torch.float32)).to(dt)
    p1 = (r1.float() + 0.02 * torch.randn_like(r1, dtype=torch.float32)).to(dt)
    ts = torch.tensor([999, 749, 499], device=dev)[torch.randint(0, 3, (B,), device=dev)]
    h = torch.tensor([[-1.0, 1.0]], device=dev).repeat(B, 1)
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
    sched = types.SimpleNamespace(alphas_cumprod=torch.cumprod(1.0 - betas, dim=0).to(dev))

    def loss(tune=(0, 0)):
        with torch.no_grad():
            return pso.pso_pair_loss(p0, p1, r0, r1, x0, x1, n0, n1, ts, ts, h, scheduler=sched, kind="dmd",
                                     step_ratio=250, tune=tune)

    if args.once:
        for _ in range(args.once):
            loss()
        torch.cuda.synchronize()
        print("ok")
        return

    esz = torch.tensor([], dtype=dt).element_size()
    alg = 10 * n * esz * B
    # reference points: torch's copy kernel on the same amount of traffic
    a = torch.empty(alg // 2, dtype=torch.uint8, device=dev)
    b = torch.empty_like(a)
    us = timed(lambda: b.copy_(a), args.reps)
    print(f"torch copy_ of {alg // 2 / 1e6:.0f} MB (read+write {alg / 1e6:.0f} MB): {us:.1f} us = {alg / us / 1e3:.0f} GB/s")
    print(f"pairs={B} latent=4x{args.hw}x{args.hw} dtype={args.dtype} algorithmic bytes={alg / 1e6:.1f} MB")
    print("threads cluster      us    GB/s")
    res = []
    # threads == 0: tensor-memory kernel, 1: TMA-ring kernel, >= 32: general LDG kernel
    for threads in (0, 1, 256):
        for cluster in (1, 2, 4, 8):
            try:
                us = timed(lambda: loss((threads, cluster)), args.reps)
            except _lib.Psob200Error as e:
                print(f"{threads:7d} {cluster:7d}  unsupported ({str(e)[:60]})")
                continue
            res.append((us, threads, cluster))
            print(f"{threads:7d} {cluster:7d} {us:7.1f} {alg / us / 1e3:7.0f}")
    us = timed(lambda: loss((0, 0)), args.reps)
    print(f"   auto    auto {us:7.1f} {alg / us / 1e3:7.0f}")
    best = min(res)
    print(f"best: threads={best[1]} cluster={best[2]} {best[0]:.1f} us {alg / best[0] / 1e3:.0f} GB/s")
    # sampler-update kernel (sampling mode): reads eps + x (+ shared noise), writes x'
    tsd = runtime.timesteps_on(ts, dev)
    sd = runtime.dmd_schedule(sched, dev, _lib.ts_dtype_code(tsd))
    noise = torch.randn(1, 4, args.hw, args.hw, device=dev).to(dt)
    salg = 3 * n * esz * B
    for tune in ((0, 0), (256, 1), (256, 2), (256, 4), (256, 8), (512, 8), (128, 8)):
        us = timed(lambda: step_ops.step_forward(sd, r0, x0, tsd, tsd - 250, noise=noise, tune=tune), args.reps)
        print(f"sampler step tune={tune}: {us:.1f} us = {salg / us / 1e3:.0f} GB/s (algorithmic {salg / 1e6:.0f} MB)")
    us = timed(lambda: step_ops.step_forward(sd, r0, x0, tsd, tsd - 250, prev_sample=n0), args.reps)
    print(f"scoring step: {us:.1f} us = {salg / us / 1e3:.0f} GB/s")


if __name__ == "__main__":
    main()
