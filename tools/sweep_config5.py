#!/usr/bin/env python
"""BASELINE config 5: PSO loss/grad + sampler + LoRA GEMM microbench sweep (batch 1-256 pairs, latents 64^2-128^2, rank 4-128,
bf16 and fp32 storage, Turbo and DMD2 schedules, LoRA forward AND backward), on every GPU of the launch:

    python tools/sweep_config5.py [--quick] [--md profiles/r02_config5_sweep_1gpu.md] [--json out.json]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/sweep_config5.py ...

Every rank runs the whole sweep on its own GPU at the same time (the kernels need no communication: pairs shard over ranks,
SURVEY.md section 8e); rank 0 gathers the per-row times and reports the SLOWEST rank (max time = min throughput) and the spread.
Device times: CUDA events around CUDA-graph replays (bench.graph_timed), inputs larger than L2 at the top of each sweep.
"""
import argparse
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pairwise_sample_optimization_b200 as pso  # noqa: E402
from pairwise_sample_optimization_b200 import _lib, lora, runtime, step_ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--json", default="")
ap.add_argument("--md", default="")
ap.add_argument("--quick", action="store_true", help="fewer points (multi-GPU runs)")
args = ap.parse_args()
rank, world, local = bench.dist_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
peaks = bench.measured_peaks()
g = torch.Generator(device=dev).manual_seed(rank)
rows = []  # (section, key tuple, us, algorithmic bytes or flops, extra)

PAIRS = (1, 4, 16, 64, 256) if args.quick else (1, 2, 4, 8, 16, 32, 64, 128, 256)
RANKS = (4, 64, 128) if args.quick else (4, 8, 16, 32, 64, 128)
DT = {"bf16": torch.bfloat16, "fp32": torch.float32}


def timed(fn, reps=15, per_graph=4):
    return bench.graph_timed(fn, reps, per_graph=per_graph) * 1e3  # us


# ---- fused loss + grad kernel: both schedules, both storage types
for kind in ("dmd", "turbo"):
    sched = bench.make_scheduler(kind, dev)
    for dname, dt in DT.items():
        for hw in (64, 128):
            for B in PAIRS + ((1024,) if hw == 64 and dname == "bf16" and not args.quick else ()):
                scale = sched.sigmas[0].item() if kind == "turbo" else 1.0
                mk = lambda s=1.0: (torch.randn(B, 4, hw, hw, device=dev, generator=g) * s).to(dt)
                x0, x1, n0, n1 = mk(scale), mk(scale), mk(scale), mk(scale)
                r0, r1 = mk(), mk()
                p0 = (r0.float() + 0.02 * torch.randn(r0.shape, device=dev, generator=g)).to(dt)
                p1 = (r1.float() + 0.02 * torch.randn(r1.shape, device=dev, generator=g)).to(dt)
                ts = torch.tensor([999, 749, 499], device=dev)[torch.randint(0, 3, (B,), device=dev, generator=g)]
                h = torch.tensor([[-1.0, 1.0]], device=dev).repeat(B, 1)
                kw = dict(scheduler=sched, kind=kind, step_ratio=250 if kind == "dmd" else None)

                def call():
                    with torch.no_grad():
                        pso.pso_pair_loss(p0, p1, r0, r1, x0, x1, n0, n1, ts, ts, h, **kw)
                us = timed(call)
                rows.append(("loss", (kind, dname, hw, B), us, 10 * 4 * hw * hw * dt.itemsize * B, None))

# ---- sampler update kernel (sampling mode + fused next-input scaling / DMD2 shared noise)
for kind in ("turbo", "dmd"):
    sched_ns = bench.make_scheduler(kind, dev)
    for dname, dt in DT.items():
        for hw in (64, 128):
            for B in PAIRS:
                mk = lambda n=B: torch.randn(n, 4, hw, hw, device=dev, generator=g).to(dt)
                pred, x = mk(), mk()
                ts = torch.full((B,), 749, device=dev)
                if kind == "turbo":
                    noise = mk()
                    sd = runtime.turbo_schedule(sched_ns, dev, _lib.ts_dtype_code(ts))
                    fn = lambda: step_ops.step_forward(sd, pred, x, ts, noise=noise, want_scaled_next=True)
                    nbytes = 5 * 4 * hw * hw * dt.itemsize * B          # 3 N read + 2 N written per sample
                else:
                    noise = mk(1)                                        # one draw shared by the batch (DS:123-124)
                    sd = runtime.dmd_schedule(sched_ns, dev, _lib.ts_dtype_code(ts))
                    tsp = ts - 250
                    fn = lambda: step_ops.step_forward(sd, pred, x, ts, tsp, noise=noise)
                    nbytes = (3 * B + 1) * 4 * hw * hw * dt.itemsize     # 2 N read + N written per sample + the shared noise
                rows.append(("sampler", (kind, dname, hw, B), timed(fn), nbytes, None))

# ---- LoRA-wrapped projection, forward and forward+backward, through the public module; vs the stock lowering
rn = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).bfloat16()
SHAPES = [(2048, 1280, 1280), (8192, 640, 640), (8192, 1280, 1280), (32768, 640, 640), (616, 2048, 1280)]
if args.quick:
    SHAPES = [(2048, 1280, 1280), (8192, 1280, 1280), (616, 2048, 1280)]
F = torch.nn.functional
for (M, K, N) in SHAPES:
    for r in RANKS:
        lay = lora.LoRALinear(torch.nn.Linear(K, N, bias=True, device=dev, dtype=torch.bfloat16), r, r)
        with torch.no_grad():
            lay.lora_B["default"].weight.normal_(std=0.02)
        w, b = lay.base_layer.weight, lay.base_layer.bias
        x, dy = rn(M, K), rn(M, N)
        xg, xa = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        Ap = lay.lora_A["default"].weight.detach().bfloat16().clone().requires_grad_(True)
        Bp = lay.lora_B["default"].weight.detach().bfloat16().clone().requires_grad_(True)

        def fwd():
            with torch.no_grad():
                return lay(x)

        def fwd_bwd():
            lay(xg).backward(dy)

        def stock_fwd():
            with torch.no_grad():
                return F.linear(x, w, b) + F.linear(F.linear(x, Ap), Bp) * 1.0

        def stock_fwd_bwd():
            (F.linear(xa, w, b) + F.linear(F.linear(xa, Ap), Bp) * 1.0).backward(dy)
        fl_f = 2.0 * M * K * N + 2.0 * M * r * (K + N)
        fl_fb = 3 * fl_f  # forward + dX (+ U) + dA, dB
        rows.append(("lora", (M, K, N, r), timed(fwd, 10, 2), fl_f, {"fwd_bwd_us": timed(fwd_bwd, 10, 2), "fl_fb": fl_fb,
                                                                     "stock_fwd_us": timed(stock_fwd, 10, 2),
                                                                     "stock_fwd_bwd_us": timed(stock_fwd_bwd, 10, 2)}))
        del lay, xg, xa

# ---- gather: slowest rank per row
flat = torch.tensor([r_[2] for r_ in rows] + [v for r_ in rows if r_[4] for v in (r_[4]["fwd_bwd_us"], r_[4]["stock_fwd_us"], r_[4]["stock_fwd_bwd_us"])],
                    device=dev, dtype=torch.float64)
lo, hi = flat.clone(), flat.clone()
if world > 1:
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
if rank == 0:
    hi, lo = hi.tolist(), lo.tolist()
    n = len(rows)
    out = []
    ex = n
    lines = [f"# config-5 sweep on {world} x B200 (every rank runs the whole sweep concurrently; slowest rank reported, spread = max/min over ranks)\n",
             f"Peaks: HBM {peaks['hbm']:.0f} GB/s, bf16 {peaks['tf_burst']:.0f} TF/s burst ({peaks['source']}).  `python tools/sweep_config5.py"
             f"{' --quick' if args.quick else ''}`\n"]
    for sec, title, hdr in (("loss", "fused PSO loss + grad kernel (algorithmic bytes 10 N sizeof per pair)", "| schedule | storage | latent | pairs | us | GB/s | of HBM | spread |"),
                            ("sampler", "sampler update kernel (Turbo: 3 N read + 2 N written; DMD2: 2 N + N/B read + N written)", "| schedule | storage | latent | samples | us | GB/s | of HBM | spread |")):
        lines += [f"\n## {title}\n", hdr, "|" + "---|" * 8]
        for i, r_ in enumerate(rows):
            if r_[0] != sec:
                continue
            kind, dname, hw, B = r_[1]
            gbs = r_[3] / hi[i] / 1e3
            out.append({"section": sec, "schedule": kind, "storage": dname, "hw": hw, "pairs": B, "us": hi[i], "gbs": gbs, "spread": hi[i] / lo[i]})
            lines.append(f"| {kind} | {dname} | 4x{hw}x{hw} | {B} | {hi[i]:.1f} | {gbs:.0f} | {100 * gbs / peaks['hbm']:.1f} % | {hi[i] / lo[i]:.2f} |")
    lines += ["\n## LoRA-wrapped projection through the public module (bf16), this library vs the stock lowering (3 cuBLAS GEMMs + scale + add, autograd)\n",
              "| M | K | N | r | fwd us | TF/s | fwd+bwd us | TF/s | stock fwd us | stock fwd+bwd us | speed-up fwd+bwd | spread |", "|" + "---|" * 12]
    for i, r_ in enumerate(rows):
        if r_[0] != "lora":
            continue
        M, K, N, r = r_[1]
        fb, sf, sfb = hi[ex], hi[ex + 1], hi[ex + 2]
        ex += 3
        out.append({"section": "lora", "M": M, "K": K, "N": N, "r": r, "fwd_us": hi[i], "fwd_bwd_us": fb, "stock_fwd_us": sf, "stock_fwd_bwd_us": sfb,
                    "fwd_tflops": r_[3] / hi[i] / 1e6, "fwd_bwd_tflops": r_[4]["fl_fb"] / fb / 1e6})
        lines.append(f"| {M} | {K} | {N} | {r} | {hi[i]:.1f} | {r_[3] / hi[i] / 1e6:.0f} | {fb:.1f} | {r_[4]['fl_fb'] / fb / 1e6:.0f} | {sf:.1f} | {sfb:.1f} | "
                     f"{sfb / fb:.2f} | {hi[i] / lo[i]:.2f} |")
    text = "\n".join(lines) + "\n"
    print(text)
    if args.md:
        with open(args.md, "w") as f:
            f.write(text)
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
