#!/usr/bin/env python
"""A few launches of the streaming kernels at the top of the config-5 sweep, for ncu:

    python tools/once_kernels.py sampler|loss|preprocess [reps]
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 1 -c 1 -o gpurun_out/<name> python tools/once_kernels.py <which>
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pairwise_sample_optimization_b200 as pso  # noqa: E402
from pairwise_sample_optimization_b200 import _lib, runtime, step_ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "sampler"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(11)
B = 256
mk = lambda: torch.randn(B, 4, 128, 128, device=dev, generator=g).bfloat16()
if which == "sampler":
    pred, x, noise = mk(), mk(), mk()
    ts = torch.tensor([999, 749, 499], device=dev)[torch.randint(0, 3, (B,), device=dev)]
    sched = runtime.turbo_schedule(bench.make_scheduler("turbo", dev), dev, _lib.ts_dtype_code(ts))
    for _ in range(reps):
        step_ops.step_forward(sched, pred, x, ts, noise=noise, want_scaled_next=True)
elif which == "loss":
    x0, x1, r0, r1, n0, n1 = mk(), mk(), mk(), mk(), mk(), mk()
    p0 = (r0.float() + 0.02 * torch.randn_like(r0, dtype=torch.float32)).bfloat16()
    p1 = (r1.float() + 0.02 * torch.randn_like(r1, dtype=torch.float32)).bfloat16()
    ts = torch.tensor([999, 749, 499], device=dev)[torch.randint(0, 3, (B,), device=dev)]
    h = torch.tensor([[-1.0, 1.0]], device=dev).repeat(B, 1)
    sched = bench.make_scheduler("dmd", dev)
    for _ in range(reps):
        with torch.no_grad():
            pso.pso_pair_loss(p0, p1, r0, r1, x0, x1, n0, n1, ts, ts, h, scheduler=sched, kind="dmd", step_ratio=250)
else:
    imgs = (torch.rand(16, 3, 512, 512, device=dev, generator=g) * 2 - 1).half()
    for _ in range(reps):
        pso.clip_image_preprocess(imgs)
torch.cuda.synchronize()
print("ok", which)
