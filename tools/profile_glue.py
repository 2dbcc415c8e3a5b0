"""Which host-side op launches each stock torch kernel of a micro-step (torch.profiler, eager, one micro-step).

Groups the CUDA kernels of one eager micro-step by (kernel, chain of parent ops, input shapes): finds the strided adds / copies /
layout conversions around this library's launches.  A diagnosis tool: profiler times are not bench values.

    python tools/profile_glue.py --config turbo64 [--match elementwise_kernel] [--top 40]
"""
import argparse
import collections
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="turbo64")
    ap.add_argument("--match", default="")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--stack", action="store_true")
    a = ap.parse_args()
    args = bench.parse_args(["--config", a.config, "--no-eager-baseline", "--no-turbo64"])
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    from pairwise_sample_optimization_b200 import _lib
    arm = bench.B200Arm(args, a.config, dev, 0, 1, _lib.lib())
    for _ in range(2):
        arm.micro(arm.d, overlap=False)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=a.stack) as prof:
        arm.micro(arm.d, overlap=False)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    for ev in prof.events():
        ks = getattr(ev, "kernels", None)
        if not ks:
            continue
        chain, p = [], ev
        while p is not None and len(chain) < 5:
            chain.append(re.sub(r"autograd::engine::evaluate_function: ", "bw:", p.name)[:48])
            p = p.cpu_parent
        shapes = str(ev.input_shapes)[:110] if ev.input_shapes else ""
        stack = ""
        if a.stack and ev.stack:
            own = [s for s in ev.stack if "/root/repo" in s or "fixtures/" in s or "pairwise_sample" in s]
            stack = " | ".join(s.split("/")[-1][:60] for s in own[:3])
        for k in ks:
            name = re.sub(r"\s+", "", k.name)
            name = re.sub(r"^void", "", name)[:90]
            t = k.duration if hasattr(k, "duration") else k.device_time
            total += t
            if a.match and a.match not in name:
                continue
            key = (name, " < ".join(chain), shapes, stack)
            agg[key][0] += 1
            agg[key][1] += t
    print(f"total device time of the micro-step's kernels: {total / 1e3:.2f} ms")
    for (name, chain, shapes, stack), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: a.top]:
        print(f"{t / 1e3:8.3f} ms {n:5d} x {t / n:7.1f} us  {name}\n      via {chain}\n      shapes {shapes}" + (f"\n      at {stack}" if stack else ""))


if __name__ == "__main__":
    main()
