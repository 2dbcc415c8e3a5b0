"""Time the exchange of the flat LoRA gradient (the only collective of the training step) on N ranks, both ways:
NCCL all-reduce + separate norm pass vs this repo's fused multimem kernel (exchange + norm, bracketed by its two barriers).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/allreduce_probe.py
"""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pairwise_sample_optimization_b200 import _lib, lora  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world = dist.get_world_size()
for n in (1451520 * 8, 1451520 * 64):
    flat = torch.randn(n, device=dev)
    ex = lora.SymmetricGradExchange()
    sym = ex.allocate(n, dev)
    sym.copy_(flat)
    for name, fn in (("nccl all_reduce(AVG) + vector_norm", lambda: (dist.all_reduce(flat, op=dist.ReduceOp.AVG), torch.linalg.vector_norm(flat))),
                     ("fused multimem exchange + norm (incl. 2 barriers)", ex.exchange)):
        times = []
        for i in range(12):
            torch.cuda.synchronize(); dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        t = torch.tensor([sorted(times[2:])[len(times[2:]) // 2]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ms = t.item()
            print(f"world {world} n={n} ({n * 4 / 1e6:.0f} MB) {name}: median {ms:.3f} ms (first call {times[0]:.2f}), "
                  f"algbw {n * 4 / ms / 1e6:.0f} GB/s", flush=True)
    del ex, sym
dist.destroy_process_group()
