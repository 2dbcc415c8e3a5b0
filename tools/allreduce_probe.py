"""Time the flat-bucket all-reduce (the only collective of the training step) on N ranks: torchrun ... tools/allreduce_probe.py"""
import os, torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for n in (1451520 * 8, 1451520 * 64):
    flat = torch.randn(n, device="cuda")
    for i in range(12):
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dist.all_reduce(flat, op=dist.ReduceOp.AVG); b.record(); torch.cuda.synchronize()
        if rank == 0 and i in (0, 1, 2, 5, 11):
            ms = a.elapsed_time(b)
            print(f"n={n} ({n * 4 / 1e6:.0f} MB) call {i}: {ms:.2f} ms, algbw {n * 4 / ms / 1e6:.0f} GB/s", flush=True)
dist.destroy_process_group()
