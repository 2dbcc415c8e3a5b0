"""Worker of tests/test_gpu_multi_exchange.py (one process per GPU, launched by torch.distributed.run): the fused
NVLink-multicast gradient exchange + optimizer boundary against NCCL all-reduce + the same optimizer."""
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pairwise_sample_optimization_b200 import lora  # noqa: E402


def make_model(dev, r):
    torch.manual_seed(5)  # same weights on every rank
    net = nn.ModuleDict({"to_q": nn.Linear(64, 96, bias=False), "to_k": nn.Linear(64, 96, bias=False),
                         "to_v": nn.Linear(128, 64, bias=False)}).to(dev).bfloat16().requires_grad_(False)
    wrapped = lora.add_adapter(net, lora.LoraConfig(r=r, lora_alpha=r, target_modules=["to_q", "to_k", "to_v"]))
    for m in wrapped:
        nn.init.normal_(m.lora_B["default"].weight, std=0.02)
    return net


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    for r in (8, 12):  # 12: operand copies with a padded pitch
        fused = lora.FusedLoRAOptimizer(make_model(dev, r), lr=1e-3, max_grad_norm=0.5, exchange=lora.SymmetricGradExchange())
        plain = lora.FusedLoRAOptimizer(make_model(dev, r), lr=1e-3, max_grad_norm=0.5)
        n = fused.bucket.flat.numel()
        assert plain.bucket.flat.numel() == n
        for boundary in range(3):
            g = torch.Generator(device=dev).manual_seed(100 * boundary + rank)
            grad = torch.randn(n, device=dev, generator=g) * (1.0 + rank)
            fused.bucket.flat.copy_(grad)
            plain.bucket.flat.copy_(grad)
            want = grad.clone()
            dist.all_reduce(want, op=dist.ReduceOp.AVG)
            fused.all_reduce()
            plain.all_reduce()
            got = fused.bucket.flat.clone()
            err = (got - want).abs().max().item() / want.abs().max().item()
            assert err <= 1e-6, f"rank {rank}: fused exchange differs from the NCCL mean by {err}"
            # every rank holds bitwise-identical gradients
            gathered = [torch.empty_like(got) for _ in range(world)]
            dist.all_gather(gathered, got)
            assert all(torch.equal(gathered[0], t) for t in gathered), "ranks disagree on the exchanged gradient"
            nf, npl = fused.step().clone(), plain.step().clone()
            want_norm = torch.linalg.vector_norm(want.double()).item()
            assert abs(nf.item() - want_norm) <= 1e-5 * want_norm, (nf.item(), want_norm)
            assert abs(npl.item() - want_norm) <= 1e-5 * want_norm
            perr = (fused.flat_param - plain.flat_param).abs().max().item()
            assert perr <= 1e-6, f"parameters after the boundary differ by {perr}"
            assert float(fused.bucket.flat.abs().max()) == 0.0  # zero_grad
            torch.cuda.synchronize()
        if rank == 0:
            print(f"exchange ok: world {world}, rank-{r} adapters, {n} gradient elements", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
