#!/usr/bin/env python
"""Kernel-time breakdown of one eager micro-step of bench.py's workload (torch.profiler / CUPTI)."""
import os, sys, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import pairwise_sample_optimization_b200 as pso
from fixtures import micro_step, sdxl_unet
from pairwise_sample_optimization_b200 import lora
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda", 0)
torch.manual_seed(1234)
cfg = sdxl_unet.sdxl_config()
with torch.device(dev):
    unet = sdxl_unet.UNet2DConditionModel(cfg)
unet = unet.to(torch.bfloat16).requires_grad_(False)
wrapped = lora.add_adapter(unet, lora.LoraConfig(r=8, lora_alpha=8))
for m in wrapped:
    torch.nn.init.normal_(m.lora_B["default"].weight, std=0.01)
unet.set_attn_processor(lora.PSOAttnProcessor2_0()); unet.train(); unet.enable_gradient_checkpointing()
bucket = lora.LoRAGradBucket(lora.lora_parameters(unet))
sched = bench.turbo_scheduler()
host = micro_step.batched_view(micro_step.synth_batch(4, 64, 2048, 1280, 100, sched.sigmas, dtype=torch.bfloat16))
d = {k: v.to(dev) for k, v in host.items()}
step = lambda: micro_step.product_micro_step_batched(pso, lora, unet, d, sched, loss_scale=1 / 6)
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
tot = collections.Counter(); cnt = collections.Counter()
for e in prof.events():
    if e.device_type.name == "CUDA":
        name = e.name.split("<")[0][:70]
        tot[name] += e.device_time if hasattr(e, "device_time") else e.cuda_time
        cnt[name] += 1
total = sum(tot.values())
print(f"total kernel time {total / 1e3:.1f} ms over {sum(cnt.values())} launches")
for name, t in tot.most_common(25):
    print(f"{t / 1e3:8.2f} ms {100 * t / total:5.1f}%  x{cnt[name]:5d}  avg {t / cnt[name]:7.1f} us  {name}")
